import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The virtual-rank multi-GPU tests run up to 8 ranks as threads on ONE GPU, each replaying its own CUDA graph whose
# kernels spin in cross-rank flag barriers.  With the default 8 hardware work queues two ranks' streams can share a
# queue, and a rank spinning at the head of a queue then blocks the peer it is waiting for.  One process per GPU (the
# deployment) never has this problem.  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_f32(case, key):
    """float32 array of a reference-GPU golden case: a JSON list, or base64 of float32 bytes (make_golden_gpu.py)."""
    import base64
    import numpy as np
    if key in case:
        return np.asarray(case[key], dtype=np.float32)
    return np.frombuffer(base64.b64decode(case[key + "_f32_b64"]), dtype=np.float32).copy()


def golden_custom_values(n):
    """The non-uniform initial ranks make_golden_gpu.py fed the reference GPU PR (values_formula in the golden files)."""
    import numpy as np
    v = np.arange(n, dtype=np.uint64)
    return (0.15 + ((v * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float64) / 1000.0).astype(np.float32)


@pytest.fixture(scope="session")
def golden():
    import json

    def load(name):
        with open(os.path.join(GOLDEN, name)) as f:
            return json.load(f)
    return load
