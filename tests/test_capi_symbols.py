"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/b200_frontier.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mini_b200 import build as b
    b.build()
    import mini_b200
    return mini_b200.load_library()


def _declared():
    names = []
    for hdr in ("include/b200_frontier.h", "include/b200/workspace.h"):
        txt = open(os.path.join(ROOT, hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 28
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the header's functions
    bound = set(lib._b200_signatures)
    assert bound == set(names), (bound ^ set(names))


def test_abi_version_and_status_strings(lib):
    assert lib.b200_abi_version() == 2
    assert lib.b200_status_string(0) == b"ok"
    assert b"capacity" in lib.b200_status_string(3)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.b200_ctx_create(C.byref(h), 0, None)
    assert rc != 0 and not h.value           # fails loudly, never computes on the CPU
    import mini_b200
    with pytest.raises(RuntimeError):
        mini_b200.Context(0)


def test_product_never_imports_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "mini_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(d, f)).read(), f
    for d, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            txt = open(os.path.join(d, f)).read()
            assert "#include \"oracle" not in txt and "liboracle" not in txt, f
