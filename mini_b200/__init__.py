"""mini_b200 -- B200-native frontier-traversal engine behind the mini-gunrock operator API.

The product is the C-ABI shared library ``libb200_frontier.so`` (include/b200_frontier.h,
hand-written sm_100a CUDA in mini_b200/csrc + include/b200) and the C++ header mirror of the
reference API in include/gunrock/.  This Python package is the ctypes binding the tests and
bench.py use; PyTorch only supplies device memory, streams and torch.distributed.
There is no CPU fallback: importing the compute API without the built library raises.
"""
from .lib import (  # noqa: F401
    B200Error, Context, Graph, Stats, load_library, library_path,
    BFS_PUSH, BFS_REF_ALPHA, BFS_BEAMER, ADV_IDEMPOTENT, ADV_NO_OUTPUT, ADV_RAW_OUTPUT,
    OP_PLUS, OP_MIN, OP_MAX, ADVANCE_QUAD, ADVANCE_LBS, ADVANCE_QUAD_RESCAN, LOOP_GRAPH, LOOP_HOST,
)

__all__ = ["B200Error", "Context", "Graph", "Stats", "load_library", "library_path"]
