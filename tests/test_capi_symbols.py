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
    assert lib.b200_abi_version() == 3
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


def test_header_is_plain_c_and_ctypes_layouts_match(tmp_path):
    """include/b200_frontier.h must compile as C99 (it is the FFI boundary) and the ctypes mirrors in
    mini_b200/lib.py must have exactly the C layouts (sizes and a few offsets) -- ABI drift shows up here, on the CPU."""
    import subprocess
    from mini_b200 import lib as L
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stddef.h>
#include <stdio.h>
#include "b200_frontier.h"
#include "b200/workspace.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200_graph), offsetof(b200_graph, no_in_arc_bitmap),
           offsetof(b200_graph, first_in_neighbor), sizeof(b200_problem), sizeof(b200_level_stat), sizeof(b200_stats),
           offsetof(b200_stats, level_loop), offsetof(b200_stats, level), offsetof(b200_workspace, launches),
           (size_t)B200_ABI_VERSION);
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(L.CGraph), L.CGraph.no_in_arc_bitmap.offset, L.CGraph.first_in_neighbor.offset, C.sizeof(L.CProblem),
            C.sizeof(L.CLevelStat), C.sizeof(L.CStats), L.CStats.level_loop.offset, L.CStats.level.offset]
    assert got[:8] == want, (got, want)
    assert got[9] == 3
    # bench_multi.py reads b200_workspace::launches through a prefix mirror of the struct
    import re
    txt = open(os.path.join(ROOT, "bench_multi.py")).read()
    assert '("launches", C.c_int64)' in txt
    fields = re.findall(r'\("(\w+)", C\.(c_\w+)\)', txt[txt.index("class _WS"):txt.index("ws = C.cast")])
    sizes = {"c_void_p": 8, "c_int32": 4, "c_int64": 8, "c_uint": 4}
    off = 0
    for name, ty in fields:
        sz = sizes[ty]
        off = (off + sz - 1) // sz * sz
        if name == "launches":
            break
        off += sz
    assert off == got[8], (off, got[8])
