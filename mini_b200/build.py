"""Builds mini_b200/libb200_frontier.so (sm_100a only) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200_frontier.so")
SOURCES = ["context.cu", "graph_build.cu", "engine.cu", "level_loop.cu", "multi_gpu.cu", "p2p_bfs.cu", "ingest.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _deps():
    out = []
    for d in (CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "b200")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(d, f))
    return out


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = "") -> str:
    """defines/out: tuning builds (e.g. defines=["B200_QUAD_NT=512"], out="libvariant.so"); the
    default library is what the package loads unless B200_FRONTIER_LIB points elsewhere."""
    nvcc = os.environ.get("NVCC", "nvcc")
    global OBJ, LIB
    obj_dir, lib = OBJ, LIB
    if out:
        lib = os.path.join(HERE, out)
        obj_dir = os.path.join(HERE, "build", os.path.splitext(out)[0])
    os.makedirs(obj_dir, exist_ok=True)
    deps = _deps()
    extra = (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if force or _stale(obj, deps):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", dest="verbose", action="store_true")
    ap.add_argument("--define", action="append", default=[])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, defines=a.define, out=a.out))
