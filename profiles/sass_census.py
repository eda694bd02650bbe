"""Mnemonic census of the hot kernels' SASS (cuobjdump, no GPU needed): which memory / atomic / shuffle instructions
each kernel is made of -- the evidence that the index stream is read with 128-bit non-allocating loads, that probes
are plain L1-allocating loads, and that output is aggregated per warp (a handful of ATOMG against hundreds of SHFL).
   python profiles/sass_census.py > profiles/r01_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("quad_advance_kernel<BfsPushQ,COMPACT>", "engine.o", r"quad_advance_kernelINS_8BfsPushQELi1ELb0"),
        ("quad_advance_kernel<SsspRelaxQ,COMPACT>", "engine.o", r"quad_advance_kernelINS_10SsspRelaxQELi1ELb0"),
        ("quad_segreduce_kernel<float,PlusF32>", "engine.o", r"quad_segreduce_kernelIfNS_7PlusF32"),
        ("bfs_pull_kernel<256>", "engine.o", r"bfs_pull_kernelILi256"),
        ("bitmap_list_kernel", "engine.o", r"bitmap_list_kernelILi256"),
        ("scan_sizes_kernel<FrontierQuads>", "engine.o", r"scan_sizes_kernelILi256ELi4ENS_13FrontierQuads"),
        ("quad_advance_kernel<BfsPushPartQDyn,ROUTED> (multi-GPU)", "p2p_bfs.o", r"quad_advance_kernelINS_15BfsPushPartQDynELi3ELb1"),
        ("p2p_stats_decide_kernel (multi-GPU)", "p2p_bfs.o", r"p2p_stats_decide_kernel")]
KEEP = re.compile(r"^(LDG|STG|ATOMG|RED|ATOMS|LDS|STS|SHFL|BAR|MATCH|VOTE|LD\.|ST\.|ATOM\.|MEMBAR|CCTL)")
for title, obj, pat in WANT:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mini_b200", "build", obj)], capture_output=True, text=True).stdout
    fn, cnt, total = None, collections.Counter(), 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and re.search(pat, fn):
            m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", line)
            if m:
                total += 1
                if KEEP.match(m.group(2)):
                    cnt[m.group(2)] += 1
    print(f"## {title}: {total} SASS instructions")
    print("   " + ", ".join(f"{k} x{v}" for k, v in sorted(cnt.items(), key=lambda x: (-x[1], x[0]))))
    print()
