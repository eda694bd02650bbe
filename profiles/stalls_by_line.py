"""Attributes an ncu report's warp-stall samples to CUDA source lines.
ncu's CSV source page lists SASS only; nvdisasm --print-line-info on the same cubin gives the line of
every SASS instruction, and the two listings are in the same order.
   python profiles/stalls_by_line.py <rep.ncu-rep> <kernel-id (1-based)> <lib.so> <mangled-substring> [top]"""
import csv, io, re, subprocess, sys, os, tempfile, collections

rep, kid, lib, pat = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
inst, seen = [], set()
for r in rows[hi + 1:]:          # the CSV page lists every instruction twice
    if len(r) >= len(hdr) and r[0] not in seen:
        seen.add(r[0])
        inst.append(r)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    # split per function
    cur, name, funcs = [], None, {}
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            name = m.group(1); funcs[name] = []
        elif name:
            funcs[name].append(l)
    for n, body in funcs.items():
        if pat in n:
            cand = []
            line = None
            for l in body:
                m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                if m:
                    line = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                    cand.append(line)
            if abs(len(cand) - len(inst)) <= 1:
                lines = cand
    if lines:
        break
if not lines:
    sys.exit(f"no function matching {pat} with {len(inst)} instructions")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for r, ln in zip(inst, lines):
    s = int(r[col["# Samples"]] or 0)
    tot += s
    a = agg[ln]
    a["samples"] += s
    a["inst"] += int(r[col["Instructions Executed"]] or 0)
    a["l2sect"] += int(r[col["L2 Theoretical Sectors Global"]] or 0)
    a["shwave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
    for st in stalls:
        a[st] += int(r[col[st]] or 0)
print(f"# {rep} kernel {kid}: {tot} samples, {len(inst)} SASS instructions, {sum(a['inst'] for a in agg.values())} warp-instructions")
srcs = {}
for ln, a in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:top]:
    if ln is None:
        text = "?"
    else:
        if ln[0] not in srcs:
            for root in ("include/b200", "mini_b200/csrc", "include"):
                p = os.path.join(root, ln[0])
                if os.path.exists(p):
                    srcs[ln[0]] = open(p).read().splitlines()
        text = srcs.get(ln[0], [""] * 100000)[ln[1] - 1].strip()[:70] if ln[0] in srcs else ""
    st = sorted(((a[s], s) for s in stalls), reverse=True)[:2]
    print(f"{100.0 * a['samples'] / tot:5.1f}%  inst={a['inst']:>9d} l2sect={a['l2sect']:>9d} smwave={a['shwave']:>9d}  {ln}  {text}   [{st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}]")
