"""Turns the B200_LOOP_TRACE lines a peer-memory BFS run prints (profiles/trace_p2p.py, stderr) into the readable
timeline committed under profiles/.    python profiles/format_trace.py gpurun_out/x/trace.err > profiles/r02_trace_x.txt"""
import re
import sys

NAMES = {20: "small-level kernel: entry", 21: "small: level start (expand the local frontier)", 22: "small: sends out, flag to peers",
         23: "small: peers' flags in, absorb inboxes", 24: "small: stats row posted", 25: "small: stats in, decision",
         2: "quad scan (entry; idle unless a big push level)", 3: "claim-only quad advance (entry)", 5: "bitmap absorb (entry)",
         13: "bitmap absorb: flag barrier passed", 9: "stats + decide (entry)", 12: "stats + decide: flag barrier passed",
         30: "pull-levels kernel: entry", 31: "pull: gather + OR of the frontier slices starts", 32: "pull: early-exit pull starts",
         33: "pull: stats row posted", 34: "pull: stats in, decision",
         26: "small: CTA 0 done expanding (grid barrier follows)", 27: "small: grid barrier passed",
         28: "small: CTA 0 done absorbing (grid barrier follows)", 29: "small: grid barrier passed",
         35: "pull: CTA 0 out of chunks (grid barrier follows)", 36: "pull: grid barrier passed",
         11: "bitmap -> list (entry)", 40: "near-far SSSP: pending-minimum pass", 41: "near-far SSSP: take pass"}
SINGLE = {2: "quad scan (entry; returns at once when the previous advance created this level's scan)", 3: "quad advance (entry)",
          9: "decide (one thread: counters -> next level's state)", 11: "bitmap -> list (entry; returns at once unless pull -> push)"}   # ids whose meaning differs in the 1-GPU loop (level_loop.cu)
if len(sys.argv) > 2 and sys.argv[2] == "single":
    NAMES.update(SINGLE)
for line in open(sys.argv[1]):
    m = re.match(r"B200_LOOP_TRACE rank(\d+) (\d+) entries.*?: (.*)", line)
    if not m:
        continue
    ent = [(float(t), int(k)) for t, k in (x.split(":") for x in m.group(3).split())]
    print(f"# rank {m.group(1)}: {m.group(2)} entries; microseconds since the first kernel of the traversal (%globaltimer)")
    prev_close = 0.0
    for i, (t, k) in enumerate(ent):
        nxt = ent[i + 1][0] if i + 1 < len(ent) else t
        if k >= 64:
            print(f"{t:9.1f}  ---- level {k - 64} closed: {t - prev_close:7.1f} us ----")
            prev_close = t
        else:
            print(f"{t:9.1f}  (+{nxt - t:6.1f})  {k:2d}  {NAMES.get(k, '?')}")
    print()
