"""Turns an .ncu-rep into the text summary committed under profiles/ (run in the build container:
ncu reads reports without a GPU).   python profiles/summarize_ncu.py gpurun_out/x.ncu-rep [kernel-index] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kid = int(sys.argv[2]) if len(sys.argv) > 2 else None
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
print(f"# ncu summary of {rep} ({len(data)} kernel launches captured; --set full --clock-control none)")
ki = hdr.index("Kernel Name")
for n, r in enumerate(data):
    if kid is not None and n != kid:
        continue
    print(f"\n## launch {n}: {r[ki][:110]}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]:>18s} {units[i]}")
# stall reasons of one launch (default: the longest)
if kid is None:
    ti = hdr.index("gpu__time_duration.sum")
    kid = max(range(len(data)), key=lambda n: float(data[n][ti].replace(",", "")))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid + 1}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    hdr = rows[hi[0]]
    col = {n: i for i, n in enumerate(hdr)}
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: 0 for s in stalls}
    seen, per = set(), []
    for r in rows[hi[0] + 1:]:
        if len(r) < len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            samp = int(r[col["# Samples"]])
        except ValueError:
            continue
        d = {s: int(r[col[s]] or 0) for s in stalls}
        for s in stalls:
            tot[s] += d[s]
        per.append((samp, r[col["Source"]], r[col["Instructions Executed"]], d))
    T = sum(p[0] for p in per) or 1
    print(f"\n## warp stall sampling, launch {kid} ({T} samples over {len(per)} SASS instructions)")
    for s, v in sorted(tot.items(), key=lambda x: -x[1])[:9]:
        print(f"{s:28s} {100.0 * v / T:5.1f} %")
    print("\ntop SASS instructions by samples:")
    for samp, s, ie, d in sorted(per, key=lambda x: -x[0])[:14]:
        top = max(d.items(), key=lambda x: x[1])
        print(f"{100.0 * samp / T:5.1f} %  exec={ie:>10s}  {s[:64]:64s} {top[0]}")
