"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
   python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("b200::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    t = float(r[14].replace(",", "")) / 1e3
    a[0] += 1; a[1] += t; a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot:.1f} us of kernel time (cold-cache, serialised under ncu: compare SHARES)")
print(f"{'kernel':90s} {'n':>4s} {'total us':>10s} {'max us':>9s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:90]:90s} {a[0]:4d} {a[1]:10.1f} {a[2]:9.1f} {100 * a[1] / tot:6.1f}%")
