"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
usage: python tools/launch_summary.py gpurun_out/r2_launches_rcca.csv [skip_first_n_of_our_kernels]"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
names = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if len(r) > vi]
# the bench runs warm-up fits first: keep the launches after the LAST occurrence of the first kernel of a step
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    # the list ends with the timed step; it starts at the first kernel of the LAST group of marker launches
    idx = max(i for i, (n, _) in enumerate(names) if marker in n)
    while idx > 0 and marker in names[idx - 1][0]:
        idx -= 1
    names = names[idx:]
agg = collections.OrderedDict()
for n, t in names:
    k = n.split("(")[0].replace("void ", "").replace("ccab::<unnamed>::", "").replace("ccab::", "")[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
print(f"# launches: {sum(v[0] for v in agg.values())}, summed kernel time {tot/1e3:.3f} ms (cold-cache, serialised: compare SHARES)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={c:4d} sum_us={t:9.1f} avg_us={t/c:8.2f} share={100*t/tot:5.1f}%")
