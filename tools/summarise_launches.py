"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and one step's sequence."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("phyx::", "").replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(r[ui], 1.0)
    seq.append((name, v))
names = [s[0] for s in seq]
# last complete step: from the last k_integrate_velocity that is followed by a k_integrate_position
ends = [i for i, n in enumerate(names) if n == "k_integrate_position"]
starts = [i for i, n in enumerate(names) if n == "k_integrate_velocity"]
end = ends[-1]
start = max(s for s in starts if s < end)
step = seq[start:end + 1]
agg = collections.OrderedDict()
for n, v in step:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(v for _, v in step)
print(f"{len(seq)} launches captured; last complete step: {len(step)} launches, {total:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} n={n:3d} total={t * 1e3:9.1f} us  share={100 * t / total:5.1f}%")
