"""Per-kernel average duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/launch_times.py gpurun_out/launches.csv"""
import collections
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.reader(io.StringIO("".join(lines))))
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t = collections.defaultdict(list)
for r in rows[1:]:
    sc = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
    t[r[ik].split("(")[0][:70]].append(float(r[iv].replace(",", "")) * sc)
tot = sum(sum(v) for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):4d} avg {sum(v) / len(v):9.2f} us  share {100 * sum(v) / tot:5.1f} %")
