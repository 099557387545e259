#!/usr/bin/env python3
"""tools/launch_summary.py <launches.csv> -- per-kernel share of ONE sweep from an `ncu --metrics gpu__time_duration.sum` launch list."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
starts = [i for i, n in enumerate(names) if "stats_kernel" in n or "rebuild_kernel" in n]
if len(starts) < 2:
    sys.exit("need at least two sweeps in the launch list")
lo, hi = starts[-2], starts[-1]          # the last complete sweep
agg = collections.OrderedDict()
if "permute" in names[lo - 1]:
    lo -= 1
    hi -= 1 if "permute" in names[hi - 1] else 0
for x in rows[lo:hi]:
    n = re.sub(r"\(.*", "", x["Kernel Name"]).replace("void ", "").replace("sbmf::", "")
    v = float(x["Metric Value"].replace(",", ""))
    v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# one sweep = launches {lo}..{hi - 1} of {sys.argv[1]}; serialised kernel time {tot / 1e3:.3f} ms in {hi - lo} launches")
print(f"{'us':>10} {'launches':>8} {'share':>7}  kernel")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]:10.1f} {a[0]:8d} {100 * a[1] / tot:6.1f}%  {n}")
