"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py FILE [steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
        k = d["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
if steps is None:   # the optimiser runs once per step
    steps = max(1, sum(v[0] for k, v in agg.items() if "adam_step" in k or "sgd_step" in k or "momentum_step" in k))
tot = sum(v[1] for v in agg.values())
print("# %d launches, %.1f us total, %.0f steps -> %.2f ms / step (serialised, cold caches)" % (sum(v[0] for v in agg.values()), tot, steps, tot / steps / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%10.1f us/step %6.1f launches/step %7.1f us avg %5.1f%%  %s" % (v[1] / steps, v[0] / steps, v[1] / v[0], 100 * v[1] / tot, k[:80]))
