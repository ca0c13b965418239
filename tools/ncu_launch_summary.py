"""per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (python bench.py --steps 2 --warmup 1)."""
import csv, re, sys
from collections import defaultdict
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:60]
    tot[k][0] += 1
    tot[k][1] += v
total = sum(v[1] for v in tot.values())
print("ncu launch list of `python bench.py --steps 2 --warmup 1 --streams 0` (cold-cache, serialised per-launch times: compare SHARES)")
print("%d launches, %.1f us total" % (sum(v[0] for v in tot.values()), total))
ours = sum(v[1] for k, v in tot.items() if k.startswith("k_"))
print("kernels of libinsmos_b200.so (k_*): %.1f %% of the time, %d launches" % (100 * ours / total, sum(v[0] for k, v in tot.items() if k.startswith("k_"))))
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-62s %5d %10.1f us %5.1f %%" % (k, n, us, 100 * us / total))
