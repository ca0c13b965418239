"""summarise a bench.py --dump-launches file: per-family totals and every sparse-conv launch."""
import json, sys
from collections import defaultdict
rows = [json.loads(l) for l in open(sys.argv[1])]
other = [json.loads(l) for l in open(sys.argv[2])] if len(sys.argv) > 2 else None
d = defaultdict(float)
for r in rows: d[r['call']] += r['ms']
for k, v in sorted(d.items(), key=lambda kv: -kv[1]): print("%-32s %.3f" % (k, v))
print("sum %.3f ms" % sum(d.values()))
for i, r in enumerate(rows):
    if 'Cin' in r:
        o = ""
        if other and i < len(other) and 'Cin' in other[i]: o = "  (other %7.1f us %s)" % (other[i]['ms'] * 1e3, other[i]['call'][-4:])
        print("%3d %-12s %7.1f us K=%3d %3d->%3d n_out=%7d pairs=%8d %6.0f GB/s%s" % (i, r['call'][19:], r['ms'] * 1e3, r['K'], r['Cin'], r['Cout'], r['n_out'], r['pairs'], r['bytes'] / r['ms'] / 1e6, o))
