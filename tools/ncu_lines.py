"""aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[h]
ix_inst, ix_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
def num(x):
    try: return int(x)
    except Exception: return 0
lines = [(r[0], r[1], num(r[ix_inst]), num(r[ix_s])) for r in rows[h + 1:] if r and r[0].isdigit()]
tot = sum(l[2] for l in lines) or 1; ts = sum(l[3] for l in lines) or 1
print("total instr", tot, "samples", ts)
for ln, src, ins, smp in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%4s %10d %5.1f%%  smp %5.1f%%  %s" % (ln, ins, 100 * ins / tot, 100 * smp / ts, src.strip()[:120]))
