"""summarise `ncu --metrics ... --csv --log-file fwd_metrics.csv` of ONE forward: per-kernel-name totals, per-family DRAM
traffic (profiles/traffic.json, read by bench.py for roofline.traffic) and a launch list."""
import csv, json, sys, re
from collections import defaultdict, OrderedDict

src, out_prefix = sys.argv[1], sys.argv[2]
rows = []
with open(src) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
launch = OrderedDict()
for r in rd:
    lid = r["ID"]
    d = launch.setdefault(lid, {"name": r["Kernel Name"], "grid": r.get("Grid Size"), "block": r.get("Block Size")})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)                # -> us
    if name.startswith("dram__bytes") or name == "lts__t_bytes.sum":
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)   # -> bytes
    d[name] = v

FAMILY = [("k_spconv_tc", "insmos_sparse_conv_fwd_tc"), ("k_spconv_umma", "insmos_sparse_conv_fwd_umma"),
          ("k_rulebook", "insmos_rulebook_build"), ("k_xblock", "insmos_rulebook_build"), ("k_leafgrid", "insmos_rulebook_build"),
          ("k_spconv_cin1", "insmos_sparse_conv_fwd_ffma"), ("k_spconv_fma", "insmos_sparse_conv_fwd_fma"),
          ("k_spconv_ff", "insmos_sparse_conv_fwd_ffma"), ("k_linear", "insmos_linear_fwd"), ("k_nms", "insmos_nms_rotated_pairs"),
          ("k_member", "insmos_box_membership"), ("k_conv_nhwc", "insmos_conv2d_nhwc_tcgen05")]
def short(n):
    return re.sub(r"\(.*", "", n).replace("void ", "")
by = defaultdict(lambda: defaultdict(float))
fam = defaultdict(lambda: defaultdict(float))
tot_us = 0.0
for d in launch.values():
    k = short(d["name"])
    b = by[k]
    b["n"] += 1
    for m in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"):
        b[m] += d.get(m, 0.0)
    for m in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"):
        b[m] += d.get(m, 0.0) * d.get("gpu__time_duration.sum", 0.0)                      # time-weighted
    tot_us += d.get("gpu__time_duration.sum", 0.0)
    for pat, f in FAMILY:
        if pat in k:
            fam[f]["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            fam[f]["l2_bytes"] += d.get("lts__t_bytes.sum", 0.0)
            fam[f]["us"] += d.get("gpu__time_duration.sum", 0.0)
            fam[f]["launches"] += 1
            break
with open(out_prefix + "_summary.txt", "w") as fh:
    fh.write("ncu --metrics (gpu__time_duration, dram bytes, issue / tensor / warps active, L2) --clock-control none over ONE forward "
             "(NVTX range of tools/one_forward.py, C2 workload); per-launch times are serialised: compare SHARES\n")
    fh.write("%d launches, %.1f us total\n" % (len(launch), tot_us))
    fh.write("%-44s %4s %9s %6s %9s %9s %7s %7s %7s %7s\n" % ("kernel", "n", "us", "share", "dramMB", "L2 MB", "issue%", "tensor%", "warps%", "L2hit%"))
    for k, b in sorted(by.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = b["gpu__time_duration.sum"] or 1e-9
        fh.write("%-44s %4d %9.1f %5.1f%% %9.1f %9.1f %7.1f %7.1f %7.1f %7.1f\n" % (
            k[:44], b["n"], t, 100 * t / tot_us, (b["dram__bytes_read.sum"] + b["dram__bytes_write.sum"]) / 1e6, b["lts__t_bytes.sum"] / 1e6,
            b["smsp__issue_active.avg.pct_of_peak_sustained_active"] / t, b["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] / t,
            b["sm__warps_active.avg.pct_of_peak_sustained_active"] / t, b["lts__t_sector_hit_rate.pct"] / t))
import hashlib, os
def _src_digest():
    h = hashlib.sha256()
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "insmos_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]
tj = {"_src_digest": _src_digest(), "_from": os.path.basename(out_prefix) + "_summary.txt (ncu over one forward of tools/one_forward.py)"}
tj.update({f: {"dram_bytes_per_forward": int(v["dram_bytes"]), "l2_bytes_per_forward": int(v["l2_bytes"]), "ncu_us": round(v["us"], 1),
               "launches": int(v["launches"])} for f, v in fam.items()})
json.dump(tj, open(out_prefix + "_traffic.json", "w"), indent=1)
print(open(out_prefix + "_summary.txt").read())
