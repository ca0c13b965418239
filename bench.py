#!/usr/bin/env python3
"""bench.py -- scans/sec of the InsMOS sparse-voxel forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one InsMOSNet.forward(batch, 'test') over one synthetic sample of BASELINE config C2:
N=10 stacked scans x 120 000 points (HDL-64E-shaped rays over a synthetic street scene), voxel 0.1 m,
random-init weights (seeded per state_dict key) with BatchNorm statistics calibrated on the input.
  value        whole-job scans/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e          same metric through the public call with HOST buffers: pinned H2D of the points and D2H of
               the per-point logits inside the timed region
  roofline     dominant C-ABI kernel family: sum of algorithmic bytes / sum of CUDA-event durations, vs measured HBM peak
  cpu_baseline the oracle (port of the reference's ME-CPU / spconv algorithms, MKL) on the host cores, bounded sample
--impl reference: the CPU port timed alone (the reference itself cannot run here: MinkowskiEngine / spconv are
un-vendored externals); rank 0 only.
Multi-GPU: samples sharded one per GPU (weak scaling), one padded NCCL all_gather of the logits per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_SCANS, N_ELEV, N_AZIM = 10, 64, 1875
WORKLOAD = "C2: N=10 stacked scans x 120k pts (1.2M points), voxel 0.1 m, full InsMOSNet forward ('test' mode)"
CPU_FULL_SCAN_NOTE = "full-size oracle forward measured once in the dev container: 128 s/scan on 8 cores"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def make_clouds(rank, count, n_azim=N_AZIM):
    from insmos_b200 import synth
    return [synth.make_sequence(seed=100 * rank + i, n_scans=N_SCANS, n_elev=N_ELEV, n_azim=n_azim) for i in range(count)]


def build_model(device, calib_pts):
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    from insmos_b200 import synth_weights
    net = InsMOSNet(default_config())
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth_weights.fill_state_dict(shapes)
    sd["model.unet.center_head.conv_cls.bias"] = torch.zeros(3)      # detections are non-empty
    net.load_state_dict(sd, strict=True)
    net = net.to(device)
    # BatchNorm calibration on the input: one pass with batch statistics written to the running buffers
    bns = [m for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    for m in bns:
        m.momentum = 1.0
    net.train()
    with torch.no_grad():
        net.forward([{"meta": None, "past_point_clouds": calib_pts, "batch_size_npast": N_SCANS}], "test")
    return net.eval()


def step(net, pts):
    boxes, _, logits = net.forward([{"meta": None, "past_point_clouds": pts, "batch_size_npast": N_SCANS}], "test")
    return logits[0], boxes[0][0]


def gather_logits(logits, world):
    """the one exchange step of the path: per-point MOS logits of every rank's sample to all ranks (padded)."""
    from insmos_b200.distributed import gather_logits as g
    parts = g(logits, world)
    return torch.cat(parts, 0), [p.shape[0] for p in parts]


def cpu_port_scans_per_sec(sd_cpu, budget_s, steps=1, warmup=0):
    """time oracle/graph.py (CPU port) on a bounded sample: every n-th azimuth of the same scene; result scaled to
    full-size scans/s by the point ratio (cost is linear in points to first order)."""
    from oracle import graph
    # the port's hot loops are small MKL GEMMs + index_add; beyond ~16 threads they get slower (measured on the
    # 128-core GPU host: 128 threads -> 30x slower than 16), so "all the threads it can use" is capped at 16
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    frac = min(max((budget_s / max(steps + warmup, 1) - 3.0) / 130.0, 1.0 / 64), 1.0 / 4)
    n_azim = max(int(round(N_AZIM * frac)), 24)
    pts = make_clouds(0, 1, n_azim=n_azim)[0]
    for _ in range(warmup):
        graph.forward(sd_cpu, pts)
    t0 = time.perf_counter()
    timing = {}
    for _ in range(steps):
        graph.forward(sd_cpu, pts, timing)
    dt = (time.perf_counter() - t0) / steps
    ratio = n_azim / N_AZIM
    return {"value": ratio / dt, "unit": "scans/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(),
            "kind": "port",
            "sample": "%d scans x %d x %d rays (%.1f%% of the C2 points), %.2f s per sample step, scaled to full-size "
                      "scans/s by the point ratio; BLAS = torch MKL; %s" % (N_SCANS, N_ELEV, n_azim, 100 * ratio, dt, CPU_FULL_SCAN_NOTE),
            "sample_seconds_per_step": dt, "rulebook_s": timing.get("me_maps_s"), "me_conv_s": timing.get("me_conv_s")}


def run_reference(args, rank):
    if rank != 0:
        return
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    from insmos_b200 import synth_weights
    shapes = {k: tuple(v.shape) for k, v in InsMOSNet(default_config()).state_dict().items()}
    sd = synth_weights.fill_state_dict(shapes)
    sd["model.unet.center_head.conv_cls.bias"] = torch.zeros(3)
    cb = cpu_port_scans_per_sec(sd, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "scans_per_sec", "value": cb["value"], "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU port of the reference's MinkowskiEngine-CPU / spconv algorithms "
                       "(oracle/graph.py); uncalibrated BatchNorm statistics (timing only)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-launches", default=None, help="write the per-C-ABI-call profile of one step (JSON lines)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU port)")
    # every distinct input cloud is seen once before the timed region (first-touch sizes hit cudaMalloc in the caching
    # allocator: measured 9.4 -> 11.4 ms/step when the 4th cloud first appeared inside the timed loop)
    args.warmup = max(args.warmup, 4)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"              # stdout carries exactly one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=device)
    from insmos_b200 import _lib

    n_clouds = 4
    host = [torch.from_numpy(c).pin_memory() for c in make_clouds(rank, n_clouds)]
    dev = [h.to(device) for h in host]
    net = build_model(device, dev[0])

    def one(i):
        logits, boxes = step(net, dev[i % n_clouds])
        if world > 1:
            logits, _ = gather_logits(logits, world)
        return logits

    # e2e: the public host-buffer path (insmos_b200.pipeline.ScanPipeline): raw scans [N_i,4] in host
    # memory + poses -> pinned H2D -> staging kernel (pose transform, time stamps) -> forward -> (NCCL gather) -> label
    # kernel -> D2H of labels + confidence.  Two samples in flight: the copies of one overlap the forward of the other.
    from insmos_b200.pipeline import ScanPipeline
    host_scans = []                                    # per cloud: (pinned [total,4] raw scans back to back, int64 offsets)
    for h in host:
        a = h.numpy()
        stamps = np.unique(a[:, 4])
        parts = [a[a[:, 4] == t][:, :4] for t in stamps]
        offs = np.concatenate([[0], np.cumsum([len(q) for q in parts])]).astype(np.int64)
        host_scans.append((torch.from_numpy(np.ascontiguousarray(np.concatenate(parts, 0))).pin_memory(), offs))
    poses = [np.eye(4)] * N_SCANS                      # synthetic clouds are already in the newest frame: T = I (same kernel work)
    max_pts = max(int(h.shape[0]) for h in host) + 1024
    pipe = ScanPipeline(net, dt_pred=0.1, n_scans=N_SCANS, max_points=max_pts,
                        post_logits=(lambda lg: gather_logits(lg, world)[0]) if world > 1 else None,
                        out_rows=max_pts * world)

    def timed(e2e):
        import gc
        gc.collect()
        gc.disable()                                   # no collector pauses inside the timed region (re-enabled below)
        try:
            return _timed(e2e)
        finally:
            gc.enable()

    def _timed(e2e):
        with torch.no_grad():
            last = None
            if e2e:
                for i in range(args.warmup):
                    pipe.result(pipe.submit_packed(*host_scans[i % n_clouds], poses))
            else:
                for i in range(args.warmup):
                    one(i)
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            launches0 = _lib.LAUNCHES
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            if e2e:
                prev = None
                for i in range(args.steps):
                    t = pipe.submit_packed(*host_scans[i % n_clouds], poses)
                    if prev is not None:
                        last = pipe.result(prev)                     # read sample i-1 on the host while sample i runs
                    prev = t
                last = pipe.result(prev)
                out = torch.from_numpy(last["labels"])
            else:
                for i in range(args.steps):
                    out = one(i)
            e.record()
            torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e)], device=device)
            if dist:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                dist.barrier()
        return float(ms.item()), _lib.LAUNCHES - launches0, out

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, out = timed(False)
    sampler.stop_flag = True
    ms_e2e, _, out_host = timed(True)
    value = world * args.steps / (ms / 1000.0)
    e2e_value = world * args.steps / (ms_e2e / 1000.0)
    n_cur = int(out_host.shape[0] // world)

    # ---- per-kernel-family breakdown: CUDA events around every C-ABI call, 2 instrumented steps after the timed region
    fam, step_ms_prof = {}, None
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.profile_start()
        for i in range(2):
            step(net, dev[i % n_clouds])
        prof = _lib.profile_stop()
        step_ms_prof = (time.perf_counter() - t0) * 1000 / 2
    for name, t, meta in prof:
        f = fam.setdefault(name, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
        f["ms"] += t / 2
        f["launches"] += 0.5
        if meta:
            f["bytes"] += meta.get("bytes", 0) / 2
            f["flops"] += meta.get("flops", 0) / 2
    conv_launches = sorted(({"ms": round(t, 4), **{k: m[k] for k in ("K", "Cin", "Cout", "n_out", "pairs")},
                             "alg_GBps": round(m["bytes"] / (t * 1e-3) / 1e9, 1)}
                            for name, t, m in prof[len(prof) // 2:] if m and "Cin" in m), key=lambda d: -d["ms"])[:16]
    if args.dump_launches and rank == 0:
        with open(args.dump_launches, "w") as fh:
            for name, t, m in prof[len(prof) // 2:]:
                fh.write(json.dumps({"call": name, "ms": round(t, 4), **(m or {})}) + "\n")
    peak, peak_src = peaks()
    top = max((k for k in fam if fam[k]["bytes"] > 0), key=lambda k: fam[k]["ms"])
    ach = fam[top]["bytes"] / (fam[top]["ms"] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per kernel family over ONE forward (ncu, tools/gpu_profile.sh)
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(top, {}).get("dram_bytes_per_forward")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "peak_source": peak_src,
                "definition": "sum of algorithmic bytes of this kernel family's launches in one step / sum of their CUDA-event "
                              "durations (2 instrumented steps after the timed region); launches per step: %d; traffic = ncu "
                              "dram__bytes_read.sum + dram__bytes_write.sum summed over the same launches of one forward "
                              "(profiles/r01_fwd_v7_summary.txt)" % round(fam[top]["launches"]),
                "alg_bytes_per_step": int(fam[top]["bytes"]), "ms_per_step": round(fam[top]["ms"], 4)}
    kernels = {k: {"ms_per_step": round(v["ms"], 4), "launch_calls": v["launches"],
                   "alg_GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] and v["ms"] > 0 else None,
                   "gflop": round(v["flops"] / 1e9, 2) if v["flops"] else None}
               for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            cpu = cpu_port_scans_per_sec(sd_cpu, budget_s=22.0)
        line = {
            "metric": "scans_per_sec", "value": round(value, 3), "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "samples sharded 1 per GPU, NCCL all_gather of logits" if world > 1 else "single GPU",
                       "l2": "no explicit flush: each step streams > 126 MB (rule books + features) and inputs rotate over %d distinct clouds" % n_clouds,
                       "arithmetic": "fp32; sparse-conv products on tensor cores as 3xTF32 (fp32-accurate); cuDNN TF32 off",
                       "points_per_step": int(dev[0].shape[0]), "current_points": n_cur},
            "e2e": {"value": round(e2e_value, 3), "unit": "scans/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": int(host[0].shape[0] * 16 + N_SCANS * 17 * 8 + (N_SCANS + 1) * 8),
                    "d2h_bytes_per_step": int(n_cur * world * 12),
                    "path": "insmos_b200.pipeline.ScanPipeline: raw scans + poses in pinned host memory -> H2D -> staging kernel -> "
                            "forward -> label kernel -> D2H (labels int32 + confidence 2 x f32 per point); 2 samples in flight"},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline,
            "kernels": kernels, "slowest_sparse_convs": conv_launches, "profiled_step_ms": round(step_ms_prof, 3),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
