#!/usr/bin/env python3
"""bench.py -- scans/sec of the InsMOS sparse-voxel forward path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4|train]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one InsMOSNet.forward(batch, 'test') over one synthetic sample of BASELINE config C2:
N=10 stacked scans x 120 000 points (HDL-64E-shaped rays over a synthetic street scene), voxel 0.1 m, seeded weights with
the BatchNorm statistics of tests/golden/insmos_c2.npz -- the exact weights and first cloud the C2 parity tests
(tests/test_gpu_c2_golden.py) check against the reference-code golden.
  value        whole-job scans/s with inputs resident in HBM (CUDA events, max over ranks); the K-step region is repeated
               until >= 2 s have been timed and the MEDIAN region is reported (min/max beside it)
  e2e          same metric through the public call with HOST buffers: pinned H2D of the raw scans, staging kernel, forward,
               label kernel and D2H of the labels inside the timed region (insmos_b200.pipeline.ScanPipeline)
  roofline     dominant C-ABI kernel family: sum of algorithmic bytes / sum of CUDA-event durations, vs measured HBM peak
  cpu_baseline the oracle (port of the reference's ME-CPU / spconv algorithms: hash-map kernel maps with OpenMP over the offsets,
               per-offset MKL sgemm) on the host cores: full-size C2 forwards (one warm-up, up to 4 timed, ~15 s), no extrapolation
--impl reference: the CPU port timed alone on the FULL C2 cloud (the reference itself cannot run here: MinkowskiEngine /
spconv are un-vendored externals); as many steps as fit the time budget, reported truthfully; rank 0 only.
--workload c4: BASELINE config 4 (300 k points per scan, voxel 0.05 m): rule-book build + gather/scatter sweep line.
--workload train: BASELINE config 5 (training step: forward + backward + gradient all-reduce + Adam on C2 samples with boxes / labels).
Multi-GPU: samples sharded one per GPU (weak scaling), one fixed-size NCCL all_gather of the logits per step, no host sync.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_SCANS, N_ELEV, N_AZIM = 10, 64, 1875
WORKLOAD = "C2: N=10 stacked scans x 120k pts (1.2M points), voxel 0.1 m, full InsMOSNet forward ('test' mode)"
MIN_TIMED_S = 2.0                      # the K-step region is repeated until this much has been timed
GATHER_PAD_ROWS = 122_880              # fixed block of the logits all_gather (>= points of one scan)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def source_digest():
    """sha256 over the CUDA sources: ties profiles/traffic.json (ncu DRAM bytes) to the build it was measured on."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "insmos_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index, enabled=True, interval=0.1):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        # ONE sampler per job (rank 0, its own GPU): eight ranks polling nvidia-smi at 10 Hz serialise on the driver and on the
        # host cores the launch threads need (8-GPU run: 6.86 ms/step with every rank sampling)
        self.enabled, self.interval = enabled, interval

    def _nvml(self):
        """in-process NVML handle (same counters as the nvidia-smi query of the profiling recipe, without spawning a process and
        taking the driver's management lock four times a second: one evidence run showed the device-resident phase -- the only
        one sampled -- at 5.6-8.7 ms per step while the unsampled end-to-end phase on the same box held 5.8 ms)"""
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        nv, h = self._nvml() if self.enabled else (None, None)
        while self.enabled and not self.stop_flag:
            try:
                if nv is not None:
                    try:
                        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    except Exception:
                        nv = None                                          # NVML unusable here: fall back to the nvidia-smi query
                        self.interval = max(self.interval, 0.25)
                        continue
                    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                    try:
                        r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    act = lambda bit: "Active" if r & bit else "Not Active"          # noqa: E731
                    self.samples.append([str(sm), str(mx), act(nv.nvmlClocksThrottleReasonHwSlowdown),
                                         act(nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                                         act(nv.nvmlClocksThrottleReasonSwThermalSlowdown), act(nv.nvmlClocksThrottleReasonSwPowerCap)])
                    self.source = "nvml"
                else:
                    r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5)
                    f = [x.strip() for x in r.stdout.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
                    self.source = "nvidia-smi"
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm),
                "source": getattr(self, "source", None)}


def make_clouds(rank, count, n_azim=N_AZIM):
    from insmos_b200 import synth
    return [synth.make_sequence(seed=100 * rank + i, n_scans=N_SCANS, n_elev=N_ELEV, n_azim=n_azim) for i in range(count)]


def golden_state_dict():
    """the C2 golden's weights: seeded per key + the BatchNorm statistics the reference code calibrated on cloud seed 0."""
    import golden_util
    meta, shapes, sd, _, _ = golden_util.load("c2", with_points=False)
    return sd


def build_model(device, calib_pts=None):
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    net = InsMOSNet(default_config())
    net.load_state_dict(golden_state_dict(), strict=True)
    return net.to(device).eval()


def step(net, pts):
    boxes, _, logits = net.forward([{"meta": None, "past_point_clouds": pts, "batch_size_npast": N_SCANS}], "test")
    return logits[0], boxes[0][0]


def cpu_port_full(sd_cpu, budget_s, max_steps, warmup=0):
    """time oracle/graph.py (CPU port) on the FULL C2 cloud (seed 0): at least one step, more while they fit the budget;
    `warmup` untimed steps first, as many of them as fit a fifth of the budget (the count actually run is returned)."""
    from oracle import graph
    # the port's hot loops are small MKL GEMMs + index_add; beyond ~16 threads they get slower (measured on the
    # 128-core GPU host: 128 threads -> 30x slower than 16), so "all the threads it can use" is capped at 16
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    from oracle import native as oracle_native
    oracle_native.set_threads(torch.get_num_threads())        # torchrun exports OMP_NUM_THREADS=1: the native loops get the same count
    pts = make_clouds(0, 1)[0]
    times, timing = [], {}
    warm_done, t_begin = 0, time.perf_counter()
    while warm_done < warmup:
        t0 = time.perf_counter()
        graph.forward(sd_cpu, pts, {})
        warm_done += 1
        if time.perf_counter() - t_begin + (time.perf_counter() - t0) > 0.2 * budget_s:
            break
    t_begin = time.perf_counter()
    while len(times) < max(max_steps, 1):
        t0 = time.perf_counter()
        graph.forward(sd_cpu, pts, timing)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin + times[-1] > budget_s:
            break
    dt = float(np.mean(times))
    return {"value": 1.0 / dt, "unit": "scans/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "port",
            "sample": "the full C2 cloud (10 scans x 64 x 1875 rays = 1.2 M points, seed 0), %d forward(s) of oracle/graph.py, "
                      "%.1f s each, no extrapolation; BLAS = torch MKL, kernel-map lookups = oracle/native (hash map, OpenMP)" % (len(times), dt),
            "steps": len(times), "warmup_steps": warm_done, "seconds_per_step": dt,
            "rulebook_s": timing.get("me_maps_s"), "me_conv_s": timing.get("me_conv_s"), "motionnet_s": timing.get("motionnet_s"),
            "unet_encoder_s": timing.get("unet_encoder_s"), "bev_s": timing.get("bev_s"), "decoder_s": timing.get("decoder_s")}


def run_reference(args, rank, out_stream):
    if rank != 0:
        return
    sd = golden_state_dict()
    cb = cpu_port_full(sd, budget_s=float(os.environ.get("INSMOS_REF_BUDGET_S", "170")), max_steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "scans_per_sec", "value": cb["value"], "unit": "scans/s", "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": cb["warmup_steps"], "requested": {"steps": args.steps, "warmup": args.warmup},
            "ms_per_step": 1000.0 * cb["seconds_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU port of the reference's MinkowskiEngine-CPU / spconv algorithms "
                       "(oracle/graph.py) on the full C2 cloud with the golden's weights: kernel maps by hash map + OpenMP over "
                       "the offsets (oracle/native), convolutions as per-offset gather -> MKL sgemm -> scatter-add; one full-size "
                       "forward takes seconds, so the run times as many of the requested steps as fit ~3 min (warm-ups within a fifth "
                       "of that) and reports the counts actually run"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out_stream.write(json.dumps(line) + "\n")
    out_stream.flush()


def run_c4(args, device, out_stream):
    """BASELINE config 4: 300 k points per scan x 10, voxel 0.05 m -- per-map build time / pairs and the gather-scatter
    kernels' algorithmic GB/s (the 100 k-voxel cap makes the full model meaningless at this density)."""
    from insmos_b200 import ops, synth, _lib
    pts = torch.from_numpy(synth.make_sequence(seed=4, n_scans=10, n_elev=160, n_azim=1875)).to(device)
    peak, peak_src = peaks()

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            r = fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps, r

    reps = max(args.steps, 5)
    rows = []
    ms, (cs, inverse, cur) = timed(lambda: ops.voxelize4d(pts, [0.05, 0.05, 0.05, 0.1]), reps)
    n_pts = pts.shape[0]
    b = 20 * n_pts + 20 * cs.n + 8 * n_pts
    rows.append({"op": "voxelize4d", "ms": round(ms, 4), "n_points": n_pts, "n_voxels": cs.n, "alg_GBps": round(b / ms / 1e6, 1)})
    ms, (c2s, parent) = timed(lambda: ops.unique_coords(cs.coords, q=[2, 2, 2, 1]), reps)
    rows.append({"op": "stride_coords 1->2", "ms": round(ms, 4), "n_out": c2s.n, "alg_GBps": round((20 * cs.n + 20 * c2s.n + 4 * cs.n) / ms / 1e6, 1)})
    g = torch.Generator().manual_seed(0)
    for name, out_set, in_set, ksize, ist, xs, chans in (
            ("5x5x5x1 ts1", cs, cs, [5, 5, 5, 1], [1, 1, 1, 1], 1, [(1, 8)]),
            ("3x3x3x3 ts1", cs, cs, [3, 3, 3, 3], [1, 1, 1, 1], 1, [(8, 8), (16, 8)]),
            ("2x2x2x1 ts1->ts2", c2s, cs, [2, 2, 2, 1], [1, 1, 1, 1], None, [(8, 8)]),
            ("3x3x3x3 ts2", c2s, c2s, [3, 3, 3, 3], [2, 2, 2, 1], 2, [(8, 16), (16, 16)])):
        spec = ops.spec_me_cube(ksize, ist)
        ms, rb = timed(lambda: ops.build_rulebook(out_set, in_set, spec, xstep=xs, step=tuple(ist)), reps)
        P, K = rb.num_pairs, int(np.prod(ksize))
        b = 4 * out_set.ncol * out_set.n + 8 * P
        rows.append({"op": "rulebook " + name, "ms": round(ms, 4), "n_out": out_set.n, "K": K, "pairs": P,
                     "alg_GBps": round(b / ms / 1e6, 1), "frac_of_peak": round(b / ms / 1e6 / peak, 4)})
        for Cin, Cout in chans:
            x = torch.randn((in_set.n, Cin), generator=g).to(device)
            W = (torch.randn((K, Cin, Cout), generator=g) / np.sqrt(Cin * 8.0)).to(device)
            ms, _ = timed(lambda: ops.sparse_conv(x, W, rb), reps)
            b = 4 * (in_set.n * Cin + out_set.n * Cout) + 8 * P + 4 * K * Cin * Cout
            rows.append({"op": "sparse_conv %s %d->%d" % (name, Cin, Cout), "ms": round(ms, 4), "pairs": P,
                         "alg_GBps": round(b / ms / 1e6, 1), "frac_of_peak": round(b / ms / 1e6 / peak, 4),
                         "gflops": round(2 * P * Cin * Cout / ms / 1e6, 1)})
    out_stream.write(json.dumps({"metric": "c4_sweep", "workload": "C4: N=10 x 300k pts (3.0 M points), voxel 0.05 m: rule-book build + "
                      "gather/scatter sparse conv sweep", "peak_GBps": peak, "peak_source": peak_src, "reps": reps, "rows": rows,
                      "gpu_launches": _lib.launch_count()}) + "\n")
    out_stream.flush()


def run_train(args, device, rank, local_rank, world, out_stream):
    """BASELINE config 5: one TRAINING step per sample of the C2 workload (+ synthetic boxes and per-point MOS labels):
    forward in train mode, losses, backward through this repository's gradient kernels, ONE NCCL all-reduce of the flat
    gradient buffer, fused Adam.  Weak scaling: every rank trains on its own sample, weights replicated ("ZeRO-0 DDP")."""
    import gc
    from insmos_b200 import _lib, synth
    from insmos_b200.config import default_config
    from insmos_b200.train import TrainStep
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    n_clouds = 3
    batches = []
    for i in range(n_clouds):
        pts, labels, boxes = synth.make_sequence(seed=100 * rank + i, n_scans=N_SCANS, n_elev=N_ELEV, n_azim=N_AZIM, return_labels=True)
        batches.append((torch.from_numpy(pts).to(device), torch.from_numpy(labels.astype(np.float32)).to(device),
                        torch.from_numpy(boxes).unsqueeze(0).to(device)))
    net = build_model(device).train()
    ts = TrainStep.from_config(net, default_config())

    def batch(i):
        pts, labels, boxes = batches[i % n_clouds]
        return [{"meta": None, "past_point_clouds": pts, "past_labels": [labels], "gt_boxes": boxes, "batch_size_npast": N_SCANS}]

    warmup = max(args.warmup, 12)         # the caching allocator / cuDNN settle after ~10 steps (measured: 90-110 ms, then 45 ms)
    losses = [ts.step(batch(i)) for i in range(warmup)]
    gc.collect()
    gc.disable()
    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    sampler.start()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    launches0 = _lib.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    walls, t_prev = [], time.perf_counter()
    for i in range(args.steps):
        last = ts.step(batch(warmup + i))                 # (returns python floats: one read-back = one sync per step)
        t_now = time.perf_counter()
        walls.append(round((t_now - t_prev) * 1000, 2))
        t_prev = t_now
    e.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    sampler.stop_flag = True
    gc.enable()
    ms = torch.tensor([s.elapsed_time(e)], device=device)
    if dist:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    ms = float(ms.item())
    # per-family breakdown of one instrumented step (CUDA events around every C-ABI call; torch kernels show up as gaps)
    _lib.profile_start()
    t0 = time.perf_counter()
    ts.step(batch(0))
    prof = _lib.profile_stop()
    prof_ms = (time.perf_counter() - t0) * 1000
    if args.dump_launches and rank == 0:
        with open(args.dump_launches, "w") as fh:
            for name, t, m in prof:
                fh.write(json.dumps({"call": name, "ms": round(t, 4), **(m or {})}) + "\n")
    fam = {}
    for name, t, meta in prof:
        f = fam.setdefault(name, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
        f["ms"] += t
        f["launches"] += 1
        if meta:
            f["bytes"] += meta.get("bytes", 0)
            f["flops"] += meta.get("flops", 0)
    peak, peak_src = peaks()
    wg = fam.get("insmos_sparse_conv_wgrad", {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
    kernels = {k: {"ms_per_step": round(v["ms"], 4), "launch_calls": v["launches"],
                   "alg_GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] and v["ms"] > 0 else None,
                   "gflop": round(v["flops"] / 1e9, 2) if v["flops"] else None}
               for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    if rank == 0:
        ach = wg["bytes"] / (wg["ms"] * 1e-3) / 1e9 if wg["ms"] > 0 else 0.0
        line = {
            "metric": "train_scans_per_sec", "value": round(world * args.steps / (ms / 1000.0), 3), "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5: training step on C2 samples (N=10 x 120k pts) + synthetic boxes [1,3,8] and per-point MOS labels: forward "
                                   "('train' mode, batch-statistics BatchNorm), loss_rpn + loss_mos + loss_motion, backward, gradient all-reduce, Adam",
                       "parallelism": ("one sample per GPU per step, weights replicated, ONE in-place NCCL all-reduce of the %.1f MB flat gradient "
                                       "buffer per step, fused Adam on the flat buffers" % (ts.flat.numel * 4 / 1e6)) if world > 1 else "single GPU",
                       "arithmetic": "fp32; sparse-conv forward and data gradient on tensor cores as 3xTF32, weight gradient fp32 FFMA; dense BEV convs cuDNN fp32 (TF32 off)",
                       "parameters": int(ts.flat.numel), "points_per_step": int(batches[0][0].shape[0]),
                       "l2": "no explicit flush: every step streams > 126 MB and inputs rotate over %d distinct clouds" % n_clouds},
            "losses_first_last": [losses[0], last], "step_wall_ms": walls,
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": "insmos_sparse_conv_wgrad", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                         "definition": "algorithmic bytes of the weight-gradient launches of one step (SURVEY 8d formula per layer) / their CUDA-event time",
                         "gflop": round(wg["flops"] / 1e9, 2), "ms_per_step": round(wg["ms"], 4)},
            "kernels": kernels, "profiled_step_ms": round(prof_ms, 3),
        }
        out_stream.write(json.dumps(line) + "\n")
        out_stream.flush()
    if dist:
        dist.destroy_process_group()


def bind_to_gpu_cpus(index):
    """pin this process to the CPU cores the driver reports as local to GPU `index` (NUMA node of its PCIe root): with one
    process per GPU and no binding, half of the ranks of an 8-GPU box queue their launches from the remote socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything else any library writes to fd 1 (NCCL's version banner, torchrun
    notices) is sent to stderr; returns the stream that still points at the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-timed-s", type=float, default=MIN_TIMED_S)
    ap.add_argument("--streams", type=int, default=2, help="forwards in flight on separate CUDA streams / host threads (0: the calling thread only)")
    ap.add_argument("--dump-launches", default=None, help="write the per-C-ABI-call profile of one step (JSON lines)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, out_stream)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    bound = bind_to_gpu_cpus(local_rank) if world > 1 else 0
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if args.workload == "c4":
        run_c4(args, device, out_stream)
        return
    if args.workload == "train":
        run_train(args, device, rank, local_rank, world, out_stream)
        return
    # every distinct input cloud is seen once before the timed region (first-touch sizes hit cudaMalloc in the caching
    # allocator: measured 9.4 -> 11.4 ms/step when the 4th cloud first appeared inside the timed loop)
    warmup = max(args.warmup, 4)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from insmos_b200 import _lib
    from insmos_b200.distributed import gather_logits_padded

    n_clouds = 4
    # weak scaling = the SAME work on every GPU: all ranks draw from the same four clouds (seeds 0-3), each rank starting at
    # a different one.  (Per-rank seeds made the ranks' work unequal: the synthetic clouds hold 296 k - 506 k 4D voxels, the
    # four-cloud mean of a rank ranged 398 k - 443 k, and a K-step region ends with the slowest rank.)
    host = [torch.from_numpy(c).pin_memory() for c in make_clouds(0, n_clouds)]
    host = host[rank % n_clouds:] + host[:rank % n_clouds]
    dev = [h.to(device) for h in host]
    net = build_model(device)
    gather_out = torch.empty((world, GATHER_PAD_ROWS + 1, 3), dtype=torch.float32, device=device) if world > 1 else None

    # `streams` forwards are kept in flight on separate CUDA streams / host threads (insmos_b200.engine.ForwardPool): the
    # data-dependent host reads of one sample overlap the kernels of the others.  The exchange step is issued from this
    # thread in step order (identical collective order on every rank).
    from insmos_b200.engine import ForwardPool
    pool = ForwardPool(net, workers=max(args.streams, 1), n_past=N_SCANS) if args.streams > 0 else None
    main_stream = torch.cuda.current_stream(device)

    def finish(job):
        logits, boxes = job.wait(main_stream)
        logits.record_stream(main_stream)
        if world > 1:                                   # the one exchange step: fixed-size all_gather, no host sync
            gather_logits_padded(logits, world, GATHER_PAD_ROWS, out=gather_out)
        return logits

    def one(i):
        logits, boxes = step(net, dev[i % n_clouds])
        if world > 1:
            gather_logits_padded(logits, world, GATHER_PAD_ROWS, out=gather_out)
        return logits

    def run_steps(k0, k):
        """k forwards (+ gathers); returns the last logits"""
        if pool is None:
            for i in range(k):
                out = one(k0 + i)
            return out
        jobs, out = [], None
        for i in range(k):
            jobs.append(pool.submit_points(dev[(k0 + i) % n_clouds]))
            if len(jobs) > args.streams:
                out = finish(jobs.pop(0))
        while jobs:
            out = finish(jobs.pop(0))
        return out

    # e2e: the public host-buffer path (insmos_b200.pipeline.ScanPipeline): raw scans [N_i,4] in host
    # memory + poses -> pinned H2D -> staging kernel (pose transform, time stamps) -> forward -> label kernel -> D2H of
    # this rank's labels + confidence (+ NCCL gather of the logits on the device).  Two samples in flight.
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
    pipe = ScanPipeline(net, dt_pred=0.1, n_scans=N_SCANS, max_points=max_pts, gather_world=world, gather_pad_rows=GATHER_PAD_ROWS,
                        workers=2 if args.streams > 0 else 0)

    def region(e2e):
        """exactly args.steps steps between two events, barrier + synchronize on both sides; -> (ms max over ranks, out)"""
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
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
            out = run_steps(0, args.steps)
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=device)
        if dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item()), out

    def timed(e2e):
        import gc
        gc.collect()
        gc.disable()                                   # no collector pauses inside the timed regions (re-enabled below)
        try:
            with torch.no_grad():
                for i in range(warmup):
                    if e2e:
                        pipe.result(pipe.submit_packed(*host_scans[i % n_clouds], poses))
                    else:
                        run_steps(i, 1)
                launches0 = _lib.launch_count()
                ms0, out = region(e2e)
                launches = _lib.launch_count() - launches0
                # repeat the K-step region until >= min_timed_s have been timed; every rank runs the same count
                reps = int(min(max(np.ceil(args.min_timed_s * 1000.0 / max(ms0, 1e-3)), 1), 400))
                if dist:
                    t = torch.tensor([reps], device=device)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    reps = int(t.item())
                regions = [ms0] + [region(e2e)[0] for _ in range(reps - 1)]
            return regions, launches, out
        finally:
            gc.enable()

    sampler = ClockSampler(local_rank, enabled=(rank == 0))
    sampler.start()
    regions, launches, out = timed(False)
    sampler.stop_flag = True
    regions_e2e, _, out_host = timed(True)
    ms, ms_e2e = float(np.median(regions)), float(np.median(regions_e2e))
    value = world * args.steps / (ms / 1000.0)
    e2e_value = world * args.steps / (ms_e2e / 1000.0)
    n_cur = int(out_host.shape[0])

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
                            for name, t, m in prof[:len(prof) // 2] if m and "Cin" in m), key=lambda d: -d["ms"])[:16]
    if args.dump_launches and rank == 0:
        with open(args.dump_launches, "w") as fh:
            for name, t, m in prof[:len(prof) // 2]:
                fh.write(json.dumps({"call": name, "ms": round(t, 4), **(m or {})}) + "\n")
    peak, peak_src = peaks()
    # `roofline` stays on the narrow-layer sparse-conv family -- the gather -> tensor-core -> scatter kernel VERDICT r01 names and
    # SURVEY 8d's "sparse-conv gather/scatter vs HBM" target; the other families are listed in `roofline_families`
    top = "insmos_sparse_conv_fwd_tc" if fam.get("insmos_sparse_conv_fwd_tc", {}).get("bytes", 0) > 0 else \
        max((k for k in fam if fam[k]["bytes"] > 0), key=lambda k: fam[k]["ms"])
    ach = fam[top]["bytes"] / (fam[top]["ms"] * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu DRAM capture of this build (profiles/traffic.json absent or made from other sources)"
    tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per kernel family over ONE forward (ncu)
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if tj.get("_src_digest") == source_digest():
                traffic = tj.get(top, {}).get("dram_bytes_per_forward")
                traffic_note = "ncu dram__bytes_read.sum + dram__bytes_write.sum over this family's launches of one forward of THIS build (%s)" % tj.get("_from", "profiles/")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": top, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                "definition": "sum of algorithmic bytes of this kernel family's launches in one step / sum of their CUDA-event "
                              "durations (2 instrumented steps after the timed region); launches per step: %d" % round(fam[top]["launches"]),
                "alg_bytes_per_step": int(fam[top]["bytes"]), "ms_per_step": round(fam[top]["ms"], 4)}
    kernels = {k: {"ms_per_step": round(v["ms"], 4), "launch_calls": v["launches"],
                   "alg_GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] and v["ms"] > 0 else None,
                   "gflop": round(v["flops"] / 1e9, 2) if v["flops"] else None}
               for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    # every family with an algorithmic-byte model against the bound that fits it: HBM for the gather/scatter and map-build
    # kernels; the wide-layer tcgen05 kernel is arithmetic-bound -- issued TF32 flops (3 products per fp32 product) against the
    # dense TF32 peak (half of the measured bf16 matmul throughput of MEASURED_PEAKS.json)
    tf32_peak = None
    try:
        tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]) / 2.0
    except Exception:
        tf32_peak = 1100.0 * 0.85
    roofline_families = []
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        if not v["bytes"] or v["ms"] <= 0:
            continue
        ent = {"kernel": k, "ms_per_step": round(v["ms"], 4), "bound": "hbm", "achieved": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
               "peak": peak, "unit": "GB/s", "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 4)}
        if k == "insmos_sparse_conv_fwd_umma" and v["flops"]:
            tf = 3.0 * v["flops"] / (v["ms"] * 1e-3) / 1e12
            ent.update({"bound": "tensor", "achieved": round(tf, 1), "peak": round(tf32_peak, 1), "unit": "TFLOP/s (TF32 issued, 3 per fp32 product)",
                        "frac": round(tf / tf32_peak, 4), "hbm_frac": ent["frac"]})
        roofline_families.append(ent)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd_cpu = {k: v.detach().cpu() for k, v in net.state_dict().items()}
            # full-size forwards of the CPU port, no extrapolation: one warm-up, then up to 4 timed steps within ~15 s
            cpu = cpu_port_full(sd_cpu, budget_s=15.0, max_steps=4, warmup=1)
        line = {
            "metric": "scans_per_sec", "value": round(value, 3), "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": "samples sharded 1 per GPU, one fixed-size NCCL all_gather of logits per step" if world > 1 else "single GPU",
                       "l2": "no explicit flush: each step streams > 126 MB (rule books + features) and inputs rotate over %d distinct clouds" % n_clouds,
                       "arithmetic": "fp32; sparse-conv products on tensor cores as 3xTF32 (fp32-accurate) or fp32 FFMA; cuDNN TF32 off",
                       "weights": "tests/golden/insmos_c2.npz (the C2 parity golden's weights)",
                       "streams": args.streams, "cpus_bound_per_rank": bound,
                       "clouds": "every rank processes the same %d synthetic clouds (seeds 0-%d), rotated by rank: identical work per GPU" % (n_clouds, n_clouds - 1),
                       "points_per_step": int(dev[0].shape[0]), "current_points": n_cur},
            "timing": {"regions": len(regions), "steps_per_region": args.steps, "timed_s": round(sum(regions) / 1000.0, 3),
                       "ms_per_step_median": round(ms / args.steps, 4), "ms_per_step_min": round(min(regions) / args.steps, 4),
                       "ms_per_step_max": round(max(regions) / args.steps, 4),
                       "note": "value = median K-step region; each region is exactly K steps between barrier+synchronize"},
            "e2e": {"value": round(e2e_value, 3), "unit": "scans/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "regions": len(regions_e2e), "ms_per_step_min": round(min(regions_e2e) / args.steps, 4),
                    "ms_per_step_max": round(max(regions_e2e) / args.steps, 4),
                    "h2d_bytes_per_step": int(host[0].shape[0] * 16 + N_SCANS * 17 * 8 + (N_SCANS + 1) * 8),
                    "d2h_bytes_per_step": int(n_cur * 12),
                    "path": "insmos_b200.pipeline.ScanPipeline: raw scans + poses in pinned host memory -> H2D -> staging kernel -> "
                            "forward -> label kernel -> D2H of this rank's sample (labels int32 + confidence 2 x f32 per point); "
                            "2 samples in flight; bytes are per rank"},
            "gpu_launches": int(launches), "gpu_launches_note": "counted inside libinsmos_b200.so at every kernel launch site over the first K-step region (insmos_launch_count)",
            "clocks": sampler.summary(), "roofline": roofline, "roofline_families": roofline_families,
            "kernels": kernels, "slowest_sparse_convs": conv_launches, "profiled_step_ms": round(step_ms_prof, 3),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        out_stream.write(json.dumps(line) + "\n")
        out_stream.flush()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
