"""A/B of how the host threads WAIT for the device, under a restricted CPU share (run on the GPU box, one GPU).

An 8-GPU job runs 8 processes on the box's 32 logical cores: 4 per rank.  Each rank has two forward threads; one holds the
GIL and queues kernels, the other typically sits in a blocking device->host size read.  With the CUDA default
(cudaDeviceScheduleAuto -> spin when cores > GPUs) the waiting thread burns a logical core, so a rank keeps ~2 cores busy
and eight ranks keep hyper-thread siblings busy -- the interpreter that feeds the launches slows down (8-GPU runs: 6.1-6.4 ms
per step against 5.5 ms alone).  cudaDeviceScheduleBlockingSync puts the waiting thread to sleep on an interrupt instead.

This script emulates one rank's CPU share with sched_setaffinity (all / 4 / 2 logical cores), switches the primary
context's schedule flag at run time and reports ms per step and the CPU seconds burnt per step (user+sys of the process).
One process, one model: every row is the same build on the same box.

    python tools/ab_host_wait.py [--steps 40] [--regions 3]
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

SCHED = {"auto": 0, "spin": 1, "yield": 2, "blocking": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--regions", type=int, default=3)
    ap.add_argument("--cpus", default="0,4,2", help="logical cores given to the process per configuration (0 = all)")
    ap.add_argument("--sched", default="auto,blocking,yield")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.zeros(1, device=device)
    rt = None
    for line in open("/proc/self/maps"):                       # the libcudart torch (and libinsmos_b200.so) already loaded
        if "libcudart.so" in line:
            rt = ctypes.CDLL(line.split()[-1])
            break
    if rt is None:
        rt = ctypes.CDLL("libcudart.so.12")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from insmos_b200.engine import ForwardPool
    from insmos_b200.pipeline import ScanPipeline
    n_clouds = 4
    host = [torch.from_numpy(c).pin_memory() for c in bench.make_clouds(0, n_clouds)]
    dev = [h.to(device) for h in host]
    net = bench.build_model(device)
    host_scans = []
    for h in host:
        a = h.numpy()
        stamps = np.unique(a[:, 4])
        parts = [a[a[:, 4] == t][:, :4] for t in stamps]
        offs = np.concatenate([[0], np.cumsum([len(q) for q in parts])]).astype(np.int64)
        host_scans.append((torch.from_numpy(np.ascontiguousarray(np.concatenate(parts, 0))).pin_memory(), offs))
    poses = [np.eye(4)] * bench.N_SCANS
    max_pts = max(int(h.shape[0]) for h in host) + 1024
    main_stream = torch.cuda.current_stream(device)
    all_cpus = sorted(os.sched_getaffinity(0))
    try:
        sib = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % all_cpus[0]).read().strip()
    except Exception:
        sib = "?"
    print(json.dumps({"logical_cpus": len(all_cpus), "siblings_of_first": sib}), flush=True)

    def timed(fn):
        import gc
        gc.collect()
        gc.disable()
        try:
            fn(6)
            torch.cuda.synchronize()
            ms = []
            c0, w0 = time.process_time(), time.perf_counter()
            for _ in range(args.regions):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn(args.steps)
                e.record()
                torch.cuda.synchronize()
                ms.append(s.elapsed_time(e) / args.steps)
            cpu_s, wall_s = time.process_time() - c0, time.perf_counter() - w0
            return float(np.median(ms)), float(min(ms)), float(max(ms)), cpu_s / max(wall_s, 1e-9)
        finally:
            gc.enable()

    rows = []
    with torch.no_grad():
        for ncpu in [int(v) for v in args.cpus.split(",")]:
            cpus = all_cpus if ncpu <= 0 else all_cpus[:ncpu]
            os.sched_setaffinity(0, cpus)                       # this thread; the worker threads created below inherit it
            for sched in args.sched.split(","):
                rc = rt.cudaSetDeviceFlags(ctypes.c_uint(SCHED[sched]))
                if rc != 0:
                    rt.cudaGetLastError()                       # do not leave a sticky-looking error for the launch checks
                got = ctypes.c_uint(0)
                rt.cudaGetDeviceFlags(ctypes.byref(got))
                pool = ForwardPool(net, workers=2, n_past=bench.N_SCANS)

                def run_pool(k):
                    jobs, out = [], None
                    for i in range(k):
                        jobs.append(pool.submit_points(dev[i % n_clouds]))
                        if len(jobs) > 2:
                            out = jobs.pop(0).wait(main_stream)
                    while jobs:
                        out = jobs.pop(0).wait(main_stream)
                    return out

                med, lo, hi, busy = timed(run_pool)
                pool.close()
                pipe = ScanPipeline(net, dt_pred=0.1, n_scans=bench.N_SCANS, max_points=max_pts, workers=2)

                def run_e2e(k):
                    prev, last = None, None
                    for i in range(k):
                        t = pipe.submit_packed(*host_scans[i % n_clouds], poses)
                        if prev is not None:
                            last = pipe.result(prev)
                        prev = t
                    return pipe.result(prev)

                med2, lo2, hi2, busy2 = timed(run_e2e)
                pipe.workers.close()
                del pipe
                rows.append({"cpus": len(cpus), "sched": sched, "set_rc": int(rc), "flags_now": int(got.value) & 7,
                             "value_ms": round(med, 4), "value_min": round(lo, 4), "value_max": round(hi, 4),
                             "value_busy_cores": round(busy, 2),
                             "e2e_ms": round(med2, 4), "e2e_min": round(lo2, 4), "e2e_max": round(hi2, 4),
                             "e2e_busy_cores": round(busy2, 2)})
                print(json.dumps(rows[-1]), flush=True)
    os.sched_setaffinity(0, all_cpus)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
