"""Where is the host floor of the step?  Same model, same launch sequence, clouds of 1/1 ... 1/16 of the C2 rays.

The number of launches and the Python work per forward barely depend on the cloud size, the kernel time does.  The step
time extrapolated to an empty cloud is therefore the host's time per forward (interpreter + launch calls + the latency of
the ~11 blocking size reads); the distance of the full-size step from it says how much of the step the kernels own.
Run on the GPU box:  python tools/ab_host_floor.py
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--regions", type=int, default=3)
    ap.add_argument("--azimuths", default="1875,940,470,235,118")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from insmos_b200 import _lib
    from insmos_b200.engine import ForwardPool
    net = bench.build_model(device)
    main_stream = torch.cuda.current_stream(device)
    n_clouds = 4
    rows = []

    def timed(fn):
        import gc
        gc.collect()
        gc.disable()
        try:
            fn(6)
            torch.cuda.synchronize()
            ms = []
            l0 = _lib.launch_count()
            c0, w0 = time.process_time(), time.perf_counter()
            for _ in range(args.regions):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn(args.steps)
                e.record()
                torch.cuda.synchronize()
                ms.append(s.elapsed_time(e) / args.steps)
            cpu_s, wall_s = time.process_time() - c0, time.perf_counter() - w0
            launches = (_lib.launch_count() - l0) / float(args.regions * args.steps)
            return float(np.median(ms)), float(min(ms)), cpu_s / max(wall_s, 1e-9), launches
        finally:
            gc.enable()

    with torch.no_grad():
        for n_azim in [int(v) for v in args.azimuths.split(",")]:
            dev = [torch.from_numpy(c).to(device) for c in bench.make_clouds(0, n_clouds, n_azim=n_azim)]
            pts = int(np.mean([d.shape[0] for d in dev]))

            def run_inline(k):
                out = None
                for i in range(k):
                    out = bench.step(net, dev[i % n_clouds])
                return out

            m0, lo0, busy0, launches = timed(run_inline)
            # kernel time of one forward: CUDA events around every C-ABI call (includes nothing of the host between calls)
            torch.cuda.synchronize()
            _lib.profile_start()
            bench.step(net, dev[0])
            kern_ms = sum(t for _, t, _ in _lib.profile_stop())
            row = {"n_azim": n_azim, "points": pts, "launches_per_step": round(launches, 1), "kernel_ms_sum_cloud0": round(kern_ms, 3),
                   "inline_ms": round(m0, 4), "inline_min": round(lo0, 4), "inline_busy_cores": round(busy0, 2)}
            for w in (2, 3):
                pool = ForwardPool(net, workers=w, n_past=bench.N_SCANS)

                def run_pool(k):
                    jobs, out = [], None
                    for i in range(k):
                        jobs.append(pool.submit_points(dev[i % n_clouds]))
                        if len(jobs) > w:
                            out = jobs.pop(0).wait(main_stream)
                    while jobs:
                        out = jobs.pop(0).wait(main_stream)
                    return out

                m, lo, busy, _ = timed(run_pool)
                pool.close()
                row.update({"pool%d_ms" % w: round(m, 4), "pool%d_min" % w: round(lo, 4), "pool%d_busy_cores" % w: round(busy, 2)})
            rows.append(row)
            print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
