"""A/B of the interpreter's GIL switch interval with several forwards in flight (run on the GPU box).

With two host threads queuing forwards under one GIL, a thread that returns from a blocking device->host size read has to
wait for the GIL: CPython asks the holder to drop it only after `sys.getswitchinterval()` (default 5 ms -- a whole forward),
so in practice the waiting thread's stream sits idle until the other thread reaches ITS next blocking read.  A short
interval hands the interpreter over within tens of microseconds.  One process, one model; the interval is changed at run
time, so every row of the table is the same build on the same box.

    python tools/ab_switch_interval.py [--steps 40] [--regions 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--regions", type=int, default=5)
    ap.add_argument("--intervals-us", default="5000,1000,300,100,30,10")
    ap.add_argument("--workers", default="2,3")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
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
    pipe = ScanPipeline(net, dt_pred=0.1, n_scans=bench.N_SCANS, max_points=max_pts, workers=2)
    main_stream = torch.cuda.current_stream(device)

    def run_pool(pool, nworkers, k):
        jobs, out = [], None
        for i in range(k):
            jobs.append(pool.submit_points(dev[i % n_clouds]))
            if len(jobs) > nworkers:
                out = jobs.pop(0).wait(main_stream)
        while jobs:
            out = jobs.pop(0).wait(main_stream)
        return out

    def run_e2e(k):
        prev, last = None, None
        for i in range(k):
            t = pipe.submit_packed(*host_scans[i % n_clouds], poses)
            if prev is not None:
                last = pipe.result(prev)
            prev = t
        return pipe.result(prev)

    def timed(fn):
        import gc
        gc.collect()
        gc.disable()
        try:
            fn(6)
            ms = []
            for _ in range(args.regions):
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn(args.steps)
                e.record()
                torch.cuda.synchronize()
                ms.append(s.elapsed_time(e) / args.steps)
            return float(np.median(ms)), float(min(ms)), float(max(ms))
        finally:
            gc.enable()

    default_interval = sys.getswitchinterval()
    rows = []
    with torch.no_grad():
        for w in [int(v) for v in args.workers.split(",")]:
            pool = ForwardPool(net, workers=w, n_past=bench.N_SCANS)
            for us in [float(v) for v in args.intervals_us.split(",")]:
                sys.setswitchinterval(us * 1e-6)
                med, lo, hi = timed(lambda k: run_pool(pool, w, k))
                rows.append({"path": "value (device-resident)", "workers": w, "switch_us": us, "ms_per_step": round(med, 4),
                             "min": round(lo, 4), "max": round(hi, 4), "scans_per_s": round(1000.0 / med, 1)})
                print(json.dumps(rows[-1]), flush=True)
            pool.close()
        for us in [float(v) for v in args.intervals_us.split(",")]:
            sys.setswitchinterval(us * 1e-6)
            med, lo, hi = timed(run_e2e)
            rows.append({"path": "e2e (ScanPipeline, host buffers)", "workers": 2, "switch_us": us, "ms_per_step": round(med, 4),
                         "min": round(lo, 4), "max": round(hi, 4), "scans_per_s": round(1000.0 / med, 1)})
            print(json.dumps(rows[-1]), flush=True)
    sys.setswitchinterval(default_interval)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
