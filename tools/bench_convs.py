#!/usr/bin/env python3
"""Per-layer A/B of the narrow sparse-conv kernels on the C2 cloud's own maps (MotionNet levels ts1..ts8, 3^4 offsets):
every (Cin, Cout) the network runs at that level, for a list of kernel variants selected through environment switches.
Prints one JSON line per (level, layer, variant): ms, algorithmic GB/s, GFLOP/s.   python tools/bench_convs.py [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from insmos_b200 import ops, synth  # noqa: E402

VARIANTS = {
    "tc4 (mma.sync 3xTF32)": dict(algo=2, env={}),
    "fma default": dict(algo=5, env={}),
    "fma F2": dict(algo=5, env={"INSMOS_FMA_F2": "1"}),
    "fma PW32": dict(algo=5, env={"INSMOS_FMA_PW": "32"}),
    "fma PW16": dict(algo=5, env={"INSMOS_FMA_PW": "16"}),
    "fma PW8": dict(algo=5, env={"INSMOS_FMA_PW": "8"}),
    "fma R128": dict(algo=5, env={"INSMOS_FMA_R": "128"}),
    "fma R512": dict(algo=5, env={"INSMOS_FMA_R": "512"}),
    "fma warps4": dict(algo=5, env={"INSMOS_FMA_WARPS": "4"}),
    "fma warps4 R128": dict(algo=5, env={"INSMOS_FMA_WARPS": "4", "INSMOS_FMA_R": "128"}),
}
LAYERS = {1: [(8, 8), (16, 8)], 2: [(8, 8), (24, 16), (16, 16)], 4: [(8, 16), (16, 16), (48, 32), (32, 32)], 8: [(16, 32), (32, 32)]}
ENV_KEYS = ["INSMOS_FMA_F2", "INSMOS_FMA_PW", "INSMOS_FMA_R", "INSMOS_FMA_WARPS"]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    dev = torch.device("cuda:0")
    pts = torch.from_numpy(synth.make_sequence(seed=0, n_scans=10, n_elev=64, n_azim=1875)).to(dev)
    cs, _, _ = ops.voxelize4d(pts, [0.1, 0.1, 0.1, 0.1])
    sets = {1: cs}
    for ts in (2, 4, 8):
        sets[ts], _ = ops.unique_coords(sets[ts // 2].coords, q=[ts, ts, ts, 1])
    g = torch.Generator().manual_seed(0)
    for ts, layers in LAYERS.items():
        s = sets[ts]
        for TM in (128, 64):
            rb = ops.build_rulebook(s, s, ops.spec_me_cube([3, 3, 3, 3], [ts, ts, ts, 1]), TM=TM, xstep=ts)
            P = rb.num_pairs
            for Cin, Cout in layers:
                x = torch.randn((s.n, Cin), generator=g).to(dev)
                W = (torch.randn((81, Cin, Cout), generator=g) / np.sqrt(Cin * 8.0)).to(dev)
                ref = None
                for name, v in VARIANTS.items():
                    if only and not any(o in name for o in only):
                        continue
                    if TM == 64 and v["algo"] == 5 and name != "fma default":
                        continue
                    for k in ENV_KEYS:
                        os.environ.pop(k, None)
                    os.environ.update(v["env"])
                    try:
                        out = ops.sparse_conv(x, W, rb, algo=v["algo"])
                        torch.cuda.synchronize()
                        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        st.record()
                        for _ in range(reps):
                            ops.sparse_conv(x, W, rb, algo=v["algo"])
                        en.record()
                        torch.cuda.synchronize()
                        ms = st.elapsed_time(en) / reps
                    except RuntimeError as e:
                        print(json.dumps({"ts": ts, "TM": TM, "Cin": Cin, "Cout": Cout, "variant": name, "error": str(e)[:80]}), flush=True)
                        continue
                    if ref is None:
                        ref = out
                    err = float((out - ref).abs().max())
                    b = 4 * (s.n * Cin + s.n * Cout) + 8 * P + 4 * 81 * Cin * Cout
                    print(json.dumps({"ts": ts, "n": s.n, "TM": TM, "pairs": P, "Cin": Cin, "Cout": Cout, "variant": name, "ms": round(ms, 4),
                                      "alg_GBps": round(b / ms / 1e6, 1), "gflops": round(2 * P * Cin * Cout / ms / 1e6, 1),
                                      "max_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
