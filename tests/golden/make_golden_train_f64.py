"""tests/golden/train_small_f64.npz: the training step of train_small.npz evaluated by oracle/train.py in FLOAT64.

Why: the gradients of this graph are ill-conditioned in fp32 at depth (train-mode BatchNorm backward; the saturated focal
loss of a randomly initialised head): the reference's own fp32 CPU run (train_small.npz) deviates from the fp64 result by
up to 1.8 % of a gradient's norm in the first encoder layers, and two fp32 CPU runs with different MKL thread counts differ
by up to 3 %.  The parity tests therefore measure every implementation -- the reference golden, the fp32 oracle, the CUDA
path -- against this fp64 result and require the CUDA path to be as close to it as the reference's own arithmetic is.
The fp64 graph is oracle/train.py (pinned to the reference graph by tests/test_oracle_train.py), fed the reference's
detections for the instance-fusion stage.  Runs anywhere (no /root/reference needed):  python tests/golden/make_golden_train_f64.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from oracle import train as otrain  # noqa: E402
from test_oracle_train import load_train_golden, projection  # noqa: E402

FULL_GRAD_MAX = 4096


def main():
    g, meta, sd, pts, labels, boxes = load_train_golden()
    ov = {"pred_boxes": torch.from_numpy(g["out:pred_boxes"]), "pred_labels": torch.from_numpy(g["out:pred_labels"])}
    r = otrain.train_step(sd, pts, labels, boxes, pred_override=ov, dtype=torch.float64)
    out = {"out:" + k: np.float64(r[k]) for k in ("loss", "loss_mos", "loss_motion_encoder", "rpn_loss_cls", "rpn_loss_loc")}
    out["out:point_seg_feature"] = r["logits"].numpy().astype(np.float32)
    out["out:current_motion_feature"] = r["motion"].numpy().astype(np.float32)
    for k, gr in r["grads"].items():
        gr = gr.double().numpy()
        out["gnorm:" + k] = np.float64(np.sqrt((gr ** 2).sum()))
        out["gproj:" + k] = np.asarray([(gr * projection(k, gr.shape, j)).sum() for j in range(int(meta["n_proj"]))])
        if gr.size <= FULL_GRAD_MAX:
            out["gfull:" + k] = gr
    path = os.path.join(HERE, "train_small_f64.npz")
    np.savez_compressed(path, meta=json.dumps({"source": "oracle/train.py float64", "of": "train_small.npz"}), **out)
    print("loss", r["loss"], "reference fp32", float(g["out:loss"]), "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
