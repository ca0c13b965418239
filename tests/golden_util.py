"""helpers to read tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own code)."""
import json
import os

import numpy as np

from insmos_b200 import synth_weights
from insmos_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, "insmos_%s.npz" % name), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    shapes = {k: tuple(v) for k, v in meta["shapes"].items()}
    bn = {k[3:]: z[k] for k in z.files if k.startswith("bn:")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out:")}
    sd = synth_weights.fill_state_dict(shapes, bn_stats=bn)
    import torch
    sd["model.unet.center_head.conv_cls.bias"] = torch.full((3,), float(meta["cls_bias"]))
    pts = synth.make_sequence(**meta["synth"])
    assert int(np.abs(pts).sum() * 1000) % (1 << 31) == meta["points_sha"], "synthetic generator drifted from the fixture"
    return meta, shapes, sd, pts, out
