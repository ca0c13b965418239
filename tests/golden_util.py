"""helpers to read tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own code)."""
import json
import os

import numpy as np

from insmos_b200 import synth_weights
from insmos_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name, with_points=True):
    z = np.load(os.path.join(GOLDEN_DIR, "insmos_%s.npz" % name), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    shapes = {k: tuple(v) for k, v in meta["shapes"].items()}
    bn = {k[3:]: z[k] for k in z.files if k.startswith("bn:")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out:")}
    sd = synth_weights.fill_state_dict(shapes, bn_stats=bn)
    import torch
    sd["model.unet.center_head.conv_cls.bias"] = torch.full((3,), float(meta["cls_bias"]))
    pts = None
    if with_points:
        pts = synth.make_sequence(**meta["synth"])
        assert int(np.abs(pts).sum() * 1000) % (1 << 31) == meta["points_sha"], "synthetic generator drifted from the fixture"
    return meta, shapes, sd, pts, out


def triple_digest(k, i, o):
    """order-independent 64-bit digest of kernel-map triples (k, in_row, out_row): (sum, xor) over the triples of a mixed
    64-bit word.  Equal multisets of triples <=> equal digests (up to 2^-64 collisions); no sort needed on either side."""
    with np.errstate(over="ignore"):
        k, i, o = (np.asarray(a).astype(np.uint64) for a in (k, i, o))
        h = k * np.uint64(0x9E3779B97F4A7C15) + i * np.uint64(0xC2B2AE3D27D4EB4F) + o * np.uint64(0x165667B19E3779F9)
        h ^= h >> np.uint64(31)
        h *= np.uint64(0xD6E8FEB86659FD93)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
        return int(h.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(h)) if len(h) else 0
