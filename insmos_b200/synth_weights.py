"""Deterministic synthetic weights keyed by state_dict key name (no checkpoint exists offline: SURVEY F4).

The same function fills the reference's model (tests/golden/make_golden.py), the oracle graph and the
CUDA mirror, so only the calibrated BatchNorm statistics have to be stored in the golden fixture.
"""
import zlib

import numpy as np
import torch


def _rng(key):
    return np.random.default_rng(zlib.crc32(key.encode()))


def fill_state_dict(shapes, bn_stats=None):
    """shapes: {key: shape tuple}.  Returns {key: tensor}.  bn_stats: optional {key: array} overriding
    running_mean / running_var."""
    sd = {}
    for key, shape in shapes.items():
        shape = tuple(shape)
        r = _rng(key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.long)
        elif leaf == "running_mean":
            t = torch.zeros(shape)
        elif leaf == "running_var":
            t = torch.ones(shape)
        elif key.endswith("MOSLoss.loss.weight"):
            t = torch.tensor([0.0, 0.5, 0.5])
        elif leaf == "kernel":                                   # ME conv: [K,Cin,Cout] or [Cin,Cout]
            fan = shape[-2] * (max(shape[0] // 4, 1) if len(shape) == 3 else 1)
            t = torch.from_numpy(r.normal(0, 1.4 / np.sqrt(fan), shape).astype(np.float32))
        elif leaf == "weight" and len(shape) == 5:               # spconv: [Cout,kz,ky,kx,Cin]
            fan = shape[4] * max(int(np.prod(shape[1:4])) // 3, 1)
            t = torch.from_numpy(r.normal(0, 1.4 / np.sqrt(fan), shape).astype(np.float32))
        elif leaf == "weight" and len(shape) == 4:               # Conv2d / ConvTranspose2d
            if "conv_box" in key:
                t = torch.from_numpy(r.normal(0, 1e-3, shape).astype(np.float32))
            elif "conv_cls" in key:
                t = torch.from_numpy(r.normal(0, 0.05, shape).astype(np.float32))
            else:
                fan = (shape[0] if "deblocks" in key else shape[1]) * shape[2] * shape[3]
                t = torch.from_numpy(r.normal(0, 1.4 / np.sqrt(fan), shape).astype(np.float32))
        elif leaf == "weight" and len(shape) == 2:               # Linear
            t = torch.from_numpy(r.normal(0, 1.0 / np.sqrt(shape[1]), shape).astype(np.float32))
        elif leaf == "weight":                                   # BatchNorm gamma
            t = torch.from_numpy(r.uniform(0.5, 1.5, shape).astype(np.float32))
        elif leaf == "bias":
            t = torch.from_numpy(r.normal(0, 0.1, shape).astype(np.float32))
        else:
            raise KeyError("unexpected state_dict key %s %s" % (key, shape))
        sd[key] = t
    if bn_stats:
        for k, v in bn_stats.items():
            sd[k] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(sd[k].shape)
    return sd


def bn_stat_keys(shapes):
    return [k for k in shapes if k.endswith("running_mean") or k.endswith("running_var")]
