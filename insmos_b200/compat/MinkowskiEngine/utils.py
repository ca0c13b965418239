"""ME.utils: sparse_collate / batched_coordinates / kaiming_normal_ (motionnet.py:33, resnet.py:90)."""
import math

import torch


def batched_coordinates(coords, dtype=torch.int32, device=None):
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        ci = torch.floor(c).to(dtype) if c.is_floating_point() else c.to(dtype)
        out.append(torch.cat([torch.full((ci.shape[0], 1), b, dtype=dtype, device=ci.device), ci], dim=1))
    r = torch.cat(out, 0)
    return r if device is None else r.to(device)


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """floor the (float) coordinates to `dtype`, prepend the batch index, concatenate features."""
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    f = torch.cat([torch.as_tensor(x) for x in feats], 0)
    if labels is not None:
        return bcoords, f, torch.cat([torch.as_tensor(x) for x in labels], 0)
    return bcoords, f


def _fans(tensor):
    if tensor.dim() == 2:
        return tensor.size(0), tensor.size(1)
    rf = tensor.size(0)                       # [K, Cin, Cout]
    return tensor.size(1) * rf, tensor.size(2) * rf


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    fan_in, fan_out = _fans(tensor)
    fan = fan_in if mode == "fan_in" else fan_out
    gain = torch.nn.init.calculate_gain(nonlinearity, a)
    with torch.no_grad():
        return tensor.normal_(0, gain / math.sqrt(fan))
