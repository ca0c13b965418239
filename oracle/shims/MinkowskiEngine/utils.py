import math

import torch

from oracle import me


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    return me.sparse_collate(coords, feats)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    return me.sparse_collate(coords, [torch.zeros(len(c), 1) for c in coords])[0]


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    if tensor.dim() == 2:
        fan_in, fan_out = tensor.size(0), tensor.size(1)
    else:
        fan_in, fan_out = tensor.size(1) * tensor.size(0), tensor.size(2) * tensor.size(0)
    fan = fan_in if mode == "fan_in" else fan_out
    with torch.no_grad():
        return tensor.normal_(0, torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan))
