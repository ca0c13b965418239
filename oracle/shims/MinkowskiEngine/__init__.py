"""CPU stand-in for MinkowskiEngine backed by oracle/me.py (TEST INFRASTRUCTURE; see oracle/__init__.py).
Lets the reference's own models/*.py run in the development container to generate golden fixtures."""
import math

import numpy as np
import torch
import torch.nn as nn

from oracle import me

from . import utils  # noqa: F401


def _tup(v, D):
    return tuple(int(x) for x in v) if isinstance(v, (list, tuple)) else (int(v),) * D


class CoordinateManager:
    def __init__(self, D):
        self.D, self.sets, self.maps = D, {}, {}

    def stride(self, key, stride):
        nk = tuple(k * s for k, s in zip(key, stride))
        if nk not in self.sets:
            self.sets[nk], _ = me.stride_coords(self.sets[key], list(nk))
        return nk

    def kernel_map(self, in_key, out_key, ksize):
        k = (in_key, out_key, ksize)
        if k not in self.maps:
            self.maps[k] = me.kernel_map(self.sets[in_key], self.sets[out_key], list(ksize), list(in_key))
        return self.maps[k]


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_manager=None, coordinate_map_key=None, **kw):
        if coordinate_manager is None:
            D = coordinates.shape[1] - 1
            coordinate_manager = CoordinateManager(D)
            coordinate_map_key = _tup(tensor_stride, D)
            coordinate_manager.sets[coordinate_map_key] = np.asarray(coordinates.cpu().numpy(), dtype=np.int32)
        self._F, self.coordinate_manager, self.coordinate_map_key = features, coordinate_manager, coordinate_map_key

    F = property(lambda s: s._F)
    features = F
    C = property(lambda s: torch.from_numpy(s.coordinate_manager.sets[s.coordinate_map_key]))
    coordinates = C
    tensor_stride = property(lambda s: list(s.coordinate_map_key))
    D = property(lambda s: s.coordinate_manager.D)

    def _like(self, f, key=None):
        return SparseTensor(f, coordinate_manager=self.coordinate_manager,
                            coordinate_map_key=self.coordinate_map_key if key is None else key)

    def slice(self, field):
        return TensorField(self._F[torch.from_numpy(field.inverse_mapping)], coordinates=field._coords.clone())


class TensorField:
    def __init__(self, features, coordinates, **kw):
        self._F, self._coords, self.inverse_mapping = features, coordinates, None

    F = property(lambda s: s._F)
    features = F
    C = property(lambda s: s._coords)
    coordinates = C

    def sparse(self):
        c = self._coords
        ci = (torch.floor(c) if c.is_floating_point() else c).to(torch.int32).numpy()
        uniq, inv = me.unique_first(ci)
        self.inverse_mapping = inv
        D = c.shape[1] - 1
        mgr = CoordinateManager(D)
        mgr.sets[(1,) * D] = uniq
        return SparseTensor(me.segment_mean(self._F, inv, len(uniq)), coordinate_manager=mgr, coordinate_map_key=(1,) * D)


def cat(*ts):
    return ts[0]._like(torch.cat([t.F for t in ts], dim=1))


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)

    def forward(self, x):
        return x._like(self.bn(x.F))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._like(torch.relu(x.F))


def _live(p):
    """parameters stay attached to autograd when a training golden is being made (tests/golden/make_golden_train.py)"""
    return p if (torch.is_grad_enabled() and p.requires_grad) else p.detach()


class _Conv(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None, **kw):
        super().__init__()
        D = dimension
        self.kernel_size, self.stride = _tup(kernel_size, D), _tup(stride, D)
        self.kernel_volume = int(math.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.randn(shape) * 0.1)
        self.bias = nn.Parameter(torch.zeros(1, out_channels)) if bias else None

    def forward(self, x):
        mgr, in_key = x.coordinate_manager, x.coordinate_map_key
        if self.kernel_volume == 1 and all(s == 1 for s in self.stride):
            return x._like(me.conv(x.F, _live(self.kernel), None, len(x.F), bias=None if self.bias is None else _live(self.bias)))
        if not self.transposed:
            out_key = in_key if all(s == 1 for s in self.stride) else mgr.stride(in_key, self.stride)
            maps = mgr.kernel_map(in_key, out_key, self.kernel_size)
        else:
            out_key = tuple(k // s for k, s in zip(in_key, self.stride))
            maps = me.transpose_map(mgr.kernel_map(out_key, in_key, self.kernel_size))
        f = me.conv(x.F, _live(self.kernel), maps, len(mgr.sets[out_key]), bias=None if self.bias is None else _live(self.bias))
        return x._like(f, out_key)


class MinkowskiConvolution(_Conv):
    pass


class MinkowskiConvolutionTranspose(_Conv):
    transposed = True


class _Stub(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


MinkowskiInstanceNorm = MinkowskiMaxPooling = MinkowskiDropout = MinkowskiGELU = MinkowskiGlobalMaxPooling = _Stub
MinkowskiLinear = MinkowskiSinusoidal = MinkowskiToSparseTensor = _Stub

from . import modules  # noqa: E402,F401
