"""CPU stand-in for spconv.pytorch backed by oracle/sp.py (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
import numpy as np
import torch
import torch.nn as nn

from oracle import me, sp

from . import utils  # noqa: F401


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, indice_dict=None, **kw):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], batch_size
        self.indice_dict = {} if indice_dict is None else indice_dict

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size, self.indice_dict)

    def dense(self):
        return sp.dense(self.features, self.indices.numpy(), self.spatial_shape, self.batch_size)


class SparseModule(nn.Module):
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args):
        super().__init__()
        for i, m in enumerate(args):
            self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


def _t3(v):
    return tuple(int(x) for x in v) if isinstance(v, (list, tuple)) else (int(v),) * 3


class SparseConvolution(SparseModule):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, subm=False, inverse=False,
                 indice_key=None, **kw):
        super().__init__()
        self.kernel_size, self.stride, self.padding = _t3(kernel_size), _t3(stride), _t3(padding)
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.randn(out_channels, *self.kernel_size, in_channels) * 0.05)
        assert not bias

    def forward(self, x):
        ind = x.indices.numpy()
        d = x.indice_dict.get(self.indice_key)
        if self.inverse:
            maps, out_ind, out_shape = me.transpose_map(d["maps"]), d["in_ind"], d["in_shape"]
        elif self.subm:
            if d is None:
                d = {"maps": sp.subm_maps(ind, self.kernel_size), "in_ind": ind, "in_shape": x.spatial_shape}
                x.indice_dict[self.indice_key] = d
            assert d["in_ind"] is ind or np.array_equal(d["in_ind"], ind)
            maps, out_ind, out_shape = d["maps"], ind, x.spatial_shape
        else:
            if d is None:
                oind, maps, oshape = sp.sparse_conv_indices(ind, x.spatial_shape, self.kernel_size, self.stride, self.padding)
                d = {"maps": maps, "in_ind": ind, "in_shape": x.spatial_shape, "out_ind": oind, "out_shape": oshape}
                x.indice_dict[self.indice_key] = d
            maps, out_ind, out_shape = d["maps"], d["out_ind"], d["out_shape"]
        f = sp.conv(x.features, (self.weight if (torch.is_grad_enabled() and self.weight.requires_grad) else self.weight.detach()), maps, len(out_ind))
        return SparseConvTensor(f, torch.from_numpy(np.ascontiguousarray(out_ind)), out_shape, x.batch_size, x.indice_dict)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, subm=True, indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias, indice_key=indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, **kw):
        super().__init__(in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)


class _ConvNS:
    SparseConvolution = SparseConvolution


conv = _ConvNS
