import math

import torch
import torch.nn as nn

from insmos_b200 import autograd as _ag
from insmos_b200 import ops

from .core import IndiceData, SparseConvTensor
from .modules import SparseModule


def _tup3(v):
    return tuple(int(x) for x in v) if isinstance(v, (list, tuple)) else (int(v),) * 3


def fold_bn(bn):
    """eval-mode BatchNorm1d as (scale, shift), cached on the module until a parameter changes."""
    ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.weight.data_ptr(), bn.running_mean.data_ptr())
    cache = getattr(bn, "_insmos_folded", None)
    if cache is None or cache[0] != ver:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
        cache = (ver, scale.contiguous(), shift.contiguous())
        bn._insmos_folded = cache
    return cache[1], cache[2]


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, **kwargs):
        super().__init__()
        assert ndim == 3 and groups == 1
        self.ndim, self.in_channels, self.out_channels = ndim, in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _tup3(kernel_size), _tup3(stride), _tup3(padding)
        if _tup3(dilation) != (1, 1, 1) or transposed:
            raise NotImplementedError("dilation / transposed SparseConvolution are not on the InsMOS path")
        self.subm, self.inverse, self.indice_key = subm, inverse, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self._wk = None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            fan_in = self.in_channels * math.prod(self.kernel_size)
            nn.init.kaiming_uniform_(self.weight.view(self.out_channels, -1), a=math.sqrt(5))
            if self.bias is not None:
                b = 1 / math.sqrt(fan_in)
                self.bias.uniform_(-b, b)

    def kernel_major_weight(self, cin=None):
        """[K, Cin, Cout] view of the spconv-layout weight, cached until the parameter changes.  cin > in_channels gives
        the same weight with zero rows appended: the fused graph pads odd channel counts (16+3 instance bits, 7 point
        features) to a multiple of 8 so that the vectorised gather / tensor-core paths apply; zero inputs times zero
        weights add exact zeros."""
        ver = (self.weight._version, self.weight.data_ptr())
        if self._wk is None or self._wk[0] != ver:
            with torch.no_grad():
                wk = self.weight.reshape(self.out_channels, -1, self.in_channels).permute(1, 2, 0).contiguous()
            self._wk = (ver, wk, {})
        if cin is None or cin == self.in_channels:
            return self._wk[1]
        pad = self._wk[2].get(cin)
        if pad is None:
            wk = self._wk[1]
            with torch.no_grad():
                pad = torch.cat([wk, wk.new_zeros((wk.shape[0], cin - self.in_channels, wk.shape[2]))], 1).contiguous()
            self._wk[2][cin] = pad
        return pad

    def prefetch_indices(self, in_set, in_shape, indice_dict):
        """queue the output-coordinate kernel of a strided SparseConv3d ahead of its forward (row count read lazily,
        see ops.CoordSet); forward() then finds the indice data under indice_key.  Returns the IndiceData."""
        assert not self.subm and not self.inverse and self.indice_key is not None
        out_shape = [(i + 2 * p - k) // s + 1 for i, k, s, p in zip(in_shape, self.kernel_size, self.stride, self.padding)]
        out_set = ops.spconv_out_coords(in_set, self.kernel_size, self.stride, self.padding, out_shape, lazy=True)
        data = IndiceData(in_set, out_set, self.kernel_size, self.stride, self.padding, in_shape, out_shape, False)
        indice_dict[self.indice_key] = data
        return data

    def forward(self, x, bn=None, relu=False, residual=None, algo=0):
        assert isinstance(x, SparseConvTensor)
        scale = shift = None
        if bn is not None:
            scale, shift = fold_bn(bn)
        data = x.find_indice_pair(self.indice_key)
        if self.inverse:
            if data is None:
                raise RuntimeError("SparseInverseConv3d: indice_key %r not found" % (self.indice_key,))
            rb = data.inverse_rulebook()
            out_set, out_shape = data.in_set, data.in_shape
        elif self.subm:
            if data is None or data.in_set is not x.coordset:
                data = IndiceData(x.coordset, x.coordset, self.kernel_size, (1, 1, 1), self.padding, x.spatial_shape,
                                  x.spatial_shape, True)
                if self.indice_key is not None:
                    x.indice_dict[self.indice_key] = data
            rb = data.forward_rulebook()
            out_set, out_shape = data.out_set, x.spatial_shape
        else:
            if data is None or data.in_set is not x.coordset:
                out_shape = [(i + 2 * p - k) // s + 1 for i, k, s, p in
                             zip(x.spatial_shape, self.kernel_size, self.stride, self.padding)]
                out_set = ops.spconv_out_coords(x.coordset, self.kernel_size, self.stride, self.padding, out_shape)
                data = IndiceData(x.coordset, out_set, self.kernel_size, self.stride, self.padding, x.spatial_shape,
                                  out_shape, False)
                if self.indice_key is not None:
                    x.indice_dict[self.indice_key] = data
            rb = data.forward_rulebook()
            out_set, out_shape = data.out_set, data.out_shape
        cin = x.features.shape[1]
        if cin < self.in_channels:
            raise ValueError("SparseConvolution: %d input channels, layer expects %d" % (cin, self.in_channels))
        if bn is None and not relu and residual is None and _ag.needs_grad(x.features, self.weight, self.bias):
            # training step: differentiable kernel-major weight, transposed map by kind
            wk = self.weight.reshape(self.out_channels, -1, self.in_channels).permute(1, 2, 0)
            if cin > self.in_channels:
                wk = torch.cat([wk, wk.new_zeros((wk.shape[0], cin - self.in_channels, wk.shape[2]))], 1)
            if self.inverse:
                rb_t, flip = data.forward_rulebook, False
            elif self.subm:
                if any(k % 2 == 0 for k in self.kernel_size):
                    raise NotImplementedError("training through a submanifold convolution with an even kernel")
                rb_t, flip = rb, True
            else:
                rb_t, flip = data.inverse_rulebook, False
            f = _ag.sparse_conv(x.features, wk.contiguous(), rb, rb_t, flip, getattr(x, "grad_cols", None))
            if self.bias is not None:
                f = f + self.bias
            return SparseConvTensor(f, out_set.coords, out_shape, x.batch_size, indice_dict=x.indice_dict,
                                    benchmark=x.benchmark, coordset=out_set)
        f = ops.sparse_conv(x.features, self.kernel_major_weight(cin), rb, scale=scale, shift=shift, bias=self.bias,
                            residual=residual, relu=relu, algo=algo)
        return SparseConvTensor(f, out_set.coords, out_shape, x.batch_size, indice_dict=x.indice_dict,
                                benchmark=x.benchmark, coordset=out_set)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, subm=True,
                         indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key=None, bias=True, algo=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)
