"""`MinkowskiEngine` operator surface re-provided over libinsmos_b200 (sm_100a CUDA, C ABI).

Not MinkowskiEngine: a from-scratch implementation of exactly the symbols the InsMOS forward path
touches (call sites models/backbones_3d/motionnet.py:21-50,
models/MinkowskiEngine/minkunet.py:52-181, resnet.py:87-126), with the same names, argument meaning,
parameter names/shapes (`.kernel [K,Cin,Cout]`, `.bias [1,Cout]`, `.bn`) and row-order convention
(first occurrence) so that checkpoints and the reference's own model code work unchanged.
CUDA only: tensors must live on the GPU; there is no CPU path.

Extension over the reference API (used by the fused inference graph in insmos_b200/net):
MinkowskiConvolution / ConvolutionTranspose.forward accept `bn=`, `relu=`, `residual=` to fuse an
eval-mode MinkowskiBatchNorm, a ReLU and a residual add into the convolution's epilogue.
"""
import math

import torch
import torch.nn as nn

from insmos_b200 import autograd as _ag
from insmos_b200 import ops

from . import utils  # noqa: F401  (ME.utils.sparse_collate, kaiming_normal_, batched_coordinates)

__version__ = "0.5.4+insmos_b200"


def _tup(v, D):
    if isinstance(v, (list, tuple)):
        assert len(v) == D
        return tuple(int(x) for x in v)
    return (int(v),) * D


class CoordinateManager:
    """coordinate sets per tensor stride + rule-book cache (ME caches kernel maps per
    (in key, out key, kernel, stride); every layer with the same geometry shares one map)."""

    def __init__(self, D):
        self.D = D
        self.sets = {}          # tensor_stride tuple -> ops.CoordSet
        self.rulebooks = {}
        self.parents = {}       # (fine key, coarse key) -> int32 [n_fine] row of each fine voxel's coarse cell

    def stride(self, key, stride, lazy=False):
        """coordinate map at tensor stride key*stride (made on first use).  lazy=True queues the kernel and leaves the
        row count on the device until the set is first used (see ops.CoordSet): callers that know the set will be
        needed later (insmos_b200.net.motion) request it one level ahead."""
        new_key = tuple(k * s for k, s in zip(key, stride))
        if new_key not in self.sets:
            cs, parent = ops.unique_coords(self.sets[key].coords, q=list(new_key), lazy=lazy)
            self.sets[new_key] = cs
            self.parents[(key, new_key)] = parent
        return new_key

    def rulebook(self, kind, in_key, out_key, ksize, stride, first_row=None):
        """first_row (device int32, optional): the caller only needs output rows >= *first_row; such a book is cached
        apart from the full one and only handed to callers that ask for it (dead-row elimination)"""
        k = (kind, in_key, out_key, ksize, stride) if first_row is None else (kind, in_key, out_key, ksize, stride, "rows_from")
        rb = self.rulebooks.get(k)
        if rb is None:
            if kind == "conv":
                spec = ops.spec_me_cube(list(ksize), list(in_key))
            else:
                spec = ops.spec_me_up(list(ksize), list(stride), list(out_key))
            # transposed map: the coarse set was made from the fine one, so every fine row already knows its parent
            parent = self.parents.get((out_key, in_key)) if kind == "up" else None
            # every set of this manager holds coordinates that are multiples of its tensor stride -> x-block probing
            rb = ops.build_rulebook(self.sets[out_key], self.sets[in_key], spec, parent=parent,
                                    xstep=in_key[0] if kind == "conv" else None, step=in_key if kind == "conv" else None,
                                    first_row=first_row if kind == "conv" else None)
            self.rulebooks[k] = rb
        return rb


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_manager=None, coordinate_map_key=None,
                 **kwargs):
        if coordinate_manager is None:
            if coordinates is None:
                raise ValueError("SparseTensor needs coordinates or a coordinate_manager + key")
            D = coordinates.shape[1] - 1
            coordinate_manager = CoordinateManager(D)
            key = _tup(tensor_stride, D)
            cs, inv = ops.unique_coords(coordinates.to(torch.int32).contiguous())
            if cs.n != coordinates.shape[0]:
                raise ValueError("SparseTensor: duplicate coordinates (use TensorField(...).sparse())")
            coordinate_manager.sets[key] = cs
            coordinate_map_key = key
        self._F = features
        self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key

    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.sets[self.coordinate_map_key].coords

    coordinates = C

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key)

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def device(self):
        return self._F.device

    def __len__(self):
        return self._F.shape[0]

    def _like(self, features, key=None):
        return SparseTensor(features, coordinate_manager=self.coordinate_manager,
                            coordinate_map_key=self.coordinate_map_key if key is None else key)

    def slice(self, field):
        """features of the voxel every point of `field` fell into (motionnet.py:38)."""
        gather = _ag.gather_rows if _ag.needs_grad(self._F) else ops.gather_rows
        return TensorField(gather(self._F, field.inverse_mapping), coordinates=field._coords.clone(), _skip_check=True)


class TensorField:
    """points with continuous coordinates [N,1+D] (batch first); .sparse() quantises with the
    UNWEIGHTED_AVERAGE mode (ME default) and remembers the point -> voxel map."""

    def __init__(self, features, coordinates, quantization_mode=None, _skip_check=False, **kwargs):
        if not _skip_check and not features.is_cuda:
            raise RuntimeError("insmos_b200 MinkowskiEngine: CUDA tensors required (no CPU fallback)")
        self._F = features
        self._coords = coordinates
        self.inverse_mapping = None

    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self._coords

    coordinates = C

    def sparse(self):
        c = self._coords
        ci = torch.floor(c).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
        cs, inv = ops.unique_coords(ci.contiguous())
        self.inverse_mapping = inv
        D = c.shape[1] - 1
        mgr = CoordinateManager(D)
        key = (1,) * D
        mgr.sets[key] = cs
        return SparseTensor(ops.segment_mean(self._F.float(), inv, cs.n), coordinate_manager=mgr, coordinate_map_key=key)


def cat(*tensors):
    out = tensors[0]
    f = out.F
    for t in tensors[1:]:
        assert t.coordinate_map_key == out.coordinate_map_key, "ME.cat: tensors must share a coordinate map"
        f = torch.cat([f, t.F], 1) if _ag.needs_grad(f, t.F) else ops.concat2(f, t.F)
    return out._like(f)


# --------------------------------------------------------------------------------------------------
class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)
        self._folded = None

    def folded(self):
        """eval-mode BatchNorm as per-channel (scale, shift); cached until a parameter changes."""
        bn = self.bn
        ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               bn.weight.data_ptr(), bn.running_mean.data_ptr())
        if self._folded is None or self._folded[0] != ver:
            with torch.no_grad():
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                shift = bn.bias - bn.running_mean * scale
            self._folded = (ver, scale.contiguous(), shift.contiguous())
        return self._folded[1], self._folded[2]

    def forward(self, x, relu=False):
        if self.training or not self.bn.track_running_stats:
            # batch statistics (training step): column-moment kernels + fused normalise(+ReLU), with autograd
            return x._like(_ag.batch_norm_train(self.bn, x.F, relu=relu))
        s, t = self.folded()
        return x._like(ops.affine_act(x.F, scale=s, shift=t, relu=relu))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x):
        f = x.F
        if _ag.needs_grad(f):
            return x._like(torch.relu(f))
        return x._like(ops.affine_act(f, relu=True, out=f if self.inplace and not f.requires_grad else None))


def _fold(bn):
    if bn is None:
        return None, None
    if bn.training:
        raise RuntimeError("fused bn= needs eval mode")
    return bn.folded()


class _ConvBase(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, dimension=None, **kwargs):
        super().__init__()
        assert dimension is not None and dimension > 0
        D = dimension
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, D
        self.kernel_size, self.stride, self.dilation = _tup(kernel_size, D), _tup(stride, D), _tup(dilation, D)
        if any(d != 1 for d in self.dilation):
            raise NotImplementedError("dilation != 1 is not on the InsMOS path")
        if expand_coordinates:
            raise NotImplementedError("expand_coordinates is not on the InsMOS path")
        self.kernel_volume = int(math.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            n = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def _forward_train(self, x):
        """differentiable path (autograd.sparse_conv): forward kernel, dgrad = forward kernel over the transposed map, wgrad kernel"""
        mgr, in_key = x.coordinate_manager, x.coordinate_map_key
        if self.kernel_volume == 1 and all(s == 1 for s in self.stride):
            return x._like(_ag.linear(x.F, self.kernel, self.bias))
        w = self.kernel if self.kernel.dim() == 3 else self.kernel.view(1, *self.kernel.shape)
        if not self.transposed:
            if all(s == 1 for s in self.stride):
                if any(k % 2 == 0 for k in self.kernel_size):
                    raise NotImplementedError("training through a stride-1 convolution with an even kernel")
                out_key = in_key
                rb = mgr.rulebook("conv", in_key, out_key, self.kernel_size, self.stride)
                rb_t, flip = rb, True                                      # own transpose with mirrored offsets
            else:
                out_key = mgr.stride(in_key, self.stride)
                rb = mgr.rulebook("conv", in_key, out_key, self.kernel_size, self.stride)
                ks, st = self.kernel_size, self.stride
                rb_t, flip = (lambda: mgr.rulebook("up", out_key, in_key, ks, st)), False
        else:
            out_key = tuple(k // s for k, s in zip(in_key, self.stride))
            if out_key not in mgr.sets or self.kernel_size != self.stride:
                raise NotImplementedError("transposed convolution off the InsMOS path")
            rb = mgr.rulebook("up", in_key, out_key, self.kernel_size, self.stride)
            ks, st = self.kernel_size, self.stride
            rb_t, flip = (lambda: mgr.rulebook("conv", out_key, in_key, ks, st)), False
        f = _ag.sparse_conv(x.F, w, rb, rb_t, flip)
        return x._like(f if self.bias is None else f + self.bias, out_key)

    def forward(self, x, bn=None, relu=False, residual=None, algo=0, first_row=None, rb_first_row=None):
        """first_row / rb_first_row (device int32 tensors, inference only): output rows below *first_row are not needed by the
        caller; rb_first_row is the smallest such bound over every layer that shares this layer's kernel map."""
        if bn is None and not relu and residual is None and _ag.needs_grad(x.F, self.kernel, self.bias):
            return self._forward_train(x)
        scale, shift = _fold(bn)
        bias = None if self.bias is None else self.bias.view(-1)
        mgr, in_key = x.coordinate_manager, x.coordinate_map_key
        if self.kernel_volume == 1 and all(s == 1 for s in self.stride):
            f = ops.linear(x.F, self.kernel, scale=scale, shift=shift, bias=bias, residual=residual, relu=relu, first_row=first_row)
            return x._like(f)
        if not self.transposed:
            out_key = in_key if all(s == 1 for s in self.stride) else mgr.stride(in_key, self.stride)
            rb = mgr.rulebook("conv", in_key, out_key, self.kernel_size, self.stride, first_row=rb_first_row)
        else:
            out_key = tuple(k // s for k, s in zip(in_key, self.stride))
            if out_key not in mgr.sets:
                raise RuntimeError("MinkowskiConvolutionTranspose: no coordinate map with tensor stride %s "
                                   "(generative transposed convolution is not on the InsMOS path)" % (out_key,))
            if self.kernel_size != self.stride:
                raise NotImplementedError("transposed convolution with kernel != stride is not on the InsMOS path")
            rb = mgr.rulebook("up", in_key, out_key, self.kernel_size, self.stride)
        w = self.kernel if self.kernel.dim() == 3 else self.kernel.view(1, *self.kernel.shape)
        f = ops.sparse_conv(x.F, w, rb, scale=scale, shift=shift, bias=bias, residual=residual, relu=relu, algo=algo,
                            first_row=first_row)
        return x._like(f, out_key)

    def extra_repr(self):
        return "in=%d, out=%d, kernel_size=%s, stride=%s" % (self.in_channels, self.out_channels,
                                                             list(self.kernel_size), list(self.stride))


class MinkowskiConvolution(_ConvBase):
    transposed = False


class MinkowskiConvolutionTranspose(_ConvBase):
    transposed = True


class _NotOnPath(nn.Module):
    """symbols the reference's files name at import / class-definition time but never execute on the
    InsMOS forward path (resnet.py:60-85,164-192)."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise NotImplementedError("%s is not on the InsMOS forward path" % type(self).__name__)


class MinkowskiInstanceNorm(_NotOnPath): pass
class MinkowskiMaxPooling(_NotOnPath): pass
class MinkowskiDropout(_NotOnPath): pass
class MinkowskiGELU(_NotOnPath): pass
class MinkowskiGlobalMaxPooling(_NotOnPath): pass
class MinkowskiLinear(_NotOnPath): pass
class MinkowskiSinusoidal(_NotOnPath): pass
class MinkowskiToSparseTensor(_NotOnPath): pass
class MinkowskiSyncBatchNorm(_NotOnPath): pass


from . import modules  # noqa: E402,F401
