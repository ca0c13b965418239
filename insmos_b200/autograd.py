"""Autograd over the C-ABI kernels: what the training step (BASELINE config 5) needs beyond the forward.

The reference trains through the autograd Functions of MinkowskiEngine / spconv (MinkowskiConvolutionFunction,
SparseConvFunction / SparseInverseConvFunction / SubMConvFunction; call sites minkunet.py:139-181,
spconv_unet.py:284-406).  Here ONE Function covers all of them, because every convolution of the path runs on a tiled
rule book:
  forward   out   = conv(feat, W, rb)                     insmos_sparse_conv_fwd_*  (tensor cores, 3xTF32)
  dgrad     dfeat = conv(dout, W^T, rb^T)                 the SAME kernels over the transposed rule book
  wgrad     dW[k] = sum_{(i,o) in bucket k} feat[i]^T dout[o]      insmos_sparse_conv_wgrad (fp32 FFMA, deterministic)
The transposed rule book is never built by a generic transposition: each map of the path has a transposed map with the
same offset index that the forward builders already produce (strided <-> transposed/inverse convolution), and a
stride-1 map over an odd kernel is its own transpose with the offsets mirrored (pair (i,o,k) <=> (o,i,K-1-k)), so the
forward rule book is reused with the weights flipped along k.
"""
import torch

from . import ops


class _SparseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weight, rb, rb_t, flip, n_grad_cols):
        ctx.save_for_backward(feat, weight)
        ctx.rb, ctx.rb_t, ctx.flip, ctx.n_grad_cols = rb, rb_t, flip, n_grad_cols
        return ops.sparse_conv(feat.detach(), weight.detach(), rb)

    @staticmethod
    def backward(ctx, dout):
        feat, weight = ctx.saved_tensors
        dout = dout.contiguous()
        K, Cin, Cout = weight.shape
        dfeat = dw = None
        if ctx.needs_input_grad[0]:
            nc = Cin if ctx.n_grad_cols is None else ctx.n_grad_cols
            rbt = ctx.rb_t() if callable(ctx.rb_t) else ctx.rb_t
            wt = weight.detach()[:, :nc, :].transpose(1, 2)                 # [K, Cout, nc]
            if ctx.flip:
                wt = wt.flip(0)
            parts = []
            for c0 in range(0, nc, 128):                                    # the tcgen05 kernel takes <= 128 output channels
                parts.append(ops.sparse_conv(dout, wt[:, :, c0:c0 + 128].contiguous(), rbt))
            dfeat = parts[0] if len(parts) == 1 else torch.cat(parts, 1)
            if nc < Cin:
                dfeat = torch.nn.functional.pad(dfeat, (0, Cin - nc))
        if ctx.needs_input_grad[1]:
            dw = ops.sparse_conv_wgrad(feat.detach(), dout, ctx.rb, K, Cin, Cout)
        return dfeat, dw, None, None, None, None


def sparse_conv(feat, weight, rb, rb_t, flip=False, n_grad_cols=None):
    """differentiable sparse convolution.  rb_t: transposed rule book (or a callable returning it, evaluated in backward);
    flip: mirror the offsets (stride-1 odd-kernel maps: rb_t is rb itself); n_grad_cols: only the first columns of feat
    need a gradient (the instance bits / zero padding appended to the features do not)."""
    return _SparseConv.apply(feat, weight, rb, rb_t, flip, n_grad_cols)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        ctx.save_for_backward(idx)
        ctx.n = src.shape[0]
        return ops.gather_rows(src.detach(), idx)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        return ops.scatter_add_rows(dout, idx, ctx.n), None


def gather_rows(src, idx):
    """out[i] = idx[i] >= 0 ? src[idx[i]] : 0, differentiable in src (backward: insmos_scatter_add_rows)."""
    return _GatherRows.apply(src, idx)


class _Linear(torch.autograd.Function):
    """out = x @ W (insmos_linear_fwd); backward: two plain GEMMs (cuBLAS through torch: dX = dY W^T, dW = X^T dY)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return ops.linear(x.detach(), weight.detach())

    @staticmethod
    def backward(ctx, dout):
        x, weight = ctx.saved_tensors
        dx = dout @ weight.t() if ctx.needs_input_grad[0] else None
        dw = x.t() @ dout if ctx.needs_input_grad[1] else None
        return dx, dw


def linear(x, weight, bias=None):
    """x [n,Cin] @ weight [Cin,Cout] (+ bias), differentiable (kernel_size-1 MinkowskiConvolution, nn.Linear)."""
    out = _Linear.apply(x, weight)
    return out if bias is None else out + bias


class _BatchNormTrain(torch.autograd.Function):
    """train-mode BatchNorm over [n, C] rows with an optional fused ReLU, three launches: column sums (sum x, sum x^2 in fp64,
    one pass), per-channel constants + running-statistics update (insmos_bn_train_finalize), y = x*scale + shift (+ReLU).
    Backward, two launches: (sum g, sum g*xhat) with the ReLU gate applied on the fly, then dx / dgamma / dbeta."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, relu, momentum, running_mean, running_var):
        x = x.contiguous()
        n = x.shape[0]
        s, ss = ops.column_moments(x, 3)
        consts = ops.bn_train_finalize(s, ss, n, None if gamma is None else gamma.detach(), None if beta is None else beta.detach(), eps,
                                       momentum, running_mean, running_var)
        y = ops.affine_act(x, scale=consts[2], shift=consts[3], relu=relu)
        ctx.save_for_backward(x, y if relu else None, consts, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, consts, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        s0, s1 = ops.column_moments(dy, 2, b=x, gate=y, mean=consts[0], invstd=consts[1])
        dx, dgamma, dbeta = ops.bn_bwd_apply(dy, x, y, consts[0], consts[1], None if gamma is None else gamma.detach(), s0, s1)
        if gamma is None:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, None, None, None, None, None


def batch_norm_train(bn, x, relu=False):
    """nn.BatchNorm1d `bn` in training mode applied to x [n, C] (+ ReLU) on the library's kernels; updates running stats."""
    if x.shape[0] < 2:
        raise ValueError("batch_norm_train: more than one row per channel required (as nn.BatchNorm1d in training mode)")
    track = bn.track_running_stats and bn.running_mean is not None
    if track and bn.momentum is None:
        raise NotImplementedError("cumulative-average BatchNorm (momentum=None) is not on the InsMOS path")
    y = _BatchNormTrain.apply(x, bn.weight, bn.bias, bn.eps, relu, bn.momentum if track else 0.0,
                              bn.running_mean if track else None, bn.running_var if track else None)
    if track:
        with torch.no_grad():
            bn.num_batches_tracked += 1
        # the finalize kernel wrote the running statistics in place: bump their version counters (the folded eval-mode
        # constants are cached on them)
        torch.autograd.graph.increment_version(bn.running_mean)
        torch.autograd.graph.increment_version(bn.running_var)
    return y


def needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)
