"""4D MotionNet: quantise+unique (a1,a2) -> CustomMinkUNet (a3,a4) -> slice + t==0 select (a5).

Mirrors models/backbones_3d/motionnet.py:12-50 and the MinkUNet14 variant of
models/MinkowskiEngine/{minkunet.py:34-187, customminkunet.py:10-12, resnet.py:43-55,87-126}
(attribute names = state_dict keys).  Encoder 3 levels + decoder 3 levels, deepest level removed.
"""
import torch
import torch.nn as nn

import MinkowskiEngine as ME
from MinkowskiEngine.modules.resnet_block import BasicBlock

from insmos_b200 import autograd, ops


class CustomMinkUNet(nn.Module):
    BLOCK = BasicBlock
    PLANES = (8, 16, 32, 64, 64, 32, 16, 8)
    LAYERS = (1, 1, 1, 1, 1, 1, 1, 1)
    INIT_DIM = 8

    def __init__(self, in_channels, out_channels, D=4):
        super().__init__()
        self.D = D
        P, L, B = self.PLANES, self.LAYERS, self.BLOCK
        space_time = lambda m, n: [m, m, m, n] if D == 4 else [m] * D          # noqa: E731
        down = dict(kernel_size=space_time(2, 1), stride=space_time(2, 1), dimension=D)
        self.inplanes = self.INIT_DIM
        self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=space_time(5, 1), dimension=D)
        self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
        self.conv1p1s2 = ME.MinkowskiConvolution(self.inplanes, self.inplanes, **down)
        self.bn1 = ME.MinkowskiBatchNorm(self.inplanes)
        self.block1 = self._make_layer(B, P[0], L[0])
        self.conv2p2s2 = ME.MinkowskiConvolution(self.inplanes, self.inplanes, **down)
        self.bn2 = ME.MinkowskiBatchNorm(self.inplanes)
        self.block2 = self._make_layer(B, P[1], L[1])
        self.conv3p4s2 = ME.MinkowskiConvolution(self.inplanes, self.inplanes, **down)
        self.bn3 = ME.MinkowskiBatchNorm(self.inplanes)
        self.block3 = self._make_layer(B, P[2], L[2])
        # (the deepest level of the original MinkUNet is not built: minkunet.py:93-95)
        self.convtr5p8s2 = ME.MinkowskiConvolutionTranspose(self.inplanes, P[5], **down)
        self.bntr5 = ME.MinkowskiBatchNorm(P[5])
        self.inplanes = P[5] + P[1] * B.expansion
        self.block6 = self._make_layer(B, P[5], L[5])
        self.convtr6p4s2 = ME.MinkowskiConvolutionTranspose(self.inplanes, P[6], **down)
        self.bntr6 = ME.MinkowskiBatchNorm(P[6])
        self.inplanes = P[6] + P[0] * B.expansion
        self.block7 = self._make_layer(B, P[6], L[6])
        self.convtr7p2s2 = ME.MinkowskiConvolutionTranspose(self.inplanes, P[7], **down)
        self.bntr7 = ME.MinkowskiBatchNorm(P[7])
        self.inplanes = P[7] + self.INIT_DIM
        self.block8 = self._make_layer(B, P[7], L[7])
        self.final = ME.MinkowskiConvolution(P[7] * B.expansion, out_channels, kernel_size=1, bias=True, dimension=D)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self._init_weights()

    def _init_weights(self):                                   # resnet.py:87-94
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):          # resnet.py:96-126
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride,
                                        dimension=self.D),
                ME.MinkowskiBatchNorm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample, dimension=self.D)]
        self.inplanes = planes * block.expansion
        layers += [block(self.inplanes, planes, stride=1, dilation=dilation, dimension=self.D) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _cbr(self, conv, bn, x, first_row=None):
        """conv -> BatchNorm -> ReLU; one fused kernel in eval mode."""
        if self.training:
            return bn(conv(x), relu=True)
        return conv(x, bn=bn, relu=True, first_row=first_row)

    @staticmethod
    def _block(seq, x, rows):
        """a _make_layer Sequential; with one BasicBlock (every level of this network) the dead-row bounds are passed down"""
        if rows is not None and len(seq) == 1 and isinstance(seq[0], BasicBlock):
            return seq[0](x, rows=rows)
        return seq(x)

    def forward(self, x):                                       # minkunet.py:139-181
        # the strided coordinate maps are requested one level ahead of their first use, so that the host read of
        # each map's row count happens while the GPU is busy with the previous level (no drained queue)
        mgr, s = x.coordinate_manager, self.conv1p1s2.stride
        k2 = mgr.stride(x.coordinate_map_key, s, lazy=True)
        p1 = self._cbr(self.conv0p1s1, self.bn0, x)
        k4 = mgr.stride(k2, s, lazy=True)
        b1p2 = self.block1(self._cbr(self.conv1p1s2, self.bn1, p1))
        mgr.stride(k4, s, lazy=True)
        b2p4 = self.block2(self._cbr(self.conv2p2s2, self.bn2, b1p2))
        out = self.block3(self._cbr(self.conv3p4s2, self.bn3, b2p4))
        if self.training or not ops.USE_TPRUNE or self.D != 4 or not getattr(self, "newest_only", False):
            out = self.block6(ME.cat(self._cbr(self.convtr5p8s2, self.bntr5, out), b2p4))
            out = self.block7(ME.cat(self._cbr(self.convtr6p4s2, self.bntr6, out), b1p2))
            out = self.block8(ME.cat(self._cbr(self.convtr7p2s2, self.bntr7, out), p1))
            return self.final(out)
        # Dead-row elimination.  The caller (MotionNet) consumes only the rows of the newest scan, time
        # index 0 (motionnet.py:42-45).  A 3x3x3x3 convolution reaches one time index back, 1x1 / strided / transposed
        # convolutions none (their kernels have extent 1 in time), so walking the decoder backwards from "rows with t >= 0":
        #   final, block8.conv2 + downsample: t >= 0;  block8.conv1: t >= -1;  convtr7: t >= -2;
        #   block7.conv2 + downsample: t >= -2;  block7.conv1: t >= -3;  convtr6: t >= -4;
        #   block6.conv2 + downsample: t >= -4;  block6.conv1: t >= -5;  convtr5: t >= -6  (the encoder needs every row).
        # starts[j] = first row with t >= -j per level, computed and kept ON THE DEVICE; kernels skip the tiles below it.
        # Results on the needed rows are bit-identical (same per-row arithmetic); the other rows are left unwritten.
        f1 = ops.time_row_starts(mgr.sets[x.coordinate_map_key])
        f2 = ops.time_row_starts(mgr.sets[k2])
        f4 = ops.time_row_starts(mgr.sets[k4])
        r = lambda f, j: f[j:j + 1]                                         # noqa: E731
        out = self._block(self.block6, ME.cat(self._cbr(self.convtr5p8s2, self.bntr5, out, r(f4, 6)), b2p4), (r(f4, 5), r(f4, 4), None))
        out = self._block(self.block7, ME.cat(self._cbr(self.convtr6p4s2, self.bntr6, out, r(f2, 4)), b1p2), (r(f2, 3), r(f2, 2), None))
        # the 3x3x3x3 map at tensor stride 1 is used by block8 only: it is BUILT for the rows t >= -1 only
        out = self._block(self.block8, ME.cat(self._cbr(self.convtr7p2s2, self.bntr7, out, r(f1, 2)), p1), (r(f1, 1), r(f1, 0), r(f1, 1)))
        return self.final(out, first_row=r(f1, 0))


class MotionNet(nn.Module):
    def __init__(self, dt_prediction, voxel_size, out_channels):
        super().__init__()
        self.dt_prediction = dt_prediction
        ds = voxel_size[0]
        self.quantization = torch.Tensor([ds, ds, ds, self.dt_prediction])     # plain attribute, not a buffer
        self.out_channels = out_channels
        self.MinkUNet = CustomMinkUNet(in_channels=1, out_channels=out_channels, D=4)

    def forward(self, batch_dict):
        pts = batch_dict["past_point_clouds"]
        if not pts.is_cuda:
            raise RuntimeError("insmos_b200: past_point_clouds must be a CUDA tensor (no CPU fallback)")
        pts = pts.float().contiguous()
        # a1+a2: fp32 true division, floor, hashed unique in first-occurrence order; also yields the
        # indices of the current-scan points (t/dt == 0), motionnet.py:22-36,42
        voxels, inverse, cur_index = ops.voxelize4d(pts, [float(q) for q in self.quantization])
        mgr = ME.CoordinateManager(4)
        key = (1, 1, 1, 1)
        mgr.sets[key] = voxels
        # every point carries the feature 0.5; the unweighted per-voxel average of 0.5s is 0.5 exactly
        feats = torch.full((voxels.n, 1), 0.5, dtype=torch.float32, device=pts.device)
        self.MinkUNet.newest_only = True              # only the t == 0 rows of its output are read below
        pred = self.MinkUNet(ME.SparseTensor(feats, coordinate_manager=mgr, coordinate_map_key=key))
        # a5: slice back to points, keep the current scan, hstack(x,y,z,intensity, motion logits)
        if autograd.needs_grad(pred.F):
            # training: the motion logits stay attached to the graph (loss_motion_encoder, models.py:321-324); the point
            # columns carry no gradient.  current_point is detached below exactly where the reference's PointToVoxel cuts
            # the graph (voxel_generate.py:27: spconv's voxel generator is not an autograd op)
            cur_inv = inverse.index_select(0, cur_index.long()).contiguous()
            motion = autograd.gather_rows(pred.F[:, :self.out_channels].contiguous(), cur_inv)
            cur = torch.cat([pts.index_select(0, cur_index.long())[:, :4], motion.detach()], 1).contiguous()
            batch_dict["current_point"] = cur
            batch_dict["current_motion_feature"] = motion
        else:
            cur = ops.build_current_points(pts, cur_index, inverse, pred.F, self.out_channels)
            batch_dict["current_point"] = cur
            batch_dict["current_motion_feature"] = cur[:, 4:]
        batch_dict["_motion_stats"] = {"n_points": pts.shape[0], "n_voxels4d": voxels.n, "manager": mgr}
        return batch_dict
