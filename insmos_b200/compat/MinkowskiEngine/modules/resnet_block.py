"""MinkowskiEngine.modules.resnet_block: BasicBlock / Bottleneck (resnet.py:96-126 builds them).

Same sub-module names as upstream (conv1, norm1, conv2, norm2[, conv3, norm3], downsample) so that
state_dict keys match.  In eval mode BatchNorm, ReLU and the residual add are fused into the
convolution epilogues; in training mode the plain op sequence runs.
"""
import torch.nn as nn

import MinkowskiEngine as ME


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = ME.MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation,
                                             dimension=dimension)
        self.norm1 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation,
                                             dimension=dimension)
        self.norm2 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x, rows=None):
        """rows (inference only): (first_row of conv1, first_row of conv2 / downsample / output, bound for a kernel map that only
        this block uses or None) -- device int32 tensors, dead-row elimination"""
        if self.training:
            out = self.norm1(self.conv1(x), relu=True)
            out = self.norm2(self.conv2(out))
            res = x if self.downsample is None else self.downsample(x)
            return self.relu(out._like(out.F + res.F))
        r1, r2, rbr = rows if rows is not None else (None, None, None)
        out = self.conv1(x, bn=self.norm1, relu=True, first_row=r1, rb_first_row=rbr)
        if self.downsample is None:
            res = x
        elif isinstance(self.downsample, nn.Sequential) and len(self.downsample) == 2 and \
                isinstance(self.downsample[1], ME.MinkowskiBatchNorm):
            res = self.downsample[0](x, bn=self.downsample[1], first_row=r2)
        else:
            res = self.downsample(x)
        return self.conv2(out, bn=self.norm2, residual=res.F, relu=True, first_row=r2, rb_first_row=rbr)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        assert dimension > 0
        self.conv1 = ME.MinkowskiConvolution(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=stride, dilation=dilation,
                                             dimension=dimension)
        self.norm2 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv3 = ME.MinkowskiConvolution(planes, planes * self.expansion, kernel_size=1, dimension=dimension)
        self.norm3 = ME.MinkowskiBatchNorm(planes * self.expansion, momentum=bn_momentum)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        res = x if self.downsample is None else self.downsample(x)
        return self.relu(out._like(out.F + res.F))
