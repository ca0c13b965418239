"""Dense BEV head (a9): HeightCompression -> BaseBEVBackbone -> CenterHead decode.

Mirrors models/backbones_2d/{height_compression.py:8-33, base_bev_backbone.py:9-115,
center_head.py:29-98,251-276}.  The dense 2D convolutions are library GEMM-shaped work (cuDNN via
torch, TF32 disabled for fp32 parity); in eval mode BatchNorm2d is folded into the convolution
weights and ReLU applied in place.  The per-cell decode + sigmoid + class max is one CUDA kernel.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from insmos_b200 import ops


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg["NUM_BEV_FEATURES"]

    def forward(self, batch_dict):
        dense = batch_dict["encoded_spconv_tensor"].dense()           # [1, C, D, H, W]
        n, c, d, h, w = dense.shape
        batch_dict["spatial_features"] = dense.view(n, c * d, h, w)
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        return batch_dict


def _fold_conv_bn(conv, bn, transposed=False):
    ver = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, conv.weight.data_ptr(), bn.running_mean.data_ptr())
    cache = getattr(conv, "_insmos_folded", None)
    if cache is None or cache[0] != ver:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            w = conv.weight * (scale.view(1, -1, 1, 1) if transposed else scale.view(-1, 1, 1, 1))
        cache = (ver, w.contiguous(), shift.contiguous())
        conv._insmos_folded = cache
    return cache[1], cache[2]


class BaseBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = model_cfg.get("LAYER_NUMS") or []
        layer_strides = model_cfg.get("LAYER_STRIDES") or []
        num_filters = model_cfg.get("NUM_FILTERS") or []
        upsample_strides = model_cfg.get("UPSAMPLE_STRIDES") or []
        num_upsample_filters = model_cfg.get("NUM_UPSAMPLE_FILTERS") or []
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        assert len(upsample_strides) == len(num_upsample_filters)
        bn2d = lambda c: nn.BatchNorm2d(c, eps=1e-3, momentum=0.01)                    # noqa: E731
        c_in = [input_channels, *num_filters[:-1]]
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        for i, n_layers in enumerate(layer_nums):
            seq = [nn.ZeroPad2d(1), nn.Conv2d(c_in[i], num_filters[i], 3, stride=layer_strides[i], padding=0, bias=False),
                   bn2d(num_filters[i]), nn.ReLU()]
            for _ in range(n_layers):
                seq += [nn.Conv2d(num_filters[i], num_filters[i], 3, padding=1, bias=False), bn2d(num_filters[i]), nn.ReLU()]
            self.blocks.append(nn.Sequential(*seq))
            if upsample_strides:
                s = upsample_strides[i]
                if s >= 1:
                    up = nn.ConvTranspose2d(num_filters[i], num_upsample_filters[i], s, stride=s, bias=False)
                else:
                    s = int(np.round(1 / s))
                    up = nn.Conv2d(num_filters[i], num_upsample_filters[i], s, stride=s, bias=False)
                self.deblocks.append(nn.Sequential(up, bn2d(num_upsample_filters[i]), nn.ReLU()))
        c_up = sum(num_upsample_filters)
        if len(upsample_strides) > len(layer_nums):
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(c_up, c_up, upsample_strides[-1], stride=upsample_strides[-1], bias=False),
                bn2d(c_up), nn.ReLU()))
        self.num_bev_features = c_up

    @staticmethod
    def _run(seq, x, training):
        if training:
            return seq(x)
        mods, i, pad = list(seq), 0, 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.ZeroPad2d):
                pad = m.padding[0]
                i += 1
                continue
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d):
                tr = isinstance(m, nn.ConvTranspose2d)
                w, b = _fold_conv_bn(m, mods[i + 1], transposed=tr)
                if tr:
                    x = F.conv_transpose2d(x, w, b, stride=m.stride)
                else:
                    x = F.conv2d(x, w, b, stride=m.stride, padding=(m.padding[0] + pad, m.padding[1] + pad))
                pad = 0
                i += 2
                if i < len(mods) and isinstance(mods[i], nn.ReLU):
                    x = torch.relu_(x)
                    i += 1
                continue
            x = m(x)
            i += 1
        return x

    def forward(self, data_dict):
        x0 = data_dict["current_bev"]
        x, ups = x0, []
        for i, blk in enumerate(self.blocks):
            x = self._run(blk, x, self.training)
            data_dict["spatial_features_%dx" % int(x0.shape[2] / x.shape[2])] = x
            ups.append(self._run(self.deblocks[i], x, self.training) if len(self.deblocks) > 0 else x)
        x = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        if len(self.deblocks) > len(self.blocks):
            x = self._run(self.deblocks[-1], x, self.training)
        data_dict["spatial_features_2d"] = x
        return data_dict


class CenterHead(nn.Module):
    """forward + box decoding only (target assignment and losses are training-only)."""

    def __init__(self, model_cfg, input_channels, num_class, class_names, grid_size, point_cloud_range,
                 predict_boxes_when_training=True):
        super().__init__()
        self.model_cfg, self.num_class, self.class_names = model_cfg, num_class, [class_names]
        self.target_cfg = model_cfg["TARGET_ASSIGNER_CONFIG"]
        self.grid_size, self.point_cloud_range = grid_size, point_cloud_range
        self.forward_ret_dict = {}
        self.conv_cls = nn.Conv2d(input_channels, num_class, kernel_size=1)
        self.conv_box = nn.Conv2d(input_channels, 8, kernel_size=1)
        nn.init.constant_(self.conv_cls.bias, -np.log((1 - 0.01) / 0.01))
        nn.init.normal_(self.conv_box.weight, mean=0, std=0.001)

    def forward(self, data_dict, Model_mode):
        if Model_mode == "train":
            raise NotImplementedError("CenterHead target assignment / losses are training-only (out of scope, SURVEY 8f N3)")
        x = data_dict["spatial_features_2d"]
        cls = self.conv_cls(x)                                         # [1, ncls, H, W]
        box = self.conv_box(x)                                         # [1, 8, H, W]
        if cls.shape[0] != 1:
            raise NotImplementedError("batch 1 per sample (models.py:313)")
        t = self.target_cfg
        boxes, scores, labels = ops.center_decode(cls[0], box[0], t["OUT_SIZE_FACTOR"], t["VOXEL_SIZE"][0],
                                                  t["VOXEL_SIZE"][1], self.point_cloud_range[0], self.point_cloud_range[1])
        data_dict["batch_cls_preds"] = cls[0].permute(1, 2, 0).reshape(1, -1, self.num_class)   # raw logits view
        data_dict["batch_box_preds"] = boxes.unsqueeze(0)
        data_dict["cls_preds_normalized"] = False
        data_dict["_decoded"] = (boxes, scores, labels)                 # sigmoid / class max already done on device
        return data_dict
