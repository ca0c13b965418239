"""Dense BEV head (a9): HeightCompression -> BaseBEVBackbone -> CenterHead decode.

Mirrors models/backbones_2d/{height_compression.py:8-33, base_bev_backbone.py:9-115,
center_head.py:29-98,251-276} (module tree / state_dict keys identical).  Inference path (eval mode):
the whole head runs channels-last on this repository's tensor-core kernels -- sparse->dense scatter straight
into NHWC, the 3x3 convs and the 2x2 transposed conv as 3xTF32 implicit GEMMs with BatchNorm folded into the
weights and ReLU in the epilogue (insmos_conv2d_nhwc_umma / _tcgen05), the two 1x1 heads as one fused linear kernel, and
decode + sigmoid + class max in one kernel.  Training mode falls back to the plain torch modules.
"""
import numpy as np
import torch
import torch.nn as nn

from insmos_b200 import ops


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg["NUM_BEV_FEATURES"]

    def forward(self, batch_dict):
        t = batch_dict["encoded_spconv_tensor"]
        if self.training:
            dense = t.dense()                                          # [1, C, D, H, W]
            n, c, d, h, w = dense.shape
            batch_dict["spatial_features"] = dense.view(n, c * d, h, w)
        else:
            D, H, W = t.spatial_shape
            batch_dict["spatial_features_nhwc"] = (ops.dense_scatter_nhwc(t.features, t.indices, D, H, W), H, W)
            batch_dict["spatial_features"] = None
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        return batch_dict


def _folded(conv, bn, transposed):
    """[taps, Cin, Cout] weight with the eval-mode BatchNorm scale folded in, and the shift as bias (cached)."""
    ver = (conv.weight._version, bn.weight._version, bn.bias._version, bn.running_mean._version,
           bn.running_var._version, conv.weight.data_ptr(), bn.running_mean.data_ptr())
    cache = getattr(conv, "_insmos_folded", None)
    if cache is None or cache[0] != ver:
        with torch.no_grad():
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            if transposed:                                              # [Cin, Cout, kh, kw] -> [kh*kw, Cin, Cout]
                w = conv.weight.permute(2, 3, 0, 1)
            else:                                                       # [Cout, Cin, kh, kw] -> [kh*kw, Cin, Cout]
                w = conv.weight.permute(2, 3, 1, 0)
            w = (w * scale.view(1, 1, 1, -1)).reshape(-1, w.shape[2], w.shape[3]).contiguous()
        cache = (ver, w, shift.contiguous())
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

    def _fast_supported(self):
        if len(self.blocks) != 1 or len(self.deblocks) != 1:
            return False
        mods = list(self.blocks[0])
        convs = [m for m in mods if isinstance(m, nn.Conv2d)]
        up = self.deblocks[0][0]
        return (all(c.kernel_size == (3, 3) and c.stride == (1, 1) and c.in_channels % 32 == 0 and c.out_channels % 128 == 0
                    for c in convs)
                and isinstance(up, nn.ConvTranspose2d) and up.kernel_size == (2, 2) and up.stride == (2, 2)
                and up.in_channels % 32 == 0 and up.out_channels % 128 == 0)

    def forward_nhwc(self, x, H, W):
        """x [H*W, C] channels-last -> ([2H*2W, C_up], 2H, 2W); the reference's layer sequence, fused."""
        mods = list(self.blocks[0])
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Conv2d):
                w, b = _folded(m, mods[i + 1], transposed=False)       # ZeroPad2d(1)+conv(p=0) == conv(p=1)
                x = ops.conv2d_nhwc(x, H, W, w, 0, bias=b, relu=True)
                i += 3
            else:
                i += 1
        up = self.deblocks[0]
        w, b = _folded(up[0], up[1], transposed=True)
        return ops.conv2d_nhwc(x, H, W, w, 2, bias=b, relu=True), 2 * H, 2 * W

    def forward(self, data_dict):
        if not self.training and data_dict.get("current_bev_nhwc") is not None and self._fast_supported():
            x, H, W = data_dict["current_bev_nhwc"]
            data_dict["spatial_features_2d_nhwc"] = self.forward_nhwc(x, H, W)
            data_dict["spatial_features_2d"] = None
            return data_dict
        x0 = data_dict.get("current_bev")
        if x0 is None:
            # eval mode produced only the channels-last map, but this configuration (strides != 1, several blocks,
            # channel counts off the tensor-core tiles) is not covered by the fused path: rebuild NCHW for the torch modules
            xn, H, W = data_dict["current_bev_nhwc"]
            x0 = xn.view(H, W, -1).permute(2, 0, 1).unsqueeze(0).contiguous()
        x, ups = x0, []
        for i, blk in enumerate(self.blocks):
            x = blk(x)
            data_dict["spatial_features_%dx" % int(x0.shape[2] / x.shape[2])] = x
            ups.append(self.deblocks[i](x) if len(self.deblocks) > 0 else x)
        x = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        if len(self.deblocks) > len(self.blocks):
            x = self.deblocks[-1](x)
        data_dict["spatial_features_2d"] = x
        return data_dict


def clip_sigmoid(x, eps=1e-4):
    """center_head.py:333-345 (the reference applies the sigmoid in place; the values are the same)"""
    return torch.clamp(torch.sigmoid(x), min=eps, max=1 - eps)


def gaussian_focal_loss(pred, gaussian_target, alpha=2.0, gamma=4.0):
    """center_head.py:590-611: element-wise focal loss against a Gaussian heat map (positives: target == 1)."""
    eps = 1e-12
    pos_weights = gaussian_target.eq(1)
    neg_weights = (1 - gaussian_target).pow(gamma)
    pos_loss = -(pred + eps).log() * (1 - pred).pow(alpha) * pos_weights
    neg_loss = -(1 - pred + eps).log() * pred.pow(alpha) * neg_weights
    return pos_loss + neg_loss


class CenterHead(nn.Module):
    """forward + box decoding (center_head.py:65-98,251-276); in 'train' mode also target assignment on device
    (insmos_center_targets replaces the Python loop of center_head.py:171-249) and the two losses (:279-331)."""

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
        self._heads = None

    def _fused_heads(self):
        """[Cin, ncls+8] weight and [ncls+8] bias of the two 1x1 heads, cached."""
        ver = (self.conv_cls.weight._version, self.conv_box.weight._version, self.conv_cls.bias._version,
               self.conv_box.bias._version, self.conv_cls.weight.data_ptr())
        if self._heads is None or self._heads[0] != ver:
            with torch.no_grad():
                w = torch.cat([self.conv_cls.weight.flatten(1), self.conv_box.weight.flatten(1)], 0).t().contiguous()
                b = torch.cat([self.conv_cls.bias, self.conv_box.bias]).contiguous()
            self._heads = (ver, w, b)
        return self._heads[1], self._heads[2]

    def assign_targets(self, gt_boxes):
        """gt_boxes [1, M, 8] (x,y,z,dx,dy,dz,yaw,class) -> dict of one-element lists, shapes as center_head.py:126-169:
        heatmaps [1,ncls,H,W], anno_boxes [1,MAX_OBJS,8], inds int64 [1,MAX_OBJS], masks uint8 [1,MAX_OBJS]."""
        if gt_boxes.dim() != 3 or gt_boxes.shape[0] != 1:
            raise NotImplementedError("batch 1 per sample (models.py:313)")
        t = self.target_cfg
        f = int(t["OUT_SIZE_FACTOR"])
        W, H = int(self.grid_size[0]) // f, int(self.grid_size[1]) // f
        heat, anno, inds, masks = ops.center_targets(
            gt_boxes[0].float().contiguous(), int(t["MAX_OBJS"]), H, W, len(self.class_names[0]),
            float(self.point_cloud_range[0]), float(self.point_cloud_range[1]), float(t["VOXEL_SIZE"][0]), float(t["VOXEL_SIZE"][1]),
            f, float(t["GAUSSIAN_OVERLAP"]), int(t["MIN_RADIUS"]),
            range_is_fp64=not np.issubdtype(np.asarray(self.point_cloud_range).dtype, np.integer))
        return {"heatmaps": [heat.unsqueeze(0)], "anno_boxes": [anno.unsqueeze(0)], "inds": [inds.unsqueeze(0)],
                "masks": [masks.unsqueeze(0)]}

    def get_cls_layer_loss(self):                                        # center_head.py:288-303
        pred = clip_sigmoid(self.forward_ret_dict["cls_preds"]).permute(0, 3, 1, 2)
        gt = self.forward_ret_dict["heatmaps"][0]
        num_pos = gt.eq(1).float().sum().clamp(min=1.0)
        w = self.model_cfg["LOSS_CONFIG"]["LOSS_WEIGHTS"]
        cls_loss = gaussian_focal_loss(pred, gt).sum() / num_pos * w["cls_weight"]
        return cls_loss, {"rpn_loss_cls": cls_loss.detach()}

    def get_box_reg_layer_loss(self):                                    # center_head.py:306-331
        target_box, inds, masks = (self.forward_ret_dict[k][0] for k in ("anno_boxes", "inds", "masks"))
        num = masks.float().sum()
        pred = self.forward_ret_dict["box_preds"]
        pred = pred.view(pred.size(0), -1, pred.size(3))
        pred = pred.gather(1, inds.unsqueeze(2).expand(inds.size(0), inds.size(1), pred.size(2)))
        mask = masks.unsqueeze(2).expand_as(target_box).float() * (~torch.isnan(target_box)).float()
        w = self.model_cfg["LOSS_CONFIG"]["LOSS_WEIGHTS"]
        bbox_weights = mask * mask.new_tensor(w["code_weights"])
        loc_loss = (torch.abs(pred - target_box) * bbox_weights).sum() / (num + 1e-4) * w["loc_weight"]
        return loc_loss, {"rpn_loss_loc": loc_loss.detach()}

    def get_loss(self):                                                  # center_head.py:279-286 (values stay on the device)
        cls_loss, tb = self.get_cls_layer_loss()
        box_loss, tb_box = self.get_box_reg_layer_loss()
        tb.update(tb_box)
        rpn_loss = cls_loss + box_loss
        tb["rpn_loss"] = rpn_loss.detach()
        return rpn_loss, tb

    def forward(self, data_dict, Model_mode):
        t = self.target_cfg
        dec = (t["OUT_SIZE_FACTOR"], t["VOXEL_SIZE"][0], t["VOXEL_SIZE"][1], self.point_cloud_range[0], self.point_cloud_range[1])
        nhwc = data_dict.get("spatial_features_2d_nhwc")
        if Model_mode == "train" and nhwc is not None:
            raise RuntimeError("CenterHead: 'train' mode needs the module in training mode (model.train())")
        if nhwc is not None:
            x, H, W = nhwc
            w, b = self._fused_heads()
            head = ops.linear(x, w, bias=b)                               # [H*W, ncls+8]
            cls_v, box_v = head[:, :self.num_class], head[:, self.num_class:]
            boxes, scores, labels = ops.center_decode(cls_v, box_v, *dec, hw=(H, W))
            data_dict["batch_cls_preds"] = cls_v.unsqueeze(0)
        else:
            x = data_dict["spatial_features_2d"]
            cls = self.conv_cls(x)                                         # [1, ncls, H, W]
            box = self.conv_box(x)                                         # [1, 8, H, W]
            if cls.shape[0] != 1:
                raise NotImplementedError("batch 1 per sample (models.py:313)")
            if Model_mode == "train":                                      # center_head.py:72-90
                self.forward_ret_dict["cls_preds"] = cls.permute(0, 2, 3, 1).contiguous()
                self.forward_ret_dict["box_preds"] = box.permute(0, 2, 3, 1).contiguous()
                self.forward_ret_dict.update(self.assign_targets(data_dict["gt_boxes"]))
            boxes, scores, labels = ops.center_decode(cls[0].detach(), box[0].detach(), *dec)
            data_dict["batch_cls_preds"] = cls[0].detach().permute(1, 2, 0).reshape(1, -1, self.num_class)
        data_dict["batch_box_preds"] = boxes.unsqueeze(0)
        data_dict["cls_preds_normalized"] = False
        data_dict["_decoded"] = (boxes, scores, labels)                     # sigmoid / class max already done on device
        return data_dict
