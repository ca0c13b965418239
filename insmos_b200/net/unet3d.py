"""3D sparse U-Net with BEV instance head and instance-feature fusion (a8-a13).

Mirrors models/backbones_3d/spconv_unet.py:71-416 (attribute names = state_dict keys).  All sparse
convolutions run on the shared tiled rule books (indice_key reuse), every conv+BN+ReLU triple is one
kernel in eval mode, the 5 host round trips of the reference (NMS mask D2H, 4x Array_Index D2H/H2D)
are replaced by device kernels.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

import spconv.pytorch as spconv
from spconv.pytorch.utils import gather_features_by_pc_voxel_id

from insmos_b200 import autograd, ops
from .bev import BaseBEVBackbone, CenterHead, HeightCompression
from .detect import InstanceBoxes, post_processing


class SparseBasicBlock(spconv.SparseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, indice_key=None, norm_fn=None):
        super().__init__()
        self.conv1 = spconv.SubMConv3d(inplanes, planes, 3, stride=stride, padding=1, bias=False, indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(planes, planes, 3, stride=1, padding=1, bias=False, indice_key=indice_key)
        self.bn2 = norm_fn(planes)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        if self.training:
            out = self.conv1(x)
            out = out.replace_feature(autograd.batch_norm_train(self.bn1, out.features, relu=True))
            out = self.conv2(out)
            out = out.replace_feature(autograd.batch_norm_train(self.bn2, out.features))
            identity = x.features if self.downsample is None else self.downsample(x).features
            return out.replace_feature(torch.relu(out.features + identity))
        identity = x.features if self.downsample is None else self.downsample(x).features
        out = self.conv1(x, bn=self.bn1, relu=True)
        return self.conv2(out, bn=self.bn2, residual=identity, relu=True)


class UNetV2(nn.Module):
    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, mos_class, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.sparse_shape = grid_size[::-1] + [1, 0, 0]                       # [41, 1000, 1200]
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        block = partial(self.post_act_block, norm_fn=norm_fn)
        M = model_cfg["MODEL"]
        self.post_process = M["POST_PROCESSING"]
        self.num_class = M["DENSE_HEAD"]["NUM_CLASS"]
        nc = self.num_class

        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1"), norm_fn(16), nn.ReLU())
        self.conv1 = spconv.SparseSequential(block(16, 16, 3, padding=1, indice_key="subm1"))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, stride=2, padding=1, indice_key="spconv2", conv_type="spconv"),
            block(32, 32, 3, padding=1, indice_key="subm2"), block(32, 32, 3, padding=1, indice_key="subm2"))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, stride=2, padding=1, indice_key="spconv3", conv_type="spconv"),
            block(64, 64, 3, padding=1, indice_key="subm3"), block(64, 64, 3, padding=1, indice_key="subm3"))
        self.conv4 = spconv.SparseSequential(
            block(64, 128, 3, stride=2, padding=1, indice_key="spconv4", conv_type="spconv"),
            block(128, 128, 3, padding=1, indice_key="subm4"), block(128, 128, 3, padding=1, indice_key="subm4"))
        if model_cfg.get("RETURN_ENCODED_TENSOR", True):
            self.conv_out = spconv.SparseSequential(
                spconv.SparseConv3d(128, 128, (3, 1, 1), stride=(2, 1, 1), padding=model_cfg.get("last_pad", 0), bias=False,
                                    indice_key="spconv_down2"), norm_fn(128), nn.ReLU())
        else:
            self.conv_out = None

        # ---- instance detection on the BEV map
        self.point_cloud_range = np.array(model_cfg["DATA"]["POINT_CLOUD_RANGE"])
        self.voxel_size = model_cfg["DATA"]["VOXEL_SIZE"]
        self.grid_size = np.round((self.point_cloud_range[3:6] - self.point_cloud_range[0:3]) / np.array(self.voxel_size)).astype(np.int64)
        self.to_bev = HeightCompression(M["MAP_TO_BEV"])
        self.bev_backbone = BaseBEVBackbone(M["BACKBONE_2D"], input_channels=self.to_bev.num_bev_features)
        self.center_head = CenterHead(M["DENSE_HEAD"], input_channels=M["BACKBONE_2D"]["NUM_UPSAMPLE_FILTERS"][0],
                                      num_class=nc if not M["DENSE_HEAD"]["CLASS_AGNOSTIC"] else 1,
                                      class_names=M["DENSE_HEAD"]["CLASE_NAME"], grid_size=self.grid_size,
                                      point_cloud_range=self.point_cloud_range,
                                      predict_boxes_when_training=model_cfg.get("ROI_HEAD", False))

        # ---- upsample fusion
        self.inv_conv_out = spconv.SparseInverseConv3d(128, 128, (3, 1, 1), bias=False, indice_key="spconv_down2")
        self.conv_up_instance_block = block(128 + nc, 128, 3, padding=1, indice_key="subm4")
        self.conv_up_instance_block_up4 = block(64 + nc, 64, 3, padding=1, indice_key="subm3")
        self.conv_up_instance_block_up3 = block(32 + nc, 32, 3, indice_key="subm2")
        self.conv_up_instance_block_up2 = block(16 + nc, 16, 3, indice_key="subm1")
        self.conv_up_instance_block_up1 = block(16 + nc, 16, 3, padding=1, indice_key="subm1")
        self.conv_up_t4 = SparseBasicBlock(128, 128, indice_key="subm4", norm_fn=norm_fn)
        self.conv_up_m4 = block(256, 128, 3, padding=1, indice_key="subm4")
        self.inv_conv4 = block(128, 64, 3, indice_key="spconv4", conv_type="inverseconv")
        self.conv_up_t3 = SparseBasicBlock(64, 64, indice_key="subm3", norm_fn=norm_fn)
        self.conv_up_m3 = block(128, 64, 3, padding=1, indice_key="subm3")
        self.inv_conv3 = block(64, 32, 3, indice_key="spconv3", conv_type="inverseconv")
        self.conv_up_t2 = SparseBasicBlock(32, 32, indice_key="subm2", norm_fn=norm_fn)
        self.conv_up_m2 = block(64, 32, 3, indice_key="subm2")
        self.inv_conv2 = block(32, 16, 3, indice_key="spconv2", conv_type="inverseconv")
        self.conv_up_t1 = SparseBasicBlock(16, 16, indice_key="subm1", norm_fn=norm_fn)
        self.conv_up_m1 = block(32, 16, 3, indice_key="subm1")
        self.conv_up_out = spconv.SparseSequential(block(16, 16, 3, padding=1, indice_key="subm1"))
        self.mos_seg_layer = nn.Linear(16, mos_class, bias=True)
        self.num_point_features = 16

    @staticmethod
    def post_act_block(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0, conv_type="subm", norm_fn=None):
        if conv_type == "subm":
            conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
        elif conv_type == "spconv":
            conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                                       indice_key=indice_key)
        elif conv_type == "inverseconv":
            conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
        else:
            raise NotImplementedError(conv_type)
        return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())

    @staticmethod
    def ur_block(x_lateral, x_bottom, conv_t, conv_m, conv_inv):
        """spconv_unet.py:213-238: transform lateral, concat with bottom, merge conv + pair-sum skip, inverse conv."""
        x_trans = conv_t(x_lateral)
        if autograd.needs_grad(x_bottom.features, x_trans.features):        # training: plain torch glue keeps the graph
            cat = torch.cat((x_bottom.features, x_trans.features), dim=1)
            x_m = conv_m(x_trans.replace_feature(cat))
            red = cat.view(cat.shape[0], x_m.features.shape[1], -1).sum(dim=2)
            return conv_inv(x_m.replace_feature(x_m.features + red))
        cat = ops.concat2(x_bottom.features, x_trans.features)
        x_m = conv_m(x_trans.replace_feature(cat))
        assert cat.shape[1] == 2 * x_m.features.shape[1]
        return conv_inv(x_m.replace_feature(ops.pairsum_add(x_m.features, cat)))

    UR_block_forward = ur_block

    def forward(self, batch_dict, Model_mode):
        feats, coords = batch_dict["voxel_features"], batch_dict["voxel_coords"]
        if feats.shape[1] % 8:                                             # 7 point features -> 8 (zero column, zero weight rows)
            feats = ops.concat2(feats, feats.new_zeros((feats.shape[0], 8 - feats.shape[1] % 8)))
        x = spconv.SparseConvTensor(features=feats, indices=coords.int(), spatial_shape=self.sparse_shape, batch_size=1,
                                    coordset=batch_dict.get("_voxel_set"))
        # the output coordinates of every strided convolution are queued one level ahead of its forward: the host read
        # of their count then overlaps the previous level's kernels instead of draining the queue
        down = [m for seq in (self.conv2, self.conv3, self.conv4, self.conv_out) for m in seq.modules()
                if isinstance(m, spconv.SparseConv3d)]
        d = down[0].prefetch_indices(x.coordset, x.spatial_shape, x.indice_dict)
        x = self.conv_input(x)
        x_conv1 = self.conv1(x)
        d = down[1].prefetch_indices(d.out_set, d.out_shape, x.indice_dict)
        x_conv2 = self.conv2(x_conv1)
        d = down[2].prefetch_indices(d.out_set, d.out_shape, x.indice_dict)
        x_conv3 = self.conv3(x_conv2)
        down[3].prefetch_indices(d.out_set, d.out_shape, x.indice_dict)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        batch_dict["encoded_spconv_tensor"] = out
        batch_dict["encoded_spconv_tensor_stride"] = 8

        # ---- instance detection
        batch_dict = self.to_bev(batch_dict)
        if batch_dict.get("spatial_features") is not None:
            batch_dict["current_bev"] = batch_dict["spatial_features"][-1].unsqueeze(0)
            batch_dict["current_bev_nhwc"] = None
        else:                                                               # eval: channels-last tensor-core path
            batch_dict["current_bev_nhwc"] = batch_dict["spatial_features_nhwc"]
        batch_dict = self.bev_backbone(batch_dict)
        batch_dict = self.center_head(batch_dict, Model_mode)
        pred_dicts, recall_dicts = post_processing(batch_dict, self.post_process, self.num_class)
        if batch_dict.get("instance_boxes_override") is not None:
            # externally supplied detections for the instance-fusion stage (e.g. a tracker's boxes; the parity
            # tests use it to compare the decoder under identical discrete decisions).  pred_dicts is still returned.
            fuse_from = batch_dict["instance_boxes_override"]
        else:
            fuse_from = pred_dicts[0]

        # ---- upsample fusion with per-level instance bits (Array_Index on device)
        inst = InstanceBoxes(fuse_from, self.point_cloud_range[0:3], self.voxel_size,
                             batch_dict["encoded_spconv_tensor_stride"], self.num_class)
        def with_bits(t, mult, bits=None):
            """features + instance bits; in training the tensor remembers how many leading columns carry a gradient"""
            c = t.features.shape[1]
            if bits is None:
                f, bits = inst.concat_bits(t.features, t.indices, mult)
            else:
                f = torch.cat([t.features, bits], 1) if autograd.needs_grad(t.features) else ops.concat2(t.features, bits)
            r = t.replace_feature(f)
            r.grad_cols = c
            return r, bits

        inv_bev = self.inv_conv_out(out)
        x_inst = self.conv_up_instance_block(with_bits(inv_bev, 1)[0])
        x_up4 = self.ur_block(x_inst, x_inst, self.conv_up_t4, self.conv_up_m4, self.inv_conv4)

        x_up4_inst = self.conv_up_instance_block_up4(with_bits(x_up4, 2)[0])
        x_up3 = self.ur_block(x_conv3, x_up4_inst, self.conv_up_t3, self.conv_up_m3, self.inv_conv3)

        x_up3_inst = self.conv_up_instance_block_up3(with_bits(x_up3, 4)[0])
        x_up2 = self.ur_block(x_conv2, x_up3_inst, self.conv_up_t2, self.conv_up_m2, self.inv_conv2)

        x_up2_b, bits1 = with_bits(x_up2, 8)
        x_up2_inst = self.conv_up_instance_block_up2(x_up2_b)
        x_up1 = self.ur_block(x_conv1, x_up2_inst, self.conv_up_t1, self.conv_up_m1, self.conv_up_out)

        # the finest level re-uses the bits computed for x_up2 (same voxel rows): spconv_unet.py:401
        x_up1_inst = self.conv_up_instance_block_up1(with_bits(x_up1, 8, bits1)[0])

        if autograd.needs_grad(x_up1_inst.features, self.mos_seg_layer.weight):
            seg = autograd.linear(x_up1_inst.features, self.mos_seg_layer.weight.t(), self.mos_seg_layer.bias)
        else:
            seg = ops.linear(x_up1_inst.features, self.mos_seg_layer.weight.t().contiguous(), bias=self.mos_seg_layer.bias)
        point_seg = gather_features_by_pc_voxel_id(seg, batch_dict["list_pc_voxel_id"][-1])
        if Model_mode == "train":
            return self.center_head.get_loss(), point_seg                   # spconv_unet.py:413-414
        return point_seg, pred_dicts, recall_dicts
