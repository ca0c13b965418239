"""Functional CPU restatement of the InsMOS forward graph over a state_dict (TEST INFRASTRUCTURE).

Follows, step by step (paths relative to the reference repository):
  models/models.py:297-376                      per-sample pipeline
  models/backbones_3d/motionnet.py:21-50        4D quantise, TensorField.sparse, MinkUNet, slice, t==0 select
  models/MinkowskiEngine/minkunet.py:139-181    MinkUNet14-variant forward (PLANES 8,16,32,64,64,32,16,8)
  models/backbones_3d/voxel_generate.py:17-31, models/backbones_2d/mean_vfe.py:47-52
  models/backbones_3d/spconv_unet.py:267-416    UNetV2 forward incl. UR blocks and instance fusion
  models/backbones_2d/{height_compression.py:14-33, base_bev_backbone.py:84-115, center_head.py:65-98,251-276}
  models/post_process.py:5-24,112-224, models/bbox_post_process/iou3d_nms_utils.py:64-79
on top of oracle/me.py, oracle/sp.py and oracle/native (C).  Pinned against the reference's own code
by tests/golden (see tests/golden/make_golden.py, tests/test_oracle_graph.py).
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import me, native, sp

PC_RANGE = [-60, -50, -3, 60, 50, 1]
VOXEL = [0.1, 0.1, 0.1]
PP = {"SCORE_THRESH": 0.1, "NMS_THRESH": 0.01, "NMS_PRE_MAXSIZE": 4096, "NMS_POST_MAXSIZE": 500}


def _bn(sd, p, x, eps):
    if sd.get("__train__") is not None:                  # training step (oracle/train.py): batch statistics, as model.train() does
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, eps)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


class _ME:
    """MinkUNet part: coordinate sets / kernel maps per tensor stride, cached like ME's coordinate manager."""

    def __init__(self, sd, prefix, coords):
        self.sd, self.p = sd, prefix
        self.sets = {1: coords}
        self.maps = {}
        self.timing = {"maps": 0.0, "conv": 0.0}

    def coords(self, ts):
        if ts not in self.sets:
            t0 = time.perf_counter()
            self.sets[ts], _ = me.stride_coords(self.coords(ts // 2), [ts, ts, ts, 1])
            self.timing["maps"] += time.perf_counter() - t0
        return self.sets[ts]

    def kmap(self, in_ts, out_ts, ksize):
        k = (in_ts, out_ts, tuple(ksize))
        if k not in self.maps:
            t0 = time.perf_counter()
            self.maps[k] = me.kernel_map(self.coords(in_ts), self.coords(out_ts), list(ksize), [in_ts, in_ts, in_ts, 1])
            self.timing["maps"] += time.perf_counter() - t0
        return self.maps[k]

    def conv(self, name, x, ts, ksize=(3, 3, 3, 3), stride=1, transposed=False):
        W = self.sd[self.p + name + ".kernel"]
        bias = self.sd.get(self.p + name + ".bias")
        if W.dim() == 2:
            t0 = time.perf_counter()
            out = me.conv(x, W, None, len(x), bias=bias)
            self.timing["conv"] += time.perf_counter() - t0
            return out, ts
        if transposed:
            out_ts = ts // 2
            maps = me.transpose_map(self.kmap(out_ts, ts, ksize))
        else:
            out_ts = ts * stride
            maps = self.kmap(ts, out_ts, ksize)
        t0 = time.perf_counter()
        out = me.conv(x, W, maps, len(self.coords(out_ts)), bias=bias)
        self.timing["conv"] += time.perf_counter() - t0
        return out, out_ts

    def bn(self, name, x):
        return _bn(self.sd, self.p + name + ".bn", x, 1e-5)

    def block(self, name, x, ts):                      # ME BasicBlock (resnet_block), downsample = conv1x1 + BN
        p = name + ".0."
        out, _ = self.conv(p + "conv1", x, ts)
        out = torch.relu(self.bn(p + "norm1", out))
        out, _ = self.conv(p + "conv2", out, ts)
        out = self.bn(p + "norm2", out)
        if (self.p + p + "downsample.0.kernel") in self.sd:
            res, _ = self.conv(p + "downsample.0", x, ts)
            res = self.bn(p + "downsample.1", res)
        else:
            res = x
        return torch.relu(out + res)


def minkunet(sd, prefix, coords, feats, timing=None):
    """minkunet.py:139-181"""
    m = _ME(sd, prefix, coords)
    down = (2, 2, 2, 1)
    out, _ = m.conv("conv0p1s1", feats, 1, ksize=(5, 5, 5, 1))
    p1 = torch.relu(m.bn("bn0", out))
    out, _ = m.conv("conv1p1s2", p1, 1, ksize=down, stride=2)
    b1p2 = m.block("block1", torch.relu(m.bn("bn1", out)), 2)
    out, _ = m.conv("conv2p2s2", b1p2, 2, ksize=down, stride=2)
    b2p4 = m.block("block2", torch.relu(m.bn("bn2", out)), 4)
    out, _ = m.conv("conv3p4s2", b2p4, 4, ksize=down, stride=2)
    out = m.block("block3", torch.relu(m.bn("bn3", out)), 8)
    out, _ = m.conv("convtr5p8s2", out, 8, ksize=down, transposed=True)
    out = m.block("block6", torch.cat([torch.relu(m.bn("bntr5", out)), b2p4], 1), 4)
    out, _ = m.conv("convtr6p4s2", out, 4, ksize=down, transposed=True)
    out = m.block("block7", torch.cat([torch.relu(m.bn("bntr6", out)), b1p2], 1), 2)
    out, _ = m.conv("convtr7p2s2", out, 2, ksize=down, transposed=True)
    out = m.block("block8", torch.cat([torch.relu(m.bn("bntr7", out)), p1], 1), 1)
    out, _ = m.conv("final", out, 1)
    if timing is not None:
        timing.update({"me_maps_s": m.timing["maps"], "me_conv_s": m.timing["conv"],
                       "n_vox": {ts: len(c) for ts, c in m.sets.items()},
                       "pairs": {str(k): int(sum(len(i) for i, _ in v)) for k, v in m.maps.items()}})
    return out


def motionnet(sd, points, dt=0.1, vs=0.1, timing=None):
    """motionnet.py:21-50.  points [N,5] (x,y,z,intensity,t) -> current_point [Nc,7]"""
    pts = torch.as_tensor(points, dtype=torch.float32)
    xyzt = torch.hstack([pts[:, 0:3], pts[:, 4].view(-1, 1)])
    coords, cur = me.quantize_points(xyzt, [vs, vs, vs, dt])
    t0 = time.perf_counter()
    uniq, inv = me.unique_first(coords)
    if timing is not None:
        timing["voxelize4d_s"] = time.perf_counter() - t0
    feats = me.segment_mean(0.5 * torch.ones(len(pts), 1), inv, len(uniq))
    out = minkunet(sd, "model.motion_encoder.MinkUNet.", uniq, feats, timing)
    sliced = out[torch.from_numpy(inv)]
    cur_t = torch.from_numpy(cur)
    return torch.hstack([pts[cur_t, :4], sliced[cur_t, :]])


class _SP:
    def __init__(self, sd, prefix):
        self.sd, self.p, self.keys = sd, prefix, {}

    def bn(self, name, x):
        return _bn(self.sd, self.p + name, x, 1e-3)

    def conv(self, wname, t, key, kind, ksize=(3, 3, 3), stride=(1, 1, 1), pad=(0, 0, 0)):
        feats, ind, shape = t
        W = self.sd[self.p + wname + ".weight"]
        d = self.keys.get(key)
        if kind == "inv":
            maps, out_ind, out_shape = me.transpose_map(d["maps"]), d["in_ind"], d["in_shape"]
        elif kind == "subm":
            if d is None:
                d = self.keys[key] = {"maps": sp.subm_maps(ind, ksize), "in_ind": ind, "in_shape": shape}
            maps, out_ind, out_shape = d["maps"], ind, shape
        else:
            if d is None:
                oind, maps, oshape = sp.sparse_conv_indices(ind, shape, ksize, stride, pad)
                d = self.keys[key] = {"maps": maps, "in_ind": ind, "in_shape": shape, "out_ind": oind, "out_shape": oshape}
            maps, out_ind, out_shape = d["maps"], d["out_ind"], d["out_shape"]
        return sp.conv(feats, W, maps, len(out_ind)), out_ind, out_shape

    def cbr(self, name, t, key, kind="subm", **kw):            # post_act_block: conv(.0) + BN(.1) + ReLU
        f, ind, shape = self.conv(name + ".0", t, key, kind, **kw)
        return torch.relu(self.bn(name + ".1", f)), ind, shape

    def basic(self, name, t, key):                              # SparseBasicBlock spconv_unet.py:71-106
        f, ind, shape = self.conv(name + ".conv1", t, key, "subm")
        f = torch.relu(self.bn(name + ".bn1", f))
        f, _, _ = self.conv(name + ".conv2", (f, ind, shape), key, "subm")
        f = self.bn(name + ".bn2", f)
        return torch.relu(f + t[0]), ind, shape


def bev_head(sd, p, spatial):
    """base_bev_backbone.py:84-115 + center_head.py:65-98,251-276"""
    x = spatial
    idx = [1, 4, 7, 10, 13, 16]
    for j, i in enumerate(idx):
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)) if j == 0 else x, sd[p + "bev_backbone.blocks.0.%d.weight" % i], padding=0 if j == 0 else 1)
        x = torch.relu(_bn(sd, p + "bev_backbone.blocks.0.%d" % (i + 1), x, 1e-3))
    x = F.conv_transpose2d(x, sd[p + "bev_backbone.deblocks.0.0.weight"], stride=2)
    x = torch.relu(_bn(sd, p + "bev_backbone.deblocks.0.1", x, 1e-3))
    cls = F.conv2d(x, sd[p + "center_head.conv_cls.weight"], sd[p + "center_head.conv_cls.bias"]).permute(0, 2, 3, 1).contiguous()
    box = F.conv2d(x, sd[p + "center_head.conv_box.weight"], sd[p + "center_head.conv_box.bias"]).permute(0, 2, 3, 1).contiguous()
    b, H, W, _ = box.shape
    bp = box.reshape(b, H * W, 8)
    ys, xs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing="ij")
    xs = xs.reshape(1, -1, 1) + bp[:, :, 0:1]
    ys = ys.reshape(1, -1, 1) + bp[:, :, 1:2]
    xs = xs * 4 * 0.1 + PC_RANGE[0]
    ys = ys * 4 * 0.1 + PC_RANGE[1]
    boxes = torch.cat([xs, ys, bp[..., 2:3], torch.exp(bp[..., 3:6]), torch.atan2(bp[..., 6:7], bp[..., 7:8])], dim=2)
    if sd.get("__train__") is not None:                  # raw head outputs for the losses (center_head.py:72-76)
        sd["__train__"]["cls_preds"], sd["__train__"]["box_preds"] = cls, box
    return cls.view(b, H * W, -1), boxes


def post_process(cls_preds, box_preds):
    """post_process.py:112-224 + class_agnostic_nms :5-24 + nms_gpu"""
    cls_preds, box_preds = cls_preds.detach(), box_preds.detach()       # (no-op at inference; the training oracle carries a graph)
    scores, labels = torch.max(torch.sigmoid(cls_preds[0]), dim=-1)
    labels = labels + 1
    boxes = box_preds[0]
    mask = scores >= PP["SCORE_THRESH"]
    s, b = scores[mask], boxes[mask]
    selected = torch.zeros(0, dtype=torch.long)
    if s.shape[0] > 0:
        top_s, idx = torch.topk(s, k=min(PP["NMS_PRE_MAXSIZE"], s.shape[0]))
        bn = b[idx]
        order = top_s.sort(0, descending=True)[1]
        keep = torch.from_numpy(native.nms(bn[order][:, :7].contiguous().numpy(), PP["NMS_THRESH"]))
        selected = idx[order[keep][:PP["NMS_POST_MAXSIZE"]]]
    selected = mask.nonzero().view(-1)[selected]
    return {"pred_boxes": boxes[selected], "pred_scores": scores[selected], "pred_labels": labels[selected],
            "n_cand": int(mask.sum()), "all_scores": scores, "all_boxes": boxes, "all_labels": labels}


def boxes_iou3d(a, b):
    """iou3d_nms_utils.py:27-61 (boxes_iou3d_gpu) in fp32: BEV overlap x height overlap / union volume.  a [N,7], b [M,7]."""
    f = np.float32
    a, b = np.asarray(a, dtype=f), np.asarray(b, dtype=f)
    bev = native.overlap_matrix(a[:, :7], b[:, :7])
    a_max, a_min = (a[:, 2] + a[:, 5] / f(2))[:, None], (a[:, 2] - a[:, 5] / f(2))[:, None]
    b_max, b_min = (b[:, 2] + b[:, 5] / f(2))[None, :], (b[:, 2] - b[:, 5] / f(2))[None, :]
    h = np.maximum(np.minimum(a_max, b_max) - np.maximum(a_min, b_min), f(0))
    o3d = bev * h
    va, vb = (a[:, 3] * a[:, 4] * a[:, 5])[:, None], (b[:, 3] * b[:, 4] * b[:, 5])[None, :]
    return o3d / np.maximum(va + vb - o3d, f(1e-6))


def recall_record(pred_boxes, gt_boxes, thresh_list=(0.3, 0.5, 0.7)):
    """post_process.py:66-109 for one sample without rois: {'gt', 'roi_<t>': 0, 'rcnn_<t>'}"""
    rec = {"gt": 0}
    for t in thresh_list:
        rec["roi_%s" % str(t)] = 0
        rec["rcnn_%s" % str(t)] = 0
    gt = np.asarray(gt_boxes, dtype=np.float32)
    k = len(gt) - 1
    while k > 0 and gt[k].sum() == 0:
        k -= 1
    gt = gt[:k + 1]
    if len(gt) > 0:
        pb = np.asarray(pred_boxes, dtype=np.float32)
        if len(pb) > 0:
            best = boxes_iou3d(pb[:, :7], gt[:, :7]).max(axis=0)
            for t in thresh_list:
                rec["rcnn_%s" % str(t)] += int((best > np.float32(t)).sum())
        rec["gt"] += len(gt)
    return rec


def instance_bits(ind, boxes8):
    xyz = np.ascontiguousarray(np.asarray(ind)[:, [3, 2, 1]], dtype=np.int32)
    return torch.from_numpy(native.find_features_by_bbox_with_yaw(xyz, boxes8.numpy())).float()


def unet(sd, voxel_features, voxel_coords, pc_voxel_id, timing=None, pred_override=None):
    """spconv_unet.py:267-416.  voxel_coords [M,4] (0,z,y,x)"""
    p = "model.unet."
    s = _SP(sd, p)
    t0 = time.perf_counter()
    x = (torch.as_tensor(voxel_features), np.asarray(voxel_coords, dtype=np.int32), [41, 1000, 1200])
    x = s.cbr("conv_input", x, "subm1")
    c1 = s.cbr("conv1.0", x, "subm1")
    c2 = s.cbr("conv2.0", c1, "spconv2", "spconv", stride=(2, 2, 2), pad=(1, 1, 1))
    c2 = s.cbr("conv2.2", s.cbr("conv2.1", c2, "subm2"), "subm2")
    c3 = s.cbr("conv3.0", c2, "spconv3", "spconv", stride=(2, 2, 2), pad=(1, 1, 1))
    c3 = s.cbr("conv3.2", s.cbr("conv3.1", c3, "subm3"), "subm3")
    c4 = s.cbr("conv4.0", c3, "spconv4", "spconv", stride=(2, 2, 2), pad=(1, 1, 1))
    c4 = s.cbr("conv4.2", s.cbr("conv4.1", c4, "subm4"), "subm4")
    out = s.cbr("conv_out", c4, "spconv_down2", "spconv", ksize=(3, 1, 1), stride=(2, 1, 1), pad=(0, 0, 0))
    t1 = time.perf_counter()
    dense = sp.dense(out[0], out[1], out[2])
    n, c, d, h, w = dense.shape
    cls, boxes = bev_head(sd, p, dense.view(n, c * d, h, w))
    t2 = time.perf_counter()
    pred = post_process(cls, boxes)
    t3 = time.perf_counter()
    fuse = pred if pred_override is None else pred_override

    b = torch.as_tensor(fuse["pred_boxes"]).clone()
    for dd in range(3):                                                   # spconv_unet.py:324-329
        b[:, dd] = (b[:, dd] - PC_RANGE[dd]) / VOXEL[dd] / 8
        b[:, 3 + dd] = b[:, 3 + dd] / VOXEL[dd] / 8
    b8 = torch.hstack([b, torch.as_tensor(fuse["pred_labels"]).view(-1, 1).to(b.dtype)])

    def with_bits(t):
        bits = instance_bits(t[1], b8)
        return (torch.cat([t[0], bits], dim=1), t[1], t[2]), bits

    def ur(lateral, bottom, name_t, name_m, name_inv, key, inv_key, inv_kind):
        tr = s.basic(name_t, lateral, key)
        cat = torch.cat((bottom[0], tr[0]), dim=1)
        xm = s.cbr(name_m, (cat, tr[1], tr[2]), key)
        red = cat.view(cat.shape[0], xm[0].shape[1], -1).sum(dim=2)
        return s.cbr(name_inv, (xm[0] + red, xm[1], xm[2]), inv_key, inv_kind)

    inv = s.conv("inv_conv_out", out, "spconv_down2", "inv")
    xin, _ = with_bits(inv)
    x_inst = s.cbr("conv_up_instance_block", xin, "subm4")
    up4 = ur(x_inst, x_inst, "conv_up_t4", "conv_up_m4", "inv_conv4", "subm4", "spconv4", "inv")
    b8[:, 0:6] = 2 * b8[:, 0:6]
    xin, _ = with_bits(up4)
    up4i = s.cbr("conv_up_instance_block_up4", xin, "subm3")
    up3 = ur(c3, up4i, "conv_up_t3", "conv_up_m3", "inv_conv3", "subm3", "spconv3", "inv")
    b8[:, 0:6] = 2 * b8[:, 0:6]
    xin, _ = with_bits(up3)
    up3i = s.cbr("conv_up_instance_block_up3", xin, "subm2")
    up2 = ur(c2, up3i, "conv_up_t2", "conv_up_m2", "inv_conv2", "subm2", "spconv2", "inv")
    b8[:, 0:6] = 2 * b8[:, 0:6]
    xin, bits1 = with_bits(up2)
    up2i = s.cbr("conv_up_instance_block_up2", xin, "subm1")
    up1 = ur(c1, up2i, "conv_up_t1", "conv_up_m1", "conv_up_out.0", "subm1", "subm1", "subm")
    up1i = s.cbr("conv_up_instance_block_up1", (torch.cat([up1[0], bits1], dim=1), up1[1], up1[2]), "subm1")
    seg = F.linear(up1i[0], sd[p + "mos_seg_layer.weight"], sd[p + "mos_seg_layer.bias"])
    logits = sp.gather_features_by_pc_voxel_id(seg, pc_voxel_id)
    if timing is not None:
        timing.update({"unet_encoder_s": t1 - t0, "bev_s": t2 - t1, "postprocess_s": t3 - t2, "decoder_s": time.perf_counter() - t3})
    return logits, pred


def forward(sd, points, timing=None, pred_override=None):
    """models.py:297-376 ('test' mode, one sample).  Returns dict with logits [Nc,3] and detections."""
    sd = {k: torch.as_tensor(v) for k, v in sd.items()}
    t0 = time.perf_counter()
    cur = motionnet(sd, points, timing=timing)
    t1 = time.perf_counter()
    vox, coords, num, ids = sp.point_to_voxel(cur.numpy(), VOXEL, PC_RANGE, 5, 100000)
    vf = sp.mean_vfe(vox, num)
    vc = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], axis=1)
    t2 = time.perf_counter()
    logits, pred = unet(sd, vf, vc, ids, timing, pred_override)
    if timing is not None:
        timing.update({"motionnet_s": t1 - t0, "voxelize3d_s": t2 - t1, "total_s": time.perf_counter() - t0})
    return {"logits": logits, "current_point": cur, "voxel_coords": vc, "voxel_features": vf, "pc_voxel_id": ids, **pred}
