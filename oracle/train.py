"""CPU restatement of ONE InsMOS training step (TEST INFRASTRUCTURE; SURVEY.md section 8f row N3, BASELINE config 5).

Follows (paths relative to the reference repository):
  models/models.py:297-347,366-368              per-sample loss = loss_rpn + loss_mos + loss_motion_encoder, mean over the list
  models/loss.py:20-34                          MOSLoss (ignored class to -inf, softmax, log(clamp 1e-8), weighted NLL)
  models/backbones_2d/center_head.py:171-249    get_targets_single (Gaussian heat map, anno boxes, indices, masks)
  models/backbones_2d/center_head.py:279-331    Gaussian focal loss / avg_factor, masked L1 / (num + 1e-4), loss weights
  models/backbones_2d/center_head.py:333-345,348-399,401-427,590-631   clip_sigmoid, gaussian_2d, draw, radius, focal, l1
on top of oracle/graph.py run with batch-statistics BatchNorm, and torch autograd for the gradients (the reference trains
through the autograd Functions of MinkowskiEngine / spconv; their CPU algorithm -- per-offset index_select -> mm ->
index_add_ -- is what oracle/me.py restates, so autograd through it is the reference's gradient).
Pinned against the reference's own model code in train mode by tests/golden/train_small.npz
(tests/golden/make_golden_train.py, tests/test_oracle_train.py).
"""
import numpy as np
import torch

from . import graph, sp

TARGET = {"MAX_OBJS": 100, "VOXEL_SIZE": [0.1, 0.1, 0.1], "OUT_SIZE_FACTOR": 4, "GAUSSIAN_OVERLAP": 0.1, "MIN_RADIUS": 2}
LOSS_WEIGHTS = {"cls_weight": 1.0, "loc_weight": 2.0, "code_weights": [1.0] * 8}
GRID = [1200, 1000, 40]
NUM_CLASS = 3


def gaussian_radius(height, width, min_overlap):
    """center_head.py:401-427 on fp32 scalars (the reference evaluates it on 0-dim float32 tensors)."""
    f = np.float32
    height, width, mo = f(height), f(width), f(min_overlap)
    b1 = height + width
    c1 = width * height * (f(1) - mo) / (f(1) + mo)
    r1 = (b1 + np.sqrt(b1 * b1 - f(4) * c1)) / f(2)
    b2 = f(2) * (height + width)
    c2 = (f(1) - mo) * width * height
    r2 = (b2 + np.sqrt(b2 * b2 - f(16) * c2)) / f(2)
    a3 = f(4) * mo
    b3 = f(-2) * mo * (height + width)
    c3 = (mo - f(1)) * width * height
    r3 = (b3 + np.sqrt(b3 * b3 - f(4) * a3 * c3)) / f(2)
    return min(r1, r2, r3)


def center_targets(gt_boxes):
    """get_targets_single for one sample.  gt_boxes [M,8] -> heatmap [ncls,H,W] f32, anno_boxes [MAX_OBJS,8] f32,
    inds int64 [MAX_OBJS], masks uint8 [MAX_OBJS]."""
    f32 = np.float32
    gt = np.asarray(gt_boxes, dtype=f32)
    fac = TARGET["OUT_SIZE_FACTOR"]
    W, H = GRID[0] // fac, GRID[1] // fac
    vx, vy = f32(TARGET["VOXEL_SIZE"][0]), f32(TARGET["VOXEL_SIZE"][1])
    heat = np.zeros((NUM_CLASS, H, W), dtype=f32)
    anno = np.zeros((TARGET["MAX_OBJS"], 8), dtype=f32)
    inds = np.zeros(TARGET["MAX_OBJS"], dtype=np.int64)
    masks = np.zeros(TARGET["MAX_OBJS"], dtype=np.uint8)
    for k in range(min(len(gt), TARGET["MAX_OBJS"])):
        cls_id = int(gt[k, 7] - f32(1))
        width = gt[k, 3] / vx / f32(fac)
        length = gt[k, 4] / vy / f32(fac)
        if not (width > 0 and length > 0 and cls_id > -1):
            continue
        radius = max(TARGET["MIN_RADIUS"], int(gaussian_radius(length, width, TARGET["GAUSSIAN_OVERLAP"])))
        # center_head.py:212-220: pc_range is an INTEGER tensor (config.yaml:6 lists integers), so fp32 box - int64 range stays
        # fp32 and the whole expression is evaluated in fp32 (a float range in the config would promote it to fp64)
        cx = (gt[k, 0] - f32(graph.PC_RANGE[0])) / vx / f32(fac)
        cy = (gt[k, 1] - f32(graph.PC_RANGE[1])) / vy / f32(fac)
        x, y = int(cx), int(cy)
        if not (0 <= x < W and 0 <= y < H):
            continue
        diameter = 2 * radius + 1
        sigma = diameter / 6
        yy, xx = np.ogrid[-radius:radius + 1, -radius:radius + 1]
        g = np.exp(-(xx * xx + yy * yy) / (2 * sigma * sigma))
        g[g < np.finfo(g.dtype).eps * g.max()] = 0
        left, right = min(x, radius), min(W - x, radius + 1)
        top, bottom = min(y, radius), min(H - y, radius + 1)
        win = heat[cls_id, y - top:y + bottom, x - left:x + right]
        np.maximum(win, g[radius - top:radius + bottom, radius - left:radius + right].astype(f32), out=win)
        inds[k] = y * W + x
        masks[k] = 1
        anno[k] = [cx - f32(x), cy - f32(y), gt[k, 2], np.log(gt[k, 3]), np.log(gt[k, 4]), np.log(gt[k, 5]),
                   np.sin(gt[k, 6]), np.cos(gt[k, 6])]
    return heat, anno, inds, masks


def mos_loss(logits, labels, weight, ignore_index=(0,)):
    logits = logits.clone()
    logits[:, list(ignore_index)] = -float("inf")
    logp = torch.log(torch.softmax(logits, dim=1).clamp(min=1e-8))
    return torch.nn.functional.nll_loss(logp, torch.as_tensor(labels).long(), weight=weight)


def rpn_loss(cls_preds, box_preds, targets):
    """cls_preds [1,H,W,ncls] raw, box_preds [1,H,W,8] raw (channels last, as center_head.py:72-73 stores them);
    targets = center_targets(...)"""
    heat, anno, inds, masks = (torch.from_numpy(np.asarray(t)) for t in targets)
    heat, anno = heat.to(cls_preds.dtype), anno.to(cls_preds.dtype)
    pred = torch.clamp(torch.sigmoid(cls_preds), min=1e-4, max=1 - 1e-4).permute(0, 3, 1, 2)
    gt = heat.unsqueeze(0)
    eps = 1e-12
    pos = -(pred + eps).log() * (1 - pred).pow(2.0) * gt.eq(1)
    neg = -(1 - pred + eps).log() * pred.pow(2.0) * (1 - gt).pow(4.0)
    cls_loss = (pos + neg).sum() / max(float(gt.eq(1).sum()), 1.0) * LOSS_WEIGHTS["cls_weight"]
    bp = box_preds.reshape(1, -1, 8)
    sel = bp.gather(1, inds.view(1, -1, 1).expand(1, -1, 8))
    target = anno.unsqueeze(0)
    mask = masks.view(1, -1, 1).expand_as(target).to(target.dtype) * (~torch.isnan(target)).to(target.dtype)
    w = mask * torch.tensor(LOSS_WEIGHTS["code_weights"], dtype=mask.dtype)
    loc_loss = (torch.abs(sel - target) * w).sum() / (masks.float().sum() + 1e-4) * LOSS_WEIGHTS["loc_weight"]
    return cls_loss, loc_loss


def train_step(sd, points, labels, gt_boxes, pred_override=None, use_motion_loss=True, dtype=torch.float32):
    """forward in train mode + backward for one sample.  sd: {key: tensor}; float parameters get requires_grad.
    dtype=torch.float64 evaluates the feature arithmetic (not the fp32 voxelisation) in double precision: the gradients
    of this graph are ill-conditioned in fp32 at depth (train-mode BatchNorm backward subtracts two projections; two fp32
    CPU runs with different MKL thread counts differ by up to 3 % in the deepest layers), so the parity tests measure every
    implementation against the fp64 result.
    Returns {"loss", "loss_mos", "loss_motion_encoder", "rpn_loss_cls", "rpn_loss_loc", "logits", "motion", "targets",
    "pred", "grads": {key: tensor}}."""
    from . import me
    params = {}
    work = {}
    me.FDTYPE, saved_dtype = dtype, me.FDTYPE
    try:
        return _train_step(sd, points, labels, gt_boxes, pred_override, use_motion_loss, dtype, params, work)
    finally:
        me.FDTYPE = saved_dtype


def _train_step(sd, points, labels, gt_boxes, pred_override, use_motion_loss, dtype, params, work):
    for k, v in sd.items():
        v = torch.as_tensor(v)
        if v.is_floating_point():
            v = v.to(dtype)
        leaf = k.rsplit(".", 1)[-1]
        if v.is_floating_point() and leaf not in ("running_mean", "running_var") and not k.endswith("MOSLoss.loss.weight"):
            v = v.detach().clone().requires_grad_(True)
            params[k] = v
        work[k] = v
    aux = {}
    work["__train__"] = aux
    wts = work["model.MOSLoss.loss.weight"]
    cur = graph.motionnet(work, points)
    motion = cur[:, 4:]
    loss_motion = mos_loss(motion, labels, wts)
    vox, coords, num, ids = sp.point_to_voxel(cur.detach().numpy(), graph.VOXEL, graph.PC_RANGE, 5, 100000)
    vf = sp.mean_vfe(vox, num)
    vc = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], axis=1)
    logits, pred = graph.unet(work, vf, vc, ids, None, pred_override)
    targets = center_targets(gt_boxes)
    cls_loss, loc_loss = rpn_loss(aux["cls_preds"], aux["box_preds"], targets)
    loss_mos = mos_loss(logits, labels, wts)
    loss = cls_loss + loc_loss + loss_mos + (loss_motion if use_motion_loss else 0.0)
    loss.backward()
    return {"loss": float(loss.detach()), "loss_mos": float(loss_mos.detach()), "loss_motion_encoder": float(loss_motion.detach()),
            "rpn_loss_cls": float(cls_loss.detach()), "rpn_loss_loc": float(loc_loss.detach()), "logits": logits.detach(), "motion": motion.detach(),
            "targets": targets, "pred": {k: (v.detach() if torch.is_tensor(v) else v) for k, v in pred.items()},
            "grads": {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}}
