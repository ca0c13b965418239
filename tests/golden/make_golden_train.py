"""Generate tests/golden/train_small.npz: ONE training step of the REFERENCE's own model code on CPU.

Run in the development container only (needs /root/reference):   python tests/golden/make_golden_train.py

What runs: /root/reference/models/models.py::InsMOSNet.forward(batch, 'train') -- the reference's training graph, its
CenterHead.assign_targets / get_loss (center_head.py:126-331), MOSLoss (loss.py:20-34), BatchNorm in train mode -- and
loss.backward(), over the same shims as make_golden.py (oracle restatement of MinkowskiEngine / spconv, now attached to
autograd; compiled reference Array_Index; C-oracle NMS standing in for the CUDA-only nms_gpu).
Two environment adaptations, neither touches the arithmetic:
  * center_head.py:150-160 builds `np.array(list of lists of tensors)`, which numpy >= 2 turns into one big ndarray
    instead of the object array the code expects: the module's `np` is wrapped so that exactly this call yields the
    object array numpy 1.x produced;
  * post_process.generate_recall_record (post_process.py:60-109) needs the CUDA-only boxes_iou3d_gpu and only fills the
    recall statistics (not the loss): replaced by a pass-through.
Stored: the five loss values, the logits of both heads, the target tensors, the detections used for instance fusion,
and for EVERY parameter the gradient's L2 norm, 4 seeded random projections and (tensors <= 4096 elements) the full
gradient; BatchNorm running statistics after the step for three layers.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

from insmos_b200 import synth, synth_weights  # noqa: E402

CASES = {
    # name: (synth kwargs, conv_cls bias).  bias 0 -> detections pass SCORE_THRESH and the instance-fusion path carries data
    "train_small": (dict(seed=11, n_scans=3, n_elev=32, n_azim=450), 0.0),
}
FULL_GRAD_MAX = 4096
N_PROJ = 4


def projection(key, shape, j):
    """seeded +-1 probe vector for parameter `key` (same on the test side)."""
    import zlib
    r = np.random.default_rng(zlib.crc32(("%s#%d" % (key, j)).encode()))
    return (r.integers(0, 2, size=int(np.prod(shape))).astype(np.float32) * 2 - 1).reshape(shape)


def adapt_reference():
    import models.backbones_2d.center_head as ch
    import models.post_process as pp

    class _NP:
        def __getattr__(self, n):
            return getattr(np, n)

        @staticmethod
        def array(x, *a, **k):
            if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], (list, tuple)) and len(x[0]) and torch.is_tensor(x[0][0]):
                o = np.empty((len(x), len(x[0])), dtype=object)
                for i, r in enumerate(x):
                    for j, v in enumerate(r):
                        o[i, j] = v
                return o
            return np.array(x, *a, **k)
    ch.np = _NP()
    pp.generate_recall_record = lambda box_preds, recall_dict, batch_index, data_dict=None, thresh_list=None: recall_dict


def run_case(net, kw, cls_bias):
    pts, labels, boxes = synth.make_sequence(return_labels=True, **kw)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth_weights.fill_state_dict(shapes)
    sd["model.unet.center_head.conv_cls.bias"] = torch.full((3,), float(cls_bias))
    net.load_state_dict(sd, strict=True)
    net.train()
    net.zero_grad(set_to_none=True)
    batch = [{"meta": None, "past_point_clouds": torch.from_numpy(pts.copy()),
              "past_labels": [torch.from_numpy(labels.astype(np.float32))],
              "gt_boxes": torch.from_numpy(boxes.copy()).unsqueeze(0), "batch_size_npast": kw["n_scans"]}]
    # the detections that gate the instance features: captured where the model computes them (a re-run after the step would
    # see cls_preds already passed through clip_sigmoid's in-place sigmoid_, center_head.py:343)
    import models.backbones_3d.spconv_unet as su
    captured = {}
    orig_pp = su.post_processing

    def capture(batch_dict, cfg, num_class):
        pd, rd = orig_pp(batch_dict, cfg, num_class)
        captured["pred"] = {k: v.detach().clone() for k, v in pd[0].items()}
        return pd, rd
    su.post_processing = capture
    try:
        loss, dicts, gt, pred = net.forward(batch, "train")
    finally:
        su.post_processing = orig_pp
    loss.backward()
    d = batch[0]
    head = net.model.unet.center_head
    out = {
        "loss": np.float64(loss.item()),
        "loss_mos": np.float64(dicts[0]["loss_mos"]), "loss_motion_encoder": np.float64(dicts[0]["loss_motion_encoder"]),
        "rpn_loss_cls": np.float64(dicts[0]["rpn_loss_cls"]), "rpn_loss_loc": np.float64(dicts[0]["rpn_loss_loc"]),
        "point_seg_feature": pred[0].detach().numpy(),                       # (column 0 is -inf: MOSLoss writes it in place)
        "current_motion_feature": d["current_motion_feature"].detach().numpy(),
        "heatmap_nonzero_index": np.flatnonzero(head.forward_ret_dict["heatmaps"][0].numpy()).astype(np.int64),
        "heatmap_nonzero_value": head.forward_ret_dict["heatmaps"][0].numpy().reshape(-1)[
            np.flatnonzero(head.forward_ret_dict["heatmaps"][0].numpy())],
        "heatmap_shape": np.asarray(head.forward_ret_dict["heatmaps"][0].shape, dtype=np.int64),
        "anno_boxes": head.forward_ret_dict["anno_boxes"][0].numpy(), "inds": head.forward_ret_dict["inds"][0].numpy(),
        "masks": head.forward_ret_dict["masks"][0].numpy(),
        "gt_boxes": boxes, "labels": labels,
    }
    out["pred_boxes"] = captured["pred"]["pred_boxes"].numpy()
    out["pred_scores"] = captured["pred"]["pred_scores"].numpy()
    out["pred_labels"] = captured["pred"]["pred_labels"].numpy()
    grads = {}
    for k, p in net.named_parameters():
        g = p.grad
        if g is None:
            grads["gnone:" + k] = np.int8(1)
            continue
        g = g.detach().numpy()
        grads["gnorm:" + k] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
        grads["gproj:" + k] = np.asarray([(g.astype(np.float64) * projection(k, g.shape, j)).sum() for j in range(N_PROJ)])
        if g.size <= FULL_GRAD_MAX:
            grads["gfull:" + k] = g
    bn_after = {}
    for k in ("model.motion_encoder.MinkUNet.bn0.bn", "model.unet.conv_input.1", "model.unet.bev_backbone.blocks.0.2"):
        bn_after["bn_after:" + k + ".running_mean"] = net.state_dict()[k + ".running_mean"].numpy().copy()
        bn_after["bn_after:" + k + ".running_var"] = net.state_dict()[k + ".running_var"].numpy().copy()
    return shapes, out, grads, bn_after


def main():
    net, cfg = mg.load_reference_model()
    adapt_reference()
    for name, (kw, cls_bias) in CASES.items():
        shapes, out, grads, bn_after = run_case(net, kw, cls_bias)
        meta = {"synth": kw, "cls_bias": cls_bias, "shapes": {k: list(v) for k, v in shapes.items()}, "n_proj": N_PROJ,
                "loss_config": cfg["MODEL"]["DENSE_HEAD"]["LOSS_CONFIG"], "target_config": cfg["MODEL"]["DENSE_HEAD"]["TARGET_ASSIGNER_CONFIG"]}
        path = os.path.join(HERE, "%s.npz" % name)
        np.savez_compressed(path, meta=json.dumps(meta), **{"out:" + k: v for k, v in out.items()}, **grads, **bn_after)
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()},
              "params with grad", sum(1 for k in grads if k.startswith("gnorm:")), "without", sum(1 for k in grads if k.startswith("gnone:")),
              "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
