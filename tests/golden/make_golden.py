"""Generate tests/golden/insmos_small.npz by running the REFERENCE's own model code on CPU.

Run in the development container only (needs /root/reference):   python tests/golden/make_golden.py

What runs: /root/reference/models/models.py::InsMOSNet (the reference's graph, unmodified) over
  * oracle/shims/{MinkowskiEngine,spconv,pytorch_lightning,easydict}  (CPU restatement of the external libs),
  * oracle/_ref/Array_Index*.so  -- the reference's own Array_Index.cpp, compiled as is,
  * the C oracle's nms (pinned bit-exact against the reference's iou3d code) standing in for the
    CUDA-only iou3d_nms_cuda.nms_gpu; Tensor.cuda() is made a no-op because the reference hard-codes it.
Weights: tests/synth_weights.py (seeded per key) + BatchNorm statistics calibrated on this input.
The fixture pins the model GRAPH (wiring, constants, op order) for oracle/graph.py and the CUDA path.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), REF, ROOT, os.path.join(ROOT, "tests")]

from oracle import build, native  # noqa: E402
from insmos_b200 import synth_weights  # noqa: E402
from insmos_b200 import synth  # noqa: E402

CASES = {
    # name: (synth kwargs, conv_cls bias)      bias 0 -> many detections; stock bias -log(99) -> none
    "small": (dict(seed=11, n_scans=3, n_elev=32, n_azim=450), 0.0),
    "small_nodet": (dict(seed=12, n_scans=2, n_elev=16, n_azim=300), -float(np.log(99.0))),
}


def load_reference_model():
    build.build_native()
    ai_path, _ = build.build_ref()
    torch.Tensor.cuda = lambda self, *a, **k: self                       # reference hard-codes .cuda() (iou3d_nms_utils.py:79)
    fake = types.ModuleType("models.bbox_post_process.iou3d_nms_cuda")

    def nms_gpu(boxes, keep, thresh):
        k = native.nms(boxes.detach().cpu().numpy(), float(thresh))
        keep[:len(k)] = torch.from_numpy(k)
        return len(k)
    fake.nms_gpu = nms_gpu
    sys.modules["models.bbox_post_process.iou3d_nms_cuda"] = fake
    import models                                                        # noqa: F401  (the reference package)
    import models.utils                                                  # namespace package: add the compiled module
    models.utils.__path__.append(os.path.dirname(ai_path))
    import models.bbox_post_process
    models.bbox_post_process.iou3d_nms_cuda = fake
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        cfg = yaml.safe_load(open(os.path.join(REF, "config", "config.yaml")))
        import models.models as mm
        net = mm.InsMOSNet(cfg)
    finally:
        os.chdir(cwd)
    return net, cfg


def run_case(net, pts, cls_bias):
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth_weights.fill_state_dict(shapes)
    sd["model.unet.center_head.conv_cls.bias"] = torch.full((3,), float(cls_bias))
    net.load_state_dict(sd, strict=True)
    batch = lambda: [{"meta": None, "past_point_clouds": torch.from_numpy(pts.copy()), "batch_size_npast": 0}]   # noqa: E731
    # BatchNorm calibration: one pass with batch statistics written into the running buffers
    bns = [m for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    for m in bns:
        m.momentum = 1.0
        m.train()
    with torch.no_grad():
        net.forward(batch(), "test")
    net.eval()
    inter = {}
    unet = net.model.unet
    hook = unet.center_head.register_forward_hook(lambda mod, inp, out: inter.update(
        n_cand=int((torch.sigmoid(out["batch_cls_preds"]).max(-1)[0] >= 0.1).sum()),
        all_scores=torch.sigmoid(out["batch_cls_preds"][0]).max(-1)[0].numpy().copy(),
        all_boxes=out["batch_box_preds"][0].numpy().copy()))
    b = batch()
    with torch.no_grad():
        boxes, recall, logits = net.forward(b, "test")
    hook.remove()
    sd2 = net.state_dict()
    d = b[0]
    return shapes, sd2, {
        "logits": logits[0].numpy(), "pred_boxes": boxes[0][0]["pred_boxes"].numpy(),
        "pred_scores": boxes[0][0]["pred_scores"].numpy(), "pred_labels": boxes[0][0]["pred_labels"].numpy(),
        "current_point": d["current_point"].numpy(), "voxel_coords": d["voxel_coords"].numpy().astype(np.int32),
        "voxel_features": d["voxel_features"].numpy(), "pc_voxel_id": d["pc_voxel_id"].numpy(),
        "n_cand": np.int64(inter["n_cand"]), "all_scores": inter["all_scores"], "all_boxes": inter["all_boxes"],
    }


def main():
    net, cfg = load_reference_model()
    for name, (kw, cls_bias) in CASES.items():
        pts = synth.make_sequence(**kw)
        shapes, sd, out = run_case(net, pts, cls_bias)
        bn = {k: sd[k].numpy() for k in synth_weights.bn_stat_keys(shapes)}
        meta = {"synth": kw, "cls_bias": cls_bias, "shapes": {k: list(v) for k, v in shapes.items()},
                "points_sha": int(np.abs(pts).sum() * 1000) % (1 << 31), "config": cfg["MODEL"]["POST_PROCESSING"]}
        path = os.path.join(HERE, "insmos_%s.npz" % name)
        np.savez_compressed(path, meta=json.dumps(meta), **{"bn:" + k: v for k, v in bn.items()},
                            **{"out:" + k: v for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()}, "boxes", len(out["pred_boxes"]), "cand", int(out["n_cand"]),
              "logit absmax %.3f" % np.abs(out["logits"]).max(), "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
