"""'eval' mode of the model mirror (models/models.py:301-306,349-359): pairwise 3D IoU kernel vs the oracle (and, on a box that
holds it, vs the reference's own boxes_iou3d_gpu compiled into oracle/_ref), the recall record, the validation losses."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from insmos_b200 import ops  # noqa: E402
from oracle import graph  # noqa: E402
from oracle import train as otrain  # noqa: E402
from test_oracle_train import load_train_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _boxes(rng, n, spread):
    return np.stack([rng.uniform(-spread, spread, n), rng.uniform(-spread, spread, n), rng.uniform(-2, 0, n), rng.uniform(0.5, 6, n),
                     rng.uniform(0.5, 3, n), rng.uniform(0.5, 2.5, n), rng.uniform(-3.2, 3.2, n)], 1).astype(np.float32)


def test_boxes_iou3d_matches_oracle(cuda):
    rng = np.random.default_rng(3)
    a, b = _boxes(rng, 300, 12.0), _boxes(rng, 40, 12.0)
    b[:5] = a[:5]                                              # identical boxes: IoU 1
    b[5:10, :2] = a[5:10, :2] + 0.3                            # heavy overlaps
    ref = graph.boxes_iou3d(a, b)
    got = ops.boxes_iou3d(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)).cpu().numpy()
    assert (ref > 0).sum() > 50 and np.abs(np.diag(ref[:5, :5]) - 1).max() < 1e-4
    # device libm sin/cos/atan2 vs glibc: last-bit differences of the corner coordinates
    assert np.abs(got - ref).max() < 2e-5
    assert ops.boxes_iou3d(torch.zeros((0, 7), device=cuda), torch.from_numpy(b).to(cuda)).shape == (0, 40)


def test_boxes_overlap_matches_reference_cuda_kernel(cuda):
    """the reference's own boxes_overlap_bev_gpu (compiled from /root/reference into oracle/_ref) on the same boxes"""
    from oracle import native
    ref = native.ref_iou3d()
    if ref is None or not hasattr(ref, "boxes_overlap_bev_gpu"):
        pytest.skip("oracle/_ref/iou3d_nms_cuda not built")
    rng = np.random.default_rng(5)
    a, b = _boxes(rng, 200, 8.0), _boxes(rng, 50, 8.0)
    at, bt = torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)
    out = torch.zeros((200, 50), device=cuda)
    ref.boxes_overlap_bev_gpu(at, bt, out)
    # IoU3D assembled from the reference's BEV overlap exactly as iou3d_nms_utils.py:41-59 does
    amax, amin = (at[:, 2] + at[:, 5] / 2).view(-1, 1), (at[:, 2] - at[:, 5] / 2).view(-1, 1)
    bmax, bmin = (bt[:, 2] + bt[:, 5] / 2).view(1, -1), (bt[:, 2] - bt[:, 5] / 2).view(1, -1)
    o3d = out * torch.clamp(torch.min(amax, bmax) - torch.max(amin, bmin), min=0)
    va, vb = (at[:, 3] * at[:, 4] * at[:, 5]).view(-1, 1), (bt[:, 3] * bt[:, 4] * bt[:, 5]).view(1, -1)
    ref_iou = o3d / torch.clamp(va + vb - o3d, min=1e-6)
    got = ops.boxes_iou3d(at, bt)
    assert float((got - ref_iou).abs().max()) < 1e-6 and int((ref_iou > 0).sum()) > 30


def test_eval_mode_losses_and_recall(cuda):
    from test_gpu_train import _batch, _train_net
    g, meta, sd, pts, labels, boxes = load_train_golden()
    net = _train_net(cuda, sd).eval()
    with torch.no_grad():
        out = net.forward(_batch(cuda, pts, labels, boxes), "eval")
    preb, recall, gts, preds, val_loss, val_motion = out
    assert len(preb) == len(recall) == 1 and isinstance(val_loss, float)
    # validation losses = MOSLoss of the two heads' logits (eval-mode BatchNorm: running statistics of the fresh weights)
    w = sd["model.MOSLoss.loss.weight"]
    ref_motion = otrain.mos_loss(_cur_motion(net, cuda, pts), labels, w)
    assert abs(float(val_motion) - float(ref_motion)) < 1e-5 * max(1.0, abs(float(ref_motion)))
    ref_val = otrain.mos_loss(preds[0].cpu(), labels, w)
    assert abs(val_loss - float(ref_val)) < 1e-5 * max(1.0, abs(float(ref_val)))
    # recall record vs the oracle on the SAME predicted boxes
    ref_rec = graph.recall_record(preb[0][0]["pred_boxes"].cpu().numpy(), boxes, (0.3, 0.5, 0.7))
    assert recall[0] == ref_rec, (recall[0], ref_rec)
    assert recall[0]["gt"] == 3
    # 'test' mode carries no recall record, unknown modes are refused
    with torch.no_grad():
        assert net.forward(_batch(cuda, pts, labels, boxes), "test")[1] == [{}]
    with pytest.raises(ValueError):
        net.forward([], "validate")


def _cur_motion(net, cuda, pts):
    with torch.no_grad():
        d = net.model.motion_encoder({"past_point_clouds": torch.from_numpy(pts).to(cuda)})
    return d["current_motion_feature"].cpu().clone()
