"""End-to-end parity of the CUDA path (models.models.InsMOSNet mirror over libinsmos_b200.so):
  * vs the golden fixtures produced by the reference's own model code (tests/golden/make_golden.py),
  * vs oracle/graph.py on a fresh seeded input.
Gates (BASELINE.json north_star): voxel indices bit-exact, MOS logits within 1e-3 abs fp32."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util  # noqa: E402
from insmos_b200 import synth_weights  # noqa: E402
from oracle import graph  # noqa: E402

pytestmark = pytest.mark.gpu
LOGIT_ATOL = 1e-3


def _net(cuda, sd):
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    net = InsMOSNet(default_config())
    net.load_state_dict(sd, strict=True)
    return net.to(cuda).eval()


def _run(net, pts, cuda):
    batch = [{"meta": None, "past_point_clouds": torch.from_numpy(pts).to(cuda), "batch_size_npast": 0}]
    with torch.no_grad():
        boxes, recall, logits = net.forward(batch, "test")
    return batch[0], boxes[0][0], logits[0]


def _compare(d, pred, logits, ref):
    assert np.array_equal(d["voxel_coords"].cpu().numpy(), np.asarray(ref["voxel_coords"])), "3D voxel indices differ"
    assert np.array_equal(d["pc_voxel_id"].cpu().numpy().astype(np.int64), np.asarray(ref["pc_voxel_id"]))
    cp = torch.as_tensor(ref["current_point"])
    err_m = (d["current_point"].cpu() - cp).abs().max().item()
    assert err_m < LOGIT_ATOL, "motion logits differ by %.3e" % err_m
    rb, rl = torch.as_tensor(ref["pred_boxes"]), torch.as_tensor(ref["pred_labels"])
    assert pred["pred_boxes"].shape == rb.shape, "box count %s vs %s" % (tuple(pred["pred_boxes"].shape), tuple(rb.shape))
    assert torch.allclose(pred["pred_boxes"].cpu(), rb, atol=1e-3), (pred["pred_boxes"].cpu() - rb).abs().max()
    assert torch.equal(pred["pred_labels"].cpu(), rl)
    rlog = torch.as_tensor(ref["logits"])
    err = (logits.cpu() - rlog).abs().max().item()
    assert logits.shape == rlog.shape
    assert err < LOGIT_ATOL, "MOS logits differ by %.3e (abs max of reference %.2f)" % (err, rlog.abs().max())
    return err


@pytest.mark.parametrize("name", ["small_nodet", "small"])
def test_cuda_forward_matches_reference_golden(cuda, name):
    meta, shapes, sd, pts, gold = golden_util.load(name)
    net = _net(cuda, sd)
    d, pred, logits = _run(net, pts, cuda)
    _compare(d, pred, logits, gold)


def test_cuda_forward_matches_oracle_graph_fresh_input(cuda):
    """a different cloud than the fixtures (weights: fixture BN statistics), checked against oracle/graph.py"""
    meta, shapes, sd, _, _ = golden_util.load("small")
    from insmos_b200 import synth
    pts, labels, _ = synth.make_sequence(seed=21, n_scans=4, n_elev=32, n_azim=500, return_labels=True)
    ref = graph.forward(sd, pts)
    net = _net(cuda, sd)
    d, pred, logits = _run(net, pts, cuda)
    _compare(d, pred, logits, ref)
    # MOS IoU (models/metrics.py:16-44 arithmetic) equal within 1e-4
    from insmos_b200.net.model import ClassificationMetrics
    cm = ClassificationMetrics(3, [0])
    lab = torch.from_numpy(labels)
    iou_gpu = cm.getIoU(cm.compute_confusion_matrix(logits.cpu(), lab).float())[2].item()
    iou_ref = cm.getIoU(cm.compute_confusion_matrix(ref["logits"], lab).float())[2].item()
    assert abs(iou_gpu - iou_ref) <= 1e-4


def test_sparse_conv_algorithms_agree_in_model(cuda):
    """same forward with the SIMT fp32 kernels forced == tensor-core 3xTF32 path within 1e-3"""
    meta, shapes, sd, pts, gold = golden_util.load("small_nodet")
    net = _net(cuda, sd)
    from insmos_b200 import ops
    _, _, logits_auto = _run(net, pts, cuda)
    orig = ops.sparse_conv
    try:
        ops.sparse_conv = lambda *a, **k: orig(*a, **{**k, "algo": 1})
        import MinkowskiEngine, spconv.pytorch.conv as spc
        _, _, logits_simt = _run(net, pts, cuda)
    finally:
        ops.sparse_conv = orig
    assert (logits_auto - logits_simt).abs().max().item() < LOGIT_ATOL


def test_reference_op_sequence_through_shims(cuda):
    """the reference-style UNFUSED op sequence (conv -> MinkowskiBatchNorm -> MinkowskiReLU as separate modules,
    the way models/MinkowskiEngine/minkunet.py calls them) gives the same result as the fused epilogue."""
    import insmos_b200
    insmos_b200.install()
    import MinkowskiEngine as ME
    g = torch.Generator().manual_seed(0)
    coords = torch.randint(-20, 20, (4000, 4), generator=g).float() + 0.5
    c, f = ME.utils.sparse_collate([coords], [torch.full((4000, 1), 0.5)])
    field = ME.TensorField(features=f.to(cuda), coordinates=c.float().to(cuda))
    st = field.sparse()
    conv = ME.MinkowskiConvolution(1, 8, kernel_size=[3, 3, 3, 3], dimension=4).to(cuda)
    bn = ME.MinkowskiBatchNorm(8).to(cuda).eval()
    with torch.no_grad():
        bn.bn.running_mean.normal_(generator=None); bn.bn.running_var.uniform_(0.5, 2.0); bn.bn.weight.uniform_(0.5, 1.5)
        unfused = ME.MinkowskiReLU()(bn(conv(st)))
        fused = conv(st, bn=bn, relu=True)
    assert torch.allclose(unfused.F, fused.F, atol=1e-5, rtol=1e-5)
    sl = fused.slice(field)
    assert sl.F.shape == (4000, 8) and torch.equal(sl.F, fused.F[field.inverse_mapping.long()])
    assert ME.cat(fused, unfused).F.shape[1] == 16
