"""CUDA detection-head kernels vs the oracle AND vs the reference's own CUDA NMS (oracle/_ref)."""
import numpy as np
import pytest
import torch

from insmos_b200 import ops
from oracle import native

pytestmark = pytest.mark.gpu


def _boxes(rng, n, spread):
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.uniform(-2, 0, n)
    b[:, 3:6] = np.exp(rng.normal(0.5, 0.5, (n, 3)))
    b[:, 6] = rng.uniform(-3.2, 3.2, n)
    return b


def _keep_equal_or_explained(keep_gpu, keep_ref, boxes, thr):
    """exact equality expected; if the device libm rounds a corner differently the first divergence must
    sit on an IoU within 1e-5 of the threshold."""
    if np.array_equal(keep_gpu, keep_ref):
        return True
    m = min(len(keep_gpu), len(keep_ref))
    d = np.nonzero(keep_gpu[:m] != keep_ref[:m])[0]
    i = int(min(keep_gpu[d[0]], keep_ref[d[0]])) if len(d) else int(max(keep_gpu[m:].tolist() + keep_ref[m:].tolist()))
    iou = native.iou_matrix(boxes[:i], boxes[i:i + 1])[:, 0]
    return bool(np.any(np.abs(iou - thr) < 1e-5))


@pytest.mark.parametrize("n,spread", [(1, 5.0), (63, 5.0), (64, 5.0), (65, 5.0), (700, 8.0), (4096, 40.0)])
def test_nms_matches_oracle(cuda, n, spread):
    rng = np.random.default_rng(n)
    b = _boxes(rng, n, spread)
    ref = native.nms(b, 0.01)
    got = ops.nms_rotated(torch.from_numpy(b).to(cuda), 0.01, 100000).cpu().numpy().astype(np.int64)
    assert _keep_equal_or_explained(got, ref, b, 0.01), (got[:20], ref[:20])
    got500 = ops.nms_rotated(torch.from_numpy(b).to(cuda), 0.01, 500).cpu().numpy().astype(np.int64)
    assert np.array_equal(got500, got[:500])
    empty = ops.nms_rotated(torch.zeros((0, 7), device=cuda), 0.01, 500)
    assert empty.numel() == 0


@pytest.mark.parametrize("n,spread", [(5, 1.0), (64, 3.0), (1000, 6.0), (4096, 40.0), (4096, 4.0), (600, 0.5)])
def test_nms_pair_list_equals_dense_mask(cuda, n, spread):
    """the pair-list build (default) and the dense upper-triangle build give the same keep list bit for bit; spread 0.5 /
    4.0 make almost every pair overlap, which overflows the list and exercises the on-device fallback."""
    rng = np.random.default_rng(1000 + n)
    b = torch.from_numpy(_boxes(rng, n, spread)).to(cuda)
    for max_keep in (500, 100000):
        a = ops.nms_rotated(b, 0.01, max_keep, dense=True)
        c = ops.nms_rotated(b, 0.01, max_keep, dense=False)
        assert torch.equal(a, c), (n, spread, max_keep, a[:10], c[:10])


def test_nms_matches_reference_cuda_kernel(cuda):
    ref = native.ref_iou3d()
    if ref is None:
        pytest.skip("oracle/_ref/iou3d_nms_cuda not built")
    rng = np.random.default_rng(7)
    for n, spread in ((900, 10.0), (3000, 30.0)):
        b = _boxes(rng, n, spread)
        bt = torch.from_numpy(b).to(cuda)
        keep = torch.zeros(n, dtype=torch.long)
        num = ref.nms_gpu(bt, keep, 0.01)                       # the reference's own kernel + host sweep
        got = ops.nms_rotated(bt, 0.01, 100000).cpu().numpy().astype(np.int64)
        assert _keep_equal_or_explained(got, keep[:num].numpy(), b, 0.01)


def test_box_membership_matches_oracle(cuda):
    rng = np.random.default_rng(3)
    for n, nb, mult in ((20000, 60, 1.0), (3000, 500, 2.0), (100000, 300, 8.0), (10, 0, 1.0)):
        zyx = np.stack([rng.integers(0, int(6 * mult), n), rng.integers(0, int(125 * mult), n),
                        rng.integers(0, int(150 * mult), n)], axis=1)
        coords = np.concatenate([np.zeros((n, 1), np.int64), zyx], axis=1).astype(np.int32)
        bx = np.zeros((nb, 8), np.float32)
        bx[:, 0] = rng.uniform(0, 150, nb); bx[:, 1] = rng.uniform(0, 125, nb); bx[:, 2] = rng.uniform(0, 6, nb)
        bx[:, 3:6] = np.exp(rng.normal(1.2, 0.6, (nb, 3))); bx[:, 6] = rng.uniform(-3.2, 3.2, nb)
        bx[:, 7] = rng.integers(0, 4, nb)
        scaled = bx.copy(); scaled[:, 0:6] *= mult
        exp = native.find_features_by_bbox_with_yaw(coords[:, [3, 2, 1]], scaled)
        got = ops.box_membership(torch.from_numpy(coords).to(cuda), torch.from_numpy(bx).to(cuda), mult).cpu().numpy()
        diff = int((got.astype(np.int32) != exp).sum())
        assert diff <= max(1, exp.sum() // 5000), "membership differs in %d cells of %d hits" % (diff, exp.sum())
        if nb:
            assert exp.sum() > 0


def test_boxes_to_voxel_units_and_decode(cuda):
    rng = np.random.default_rng(4)
    b7 = _boxes(rng, 300, 40.0)
    lab = rng.integers(1, 4, 300).astype(np.int32)
    got = ops.boxes_to_voxel_units(torch.from_numpy(b7).to(cuda), torch.from_numpy(lab).to(cuda), [-60, -50, -3],
                                   [0.1, 0.1, 0.1], 8).cpu()
    t = torch.from_numpy(b7).clone()
    for d, lo in enumerate((-60, -50, -3)):                      # spconv_unet.py:324-329, same op order
        t[:, d] = (t[:, d] - lo) / 0.1 / 8
        t[:, 3 + d] = t[:, 3 + d] / 0.1 / 8
    assert torch.equal(got[:, :7], t) and torch.equal(got[:, 7], torch.from_numpy(lab).float())
    H, W = 50, 60
    g = torch.Generator().manual_seed(1)
    cls, box = torch.randn((3, H, W), generator=g), torch.randn((8, H, W), generator=g) * 0.5
    boxes, scores, labels = ops.center_decode(cls.to(cuda), box.to(cuda), 4, 0.1, 0.1, -60, -50)
    bp = box.permute(1, 2, 0).reshape(1, H * W, 8)                # center_head.py:251-276 on CPU
    ys, xs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing="ij")
    xs = xs.reshape(1, -1, 1) + bp[:, :, 0:1]
    ys = ys.reshape(1, -1, 1) + bp[:, :, 1:2]
    xs = xs * 4 * 0.1 + (-60)
    ys = ys * 4 * 0.1 + (-50)
    ref = torch.cat([xs, ys, bp[..., 2:3], torch.exp(bp[..., 3:6]), torch.atan2(bp[..., 6:7], bp[..., 7:8])], dim=2)[0]
    err = (boxes.cpu() - ref).abs().max(dim=0)[0]
    # positions/heights: same fp32 op order -> exact; exp / atan2: device libm within a few ulp
    assert torch.equal(boxes.cpu()[:, :3], ref[:, :3]), err
    assert torch.allclose(boxes.cpu()[:, 3:], ref[:, 3:], rtol=2e-6, atol=2e-6), err
    sc, lb = torch.max(torch.sigmoid(cls.permute(1, 2, 0).reshape(-1, 3)), dim=-1)
    assert torch.allclose(scores.cpu(), sc, rtol=1e-6, atol=1e-7)
    assert torch.equal(labels.cpu().long(), lb + 1)
