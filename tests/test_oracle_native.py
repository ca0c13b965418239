"""Pin the C restatement (oracle/native/oracle_native.c) against the reference's own compiled native
code (oracle/_ref, built by oracle/build.py from /root/reference) -- bit for bit."""
import numpy as np
import pytest
import torch

from oracle import native


def _boxes(rng, n, spread=20.0):
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.uniform(-2, 0, n)
    b[:, 3:6] = np.exp(rng.normal(0.5, 0.5, (n, 3)))
    b[:, 6] = rng.uniform(-3.2, 3.2, n)
    return b


def test_iou_matrix_matches_reference_cpu_twin():
    ref = native.ref_iou3d()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(0)
    for seed_spread in (20.0, 5.0):
        a, b = _boxes(rng, 200, seed_spread), _boxes(rng, 250, seed_spread)
        out = torch.zeros(200, 250)
        ref.boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), out)
        mine = native.iou_matrix(a, b)
        assert (mine > 0).sum() > 100
        assert np.array_equal(mine, out.numpy())
    # identical and axis-aligned boxes: known answers
    sq = np.array([[0, 0, 0, 2, 2, 1, 0]], np.float32)
    assert abs(native.iou_matrix(sq, sq)[0, 0] - 1.0) < 1e-2          # 1e-2 corner margin inflates slightly
    half = np.array([[1, 0, 0, 2, 2, 1, 0]], np.float32)
    assert abs(native.iou_matrix(sq, half)[0, 0] - 1.0 / 3.0) < 2e-2


def test_overlap_and_iou3d_are_consistent_with_the_pinned_bev_iou():
    """oracle.graph.boxes_iou3d (recall record of the 'eval' mode) rests on the same box_overlap the pinned BEV IoU uses"""
    from oracle import graph
    rng = np.random.default_rng(2)
    a, b = _boxes(rng, 60, 6.0), _boxes(rng, 70, 6.0)
    ov, iou = native.overlap_matrix(a, b), native.iou_matrix(a, b)
    sa, sb = (a[:, 3] * a[:, 4])[:, None], (b[:, 3] * b[:, 4])[None, :]
    assert np.array_equal(iou, ov / np.maximum(sa + sb - ov, np.float32(1e-8)))
    cube = np.array([[0, 0, 0, 2, 2, 2, 0]], np.float32)
    up = np.array([[0, 0, 1, 2, 2, 2, 0]], np.float32)                    # shifted by half its height: IoU3D = 1/3
    assert abs(graph.boxes_iou3d(cube, up)[0, 0] - 1.0 / 3.0) < 2e-2
    assert graph.boxes_iou3d(cube, np.array([[0, 0, 5, 2, 2, 2, 0]], np.float32))[0, 0] == 0.0
    rec = graph.recall_record(np.concatenate([cube, up]), np.array([[0, 0, 0, 2, 2, 2, 0, 1], [9, 9, 0, 2, 2, 2, 0, 2], [0] * 8], np.float32))
    assert rec == {"gt": 2, "roi_0.3": 0, "rcnn_0.3": 1, "roi_0.5": 0, "rcnn_0.5": 1, "roi_0.7": 0, "rcnn_0.7": 1}


def test_array_index_matches_reference():
    ai = native.ref_array_index()
    if ai is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(1)
    for n, nb in ((20000, 60), (3000, 500), (50, 0)):
        vox = rng.integers(0, 150, (n, 3)).astype(np.int32)
        vox[:, 2] = rng.integers(0, 6, n)
        bx = np.zeros((nb, 8), np.float32)
        bx[:, 0:2] = rng.uniform(0, 150, (nb, 2)); bx[:, 2] = rng.uniform(0, 6, nb)
        bx[:, 3:6] = np.exp(rng.normal(1.5, 0.6, (nb, 3))); bx[:, 6] = rng.uniform(-3.2, 3.2, nb)
        bx[:, 7] = rng.integers(0, 4, nb)
        r = ai.find_features_by_bbox_with_yaw(vox, bx, np.zeros((n, 3), dtype=np.int32))
        m = native.find_features_by_bbox_with_yaw(vox, bx)
        assert np.array_equal(np.asarray(r), m)
        if nb:
            assert m.sum() > 0


def test_first_hit_pruning_quirk_is_reproduced():
    # a long thin rotated box: voxels inside the box but outside the axis-aligned +-extent window of
    # the first hit are NOT marked (Array_Index.cpp:48-51)
    vox = np.array([[0, 0, 0], [9, 9, 0], [1, 1, 0]], np.int32)
    box = np.array([[5, 5, 0, 16, 1.5, 2, np.pi / 4, 1]], np.float32)
    m = native.find_features_by_bbox_with_yaw(vox, box)
    assert m[:, 0].tolist() == [1, 0, 1]
    ai = native.ref_array_index()
    if ai is not None:
        assert np.array_equal(np.asarray(ai.find_features_by_bbox_with_yaw(vox, box, np.zeros((3, 3), dtype=np.int32))), m)


def test_nms_greedy_properties():
    rng = np.random.default_rng(2)
    b = _boxes(rng, 700, 8.0)
    keep = native.nms(b, 0.01)
    iou = native.iou_matrix(b, b)
    kept = set(keep.tolist())
    assert keep[0] == 0 and np.all(np.diff(keep) > 0)
    for i in range(len(b)):
        suppressed_by = [j for j in keep if j < i and iou[j, i] > 0.01]
        assert (i in kept) == (len(suppressed_by) == 0)
