"""Dead-row elimination in the 4D MotionNet decoder: only the rows of the newest scan leave MotionNet,
so the last decoder layers compute the rows whose time index their consumers can reach.  The outputs the model reads must be
BIT-IDENTICAL with and without it; the row bounds are checked against numpy, including unordered input."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util  # noqa: E402
from insmos_b200 import ops, synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _starts_ref(t, n):
    out = np.full(16, n, dtype=np.int64)
    for j in range(16):
        idx = np.flatnonzero(t >= -j)
        if len(idx):
            out[j] = idx[0]
    return out


@pytest.mark.parametrize("shuffle", [False, True])
def test_time_row_starts_match_numpy(cuda, shuffle):
    pts = synth.make_sequence(seed=5, n_scans=6, n_elev=16, n_azim=300)
    if shuffle:
        pts = pts[np.random.default_rng(0).permutation(len(pts))]
    cs, _, _ = ops.voxelize4d(torch.from_numpy(pts).to(cuda), [0.1, 0.1, 0.1, 0.1])
    got = ops.time_row_starts(cs)[:16].cpu().numpy()
    ref = _starts_ref(cs.coords[:, 4].cpu().numpy(), cs.n)
    assert np.array_equal(got, ref)
    if not shuffle:                                    # time-ordered input: the needed rows are exactly a suffix
        t = cs.coords[:, 4].cpu().numpy()
        assert np.all(np.diff(t) >= 0) and got[0] == np.flatnonzero(t == 0)[0]
    empty = ops.time_row_starts(ops.CoordSet(torch.zeros((0, 5), dtype=torch.int32, device=cuda), 0, cs.table, cs.cap))
    assert np.array_equal(empty[:16].cpu().numpy(), np.zeros(16))


def _motion(cuda, sd, pts, prune):
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    net = InsMOSNet(default_config())
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda).eval()
    old = ops.USE_TPRUNE
    ops.USE_TPRUNE = prune
    try:
        with torch.no_grad():
            d = net.model.motion_encoder({"past_point_clouds": torch.from_numpy(pts).to(cuda)})
            calls = None
        return d["current_point"].cpu()
    finally:
        ops.USE_TPRUNE = old


@pytest.mark.parametrize("case,shuffle", [("small", False), ("small", True), ("small_nodet", False)])
def test_motionnet_output_is_bit_identical_with_dead_row_elimination(cuda, case, shuffle):
    meta, shapes, sd, pts, gold = golden_util.load(case)
    if shuffle:                                        # unordered points: bounds stay valid (fewer rows are skipped)
        pts = pts[np.random.default_rng(1).permutation(len(pts))]
    full = _motion(cuda, sd, pts, prune=False)
    pruned = _motion(cuda, sd, pts, prune=True)
    assert torch.equal(full, pruned)
    if not shuffle:
        assert (pruned - torch.from_numpy(gold["current_point"])).abs().max() < 1e-3


def test_full_size_forward_is_bit_identical_and_skips_work(cuda):
    """C2-size cloud: same logits and boxes; the pruned forward launches the same kernels on fewer tiles (fewer pairs built)"""
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    from insmos_b200 import _lib
    meta, shapes, sd, _, _ = golden_util.load("c2", with_points=False)
    net = InsMOSNet(default_config())
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda).eval()
    pts = torch.from_numpy(synth.make_sequence(seed=3, n_scans=10, n_elev=64, n_azim=1875)).to(cuda)
    res = {}
    for prune in (False, True):
        old, ops.USE_TPRUNE = ops.USE_TPRUNE, prune
        try:
            with torch.no_grad():
                _lib.profile_start()
                boxes, _, logits = net.forward([{"meta": None, "past_point_clouds": pts, "batch_size_npast": 10}], "test")
                prof = _lib.profile_stop()
        finally:
            ops.USE_TPRUNE = old
        pairs = sum(m.get("pairs", 0) for n, t, m in prof if n.startswith("insmos_rulebook_build") and m)
        res[prune] = (logits[0].cpu(), boxes[0][0]["pred_boxes"].cpu(), pairs)
    assert torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1])
    assert res[True][2] < 0.8 * res[False][2], (res[True][2], res[False][2])
