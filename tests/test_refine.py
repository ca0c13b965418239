"""N4: instance-level refinement after the forward path, scripts/refine.py:169-302.
  * the C restatement of Array_Index.find_point_in_instance_bbox_with_yaw vs the reference's own compiled module (bit exact),
  * oracle/refine.py vs the golden written by the reference's own scripts/refine.py (tests/golden/make_golden_refine.py),
  * (GPU) the device kernels vs the C oracle and insmos_b200.refine.InstanceRefiner vs the same golden -- labels exact."""
import os

import numpy as np
import pytest
import torch

from oracle import native, refine as oref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refine_small.npz")


def _gold():
    z = np.load(GOLD)
    F = int(z["n_frames"])
    frames = [{k: z["%s%d" % (k, f)] for k in ("scan", "boxes", "labels", "mos", "conf", "out")} for f in range(F)]
    return z["poses"], frames


def _to12(mos):
    m = mos.astype(np.int32).copy()
    m[m == 251] = 2
    m[m == 9] = 1
    return m


def _back(lab):
    out = lab.astype(np.int32).copy()
    out[lab == 1] = 9
    out[lab == 2] = 251
    return out


def _random_case(seed, n=6000, nb=40):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 1, (n, 1)), rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    boxes = np.zeros((nb, 8), np.float32)
    boxes[:, 0:2] = rng.uniform(-18, 18, (nb, 2))
    boxes[:, 2] = rng.uniform(-1.5, 0.5, nb)
    boxes[:, 3:6] = np.exp(rng.normal(0.9, 0.4, (nb, 3)))
    boxes[:, 6] = rng.uniform(-3.2, 3.2, nb)
    boxes[:, 7] = rng.integers(0, 4, nb)                     # label 0 = ignored box
    return pts, boxes


def test_c_oracle_matches_the_compiled_reference():
    ai = native.ref_array_index()
    if ai is None:
        pytest.skip("oracle/_ref/Array_Index not built (reference absent)")
    rng = np.random.default_rng(3)
    n, nb = 20000, 12
    pts = np.concatenate([rng.uniform(-20, 20, (n, 2)), rng.uniform(-2, 1, (n, 1)), rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    boxes = np.zeros((nb, 8), np.float32)
    boxes[:, 0] = np.arange(nb) * 3.3 - 19                   # disjoint boxes: the reference's OpenMP loop over boxes cannot race
    boxes[:, 1] = rng.uniform(-15, 15, nb)
    boxes[:, 2] = rng.uniform(-1, 0, nb)
    boxes[:, 3:6] = [2.8, 2.0, 1.8]
    boxes[:, 6] = rng.uniform(-3, 3, nb)
    boxes[:, 7] = rng.integers(0, 4, nb)
    ref = ai.find_point_in_instance_bbox_with_yaw(pts, boxes, np.zeros((n, 3), dtype=np.int32), 0.03)
    got = native.find_point_in_instance_bbox_with_yaw(pts, boxes, 0.03)
    assert np.array_equal(ref, got) and (got > 0).sum() > 100, int((got > 0).sum())


def test_oracle_refine_matches_reference_script_golden():
    poses, frames = _gold()
    r = oref.Refiner()
    changed = 0
    for f, fr in enumerate(frames):
        got = r.step(f, fr["scan"], fr["boxes"], fr["labels"], _to12(fr["mos"]), fr["conf"], poses)
        assert np.array_equal(_back(got), fr["out"]), "frame %d: %d labels differ" % (f, int((_back(got) != fr["out"]).sum()))
        changed += int((fr["out"] != fr["mos"].astype(np.int32)).sum())
    assert changed > 1000                                    # the fixture exercises the relabelling branches


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_point_instance_kernels_match_c_oracle(cuda, seed):
    from insmos_b200 import refine
    pts, boxes = _random_case(seed)
    want = native.find_point_in_instance_bbox_with_yaw(pts, boxes, 0.03)
    ids = refine.point_instance_ids(torch.from_numpy(pts).to(cuda), torch.from_numpy(boxes).to(cuda))
    assert np.array_equal(ids.cpu().numpy(), want), int((ids.cpu().numpy() != want).sum())
    rng = np.random.default_rng(seed)
    lab = rng.integers(1, 3, len(pts)).astype(np.int32)
    conf = rng.uniform(0, 1, (len(pts), 2)).astype(np.float32) * (rng.uniform(0, 1, (len(pts), 1)) < 0.5)
    st = refine.instance_stats(ids, 0, torch.from_numpy(lab).to(cuda), torch.from_numpy(conf).to(cuda), len(boxes)).cpu().numpy()
    for b in range(len(boxes)):
        sel = want[:, 0] == b + 1
        assert st[b, 0] == sel.sum() and st[b, 1] == (lab[sel] == 2).sum() and st[b, 2] == (conf[sel, 1] >= 1e-5).sum()
    new = np.full(len(boxes) + 1, -1, dtype=np.int32)
    new[1::2] = 2
    new[2::4] = 1
    out = refine.relabel_instances(ids, 0, torch.from_numpy(new).to(cuda), torch.from_numpy(lab).to(cuda).clone()).cpu().numpy()
    exp = lab.copy()
    hit = (want[:, 0] > 0) & (new[want[:, 0]] >= 0)
    exp[hit] = new[want[hit, 0]]
    assert np.array_equal(out, exp)
    empty = refine.point_instance_ids(torch.from_numpy(pts).to(cuda), torch.zeros((0, 8), device=cuda))
    assert int(empty.abs().sum()) == 0


@pytest.mark.gpu
def test_instance_refiner_matches_reference_script_golden(cuda):
    from insmos_b200 import refine
    poses, frames = _gold()
    r = refine.InstanceRefiner()
    for f, fr in enumerate(frames):
        got = r.step(f, torch.from_numpy(fr["scan"]).to(cuda), fr["boxes"], fr["labels"],
                     torch.from_numpy(_to12(fr["mos"])).to(cuda), torch.from_numpy(fr["conf"]).to(cuda), poses)
        got = _back(got.cpu().numpy())
        assert np.array_equal(got, fr["out"]), "frame %d: %d labels differ from the reference script" % (f, int((got != fr["out"]).sum()))
