"""oracle/train.py (CPU restatement of one training step) pinned against tests/golden/train_small.npz, which
was written by the REFERENCE's own model code in train mode (tests/golden/make_golden_train.py): losses, targets,
logits, and the gradient of every one of the 201 parameter tensors (L2 norm + 4 seeded projections, full tensors <= 4096)."""
import json
import os
import sys
import zlib

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from insmos_b200 import synth, synth_weights  # noqa: E402
from oracle import train as otrain  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_small.npz")
LOSS_RTOL = 1e-5


def projection(key, shape, j):
    r = np.random.default_rng(zlib.crc32(("%s#%d" % (key, j)).encode()))
    return (r.integers(0, 2, size=int(np.prod(shape))).astype(np.float32) * 2 - 1).reshape(shape)


def load_train_golden():
    g = np.load(GOLDEN, allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    pts, labels, boxes = synth.make_sequence(return_labels=True, **meta["synth"])
    assert np.array_equal(boxes, g["out:gt_boxes"]) and np.array_equal(labels, g["out:labels"]), "synthetic generator drifted"
    sd = synth_weights.fill_state_dict({k: tuple(v) for k, v in meta["shapes"].items()})
    sd["model.unet.center_head.conv_cls.bias"] = torch.full((3,), float(meta["cls_bias"]))
    return g, meta, sd, pts, labels, boxes


GOLDEN_F64 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_small_f64.npz")


def grad_summary(grads, n_proj=4):
    """{key: tensor} -> {key: (norm, projections, full tensor or None)} in float64"""
    out = {}
    for k, gr in grads.items():
        gr = gr.detach().double().cpu().numpy()
        out[k] = (float(np.sqrt((gr ** 2).sum())), np.asarray([(gr * projection(k, gr.shape, j)).sum() for j in range(n_proj)]),
                  gr if gr.size <= 4096 else None)
    return out


def npz_summary(g):
    return {k[6:]: (float(g[k]), np.asarray(g["gproj:" + k[6:]]), g["gfull:" + k[6:]].astype(np.float64) if "gfull:" + k[6:] in g.files else None)
            for k in g.files if k.startswith("gnorm:")}


def grad_errors(got, truth):
    """per parameter: max over (norm, 4 projections, all elements of small tensors) of |got - truth| / ||truth||"""
    assert set(got) >= set(truth) and len(truth) == 201
    err = {}
    for k, (n, proj, full) in truth.items():
        gn, gproj, gfull = got[k]
        scale = max(n, 1e-20)
        e = max(abs(gn - n), float(np.abs(gproj - proj).max())) / scale
        if full is not None and gfull is not None:
            e = max(e, float(np.sqrt(((gfull - full) ** 2).sum())) / scale)
        err[k] = e
    return err


def check_grads_vs_f64(grads, factor=3.0, floor=2e-3):
    """the implementation must be as close to the fp64 gradients as the REFERENCE's own fp32 run is: per parameter
    err <= max(factor * err_reference, floor).  Returns (worst key, its error, the reference's error there)."""
    truth = npz_summary(np.load(GOLDEN_F64, allow_pickle=False))
    ref_err = grad_errors(npz_summary(np.load(GOLDEN, allow_pickle=False)), truth)
    err = grad_errors(grad_summary(grads), truth)
    worst = max(err, key=lambda k: err[k] / max(factor * ref_err[k], floor))
    for k, e in err.items():
        assert e <= max(factor * ref_err[k], floor), \
            "gradient of %s: relative error %.3e vs fp64, the reference's fp32 run has %.3e there" % (k, e, ref_err[k])
    return worst, err[worst], ref_err[worst]


@pytest.fixture(scope="module")
def step():
    g, meta, sd, pts, labels, boxes = load_train_golden()
    # the instance-fusion stage is fed the reference's own detections: the greedy NMS over ~75 k near-tied candidates is
    # sensitive to the last bit of the scores (MKL thread count changes the summation order), see tests/test_gpu_model.py
    override = {"pred_boxes": torch.from_numpy(g["out:pred_boxes"]), "pred_labels": torch.from_numpy(g["out:pred_labels"])}
    return g, otrain.train_step(sd, pts, labels, boxes, pred_override=override)


def test_losses_match_reference(step):
    g, r = step
    for k in ("loss", "loss_mos", "loss_motion_encoder", "rpn_loss_cls", "rpn_loss_loc"):
        ref = float(g["out:" + k])
        assert abs(r[k] - ref) <= LOSS_RTOL * abs(ref), (k, r[k], ref)


def test_targets_match_reference(step):
    g, r = step
    heat, anno, inds, masks = r["targets"]
    assert list(heat.shape) == list(g["out:heatmap_shape"][1:])
    idx = np.flatnonzero(heat)
    assert np.array_equal(idx, g["out:heatmap_nonzero_index"])
    assert np.array_equal(heat.reshape(-1)[idx], g["out:heatmap_nonzero_value"])          # bit-exact Gaussian peaks
    assert np.array_equal(inds, g["out:inds"][0]) and np.array_equal(masks, g["out:masks"][0])
    assert np.abs(anno - g["out:anno_boxes"][0]).max() <= 1e-6


def test_forward_outputs_and_detections_match_reference(step):
    g, r = step
    assert np.abs(r["logits"].numpy()[:, 1:] - g["out:point_seg_feature"][:, 1:]).max() < 1e-4
    assert np.abs(r["motion"].numpy()[:, 1:] - g["out:current_motion_feature"][:, 1:]).max() < 1e-4
    # free-running detections of the oracle: same count, and (up to rank swaps of near-tied candidates) the same boxes
    assert r["pred"]["pred_boxes"].shape == g["out:pred_boxes"].shape
    same = np.abs(r["pred"]["pred_boxes"].numpy() - g["out:pred_boxes"]).max(axis=1) < 1e-3
    assert same.mean() > 0.9, same.mean()


def test_reference_gradients_agree_with_the_fp64_oracle():
    """pins the GRAPH of oracle/train.py: the reference's fp32 gradients (golden) against the oracle's fp64 gradients.  A wiring
    or formula difference shows as an O(1) error; fp32 rounding shows as <= 2 % in the deepest layers, ~1e-5 at the median."""
    truth = npz_summary(np.load(GOLDEN_F64, allow_pickle=False))
    err = grad_errors(npz_summary(np.load(GOLDEN, allow_pickle=False)), truth)
    assert max(err.values()) < 5e-2, max(err, key=err.get)
    assert float(np.median(list(err.values()))) < 1e-4
    g64 = np.load(GOLDEN_F64, allow_pickle=False)
    g32 = np.load(GOLDEN, allow_pickle=False)
    for k in ("loss", "loss_mos", "loss_motion_encoder", "rpn_loss_cls", "rpn_loss_loc"):
        assert abs(float(g32["out:" + k]) - float(g64["out:" + k])) <= 1e-5 * abs(float(g64["out:" + k])), k


def test_gradients_match_reference(step):
    """the fp32 oracle is as close to the fp64 gradients as the reference's own fp32 run"""
    g, r = step
    worst = check_grads_vs_f64(r["grads"])
    print("worst: %s err %.3e (reference fp32: %.3e)" % worst)


def test_center_targets_edge_cases():
    """boxes outside the map, degenerate sizes and class 0 are skipped; a box at the border is clipped (center_head.py:204-227,380-392)"""
    boxes = np.array([[-59.9, -49.9, 0, 4, 2, 1.5, 0.3, 1],       # corner: window clipped on two sides
                      [100.0, 0, 0, 4, 2, 1.5, 0, 2],             # outside the range
                      [0, 0, 0, 0.0, 2, 1.5, 0, 1],               # zero width
                      [5, 5, 0, 4, 2, 1.5, 0, 0],                 # class 0 -> cls_id -1
                      [5.03, -7.21, -1, 0.8, 0.6, 1.7, -2.0, 3]], dtype=np.float32)
    heat, anno, inds, masks = otrain.center_targets(boxes)
    assert masks[:5].tolist() == [1, 0, 0, 0, 1]
    assert heat[0, 0, 0] == 1.0 and heat[2].max() == 1.0 and heat[1].max() == 0.0
    assert inds[0] == 0 and inds[4] == int((-7.21 + 50) / 0.4) * 300 + int((5.03 + 60) / 0.4)
    assert np.allclose(anno[4, 3:6], np.log([0.8, 0.6, 1.7]), atol=1e-6) and anno[1].sum() == 0
