"""Training step on the GPU (BASELINE config 5) vs the oracle and vs the golden written by the reference's own
model code in train mode: the gradient kernels one by one (wgrad, dgrad through the transposed maps, train-mode BatchNorm,
scatter-add, target assignment, Adam), then the whole step -- losses, targets, logits and the gradient of every parameter."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from insmos_b200 import autograd as ag  # noqa: E402
from insmos_b200 import ops, synth  # noqa: E402
from oracle import me, sp  # noqa: E402
from oracle import train as otrain  # noqa: E402
from test_oracle_train import check_grads_vs_f64, load_train_golden  # noqa: E402

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / max(float(b.norm()), 1e-20))


def _me_setup(cuda, seed=7):
    pts = synth.make_sequence(seed=seed, n_scans=3, n_elev=32, n_azim=400)
    cs, _, _ = ops.voxelize4d(torch.from_numpy(pts).to(cuda), [0.1, 0.1, 0.1, 0.1])
    return cs, cs.coords.cpu().numpy()


@pytest.mark.parametrize("Cin,Cout", [(1, 8), (8, 8), (16, 8), (24, 16), (48, 32), (7, 3), (19, 16), (136, 128), (64, 64), (32, 96), (256, 128)])
def test_wgrad_matches_autograd_of_oracle(cuda, Cin, Cout):
    cs, c = _me_setup(cuda)
    ks = [3, 3, 3, 1] if Cin * Cout > 8192 else [3, 3, 3, 3]
    K = int(np.prod(ks))
    maps = me.kernel_map(c, c, ks, [1, 1, 1, 1])
    rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ks, [1, 1, 1, 1]))
    g = torch.Generator().manual_seed(Cin * 1000 + Cout)
    x = torch.randn((len(c), Cin), generator=g)
    dy = torch.randn((len(c), Cout), generator=g)
    W = torch.zeros((K, Cin, Cout), requires_grad=True)
    (me.conv(x, W, maps, len(c)) * dy).sum().backward()
    dw = ops.sparse_conv_wgrad(x.to(cuda), dy.to(cuda), rb, K, Cin, Cout)
    assert _rel(dw, W.grad) < 1e-5
    dw2 = ops.sparse_conv_wgrad(x.to(cuda), dy.to(cuda), rb, K, Cin, Cout)
    assert torch.equal(dw, dw2), "wgrad must be deterministic"


def _check_conv_grads(cuda, f, W, maps, n_out, rb, rb_t, flip, tol=2e-5):
    x_ref = f.clone().requires_grad_(True)
    W_ref = W.clone().requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    out_ref = me.conv(x_ref, W_ref, maps, n_out)
    dy = torch.randn(out_ref.shape, generator=g)
    (out_ref * dy).sum().backward()
    x = f.to(cuda).requires_grad_(True)
    Wd = W.to(cuda).requires_grad_(True)
    out = ag.sparse_conv(x, Wd, rb, rb_t, flip)
    assert torch.allclose(out.detach().cpu(), out_ref.detach(), rtol=2e-5, atol=2e-5)
    (out * dy.to(cuda)).sum().backward()
    assert _rel(x.grad, x_ref.grad) < tol, "dgrad"
    assert _rel(Wd.grad, W_ref.grad) < tol, "wgrad"


@pytest.mark.parametrize("ks,Cin,Cout", [([3, 3, 3, 3], 8, 16), ([5, 5, 5, 1], 8, 8), ([3, 3, 3, 3], 48, 32)])
def test_stride1_conv_backward_reuses_the_forward_rulebook_flipped(cuda, ks, Cin, Cout):
    cs, c = _me_setup(cuda)
    maps = me.kernel_map(c, c, ks, [1, 1, 1, 1])
    rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ks, [1, 1, 1, 1]))
    g = torch.Generator().manual_seed(1)
    K = int(np.prod(ks))
    _check_conv_grads(cuda, torch.randn((len(c), Cin), generator=g), torch.randn((K, Cin, Cout), generator=g) / np.sqrt(Cin * 20.0),
                      maps, len(c), rb, rb, True)


def test_me_layers_backward_strided_and_transposed(cuda):
    """MinkowskiConvolution (2,2,2,1)/2 and MinkowskiConvolutionTranspose through the compat modules vs autograd over oracle/me.py"""
    import insmos_b200
    insmos_b200.install()
    import MinkowskiEngine as ME
    cs, c = _me_setup(cuda)
    mgr = ME.CoordinateManager(4)
    mgr.sets[(1, 1, 1, 1)] = cs
    g = torch.Generator().manual_seed(5)
    f = torch.randn((len(c), 8), generator=g)
    down = ME.MinkowskiConvolution(8, 16, kernel_size=[2, 2, 2, 1], stride=[2, 2, 2, 1], dimension=4).to(cuda)
    up = ME.MinkowskiConvolutionTranspose(16, 8, kernel_size=[2, 2, 2, 1], stride=[2, 2, 2, 1], dimension=4).to(cuda)
    x = ME.SparseTensor(f.to(cuda).requires_grad_(True), coordinate_manager=mgr, coordinate_map_key=(1, 1, 1, 1))
    y = down(x)
    z = up(y)
    dz = torch.randn(z.F.shape, generator=g)
    (z.F * dz.to(cuda)).sum().backward()
    # oracle
    c2, _ = me.stride_coords(c, [2, 2, 2, 1])
    assert np.array_equal(c2, mgr.sets[(2, 2, 2, 1)].coords.cpu().numpy())
    maps = me.kernel_map(c, c2, [2, 2, 2, 1], [1, 1, 1, 1])
    fr = f.clone().requires_grad_(True)
    Wd = down.kernel.detach().cpu().clone().requires_grad_(True)
    Wu = up.kernel.detach().cpu().clone().requires_grad_(True)
    yr = me.conv(fr, Wd, maps, len(c2))
    zr = me.conv(yr, Wu, me.transpose_map(maps), len(c))
    (zr * dz).sum().backward()
    assert torch.allclose(z.F.detach().cpu(), zr.detach(), rtol=2e-5, atol=2e-5)
    assert _rel(x.F.grad, fr.grad) < 2e-5 and _rel(down.kernel.grad, Wd.grad) < 2e-5 and _rel(up.kernel.grad, Wu.grad) < 2e-5


def test_spconv_layers_backward(cuda):
    """SubMConv3d -> SparseConv3d (stride 2) -> SparseInverseConv3d through the compat modules vs autograd over oracle/sp.py"""
    import insmos_b200
    insmos_b200.install()
    import spconv.pytorch as spconv
    rng = np.random.default_rng(0)
    ind = np.unique(np.concatenate([np.zeros((4000, 1), np.int64), rng.integers(0, [9, 60, 60], (4000, 3))], 1), axis=0).astype(np.int32)
    rng.shuffle(ind)
    shape = [9, 60, 60]
    g = torch.Generator().manual_seed(9)
    f = torch.randn((len(ind), 16), generator=g)
    subm = spconv.SubMConv3d(16, 16, 3, padding=1, bias=False, indice_key="s").to(cuda)
    down = spconv.SparseConv3d(16, 32, 3, stride=2, padding=1, bias=False, indice_key="d").to(cuda)
    inv = spconv.SparseInverseConv3d(32, 8, 3, indice_key="d", bias=False).to(cuda)
    for m in (subm, down, inv):
        m.train()
    x = spconv.SparseConvTensor(f.to(cuda).requires_grad_(True), torch.from_numpy(ind).to(cuda), shape, 1)
    z = inv(down(subm(x)))
    dz = torch.randn(z.features.shape, generator=g)
    (z.features * dz.to(cuda)).sum().backward()
    fr = f.clone().requires_grad_(True)
    Ws, Wd, Wi = (m.weight.detach().cpu().clone().requires_grad_(True) for m in (subm, down, inv))
    m1 = sp.subm_maps(ind, (3, 3, 3))
    a = sp.conv(fr, Ws, m1, len(ind))
    oind, m2, oshape = sp.sparse_conv_indices(ind, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    b = sp.conv(a, Wd, m2, len(oind))
    zr = sp.conv(b, Wi, me.transpose_map(m2), len(ind))
    (zr * dz).sum().backward()
    assert torch.allclose(z.features.detach().cpu(), zr.detach(), rtol=1e-4, atol=1e-4)
    assert _rel(x.features.grad, fr.grad) < 1e-4
    for m, w in ((subm, Ws), (down, Wd), (inv, Wi)):
        assert _rel(m.weight.grad, w.grad) < 1e-4


@pytest.mark.parametrize("n,C,relu", [(50000, 8, True), (3001, 16, False), (7, 128, True), (20000, 256, True), (1000, 3, False)])
def test_batch_norm_train_matches_torch(cuda, n, C, relu):
    g = torch.Generator().manual_seed(n + C)
    x = (torch.randn((n, C), generator=g) * 3 + 1.5)
    dy = torch.randn((n, C), generator=g)
    bn_ref = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01)
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.normal_()
    import copy
    bn = copy.deepcopy(bn_ref).to(cuda)
    xr = x.clone().requires_grad_(True)
    yr = bn_ref(xr)
    yr = torch.relu(yr) if relu else yr
    (yr * dy).sum().backward()
    xd = x.to(cuda).requires_grad_(True)
    y = ag.batch_norm_train(bn, xd, relu=relu)
    (y * dy.to(cuda)).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-5, atol=1e-5)
    assert _rel(xd.grad, xr.grad) < 1e-4
    assert _rel(bn.weight.grad, bn_ref.weight.grad) < 1e-4 and _rel(bn.bias.grad, bn_ref.bias.grad) < 1e-4
    assert torch.allclose(bn.running_mean.cpu(), bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn.running_var.cpu(), bn_ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1


def test_gather_rows_backward_is_scatter_add(cuda):
    g = torch.Generator().manual_seed(2)
    src = torch.randn((500, 5), generator=g)
    idx = torch.randint(-1, 500, (4000,), generator=g, dtype=torch.int32)
    dy = torch.randn((4000, 5), generator=g)
    s = src.to(cuda).requires_grad_(True)
    out = ag.gather_rows(s, idx.to(cuda))
    (out * dy.to(cuda)).sum().backward()
    ref = torch.zeros_like(src).index_add_(0, idx[idx >= 0].long(), dy[idx >= 0])
    assert torch.allclose(s.grad.cpu(), ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(out.detach().cpu()[idx >= 0], src[idx[idx >= 0].long()]) and float(out.detach().cpu()[idx < 0].abs().sum()) == 0


def test_center_targets_match_oracle(cuda):
    rng = np.random.default_rng(4)
    n = 60
    boxes = np.stack([rng.uniform(-70, 70, n), rng.uniform(-60, 60, n), rng.uniform(-2, 0, n), rng.uniform(0.3, 12, n),
                      rng.uniform(0.3, 5, n), rng.uniform(0.5, 3, n), rng.uniform(-3.2, 3.2, n), rng.integers(0, 4, n)], 1).astype(np.float32)
    boxes[3, 3] = 0.0
    boxes[5, 0:2] = [-59.95, 49.9]
    heat_o, anno_o, inds_o, masks_o = otrain.center_targets(boxes)
    heat, anno, inds, masks = ops.center_targets(torch.from_numpy(boxes).to(cuda), 100, 250, 300, 3, -60, -50, 0.1, 0.1, 4, 0.1, 2)
    assert np.array_equal(masks.cpu().numpy(), masks_o) and np.array_equal(inds.cpu().numpy(), inds_o)
    h = heat.cpu().numpy()
    assert np.array_equal(h != 0, heat_o != 0)
    assert np.abs(h - heat_o).max() <= 6e-8                      # fp64 exp on the device vs numpy, rounded to fp32
    assert np.abs(anno.cpu().numpy() - anno_o).max() <= 1e-6


def test_adam_step_matches_torch(cuda):
    g = torch.Generator().manual_seed(8)
    p0 = torch.randn(10001, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4)
    p = p0.to(cuda)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(10001, generator=g)
        ref.grad = gr.clone()
        opt.step()
        ops.adam_step(p, (2 * gr).to(cuda), m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step, grad_scale=0.5)
        assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-6, atol=1e-7), step


# ---- the whole step ---------------------------------------------------------------------------------------------
def _train_net(cuda, sd):
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    net = InsMOSNet(default_config())
    net.load_state_dict(sd, strict=True)
    return net.to(cuda).train()


def _batch(cuda, pts, labels, boxes, override=None):
    d = {"meta": None, "past_point_clouds": torch.from_numpy(pts).to(cuda), "past_labels": [torch.from_numpy(labels.astype(np.float32)).to(cuda)],
         "gt_boxes": torch.from_numpy(boxes).unsqueeze(0).to(cuda), "batch_size_npast": 0}
    if override is not None:
        d["instance_boxes_override"] = {"pred_boxes": torch.as_tensor(override["pred_boxes"]).float().to(cuda),
                                        "pred_labels": torch.as_tensor(override["pred_labels"]).long().to(cuda)}
    return [d]


@pytest.fixture(scope="module")
def golden_step(cuda):
    g, meta, sd, pts, labels, boxes = load_train_golden()
    net = _train_net(cuda, sd)
    batch = _batch(cuda, pts, labels, boxes, override={"pred_boxes": g["out:pred_boxes"], "pred_labels": g["out:pred_labels"]})
    loss, dicts, gts, preds = net.forward(batch, "train")
    loss.backward()
    torch.cuda.synchronize()
    return g, net, batch, loss, dicts, preds


def test_train_step_losses_match_reference_golden(golden_step):
    g, net, batch, loss, dicts, preds = golden_step
    got = {"loss": float(loss), "loss_mos": float(dicts[0]["loss_mos"]), "loss_motion_encoder": float(dicts[0]["loss_motion_encoder"]),
           "rpn_loss_cls": float(dicts[0]["rpn_loss_cls"]), "rpn_loss_loc": float(dicts[0]["rpn_loss_loc"])}
    for k, v in got.items():
        ref = float(g["out:" + k])
        assert abs(v - ref) <= 1e-4 * abs(ref), (k, v, ref)
    assert np.abs(preds[0].detach().cpu().numpy()[:, 1:] - g["out:point_seg_feature"][:, 1:]).max() < 1e-3
    assert np.abs(batch[0]["current_motion_feature"].detach().cpu().numpy()[:, 1:] - g["out:current_motion_feature"][:, 1:]).max() < 1e-3


def test_train_step_targets_match_reference_golden(golden_step):
    g, net, *_ = golden_step
    fr = net.model.unet.center_head.forward_ret_dict
    heat = fr["heatmaps"][0].cpu().numpy().reshape(-1)
    idx = np.flatnonzero(heat)
    assert np.array_equal(idx, g["out:heatmap_nonzero_index"])
    assert np.abs(heat[idx] - g["out:heatmap_nonzero_value"]).max() <= 6e-8
    assert np.array_equal(fr["inds"][0].cpu().numpy(), g["out:inds"]) and np.array_equal(fr["masks"][0].cpu().numpy(), g["out:masks"])
    assert np.abs(fr["anno_boxes"][0].cpu().numpy() - g["out:anno_boxes"]).max() <= 1e-6


def test_train_step_gradients_match_reference_golden(golden_step):
    g, net, *_ = golden_step
    # measured against the fp64 gradients (tests/golden/train_small_f64.npz): per parameter no further from them than
    # max(3x the reference's own fp32 deviation, 2e-3 of the gradient's norm)
    grads = {k: p.grad for k, p in net.named_parameters()}
    worst = check_grads_vs_f64(grads)
    print("worst: %s err %.3e (reference fp32: %.3e)" % worst)


def test_train_step_running_statistics_match_reference_golden(golden_step):
    g, net, *_ = golden_step
    sd = net.state_dict()
    for k in (k for k in g.files if k.startswith("bn_after:")):
        assert torch.allclose(sd[k[9:]].cpu(), torch.from_numpy(g[k]), rtol=1e-4, atol=1e-5), k


def test_free_running_train_forward_keeps_the_reference_boxes(cuda):
    g, meta, sd, pts, labels, boxes = load_train_golden()
    net = _train_net(cuda, sd)
    captured = {}
    import insmos_b200.net.unet3d as u
    orig = u.post_processing

    def cap(bd, cfg, nc):
        r = orig(bd, cfg, nc)
        captured["pred"] = r[0][0]
        return r
    u.post_processing = cap
    try:
        loss, dicts, _, _ = net.forward(_batch(cuda, pts, labels, boxes), "train")
    finally:
        u.post_processing = orig
    pb = captured["pred"]["pred_boxes"].cpu().numpy()
    assert pb.shape == g["out:pred_boxes"].shape
    same = np.abs(pb - g["out:pred_boxes"]).max(axis=1) < 1e-2
    assert same.mean() > 0.98, "free-running boxes agree on %.3f of the rows" % same.mean()
    assert abs(float(loss) - float(g["out:loss"])) <= 1e-3 * abs(float(g["out:loss"]))


def test_optimizer_step_updates_the_model_and_invalidates_weight_caches(cuda):
    from insmos_b200.train import TrainStep
    g, meta, sd, pts, labels, boxes = load_train_golden()
    net = _train_net(cuda, sd)
    ts = TrainStep(net, lr=1e-3, weight_decay=1e-4)
    before = ts.flat.data.clone()
    r1 = ts.step(_batch(cuda, pts, labels, boxes), want_confusion=True)
    assert abs(r1["loss"] - float(g["out:loss"])) <= 1e-3 * abs(float(g["out:loss"]))
    assert r1["confusion_matrix"].sum().item() == len(labels)
    # first Adam step moves every coordinate with a non-zero gradient by ~lr
    moved = (ts.flat.data - before).abs()
    assert float(moved.max()) <= 1.01e-3 + 1e-4 * float(before.abs().max()) and float(moved.mean()) > 1e-4
    # the parameters of the module ARE the flat buffer, and the second step sees the new weights (caches keyed on _version)
    k0 = net.model.motion_encoder.MinkUNet.conv0p1s1.kernel
    assert k0.data_ptr() >= ts.flat.data.data_ptr() and k0.grad.data_ptr() >= ts.flat.grad.data_ptr()
    r2 = ts.step(_batch(cuda, pts, labels, boxes))
    assert r2["loss"] != r1["loss"] and np.isfinite(r2["loss"])
    # reference: the same two steps with torch.optim.Adam over the same module gradients
    net2 = _train_net(cuda, sd)
    opt = torch.optim.Adam(net2.parameters(), lr=1e-3, weight_decay=1e-4)
    for _ in range(2):
        opt.zero_grad()
        l2, *_ = net2.forward(_batch(cuda, pts, labels, boxes), "train")
        l2.backward()
        opt.step()
    assert abs(float(l2) - r2["loss"]) <= 2e-3 * abs(r2["loss"]), (float(l2), r2["loss"])
    # Adam normalises every coordinate's step to ~lr, so a coordinate whose gradient is pure rounding noise may move the other
    # way (2 steps x lr = 2e-3 apart at most); all but a handful must agree closely
    p2 = torch.cat([p.detach().reshape(-1) for p in net2.parameters()])
    d = (p2 - ts.flat.data).abs()
    assert float(d.max()) <= 4.1e-3 and float((d > 1e-4).float().mean()) < 1e-2 and float(d.median()) < 1e-6, (float(d.max()), float((d > 1e-4).float().mean()), float(d.median()))
