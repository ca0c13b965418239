"""Size-independent properties at the BASELINE.json sizes (C2: 10 x 120 k points, voxel 0.1; C4-like: 10 x 300 k points,
voxel 0.05), where the CPU oracle would take minutes to hours: adjointness of the sparse convolution over its own rule
book (<conv(x;W), y> = <x, conv(y; W'), W'[k] = W[K-1-k]^T, true iff the submanifold map is symmetric and every pair is
applied exactly once), linearity, partition property of the strided maps, idempotence of the voxel set, run-to-run
determinism of the whole forward."""
import numpy as np
import pytest
import torch

from insmos_b200 import ops, synth

pytestmark = pytest.mark.gpu


def _voxels(cuda, n_elev, n_azim, voxel, seed=3):
    pts = synth.make_sequence(seed=seed, n_scans=10, n_elev=n_elev, n_azim=n_azim)
    cs, inverse, cur = ops.voxelize4d(torch.from_numpy(pts).to(cuda), [voxel, voxel, voxel, 0.1])
    return pts, cs, inverse, cur


def _adjoint_gap(cuda, rb, K, Cin, Cout, n, algo, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((n, Cin), generator=g).to(cuda)
    y = torch.randn((n, Cout), generator=g).to(cuda)
    W = (torch.randn((K, Cin, Cout), generator=g) / np.sqrt(Cin * 8.0)).to(cuda)
    Wt = W.flip(0).transpose(1, 2).contiguous()
    lhs = (ops.sparse_conv(x, W, rb, algo=algo).double() * y.double()).sum().item()
    rhs = (x.double() * ops.sparse_conv(y, Wt, rb, algo=algo if Cin % 4 == 0 else 3).double()).sum().item()   # (FFMA path needs Cout % 4 == 0)
    scale = (ops.sparse_conv(x.abs(), W.abs(), rb, algo=algo).double() * y.abs().double()).sum().item()
    return abs(lhs - rhs) / scale


@pytest.mark.parametrize("ksize,Cin,Cout,algo", [([3, 3, 3, 3], 16, 8, 2), ([5, 5, 5, 1], 8, 8, 2), ([3, 3, 3, 1], 32, 32, 4),
                                                 ([3, 3, 3, 1], 32, 32, 2), ([3, 3, 3, 3], 1, 8, 1)])
def test_c2_conv_is_adjoint_over_its_rulebook(cuda, ksize, Cin, Cout, algo):
    pts, cs, _, _ = _voxels(cuda, 64, 1875, 0.1)
    assert pts.shape[0] > 1_100_000 and cs.n > 300_000                     # BASELINE config 2 size
    K = int(np.prod(ksize))
    rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ksize, [1, 1, 1, 1]), xstep=1)
    assert rb.num_pairs >= cs.n                                             # the centre offset pairs every row with itself
    # 3xTF32 / fp32 accumulate: the two sides differ by rounding only
    assert _adjoint_gap(cuda, rb, K, Cin, Cout, cs.n, algo, seed=K + Cin) < 2e-6


def test_c2_strided_maps_partition_and_voxel_set_is_idempotent(cuda):
    _, cs, inverse, cur = _voxels(cuda, 64, 1875, 0.1)
    again, inv2 = ops.unique_coords(cs.coords)
    assert again.n == cs.n and torch.equal(again.coords, cs.coords)
    assert torch.equal(inv2, torch.arange(cs.n, dtype=torch.int32, device=cuda))
    assert int(inverse.max()) == cs.n - 1 and int(inverse.min()) == 0
    fine, ts = cs, 1
    for _ in range(3):
        coarse, parent = ops.unique_coords(fine.coords, q=[2 * ts, 2 * ts, 2 * ts, 1])
        up = ops.build_rulebook(fine, coarse, ops.spec_me_up([2, 2, 2, 1], [2, 2, 2, 1], [ts, ts, ts, 1]), parent=parent)
        down = ops.build_rulebook(coarse, fine, ops.spec_me_cube([2, 2, 2, 1], [ts, ts, ts, 1]))
        assert up.num_pairs == fine.n == down.num_pairs                    # every fine voxel has exactly one parent
        # transposed conv of ones with all-ones 1x1 weights = 1 per fine row; strided conv of ones = children count, sums to n_fine
        ones_c = torch.ones((coarse.n, 1), device=cuda)
        w = torch.ones((8, 1, 4), device=cuda)
        assert torch.equal(ops.sparse_conv(ones_c, w, up, algo=1), torch.ones((fine.n, 4), device=cuda))
        kids = ops.sparse_conv(torch.ones((fine.n, 1), device=cuda), w, down, algo=1)
        assert float(kids[:, 0].sum()) == float(fine.n) and float(kids.max()) <= 8.0
        fine, ts = coarse, 2 * ts


def test_c4_dense_scene_rulebook_and_conv_sweep(cuda):
    """BASELINE config 4 shape: 300 k points per scan (160 x 1875 rays), N = 10, voxel 0.05 m: map build + gather/scatter only."""
    pts, cs, _, _ = _voxels(cuda, 160, 1875, 0.05, seed=4)
    assert pts.shape[0] > 2_800_000 and cs.n > 1_000_000
    for ksize, Cin, Cout in (([5, 5, 5, 1], 1, 8), ([3, 3, 3, 3], 16, 8)):
        K = int(np.prod(ksize))
        rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ksize, [1, 1, 1, 1]), xstep=1)
        rb_plain = ops.build_rulebook(cs, cs, ops.spec_me_cube(ksize, [1, 1, 1, 1]))
        assert rb.num_pairs == rb_plain.num_pairs >= cs.n and torch.equal(rb.seg, rb_plain.seg)
        assert _adjoint_gap(cuda, rb, K, Cin, Cout, cs.n, 0, seed=K) < 2e-6
        # linearity
        g = torch.Generator().manual_seed(K)
        x1, x2 = torch.randn((cs.n, Cin), generator=g).to(cuda), torch.randn((cs.n, Cin), generator=g).to(cuda)
        W = (torch.randn((K, Cin, Cout), generator=g) / 4.0).to(cuda)
        a = ops.sparse_conv(2.0 * x1 - 0.5 * x2, W, rb)
        b = 2.0 * ops.sparse_conv(x1, W, rb) - 0.5 * ops.sparse_conv(x2, W, rb)
        assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item())


def test_c2_forward_is_deterministic_and_finite(cuda):
    import bench
    clouds = [torch.from_numpy(c).to(cuda) for c in bench.make_clouds(0, 1)]
    net = bench.build_model(cuda, clouds[0])
    with torch.no_grad():
        l1, b1 = bench.step(net, clouds[0])
        l2, b2 = bench.step(net, clouds[0])
    assert l1.shape == (120_000, 3) and torch.isfinite(l1).all()
    assert torch.equal(l1, l2), "two forwards over the same scan must agree bit for bit (no atomics on the value path)"
    assert torch.equal(b1["pred_boxes"], b2["pred_boxes"]) and b1["pred_boxes"].shape[0] > 0
    labels, conf = ops.mos_labels(l1, 1, None)
    assert torch.equal(labels.long(), 1 + l1[:, 1:].argmax(1))              # class 0 ignored: argmax over the live classes
