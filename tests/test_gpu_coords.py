"""CUDA coordinate ops vs the oracle: voxel indices, inverse maps, row order -- bit exact."""
import numpy as np
import pytest
import torch

from insmos_b200 import ops, synth
from oracle import me, sp

pytestmark = pytest.mark.gpu


def _seq(seed, n_scans, n_elev, n_azim):
    return synth.make_sequence(seed=seed, n_scans=n_scans, n_elev=n_elev, n_azim=n_azim)


@pytest.mark.parametrize("n_scans,n_elev,n_azim", [(3, 16, 200), (10, 32, 600), (1, 64, 1875)])
def test_voxelize4d_bit_exact(cuda, n_scans, n_elev, n_azim):
    pts = _seq(1, n_scans, n_elev, n_azim)
    quant = [0.1, 0.1, 0.1, 0.1]
    cs, inverse, cur = ops.voxelize4d(torch.from_numpy(pts).to(cuda), quant)
    oc, ocur = me.quantize_points(np.concatenate([pts[:, :3], pts[:, 4:5]], 1), quant)
    ou, oinv = me.unique_first(oc)
    assert cs.n == len(ou)
    assert np.array_equal(cs.coords.cpu().numpy(), ou)
    assert np.array_equal(inverse.cpu().numpy().astype(np.int64), oinv)
    assert np.array_equal(cur.cpu().numpy().astype(np.int64), np.nonzero(ocur)[0])


def test_voxelize4d_empty_and_duplicates(cuda):
    cs, inv, cur = ops.voxelize4d(torch.zeros((0, 5), device=cuda), [0.1, 0.1, 0.1, 0.1])
    assert cs.n == 0 and inv.numel() == 0 and cur.numel() == 0
    p = torch.tensor([[0.05, 0.05, 0.05, 1.0, 0.0]] * 1000 + [[-0.05, 0.05, 0.05, 1.0, -0.1]], device=cuda)
    cs, inv, cur = ops.voxelize4d(p, [0.1, 0.1, 0.1, 0.1])
    assert cs.n == 2 and cs.coords.cpu().tolist() == [[0, 0, 0, 0, 0], [0, -1, 0, 0, -1]]
    assert inv.cpu().tolist() == [0] * 1000 + [1] and cur.numel() == 1000


def test_voxelize4d_range_error_is_loud(cuda):
    p = torch.tensor([[1.0e6, 0, 0, 0, 0]], device=cuda)
    with pytest.raises(RuntimeError, match="packable range"):
        ops.voxelize4d(p, [0.1, 0.1, 0.1, 0.1])


def test_unique_and_stride_chain(cuda):
    pts = _seq(2, 4, 32, 500)
    oc, _ = me.quantize_points(np.concatenate([pts[:, :3], pts[:, 4:5]], 1), [0.1, 0.1, 0.1, 0.1])
    cs, inv = ops.unique_coords(torch.from_numpy(oc).to(cuda))
    ou, oinv = me.unique_first(oc)
    assert np.array_equal(cs.coords.cpu().numpy(), ou) and np.array_equal(inv.cpu().numpy(), oinv)
    cur_o, cur_g = ou, cs
    for ts in (2, 4, 8):
        nxt_o, par_o = me.stride_coords(cur_o, [ts, ts, ts, 1])
        nxt_g, par_g = ops.unique_coords(cur_g.coords, q=[ts, ts, ts, 1])
        assert np.array_equal(nxt_g.coords.cpu().numpy(), nxt_o), "stride %d" % ts
        assert np.array_equal(par_g.cpu().numpy(), par_o)
        cur_o, cur_g = nxt_o, nxt_g


@pytest.mark.parametrize("max_vox,max_pts", [(100000, 5), (500, 3)])
def test_voxelize3d_bit_exact(cuda, max_vox, max_pts):
    pts = _seq(3, 1, 64, 1875)
    rng = np.random.default_rng(0)
    p7 = np.concatenate([pts[:, :4], rng.normal(size=(len(pts), 3)).astype(np.float32)], axis=1)
    pc_range = [-60, -50, -3, 60, 50, 1]
    vs = [0.1, 0.1, 0.1]
    r = ops.voxelize3d(torch.from_numpy(p7).to(cuda), pc_range, vs, [1200, 1000, 40], max_vox, max_pts, want_voxels=True)
    vox, coords, num, ids = sp.point_to_voxel(p7, vs, pc_range, max_pts, max_vox)
    assert r["set"].n == len(coords)
    assert np.array_equal(r["set"].coords.cpu().numpy()[:, 1:], coords)
    assert np.all(r["set"].coords.cpu().numpy()[:, 0] == 0)
    assert np.array_equal(r["pc_voxel_id"].cpu().numpy().astype(np.int64), ids)
    assert np.array_equal(r["num_points"].cpu().numpy(), num)
    assert torch.equal(r["voxels"].cpu(), vox)
    mean = sp.mean_vfe(vox, num)
    assert torch.allclose(r["mean"].cpu(), mean, atol=1e-6, rtol=1e-6)
    assert (ids == -1).sum() > 0                                     # out-of-range points exist in the scene


def test_spconv_out_coords_creation_order(cuda):
    pts = _seq(4, 1, 64, 1000)
    _, coords, _, _ = sp.point_to_voxel(pts[:, :4], [0.1] * 3, [-60, -50, -3, 60, 50, 1], 5, 100000)
    ind = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], axis=1)
    gin, _ = ops.unique_coords(torch.from_numpy(ind).to(cuda))
    shape = [41, 1000, 1200]
    for ks, st, pd in (([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 1, 1], [2, 1, 1], [0, 0, 0])):
        oind, _, oshape = sp.sparse_conv_indices(ind, shape, ks, st, pd)
        g = ops.spconv_out_coords(gin, ks, st, pd, oshape)
        assert g.n == len(oind)
        assert np.array_equal(g.coords.cpu().numpy(), oind)
