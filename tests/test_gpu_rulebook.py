"""CUDA rule books vs the oracle's kernel maps -- bit exact after canonical sort (SURVEY 8d parity gate)."""
import numpy as np
import pytest
import torch

from insmos_b200 import ops, synth
from oracle import me, sp

pytestmark = pytest.mark.gpu


def _levels(cuda, seed=5, n_scans=4, n_elev=32, n_azim=500):
    pts = synth.make_sequence(seed=seed, n_scans=n_scans, n_elev=n_elev, n_azim=n_azim)
    cs, _, _ = ops.voxelize4d(torch.from_numpy(pts).to(cuda), [0.1, 0.1, 0.1, 0.1])
    return cs, cs.coords.cpu().numpy()


def _check(rb, maps, n_in, n_out):
    got = rb.to_coo().numpy()
    exp = me.maps_to_triples(maps, n_in, n_out)
    assert rb.num_pairs == len(exp), "pair count %d vs oracle %d" % (rb.num_pairs, len(exp))
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("ksize", [[3, 3, 3, 3], [5, 5, 5, 1]])
@pytest.mark.parametrize("TM", [None, 16, 128])
def test_me_cube_maps(cuda, ksize, TM):
    cs, c = _levels(cuda)
    rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ksize, [1, 1, 1, 1]), TM=TM)
    _check(rb, me.kernel_map(c, c, ksize, [1, 1, 1, 1]), len(c), len(c))


@pytest.mark.parametrize("ksize", [[3, 3, 3, 3], [5, 5, 5, 1], [3, 3, 3, 1]])
def test_me_cube_maps_xblock_table_identical(cuda, ksize, monkeypatch):
    """x-block probing must reproduce the voxel-table rule book bit for bit, at every level (coordinates = multiples of ts)."""
    monkeypatch.setattr(ops, "XBLOCK_MIN_KX", 3)          # also exercise the 3-wide runs (off by default: no speed-up)
    cs, c = _levels(cuda)
    g, ts = cs, 1
    for _ in range(3):
        for TM in (None, 16):
            spec = ops.spec_me_cube(ksize, [ts, ts, ts, 1])
            a = ops.build_rulebook(g, g, spec, TM=TM)
            b = ops.build_rulebook(g, g, spec, TM=TM, xstep=ts)
            assert a.num_pairs == b.num_pairs
            assert torch.equal(a.seg, b.seg)
            tot = int(a.seg.view(torch.int16).to(torch.int32).view(-1, a.K + 1)[:, -1].sum())   # occupied prefixes only
            assert np.array_equal(a.to_coo().numpy(), b.to_coo().numpy())
        if ts == 1:
            _check(b, me.kernel_map(c, c, ksize, [1, 1, 1, 1]), len(c), len(c))
        g, _ = ops.unique_coords(g.coords, q=[2 * ts, 2 * ts, 2 * ts, 1])
        ts *= 2


def test_me_strided_and_transposed_maps_all_levels(cuda):
    cs, c = _levels(cuda)
    fine_g, fine_o, ts = cs, c, 1
    for _ in range(3):
        coarse_o, _ = me.stride_coords(fine_o, [2 * ts, 2 * ts, 2 * ts, 1])
        coarse_g, parent = ops.unique_coords(fine_g.coords, q=[2 * ts, 2 * ts, 2 * ts, 1])
        maps = me.kernel_map(fine_o, coarse_o, [2, 2, 2, 1], [ts, ts, ts, 1])
        rb = ops.build_rulebook(coarse_g, fine_g, ops.spec_me_cube([2, 2, 2, 1], [ts, ts, ts, 1]))
        _check(rb, maps, len(fine_o), len(coarse_o))
        rbt = ops.build_rulebook(fine_g, coarse_g, ops.spec_me_up([2, 2, 2, 1], [2, 2, 2, 1], [ts, ts, ts, 1]))
        _check(rbt, me.transpose_map(maps), len(coarse_o), len(fine_o))
        # probe-free build from the parent map: identical rule book, bit for bit, at every tile size
        for TM in (None, 16, 64):
            a = ops.build_rulebook(fine_g, coarse_g, ops.spec_me_up([2, 2, 2, 1], [2, 2, 2, 1], [ts, ts, ts, 1]), TM=TM)
            b = ops.build_rulebook(fine_g, coarse_g, ops.spec_me_up([2, 2, 2, 1], [2, 2, 2, 1], [ts, ts, ts, 1]), TM=TM, parent=parent)
            assert a.num_pairs == b.num_pairs == len(fine_o)
            assert torch.equal(a.seg, b.seg)
            assert np.array_equal(a.to_coo().numpy(), b.to_coo().numpy())
        # 3^4 map at the coarse level (offsets scaled by the tensor stride, time stride stays 1)
        ts2 = 2 * ts
        rb3 = ops.build_rulebook(coarse_g, coarse_g, ops.spec_me_cube([3, 3, 3, 3], [ts2, ts2, ts2, 1]))
        _check(rb3, me.kernel_map(coarse_o, coarse_o, [3, 3, 3, 3], [ts2, ts2, ts2, 1]), len(coarse_o), len(coarse_o))
        fine_g, fine_o, ts = coarse_g, coarse_o, ts2


def test_spconv_maps(cuda):
    pts = synth.make_sequence(seed=6, n_scans=1, n_elev=64, n_azim=1000)
    _, coords, _, _ = sp.point_to_voxel(pts[:, :4], [0.1] * 3, [-60, -50, -3, 60, 50, 1], 5, 100000)
    ind = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], axis=1)
    gin, _ = ops.unique_coords(torch.from_numpy(ind).to(cuda))
    rb = ops.build_rulebook(gin, gin, ops.spec_sp_subm([3, 3, 3]))
    _check(rb, sp.subm_maps(ind, [3, 3, 3]), len(ind), len(ind))
    shape = [41, 1000, 1200]
    cur_g, cur_o = gin, ind
    for ks, st, pd in (([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 1, 1], [2, 1, 1], [0, 0, 0])):
        oind, maps, oshape = sp.sparse_conv_indices(cur_o, shape, ks, st, pd)
        og = ops.spconv_out_coords(cur_g, ks, st, pd, oshape)
        rb = ops.build_rulebook(og, cur_g, ops.spec_sp_conv(ks, st, pd))
        _check(rb, maps, len(cur_o), len(oind))
        rbi = ops.build_rulebook(cur_g, og, ops.spec_sp_inverse(ks, st, pd))
        _check(rbi, me.transpose_map(maps), len(oind), len(cur_o))
        cur_g, cur_o, shape = og, oind, oshape


def _same_book(a, b):
    assert a.num_pairs == b.num_pairs
    assert torch.equal(a.seg, b.seg), "bucket offsets differ"
    assert np.array_equal(a.to_coo().numpy(), b.to_coo().numpy())


@pytest.mark.parametrize("ksize", [[3, 3, 3, 3], [5, 5, 5, 1], [3, 3, 3, 1], [2, 2, 2, 1]])
def test_leafgrid_maps_identical_at_every_level(cuda, ksize, monkeypatch):
    """the leaf-grid build (4x4x4x1 leaves, <= 24 leaf probes per row, indexed loads per offset) reproduces the voxel-table
    rule book bit for bit -- same seg, same entries -- at every tensor stride, and matches the oracle."""
    monkeypatch.setattr(ops, "LEAFGRID_MIN_ROWS", 0)
    monkeypatch.setattr(ops, "LEAFGRID_MIN_K", 1)
    cs, c = _levels(cuda)
    g, ts = cs, 1
    for lvl in range(3):
        step = (ts, ts, ts, 1)
        for TM in (None, 16):
            spec = ops.spec_me_cube(ksize, [ts, ts, ts, 1])
            assert ops.leafgrid_eligible(spec, step)
            a = ops.build_rulebook(g, g, spec, TM=TM)                      # voxel table
            b = ops.build_rulebook(g, g, spec, TM=TM, step=step)           # leaf grid
            assert "_leafgrid" in g.__dict__
            _same_book(a, b)
        if ts == 1:
            _check(b, me.kernel_map(c, c, ksize, [1, 1, 1, 1]), len(c), len(c))
        coarse, _ = ops.unique_coords(g.coords, q=[2 * ts, 2 * ts, 2 * ts, 1])
        # strided map: out = coarse level, in = this level (children of a coarse cell sit in one leaf)
        sd = ops.spec_me_cube([2, 2, 2, 1], [ts, ts, ts, 1])
        _same_book(ops.build_rulebook(coarse, g, sd), ops.build_rulebook(coarse, g, sd, step=step))
        g, ts = coarse, 2 * ts


def test_leafgrid_spconv_maps_negative_and_edge_coordinates(cuda, monkeypatch):
    monkeypatch.setattr(ops, "LEAFGRID_MIN_ROWS", 0)
    monkeypatch.setattr(ops, "LEAFGRID_MIN_K", 1)
    pts = synth.make_sequence(seed=6, n_scans=1, n_elev=64, n_azim=1000)
    _, coords, _, _ = sp.point_to_voxel(pts[:, :4], [0.1] * 3, [-60, -50, -3, 60, 50, 1], 5, 100000)
    ind = np.concatenate([np.zeros((len(coords), 1), np.int32), coords], axis=1)
    gin, _ = ops.unique_coords(torch.from_numpy(ind).to(cuda))
    a = ops.build_rulebook(gin, gin, ops.spec_sp_subm([3, 3, 3]))
    b = ops.build_rulebook(gin, gin, ops.spec_sp_subm([3, 3, 3]), step=(1, 1, 1))
    _same_book(a, b)
    _check(b, sp.subm_maps(ind, [3, 3, 3]), len(ind), len(ind))
    oind, maps, oshape = sp.sparse_conv_indices(ind, [41, 1000, 1200], [3, 3, 3], [2, 2, 2], [1, 1, 1])
    og = ops.spconv_out_coords(gin, [3, 3, 3], [2, 2, 2], [1, 1, 1], oshape)
    rb = ops.build_rulebook(og, gin, ops.spec_sp_conv([3, 3, 3], [2, 2, 2], [1, 1, 1]), step=(1, 1, 1))
    _check(rb, maps, len(ind), len(oind))
    assert not ops.leafgrid_eligible(ops.spec_sp_inverse([3, 3, 3], [2, 2, 2], [1, 1, 1]), (1, 1, 1))
    # negative coordinates, leaf boundaries, and rows at the edge of the packable range (checked path inside the kernel)
    g = torch.Generator().manual_seed(1)
    xyz = torch.randint(-9, 9, (3000, 3), generator=g)
    t = torch.randint(-3, 1, (3000, 1), generator=g)
    far = torch.tensor([[32760, 0, 0, 0], [32761, 0, 0, 0], [-32760, 5, 5, -1], [-32759, 5, 5, -1], [0, 32765, -32765, 0]])
    c4 = torch.cat([torch.cat([xyz, t], 1), far], 0)
    c5 = torch.cat([torch.zeros((len(c4), 1), dtype=torch.long), c4], 1).to(torch.int32)
    cs, _ = ops.unique_coords(c5.to(cuda))
    cn = cs.coords.cpu().numpy()
    for ksize in ([3, 3, 3, 3], [5, 5, 5, 1]):
        spec = ops.spec_me_cube(ksize, [1, 1, 1, 1])
        p = ops.build_rulebook(cs, cs, spec)
        q = ops.build_rulebook(cs, cs, spec, step=(1, 1, 1, 1))
        _same_book(p, q)
        _check(q, me.kernel_map(cn, cn, ksize, [1, 1, 1, 1]), len(cn), len(cn))


def test_leafgrid_overflow_falls_back_to_the_voxel_table(cuda, monkeypatch):
    """isolated voxels: every voxel is its own leaf, more leaves than the bounded probing tolerates -> the grid raises its
    overflow word and the builder probes the voxel table instead (decided on the device); result unchanged."""
    monkeypatch.setattr(ops, "LEAFGRID_MIN_ROWS", 0)
    g = torch.Generator().manual_seed(2)
    n = 6000
    xyz = torch.randint(-1000, 1000, (n, 3), generator=g) * 8            # 8 apart: one voxel per 4x4x4 leaf
    c5 = torch.cat([torch.zeros((n, 1), dtype=torch.long), xyz, torch.zeros((n, 1), dtype=torch.long)], 1).to(torch.int32)
    cs, _ = ops.unique_coords(c5.to(cuda))
    spec = ops.spec_me_cube([3, 3, 3, 3], [1, 1, 1, 1])
    a = ops.build_rulebook(cs, cs, spec)
    b = ops.build_rulebook(cs, cs, spec, step=(1, 1, 1, 1))
    _same_book(a, b)
    assert a.num_pairs == cs.n                                            # only the centre offset pairs anything
