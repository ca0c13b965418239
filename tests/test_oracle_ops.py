"""Self-validation of the oracle's ME / spconv restatement (nothing upstream pins it: SURVEY.md F3/F4).

Independent checks: dense torch convolutions on small grids with occupancy masks, brute-force
python loops, and algebraic properties (adjointness of transposed / inverse convolution,
first-occurrence numbering, map symmetry).
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import me, sp


def _rand_coords3(rng, n, size, ncol=4):
    c = rng.integers(0, size, (n, 3))
    c = np.unique(c, axis=0)
    rng.shuffle(c)
    b = np.zeros((len(c), 1), dtype=np.int64)
    cols = [b, c]
    if ncol == 5:
        cols.append(rng.integers(-2, 1, (len(c), 1)))
    return np.concatenate(cols, axis=1).astype(np.int32)


def test_unique_first_is_first_occurrence_order():
    rng = np.random.default_rng(0)
    c = rng.integers(-5, 5, (500, 5)).astype(np.int32)
    c[:, 0] = 0
    u, inv = me.unique_first(c)
    seen, order = {}, []
    for row in map(tuple, c):
        if row not in seen:
            seen[row] = len(order)
            order.append(row)
    assert [tuple(r) for r in u] == order
    assert [seen[tuple(r)] for r in c] == inv.tolist()


def test_quantize_uses_fp32_true_division_known_answers():
    # t/dt for t in {-0.9..0}: exactly {-9..0}; only the last is == 0  
    t = np.array([round((i - 9) * 0.1, 3) for i in range(10)], dtype=np.float32)
    pts = np.stack([np.zeros(10), np.zeros(10), np.zeros(10), t], axis=1).astype(np.float32)
    c, cur = me.quantize_points(pts, [0.1, 0.1, 0.1, 0.1])
    assert c[:, 4].tolist() == list(range(-9, 1))
    assert cur.tolist() == [False] * 9 + [True]
    # division differs from multiplication by the reciprocal on some inputs
    x = np.random.default_rng(1).uniform(-80, 80, 200000).astype(np.float32)
    d = np.floor(torch.div(torch.from_numpy(x), torch.tensor(0.1)).numpy())
    m = np.floor(x * np.float32(10.0))
    assert (d != m).sum() > 0


def test_me_conv3_matches_dense_conv3d():
    rng = np.random.default_rng(2)
    S, Cin, Cout = 9, 3, 5
    c = _rand_coords3(rng, 200, S)                       # (b,x,y,z)
    feats = torch.from_numpy(rng.normal(size=(len(c), Cin)).astype(np.float32))
    W = torch.from_numpy(rng.normal(size=(27, Cin, Cout)).astype(np.float32))
    maps = me.kernel_map(c, c, [3, 3, 3], [1, 1, 1])
    out = me.conv(feats, W, maps, len(c))
    dense = torch.zeros(1, Cin, S, S, S)                 # [z,y,x]
    dense[0, :, c[:, 3], c[:, 2], c[:, 1]] = feats.t()
    # offset index: x fastest -> k = kx + 3*ky + 9*kz
    wd = W.reshape(3, 3, 3, Cin, Cout).permute(4, 3, 0, 1, 2).contiguous()    # [Cout,Cin,kz,ky,kx]
    ref = F.conv3d(dense, wd, padding=1)[0, :, c[:, 3], c[:, 2], c[:, 1]].t()
    assert torch.allclose(out, ref, atol=1e-4)


def test_me_strided_conv_and_transpose():
    rng = np.random.default_rng(3)
    S, Cin, Cout = 8, 4, 6
    c = _rand_coords3(rng, 150, S)
    feats = torch.from_numpy(rng.normal(size=(len(c), Cin)).astype(np.float32))
    W = torch.from_numpy(rng.normal(size=(8, Cin, Cout)).astype(np.float32))
    oc, parent = me.stride_coords(c, [2, 2, 2])
    assert np.all(oc[:, 1:] % 2 == 0)
    maps = me.kernel_map(c, oc, [2, 2, 2], [1, 1, 1])
    # every fine voxel has exactly one parent and the map's out row is that parent
    tri = me.maps_to_triples(maps, len(c), len(oc))
    assert len(tri) == len(c)
    assert np.array_equal(tri[np.argsort(tri[:, 1]), 2], parent)
    out = me.conv(feats, W, maps, len(oc))
    dense = torch.zeros(1, Cin, S, S, S)
    dense[0, :, c[:, 3], c[:, 2], c[:, 1]] = feats.t()
    wd = W.reshape(2, 2, 2, Cin, Cout).permute(4, 3, 0, 1, 2).contiguous()
    ref = F.conv3d(dense, wd, stride=2)[0, :, oc[:, 3] // 2, oc[:, 2] // 2, oc[:, 1] // 2].t()
    assert torch.allclose(out, ref, atol=1e-4)
    # transposed conv with the swapped map is the adjoint: <conv(x), y> == <x, convT(y)>
    y = torch.from_numpy(rng.normal(size=(len(oc), Cout)).astype(np.float32))
    xt = me.conv(y, W.transpose(1, 2).contiguous(), me.transpose_map(maps), len(c))
    assert abs(float((out * y).sum() - (feats * xt).sum())) < 1e-2


def test_me_kernel_map_4d_brute_force_and_symmetry():
    rng = np.random.default_rng(4)
    c = _rand_coords3(rng, 120, 6, ncol=5)
    ks = [3, 3, 3, 3]
    maps = me.kernel_map(c, c, ks, [1, 1, 1, 1])
    lut = {tuple(r): i for i, r in enumerate(c)}
    offs = me.kernel_offsets(ks, [1, 1, 1, 1])
    brute = []
    for k, off in enumerate(offs):
        for o, r in enumerate(c):
            q = (r[0], r[1] + off[0], r[2] + off[1], r[3] + off[2], r[4] + off[3])
            if q in lut:
                brute.append((k, lut[q], o))
    tri = me.maps_to_triples(maps, len(c), len(c))
    assert sorted(brute) == [tuple(t) for t in tri]
    # offset index: x fastest, t slowest; k=0 is (-1,-1,-1,-1), k=1 is (0,-1,-1,-1), centre is k=40
    assert offs[0].tolist() == [-1, -1, -1, -1] and offs[1].tolist() == [0, -1, -1, -1] and offs[40].tolist() == [0, 0, 0, 0]
    # symmetry: (k, i, o) <-> (K-1-k, o, i)
    s = {tuple(t) for t in tri}
    assert all((80 - k, o, i) in s for (k, i, o) in s)
    # even kernels are anchored at 0 (not centred)
    assert me.kernel_offsets([2, 2, 2, 1], [4, 4, 4, 1])[7].tolist() == [4, 4, 4, 0]


def test_point_to_voxel_matches_python_loop():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-3, 3, (4000, 7)).astype(np.float32)
    rng_ = [-2.0, -2.0, -1.0, 2.0, 2.0, 1.0]
    vs = [0.5, 0.5, 0.5]
    for max_vox, max_pts in ((10000, 5), (37, 3)):
        vox, coords, num, ids = sp.point_to_voxel(pts, vs, rng_, max_pts, max_vox)
        table, order, members, ref_ids = {}, [], [], []
        for i, p in enumerate(pts):
            c = np.floor((p[:3] - np.float32(rng_[:3])) / np.float32(vs)).astype(np.int64)
            if np.any(c < 0) or np.any(c >= np.array([8, 8, 4])):
                ref_ids.append(-1)
                continue
            key = (c[2], c[1], c[0])
            if key not in table:
                if len(order) >= max_vox:
                    ref_ids.append(-1)
                    continue
                table[key] = len(order)
                order.append(key)
                members.append([])
            v = table[key]
            ref_ids.append(v)
            if len(members[v]) < max_pts:
                members[v].append(i)
        assert ids.tolist() == ref_ids
        assert [tuple(r) for r in coords] == order
        assert num.tolist() == [len(m) for m in members]
        for v, m in enumerate(members):
            assert torch.equal(vox[v, :len(m)], torch.from_numpy(pts[m]))
            assert float(vox[v, len(m):].abs().sum()) == 0.0
        mean = sp.mean_vfe(vox, num)
        assert torch.allclose(mean[0], torch.from_numpy(pts[members[0]]).sum(0) / len(members[0]), atol=1e-6)


def _spconv_dense_weight(W):                               # [Cout,kz,ky,kx,Cin] -> conv3d [Cout,Cin,kz,ky,kx]
    return W.permute(0, 4, 1, 2, 3).contiguous()


def test_spconv_subm_strided_inverse_match_dense():
    rng = np.random.default_rng(6)
    shape, Cin, Cout = [7, 10, 12], 3, 4
    zyx = np.stack([rng.integers(0, s, 150) for s in shape], axis=1)
    zyx = np.unique(zyx, axis=0)
    rng.shuffle(zyx)
    ind = np.concatenate([np.zeros((len(zyx), 1), dtype=np.int64), zyx], axis=1).astype(np.int32)
    feats = torch.from_numpy(rng.normal(size=(len(ind), Cin)).astype(np.float32))
    dense = sp.dense(feats, ind, shape)                                     # [1,Cin,Z,Y,X]
    # SubM: out set == in set, centred kernel
    W = torch.from_numpy(rng.normal(size=(Cout, 3, 3, 3, Cin)).astype(np.float32))
    out = sp.conv(feats, W, sp.subm_maps(ind, [3, 3, 3]), len(ind))
    ref = F.conv3d(dense, _spconv_dense_weight(W), padding=1)[0, :, ind[:, 1], ind[:, 2], ind[:, 3]].t()
    assert torch.allclose(out, ref, atol=1e-4)
    # strided SparseConv3d(k=3,s=2,p=1) and the (3,1,1)/(2,1,1)/p0 conv_out
    for ks, st, pd in (([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 1, 1], [2, 1, 1], [0, 0, 0])):
        W = torch.from_numpy(rng.normal(size=(Cout, *ks, Cin)).astype(np.float32))
        oind, maps, oshape = sp.sparse_conv_indices(ind, shape, ks, st, pd)
        out = sp.conv(feats, W, maps, len(oind))
        refd = F.conv3d(dense, _spconv_dense_weight(W), stride=st, padding=pd)
        assert list(refd.shape[2:]) == oshape
        assert torch.allclose(out, refd[0, :, oind[:, 1], oind[:, 2], oind[:, 3]].t(), atol=1e-4)
        # the output set covers every position a dense conv of the occupancy can reach
        occ = F.conv3d((dense.abs().sum(1, keepdim=True) > 0).float(), torch.ones(1, 1, *ks), stride=st, padding=pd)
        assert int((occ > 0).sum()) == len(oind)
        # creation order: scanning inputs ascending, offsets ascending
        seen, order = set(), []
        offs = sp._offsets3(ks)
        for r in ind:
            for o in offs:
                num = r[1:4] + np.asarray(pd) - o
                if np.any(num % np.asarray(st)):
                    continue
                q = num // np.asarray(st)
                if np.any(q < 0) or np.any(q >= np.asarray(oshape)):
                    continue
                t = (int(q[0]), int(q[1]), int(q[2]))
                if t not in seen:
                    seen.add(t)
                    order.append(t)
        assert [tuple(r[1:]) for r in oind.tolist()] == order
        # inverse conv (pairs swapped, same k) is the adjoint of the forward conv with transposed weights
        y = torch.from_numpy(rng.normal(size=(len(oind), Cout)).astype(np.float32))
        Winv = W.permute(4, 1, 2, 3, 0).contiguous()                         # [Cin,kz,ky,kx,Cout] : Cout -> Cin
        xt = sp.conv(y, Winv, me.transpose_map(maps), len(ind))
        assert abs(float((out * y).sum() - (feats * xt).sum())) < 1e-2
    # geometry chain of the model
    s = [41, 1000, 1200]
    chain = [s]
    for _ in range(3):
        s = sp.conv_out_shape(s, [3, 3, 3], [2, 2, 2], [1, 1, 1])
        chain.append(s)
    chain.append(sp.conv_out_shape(s, [3, 1, 1], [2, 1, 1], [0, 0, 0]))
    assert chain == [[41, 1000, 1200], [21, 500, 600], [11, 250, 300], [6, 125, 150], [2, 125, 150]]


def _same_maps(a, b):
    assert len(a) == len(b)
    for (i1, o1), (i2, o2) in zip(a, b):
        assert i1.dtype == i2.dtype == np.int64
        assert np.array_equal(i1, i2) and np.array_equal(o1, o2)


def test_native_neighbor_table_equals_the_numpy_statement():
    """large kernel maps are looked up in oracle/native (hash map + OpenMP over the offsets); the numpy statement (stable sort +
    binary search) is its checker: same pairs in the same order, incl. duplicated input rows (first row wins), coordinates at
    the edge of the packable range, strided maps, even kernels, empty sets, and the brute-force loop on a small 4D set."""
    rng = np.random.default_rng(11)
    for trial in range(4):
        n = 2500
        c = np.concatenate([np.zeros((n, 1), int), rng.integers(-6, 7, (n, 3)), rng.integers(-3, 1, (n, 1))], 1)
        if trial % 2:
            c[:50, 1] = 32767; c[50:100, 2] = -32767; c[100:120, 4] = 127; c[120:140, 4] = -127
        if trial == 2:
            c = np.concatenate([c, c[:400]], 0)
        out = c[rng.permutation(len(c))[:1500]]
        for ks, st in (([3, 3, 3, 3], [1, 1, 1, 1]), ([2, 2, 2, 1], [1, 1, 1, 1]), ([5, 5, 5, 1], [1, 1, 1, 1]), ([3, 3, 3, 3], [2, 2, 2, 1])):
            _same_maps(me.kernel_map(c, out, ks, st, native=False), me.kernel_map(c, out, ks, st, native=True))
    e = np.zeros((0, 5), int)
    for a, b in ((e, e), (c, e), (e, c)):
        _same_maps(me.kernel_map(a, b, [3, 3, 3, 3], [1, 1, 1, 1], native=False), me.kernel_map(a, b, [3, 3, 3, 3], [1, 1, 1, 1], native=True))
    ind = _rand_coords3(rng, 4000, 20).astype(np.int64)
    _same_maps(sp.subm_maps(ind, (3, 3, 3), native=False), sp.subm_maps(ind, (3, 3, 3), native=True))
    ind[:10, 3] = 32767
    _same_maps(sp.subm_maps(ind, (3, 3, 3), native=False), sp.subm_maps(ind, (3, 3, 3), native=True))
    # brute force, through the native path
    c = _rand_coords3(rng, 120, 6, ncol=5)
    maps = me.kernel_map(c, c, [3, 3, 3, 3], [1, 1, 1, 1], native=True)
    lut = {tuple(r): i for i, r in enumerate(c)}
    brute = sorted((k, lut[q], o) for k, off in enumerate(me.kernel_offsets([3, 3, 3, 3], [1, 1, 1, 1])) for o, r in enumerate(c)
                   for q in [(r[0], r[1] + off[0], r[2] + off[1], r[3] + off[2], r[4] + off[3])] if q in lut)
    assert brute == [tuple(t) for t in me.maps_to_triples(maps, len(c), len(c))]
    # the automatic switch takes the native path only for large maps
    assert 120 * 81 < me.NATIVE_MIN_PROBES


def test_c4_size_sets_and_maps_equal_the_stored_digests():
    """BASELINE config 4 scale (3 M points, voxel 0.05 m: 1.4 M voxels, maps of 1.4-25.7 M pairs): the oracle with native kernel-map
    lookups reproduces tests/golden/maps_c4.json, which was written through the numpy statement (make_golden_c2.py --c4-only).
    These are the digests the CUDA rule books are compared with on the GPU (tests/test_gpu_c2_golden.py::test_c4_size_maps_bit_exact)."""
    import hashlib
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import golden_util
    from insmos_b200 import synth
    want = json.load(open(os.path.join(golden_util.GOLDEN_DIR, "maps_c4.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()          # noqa: E731
    pts = synth.make_sequence(**want["synth"])
    assert len(pts) == want["n_points"]
    coords, _ = me.quantize_points(np.concatenate([pts[:, 0:3], pts[:, 4:5]], axis=1), [want["voxel"]] * 3 + [0.1])
    uniq, inv = me.unique_first(coords)
    assert {"n": int(len(uniq)), "sha": sha(uniq.astype(np.int32))} == want["sets"]["ts1"]
    assert sha(inv.astype(np.int32)) == want["inverse_sha"]
    c2, _ = me.stride_coords(uniq, [2, 2, 2, 1])
    assert {"n": int(len(c2)), "sha": sha(c2.astype(np.int32))} == want["sets"]["ts2"]
    for name, (ic, oc, ks, st) in {"ts1_5x5x5x1": (uniq, uniq, [5, 5, 5, 1], [1, 1, 1, 1]),
                                   "ts1_3x3x3x3": (uniq, uniq, [3, 3, 3, 3], [1, 1, 1, 1]),
                                   "ts1_to_ts2_2x2x2x1": (uniq, c2, [2, 2, 2, 1], [1, 1, 1, 1]),
                                   "ts2_3x3x3x3": (c2, c2, [3, 3, 3, 3], [2, 2, 2, 1])}.items():
        maps = me.kernel_map(ic, oc, ks, st, native=True)
        ks_ = np.concatenate([np.full(len(i), k, dtype=np.int64) for k, (i, o) in enumerate(maps)])
        s, x = golden_util.triple_digest(ks_, np.concatenate([i for i, o in maps]), np.concatenate([o for i, o in maps]))
        assert {"pairs": int(len(ks_)), "sum": s, "xor": x, "n_in": int(len(ic)), "n_out": int(len(oc))} == want["maps"][name], name
