"""oracle/graph.py (the travelling restatement of the model graph) vs the golden fixtures produced by the
reference's OWN models/*.py (tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util  # noqa: E402
from oracle import graph  # noqa: E402


@pytest.mark.parametrize("name", ["small_nodet", "small"])
def test_graph_matches_reference_code(name):
    meta, shapes, sd, pts, gold = golden_util.load(name)
    r = graph.forward(sd, pts)
    assert np.array_equal(r["voxel_coords"], gold["voxel_coords"])
    assert np.array_equal(r["pc_voxel_id"], gold["pc_voxel_id"])
    assert torch.allclose(r["current_point"], torch.from_numpy(gold["current_point"]), atol=1e-5, rtol=1e-5)
    assert r["n_cand"] == int(gold["n_cand"])
    assert torch.allclose(r["pred_boxes"], torch.from_numpy(gold["pred_boxes"]), atol=1e-5)
    assert torch.equal(r["pred_labels"], torch.from_numpy(gold["pred_labels"]))
    assert torch.allclose(r["logits"], torch.from_numpy(gold["logits"]), atol=1e-4, rtol=1e-5), \
        (r["logits"] - torch.from_numpy(gold["logits"])).abs().max()


def test_state_dict_contract_matches_reference():
    """the product mirror must expose exactly the reference's state_dict keys and shapes."""
    meta, shapes, _, _, _ = golden_util.load("small_nodet")
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    mine = {k: tuple(v.shape) for k, v in InsMOSNet(default_config()).state_dict().items()}
    assert set(mine) == set(shapes), (sorted(set(mine) ^ set(shapes))[:10])
    assert all(mine[k] == shapes[k] for k in shapes)
    assert meta["config"] == default_config()["MODEL"]["POST_PROCESSING"]


def test_graph_matches_reference_code_at_c2_size(monkeypatch):
    """the benchmark's own sample (BASELINE config 2: 10 x 120 000 points, cloud seed 0, the weights bench.py uses) through
    oracle/graph.py -- with the kernel-map lookups in oracle/native -- against the golden written by the reference's own model
    code (tests/golden/make_golden_c2.py): every coordinate set and voxel id bit for bit (sha256), every one of the 16 kernel maps
    by pair count + order-independent digest, the same 500 boxes, logits to 1e-5."""
    import hashlib
    from oracle import me, sp
    meta, shapes, sd, pts, gold = golden_util.load("c2")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()          # noqa: E731
    sets, maps = {}, {}

    def digest(name, m, n_in, n_out):
        ks = np.concatenate([np.full(len(i), k, dtype=np.int64) for k, (i, o) in enumerate(m)])
        s, x = golden_util.triple_digest(ks, np.concatenate([i for i, o in m]), np.concatenate([o for i, o in m]))
        maps[name] = {"pairs": int(len(ks)), "sum": s, "xor": x, "n_in": int(n_in), "n_out": int(n_out), "K": len(m)}

    o_km, o_sc, o_subm, o_sci = me.kernel_map, me.stride_coords, sp.subm_maps, sp.sparse_conv_indices

    def kernel_map(in_coords, out_coords, ksize, in_stride):
        m = o_km(in_coords, out_coords, ksize, in_stride)
        sets.setdefault("me_ts%d" % in_stride[0], {"n": int(len(in_coords)), "sha": sha(np.asarray(in_coords, dtype=np.int32))})
        digest("me_ts%d_to_n%d_k%s" % (in_stride[0], len(out_coords), "x".join(str(k) for k in ksize)), m, len(in_coords), len(out_coords))
        return m

    def stride_coords(coords, new_stride):
        u, inv = o_sc(coords, new_stride)
        sets["me_ts%d" % new_stride[0]] = {"n": int(len(u)), "sha": sha(np.asarray(u, dtype=np.int32))}
        return u, inv

    def subm_maps(indices, ksize):
        m = o_subm(indices, ksize)
        sets["sp_n%d" % len(indices)] = {"n": int(len(indices)), "sha": sha(np.asarray(indices, dtype=np.int32))}
        digest("sp_subm_n%d_k%s" % (len(indices), "x".join(str(k) for k in ksize)), m, len(indices), len(indices))
        return m

    def sparse_conv_indices(indices, in_shape, ksize, stride, pad):
        oind, m, oshape = o_sci(indices, in_shape, ksize, stride, pad)
        sets["sp_n%d" % len(oind)] = {"n": int(len(oind)), "sha": sha(np.asarray(oind, dtype=np.int32))}
        digest("sp_conv_n%d_k%s_s%s" % (len(indices), "x".join(str(k) for k in ksize), "x".join(str(k) for k in stride)), m, len(indices), len(oind))
        return oind, m, oshape

    monkeypatch.setattr(me, "kernel_map", kernel_map)
    monkeypatch.setattr(me, "stride_coords", stride_coords)
    monkeypatch.setattr(sp, "subm_maps", subm_maps)
    monkeypatch.setattr(sp, "sparse_conv_indices", sparse_conv_indices)
    r = graph.forward(sd, pts)
    assert sets == meta["sets"]
    assert set(maps) == set(meta["maps"]) and len(maps) == 16
    for name, want in meta["maps"].items():
        assert maps[name] == want, name
    assert sha(np.asarray(r["pc_voxel_id"], dtype=np.int64)) == meta["pc_voxel_id_sha"]
    assert {"n": int(len(r["voxel_coords"])), "sha": sha(np.asarray(r["voxel_coords"], dtype=np.int32))} == meta["voxel_coords"]
    assert r["n_cand"] == meta["n_cand"]
    assert torch.equal(r["pred_labels"], torch.from_numpy(gold["pred_labels"]))
    assert torch.allclose(r["pred_boxes"], torch.from_numpy(gold["pred_boxes"]), atol=1e-5)
    assert float((r["logits"] - torch.from_numpy(gold["logits"])).abs().max()) <= 1e-5
