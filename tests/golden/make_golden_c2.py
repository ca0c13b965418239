"""Generate tests/golden/insmos_c2.npz: the full BASELINE-config-2 sample (10 scans x 120 000 points, seed 0 -- the
first cloud bench.py times) run through the REFERENCE's own model code on CPU, plus tests/golden/maps_c4.npz (pair counts
and hashes of C4-size kernel maps from the oracle).

Run in the development container only (needs /root/reference):   python tests/golden/make_golden_c2.py [--c4-only|--c2-only]

Same machinery as make_golden.py (reference models/models.py::InsMOSNet over the oracle-backed ME/spconv shims, the
reference's own Array_Index.cpp compiled as is, C NMS pinned against the reference's iou3d code).  Because a fixture of
every intermediate would be hundreds of MB, the integer artefacts are stored as digests:
  * every coordinate set (row order matters): sha256 of the int32 rows + row count;
  * every kernel map: pair count + an order-independent 64-bit digest of its (k, in_row, out_row) triples
    (golden_util.triple_digest: sum and xor of a mixed 64-bit word per triple);
  * 3D voxel coords / pc_voxel_id: sha256;
and the floating-point outputs in full: MOS logits [120000,3], motion features [120000,3], the detections.
The BatchNorm statistics calibrated by the reference code on this input are stored, so that bench.py and the GPU tests
run the forward on exactly these weights.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets sys.path for oracle / reference / shims)

import golden_util  # noqa: E402
from oracle import me, sp  # noqa: E402
from insmos_b200 import synth, synth_weights  # noqa: E402

C2_SYNTH = dict(seed=0, n_scans=10, n_elev=64, n_azim=1875)
C4_SYNTH = dict(seed=4, n_scans=10, n_elev=160, n_azim=1875)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Recorder:
    """wraps the oracle's map / coordinate builders so that every map the reference graph requests is digested."""

    def __init__(self):
        self.maps, self.sets = {}, {}
        self._orig = {}

    def _digest(self, name, maps, n_in, n_out):
        ks = np.concatenate([np.full(len(i), k, dtype=np.int64) for k, (i, o) in enumerate(maps)])
        ins = np.concatenate([np.asarray(i, dtype=np.int64) for i, o in maps])
        outs = np.concatenate([np.asarray(o, dtype=np.int64) for i, o in maps])
        s, x = golden_util.triple_digest(ks, ins, outs)
        self.maps[name] = {"pairs": int(len(ks)), "sum": s, "xor": x, "n_in": int(n_in), "n_out": int(n_out), "K": len(maps)}

    def install(self):
        rec = self
        o_km, o_sc, o_subm, o_sci = me.kernel_map, me.stride_coords, sp.subm_maps, sp.sparse_conv_indices
        self._orig = dict(km=o_km, sc=o_sc, subm=o_subm, sci=o_sci)

        def kernel_map(in_coords, out_coords, ksize, in_stride):
            maps = o_km(in_coords, out_coords, ksize, in_stride)
            rec.sets.setdefault("me_ts%d" % in_stride[0], {"n": int(len(in_coords)), "sha": sha(np.asarray(in_coords, dtype=np.int32))})
            name = "me_ts%d_to_n%d_k%s" % (in_stride[0], len(out_coords), "x".join(str(k) for k in ksize))
            rec._digest(name, maps, len(in_coords), len(out_coords))
            return maps

        def stride_coords(coords, new_stride):
            u, inv = o_sc(coords, new_stride)
            rec.sets["me_ts%d" % new_stride[0]] = {"n": int(len(u)), "sha": sha(np.asarray(u, dtype=np.int32))}
            return u, inv

        def subm_maps(indices, ksize):
            maps = o_subm(indices, ksize)
            name = "sp_subm_n%d_k%s" % (len(indices), "x".join(str(k) for k in ksize))
            rec.sets["sp_n%d" % len(indices)] = {"n": int(len(indices)), "sha": sha(np.asarray(indices, dtype=np.int32))}
            rec._digest(name, maps, len(indices), len(indices))
            return maps

        def sparse_conv_indices(indices, in_shape, ksize, stride, pad):
            oind, maps, oshape = o_sci(indices, in_shape, ksize, stride, pad)
            name = "sp_conv_n%d_k%s_s%s" % (len(indices), "x".join(str(k) for k in ksize), "x".join(str(k) for k in stride))
            rec.sets["sp_n%d" % len(oind)] = {"n": int(len(oind)), "sha": sha(np.asarray(oind, dtype=np.int32))}
            rec._digest(name, maps, len(indices), len(oind))
            return oind, maps, oshape

        me.kernel_map, me.stride_coords, sp.subm_maps, sp.sparse_conv_indices = kernel_map, stride_coords, subm_maps, sparse_conv_indices

    def uninstall(self):
        me.kernel_map, me.stride_coords = self._orig["km"], self._orig["sc"]
        sp.subm_maps, sp.sparse_conv_indices = self._orig["subm"], self._orig["sci"]


def make_c2():
    net, cfg = mg.load_reference_model()
    pts, labels, _ = synth.make_sequence(return_labels=True, **C2_SYNTH)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    rec = Recorder()
    rec.install()
    try:
        shapes, sd, out = mg.run_case(net, pts, 0.0)
    finally:
        rec.uninstall()
    print("reference graph over the oracle shims: %.0f s (calibration pass + eval pass)" % (time.time() - t0))
    bn = {k: sd[k].numpy() for k in synth_weights.bn_stat_keys(shapes)}
    meta = {"synth": C2_SYNTH, "cls_bias": 0.0, "shapes": {k: list(v) for k, v in shapes.items()},
            "points_sha": int(np.abs(pts).sum() * 1000) % (1 << 31), "config": cfg["MODEL"]["POST_PROCESSING"],
            "maps": rec.maps, "sets": rec.sets,
            "voxel_coords": {"n": int(len(out["voxel_coords"])), "sha": sha(out["voxel_coords"].astype(np.int32))},
            "pc_voxel_id_sha": sha(out["pc_voxel_id"].astype(np.int64)),
            "n_dropped_points": int((out["pc_voxel_id"] < 0).sum()),
            "n_cand": int(out["n_cand"])}
    keep = {"logits": out["logits"].astype(np.float32), "motion": out["current_point"][:, 4:].astype(np.float32),
            "pred_boxes": out["pred_boxes"], "pred_scores": out["pred_scores"], "pred_labels": out["pred_labels"],
            "mos_labels": labels.astype(np.int8), "all_scores": out["all_scores"].astype(np.float32),
            "all_boxes": out["all_boxes"].astype(np.float32)}
    path = os.path.join(HERE, "insmos_c2.npz")
    np.savez_compressed(path, meta=json.dumps(meta), **{"bn:" + k: v for k, v in bn.items()},
                        **{"out:" + k: v for k, v in keep.items()})
    print("c2", {k: v.shape for k, v in keep.items()}, "boxes", len(out["pred_boxes"]), "cand", int(out["n_cand"]),
          "maps", {k: v["pairs"] for k, v in rec.maps.items()}, "%.1f KB" % (os.path.getsize(path) / 1024))


def make_c4():
    """C4-size maps (300 k points per scan x 10, voxel 0.05 m): counts + digests only."""
    pts = synth.make_sequence(**C4_SYNTH)
    xyzt = np.concatenate([pts[:, 0:3], pts[:, 4:5]], axis=1)
    coords, _ = me.quantize_points(xyzt, [0.05, 0.05, 0.05, 0.1])
    uniq, inv = me.unique_first(coords)
    res = {"synth": C4_SYNTH, "voxel": 0.05, "n_points": int(len(pts)), "sets": {}, "maps": {}}
    res["sets"]["ts1"] = {"n": int(len(uniq)), "sha": sha(uniq.astype(np.int32))}
    res["inverse_sha"] = sha(inv.astype(np.int32))
    c2, _ = me.stride_coords(uniq, [2, 2, 2, 1])
    res["sets"]["ts2"] = {"n": int(len(c2)), "sha": sha(c2.astype(np.int32))}
    for name, (ic, oc, ks, st) in {"ts1_5x5x5x1": (uniq, uniq, [5, 5, 5, 1], [1, 1, 1, 1]),
                                   "ts1_3x3x3x3": (uniq, uniq, [3, 3, 3, 3], [1, 1, 1, 1]),
                                   "ts1_to_ts2_2x2x2x1": (uniq, c2, [2, 2, 2, 1], [1, 1, 1, 1]),
                                   "ts2_3x3x3x3": (c2, c2, [3, 3, 3, 3], [2, 2, 2, 1])}.items():
        t0 = time.time()
        maps = me.kernel_map(ic, oc, ks, st)
        ks_ = np.concatenate([np.full(len(i), k, dtype=np.int64) for k, (i, o) in enumerate(maps)])
        ins = np.concatenate([np.asarray(i, dtype=np.int64) for i, o in maps])
        outs = np.concatenate([np.asarray(o, dtype=np.int64) for i, o in maps])
        s, x = golden_util.triple_digest(ks_, ins, outs)
        res["maps"][name] = {"pairs": int(len(ks_)), "sum": s, "xor": x, "n_in": int(len(ic)), "n_out": int(len(oc))}
        print("c4", name, res["maps"][name], "%.0f s" % (time.time() - t0))
    with open(os.path.join(HERE, "maps_c4.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    if "--c4-only" not in sys.argv:
        make_c2()
    if "--c2-only" not in sys.argv:
        make_c4()
