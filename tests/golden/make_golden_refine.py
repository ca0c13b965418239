"""Generate tests/golden/refine_small.npz by running the REFERENCE's own scripts/refine.py main() on a synthetic sequence.

Run in the development container only (needs /root/reference):   python tests/golden/make_golden_refine.py

What runs: /root/reference/scripts/refine.py (unmodified) with
  * models.utils.Array_Index = the reference's own Array_Index.cpp compiled as is (oracle/_ref),
  * `open3d`, `spconv` import stubs (imported by the script, never used by main()),
  * a temporary working directory holding ./config/semantic-kitti-mos.yaml (copied at run time), the synthetic KITTI-layout
    sequence 08 (velodyne/*.bin, poses.txt, calib.txt) and ./preb_out/InsMOS/{bbox_preb,mos_preb,confidence}/... as
    predict_mos.py would have written them.
Scene: 8 cars fixed in the world or moving at 1 m/frame, a pedestrian and a cyclist box, 11 frames, ego motion with a slow yaw;
the per-point MOS predictions are drawn so that every branch of the frame logic fires (highly dynamic bottom-up relabelling in
the first 5 frames, window tracking afterwards, > 6 cars top-down reset, confidences from frame 9).
The fixture stores the inputs and the .label files the reference wrote; oracle/refine.py and insmos_b200/refine.py must
reproduce them exactly.
"""
import importlib.util
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), ROOT]

from oracle import build  # noqa: E402

F = 11
CAR = np.array([4.2, 1.8, 1.6], dtype=np.float32)


def velo_pose(i):
    yaw = 0.015 * i
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[0, 3] = 0.5 * i
    T[1, 3] = 0.02 * i
    return T


def make_sequence(rng):
    """-> per frame: scan [N,4] f32 (sensor frame), boxes [nb,7] f32, labels [nb] i64, mos uint32 [N] (9/251), conf [N,2] f32"""
    n_car = 8
    car_c = np.stack([rng.uniform(-25, 25, n_car), rng.uniform(-8, 8, n_car), np.full(n_car, -0.9)], 1)
    car_c[:, 0] = np.linspace(-28, 28, n_car) + rng.uniform(-1, 1, n_car)                # well separated along x
    car_v = np.zeros((n_car, 3)); car_v[:4, 0] = [1.0, -1.0, 0.8, 1.2]                    # the first four move
    car_yaw = rng.uniform(-0.3, 0.3, n_car)
    others = np.array([[5.0, 12.0, -0.9, 0.8, 0.8, 1.8, 0.1, 2], [-7.0, -13.0, -0.9, 1.8, 0.7, 1.7, -0.4, 3]])
    p_moving = np.array([0.95, 0.8, 0.65, 0.4, 0.0, 0.02, 0.0, 0.2])                     # P(point predicted moving) per car
    frames = []
    for f in range(F):
        Tinv = np.linalg.inv(velo_pose(f))
        pts, mos, conf = [], [], []
        boxes, labels = [], []
        for c in range(n_car):
            centre_w = car_c[c] + car_v[c] * f
            local = rng.uniform(-0.5, 0.5, (50, 3)) * CAR * 0.96
            cy, sy = np.cos(car_yaw[c]), np.sin(car_yaw[c])
            world = np.stack([local[:, 0] * cy - local[:, 1] * sy, local[:, 0] * sy + local[:, 1] * cy, local[:, 2]], 1) + centre_w
            pts.append((Tinv @ np.hstack([world, np.ones((50, 1))]).T).T[:, :3])
            pm = p_moving[c] if f not in (3, 7) else min(1.0, p_moving[c] + 0.3)         # two frames with more "moving" votes
            m = rng.uniform(0, 1, 50) < pm
            mos.append(np.where(m, 251, 9))
            conf.append(np.stack([np.where(m, 0.2, 0.9), np.where(m, 0.8, 0.0) * (rng.uniform(0, 1, 50) < 0.9)], 1))
            cs = (Tinv @ np.append(centre_w, 1.0))[:3] + rng.normal(0, 0.05, 3)
            yaw_s = car_yaw[c] - 0.015 * f
            boxes.append(np.concatenate([cs, CAR + rng.normal(0, 0.03, 3), [yaw_s]]))
            labels.append(1)
        for o in others:
            local = rng.uniform(-0.5, 0.5, (20, 3)) * o[3:6] * 0.9
            world = local + o[:3]
            pts.append((Tinv @ np.hstack([world, np.ones((20, 1))]).T).T[:, :3])
            mos.append(np.full(20, 251 if o[7] == 2 else 9))
            conf.append(np.tile([0.5, 0.5], (20, 1)))
            boxes.append(np.concatenate([(Tinv @ np.append(o[:3], 1.0))[:3], o[3:6], [o[6] - 0.015 * f]]))
            labels.append(int(o[7]))
        bg = np.stack([rng.uniform(-40, 40, 600), rng.uniform(-20, 20, 600), rng.uniform(-1.8, -1.6, 600)], 1)
        pts.append(bg); mos.append(np.full(600, 9)); conf.append(np.tile([1.0, 0.0], (600, 1)))
        P = np.concatenate(pts, 0)
        perm = rng.permutation(len(P))                                                   # first-hit pruning sees a shuffled order
        scan = np.concatenate([P, rng.uniform(0, 1, (len(P), 1))], 1).astype(np.float32)[perm]
        if f == 2:                                                                       # a frame where one car has no points at all
            boxes[5][:3] += [0.0, 30.0, 0.0]
        frames.append({"scan": scan, "boxes": np.asarray(boxes, dtype=np.float32), "labels": np.asarray(labels, dtype=np.int64),
                       "mos": np.concatenate(mos).astype(np.uint32)[perm], "conf": np.concatenate(conf, 0).astype(np.float32)[perm]})
    return frames


def run_reference(frames):
    build.build_native()
    ai_path, _ = build.build_ref()
    for name in ("open3d",):
        sys.modules[name] = types.ModuleType(name)
    import spconv.pytorch  # noqa: F401  (oracle shim)
    import spconv.pytorch.utils as su
    if not hasattr(su, "PointToVoxel"):
        su.PointToVoxel = object
    pkg = types.ModuleType("models"); pkg.__path__ = []
    util = types.ModuleType("models.utils"); util.__path__ = [os.path.dirname(ai_path)]
    sys.modules["models"], sys.modules["models.utils"] = pkg, util
    spec = importlib.util.spec_from_file_location("ref_refine", os.path.join(REF, "scripts", "refine.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tmp = tempfile.mkdtemp(prefix="refine_golden_")
    cwd = os.getcwd()
    try:
        os.makedirs(os.path.join(tmp, "config"))
        shutil.copy(os.path.join(REF, "config", "semantic-kitti-mos.yaml"), os.path.join(tmp, "config"))
        seq = os.path.join(tmp, "data", "08")
        os.makedirs(os.path.join(seq, "velodyne"))
        Tr = np.array([[0.0, -1.0, 0.0, 0.1], [0.0, 0.0, -1.0, -0.2], [1.0, 0.0, 0.0, 0.3], [0, 0, 0, 1.0]])
        lines = []
        out = {k: os.path.join(tmp, "preb_out", "InsMOS", k, "sequences", "08", "predictions") for k in ("bbox_preb", "mos_preb", "confidence")}
        for d in out.values():
            os.makedirs(d)
        for f, fr in enumerate(frames):
            fr["scan"].tofile(os.path.join(seq, "velodyne", "%06d.bin" % f))
            cam = Tr @ velo_pose(f) @ np.linalg.inv(Tr)
            lines.append(" ".join("%.12e" % v for v in cam[:3].reshape(-1)))
            np.save(os.path.join(out["bbox_preb"], "%06d.npy" % f), {"pred_boxes": fr["boxes"], "pred_labels": fr["labels"],
                                                                    "pred_scores": np.ones(len(fr["labels"]), np.float32)})
            fr["mos"].tofile(os.path.join(out["mos_preb"], "%06d.label" % f))
            np.save(os.path.join(out["confidence"], "%06d.npy" % f), fr["conf"])
        open(os.path.join(seq, "poses.txt"), "w").write("\n".join(lines) + "\n")
        open(os.path.join(seq, "calib.txt"), "w").write("Tr: " + " ".join("%.12e" % v for v in Tr[:3].reshape(-1)) + "\n")
        os.chdir(tmp)
        mod.main(os.path.join(tmp, "data"), "valid")
        poses = mod.get_lidar_pose(os.path.join(seq, "poses.txt"), os.path.join(seq, "calib.txt"))
        res = [np.fromfile(os.path.join(tmp, "preb_out_refine", "mos_preb", "sequences", "08", "predictions", "%06d.label" % f),
                           dtype=np.int32) for f in range(len(frames))]
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return poses, res


def main():
    rng = np.random.default_rng(2024)
    frames = make_sequence(rng)
    poses, res = run_reference(frames)
    data = {"poses": poses, "n_frames": np.int64(F)}
    changed = 0
    for f, fr in enumerate(frames):
        for k, v in fr.items():
            data["%s%d" % (k, f)] = v
        data["out%d" % f] = res[f]
        changed += int((res[f] != fr["mos"].astype(np.int32)).sum())
    path = os.path.join(HERE, "refine_small.npz")
    np.savez_compressed(path, **data)
    print("frames", F, "points/frame", len(frames[0]["scan"]), "labels changed by the refinement:", changed,
          "per frame:", [int((res[f] != frames[f]["mos"].astype(np.int32)).sum()) for f in range(F)], "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
