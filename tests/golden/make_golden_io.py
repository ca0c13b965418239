"""Generate tests/golden/io_small.npz by calling the REFERENCE's own staging / labelling code.

Run in the development container only (needs /root/reference):   python tests/golden/make_golden_io.py

What runs: scripts/predict_mos.py::DemoDataset.transform_point_cloud / timestamp_tensor (unbound, exactly the loop of
__getitem__ :130-150) and the label post-step of main() (:440-454 with to_original_labels), imported from the reference
file over the same shims as make_golden.py.  The semantic config is the reference's config/semantic-kitti-mos.yaml.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (sets sys.path for the shims + reference, compiles oracle/_ref)

REF = "/root/reference"


def main():
    make_golden.load_reference_model()                                   # registers the fake iou3d module, Array_Index path
    sys.path.insert(0, os.path.join(REF, "scripts"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import predict_mos as pm
    finally:
        os.chdir(cwd)
    sem = yaml.safe_load(open(os.path.join(REF, "config", "semantic-kitti-mos.yaml")))
    rng = np.random.default_rng(21)
    n_scans, dt = 4, 0.1
    scans = [np.concatenate([rng.uniform(-60, 60, (n, 2)), rng.uniform(-3, 2, (n, 1)), rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
             for n in (1500, 1733, 1200, 1601)]
    poses = []
    for i in range(n_scans):                                             # KITTI-like ego motion: yaw drift + forward translation
        a = 0.02 * i + rng.normal(0, 1e-3)
        T = np.eye(4)
        T[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        T[:3, 3] = [1.1 * i + rng.normal(0, 0.01), 0.05 * i, rng.normal(0, 0.01)]
        poses.append(T)
    ds = pm.DemoDataset.__new__(pm.DemoDataset)                          # no files: only the pure methods are used
    staged = []
    for i, pcd in enumerate(scans):
        pcd = pcd.copy()
        pcd[:, :3] = ds.transform_point_cloud(pcd[:, :3], poses[i], poses[-1])      # :132-136
        t = round((i - n_scans + 1) * dt, 3)
        staged.append(ds.timestamp_tensor(torch.from_numpy(pcd)[:, :4], t))          # :141-148
    past = torch.cat(staged, dim=0).numpy()
    # ---- N2: the post step of main() on random logits (incl. exact ties and large magnitudes)
    logits = rng.normal(0, 3, (5000, 3)).astype(np.float32)
    logits[:50, 2] = logits[:50, 1]
    logits[50:60] *= 40
    ignore_index = [k for k, ign in sem["learning_ignore"].items() if ign]
    m = logits.copy()
    m[:, ignore_index] = -float("inf")                                   # :441
    p = F.softmax(torch.from_numpy(m), dim=1)                            # :444
    conf = p.detach().cpu().numpy()[:, 1:]                               # :446-447
    lab = torch.argmax(p, axis=1).long().cpu().numpy()                   # :451-452
    lab = pm.to_original_labels(lab, sem).reshape((-1)).astype(np.int32)  # :453-454
    path = os.path.join(HERE, "io_small.npz")
    np.savez_compressed(path, **{"scan%d" % i: s for i, s in enumerate(scans)}, poses=np.stack(poses), dt=np.float64(dt),
                        past_point_clouds=past, logits=logits, labels=lab, confidence=conf,
                        learning_ignore=np.array([int(sem["learning_ignore"][k]) for k in range(3)]),
                        learning_map_inv=np.array([sem["learning_map_inv"][k] for k in range(3)], dtype=np.int32))
    print("wrote", path, past.shape, lab.shape, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
