"""ORACLE (test infrastructure, CPU): restatement of the steps either side of the forward path.

N1  input staging        scripts/predict_mos.py:114-159 (DemoDataset.__getitem__), :161-166 (transform_point_cloud),
                         :174-179 (timestamp_tensor)
N2  output labelling     scripts/predict_mos.py:440-454 (mask / softmax / confidence / argmax), :279-283 (to_original_labels)
Pinned by tests/golden/io_small.npz, produced by calling the reference's own functions (tests/golden/make_golden_io.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F


def scan_transforms(poses):
    """inv(to_pose) @ from_pose for every scan, newest scan = target frame (predict_mos.py:132-136,162)."""
    to_inv = np.linalg.inv(poses[-1])
    return [to_inv @ p for p in poses]


def stage_scans(scans, poses, dt_pred, transform=True):
    """list of [Ni,4] float32 (x,y,z,intensity) + 4x4 float64 poses -> [sum Ni, 5] float32 (x,y,z,intensity,t)."""
    n = len(scans)
    out = []
    for i, pcd in enumerate(scans):
        pcd = pcd.copy()
        if transform:
            T = np.linalg.inv(poses[-1]) @ poses[i]                            # :162
            xyz1 = np.hstack([pcd[:, :3], np.ones((pcd.shape[0], 1))]).T       # float64 (:164)
            pcd[:, :3] = (T @ xyz1).T[:, :3]                                   # rounded to float32 on assignment (:135)
        t = round((i - n + 1) * dt_pred, 3)                                     # :147-148
        tt = torch.from_numpy(pcd)[:, :4]
        out.append(torch.hstack([tt, t * torch.ones((tt.shape[0], 1))]))       # :176-178
    return torch.cat(out, dim=0).numpy()


def mos_labels(logits, learning_ignore, learning_map_inv):
    """logits [N,C] float32 -> (labels int32 [N] in original ids, confidence [N,C-1] float32)."""
    x = np.array(logits, dtype=np.float32, copy=True)
    ignore_index = [k for k, ign in learning_ignore.items() if ign]
    x[:, ignore_index] = -float("inf")                                         # :441
    p = F.softmax(torch.from_numpy(x), dim=1)                                  # :444
    conf = p.numpy()[:, 1:]                                                     # :446-447
    lab = torch.argmax(p, axis=1).long().numpy()                                # :451-452
    orig = lab.copy()
    for k, v in learning_map_inv.items():                                       # :279-283
        orig[lab == k] = v
    return orig.reshape(-1).astype(np.int32), conf
