"""Instance-level refinement of the MOS labels after the forward path: the host-side mirror of
scripts/refine.py:169-302 over the device kernels of csrc/refine.cu.

Per frame the reference loads the scan, the predicted boxes, the predicted labels and confidences from disk, finds the
points of every predicted box on the host (Array_Index.find_point_in_instance_bbox_with_yaw, OpenMP over boxes), counts
moving points per car with numpy `where` scans, and rewrites the labels of whole instances.  Here the points, labels and
confidences stay on the GPU: one kernel pair assigns instance ids, one kernel reduces the per-instance statistics (integer
atomics: deterministic), the tracking logic over the 5-frame window -- a few dozen boxes -- runs on the host on a [nb,3]
int32 read-back, and one kernel applies the per-instance label decisions.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from insmos_b200 import ops
from insmos_b200._lib import call

OUT_GROUND = 0.03                      # refine.py:196
INSTANCE_WINDOW = 5                    # refine.py:168
CONF_THRESH = 0.00001                  # refine.py:217


def point_instance_ids(points, boxes8, out_ground=OUT_GROUND, n_class=3):
    """points [n,>=3] f32 CUDA, boxes8 [nb,8] f32 CUDA -> ids [n,n_class] int32 (box index + 1 per class column)."""
    points = ops._req(points, torch.float32, "point_instance_ids")
    n, stride = points.shape
    nb = int(boxes8.shape[0])
    ids = torch.empty((max(n, 1), n_class), dtype=torch.int32, device=points.device)
    first = torch.empty(max(nb, 1), dtype=torch.int32, device=points.device)
    if nb:
        boxes8 = ops._req(boxes8, torch.float32, "point_instance_ids")
    call("insmos_point_instance_ids", ops._p(points), n, stride, ops._p(boxes8) if nb else None, nb, float(out_ground),
         ops._p(ids), n_class, ops._p(first), ops._stream())
    return ids[:n]


def instance_stats(ids, col, labels, conf, nb, moving_label=2, conf_thresh=CONF_THRESH):
    """-> int32 [nb,3] on the device: points, moving points, confident points of every instance of class column `col`."""
    ids = ops._req(ids, torch.int32, "instance_stats")
    labels = ops._req(labels, torch.int32, "instance_stats")
    n, ncls = ids.shape
    stats = torch.empty((max(nb, 1), 3), dtype=torch.int32, device=ids.device)
    cs = 0
    if conf is not None:
        conf = ops._req(conf, torch.float32, "instance_stats")
        cs = conf.shape[1] if conf.dim() == 2 else 1
    cptr = None if conf is None else C.c_void_p(conf.data_ptr() + (4 if cs == 2 else 0))     # column 1 of [n,2] (refine.py:216)
    call("insmos_instance_stats", ops._p(ids), ncls, col, n, ops._p(labels), moving_label, cptr, cs, float(conf_thresh), nb,
         ops._p(stats), ops._stream())
    return stats[:nb]


def relabel_instances(ids, col, new_label, labels):
    """labels[j] = new_label[id] (in place) where id = ids[j,col] > 0 and new_label[id] >= 0; new_label int32 [nb+1] CUDA."""
    ids = ops._req(ids, torch.int32, "relabel_instances")
    n, ncls = ids.shape
    call("insmos_relabel_instances", ops._p(ids), ncls, col, n, ops._p(new_label), int(new_label.shape[0]) - 1, ops._p(labels),
         ops._stream())
    return labels


def _transform(pts, from_pose, to_pose):
    T = np.linalg.inv(to_pose) @ from_pose
    return (T @ np.hstack([pts, np.ones((pts.shape[0], 1))]).T).T[:, :3]


def _matches(c, a, b):
    return (abs(c[0] - b[0]) < 1 and abs(c[1] - b[1]) < 1 and abs(c[2] - b[2]) < 0.5 and
            abs(a[3] - b[3]) < 0.3 and abs(a[4] - b[4]) < 0.3 and abs(a[5] - b[5]) < 0.3)


class InstanceRefiner:
    """step(frame_idx, scan, pred_boxes, pred_labels, mos_label, moving_confidence, poses) -> refined labels (CUDA int32).
    scan [N,4] f32 CUDA; pred_boxes [nb,7] / pred_labels [nb] (host or device); mos_label [N] int32 CUDA in {1,2};
    moving_confidence [N,2] f32 CUDA or None; poses [F,4,4] float64 (host)."""

    def __init__(self, instance_window=INSTANCE_WINDOW):
        self.window, self.W = [], instance_window

    def step(self, frame_idx, scan, pred_boxes, pred_labels, mos_label, moving_confidence, poses):
        dev, W = scan.device, self.W
        boxes = torch.as_tensor(pred_boxes, dtype=torch.float32).cpu().numpy().copy()
        labels_b = torch.as_tensor(pred_labels).cpu().numpy().astype(np.int64)
        nb = len(labels_b)
        mos_label = mos_label.to(torch.int32).clone()
        boxes8 = torch.from_numpy(np.concatenate([boxes, labels_b.reshape(-1, 1).astype(np.float32)], 1)).to(dev)
        ids = point_instance_ids(scan, boxes8)
        conf = None if (frame_idx < 9 or moving_confidence is None) else moving_confidence           # refine.py:176-177
        stats = instance_stats(ids, 0, mos_label, conf, nb).cpu().numpy()                            # the one read-back
        # ---- bottom-up (refine.py:198-258); `cars` = instances with label 1 and at least one point, in box order
        cars, attrs, idx_list, idx_moving_list, moving_car_num = [], [], [], [], 0
        for b in range(nb):
            if labels_b[b] == 1 and stats[b, 0] != 0:
                n_pts, n_mov, n_conf = (int(v) for v in stats[b])
                a = boxes[b]
                a[-1] = 1 if (n_mov / n_pts) > 0.6 else 0                                            # the yaw slot becomes the flag
                cars.append(b)
                attrs.append(a)
                if (n_mov / n_pts) > 0.3:
                    moving_car_num += 1
                if (n_mov / n_pts) > 0.001:
                    idx_list.append(len(cars) - 1)
                if (n_conf / n_pts) > 0.5:
                    idx_moving_list.append(len(cars) - 1)
        new_label = np.full(nb + 1, -1, dtype=np.int32)                                              # per-instance decision
        if frame_idx != 0:
            for lst, need in ((idx_list, 3), (idx_moving_list, 5)):
                if moving_car_num >= need:
                    for c in lst:
                        if frame_idx < W:
                            new_label[cars[c] + 1] = 2
                        attrs[c][-1] = 1
        elif moving_car_num >= 5:
            for c in idx_list + idx_moving_list:
                new_label[cars[c] + 1] = 2
        # ---- tracking over the window + top-down (refine.py:260-294)
        self.window.append(attrs)
        if frame_idx >= W:
            assert len(self.window) == W + 1
            for a in attrs:
                find_flag = moving_flag = 0
                for i in range(W):
                    c = _transform(a[0:3].reshape(-1, 3).astype(np.float64), poses[frame_idx], poses[frame_idx - i - 1]).reshape(-1)
                    for prev in self.window[W - 1 - i]:
                        if _matches(c, a, prev):
                            find_flag += 1
                            moving_flag += int(prev[-1] == 1)
                            break
                if find_flag == 5:
                    if moving_flag > 3:
                        a[-1] = 1
                elif moving_flag > 1 or (moving_flag > 0 and moving_car_num >= 3):
                    a[-1] = 1
            for j, a in enumerate(attrs):
                if a[-1] == 1:
                    new_label[cars[j] + 1] = 2
                if a[-1] == 0 and len(attrs) > 6:
                    new_label[cars[j] + 1] = 1
            self.window.pop(0)
        if nb and (new_label >= 0).any():
            relabel_instances(ids, 0, torch.from_numpy(new_label).to(dev), mos_label)
        return mos_label
