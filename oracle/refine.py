"""CPU restatement of the instance-level refinement step, scripts/refine.py:169-302 (TEST INFRASTRUCTURE; see
oracle/__init__.py).  One `step` per frame, following the reference line by line -- including its quirks:
  * the per-instance attribute vector is the predicted box row itself (x,y,z,dx,dy,dz,yaw) whose LAST entry (the yaw) is
    overwritten by the moving flag (refine.py:222-227);
  * only cars (label 1, column 0 of the instance ids) are considered (refine.py:206);
  * the bottom-up relabelling for highly dynamic scenes only happens during the first `instance_window` frames
    (refine.py:240-250), the tracking over the window afterwards (refine.py:260-292).
Pinned by tests/golden/refine_small.npz, produced by running the reference's own scripts/refine.py on a synthetic sequence
(tests/golden/make_golden_refine.py)."""
import numpy as np

from . import native

OUT_GROUND = 0.03                      # refine.py:196
INSTANCE_WINDOW = 5                    # refine.py:168


def transform_point_cloud(pts, from_pose, to_pose):
    """refine.py:124-129"""
    transformation = np.linalg.inv(to_pose) @ from_pose
    xyz1 = np.hstack([pts, np.ones((pts.shape[0], 1))]).T
    return (transformation @ xyz1).T[:, :3]


def matches(center, attr, attr_pre):
    """refine.py:272-273: box match between the current instance (centre moved into the past frame) and a past instance"""
    return (abs(center[0] - attr_pre[0]) < 1 and abs(center[1] - attr_pre[1]) < 1 and abs(center[2] - attr_pre[2]) < 0.5 and
            abs(attr[3] - attr_pre[3]) < 0.3 and abs(attr[4] - attr_pre[4]) < 0.3 and abs(attr[5] - attr_pre[5]) < 0.3)


class Refiner:
    def __init__(self, instance_window=INSTANCE_WINDOW):
        self.window, self.instance_window = [], instance_window

    def step(self, frame_idx, scan, pred_boxes, pred_labels, mos_label, moving_confidence, lidar_pose):
        """scan [N,4] f32; pred_boxes [nb,7], pred_labels [nb]; mos_label [N] in {1 static, 2 moving} (refine.py:180-182);
        moving_confidence [N,2]; lidar_pose: all poses [F,4,4].  Returns the refined labels [N] int32 (1/2)."""
        W = self.instance_window
        mos_label = mos_label.astype(np.int32).copy()
        pred_boxes = np.array(pred_boxes, copy=True)
        if frame_idx < 9:                                               # refine.py:176-177
            moving_confidence = np.zeros([mos_label.shape[0], 2])
        boxes = np.concatenate((pred_boxes, np.asarray(pred_labels).reshape(-1, 1)), axis=1)
        index = native.find_point_in_instance_bbox_with_yaw(scan, boxes, OUT_GROUND)
        moving_car_num, car_idx = 0, -1
        car_idx_list, car_idx_moving_list, car_all_list, attribute_list = [], [], [], []
        for instance_idx in range(len(pred_labels)):
            if pred_labels[instance_idx] == 1:
                pts_idx = np.where(index[:, 0] == instance_idx + 1)[0]
                n_pts = len(pts_idx)
                if n_pts != 0:
                    n_moving = len(np.where(mos_label[pts_idx] == 2)[0])
                    conf = moving_confidence[pts_idx][:, 1]
                    n_conf = len(np.where(conf >= 0.00001)[0])
                    car_idx += 1
                    car_all_list.append(pts_idx)
                    attr = pred_boxes[instance_idx]                     # a view: the yaw slot becomes the moving flag
                    attr[-1] = 1 if (n_moving / n_pts) > 0.6 else 0
                    attribute_list.append(attr)
                    if (n_moving / n_pts) > 0.3:
                        moving_car_num += 1
                    if (n_moving / n_pts) > 0.001:
                        car_idx_list.append(car_idx)
                    if (n_conf / conf.shape[0]) > 0.5:
                        car_idx_moving_list.append(car_idx)
        if frame_idx != 0:
            if moving_car_num >= 3:
                for c in car_idx_list:
                    if frame_idx < W:
                        mos_label[car_all_list[c]] = 2
                    attribute_list[c][-1] = 1
            if moving_car_num >= 5:
                for c in car_idx_moving_list:
                    if frame_idx < W:
                        mos_label[car_all_list[c]] = 2
                    attribute_list[c][-1] = 1
        else:
            if moving_car_num >= 5:
                for c in car_idx_list:
                    mos_label[car_all_list[c]] = 2
                for c in car_idx_moving_list:
                    mos_label[car_all_list[c]] = 2
        self.window.append(attribute_list)
        if frame_idx >= W:
            assert len(self.window) == W + 1
            current = self.window[-1].copy()
            for attr in current:
                find_flag = moving_flag = 0
                for i in range(W):
                    c = transform_point_cloud(attr[0:3].reshape(-1, 3), lidar_pose[frame_idx], lidar_pose[frame_idx - i - 1]).reshape(-1)
                    for attr_pre in self.window[W - 1 - i]:
                        if matches(c, attr, attr_pre):
                            find_flag += 1
                            if attr_pre[-1] == 1:
                                moving_flag += 1
                            break
                if find_flag == 5:
                    if moving_flag > 3:
                        attr[-1] = 1
                else:
                    if (moving_flag > 1) or (moving_flag > 0 and moving_car_num >= 3):
                        attr[-1] = 1
            for j in range(len(current)):
                if current[j][-1] == 1:
                    mos_label[car_all_list[j]] = 2
                if current[j][-1] == 0 and len(current) > 6:
                    mos_label[car_all_list[j]] = 1
            self.window.pop(0)
        return mos_label
