import numpy as np
import torch

from oracle import sp


class PointToVoxel:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel, device=None):
        self.vs, self.rng = list(vsize_xyz), list(coors_range_xyz)
        self.mv, self.mp = max_num_voxels, max_num_points_per_voxel

    def generate_voxel_with_id(self, pc):
        vox, coords, num, ids = sp.point_to_voxel(pc.detach().cpu().numpy(), self.vs, self.rng, self.mp, self.mv)
        return vox, torch.from_numpy(coords), torch.from_numpy(num), torch.from_numpy(ids)


def gather_features_by_pc_voxel_id(seg, ids):
    return sp.gather_features_by_pc_voxel_id(seg, ids.numpy())
