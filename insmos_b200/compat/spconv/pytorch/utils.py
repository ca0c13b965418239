"""spconv.pytorch.utils: PointToVoxel, gather_features_by_pc_voxel_id (voxel_generate.py:19-27, spconv_unet.py:410)."""
import torch

from insmos_b200 import ops


class PointToVoxel:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel,
                 device=torch.device("cuda")):
        self.vsize = [float(v) for v in vsize_xyz]
        self.range = [float(v) for v in coors_range_xyz]
        self.grid = [int(round((self.range[3 + d] - self.range[d]) / self.vsize[d])) for d in range(3)]
        self.num_point_features = num_point_features
        self.max_num_voxels = int(max_num_voxels)
        self.max_num_points_per_voxel = int(max_num_points_per_voxel)
        self.last = None                       # full result of the last call (CoordSet, fused mean ...)

    def _run(self, pc, want_voxels=True):
        r = ops.voxelize3d(pc.float().contiguous(), self.range, self.vsize, self.grid, self.max_num_voxels,
                           self.max_num_points_per_voxel, want_voxels=want_voxels)
        self.last = r
        return r

    def generate_voxel_with_id(self, pc, clear_voxels=True, empty_mean=False):
        r = self._run(pc)
        return r["voxels"], r["set"].coords[:, 1:4], r["num_points"], r["pc_voxel_id"].to(torch.int64)

    def __call__(self, pc, clear_voxels=True, empty_mean=False):
        r = self._run(pc)
        return r["voxels"], r["set"].coords[:, 1:4], r["num_points"]


def gather_features_by_pc_voxel_id(seg_res_features, pc_voxel_id, invalid_value=0):
    if invalid_value != 0:
        raise NotImplementedError
    from insmos_b200 import autograd as _ag
    gather = _ag.gather_rows if _ag.needs_grad(seg_res_features) else ops.gather_rows
    return gather(seg_res_features.contiguous(), pc_voxel_id.to(torch.int32).contiguous())
