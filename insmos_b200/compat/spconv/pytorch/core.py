import torch

from insmos_b200 import ops


class IndiceData:
    """what spconv stores under an indice_key: both coordinate sets, geometry, and the rule books
    (forward pairs, and lazily the swapped pairs used by SparseInverseConv3d)."""

    def __init__(self, in_set, out_set, ksize, stride, padding, in_shape, out_shape, subm):
        self.in_set, self.out_set = in_set, out_set
        self.ksize, self.stride, self.padding = ksize, stride, padding
        self.in_shape, self.out_shape, self.subm = in_shape, out_shape, subm
        self._fwd = None
        self._inv = None

    def forward_rulebook(self):
        if self._fwd is None:
            spec = ops.spec_sp_subm(self.ksize) if self.subm else ops.spec_sp_conv(self.ksize, self.stride, self.padding)
            self._fwd = ops.build_rulebook(self.out_set, self.in_set, spec, step=(1, 1, 1))
        return self._fwd

    def inverse_rulebook(self):
        if self._inv is None:
            self._inv = ops.build_rulebook(self.in_set, self.out_set, ops.spec_sp_inverse(self.ksize, self.stride, self.padding))
        return self._inv


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False, coordset=None, **kwargs):
        if not features.is_cuda:
            raise RuntimeError("insmos_b200 spconv: CUDA tensors required (no CPU fallback)")
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = batch_size
        self.indice_dict = {} if indice_dict is None else indice_dict
        self.benchmark = benchmark
        self._coordset = coordset

    @property
    def coordset(self):
        """hash table over the indices (built once; reused by every layer on this index set)."""
        if self._coordset is None:
            cs, _ = ops.unique_coords(self.indices.to(torch.int32).contiguous())
            if cs.n != self.indices.shape[0]:
                raise ValueError("SparseConvTensor: duplicate indices")
            self._coordset = cs
        return self._coordset

    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, indice_dict=self.indice_dict,
                             benchmark=self.benchmark, coordset=self._coordset)
        return t

    def find_indice_pair(self, key):
        return self.indice_dict.get(key) if key is not None else None

    def dense(self, channels_first=True):
        if self.batch_size != 1:
            raise NotImplementedError("dense(): batch_size 1 on the InsMOS path (spconv_unet.py:282)")
        D, H, W = self.spatial_shape
        if torch.is_grad_enabled() and self.features.requires_grad:
            # training: index_put keeps the graph (the scatter kernel has no backward; this is plumbing, one call per step)
            ind = self.indices.long()
            out = self.features.new_zeros((self.features.shape[1], D, H, W))
            out[:, ind[:, 1], ind[:, 2], ind[:, 3]] = self.features.t()
            out = out.unsqueeze(0)
            return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()
        out = ops.dense_scatter(self.features, self.indices.to(torch.int32).contiguous(), D, H, W).unsqueeze(0)
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()

    @property
    def spatial_size(self):
        s = 1
        for v in self.spatial_shape:
            s *= v
        return s
