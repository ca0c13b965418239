"""`spconv.pytorch` operator surface re-provided over libinsmos_b200 (sm_100a CUDA, C ABI).

Not spconv: a from-scratch implementation of the symbols the InsMOS forward path touches
(call sites models/backbones_3d/spconv_unet.py:13-18,71-106,120-208,284-410,
models/backbones_3d/voxel_generate.py:6,19-27, models/backbones_2d/height_compression.py:26) with the
same names, arguments, parameter layout (`.weight [Cout,kz,ky,kx,Cin]`) and `indice_key` semantics.
Row order is first-occurrence / creation order (spconv's CPU semantics; its GPU order is hash order).
CUDA only.
"""
from .core import SparseConvTensor
from .modules import SparseModule, SparseSequential
from .conv import SparseConvolution, SubMConv3d, SparseConv3d, SparseInverseConv3d
from . import conv, utils  # noqa: F401

__all__ = ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SubMConv3d", "SparseConv3d",
           "SparseInverseConv3d", "conv", "utils"]
