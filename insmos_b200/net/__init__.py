"""Host-side mirror of the reference's model graph for the hot path (models.models.InsMOSNet).

Same module tree / parameter names and forward() contract as the reference,
so Lightning checkpoints load strictly and scripts/predict_mos.py runs unchanged; the graph itself
is written for the B200 path: BatchNorm/ReLU/residual fused into convolution epilogues, rule books
shared per geometry, voxeliser + mean-VFE fused, NMS sweep and box membership on device.
"""
import os
import sys

COMPAT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat")


def install_compat(prepend=True):
    """make `MinkowskiEngine`, `spconv`, `pytorch_lightning`, `easydict`, `models` importable from
    insmos_b200/compat (the reference imports them by these names: predict_mos.py:4,19,23)."""
    if COMPAT_DIR not in sys.path:
        if prepend:
            sys.path.insert(0, COMPAT_DIR)
        else:
            sys.path.append(COMPAT_DIR)
    return COMPAT_DIR
