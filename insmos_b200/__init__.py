"""insmos_b200: B200-native (sm_100a) implementation of the InsMOS sparse-voxel forward path.

    import insmos_b200
    insmos_b200.install()                 # MinkowskiEngine / spconv / pytorch_lightning / models import names
    from models.models import InsMOSNet   # the reference's entry point, served by libinsmos_b200.so
"""
from .net import install_compat as install  # noqa: F401

__version__ = "0.1.0"
