"""models.models.InsMOSNet -- the reference's entry point (models/models.py:27-59), served by the B200 path."""
from insmos_b200.net.model import InsMOSNet, InsMOS_Model, ClassificationMetrics  # noqa: F401
