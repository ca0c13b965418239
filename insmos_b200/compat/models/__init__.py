"""`models` import name of the reference (predict_mos.py:23: `import models.models as models`)."""
