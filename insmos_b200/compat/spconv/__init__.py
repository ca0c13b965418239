"""`spconv` import name re-provided over libinsmos_b200 (see spconv/pytorch/__init__.py)."""
__version__ = "2.3.6+insmos_b200"
from . import pytorch  # noqa: F401
