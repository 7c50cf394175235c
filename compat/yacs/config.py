"""Import shim: `from yacs.config import CfgNode` (yacs is not installed in this image).
Put <repo>/compat on sys.path to use it."""
from mickey_b200.config import CfgNode  # noqa: F401
