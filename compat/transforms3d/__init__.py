"""Import shim for the subset of transforms3d the reference's scripts use (quaternions w,x,y,z)."""
from . import quaternions  # noqa: F401
