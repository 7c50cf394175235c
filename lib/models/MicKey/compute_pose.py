"""Drop-in module path of the reference's lib/models/MicKey/compute_pose.py (MickeyRelativePose, :6-60)."""
from mickey_b200.model import MickeyRelativePose  # noqa: F401
