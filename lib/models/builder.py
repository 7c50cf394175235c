"""Drop-in module path of the reference's lib/models/builder.py (build_model, :5-20).
The implementation lives in mickey_b200.model (CUDA-backed)."""
from mickey_b200.model import build_model  # noqa: F401
