"""Import shim for `import matplotlib.pyplot as plt` (colour maps only)."""
from . import cm, _Cmap  # noqa: F401


def get_cmap(name="viridis", lut=None):
    return _Cmap(name)
