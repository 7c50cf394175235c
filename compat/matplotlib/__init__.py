"""Import shim: the reference imports matplotlib only for colour maps (training_utils.py:4, visualization.py:10).
`cm.get_cmap(name)` returns a callable mapping [0,1] -> RGBA with a smooth blue-green-yellow-red ramp."""
import numpy as np


class _Cmap:
    def __init__(self, name="viridis"):
        self.name = name

    def __call__(self, x, bytes=False):
        x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
        r = np.clip(1.5 - np.abs(4.0 * x - 3.0), 0, 1)
        g = np.clip(1.5 - np.abs(4.0 * x - 2.0), 0, 1)
        b = np.clip(1.5 - np.abs(4.0 * x - 1.0), 0, 1)
        rgba = np.stack([r, g, b, np.ones_like(x)], axis=-1)
        return (rgba * 255).astype(np.uint8) if bytes else rgba


class _CM:
    @staticmethod
    def get_cmap(name="viridis", lut=None):
        return _Cmap(name)

    def __getattr__(self, name):
        return _Cmap(name)


cm = _CM()
colormaps = {"viridis": _Cmap("viridis")}


def use(*a, **k):
    return None
