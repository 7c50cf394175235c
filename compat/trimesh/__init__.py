"""Import-only shim: trimesh is imported at module top by the reference's lib/utils/visualization.py but only used by the
optional 3-D rendering (--generate_3D_vis); any attribute access raises."""


def __getattr__(name):
    raise ImportError("trimesh is not installed in this image; 3-D visualisation is unavailable")
