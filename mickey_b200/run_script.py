"""Run one of the reference's own scripts (demo_inference.py, submission.py) UNCHANGED against mickey_b200.

    python -m mickey_b200.run_script /path/to/mickey/demo_inference.py --config ... --checkpoint ...

The reference's `lib`, `config` packages have no __init__.py, i.e. they are namespace packages.  Putting this
repo in front of the reference root on sys.path makes `lib.models.builder`, `lib.models.MicKey.compute_pose`,
`lib.utils.data` and `config.default` resolve to the CUDA-backed mirrors here, while everything this repo does not
provide (datasets, visualisation, benchmarks) still resolves to the reference's own files.  `compat/` supplies
import shims for the third-party modules missing from this image (pytorch_lightning, yacs, transforms3d, matplotlib,
pyrender, trimesh).  Nothing of the reference is copied or modified.
"""
import os
import runpy
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    ref_root = os.path.dirname(script)
    shims = os.path.join(REPO, "compat")
    needed = [m for m in ("pytorch_lightning", "yacs", "transforms3d", "matplotlib", "pyrender", "trimesh")]
    sys.path[:0] = [REPO] + ([shims] if any(_missing(m) for m in needed) else []) + [ref_root]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


def _missing(mod):
    import importlib.util
    try:
        return importlib.util.find_spec(mod) is None
    except Exception:      # noqa: BLE001
        return True


if __name__ == "__main__":
    main()
