"""Harness that imports the UNMODIFIED reference (nianticlabs/mickey) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the oracle restatement in
oracle/mickey_oracle.py against the real reference code and (b) generate the golden fixtures under
tests/golden/.  /root/reference does not exist on the GPU box, so nothing on the GPU path imports
this module; `available()` says whether the reference tree is present.

What it has to work around (none of it touches /root/reference):
  * pytorch_lightning / matplotlib / cv2-free imports -> tiny stub modules in sys.modules
    (reference imports them at compute_pose.py:1 and training_utils.py:3-4)
  * torch.hub.load_state_dict_from_url (mickey_extractor.py:15-17, no network) -> returns the
    state dict of a freshly constructed ViT of the requested variant
  * the hard-coded `vit_large` symbol (mickey_extractor.py:25) -> rebound to vit_small/base/large
"""
import os
import sys
import types
import contextlib

import torch

REF_ROOT = os.environ.get("MICKEY_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "models", "MicKey"))


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        try:
            import pytorch_lightning  # noqa: F401
        except Exception:
            pl = types.ModuleType("pytorch_lightning")

            class LightningModule(torch.nn.Module):
                pass

            pl.LightningModule = LightningModule
            sys.modules["pytorch_lightning"] = pl
    for name in ("matplotlib", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)


@contextlib.contextmanager
def _ref_on_path():
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)


def build_reference_model(cfg, state_dict=None, variant="vits"):
    """Construct the reference MickeyRelativePose and (optionally) load `state_dict` into it.

    cfg: mickey_b200.config.CfgNode (attribute + item access, like yacs).
    """
    assert available(), "reference tree not present"
    _install_stubs()
    # our repo also has a top-level `lib` package (the drop-in module paths); make sure the
    # reference's `lib` namespace is the one imported here.
    saved = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.")}
    for k in saved:
        del sys.modules[k]
    with _ref_on_path():
        import lib.models.MicKey.modules.DINO_modules.dinov2 as dinov2
        import lib.models.MicKey.modules.mickey_extractor as mx
        from lib.models.MicKey.compute_pose import MickeyRelativePose

        factory = {"vits": dinov2.vit_small, "vitb": dinov2.vit_base, "vitl": dinov2.vit_large}[variant]
        mx.vit_large = factory
        orig_hub = torch.hub.load_state_dict_from_url
        torch.hub.load_state_dict_from_url = lambda *a, **k: factory(
            img_size=518, patch_size=14, init_values=1.0, ffn_layer="mlp", block_chunks=0).state_dict()
        try:
            model = MickeyRelativePose(cfg)
        finally:
            torch.hub.load_state_dict_from_url = orig_hub
        ref_modules = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.")}
    # keep the reference modules reachable from the model but restore our own `lib`
    for k in ref_modules:
        del sys.modules[k]
    sys.modules.update(saved)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model
