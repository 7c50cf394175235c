"""Synthetic (seeded) MicKey state dicts with the reference's tensor names, and checkpoint plumbing.

There are no pretrained weights on the box (no network), and BASELINE.json asks for random-init
weights of the named architecture.  `synthetic_state_dict` produces a full, strict-loadable state
dict (names/shapes probed from the reference: compute_matches.extractor.{dinov2_vitl14,depth_head,
det_offset,dsc_head,det_head}.* and compute_matches.matcher.matching_mat.dustbin_score)
whose values are a pure function of (tensor name, seed) — so the same weights can be loaded into the
unmodified reference (to make golden fixtures), into the CPU oracle and into the CUDA engine without
ever shipping a checkpoint file.  Values are deliberately non-degenerate (non-unit LayerNorm/BN
statistics, non-zero biases, LayerScale around 1) so that a parity test catches a dropped bias or a
mis-folded BatchNorm.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict

import torch

from .config import VARIANTS, backbone_variant

EXTRACTOR = "compute_matches.extractor."
BACKBONE = EXTRACTOR + "dinov2_vitl14."
DUSTBIN = "compute_matches.matcher.matching_mat.dustbin_score"
HEADS = ("depth_head", "det_offset", "dsc_head", "det_head")
HEAD_OUT = {"depth_head": ("depth", 1), "det_offset": ("xy_offset", 2), "det_head": ("score", 1)}
POS_GRID = 37            # img_size 518 / patch 14 (reference mickey_extractor.py:18)
PATCH = 14


def _gen(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def _normal(name, seed, shape, std, mean=0.0, clip=2.0):
    x = torch.randn(shape, generator=_gen(name, seed), dtype=torch.float32)
    return (x.clamp_(-clip, clip) * std + mean).contiguous()


def _uniform(name, seed, shape, lo, hi):
    return (torch.rand(shape, generator=_gen(name, seed), dtype=torch.float32) * (hi - lo) + lo).contiguous()


def synthetic_state_dict(cfg, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Full fp32 state dict for MickeyRelativePose(cfg) (reference compute_pose.py:6-18)."""
    variant = backbone_variant(cfg)
    D, depth, _ = VARIANTS[variant]
    sd: Dict[str, torch.Tensor] = {}

    def put(name, t):
        sd[name] = t

    # --- DINOv2 backbone (reference dinov2.py:95-152; layers/*.py) ---------------------------------
    b = BACKBONE
    put(b + "cls_token", _normal(b + "cls_token", seed, (1, 1, D), 0.02))
    put(b + "pos_embed", _normal(b + "pos_embed", seed, (1, 1 + POS_GRID * POS_GRID, D), 0.02))
    put(b + "mask_token", torch.zeros(1, D))
    put(b + "patch_embed.proj.weight", _normal(b + "pe.w", seed, (D, 3, PATCH, PATCH), 0.02))
    put(b + "patch_embed.proj.bias", _normal(b + "pe.b", seed, (D,), 0.02))
    for i in range(depth):
        p = f"{b}blocks.{i}."
        for ln in ("norm1", "norm2"):
            put(p + ln + ".weight", _normal(p + ln + ".w", seed, (D,), 0.1, mean=1.0))
            put(p + ln + ".bias", _normal(p + ln + ".b", seed, (D,), 0.02))
        put(p + "attn.qkv.weight", _normal(p + "qkv.w", seed, (3 * D, D), 0.02))
        put(p + "attn.qkv.bias", _normal(p + "qkv.b", seed, (3 * D,), 0.02))
        put(p + "attn.proj.weight", _normal(p + "proj.w", seed, (D, D), 0.02))
        put(p + "attn.proj.bias", _normal(p + "proj.b", seed, (D,), 0.02))
        put(p + "ls1.gamma", _normal(p + "ls1", seed, (D,), 0.1, mean=1.0))
        put(p + "mlp.fc1.weight", _normal(p + "fc1.w", seed, (4 * D, D), 0.02))
        put(p + "mlp.fc1.bias", _normal(p + "fc1.b", seed, (4 * D,), 0.02))
        put(p + "mlp.fc2.weight", _normal(p + "fc2.w", seed, (D, 4 * D), 0.02))
        put(p + "mlp.fc2.bias", _normal(p + "fc2.b", seed, (D,), 0.02))
        put(p + "ls2.gamma", _normal(p + "ls2", seed, (D,), 0.1, mean=1.0))
    put(b + "norm.weight", _normal(b + "norm.w", seed, (D,), 0.1, mean=1.0))
    put(b + "norm.bias", _normal(b + "norm.b", seed, (D,), 0.02))

    # --- four heads (reference mickey_extractor.py:67-251, extractor_utils.py:12-35) ---------------
    kp = cfg["MICKEY"]["KP_HEADS"]
    dims = list(kp["BLOCKS_DIM"])
    last_dim = cfg["MICKEY"]["DSC_HEAD"]["LAST_DIM"]
    use_bn = kp["BN"]
    for head in HEADS:
        hp = EXTRACTOR + head + "."
        chans = [D] + dims
        if head == "dsc_head":
            chans = [D] + dims[:3] + [last_dim]
        if head == "det_head":       # constants the reference registers as frozen parameters (:88-91)
            put(hp + "eps", torch.tensor(1e-16))
            put(hp + "offset_par1", torch.tensor(0.5))
            put(hp + "offset_par2", torch.tensor(2.0))
            put(hp + "ones_kernel", torch.ones(1, 1, 3, 3))
        for r in range(4):
            cin, cout = chans[r], chans[r + 1]
            rp = f"{hp}resblock{r + 1}."
            put(rp + "conv1.weight", _normal(rp + "c1", seed, (cout, cin, 3, 3), math.sqrt(2.0 / (9 * cin))))
            put(rp + "conv2.weight", _normal(rp + "c2", seed, (cout, cout, 3, 3), math.sqrt(1.0 / (9 * cout))))
            if use_bn:
                for bn in ("bn1", "bn2"):
                    put(rp + bn + ".weight", _normal(rp + bn + ".w", seed, (cout,), 0.1, mean=1.0))
                    put(rp + bn + ".bias", _normal(rp + bn + ".b", seed, (cout,), 0.05))
                    put(rp + bn + ".running_mean", _normal(rp + bn + ".m", seed, (cout,), 0.1))
                    put(rp + bn + ".running_var", _uniform(rp + bn + ".v", seed, (cout,), 0.6, 1.4))
                    put(rp + bn + ".num_batches_tracked", torch.tensor(0, dtype=torch.int64))
            if cin != cout:
                put(rp + "shortcut.0.weight", _normal(rp + "sc", seed, (cout, cin, 1, 1), math.sqrt(1.0 / cin)))
        if head in HEAD_OUT:
            nm, oc = HEAD_OUT[head]
            put(hp + nm + ".weight", _normal(hp + nm, seed, (oc, chans[4], 1, 1), math.sqrt(1.0 / chans[4])))
        for li in range(3):          # Transformer_self_att(d_model=128, num_layers=3)
            lp = f"{hp}att_layer.layers.{li}."
            xav = lambda o, i: math.sqrt(6.0 / (i + o))
            for nm, (o, i) in {"q_proj": (128, 128), "k_proj": (128, 128), "v_proj": (128, 128),
                               "merge": (128, 128), "mlp.0": (256, 256), "mlp.2": (128, 256)}.items():
                a = xav(o, i)
                put(lp + nm + ".weight", _uniform(lp + nm, seed, (o, i), -a, a))
            for ln in ("norm1", "norm2"):
                put(lp + ln + ".weight", _normal(lp + ln + ".w", seed, (128,), 0.1, mean=1.0))
                put(lp + ln + ".bias", _normal(lp + ln + ".b", seed, (128,), 0.02))

    put(DUSTBIN, torch.tensor(1.0))
    return sd


def reorder_like(sd: Dict[str, torch.Tensor], reference_keys) -> Dict[str, torch.Tensor]:
    """Return `sd` in the key order of `reference_keys` (strict key-set equality is asserted)."""
    assert set(sd) == set(reference_keys), (
        f"missing: {sorted(set(reference_keys) - set(sd))[:5]} extra: {sorted(set(sd) - set(reference_keys))[:5]}")
    return {k: sd[k] for k in reference_keys}


def synthetic_checkpoint(cfg, seed: int = 0, with_backbone: bool = False) -> dict:
    """A dict shaped like the reference's mickey.ckpt: {'state_dict': ...}.  Real MicKey checkpoints
    omit the frozen DINOv2 tensors (reference model.py:291-298); with_backbone=False mimics that."""
    sd = synthetic_state_dict(cfg, seed)
    if not with_backbone:
        sd = {k: v for k, v in sd.items() if "dinov2" not in k}
    return {"state_dict": sd}
