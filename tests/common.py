"""Shared test inputs (seeded, synthetic) and golden-case definitions."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# toy intrinsics of the reference demo (data/toy_example/intrinsics.txt:1), used by every synthetic pair
K_TOY = [[549.7, 0.0, 268.7], [0.0, 549.7, 351.8], [0.0, 0.0, 1.0]]

GOLDEN_CASES = {
    # small: 15x14 token grid (N=210), every tensor stored in full
    "vits_small": dict(variant="vits", it_matches=4, it_ransac=16, batch=2, height=210, width=196,
                       weight_seed=0, data_seed=5, rng_seed=3, stride=1),
    "vitb_small": dict(variant="vitb", it_matches=2, it_ransac=8, batch=1, height=224, width=182,
                       weight_seed=1, data_seed=6, rng_seed=4, stride=1),
    # the reference's default backbone (ViT-L/14: 24 blocks, 16 heads, D = 1024); 15x13 grid: N = 195, N*N odd
    "vitl_small": dict(variant="vitl", it_matches=2, it_ransac=8, batch=1, height=210, width=182,
                       weight_seed=2, data_seed=8, rng_seed=5, stride=1),
    # BASELINE config 2 at full size (N=1938): N x N tensors stored with stride 17
    "vits_720x540": dict(variant="vits", it_matches=8, it_ransac=64, batch=1, height=720, width=540,
                         weight_seed=0, data_seed=7, rng_seed=11, stride=17),
    # BASELINE config 3's model at full size: ViT-B/14, 1024 hypotheses (16x64), two pairs (T = 1939 attention tails,
    # 128x256 persistent GEMM tiles, batched matcher) — N x N tensors strided by 17
    "vitb_720x540": dict(variant="vitb", it_matches=16, it_ransac=64, batch=2, height=720, width=540,
                         weight_seed=3, data_seed=9, rng_seed=12, stride=17),
    # the reference's default model at full size: ViT-L/14, 2000 hypotheses (20x100)
    "vitl_720x540": dict(variant="vitl", it_matches=20, it_ransac=100, batch=1, height=720, width=540,
                         weight_seed=4, data_seed=10, rng_seed=13, stride=17),
}


def synthetic_pair(batch, height, width, seed, device="cpu"):
    """Two uniform-random RGB images in [0,1] plus toy intrinsics (BASELINE.md §3 'Inputs')."""
    g = torch.Generator().manual_seed(seed)
    im0 = torch.rand(batch, 3, height, width, generator=g)
    im1 = torch.rand(batch, 3, height, width, generator=g)
    K = torch.tensor(K_TOY, dtype=torch.float32)[None].repeat(batch, 1, 1)
    return {"image0": im0.to(device), "image1": im1.to(device),
            "K_color0": K.clone().to(device), "K_color1": K.clone().to(device)}


def load_golden(name):
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    with np.load(path) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a, b):
    """Relative Frobenius error ||a-b|| / ||b|| in fp64 (the 'relative' of DESIGN.md tolerances)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def rotation_angle_deg(R1, R2):
    """Geodesic angle between rotations, fp64, via 2*asin(||R1-R2||_F / (2*sqrt(2))) (acos of the
    trace quantises at ~0.03 deg in fp32)."""
    d = (R1.detach().double().cpu() - R2.detach().double().cpu()).flatten(-2).norm(dim=-1)
    return torch.rad2deg(2 * torch.asin((d / (2 * 2 ** 0.5)).clamp(max=1.0)))
