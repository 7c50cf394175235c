"""BASELINE configs[0] (C1): the reference's own `demo_inference.py` call sequence on its toy_example pair, on the
reference's PyTorch CPU path, single pair, no GPU.  Runs in the build container only (needs /root/reference): the
UNMODIFIED reference model (ViT-L/14 backbone, 2000 hypotheses: its released configuration, fp32 because fp16 CPU
kernels are not what a CPU user would run) with seeded random-init weights, images read and resized exactly as
demo_inference.py:12-29,94-98 does.  Writes profiles/r02_c1_reference_cpu.json.

    python tools/time_c1_reference.py [n_timed]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200.config import mickey_cfg            # noqa: E402
from mickey_b200.io import read_color_image          # noqa: E402
from mickey_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_harness                        # noqa: E402
from lib.datasets.utils import correct_intrinsic_scale  # noqa: E402

TOY = os.path.join(ref_harness.REF_ROOT, "data", "toy_example")


def main():
    assert ref_harness.available(), "needs /root/reference"
    n_timed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cfg = mickey_cfg("vitl", 20, 100, float16=False)
    model = ref_harness.build_reference_model(cfg, synthetic_state_dict(cfg, seed=0), variant="vitl")
    resize = (540, 720)
    im0 = read_color_image(os.path.join(TOY, "im0.jpg"), resize)[None]
    im1 = read_color_image(os.path.join(TOY, "im1.jpg"), resize)[None]
    Ks = {}
    for line in open(os.path.join(TOY, "intrinsics.txt")):
        if "#" in line or not line.strip():
            continue
        parts = line.strip().split(" ")
        fx, fy, cx, cy, W, H = map(float, parts[1:])
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
        Ks[parts[0]] = correct_intrinsic_scale(K, resize[0] / W, resize[1] / H)
    durs = []
    for i in range(1 + n_timed):
        data = {"image0": im0, "image1": im1, "K_color0": Ks["im0.jpg"][None], "K_color1": Ks["im1.jpg"][None]}
        torch.manual_seed(i)
        t0 = time.perf_counter()
        with torch.no_grad():
            R, t = model(data, return_inliers=False)
        durs.append(time.perf_counter() - t0)
    timed = durs[1:]
    out = {"config": "BASELINE configs[0]: demo_inference.py toy_example pair on the reference PyTorch CPU path (single pair, no GPU)",
           "model": "unmodified reference, ViT-L/14, 2000 hypotheses (20x100), fp32, seeded random-init weights",
           "host": f"build container, {os.cpu_count()} vCPU, torch {torch.__version__}, threads {torch.get_num_threads()}",
           "image": list(im0.shape), "s_per_pair": sum(timed) / len(timed), "pairs_per_s": len(timed) / sum(timed),
           "warmup_s": durs[0], "timed_s": timed, "pose_finite": bool(torch.isfinite(R).all() and torch.isfinite(t).all())}
    with open(os.path.join(ROOT, "profiles", "r02_c1_reference_cpu.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
