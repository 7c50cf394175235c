"""Generate a synthetic Map-free tree (BASELINE config 5: "random-init weights, synthetic GT"):

    python tools/make_synthetic_mapfree.py --root data --split val --scenes 4 --queries 20

<root>/<split>/s0000N/{intrinsics.txt, poses.txt, seq0/frame_00000.jpg, seq1/frame_000NN.jpg}   in the formats
lib/datasets/mapfree.py parses (reference mapfree.py:31-69,94-103): 540x720 JPEGs of smooth random textures, the
toy intrinsics of the reference demo, an identity pose for the seq0 keyframe and random small motions for the queries.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def random_quaternion(rng, max_deg):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, max_deg))
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * axis])


def texture(rng, h, w):
    import cv2
    small = rng.integers(0, 256, size=(h // 12, w // 12, 3), dtype=np.uint8)
    img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
    noise = rng.integers(-12, 13, size=img.shape, dtype=np.int16)
    return np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def make_tree(root, split="val", scenes=2, queries=10, seed=0, width=540, height=720, frame_step=1):
    import cv2
    rng = np.random.default_rng(seed)
    out = []
    for s in range(scenes):
        d = os.path.join(root, split, f"s{s:05d}")
        os.makedirs(os.path.join(d, "seq0"), exist_ok=True)
        os.makedirs(os.path.join(d, "seq1"), exist_ok=True)
        names = ["seq0/frame_00000.jpg"] + [f"seq1/frame_{q * frame_step:05d}.jpg" for q in range(queries)]
        with open(os.path.join(d, "intrinsics.txt"), "w") as f:
            for n in names:
                f.write(f"{n} 590.0 590.0 269.2 352.2 {width} {height}\n")
        with open(os.path.join(d, "poses.txt"), "w") as f:
            f.write("# frame qw qx qy qz tx ty tz (world to camera)\n")
            for i, n in enumerate(names):
                q = np.array([1.0, 0, 0, 0]) if i == 0 else random_quaternion(rng, 25.0)
                t = np.zeros(3) if i == 0 else rng.uniform(-1.5, 1.5, size=3)
                f.write(n + " " + " ".join(f"{v:.6f}" for v in np.concatenate([q, t])) + "\n")
        for n in names:
            cv2.imwrite(os.path.join(d, n), texture(rng, height, width), [cv2.IMWRITE_JPEG_QUALITY, 90])
        out.append(d)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default="data")
    ap.add_argument("--split", default="val")
    ap.add_argument("--scenes", type=int, default=2)
    ap.add_argument("--queries", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    for d in make_tree(args.root, args.split, args.scenes, args.queries, args.seed):
        print(d)
