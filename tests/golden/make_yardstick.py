"""How far the reference's OWN released configuration (`FLOAT16: True`, mickey_extractor.py:31-35: fp16 DINOv2
backbone) is from its fp32 path on the golden cases, with the fp32 run's two multinomial draws replayed.

Run in the build container (needs /root/reference):   python tests/golden/make_yardstick.py [case ...]
Writes tests/golden/fp16_yardstick.json.  The GPU parity tests read it: the end-to-end pose of the CUDA path (fp16
tensor-core operands) is graded against the fp32 fixture with the north-star tolerance OR this yardstick, whichever is
larger — the synthetic random-weight problems are ill-conditioned (a pose from 3 sampled points and few inliers), so
the reference itself moves by more than 1e-2 deg / 1e-3 m when only its backbone precision changes.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mickey_b200.config import mickey_cfg            # noqa: E402
from mickey_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_harness                        # noqa: E402
from tests.common import synthetic_pair, GOLDEN_CASES, load_golden, rel_err, rotation_angle_deg  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fp16_yardstick.json")


def run_case(name, spec):
    gold = load_golden(name)
    cfg = mickey_cfg(spec["variant"], spec["it_matches"], spec["it_ransac"], float16=True)
    sd = synthetic_state_dict(cfg, seed=spec["weight_seed"])
    model = ref_harness.build_reference_model(cfg, sd, variant=spec["variant"])
    data = synthetic_pair(spec["batch"], spec["height"], spec["width"], seed=spec["data_seed"])
    replay = [gold["outer_idx"].long(), gold["inner_idx"].long()]
    picks = []
    orig_mn, orig_am = torch.multinomial, torch.argmax

    def replay_multinomial(*a, **k):
        return replay.pop(0)

    def recording_argmax(inp, *a, **k):
        out = orig_am(inp, *a, **k)
        if inp.dim() == 2 and (k.get("dim") == 1 or (a and a[0] == 1)):
            picks.append((inp.clone(), out.clone()))
        return out

    torch.multinomial, torch.argmax = replay_multinomial, recording_argmax
    try:
        with torch.no_grad():
            R, t = model(data, return_inliers=True)
    finally:
        torch.multinomial, torch.argmax = orig_mn, orig_am
    st = spec.get("stride", 1)
    hyp, best = picks[0]
    e = {
        "dsc": max(rel_err(data[k][:, :, ::st], gold[k]) for k in ("dsc0", "dsc1")),
        "depth": max(rel_err(data[k], gold[k]) for k in ("depth_kp0", "depth_kp1")),
        "kps_px": max(float((data[k] - gold[k]).abs().max()) for k in ("kps0", "kps1")),
        "scr": max(rel_err(data[k], gold[k]) for k in ("scr0", "scr1")),
        "scores": rel_err(data["scores"][:, ::st, ::st], gold["scores"]),
        "final_scores": rel_err(data["final_scores"][:, ::st, ::st], gold["final_scores"]),
        "hyp_scores": rel_err(hyp, gold["hyp_scores"]),
        "same_winner": float(bool((best.int() == gold["best"]).all())),
        "rot_deg": float(rotation_angle_deg(R, gold["R"]).max()),
        "t_m": float((t - gold["t"]).abs().max()),
        "inliers": rel_err(data["inliers"].reshape(-1), gold["inliers"].reshape(-1)),
    }
    print(name, {k: float(f"{v:.3e}") for k, v in e.items()}, flush=True)
    return e


if __name__ == "__main__":
    assert ref_harness.available(), "needs /root/reference"
    only = sys.argv[1:]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name, spec in GOLDEN_CASES.items():
        if not only or name in only:
            res[name] = run_case(name, spec)
            with open(OUT, "w") as f:
                json.dump(res, f, indent=1, sort_keys=True)
