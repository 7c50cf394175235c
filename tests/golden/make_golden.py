"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as small .npz files and
this script is committed next to them.  Inputs are not stored: they are a pure function of the seed
(`tests/common.py::synthetic_pair`), and the weights are `mickey_b200.weights.synthetic_state_dict`.

For the stochastic solver the two torch.multinomial draws of the reference
(probabilisticProcrustes.py:231,251) are recorded while it runs and stored with the fixture, so a
test can inject them into the oracle / the CUDA solver and compare everything downstream.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mickey_b200.config import mickey_cfg            # noqa: E402
from mickey_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_harness                        # noqa: E402
from tests.common import synthetic_pair, GOLDEN_CASES  # noqa: E402


def run_case(name, spec):
    cfg = mickey_cfg(spec["variant"], spec["it_matches"], spec["it_ransac"], float16=False)
    sd = synthetic_state_dict(cfg, seed=spec["weight_seed"])
    model = ref_harness.build_reference_model(cfg, sd, variant=spec["variant"])
    data = synthetic_pair(spec["batch"], spec["height"], spec["width"], seed=spec["data_seed"])

    draws = []
    orig = torch.multinomial

    def recording_multinomial(*a, **k):
        out = orig(*a, **k)
        draws.append(out.clone())
        return out

    # the hypothesis scores and the winner (probabilisticProcrustes.py:265,275) are internal to the solver: record the
    # argument and the result of its single torch.argmax(score_k, dim=1) call
    picks = []
    orig_argmax = torch.argmax

    def recording_argmax(inp, *a, **k):
        out = orig_argmax(inp, *a, **k)
        if inp.dim() == 2 and (k.get("dim") == 1 or (a and a[0] == 1)):
            picks.append((inp.clone(), out.clone()))
        return out

    torch.manual_seed(spec["rng_seed"])
    torch.multinomial = recording_multinomial
    torch.argmax = recording_argmax
    try:
        with torch.no_grad():
            R, t = model(data, return_inliers=True)
    finally:
        torch.multinomial = orig
        torch.argmax = orig_argmax
    assert len(draws) == 2 and len(picks) == 1
    st = spec.get("stride", 1)
    out = {
        "kps0": data["kps0"], "kps1": data["kps1"],
        "depth_kp0": data["depth_kp0"], "depth_kp1": data["depth_kp1"],
        "scr0": data["scr0"], "scr1": data["scr1"],
        "dsc0": data["dsc0"][:, :, ::st], "dsc1": data["dsc1"][:, :, ::st],
        "scores": data["scores"][:, ::st, ::st], "kp_scores": data["kp_scores"][:, ::st, ::st],
        "final_scores": data["final_scores"][:, ::st, ::st],
        "scores_rowsum": data["scores"].sum(-1), "final_rowsum": data["final_scores"].double().sum(-1),
        "outer_idx": draws[0].to(torch.int32), "inner_idx": draws[1].to(torch.int16),
        "hyp_scores": picks[0][0], "best": picks[0][1].to(torch.int32),
        "R": R, "t": t, "inliers": data["inliers"],
        "n_inliers_list": torch.tensor([len(x) for x in data["inliers_list"]]),
        "inliers_list0": data["inliers_list"][0],
    }
    out = {k: v.detach().cpu().numpy() for k, v in out.items()}
    path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    assert ref_harness.available(), "needs /root/reference"
    only = sys.argv[1:]
    for name, spec in GOLDEN_CASES.items():
        if not only or name in only:
            run_case(name, spec)
