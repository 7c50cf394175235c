"""The two callers the drop-in must satisfy, executed on the GPU.  The reference tree does not exist
on the GPU box, so the CALLERS are restated here line by line — demo_inference.py:79-138 and submission.py:32-96 —
while everything they import (`config.default.cfg`, `lib.models.builder.build_model`, `lib.utils.data`,
`lib.datasets.datamodules.DataModule`, `transforms3d`) resolves to this repo exactly as it does for the unmodified
scripts under `python -m mickey_b200.run_script`."""
import os
import sys
import zipfile

import numpy as np
import pytest
import torch

from tests.common import ROOT, K_TOY

pytestmark = pytest.mark.gpu
if __import__("importlib").util.find_spec("transforms3d") is None:
    sys.path.insert(0, os.path.join(ROOT, "compat"))

from mickey_b200.config import mickey_cfg                     # noqa: E402
from mickey_b200.weights import synthetic_checkpoint          # noqa: E402
from tools.make_synthetic_mapfree import make_tree, texture   # noqa: E402


def _write_model_files(tmp_path, variant="vits", im=2, ir=8):
    cfg = mickey_cfg(variant, im, ir)
    keep = {k: cfg[k] for k in ("MODEL", "MICKEY", "FEATURE_MATCHER", "PROCRUSTES")}
    from mickey_b200.config import CfgNode
    (tmp_path / "model.yaml").write_text(CfgNode(keep).dump())
    torch.save(synthetic_checkpoint(cfg, seed=0), tmp_path / "mickey.ckpt")      # like a real one: no dinov2 tensors
    return str(tmp_path / "model.yaml"), str(tmp_path / "mickey.ckpt")


def test_demo_inference_call_sequence(tmp_path):
    import cv2
    from config.default import cfg as default_cfg
    from lib.models.builder import build_model
    from lib.datasets.utils import correct_intrinsic_scale
    cfg = default_cfg.clone()
    config, ckpt = _write_model_files(tmp_path)
    rng = np.random.default_rng(0)
    for n in ("im0.jpg", "im1.jpg"):
        cv2.imwrite(str(tmp_path / n), texture(rng, 360, 270))
    (tmp_path / "intrinsics.txt").write_text("im0.jpg 590.0 590.0 134.6 176.1 270 360\nim1.jpg 590.0 590.0 134.6 176.1 270 360\n")
    resize = (196, 224)

    # ---- demo_inference.py:12-29 read_color_image, :31-47 read_intrinsics (restated)
    def read_color_image(path, resize):
        image = cv2.cvtColor(cv2.imread(str(path), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
        if resize is not None:
            image = cv2.resize(image, resize)
        return (torch.from_numpy(image).float().permute(2, 0, 1) / 255).unsqueeze(0)

    def read_intrinsics(path, resize):
        Ks = {}
        for line in open(path).readlines():
            parts = line.strip().split(" ")
            fx, fy, cx, cy, W, H = map(float, parts[1:])
            K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
            Ks[parts[0]] = correct_intrinsic_scale(K, resize[0] / W, resize[1] / H).numpy()
        return Ks

    # ---- demo_inference.py:79-138
    device = torch.device("cuda:0")
    cfg.merge_from_file(config)
    model = build_model(cfg, checkpoint=ckpt)
    im0 = read_color_image(tmp_path / "im0.jpg", resize).to(device)
    im1 = read_color_image(tmp_path / "im1.jpg", resize).to(device)
    K = read_intrinsics(tmp_path / "intrinsics.txt", resize)
    data = {"image0": im0, "image1": im1,
            "K_color0": torch.from_numpy(K["im0.jpg"]).unsqueeze(0).to(device),
            "K_color1": torch.from_numpy(K["im1.jpg"]).unsqueeze(0).to(device)}
    model(data, return_inliers=True)
    batch_id = 0
    gh, gw = 224 // 14, 196 // 14
    assert data["depth0_map"][batch_id].shape == (1, gh, gw) and data["depth1_map"][batch_id].shape == (1, gh, gw)
    invalid = (data["depth0_map"][batch_id] < 0.001).cpu()[0]                    # :121
    assert invalid.shape == (gh, gw)
    assert data["scr0"][batch_id].shape == (1, gh * gw) and data["image0"][batch_id].shape == (3, 224, 196)   # prepare_score_map inputs
    assert data["R"].shape == (1, 3, 3) and data["t"].shape == (1, 1, 3) and data["inliers"].numel() == 1
    assert model.e2e_Procrustes.num_samples_matches == 2048                     # :136
    lst = data["inliers_list"][batch_id]                                          # create_point_cloud_from_inliers input
    assert lst.dim() == 2 and lst.shape[1] in (5, 7)
    P = np.eye(4)
    P[:3, :3] = data["R"][batch_id][np.newaxis].detach().cpu().numpy()
    P[:3, 3] = data["t"][batch_id].detach().cpu().numpy().reshape(-1)
    assert np.isfinite(P).all()


def test_submission_call_sequence(tmp_path):
    from collections import defaultdict
    from config.default import cfg as default_cfg
    from lib.datasets.datamodules import DataModule
    from lib.models.builder import build_model
    from lib.utils.data import data_to_model_device
    from transforms3d.quaternions import mat2quat
    from mickey_b200 import submission as mksub
    cfg = default_cfg.clone()
    config, ckpt = _write_model_files(tmp_path)
    make_tree(str(tmp_path / "data"), "val", scenes=2, queries=11, seed=3, width=196, height=224)

    # ---- submission.py:71-96 eval()
    cfg.merge_from_file(os.path.join(ROOT, "config", "datasets", "mapfree.yaml"))
    cfg.merge_from_file(config)
    cfg.DATASET.DATA_ROOT, cfg.DATASET.HEIGHT, cfg.DATASET.WIDTH = str(tmp_path / "data"), 224, 196
    cfg.TRAINING.BATCH_SIZE, cfg.TRAINING.NUM_WORKERS = 4, 0
    dataloader = DataModule(cfg, drop_last_val=False).val_dataloader()
    model = build_model(cfg, ckpt)

    # ---- submission.py:32-61 predict(), the per-pair host loop as the reference writes it
    results_dict, packed_all, names_all = defaultdict(list), [], []
    torch.manual_seed(0)
    for data in dataloader:
        data = data_to_model_device(data, model)
        with torch.no_grad():
            R_batched, t_batched = model(data)
        packed_all.append(mksub.pack_poses(R_batched, t_batched, data["inliers"]).clone())
        for i_batch in range(len(data["scene_id"])):
            R = R_batched[i_batch].unsqueeze(0).detach().cpu().numpy()
            t = t_batched[i_batch].reshape(-1).detach().cpu().numpy()
            inliers = data["inliers"][i_batch].item()
            scene = data["scene_id"][i_batch]
            query_img = data["pair_names"][1][i_batch]
            names_all.append((scene, query_img))
            if np.isnan(R).any() or np.isnan(t).any() or np.isinf(t).any():
                continue
            results_dict[scene].append(mksub.Pose(image_name=query_img, q=mat2quat(R).reshape(-1), t=t.reshape(-1), inliers=inliers))
    mksub.save_submission(results_dict, tmp_path / "submission.zip")
    with zipfile.ZipFile(tmp_path / "submission.zip") as z:
        assert sorted(z.namelist()) == ["pose_s00000.txt", "pose_s00001.txt"]
        ref_lines = {n: z.read(n).decode().split("\n") for n in z.namelist()}
    assert sum(len(v) for v in ref_lines.values()) == 6                          # 2 scenes x frames 0, 5, 10
    first = ref_lines["pose_s00000.txt"][0].split(" ")
    assert first[0] == "seq1/frame_00000.jpg" and len(first) == 9 and all(np.isfinite(float(x)) for x in first[1:])

    # ---- the batched device writer (mk_pose_to_submission) produces the same lines from the same poses
    recs = mksub.poses_to_records(torch.cat(packed_all))
    ours = defaultdict(list)
    for rec, (scene, q) in zip(recs, names_all):
        for pose in mksub.records_to_poses(rec[None], [q]):
            ours[scene].append(str(pose))
    for scene, lines in ours.items():
        assert lines == ref_lines[f"pose_{scene}.txt"]
