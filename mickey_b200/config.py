"""Configuration tree for the MicKey hot path.

Mirrors the *surface* of the reference's yacs tree (reference config/default.py:1-141: same key
names, `None` defaults, attribute and item access, `merge_from_file`) without depending on yacs,
which is not installed in this image.  Only the keys the inference hot path reads are interpreted
by mickey_b200; the remaining keys are carried so that reference YAML files merge without a
"non-existent key" error (yacs semantics, reference submission.py:73-74).

One optional key is new: MICKEY.DINOV2.VARIANT ('vits' | 'vitb' | 'vitl').  The reference has no
backbone-variant key (it hard-codes vit_large, mickey_extractor.py:25); when VARIANT is None it is
derived from CHANNEL_DIM (384 / 768 / 1024).
"""
from __future__ import annotations

import ast
import copy
import yaml


def _decode(v):
    """yacs semantics (yacs/config.py `_decode_cfg_value`): a string that parses as a Python literal becomes that
    literal ('None' -> None, '[1, 2]' -> [1, 2]); anything else stays a string."""
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


class CfgNode(dict):
    """dict with attribute access and yacs-like merge semantics."""

    def __init__(self, init=None):
        super().__init__()
        if init:
            for k, v in init.items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_other_cfg(self, other, _path=""):
        for k, v in other.items():
            if k not in self:
                raise KeyError(f"Non-existent config key: {_path}{k}")
            if isinstance(v, dict) and isinstance(self[k], CfgNode):
                self[k].merge_from_other_cfg(v, _path + k + ".")
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else _decode(v)

    def merge_from_file(self, path):
        with open(path, "r") as f:
            loaded = yaml.safe_load(f) or {}
        self.merge_from_other_cfg(loaded)

    def merge_from_list(self, kv):
        assert len(kv) % 2 == 0
        for key, val in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent config key: {key}")
            node[parts[-1]] = _decode(val)

    def dump(self):
        def plain(n):
            return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}
        return yaml.safe_dump(plain(self))


def _tree(spec):
    """spec: {'A.B.C': default, ...} -> nested CfgNode."""
    root = CfgNode()
    for dotted, default in spec.items():
        node = root
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node:
                node[p] = CfgNode()
            node = node[p]
        node[parts[-1]] = default
    return root


# Keys on the hot path (read by mickey_b200) ------------------------------------------------------
_HOT = {
    "MODEL": None, "DEBUG": False,
    "MICKEY.DINOV2.DOWN_FACTOR": None, "MICKEY.DINOV2.CHANNEL_DIM": None,
    "MICKEY.DINOV2.FLOAT16": None, "MICKEY.DINOV2.VARIANT": None,
    "MICKEY.KP_HEADS.BLOCKS_DIM": None, "MICKEY.KP_HEADS.BN": None,
    "MICKEY.KP_HEADS.USE_SOFTMAX": None, "MICKEY.KP_HEADS.USE_DEPTHSIGMOID": None,
    "MICKEY.KP_HEADS.MAX_DEPTH": None, "MICKEY.KP_HEADS.POS_ENCODING": None,
    "MICKEY.DSC_HEAD.LAST_DIM": None, "MICKEY.DSC_HEAD.BLOCKS_DIM": None,
    "MICKEY.DSC_HEAD.BN": None, "MICKEY.DSC_HEAD.NORM_DSC": None,
    "MICKEY.DSC_HEAD.POS_ENCODING": None,
    "FEATURE_MATCHER.TYPE": None, "FEATURE_MATCHER.DUAL_SOFTMAX.TEMPERATURE": None,
    "FEATURE_MATCHER.DUAL_SOFTMAX.USE_DUSTBIN": None,
    "FEATURE_MATCHER.SINKHORN.NUM_IT": None, "FEATURE_MATCHER.SINKHORN.DUSTBIN_SCORE_INIT": None,
    "FEATURE_MATCHER.USE_TRANSFORMER": None, "FEATURE_MATCHER.TOP_KEYPOINTS": False,
    "PROCRUSTES.IT_MATCHES": None, "PROCRUSTES.IT_RANSAC": None,
    "PROCRUSTES.NUM_SAMPLED_MATCHES": None, "PROCRUSTES.NUM_CORR_3D_3D": None,
    "PROCRUSTES.NUM_REFINEMENTS": None, "PROCRUSTES.TH_INLIER": None,
    "PROCRUSTES.TH_SOFT_INLIER": None,
}

# Keys carried only so reference YAMLs merge (training / dataset side; out of scope here) ----------
_CARRIED = {
    **{f"LOSS_CLASS.{k}": None for k in (
        "LOSS_FUNCTION", "SOFT_CLIPPING", "POSE_ERR.MAX_LOSS_VALUE", "POSE_ERR.MAX_LOSS_SOFTVALUE",
        "VCRE.MAX_LOSS_VALUE", "VCRE.MAX_LOSS_SOFTVALUE",
        "GENERATE_HYPOTHESES.SCORE_TEMPERATURE", "GENERATE_HYPOTHESES.IT_MATCHES",
        "GENERATE_HYPOTHESES.IT_RANSAC", "GENERATE_HYPOTHESES.INLIER_3D_TH",
        "GENERATE_HYPOTHESES.INLIER_REF_TH", "GENERATE_HYPOTHESES.NUM_REF_STEPS",
        "GENERATE_HYPOTHESES.NUM_CORR_3d3d", "CURRICULUM_LEARNING.TRAIN_CURRICULUM",
        "CURRICULUM_LEARNING.TRAIN_WITH_TOPK", "CURRICULUM_LEARNING.TOPK_INIT",
        "CURRICULUM_LEARNING.TOPK", "NULL_HYPOTHESIS.ADD_NULL_HYPOTHESIS",
        "NULL_HYPOTHESIS.TH_OUTLIERS", "SAMPLER.NUM_SAMPLES_MATCHES")},
    "PROCRUSTES_TRAINING.MAX_CORR_DIST": None, "PROCRUSTES_TRAINING.REFINE": False,
    **{f"DATASET.{k}": None for k in (
        "DATA_SOURCE", "SCENES", "DATA_ROOT", "SEED", "NPZ_ROOT", "MIN_OVERLAP_SCORE",
        "MAX_OVERLAP_SCORE", "CONSECUTIVE_PAIRS", "FRAME_RATE", "AUGMENTATION_TYPE",
        "PAIRS_TXT.TRAIN", "PAIRS_TXT.VAL", "PAIRS_TXT.TEST", "HEIGHT", "WIDTH")},
    "DATASET.BLACK_WHITE": False, "DATASET.PAIRS_TXT.ONE_NN": False,
    **{f"TRAINING.{k}": None for k in (
        "BATCH_SIZE", "NUM_WORKERS", "NUM_GPUS", "SAMPLER", "N_SAMPLES_SCENE",
        "SAMPLE_WITH_REPLACEMENT", "LR", "LR_STEP_INTERVAL", "LR_STEP_GAMMA", "VAL_INTERVAL",
        "VAL_BATCHES", "LOG_INTERVAL", "EPOCHS")},
    "TRAINING.GRAD_CLIP": 0.0,
}


def default_cfg() -> CfgNode:
    return _tree({**_HOT, **_CARRIED})


VARIANTS = {
    # name: (embed_dim, depth, heads)     reference dinov2.py:306-342
    "vits": (384, 12, 6),
    "vitb": (768, 12, 12),
    "vitl": (1024, 24, 16),
}


def backbone_variant(cfg) -> str:
    v = cfg["MICKEY"]["DINOV2"].get("VARIANT", None)
    if v is not None:
        assert v in VARIANTS, f"unknown backbone variant {v}"
        return v
    dim = cfg["MICKEY"]["DINOV2"]["CHANNEL_DIM"]
    for name, (d, _, _) in VARIANTS.items():
        if d == dim:
            return name
    raise ValueError(f"cannot derive backbone variant from CHANNEL_DIM={dim}")


def mickey_cfg(variant="vitl", it_matches=20, it_ransac=100, float16=True) -> CfgNode:
    """The MicKey inference configuration (values of the reference's released config,
    config/MicKey/curriculum_learning.yaml:1-32,89-96) for a given backbone / hypothesis budget.

    BASELINE configs: C2 = mickey_cfg('vits', 8, 64); C3 = mickey_cfg('vitb', 16, 64);
    repo default = mickey_cfg('vitl', 20, 100).
    """
    cfg = default_cfg()
    cfg.MODEL = "MicKey"
    d = cfg.MICKEY.DINOV2
    d.DOWN_FACTOR, d.CHANNEL_DIM, d.FLOAT16, d.VARIANT = 14, VARIANTS[variant][0], float16, variant
    k = cfg.MICKEY.KP_HEADS
    k.BLOCKS_DIM, k.BN, k.USE_SOFTMAX, k.USE_DEPTHSIGMOID = [512, 256, 128, 64], True, True, False
    k.MAX_DEPTH, k.POS_ENCODING = 60, True
    s = cfg.MICKEY.DSC_HEAD
    s.LAST_DIM, s.BLOCKS_DIM, s.BN, s.NORM_DSC, s.POS_ENCODING = 128, [512, 256, 128], True, True, True
    m = cfg.FEATURE_MATCHER
    m.TYPE, m.USE_TRANSFORMER = "DualSoftmax", False
    m.DUAL_SOFTMAX.TEMPERATURE, m.DUAL_SOFTMAX.USE_DUSTBIN = 0.1, True
    m.SINKHORN.NUM_IT, m.SINKHORN.DUSTBIN_SCORE_INIT = 10, 1.0
    p = cfg.PROCRUSTES
    p.IT_MATCHES, p.IT_RANSAC, p.NUM_SAMPLED_MATCHES, p.NUM_CORR_3D_3D = it_matches, it_ransac, 2048, 3
    p.NUM_REFINEMENTS, p.TH_INLIER, p.TH_SOFT_INLIER = 4, 0.15, 0.3
    cfg.DATASET.HEIGHT, cfg.DATASET.WIDTH = 720, 540
    return cfg
