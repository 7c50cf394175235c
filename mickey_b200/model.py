"""Host-side mirror of the reference's model surface for the inference hot path.

Same names, arguments, data-dict keys and error behaviour as
    lib/models/builder.py:5-20            build_model(cfg, checkpoint)
    lib/models/MicKey/compute_pose.py:6-60 MickeyRelativePose
    .../modules/compute_correspondences.py ComputeCorrespondences
    .../modules/utils/probabilisticProcrustes.py e2eProbabilisticProcrustesSolver
so that demo_inference.py / submission.py run against it unchanged — but every stage executes inside
libmickey_b200.so (hand-written sm_100a CUDA).  The nn.Module tree below only *stores* the parameters
under the reference's state-dict names (so a real mickey.ckpt loads with strict=True); it has no
forward arithmetic of its own.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .config import backbone_variant
from .engine import Engine, PATCH
from .weights import synthetic_state_dict

_BUFFER_SUFFIXES = ("running_mean", "running_var", "num_batches_tracked")


def synthetic_backbone_allowed() -> bool:
    """Benchmarks and tests run on seeded random-init weights (BASELINE.json: there is no network for the DINOv2
    download the reference performs, mickey_extractor.py:15-17); they opt in with MICKEY_SYNTHETIC_BACKBONE=1.
    Anyone else loading a real mickey.ckpt (which omits the frozen DINOv2 tensors) without supplying DINOv2 weights
    would silently get poses from a random backbone, so that is an error."""
    import os
    return os.environ.get("MICKEY_SYNTHETIC_BACKBONE", "0") == "1"


class _ParamTree(nn.Module):
    """A module whose only job is to hold tensors under dotted names."""

    def add(self, dotted: str, tensor: torch.Tensor):
        parts = dotted.split(".")
        node = self
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _ParamTree())
            node = node._modules[p]
        leaf = parts[-1]
        if leaf in _BUFFER_SUFFIXES:
            node.register_buffer(leaf, tensor.clone())
        else:
            node.register_parameter(leaf, nn.Parameter(tensor.clone(), requires_grad=False))

    def eval(self):            # the reference calls .eval()/.train() on the heads; nothing to switch here
        return super().eval()


class e2eProbabilisticProcrustesSolver:
    """Test-time metric pose solver (reference probabilisticProcrustes.py:5-20, 183-348), CUDA-backed."""

    def __init__(self, cfg, owner: "MickeyRelativePose"):
        p = cfg.PROCRUSTES
        self.it_RANSAC = p.IT_RANSAC
        self.it_matches = p.IT_MATCHES
        self.num_samples_matches = p.NUM_SAMPLED_MATCHES
        self.num_corr_3d_3d = p.NUM_CORR_3D_3D
        self.num_refinements = p.NUM_REFINEMENTS
        self.th_inlier = p.TH_INLIER
        self.th_soft_inlier = p.TH_SOFT_INLIER
        self._owner = owner

    def estimate_pose_vectorized(self, batch, return_inliers=False, outer_idx=None, inner_idx=None, seed=None):
        eng = self._owner._engine()
        final = batch["final_scores"].detach().float()      # a padded-pitch view goes to the kernels as it is
        B, N, _ = final.shape
        kps = torch.cat([batch["kps0"], batch["kps1"]], 0).detach().float().contiguous()
        depth = torch.cat([batch["depth_kp0"], batch["depth_kp1"]], 0).detach().float().contiguous()
        if seed is None:
            seed = int(torch.randint(1, 2 ** 62, (1,)).item())     # follows torch.manual_seed like the reference
        res = eng.solve(final, kps, depth, batch["K_color0"], batch["K_color1"], seed, outer_idx=outer_idx,
                        inner_idx=inner_idx, want_extras=True)
        pose = res["pose"]
        R = pose[:, :9].reshape(B, 3, 3).contiguous()
        t = pose[:, 9:12].reshape(B, 1, 3).contiguous()
        inliers = pose[:, 12:13].contiguous()
        batch["_solver"] = res
        if not return_inliers:
            return R, t, inliers
        # inlier list of the winning sampled set (probabilisticProcrustes.py:305-327): rows
        # [x0, y0, x1, y1, score, d0, d1] sorted by score, assembled from the kernel's mask (plumbing only)
        n_s = self.num_samples_matches
        sets = res["best_set"].long()
        cells = res["sampled_idx"].long()[sets]                           # [B, n_s]
        mask = res["inlier_mask"] > 0.5
        i0, i1 = torch.div(cells, N, rounding_mode="trunc"), cells % N
        bidx = torch.arange(B, device=final.device)[:, None].expand(-1, n_s)
        w = final[bidx, i0, i1]
        rows = torch.cat([batch["kps0"][bidx, :, i0], batch["kps1"][bidx, :, i1], w[..., None],
                          batch["depth_kp0"][bidx, :, i0], batch["depth_kp1"][bidx, :, i1]], dim=-1)
        zero_pose = bool((res["status"].item() & 7) != 0)
        out = []
        for b in range(B):
            if zero_pose:
                out.append(torch.zeros([0, 5]))
                continue
            rb = rows[b][mask[b]]
            out.append(rb[torch.argsort(rb[:, 4], descending=True)])
        return R, t, inliers, out


class ComputeCorrespondences(nn.Module):
    """Extraction + matching (reference compute_correspondences.py:6-92), CUDA-backed."""

    def __init__(self, cfg, owner: "MickeyRelativePose"):
        super().__init__()
        object.__setattr__(self, "_owner", owner)
        self.dsc_dim = cfg["MICKEY"]["DSC_HEAD"]["LAST_DIM"]
        self.down_factor = cfg["MICKEY"]["DINOV2"]["DOWN_FACTOR"]
        self.extractor = _ParamTree()
        self.matcher = _ParamTree()

    def forward(self, data):
        eng = self._owner._engine()
        im0, im1 = data["image0"], data["image1"]
        B = im0.shape[0]
        if im0.shape != im1.shape:
            raise ValueError(f"image0 {tuple(im0.shape)} and image1 {tuple(im1.shape)} must have the same shape (both images "
                             "of a batch are extracted in one call)")
        images = torch.cat([im0, im1], dim=0)
        kps, depth, scr, dsc = eng.extract(images)
        N = kps.shape[-1]
        H, W = eng.geo
        gh, gw = H // PATCH, W // PATCH
        scores, kp_scores, final = eng.match(B, N, lean=bool(getattr(self._owner, "lean_outputs", False)))
        data["kps0_shape"], data["kps1_shape"] = [gh, gw], [gh, gw]
        data["depth0_map"] = depth[:B].reshape(B, 1, gh, gw)
        data["depth1_map"] = depth[B:].reshape(B, 1, gh, gw)
        data["down_factor"] = self.down_factor
        data["kps0"], data["kps1"] = kps[:B], kps[B:]
        data["depth_kp0"], data["depth_kp1"] = depth[:B], depth[B:]
        data["scr0"], data["scr1"] = scr[:B], scr[B:]
        data["dsc0"], data["dsc1"] = dsc[:B], dsc[B:]
        if scores is not None:
            data["scores"] = scores
            data["kp_scores"] = kp_scores
        data["_final_scores_fused"] = final
        return data["kps0"], data["dsc0"], data["kps1"], data["dsc1"]


class MickeyRelativePose(nn.Module):
    """Metric relative pose between two images (reference compute_pose.py:6-60)."""

    def __init__(self, cfg, dinov2_weights=None):
        """dinov2_weights: optional DINOv2 state dict (native names: 'cls_token', 'blocks.0.attn.qkv.weight', ...)
        or a path to one — the stand-in for the reference's download (mickey_extractor.py:15-17; there is no network
        here).  Also read from $MICKEY_DINOV2_WEIGHTS.  Without it the backbone keeps seeded random-init values."""
        super().__init__()
        if cfg.MODEL is not None and cfg.MODEL != "MicKey":
            raise NotImplementedError()
        self.cfg = cfg
        self.variant = backbone_variant(cfg)
        self.compute_matches = ComputeCorrespondences(cfg, self)
        self.e2e_Procrustes = e2eProbabilisticProcrustesSolver(cfg, self)
        # parameter storage under the reference's names; values are placeholders until a checkpoint loads
        # (the reference downloads DINOv2 here, mickey_extractor.py:15-17 — there is no network on the box)
        for name, t in synthetic_state_dict(cfg, seed=0).items():
            assert name.startswith("compute_matches.")
            self.compute_matches.__getattr__(name.split(".")[1]).add(".".join(name.split(".")[2:]), t)
        import os
        dinov2_weights = dinov2_weights or os.environ.get("MICKEY_DINOV2_WEIGHTS")
        self.__dict__["_backbone_is_synthetic"] = dinov2_weights is None
        if dinov2_weights is not None:
            if isinstance(dinov2_weights, str):
                dinov2_weights = torch.load(dinov2_weights, map_location="cpu")
            own = self.state_dict()
            pre = "compute_matches.extractor.dinov2_vitl14."
            missing = [k for k in own if k.startswith(pre) and k[len(pre):] not in dinov2_weights]
            if missing:
                raise KeyError(f"DINOv2 weights lack {missing[:3]} ...")
            super().load_state_dict({**own, **{pre + k: v for k, v in dinov2_weights.items() if pre + k in own}})
        self.__dict__["_eng"] = None
        self.__dict__["_eng_version"] = -1
        self.__dict__["_param_version"] = 0
        self.is_eval_model(True)

    # -- checkpoint plumbing (compute_pose.py:39-48, builder.py:11-13) ---------------------------------------
    def on_load_checkpoint(self, checkpoint):
        """compute_pose.py:39-48: the checkpoint's (absent) DINOv2 tensors are filled from the module's own backbone.
        The reference's own backbone is the downloaded pretrained DINOv2; here it must have been supplied
        (`dinov2_weights=` / $MICKEY_DINOV2_WEIGHTS) unless synthetic weights were explicitly allowed."""
        if self.__dict__.get("_backbone_is_synthetic", True) and not synthetic_backbone_allowed():
            raise RuntimeError(
                "MickeyRelativePose holds seeded RANDOM DINOv2 weights: the reference downloads the pretrained ViT at this "
                "point (mickey_extractor.py:15-17) and there is no network here.  Pass dinov2_weights= (or set "
                "$MICKEY_DINOV2_WEIGHTS to a dinov2_vit*14_pretrain.pth), or set MICKEY_SYNTHETIC_BACKBONE=1 to run on "
                "synthetic weights on purpose (benchmarks / tests).")
        own = self.compute_matches.state_dict()
        for k in own:
            if "dinov2" in k:
                checkpoint["state_dict"]["compute_matches." + k] = own[k]

    def load_state_dict(self, state_dict, strict=True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self.__dict__["_param_version"] += 1
        return res

    def _apply(self, fn, *a, **k):
        res = super()._apply(fn, *a, **k)
        self.__dict__["_param_version"] += 1
        return res

    def is_eval_model(self, is_eval):
        return None        # BatchNorm is folded at load time: always eval semantics

    def _engine_pool(self):
        """pipeline_depth engines (default 1).  With depth 2, consecutive forward() calls alternate between two engines
        that own separate workspaces, CUDA graphs, RNG state and streams but share one copy of the packed weights."""
        depth = int(getattr(self, "pipeline_depth", 1))
        first = self._engine()
        pool = self.__dict__.setdefault("_pool", [])
        if len(pool) != depth - 1 or self.__dict__.get("_pool_version") != self._eng_version or (pool and pool[0].device != first.device):
            pool.clear()
            for _ in range(depth - 1):
                e = Engine(self.cfg, first.device, side_stream=True)
                e.load_state_dict(None, share_with=first)
                pool.append(e)
            self.__dict__["_pool_version"] = self._eng_version
        if depth > 1 and first.stream is None:
            first.stream = torch.cuda.Stream(device=first.device)
        engines = [first] + pool
        for e in engines:
            e.assume_inputs_ready = bool(getattr(self, "assume_inputs_ready", False))
        return engines

    def _engine(self) -> Engine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("mickey_b200 runs on CUDA only: move the model to a B200 with .cuda() "
                               "(there is no CPU fallback)")
        if self._eng is None or self._eng.device != dev:
            self.__dict__["_eng"] = Engine(self.cfg, dev)
            self.__dict__["_eng_version"] = -1
        if self._eng_version != self._param_version:
            self._eng.load_state_dict(self.state_dict())
            self.__dict__["_eng_version"] = self._param_version
        return self._eng

    # -- the hot path ------------------------------------------------------------------------------------------
    def _inlier_list(self, data, st, B, N):
        """probabilisticProcrustes.py:305-327 from the kernel's winning set + hard-inlier mask (plumbing)."""
        n_s = self.e2e_Procrustes.num_samples_matches
        cells = st["sampled_idx"].long()[st["best_set"].long()]
        mask = st["inlier_mask"] > 0.5
        i0, i1 = torch.div(cells, N, rounding_mode="trunc"), cells % N
        bidx = torch.arange(B, device=cells.device)[:, None].expand(-1, n_s)
        w = data["final_scores"][bidx, i0, i1]
        rows = torch.cat([data["kps0"][bidx, :, i0], data["kps1"][bidx, :, i1], w[..., None],
                          data["depth_kp0"][bidx, :, i0], data["depth_kp1"][bidx, :, i1]], dim=-1)
        if int(st["status"].item()) & 7:
            return [torch.zeros([0, 5])] * B
        out = []
        for b in range(B):
            rb = rows[b][mask[b]]
            out.append(rb[torch.argsort(rb[:, 4], descending=True)])
        return out

    @torch.no_grad()
    def forward(self, data, return_inliers=False):
        """One C call (mk_forward) per batch, replayed from a CUDA graph after the first two calls.
        `self.static_outputs = True` hands out the engine's static output buffers directly (they are overwritten
        by the next forward of the same geometry); the default clones them so that every call returns fresh
        tensors like the reference does.  `self.lean_outputs = True` skips data['scores'] / data['kp_scores'] (the solver
        only reads final_scores): 17 instead of 47 MB of N x N traffic per 720x540 pair."""
        if getattr(self, "staged", False):
            return self.forward_staged(data, return_inliers)
        pool = self._engine_pool()
        turn = self.__dict__.get("_turn", 0)
        self.__dict__["_turn"] = turn + 1
        eng = pool[turn % len(pool)]
        im0, im1 = data["image0"], data["image1"]
        B = im0.shape[0]
        seed = int(torch.randint(1, 2 ** 62, (1,)).item())
        if im0.dtype != torch.uint8:                     # uint8 [B,H,W,3] goes to the ingest kernel as it is (mickey_b200.io)
            im0, im1 = im0.float(), im1.float()
        st = eng.forward(im0, im1, data["K_color0"].float(), data["K_color1"].float(), seed,
                         use_graph=getattr(self, "use_graph", True), lean=bool(getattr(self, "lean_outputs", False)))
        keep = (lambda t: t) if getattr(self, "static_outputs", False) else (lambda t: t.clone())
        H, W = eng.geo
        gh, gw = H // PATCH, W // PATCH
        N = gh * gw
        kps, depth, scr, dsc = keep(st["kps"]), keep(st["depth"]), keep(st["scr"]), keep(st["dsc"])
        data["kps0_shape"], data["kps1_shape"], data["down_factor"] = [gh, gw], [gh, gw], self.compute_matches.down_factor
        data["depth0_map"], data["depth1_map"] = depth[:B].reshape(B, 1, gh, gw), depth[B:].reshape(B, 1, gh, gw)
        data["kps0"], data["kps1"] = kps[:B], kps[B:]
        data["depth_kp0"], data["depth_kp1"] = depth[:B], depth[B:]
        data["scr0"], data["scr1"] = scr[:B], scr[B:]
        data["dsc0"], data["dsc1"] = dsc[:B], dsc[B:]
        if st["scores"] is not None:             # lean_outputs: only final_scores (what the solver reads) is materialised
            data["scores"], data["kp_scores"] = keep(st["scores"]), keep(st["kp_scores"])
        data["final_scores"] = keep(st["final_scores"])
        pose = keep(st["pose"])
        R, t, inliers = pose[:, :9].reshape(B, 3, 3), pose[:, 9:12].reshape(B, 1, 3), pose[:, 12:13]
        if return_inliers:
            data["inliers_list"] = self._inlier_list(data, st, B, N)
        data["R"], data["t"], data["inliers"] = R, t, inliers
        return R, t

    @torch.no_grad()
    def forward_staged(self, data, return_inliers=False):
        """The same path as three C calls (mk_extract / mk_match / mk_solve_pose), mirroring the reference's
        structure (compute_pose.py:20-37); used by the stage-wise tests."""
        self.compute_matches(data)
        # final_scores = scores * kp_scores (compute_pose.py:23) is produced by the matcher kernel's epilogue
        data["final_scores"] = data.pop("_final_scores_fused")
        if return_inliers:
            R, t, inliers, inliers_list = self.e2e_Procrustes.estimate_pose_vectorized(data, return_inliers=True)
            data["inliers_list"] = inliers_list
        else:
            R, t, inliers = self.e2e_Procrustes.estimate_pose_vectorized(data, return_inliers=False)
        data.pop("_solver", None)
        data["R"] = R
        data["t"] = t
        data["inliers"] = inliers
        return R, t


def build_model(cfg, checkpoint=""):
    """Mirror of reference lib/models/builder.py:5-20.  `checkpoint` may be a path (torch.load) or an
    already loaded dict {'state_dict': ...}."""
    if cfg.MODEL == "MicKey":
        model = MickeyRelativePose(cfg)
        ckpt = checkpoint if isinstance(checkpoint, dict) else torch.load(checkpoint, map_location="cpu")
        model.on_load_checkpoint(ckpt)
        model.load_state_dict(ckpt["state_dict"])
        if torch.cuda.is_available():
            model = model.cuda()
        model.eval()
        return model
    raise NotImplementedError()
