"""Device-side pipeline driver: packs a reference-named state dict into the layouts the CUDA kernels
consume, owns the workspace, and calls the C ABI (include/mickey_b200.h) on the current CUDA stream.

PyTorch is plumbing here (device memory, streams); all arithmetic of the hot path happens inside
libmickey_b200.so.  The only torch arithmetic in this file is weight preparation at load time:
fp16 casts, BatchNorm folding into the conv weights, stacking the four heads, the bicubic resize of
the position embedding (same torch call as the reference, dinov2.py:165-189) and the sine table.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _lib
from .config import VARIANTS, backbone_variant
from .weights import BACKBONE, DUSTBIN, EXTRACTOR

HEAD_ORDER = ("depth_head", "det_offset", "det_head", "dsc_head")     # group order inside the kernels
KPAD = 640
PATCH = 14


def nn_pitch(N: int) -> int:
    """Row pitch (floats) of the N x N outputs the engine allocates: N rounded up to 32, so that every row starts on a
    128-byte line and the matcher's outputs can leave through TMA tensor stores (N = 1938 rows are only 8-byte aligned).
    The tensors handed out are the [.., :N] views; MICKEY_NN_CONTIGUOUS=1 keeps the reference's contiguous layout."""
    import os
    if os.environ.get("MICKEY_NN_CONTIGUOUS") == "1":
        return N
    return (N + 31) // 32 * 32


def nn_empty(B: int, N: int, device):
    """fp32 [B, N, N] view of a [B, N, nn_pitch(N)] buffer."""
    return torch.empty(B, N, nn_pitch(N), device=device)[:, :, :N]


def make_mk_config(cfg) -> _lib.MkConfig:
    variant = backbone_variant(cfg)
    D, depth, heads = VARIANTS[variant]
    m, p = cfg["MICKEY"], cfg["PROCRUSTES"]
    if cfg["FEATURE_MATCHER"]["TYPE"] != "DualSoftmax":
        # the reference's Sinkhorn branch is unreachable (feature_matcher.py:50 vs :125)
        raise NotImplementedError("only FEATURE_MATCHER.TYPE == 'DualSoftmax' is supported")
    c = _lib.MkConfig()
    c.embed_dim, c.depth, c.heads = D, depth, heads
    c.down_factor = int(m["DINOV2"]["DOWN_FACTOR"])
    for i, v in enumerate(m["KP_HEADS"]["BLOCKS_DIM"]):
        c.block_dims[i] = int(v)
    c.desc_dim = int(m["DSC_HEAD"]["LAST_DIM"])
    c.use_softmax = int(bool(m["KP_HEADS"]["USE_SOFTMAX"]))
    c.depth_sigmoid = int(bool(m["KP_HEADS"]["USE_DEPTHSIGMOID"]))
    c.max_depth = float(m["KP_HEADS"]["MAX_DEPTH"])
    c.kp_pos_enc = int(bool(m["KP_HEADS"]["POS_ENCODING"]))
    c.dsc_pos_enc = int(bool(m["DSC_HEAD"]["POS_ENCODING"]))
    c.norm_dsc = int(bool(m["DSC_HEAD"]["NORM_DSC"]))
    c.temperature = float(cfg["FEATURE_MATCHER"]["DUAL_SOFTMAX"]["TEMPERATURE"])
    c.use_dustbin = int(bool(cfg["FEATURE_MATCHER"]["DUAL_SOFTMAX"]["USE_DUSTBIN"]))
    c.it_matches, c.it_ransac = int(p["IT_MATCHES"]), int(p["IT_RANSAC"])
    c.num_sampled, c.num_corr, c.num_refine = int(p["NUM_SAMPLED_MATCHES"]), int(p["NUM_CORR_3D_3D"]), int(p["NUM_REFINEMENTS"])
    c.th_inlier, c.th_soft_inlier = float(p["TH_INLIER"]), float(p["TH_SOFT_INLIER"])
    if c.down_factor != PATCH:
        raise NotImplementedError("DOWN_FACTOR must be 14 (DINOv2 patch size)")
    return c


# ---------------------------------------------------------------------------------------------------------
# weight packing
# ---------------------------------------------------------------------------------------------------------
def _fold_conv3x3(w: torch.Tensor, bn: Optional[Dict[str, torch.Tensor]]):
    """[cout, cin, 3, 3] (+ eval BatchNorm) -> ([cout, 9*cin] with column = (ky*3+kx)*cin + ci, shift[cout])."""
    cout, cin = w.shape[:2]
    w = w.float()
    if bn is not None:
        scale = bn["weight"].float() / torch.sqrt(bn["running_var"].float() + 1e-5)
        shift = bn["bias"].float() - bn["running_mean"].float() * scale
    else:
        scale = torch.ones(cout, device=w.device)
        shift = torch.zeros(cout, device=w.device)
    wp = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    return wp, shift


def pack_weights(sd: Dict[str, torch.Tensor], cfg, device) -> Dict[str, torch.Tensor]:
    """state dict with reference names (fp32 or fp16) -> {packed name: device tensor}."""
    variant = backbone_variant(cfg)
    D, depth, _ = VARIANTS[variant]
    use_bn = bool(cfg["MICKEY"]["KP_HEADS"]["BN"])
    out: Dict[str, torch.Tensor] = {}

    def g(name):
        return sd[name].detach().to(device)

    def h16(t):
        return t.to(torch.float16).contiguous()

    def f32(t):
        return t.to(torch.float32).contiguous()

    b = BACKBONE
    pw = g(b + "patch_embed.proj.weight").float().reshape(D, 3 * PATCH * PATCH)
    out["patch.w"] = h16(F.pad(pw, (0, KPAD - pw.shape[1])))
    for i in range(depth):
        p, q = f"{b}blocks.{i}.", f"blk{i}."
        out[q + "ln1.w"], out[q + "ln1.b"] = f32(g(p + "norm1.weight")), f32(g(p + "norm1.bias"))
        out[q + "ln2.w"], out[q + "ln2.b"] = f32(g(p + "norm2.weight")), f32(g(p + "norm2.bias"))
        out[q + "qkv.w"], out[q + "qkv.b"] = h16(g(p + "attn.qkv.weight")), f32(g(p + "attn.qkv.bias"))
        out[q + "proj.w"], out[q + "proj.b"] = h16(g(p + "attn.proj.weight")), f32(g(p + "attn.proj.bias"))
        out[q + "fc1.w"], out[q + "fc1.b"] = h16(g(p + "mlp.fc1.weight")), f32(g(p + "mlp.fc1.bias"))
        out[q + "fc2.w"], out[q + "fc2.b"] = h16(g(p + "mlp.fc2.weight")), f32(g(p + "mlp.fc2.bias"))
        out[q + "ls1"], out[q + "ls2"] = f32(g(p + "ls1.gamma")), f32(g(p + "ls2.gamma"))
    out["norm.w"], out["norm.b"] = f32(g(b + "norm.weight")), f32(g(b + "norm.bias"))

    def bn_of(prefix):
        if not use_bn:
            return None
        return {k: g(prefix + k) for k in ("weight", "bias", "running_mean", "running_var")}

    def block(head, r):
        rp = f"{EXTRACTOR}{head}.resblock{r}."
        w1, s1 = _fold_conv3x3(g(rp + "conv1.weight"), bn_of(rp + "bn1."))
        w2, s2 = _fold_conv3x3(g(rp + "conv2.weight"), bn_of(rp + "bn2."))
        sc = sd.get(rp + "shortcut.0.weight")
        sc = None if sc is None else sc.detach().to(device).float().reshape(sc.shape[0], sc.shape[1])
        return w1, s1, w2, s2, sc

    for r in (1, 2, 3):
        parts = [block(hd, r) for hd in HEAD_ORDER]
        assert all(p[4] is not None for p in parts), "resblocks 1-3 change width and must have a shortcut conv"
        out[f"rb{r}.c1.w"], out[f"rb{r}.c1.b"] = h16(torch.cat([p[0] for p in parts])), f32(torch.cat([p[1] for p in parts]))
        out[f"rb{r}.c2.w"], out[f"rb{r}.c2.b"] = h16(torch.cat([p[2] for p in parts])), f32(torch.cat([p[3] for p in parts]))
        out[f"rb{r}.sc.w"] = h16(torch.cat([p[4] for p in parts]))
    kparts = [block(hd, 4) for hd in HEAD_ORDER[:3]]
    out["rb4k.c1.w"], out["rb4k.c1.b"] = h16(torch.cat([p[0] for p in kparts])), f32(torch.cat([p[1] for p in kparts]))
    out["rb4k.c2.w"], out["rb4k.c2.b"] = h16(torch.cat([p[2] for p in kparts])), f32(torch.cat([p[3] for p in kparts]))
    out["rb4k.sc.w"] = h16(torch.cat([p[4] for p in kparts]))
    d = block("dsc_head", 4)
    if d[4] is not None:
        raise NotImplementedError("descriptor head with LAST_DIM != 128 (shortcut conv in resblock4) is not supported")
    out["rb4d.c1.w"], out["rb4d.c1.b"], out["rb4d.c2.w"], out["rb4d.c2.b"] = h16(d[0]), f32(d[1]), h16(d[2]), f32(d[3])

    for l in range(3):
        def lw(name):
            return [g(f"{EXTRACTOR}{hd}.att_layer.layers.{l}.{name}").float() for hd in HEAD_ORDER]
        qkv = [torch.cat([q_, k_, v_]) for q_, k_, v_ in zip(lw("q_proj.weight"), lw("k_proj.weight"), lw("v_proj.weight"))]
        out[f"att{l}.qkv.w"] = h16(torch.cat(qkv))
        out[f"att{l}.merge.w"] = h16(torch.cat(lw("merge.weight")))
        out[f"att{l}.mlp0.w"] = h16(torch.cat(lw("mlp.0.weight")))
        out[f"att{l}.mlp2.w"] = h16(torch.cat(lw("mlp.2.weight")))
        out[f"att{l}.n1.w"], out[f"att{l}.n1.b"] = f32(torch.cat(lw("norm1.weight"))), f32(torch.cat(lw("norm1.bias")))
        out[f"att{l}.n2.w"], out[f"att{l}.n2.b"] = f32(torch.cat(lw("norm2.weight"))), f32(torch.cat(lw("norm2.bias")))

    out["out.depth.w"] = f32(g(EXTRACTOR + "depth_head.depth.weight").reshape(-1))
    out["out.xy.w"] = f32(g(EXTRACTOR + "det_offset.xy_offset.weight").reshape(-1))
    out["out.score.w"] = f32(g(EXTRACTOR + "det_head.score.weight").reshape(-1))
    if DUSTBIN in sd:
        out["dustbin"] = f32(g(DUSTBIN).reshape(1))
    return out


def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """Resize the square position-embedding grid to (gh, gw) exactly as the reference does
    (dinov2.py:165-189: bicubic, scale_factor with the +0.1 trick).  Returns [1 + gh*gw, D] fp32."""
    pe = pos_embed.float()
    n = pe.shape[1] - 1
    gs = int(math.sqrt(n))
    dim = pe.shape[-1]
    if gh * gw == n and gh == gw:
        return pe[0]
    grid = pe[:, 1:].reshape(1, gs, gs, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, scale_factor=((gh + 0.1) / gs, (gw + 0.1) / gs), mode="bicubic")
    assert grid.shape[-2:] == (gh, gw)
    return torch.cat([pe[0, :1], grid.permute(0, 2, 3, 1).reshape(gh * gw, dim)], dim=0)


def sine_table_padded(gh: int, gw: int, d_model: int = 128) -> torch.Tensor:
    """2-D sine position encoding (att_layers/transformer.py:25-36, positions start at 1) laid out on
    the zero-padded token grid: [(gh+2)*(gw+2), d_model], zeros on the pad ring."""
    y = torch.arange(1, gh + 1, dtype=torch.float32).view(gh, 1).expand(gh, gw)
    x = torch.arange(1, gw + 1, dtype=torch.float32).view(1, gw).expand(gh, gw)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))
    pe = torch.zeros(gh, gw, d_model)
    pe[..., 0::4] = torch.sin(x[..., None] * div)
    pe[..., 1::4] = torch.cos(x[..., None] * div)
    pe[..., 2::4] = torch.sin(y[..., None] * div)
    pe[..., 3::4] = torch.cos(y[..., None] * div)
    out = torch.zeros(gh + 2, gw + 2, d_model)
    out[1:-1, 1:-1] = pe
    return out.reshape(-1, d_model).contiguous()


# ---------------------------------------------------------------------------------------------------------
class Engine:
    """One C handle + packed weights + workspace for a fixed (cfg, device)."""
    MAX_GEOMETRIES = 4          # image geometries whose tables / workspace / graphs are kept alive at once

    def __init__(self, cfg, device, side_stream: bool = False):
        """side_stream=True gives the engine its own CUDA stream for forward(): two such engines (see
        MickeyRelativePose.pipeline_depth) keep two steps in flight, so the many kernels of one step that cannot fill
        148 SMs at B=1 share the GPU with the next step's."""
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MickeyB200Error("mickey_b200 runs on a CUDA device only (sm_100a); there is no CPU path")
        self.mkcfg = make_mk_config(cfg)
        h = C.c_void_p()
        _lib.check(self.lib.mk_create(self.device.index or 0, C.byref(self.mkcfg), C.byref(h)), "mk_create")
        self.h = h
        self.packed: Dict[str, torch.Tensor] = {}
        self.stream = torch.cuda.Stream(device=self.device) if side_stream else None
        self.assume_inputs_ready = False
        self._raw_pos = None
        self.geo = None
        self.ws = None
        self.ws_pairs = 0
        # Per-geometry state (size-dependent tables, workspace, captured graphs) stays alive while graphs that baked
        # its pointers may still be replayed; a weight (re)load drops all of it.
        self._geo_state: Dict[tuple, dict] = {}
        self._graphs, self._slot = {}, {}
        self._copy_stream = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.mk_destroy(self.h)
        except Exception:
            pass

    # -- weights ---------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], share_with: "Engine" = None):
        """share_with: another engine on the same device whose packed weight tensors are registered here too
        (one copy of the weights in HBM for all pipeline slots)."""
        if share_with is not None:
            self.packed = {k: v for k, v in share_with.packed.items() if k not in ("patch.posb", "patch.clspos", "head.pe")}
            self._raw_pos = share_with._raw_pos
        else:
            with torch.no_grad():
                self.packed = pack_weights(sd, self.cfg, self.device)
                self._raw_pos = (sd[BACKBONE + "pos_embed"].detach().to(self.device).float(),
                                 sd[BACKBONE + "cls_token"].detach().to(self.device).float(),
                                 sd[BACKBONE + "patch_embed.proj.bias"].detach().to(self.device).float())
        for name, t in self.packed.items():
            self._register(name, t)
        # captured graphs hold raw pointers of the previous packed weights and tables: none of them may be replayed
        self._graphs.clear()
        self._slot.clear()
        self._geo_state.clear()
        self.geo = None
        self.ws = None
        self.ws_pairs = 0

    def _register(self, name, t):
        assert t.is_contiguous() and t.device == self.device
        dt = {torch.float32: 0, torch.float16: 1}[t.dtype]
        _lib.check(self.lib.mk_set_tensor(self.h, name.encode(), _lib.ptr(t), dt, t.numel()), f"mk_set_tensor({name})")

    def prepare(self, n_pairs: int, H: int, W: int):
        """Size-dependent tables + workspace for images cropped to (H, W) (multiples of 14)."""
        assert self.packed, "load_state_dict first"
        if self.geo != (H, W):
            if self.geo is not None:                       # park the outgoing geometry's workspace with its state
                self._geo_state[self.geo].update(ws=self.ws, ws_pairs=self.ws_pairs)
            gs = self._geo_state.get((H, W))
            if gs is None:
                gh, gw = H // PATCH, W // PATCH
                with torch.no_grad():
                    pos, cls, pbias = self._raw_pos
                    full = interpolate_pos_embed(pos, gh, gw)
                    gs = {"patch.posb": (full[1:] + pbias[None]).contiguous(),
                          "patch.clspos": (cls.reshape(-1) + full[0]).contiguous(),
                          "head.pe": sine_table_padded(gh, gw).to(self.device), "ws": None, "ws_pairs": 0}
                if len(self._geo_state) >= self.MAX_GEOMETRIES:      # evict the oldest geometry together with its graphs
                    old = next(iter(self._geo_state))
                    del self._geo_state[old]
                    for k in [k for k in self._graphs if (k[1], k[2]) == old]:
                        del self._graphs[k]
                self._geo_state[(H, W)] = gs
            for n in ("patch.posb", "patch.clspos", "head.pe"):
                self.packed[n] = gs[n]
                self._register(n, gs[n])
            _lib.check(self.lib.mk_finalize(self.h, H, W), "mk_finalize")
            self.geo = (H, W)
            self.ws, self.ws_pairs = gs["ws"], gs["ws_pairs"]
        if self.ws is None or self.ws_pairs < n_pairs:
            nbytes = self.lib.mk_workspace_bytes(self.h, n_pairs, H, W)
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.ws_pairs = n_pairs
        return self.ws

    @property
    def launch_count(self) -> int:
        """Kernel launches issued by the library (eager calls) ..."""
        return int(self.lib.mk_launch_count(self.h))

    @property
    def total_kernel_launches(self) -> int:
        """... plus the kernel nodes executed by CUDA-graph replays of mk_forward."""
        return self.launch_count + getattr(self, "graph_launches", 0)

    def ws_view(self, name: str, dtype, shape):
        """Typed view of a named intermediate buffer of the last call's workspace (debugging / tests)."""
        H, W = self.geo
        off = self.lib.mk_workspace_offset(self.h, name.encode(), self.ws_pairs, H, W)
        if off < 0:
            raise _lib.MickeyB200Error(self.lib.mk_last_error().decode())
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.ws[off:off + nbytes].view(dtype).reshape(shape)

    def profile(self, enable: bool):
        _lib.check(self.lib.mk_profile_enable(self.h, int(enable)), "mk_profile_enable")

    def profile_read(self) -> Dict[str, tuple]:
        """{kernel class: (launch scopes, total device ms)} measured with CUDA events on the launch stream."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check(self.lib.mk_profile_read(self.h, buf, len(buf)), "mk_profile_read")
        out = {}
        for line in buf.value.decode().splitlines():
            tag, n, ms = line.split()
            out[tag] = (int(n), float(ms))
        return out

    def _ws_for(self, n_pairs, H, W):
        # the workspace layout depends on n_pairs: carve exactly for this call's batch
        self.prepare(n_pairs, H, W)
        if self.ws_pairs != n_pairs:
            nbytes = self.lib.mk_workspace_bytes(self.h, n_pairs, H, W)
            if self.ws.numel() < nbytes:
                self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.ws_pairs = n_pairs
        return self.ws

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # -- stages ------------------------------------------------------------------------------------------------
    def extract(self, images: torch.Tensor):
        """images fp32 [2B, 3, H, W] (image0 batch then image1 batch) -> kps, depth, scr, dsc.
        H, W need not be multiples of 14: the patch gather reads only the top-left 14*(H//14) x 14*(W//14) crop
        (reference mickey_extractor.py:46), so no cropped copy is made."""
        u8 = images.dtype == torch.uint8
        if u8:                                           # [2B, H, W, 3] RGB as decoded (mk_extract_u8)
            assert images.dim() == 4 and images.shape[-1] == 3, "uint8 images must be [n, H, W, 3] (HWC, RGB)"
            images = images.contiguous()
            n_img, H, W, _ = images.shape
        else:
            images = images.float().contiguous()
            n_img, _, H, W = images.shape
        assert n_img % 2 == 0
        B, N = n_img // 2, (H // PATCH) * (W // PATCH)
        ws = self._ws_for(B, H, W)
        dev = self.device
        kps = torch.empty(n_img, 2, N, device=dev)
        depth = torch.empty(n_img, 1, N, device=dev)
        scr = torch.empty(n_img, 1, N, device=dev)
        dsc = torch.empty(n_img, self.mkcfg.desc_dim, N, device=dev)
        fn = self.lib.mk_extract_u8 if u8 else self.lib.mk_extract
        _lib.check(fn(self.h, _lib.ptr(images), B, H, W, _lib.ptr(kps), _lib.ptr(depth), _lib.ptr(scr),
                      _lib.ptr(dsc), _lib.ptr(ws), ws.numel(), self._stream()), "mk_extract")
        return kps, depth, scr, dsc

    def match(self, B: int, N: int, lean: bool = False):
        """lean: only final_scores (the matrix the solver reads) is materialised; scores / kp_scores come back as None."""
        dev = self.device
        scores = None if lean else nn_empty(B, N, dev)
        kp_scores = None if lean else nn_empty(B, N, dev)
        final = nn_empty(B, N, dev)
        _lib.check(self.lib.mk_match(self.h, B, _lib.ptr(scores), _lib.ptr(kp_scores), _lib.ptr(final), final.stride(1),
                                     _lib.ptr(self.ws), self.ws.numel(), self._stream()), "mk_match")
        return scores, kp_scores, final

    # -- whole path in one C call, optionally replayed from a CUDA graph -------------------------------------------
    def _static_buffers(self, B, H, W, u8=False, lean=False):
        dev, c = self.device, self.mkcfg
        N = (H // PATCH) * (W // PATCH)
        f = lambda *s: torch.empty(*s, device=dev)                       # noqa: E731
        return {
            "images": torch.empty(2 * B, H, W, 3, dtype=torch.uint8, device=dev) if u8 else f(2 * B, 3, H, W),
            "K0": f(B, 3, 3), "K1": f(B, 3, 3),
            "kps": f(2 * B, 2, N), "depth": f(2 * B, 1, N), "scr": f(2 * B, 1, N), "dsc": f(2 * B, c.desc_dim, N),
            "scores": None if lean else nn_empty(B, N, dev), "kp_scores": None if lean else nn_empty(B, N, dev),
            "final_scores": nn_empty(B, N, dev), "pose": f(B, 13),
            "best_set": torch.empty(B, dtype=torch.int32, device=dev), "inlier_mask": f(B, c.num_sampled),
            "sampled_idx": torch.empty(B * c.it_matches, c.num_sampled, dtype=torch.int32, device=dev),
            "status": torch.zeros(1, dtype=torch.int32, device=dev),
        }

    def _call_forward(self, st, B, H, W, seed):
        ws = self.ws
        fn = self.lib.mk_forward_u8 if st["images"].dtype == torch.uint8 else self.lib.mk_forward
        _lib.check(fn(
            self.h, _lib.ptr(st["images"]), _lib.ptr(st["K0"]), _lib.ptr(st["K1"]), B, H, W, C.c_ulonglong(seed),
            _lib.ptr(st["kps"]), _lib.ptr(st["depth"]), _lib.ptr(st["scr"]), _lib.ptr(st["dsc"]), _lib.ptr(st["scores"]),
            _lib.ptr(st["kp_scores"]), _lib.ptr(st["final_scores"]), st["final_scores"].stride(1), _lib.ptr(st["pose"]), _lib.ptr(st["best_set"]),
            _lib.ptr(st["inlier_mask"]), _lib.ptr(st["sampled_idx"]), _lib.ptr(st["status"]), _lib.ptr(ws), ws.numel(),
            self._stream()), "mk_forward")

    def forward(self, image0, image1, K0, K1, seed: int, use_graph: bool = True, lean: bool = False):
        """Whole hot path (extract -> match -> solve) for a batch of pairs.

        Returns the dict of STATIC output tensors of this (B, H, W) geometry.  Two buffer sets alternate, so the
        tensors of one call stay valid until the call after the next one with the same geometry.  Host (pinned)
        inputs are copied H2D on a side stream into the other buffer set while the previous call is still
        computing; device inputs are copied D2D on the main stream."""
        B = image0.shape[0]
        u8 = image0.dtype == torch.uint8                 # [B, H, W, 3] RGB straight from the decoder (mk_forward_u8)
        if u8:
            if image0.dim() != 4 or image0.shape[-1] != 3 or image1.dtype != torch.uint8:
                raise _lib.MickeyB200Error("uint8 images must both be [B, H, W, 3] (HWC, RGB)")
            H, W = image0.shape[1], image0.shape[2]
        else:
            H, W = image0.shape[-2], image0.shape[-1]
        if image1.shape != image0.shape:
            raise _lib.MickeyB200Error(f"image0 {tuple(image0.shape)} and image1 {tuple(image1.shape)} must have the same shape: "
                                       "both images of a batch go through one extraction call")
        self._ws_for(B, H, W)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        fmt = (bool(u8), bool(lean))
        slot = self._slot.get((B, H, W, fmt), 0)
        self._slot[(B, H, W, fmt)] = slot ^ 1
        key = (B, H, W, slot, fmt)
        ent = self._graphs.get(key)
        if ent is None or ent["ws_ptr"] != self.ws.data_ptr():
            ent = {"st": self._static_buffers(B, H, W, u8, lean), "graph": None, "launches": 0, "ws_ptr": self.ws.data_ptr(),
                   "calls": 0, "done": None}
            self._graphs[key] = ent
        st = ent["st"]
        caller = torch.cuda.current_stream()
        main = self.stream if self.stream is not None else caller
        if self.stream is not None and image0.device.type != "cpu" and not self.assume_inputs_ready:
            main.wait_stream(caller)                    # device inputs may still be being produced on the caller's stream
        with torch.cuda.stream(main):
            self._forward_on(main, ent, st, image0, image1, K0, K1, B, H, W, seed, use_graph)
        if self.stream is not None:
            caller.wait_event(ent["done"])              # outputs are safe to consume on the caller's stream
        return st

    def _forward_on(self, main, ent, st, image0, image1, K0, K1, B, H, W, seed, use_graph):
        if image0.device.type == "cpu":
            cs = self._copy_stream
            if ent["done"] is not None:
                cs.wait_event(ent["done"])              # the graph that last read this input buffer has finished
            with torch.cuda.stream(cs):
                st["images"][:B].copy_(image0, non_blocking=True)
                st["images"][B:].copy_(image1, non_blocking=True)
                st["K0"].copy_(K0, non_blocking=True)
                st["K1"].copy_(K1, non_blocking=True)
            main.wait_stream(cs)
        else:
            st["images"][:B].copy_(image0, non_blocking=True)
            st["images"][B:].copy_(image1, non_blocking=True)
            st["K0"].copy_(K0, non_blocking=True)
            st["K1"].copy_(K1, non_blocking=True)
        seed = (int(seed) & (2 ** 64 - 1)) or 1
        if not use_graph:
            self._call_forward(st, B, H, W, seed)
        elif ent["graph"] is None and ent["calls"] == 0:
            # first call of this buffer set runs eagerly (one-time lazy initialisation inside the library:
            # function attributes, TMA descriptors); the second call is captured, later calls replay
            l0 = self.launch_count
            self._call_forward(st, B, H, W, seed)
            ent["launches"] = self.launch_count - l0
            ent["calls"] = 1
        else:
            _lib.check(self.lib.mk_set_seed(self.h, C.c_ulonglong(seed), self._stream()), "mk_set_seed")
            if ent["graph"] is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._call_forward(st, B, H, W, 0)      # seed 0 = continue the device-side sequence
                ent["graph"] = g
            ent["graph"].replay()
            self.graph_replays = getattr(self, "graph_replays", 0) + 1
            self.graph_launches = getattr(self, "graph_launches", 0) + ent["launches"]
        if ent["done"] is None:
            ent["done"] = torch.cuda.Event()
        ent["done"].record(main)

    def solve(self, final_scores, kps, depth, K0, K1, seed: int, outer_idx=None, inner_idx=None, want_extras=False):
        """kps [2B,2,N], depth [2B,1,N] as produced by extract (image0 rows first).  final_scores [B,N,N] may be a padded
        view (last dim contiguous, rows `stride(1)` floats apart) or any tensor (made contiguous)."""
        B, N, _ = final_scores.shape
        if final_scores.stride(2) != 1 or final_scores.stride(0) != N * final_scores.stride(1):
            final_scores = final_scores.contiguous()
        dev = self.device
        c = self.mkcfg
        pose = torch.empty(B, 13, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        best_set = torch.empty(B, dtype=torch.int32, device=dev) if want_extras else None
        mask = torch.empty(B, c.num_sampled, device=dev) if want_extras else None
        sampled = torch.empty(B * c.it_matches, c.num_sampled, dtype=torch.int32, device=dev) if want_extras else None
        hyp = torch.empty(B, c.it_matches * c.it_ransac, device=dev) if want_extras else None
        if outer_idx is not None:
            outer_idx = outer_idx.to(dev, torch.int32).contiguous()
        if inner_idx is not None:
            inner_idx = inner_idx.to(dev, torch.int32).contiguous()
        K0 = K0.to(dev, torch.float32).contiguous()
        K1 = K1.to(dev, torch.float32).contiguous()
        _lib.check(self.lib.mk_solve_pose(
            self.h, _lib.ptr(final_scores), final_scores.stride(1), _lib.ptr(kps), _lib.ptr(depth), _lib.ptr(K0), _lib.ptr(K1), B, N,
            C.c_ulonglong(seed & (2 ** 64 - 1)), _lib.ptr(outer_idx), _lib.ptr(inner_idx), _lib.ptr(pose),
            _lib.ptr(best_set), _lib.ptr(mask), _lib.ptr(sampled), _lib.ptr(hyp), _lib.ptr(status),
            _lib.ptr(self.ws), self.ws.numel(), self._stream()), "mk_solve_pose")
        return {"pose": pose, "status": status, "best_set": best_set, "inlier_mask": mask, "sampled_idx": sampled,
                "hyp_scores": hyp}
