"""Launch the hot kernels in isolation on BASELINE-size operands (for `ncu --set full`).
    ncu --set full --clock-control none --import-source on -o gpurun_out/prof python tools/ncu_targets.py [names...]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200 import _lib  # noqa: E402
from tests.gpu_util import gemm, stream  # noqa: E402

lib = _lib.load()
dev = "cuda"
which = set(sys.argv[1:]) or {"fc1", "proj", "attention", "conv", "dual", "sampler", "linattn"}
M, D, T, N = 3878, 384, 1939, 1938
torch.manual_seed(0)
reps = int(os.environ.get('NCU_REPS', '2'))

if "fc1" in which:          # ViT-S fc1: [3878,384] x [1536,384]^T + bias + GELU -> fp16
    a, w, b = torch.randn(M, D, device=dev).half(), (torch.randn(4 * D, D, device=dev) * 0.02).half(), torch.randn(4 * D, device=dev)
    out = torch.empty(M, 4 * D, dtype=torch.float16, device=dev)
    for _ in range(reps):
        gemm("STORE_H", a, w, M, 4 * D, D, bias=b, act=1, out_h=out, out_h_ld=4 * D)
if "proj" in which:         # ViT-S fc2: K=1536 -> N=384, LayerScale + residual fp32 in place
    a, w = torch.randn(M, 4 * D, device=dev).half(), (torch.randn(D, 4 * D, device=dev) * 0.02).half()
    b, g, x = torch.randn(D, device=dev), torch.randn(D, device=dev), torch.randn(M, D, device=dev)
    for _ in range(reps):
        gemm("RESID_F", a, w, M, D, 4 * D, bias=b, gamma=g, out_f=x, out_f_ld=D)
if "attention" in which:
    qkv = torch.randn(2 * T, 3 * D, device=dev).half()
    out = torch.empty(2 * T, D, dtype=torch.float16, device=dev)
    for _ in range(reps):
        _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(out), 2, T, D, 6, 0, stream()))
if "attention_big" in which:   # ViT-B shape, full waves (8 images x 12 heads)
    qkv = torch.randn(8 * T, 3 * 768, device=dev).half()
    out = torch.empty(8 * T, 768, dtype=torch.float16, device=dev)
    for _ in range(reps):
        _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(out), 8, T, 768, 12, 0, stream()))
if "conv" in which:         # heads resblock1 conv2: 4 groups x (512 -> 512, 3x3) over 2 padded 53x40 images
    h2, w2, G, Cc = 53, 40, 4, 512
    R = 2 * h2 * w2
    a = torch.randn(R, G * Cc, device=dev).half()
    w = (torch.randn(G * Cc, 9 * Cc, device=dev) * 0.01).half()
    b = torch.randn(G * Cc, device=dev)
    out = torch.empty(R, G * Cc, dtype=torch.float16, device=dev)
    taps = [(ky - 1) * w2 + (kx - 1) for ky in range(3) for kx in range(3)]
    for _ in range(reps):
        gemm("CONV", a, w, R, Cc, taps=taps, chunks_per_tap=Cc // 64, groups=G, a_col_group_off=Cc, b_row_group_off=Cc,
             bias=b, bias_group_off=Cc, act=2, pad_h2=h2, pad_w2=w2, out_h=out, out_h_ld=G * Cc, out_h_group_off=Cc)
if "dual" in which:
    d0 = torch.nn.functional.normalize(torch.randn(1, N, 128, device=dev), dim=-1)
    d1 = torch.nn.functional.normalize(torch.randn(1, N, 128, device=dev), dim=-1)

    def split(d, role):
        hi = d.half(); lo = (d - hi.float()).half()
        return torch.cat([hi, lo, hi] if role == 0 else [hi, hi, lo], dim=-1).reshape(N, 384).contiguous()
    a0, a1 = split(d0, 0), split(d1, 1)
    dust = torch.ones(1, device=dev)
    NP = (N + 127) // 128 * 128
    pr, pc = torch.zeros(1, NP // 64, NP, 2, device=dev), torch.zeros(1, NP // 32, NP, 2, device=dev)
    lr, lc = torch.zeros(1, NP, device=dev), torch.zeros(1, NP, device=dev)
    s0, s1 = torch.rand(1, N, device=dev), torch.rand(1, N, device=dev)
    sc, kp, fin = (torch.empty(1, N, N, device=dev) for _ in range(3))
    common = dict(groups=1, a_row_group_off=N, b_row_group_off=N, n_valid=N, inv_temp=10.0, part_ld=NP)
    for _ in range(reps):
        gemm("LSE", a0, a1, N, N, 384, part_row=pr, part_col=pc, **common)
        _lib.check(lib.mk_op_matcher_reduce(_lib.ptr(pr), _lib.ptr(pc), _lib.ptr(dust), 1, N, NP, _lib.ptr(lr), _lib.ptr(lc), stream()))
        gemm("DUAL", a0, a1, N, N, 384, lse_r=lr, lse_c=lc, scr0=s0, scr1=s1, scores=sc, kp_scores=kp, final_scores=fin, **common)
if "dual_b" in which:       # the matcher at batch 8 (2048 tiles -> persistent kernels; 360 MB of outputs: HBM, not L2)
    Bm = 8
    d0 = torch.nn.functional.normalize(torch.randn(Bm, N, 128, device=dev), dim=-1)
    d1 = torch.nn.functional.normalize(torch.randn(Bm, N, 128, device=dev), dim=-1)

    def split_b(d, role):
        hi = d.half(); lo = (d - hi.float()).half()
        return torch.cat([hi, lo, hi] if role == 0 else [hi, hi, lo], dim=-1).reshape(Bm * N, 384).contiguous()
    a0, a1 = split_b(d0, 0), split_b(d1, 1)
    dust = torch.ones(1, device=dev)
    NP = (N + 127) // 128 * 128
    pr, pc = torch.zeros(Bm, NP // 64, NP, 2, device=dev), torch.zeros(Bm, NP // 32, NP, 2, device=dev)
    lr, lc = torch.zeros(Bm, NP, device=dev), torch.zeros(Bm, NP, device=dev)
    s0, s1 = torch.rand(Bm, N, device=dev), torch.rand(Bm, N, device=dev)
    PITCH = N if os.environ.get("NCU_DUAL_CONTIGUOUS") == "1" else (N + 31) // 32 * 32
    sc, kp, fin = (torch.empty(Bm, N, PITCH, device=dev)[:, :, :N] for _ in range(3))
    common = dict(groups=Bm, a_row_group_off=N, b_row_group_off=N, n_valid=N, inv_temp=10.0, part_ld=NP, out_pitch=PITCH)
    for _ in range(reps):
        gemm("LSE", a0, a1, N, N, 384, part_row=pr, part_col=pc, **common)
        _lib.check(lib.mk_op_matcher_reduce(_lib.ptr(pr), _lib.ptr(pc), _lib.ptr(dust), Bm, N, NP, _lib.ptr(lr), _lib.ptr(lc), stream()))
        gemm("DUAL", a0, a1, N, N, 384, lse_r=lr, lse_c=lc, scr0=s0, scr1=s1, scores=sc, kp_scores=kp, final_scores=fin, **common)
if "fc1_b" in which or "qkv_b" in which or "fc2_b" in which:    # ViT-B GEMMs of the B=32 workload (half the batch: 32 images)
    Mb, Db = 32 * T, 768
    a = torch.randn(Mb, Db, device=dev).half()
    if "fc1_b" in which:
        w, b = (torch.randn(4 * Db, Db, device=dev) * 0.02).half(), torch.randn(4 * Db, device=dev)
        out = torch.empty(Mb, 4 * Db, dtype=torch.float16, device=dev)
        for _ in range(reps):
            gemm("STORE_H", a, w, Mb, 4 * Db, Db, bias=b, act=1, out_h=out, out_h_ld=4 * Db)
    if "qkv_b" in which:
        w, b = (torch.randn(3 * Db, Db, device=dev) * 0.02).half(), torch.randn(3 * Db, device=dev)
        out = torch.empty(Mb, 3 * Db, dtype=torch.float16, device=dev)
        for _ in range(reps):
            gemm("STORE_H", a, w, Mb, 3 * Db, Db, bias=b, out_h=out, out_h_ld=3 * Db)
    if "fc2_b" in which:
        a4 = torch.randn(Mb, 4 * Db, device=dev).half()
        w, b, g, x = (torch.randn(Db, 4 * Db, device=dev) * 0.02).half(), torch.randn(Db, device=dev), torch.randn(Db, device=dev), torch.randn(Mb, Db, device=dev)
        for _ in range(reps):
            gemm("RESID_F", a4, w, Mb, Db, 4 * Db, bias=b, gamma=g, out_f=x, out_f_ld=Db)
if "conv_b" in which:       # heads resblock1 conv2 at batch: 32 padded 53x40 images
    h2, w2, G, Cc = 53, 40, 4, 512
    R = 32 * h2 * w2
    a = torch.randn(R, G * Cc, device=dev).half()
    w = (torch.randn(G * Cc, 9 * Cc, device=dev) * 0.01).half()
    b = torch.randn(G * Cc, device=dev)
    out = torch.empty(R, G * Cc, dtype=torch.float16, device=dev)
    taps = [(ky - 1) * w2 + (kx - 1) for ky in range(3) for kx in range(3)]
    for _ in range(reps):
        gemm("CONV", a, w, R, Cc, taps=taps, chunks_per_tap=Cc // 64, groups=G, a_col_group_off=Cc, b_row_group_off=Cc,
             bias=b, bias_group_off=Cc, act=2, pad_h2=h2, pad_w2=w2, out_h=out, out_h_ld=G * Cc, out_h_group_off=Cc)
if "attention_b" in which:  # ViT-B attention at batch: 32 images x 12 heads
    qkv = torch.randn(32 * T, 3 * 768, device=dev).half()
    out = torch.empty(32 * T, 768, dtype=torch.float16, device=dev)
    for _ in range(reps):
        _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(out), 32, T, 768, 12, 0, stream()))
if "sampler" in which:
    p = torch.rand(1, N * N, device=dev) * 1e-9
    nb = lib.mk_op_sample_workspace_bytes(1, 8)
    ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
    idx = torch.zeros(8, 2048, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    for _ in range(reps):
        _lib.check(lib.mk_op_sample(_lib.ptr(p), 1, N, 0, 8, 2048, 77, _lib.ptr(ws), nb, _lib.ptr(idx), _lib.ptr(status), stream()))
if "linattn" in which:
    h2, w2, G = 53, 40, 4
    qkv = torch.randn(2 * h2 * w2, G * 384, device=dev)
    kv = torch.zeros(2, G, 8, 272, device=dev)
    kvp = torch.zeros(2, G, (h2 * w2 + 31) // 32, 8, 272, device=dev)
    msg = torch.zeros(2 * h2 * w2, G * 128, dtype=torch.float16, device=dev)
    for _ in range(reps):
        _lib.check(lib.mk_op_linattn(_lib.ptr(qkv), _lib.ptr(kvp), _lib.ptr(kv), _lib.ptr(msg), 2, G, h2, w2, 1e-6, stream()))
torch.cuda.synchronize()
print("done", sorted(which))
