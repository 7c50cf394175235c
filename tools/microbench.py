"""Per-kernel device times (CUDA events, warm, averaged) of the hot kernels on BASELINE-size operands.
    python tools/microbench.py            -> prints one line per kernel: name, us, achieved TFLOP/s or GB/s"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200 import _lib  # noqa: E402
from tests.gpu_util import gemm, stream  # noqa: E402

lib = _lib.load()
dev = "cuda"
M, D, T, N = 3878, 384, 1939, 1938
torch.manual_seed(0)


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def report(name, us, flops=None, nbytes=None):
    extra = f"{flops / us / 1e6:8.1f} TFLOP/s" if flops else (f"{nbytes / us / 1e3:8.1f} GB/s" if nbytes else "")
    print(f"{name:32s} {us:9.2f} us  {extra}", flush=True)


def lin(M_, N_, K_, epi, **kw):
    a = torch.randn(M_, K_, device=dev).half()
    w = (torch.randn(N_, K_, device=dev) * 0.02).half()
    return lambda: gemm(epi, a, w, M_, N_, K_, **kw)


b1536, b384, b1152 = (torch.randn(n, device=dev) for n in (1536, 384, 1152))
g384 = torch.randn(384, device=dev)
x32 = torch.randn(M, D, device=dev)
o1152 = torch.empty(M, 1152, dtype=torch.float16, device=dev)
o1536 = torch.empty(M, 1536, dtype=torch.float16, device=dev)
report("vit.qkv  3878x1152x384", timeit(lin(M, 1152, 384, "STORE_H", bias=b1152, out_h=o1152, out_h_ld=1152)), 2 * M * 1152 * 384)
report("vit.proj 3878x384x384", timeit(lin(M, 384, 384, "RESID_F", bias=b384, gamma=g384, out_f=x32, out_f_ld=384)), 2 * M * 384 * 384)
report("vit.fc1  3878x1536x384 gelu", timeit(lin(M, 1536, 384, "STORE_H", bias=b1536, act=1, out_h=o1536, out_h_ld=1536)), 2 * M * 1536 * 384)
report("vit.fc1  (no act, no bias)", timeit(lin(M, 1536, 384, "STORE_H", out_h=o1536, out_h_ld=1536)), 2 * M * 1536 * 384)
report("vit.fc2  3878x384x1536", timeit(lin(M, 384, 1536, "RESID_F", bias=b384, gamma=g384, out_f=x32, out_f_ld=384)), 2 * M * 384 * 1536)
big = 16384
obig = torch.empty(big, 4096, dtype=torch.float16, device=dev)
report("gemm 16384x4096x4096", timeit(lin(big, 4096, 4096, "STORE_H", out_h=obig, out_h_ld=4096), iters=10), 2 * big * 4096 * 4096)

qkv = torch.randn(2 * T, 3 * D, device=dev).half()
att = torch.empty(2 * T, D, dtype=torch.float16, device=dev)
fl_att = 2 * 6 * 4 * T * T * 64
for impl, nm in ((1, "tcgen05"), (2, "mma.sync")):
    report(f"vit.attention {nm}", timeit(lambda: _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(att), 2, T, D, 6, impl, stream()))), fl_att)

xn = torch.empty(M, D, dtype=torch.float16, device=dev)
w_ln, b_ln = torch.randn(D, device=dev), torch.randn(D, device=dev)
report("vit.layernorm 3878x384", timeit(lambda: _lib.check(lib.mk_op_layernorm(_lib.ptr(x32), _lib.ptr(w_ln), _lib.ptr(b_ln), _lib.ptr(xn), M, D, 1e-6, 0, 0, 0, stream()))),
       nbytes=M * D * 6)

h2, w2, G, Cc = 53, 40, 4, 512
R = 2 * h2 * w2
a = torch.randn(R, G * Cc, device=dev).half()
w = (torch.randn(G * Cc, 9 * Cc, device=dev) * 0.01).half()
bb = torch.randn(G * Cc, device=dev)
out = torch.empty(R, G * Cc, dtype=torch.float16, device=dev)
taps = [(ky - 1) * w2 + (kx - 1) for ky in range(3) for kx in range(3)]
report("head.conv3x3 4x(512->512)", timeit(lambda: gemm("CONV", a, w, R, Cc, taps=taps, chunks_per_tap=Cc // 64, groups=G, a_col_group_off=Cc,
       b_row_group_off=Cc, bias=bb, bias_group_off=Cc, act=2, pad_h2=h2, pad_w2=w2, out_h=out, out_h_ld=G * Cc, out_h_group_off=Cc)),
       2 * R * G * Cc * 9 * Cc)

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
lse = lambda: gemm("LSE", a0, a1, N, N, 384, part_row=pr, part_col=pc, lse_bound=1.001, **common)      # normalised descriptors
lse_max = lambda: gemm("LSE", a0, a1, N, N, 384, part_row=pr, part_col=pc, **common)                   # true row / column maxima
red = lambda: _lib.check(lib.mk_op_matcher_reduce(_lib.ptr(pr), _lib.ptr(pc), _lib.ptr(dust), 1, N, NP, _lib.ptr(lr), _lib.ptr(lc), stream()))
PITCH = (N + 31) // 32 * 32
scp, kpp, finp = (torch.empty(1, N, PITCH, device=dev)[:, :, :N] for _ in range(3))       # 128-byte aligned rows: TMA tensor stores
dual_c = lambda: gemm("DUAL", a0, a1, N, N, 384, lse_r=lr, lse_c=lc, scr0=s0, scr1=s1, scores=sc, kp_scores=kp, final_scores=fin, out_pitch=N, **common)
dual = lambda: gemm("DUAL", a0, a1, N, N, 384, lse_r=lr, lse_c=lc, scr0=s0, scr1=s1, scores=scp, kp_scores=kpp, final_scores=finp, out_pitch=PITCH, **common)
dual_lean = lambda: gemm("DUAL", a0, a1, N, N, 384, lse_r=lr, lse_c=lc, scr0=s0, scr1=s1, final_scores=finp, out_pitch=PITCH, **common)
lse(); red()
report("match.lse (rows + columns, fixed shift)", timeit(lse), 2 * N * N * 384)
report("match.lse (rows + columns, true maxima)", timeit(lse_max), 2 * N * N * 384)
report("match.reduce", timeit(red), nbytes=(NP // 64 + NP // 32) * NP * 8)
report("match.dual_softmax (contiguous, st.global)", timeit(dual_c), nbytes=3 * N * N * 4 + 2 * N * 384 * 2)
report("match.dual_softmax (pitch 1952, TMA stores)", timeit(dual), nbytes=3 * N * N * 4 + 2 * N * 384 * 2)
report("match.dual_softmax (lean)", timeit(dual_lean), nbytes=N * N * 4 + 2 * N * 384 * 2)
report("matcher, all three launches", timeit(lambda: (lse(), red(), dual())), nbytes=3 * N * N * 4 + 2 * N * 128 * 4 + 2 * N * 4)
wr = torch.empty(32 * 3 * N * N, device=dev)
report("(torch fill of 32*3*N*N fp32: write-only stream)", timeit(lambda: wr.fill_(1.0), iters=10), nbytes=32 * 3 * N * N * 4)
del wr
cp_src, cp_dst = torch.empty(3 * N * N, device=dev), torch.empty(3 * N * N, device=dev)
report("(torch copy of 3*N*N fp32)", timeit(lambda: cp_dst.copy_(cp_src)), nbytes=2 * 3 * N * N * 4)

p = torch.rand(1, N * N, device=dev) * 1e-9
nb = lib.mk_op_sample_workspace_bytes(1, 8)
ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
idx = torch.zeros(8, 2048, dtype=torch.int32, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
report("solve.sample_outer (8 streams)", timeit(lambda: _lib.check(lib.mk_op_sample(_lib.ptr(p), 1, N, 0, 8, 2048, 77, _lib.ptr(ws), nb, _lib.ptr(idx), _lib.ptr(status), stream())), iters=20),
       nbytes=N * N * 4)
pp = (torch.rand(1, N, PITCH, device=dev) * 1e-9)
report("solve.sample_outer (8 streams, row pitch 1952)", timeit(lambda: _lib.check(lib.mk_op_sample(_lib.ptr(pp), 1, N, PITCH, 8, 2048, 77, _lib.ptr(ws), nb, _lib.ptr(idx), _lib.ptr(status), stream())), iters=20),
       nbytes=N * N * 4)
