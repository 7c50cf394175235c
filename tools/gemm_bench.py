"""Device times of the ViT-B GEMMs at the C3 batch (64 images = 124096 tokens) and of the head convolution, for A/B runs of
GEMM build options (MICKEY_GEMM_2SM, MICKEY_GEMM_2SM_STAGES):   python tools/gemm_bench.py [n_images]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200 import _lib  # noqa: E402
from tests.gpu_util import gemm  # noqa: E402

lib = _lib.load()
dev = "cuda"
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M, D = n_img * 1939, 768
tag = f"2sm={os.environ.get('MICKEY_GEMM_2SM', '1')} stages={os.environ.get('MICKEY_GEMM_2SM_STAGES', 'default')}"
torch.manual_seed(0)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def lin(N_, K_, epi, **kw):
    a = torch.randn(M, K_, device=dev).half()
    w = (torch.randn(N_, K_, device=dev) * 0.02).half()
    return lambda: gemm(epi, a, w, M, N_, K_, **kw)


x32 = torch.randn(M, D, device=dev)
g = torch.randn(D, device=dev) * 0.1
total = 0.0
for name, N_, K_, epi, kw in (
        ("qkv", 3 * D, D, "STORE_H", dict(bias=torch.randn(3 * D, device=dev), out_h=torch.empty(M, 3 * D, dtype=torch.float16, device=dev), out_h_ld=3 * D)),
        ("proj", D, D, "RESID_F", dict(bias=torch.randn(D, device=dev), gamma=g, out_f=x32, out_f_ld=D)),
        ("fc1+gelu", 4 * D, D, "STORE_H", dict(bias=torch.randn(4 * D, device=dev), act=1, out_h=torch.empty(M, 4 * D, dtype=torch.float16, device=dev), out_h_ld=4 * D)),
        ("fc2", D, 4 * D, "RESID_F", dict(bias=torch.randn(D, device=dev), gamma=g, out_f=x32, out_f_ld=D))):
    us = timeit(lin(N_, K_, epi, **kw))
    total += us
    print(f"{tag} vit-b.{name:9s} {M}x{N_}x{K_}: {us:9.1f} us  {2 * M * N_ * K_ / us / 1e6:7.1f} TFLOP/s", flush=True)
    del kw
print(f"{tag} vit-b block GEMMs: {total:9.1f} us", flush=True)
big = 16384
obig = torch.empty(big, 4096, dtype=torch.float16, device=dev)
a = torch.randn(big, 4096, device=dev).half(); w = (torch.randn(4096, 4096, device=dev) * 0.02).half()
us = timeit(lambda: gemm("STORE_H", a, w, big, 4096, 4096, out_h=obig, out_h_ld=4096))
print(f"{tag} gemm 16384x4096x4096: {us:9.1f} us  {2 * big * 4096 * 4096 / us / 1e6:7.1f} TFLOP/s", flush=True)
h2, w2, G, Cc = 53, 40, 4, 512
R = n_img * h2 * w2
a = torch.randn(R, G * Cc, device=dev).half()
w = (torch.randn(G * Cc, 9 * Cc, device=dev) * 0.01).half()
bb = torch.randn(G * Cc, device=dev)
out = torch.empty(R, G * Cc, dtype=torch.float16, device=dev)
taps = [(ky - 1) * w2 + (kx - 1) for ky in range(3) for kx in range(3)]
us = timeit(lambda: gemm("CONV", a, w, R, Cc, taps=taps, chunks_per_tap=Cc // 64, groups=G, a_col_group_off=Cc, b_row_group_off=Cc, bias=bb,
                         bias_group_off=Cc, act=2, pad_h2=h2, pad_w2=w2, out_h=out, out_h_ld=G * Cc, out_h_group_off=Cc))
print(f"{tag} head.conv3x3 4x(512->512), {n_img} images: {us:9.1f} us  {2 * R * G * Cc * 9 * Cc / us / 1e6:7.1f} TFLOP/s", flush=True)
