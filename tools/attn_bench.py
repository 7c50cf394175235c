"""Attention kernel timing at the C2 (2 images, ViT-S) and C3 (64 images, ViT-B) shapes; MICKEY_ATTN_POLY selects the
fraction of exp2 evaluated on the FMA pipe (0, 8, 4)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200 import _lib  # noqa: E402
from tests.gpu_util import stream  # noqa: E402
lib = _lib.load()
T = 1939
IMPL = int(os.environ.get('ATTN_IMPL', '1'))
for n_img, heads in ((2, 6), (16, 12), (64, 12)):
    D = heads * 64
    qkv = torch.randn(n_img * T, 3 * D, device="cuda").half()
    att = torch.empty(n_img * T, D, dtype=torch.float16, device="cuda")
    fn = lambda: _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(att), n_img, T, D, heads, IMPL, stream()))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20 if n_img < 64 else 5
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    fl = n_img * heads * 4 * T * T * 64
    print(f"impl={IMPL} poly={os.environ.get('MICKEY_ATTN_POLY','4')} imgs={n_img} heads={heads}: {us:9.1f} us  {fl/us/1e6:7.1f} TFLOP/s  exp-rate {n_img*heads*T*2048/us/1e3/148/1.965:.2f} /clk/SM", flush=True)
