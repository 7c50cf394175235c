"""Attention kernel timing at the C2 (2 images) and C3 (64 images, ViT-B) shapes.  MICKEY_ATTN_DBG selects
timing-only variants (results are wrong on purpose)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mickey_b200 import _lib  # noqa: E402
from tests.gpu_util import stream  # noqa: E402
lib = _lib.load()
T = 1939
for n_img, heads in ((2, 6), (16, 12), (64, 12)):
    D = heads * 64
    qkv = torch.randn(n_img * T, 3 * D, device="cuda").half()
    att = torch.empty(n_img * T, D, dtype=torch.float16, device="cuda")
    fn = lambda: _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(att), n_img, T, D, heads, 1, stream()))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20 if n_img < 64 else 5
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    fl = n_img * heads * 4 * T * T * 64
    print(f"dbg={os.environ.get('MICKEY_ATTN_DBG','0')} imgs={n_img} heads={heads}: {us:9.1f} us  {fl/us/1e6:7.1f} TFLOP/s  exp-rate {n_img*heads*T*2048/us/1e3/148/1.965:.2f} /clk/SM", flush=True)
if os.environ.get("MICKEY_ATTN_DBG") == "2":
    import numpy as np
    n_img, heads = 2, 6
    D = heads * 64
    qkv = torch.randn(n_img * T, 3 * D, device="cuda").half()
    att = torch.zeros(n_img * T, D, dtype=torch.float16, device="cuda")
    _lib.check(lib.mk_op_attention(_lib.ptr(qkv), _lib.ptr(att), n_img, T, D, heads, 1, stream()))
    torch.cuda.synchronize()
    raw = att.view(torch.int32).flatten()[: 192 * 160].cpu().numpy().astype(np.int64).reshape(192, 160)
    g0 = raw[:, 150].min()
    for cta in (0, 5, 100, 150, 191):
        r = raw[cta]
        print(f"cta {cta} smid {r[158]} flag {r[159]}: entry->loop {(r[157]-r[152]) & 0xffffffff} clk, loop start->tile0 S ready {(r[1]-r[157]) & 0xffffffff}, "
              f"tiles {(r[15*8+7]-r[0]) & 0xffffffff}, last pv wait {(r[154]-r[15*8+7]) & 0xffffffff}; globaltimer start +{r[150]-g0} ns, end +{r[151]-g0} ns")
    print("all CTA start ns:", sorted((raw[:, 150] - g0).tolist())[::8])
    print("all CTA end ns:", sorted((raw[:, 151] - g0).tolist())[::8])
    st = raw.reshape(192, 20, 8)[:, :16]
    names = ["wait_s", "ldtm", "max", "exp", "wait_pv", "sttm_issue", "st_wait+arrive", "loop"]
    for cta in (0, 5, 100, 150, 191):
        d = st[cta]
        seg = np.concatenate([(d[:, 1:] - d[:, :-1]) & 0xffffffff, ((np.roll(d[:, 0], -1) - d[:, 7]) & 0xffffffff)[:, None]], axis=1)
        print("cta", cta, "tile time", ((d[1:, 0] - d[:-1, 0]) & 0xffffffff)[2:12].mean())
        print("   " + "  ".join(f"{n}={v:.0f}" for n, v in zip(names, seg[2:12].mean(axis=0))))
