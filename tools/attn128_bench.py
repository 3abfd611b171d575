"""head_dim-128 causal attention at the LLaVA-v1.6 bench shape: one-tile vs two-tile tcgen05 variants vs mma.sync."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L, ops  # noqa: E402

bf = torch.bfloat16
nseq, T, heads, hd = 64, 3057, 32, 128
D = heads * hd
qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
lens = torch.randint(2970, 3058, (nseq,), dtype=torch.int32)
ss = (T - lens).to("cuda")
sl = lens.to("cuda")
flops = sum(2.0 * int(n) * int(n) * D for n in lens)  # causal: half of 4 n^2 d per head
outs = {}
for name, impl in (("1tile", L.ATTN_TCGEN05_1TILE), ("2tile", L.ATTN_TCGEN05_2TILE), ("mma.sync", L.ATTN_MMA_SYNC)):
    o = torch.zeros(nseq * T, D, device="cuda", dtype=bf)
    for _ in range(2):
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, ss, sl, heads, hd, True, hd ** -0.5, impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, ss, sl, heads, hd, True, hd ** -0.5, impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    outs[name] = o
    print(f"{name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TF/s", flush=True)
for name in ("2tile", "mma.sync"):
    d = (outs[name].float() - outs["1tile"].float()).abs().max().item()
    print(f"max |{name} - 1tile| = {d:.4g}")
