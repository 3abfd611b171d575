"""Smallest multi-tile attention launches (for a quick compute-sanitizer pass): a causal pair CTA and a 5-tile CLIP CTA."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L, ops
bf = torch.bfloat16
for (T, heads, hd, causal) in ((300, 1, 96, True), (577, 1, 64, False)):
    D = heads * hd
    qkv = torch.randn(T, 3 * D, device="cuda", dtype=bf)
    outs = []
    for impl in (L.ATTN_TCGEN05_1TILE, L.ATTN_TCGEN05_MULTITILE):
        o = torch.zeros(T, D, device="cuda", dtype=bf)
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, 1, T, None, None, heads, hd, causal, hd ** -0.5, impl)
        torch.cuda.synchronize()
        outs.append(o)
    print("T", T, "identical", torch.equal(outs[0], outs[1]), flush=True)
