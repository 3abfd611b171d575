"""Time the attention implementations of the product library on the product shapes. usage: python tools/attn_impl_bench.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L, ops
bf = torch.bfloat16
shapes = {"clip hd64": (416, 577, 16, 64, False), "decoder hd96 causal": (32, 2048, 32, 96, True),
          "qwen vit hd96 full (4900 tokens)": (8, 4900, 16, 96, False)}
for name, (nseq, T, heads, hd, causal) in shapes.items():
    D = heads * hd
    qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
    o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
    fl = 4.0 * nseq * heads * T * T * hd * (0.5 if causal else 1.0)
    for iname, impl in (("one tile per CTA", L.ATTN_TCGEN05_1TILE), ("multi-tile", L.ATTN_TCGEN05_MULTITILE)):
        run = lambda: ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, None, None, heads, hd, causal, hd ** -0.5, impl)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"{name} | {iname}: {best:.3f} ms = {fl / best / 1e9:.0f} TF/s", flush=True)
