import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L, ops
bf = torch.bfloat16
for name, (nseq, T, heads, hd, causal) in {"clip": (416, 577, 16, 64, False), "dec": (32, 2048, 32, 96, True)}.items():
    D = heads * hd
    qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
    o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
    for _ in range(2):
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, None, None, heads, hd, causal, hd ** -0.5, L.ATTN_TCGEN05)
    torch.cuda.synchronize()
