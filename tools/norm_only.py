import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200 import ops
bf = torch.bfloat16
x = torch.randn(65536, 3072, device="cuda", dtype=bf); w = torch.ones(3072, device="cuda", dtype=bf); y = torch.empty(65536, 3200, device="cuda", dtype=bf)
ops.rmsnorm(x, w, y, 65536, 3072, 1e-5)
xc = torch.randn(240032, 1024, device="cuda", dtype=bf); wc = torch.ones(1024, device="cuda", dtype=bf); yc = torch.empty_like(xc)
ops.layernorm(xc, wc, wc, yc, 240032, 1024, 1e-5)
qkv = torch.randn(65536, 9216, device="cuda", dtype=bf); pos = torch.zeros(65536, device="cuda", dtype=torch.int32)
tab = torch.ones(2048, 48, device="cuda", dtype=bf)
ops.rope_su(qkv, pos, tab, tab, 65536, 32, 96)
torch.cuda.synchronize()
