"""SM clock / power while the attention kernel runs back to back (is the micro-benchmark itself power-capped?).
usage (GPU box): python tools/attn_clock.py"""
import os, sys, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L, ops
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
bf = torch.bfloat16
for name, (nseq, T, heads, hd, causal) in {"clip": (416, 577, 16, 64, False), "dec": (32, 2048, 32, 96, True)}.items():
    D = heads * hd
    qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
    o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
    samples, stop = [], False
    def sample():
        while not stop:
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
            time.sleep(0.05)
    th = threading.Thread(target=sample); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 1500
    e0.record()
    for _ in range(n):
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, None, None, heads, hd, causal, hd ** -0.5, L.ATTN_TCGEN05)
    e1.record(); torch.cuda.synchronize(); stop = True; th.join()
    ms = e0.elapsed_time(e1) / n
    tail = samples[len(samples) // 2:]
    print(f"{name}: {ms:.3f} ms/launch over {n} launches; SM MHz median {sorted(c for c, _ in tail)[len(tail)//2]}, "
          f"power W median {sorted(p for _, p in tail)[len(tail)//2]:.0f}; first samples {samples[:3]} last {samples[-3:]}")
