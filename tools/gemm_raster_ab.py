#!/usr/bin/env python
"""Interleaved raster sweep for the CTA-pair GEMM (LR_GEMM_GROUP_M is read at every launch): all group sizes of a
shape are timed round-robin, the order rotating per round, after a few seconds of heating - the sequential sweep of
tools/gemm_raster_bench.py favours whatever runs first under the power cap. python tools/gemm_raster_ab.py"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402

SHAPES = [("phi down", 64220, 3072, 8320, L.EPI_RESIDUAL), ("phi o", 64220, 3072, 3200, L.EPI_RESIDUAL),
          ("phi gate_up", 64220, 16384, 3200, L.EPI_SWIGLU), ("phi qkv", 64220, 9216, 3200, L.EPI_NONE),
          ("clip fc2", 240032, 1024, 4096, L.EPI_RESIDUAL), ("clip fc1", 240032, 4096, 1024, L.EPI_NONE),
          ("clip qkv", 240032, 3072, 1024, L.EPI_NONE), ("clip out", 240032, 1024, 1024, L.EPI_RESIDUAL)]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default="0,2,4,6,8,12,16,20,32")
    ap.add_argument("--rounds", type=int, default=9)
    ap.add_argument("--iters", type=int, default=4)
    a = ap.parse_args()
    groups = a.groups.split(",")
    dev = "cuda"
    for name, M, N, K, epi in SHAPES:
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
        No = N // 2 if epi == L.EPI_SWIGLU else N
        C = torch.empty(M, No, device=dev, dtype=torch.bfloat16)
        R = torch.randn(M, No, device=dev, dtype=torch.bfloat16) if epi == L.EPI_RESIDUAL else None
        os.environ["LR_GEMM_GROUP_M"] = "0"
        for _ in range(300 if name == SHAPES[0][0] else 30):   # heat
            ops.gemm(A, W, C, M, N, K, epi, None, R)
        torch.cuda.synchronize()
        t = {g: [] for g in groups}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(a.rounds):
            order = groups[r % len(groups):] + groups[:r % len(groups)]
            if r % 2:
                order = order[::-1]
            for g in order:
                os.environ["LR_GEMM_GROUP_M"] = g
                ops.gemm(A, W, C, M, N, K, epi, None, R)
                e0.record()
                for _ in range(a.iters):
                    ops.gemm(A, W, C, M, N, K, epi, None, R)
                e1.record()
                torch.cuda.synchronize()
                t[g].append(e0.elapsed_time(e1) / a.iters)
        heur = int((32 << 20) / (256 * K * 2))
        heur = min(max(heur, 4), 32)
        res = "  ".join(f"g{g}: {2.0 * M * N * K / statistics.median(t[g]) / 1e9:.0f}" for g in groups)
        print(f"{name:12s} M={M} N={N} K={K} (g0 = heuristic {heur})  TF/s median of {a.rounds}  {res}", flush=True)
        del A, W, C, R
    os.environ.pop("LR_GEMM_GROUP_M", None)


if __name__ == "__main__":
    main()
