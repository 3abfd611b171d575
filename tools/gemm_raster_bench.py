#!/usr/bin/env python
"""Raster experiment for the CTA-pair GEMM: time the large-K decoder shapes for several m-group sizes
(LR_GEMM_GROUP_M overrides the launch heuristic). python tools/gemm_raster_bench.py

CAVEAT (r02): the group sizes are timed one after the other, so on a power-capped B200 the FIRST entry is favoured by
up to 10 % ("g0" = the heuristic and the explicit value of the same heuristic differ by that much,
gpurun_out/r02_gemm_raster_hints.txt). Use it for > 15 % effects only, or interleave as tools/attn_ab.py does.
--once: one launch per shape for `ncu --metrics dram__bytes_read.sum,...` (tools/gpu_gemm_l2_hints.sh)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402

SHAPES = [("qwen down", 86912, 3584, 19072, L.EPI_RESIDUAL), ("phi down", 65536, 3072, 8320, L.EPI_RESIDUAL),
          ("llava7b down", 195648, 4096, 11136, L.EPI_RESIDUAL), ("phi gate_up", 65536, 16384, 3200, L.EPI_SWIGLU),
          ("qwen gate_up", 86912, 37888, 3840, L.EPI_SWIGLU), ("phi qkv", 65536, 9216, 3200, L.EPI_NONE),
          ("phi o", 65536, 3072, 3200, L.EPI_RESIDUAL), ("qwen vit down", 313600, 1280, 3456, L.EPI_RESIDUAL)]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default=None)
    ap.add_argument("--groups", default="0,2,4,6,8,12,16,24,32,64")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--once", action="store_true", help="one launch per shape, no timing (for ncu --metrics dram__bytes_*)")
    a = ap.parse_args()
    dev = "cuda"
    for name, M, N, K, epi in SHAPES:
        if a.shape and a.shape != name:
            continue
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        W = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * K ** -0.5
        No = N // 2 if epi == L.EPI_SWIGLU else N
        C = torch.empty(M, No, device=dev, dtype=torch.bfloat16)
        R = torch.randn(M, No, device=dev, dtype=torch.bfloat16) if epi == L.EPI_RESIDUAL else None
        res = []
        if a.once:
            ops.gemm(A, W, C, M, N, K, epi, None, R)
            torch.cuda.synchronize()
            continue
        for g in a.groups.split(","):
            os.environ["LR_GEMM_GROUP_M"] = g
            for _ in range(a.warm):
                ops.gemm(A, W, C, M, N, K, epi, None, R)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                ops.gemm(A, W, C, M, N, K, epi, None, R)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            res.append(f"g{g}: {2.0 * M * N * K / ms / 1e9:.0f}")
        print(f"{name:14s} M={M} N={N} K={K}  TF/s  " + "  ".join(res), flush=True)
        del A, W, C, R


if __name__ == "__main__":
    main()
