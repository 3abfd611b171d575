"""Time the SkipCA head kernels alone at the config-2 shape (32 samples, N_v = 1921, H = 3072: 378 MB of K|V rows per
launch, larger than L2) with CUDA events: achieved HBM GB/s = algorithmic bytes (the V half for lr_skipca_head, the K
half for lr_skipca_scores) / time. usage (GPU box): python tools/head_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402


def main():
    B, nv, H, vhd = 32, 1921, 3072, 2
    dev = "cuda"
    torch.manual_seed(0)
    kv = torch.randn(B * nv, 2 * H, device=dev, dtype=torch.bfloat16)
    q = torch.randn(B, H, device=dev, dtype=torch.bfloat16)
    x = torch.randn(B, H, device=dev, dtype=torch.bfloat16)
    ln = torch.ones(H, device=dev, dtype=torch.bfloat16)
    vh = torch.randn(vhd, H, device=dev, dtype=torch.bfloat16) * 0.02
    plan = torch.zeros(B, L.PLAN_STRIDE, dtype=torch.int32)
    plan[:, L.PLAN_ROW_BASE] = torch.arange(B, dtype=torch.int32) * nv
    plan[:, L.PLAN_NV] = nv
    plan = plan.to(dev).view(-1)
    scores = torch.empty(B, nv, device=dev, dtype=torch.float32)
    reward = torch.empty(B, vhd, device=dev, dtype=torch.bfloat16)
    out = {}
    for name, fn, nbytes in (("lr_skipca_scores", lambda: ops.skipca_scores(q, kv, plan, scores, B, H, nv), B * nv * H * 2),
                             ("lr_skipca_head", lambda: ops.skipca_head(scores, kv, plan, x, ln, vh, reward, B, H, nv, vhd,
                                                                        1e-5), B * nv * H * 2)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        out[name] = {"us": best * 1e3, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / best / 1e6}
    peak = 6555.8
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = json.load(f).get("hbm_gbs", peak)
    except OSError:
        pass
    for k in out:
        out[k]["frac_of_measured_copy_peak"] = out[k]["GBps"] / peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()
