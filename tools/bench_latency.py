"""Single-pair scoring latency through the public API (the call eval/simple_inference.py makes: two custom_forward
calls of batch 1 + preference_compute), host inputs in pinned memory, H2D and the D2H of the probability inside the
timed region. Shapes: BASELINE.json configs[0] ((1344,1344) -> 17 crops, BT head, no SkipCA / LoRA) and the
configs[1] sample shape ((1008,1344), S=2048, SkipCA + LoRA + GPM), plus small batches of the latter.
usage (GPU box): python tools/bench_latency.py [--iters 20]"""
import argparse
import json
import os
import statistics
import sys
import time
import types

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from llava_reward_b200.synth import synth_batch  # noqa: E402


def build(gpm: bool):
    ypath = f"/tmp/llava_reward_b200_lat_{int(gpm)}.yaml"
    with open(ypath, "w") as f:
        yaml.safe_dump({"is_general_preference": gpm, "add_cross_attention": gpm, "value_head_dim": 2 if gpm else 1,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:1234", pm_path=None, cache_dir=None, ft_projector=False,
                                 config_overrides={} if gpm else {"use_lora": False})
    args, model = load_reward_adaptor(args, "phi3v", ypath)
    return args, model.to("cuda").eval()


def measure(args, model, B, hw, seq_len, iters):
    cfg = model.config
    host = {}
    for tag in ("c", "r"):
        ids, mask, pix, sizes = synth_batch(cfg, B, hw, seq_len, seed=7, tag=tag, device="cuda")
        host[tag] = tuple(t.cpu().pin_memory() for t in (ids, mask, pix, sizes))

    def once():
        rs = {}
        for tag in ("c", "r"):
            ids, mask, pix, sizes = host[tag]
            rs[tag], _ = model.custom_forward(ids.to("cuda", non_blocking=True), mask.to("cuda", non_blocking=True),
                                              pix.to("cuda", non_blocking=True), sizes)
        return preference_compute(args, rs["c"], rs["r"])  # synchronises (D2H), like the reference

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    ts = []
    n0 = L.launch_count()
    for _ in range(iters):
        t0 = time.perf_counter()
        once()
        ts.append((time.perf_counter() - t0) * 1e3)
    launches = (L.launch_count() - n0) // iters
    S = host["c"][0].shape[1]
    # where the time of ONE forward goes: host time until custom_forward returns (it blocks once, on the token-plan
    # D2H at its top, then only enqueues) vs the device time between two events around it
    dev_in = tuple(t.to("cuda") for t in host["c"][:3]) + (host["c"][3],)
    host_ms, gpu_ms = [], []
    for _ in range(max(5, iters // 2)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        model.custom_forward(*dev_in)
        host_ms.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        gpu_ms.append(e0.elapsed_time(e1))
    return {"batch_pairs": B, "image_hw": list(hw), "S": S, "ms_per_call_median": statistics.median(ts),
            "ms_min": min(ts), "ms_max": max(ts), "pairs_per_s": B * 1e3 / statistics.median(ts),
            "launches_per_call": launches, "forward_host_enqueue_ms": statistics.median(host_ms),
            "forward_device_ms": statistics.median(gpu_ms)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    out = []
    args, model = build(gpm=False)
    r = measure(args, model, 1, (1344, 1344), None, a.iters)
    r["workload"] = "configs[0] shape: BT head, no SkipCA/LoRA, (1344,1344) -> 17 crops, 1 pair per call"
    out.append(r)
    print(json.dumps(r), flush=True)
    del model
    torch.cuda.empty_cache()
    args, model = build(gpm=True)
    for B in (1, 2, 4, 8):
        r = measure(args, model, B, (1008, 1344), 2048, a.iters if B <= 2 else max(5, a.iters // 2))
        r["workload"] = f"configs[1] sample shape: SkipCA + LoRA r128 + GPM, (1008,1344), S=2048, {B} pair(s) per call"
        out.append(r)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
