#!/usr/bin/env python
"""Benchmark of the Qwen2.5-VL branch (BASELINE.json configs[3]: "Qwen2.5-VL-7B LLaVA-Reward backbone (window-attn ViT,
M-RoPE) BT scoring bf16"). Same JSON-line schema as bench.py; bench.py itself stays on configs[1].

    python tools/bench_qwen.py [--steps K] [--warmup W] [--pairs P]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_qwen.py --gpus N

One step = P pairs = 2P samples through custom_forward(inputs_batch): every image is a 70 x 70 patch grid (980 x 980
px, what smart_resize makes of a 1024 x 1024 image under the reference's max_pixels = 1280*28*28,
llava_reward/utils/utils.py:35-37) = 4900 patches -> 1225 image tokens, text U[40,127] tokens, S = 1358 left-padded,
BT head, LoRA r128 on all seven decoder linears; rewards and pair probabilities are read back. Weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PAIRS = 32
GRID = (70, 70)
TEXT_LEN_RANGE = (40, 128)


def tflop_per_sample(cfg, grid, S: int) -> float:
    """Algorithmic FLOPs (SURVEY.md 8(d) conventions: 2MNK, causal attention at half cost, LoRA unmerged, un-padded
    head_dim 80 / intermediate 3420, window attention counted per window, the reference's redundant second vision pass
    and discarded lm_head NOT credited)."""
    from llava_reward_b200.config import qwen_window_plan
    import numpy as np
    h, w = grid
    T = h * w
    D, DI, Lv = cfg.vit_hidden, cfg.vit_intermediate, cfg.vit_depth
    H, I, Lyr, r = cfg.hidden_size, cfg.intermediate_size, cfg.num_layers, (cfg.lora_rank if cfg.use_lora else 0)
    kvw = cfg.num_kv_heads * cfg.head_dim
    plan = qwen_window_plan([(1, h, w)], cfg.vit_merge, cfg.vit_window, cfg.vit_patch)
    win = np.diff(plan["win_cu"]).astype(np.float64)
    n_full = len([i for i in cfg.vit_fullatt if i < Lv])
    att = (Lv - n_full) * 4 * float((win ** 2).sum()) * D + n_full * 4 * float(T) * T * D
    vit = T * (2 * cfg.patch_dim * D + Lv * 2 * (4 * D * D + 3 * D * DI)) + att
    unit = cfg.vit_merge ** 2
    merger = (T // unit) * 2 * (D * unit * D * unit + D * unit * H)
    lin = Lyr * 2 * (H * (H + 2 * kvw) + H * H + 3 * H * I) * S
    lora = Lyr * 2 * r * ((H + H) + 2 * (H + kvw) + (H + H) + 2 * (H + I) + (I + H)) * S
    attn = Lyr * 2 * S * S * H
    return (vit + merger + lin + lora + attn + 2 * S * H * cfg.vhd) / 1e12

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=PAIRS)
    ap.add_argument("--profile-run", action="store_true")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import yaml

    sys.path.insert(0, ROOT)
    from bench import ClockSampler, load_peaks
    from llava_reward_b200 import _lib as L
    from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor
    from llava_reward_b200.synth import synth_batch_qwen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    ypath = f"/tmp/llava_reward_b200_bench_qwen_{rank}.yaml"
    with open(ypath, "w") as f:
        yaml.safe_dump({"is_general_preference": False, "add_cross_attention": False, "value_head_dim": 1,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:1234", pm_path=None,
                                 cache_dir=None, ft_projector=False, config_overrides={})
    args, model = load_reward_adaptor(args, "qwen", ypath)
    model = model.to(dev).eval()
    eng, cfg = model.engine, model.config
    P = a.pairs
    B = 2 * P
    n_img_tok = (GRID[0] // cfg.vit_merge) * (GRID[1] // cfg.vit_merge)
    S = 4 + n_img_tok + 1 + (TEXT_LEN_RANGE[1] - 1) + 1
    batch = synth_batch_qwen(cfg, [GRID] * B, S, seed=7 + rank, tag="pair", device=dev, text_len_range=TEXT_LEN_RANGE)
    keys = ("input_ids", "attention_mask", "pixel_values", "image_grid_thw")
    host = {k: batch[k].cpu().pin_memory() for k in keys}
    resident = {k: host[k].to(dev) for k in keys}
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    out_w = 2 * cfg.vhd + 1
    gather = torch.empty(world * P, out_w, dtype=torch.float32, device=dev) if world > 1 else None

    def step(from_host: bool):
        ib = {k: host[k].to(dev, non_blocking=True) for k in keys} if from_host else resident
        r, _ = model.custom_forward(inputs_batch=ib)
        rc, rr = r[0::2].contiguous(), r[1::2].contiguous()    # samples 2i / 2i+1 = chosen / rejected of pair i
        prob = eng.preference(rc, rr)
        res = torch.cat([rc.float(), rr.float(), prob[:, None]], dim=1)
        if world > 1:
            dist.all_gather_into_tensor(gather, res)
            res = gather
        return res.cpu() if from_host else res

    def timed(from_host: bool, steps: int):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(from_host)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return ms.item()

    if a.profile_run:
        step(False)
        torch.cuda.synchronize()
        print("PROFILE-RUN ms", timed(False, 1), flush=True)
        return
    for _ in range(max(a.warmup, 3)):
        step(False)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    L.reset_launch_count()
    eng.profile = {"gate_up": []}
    ms_dev = timed(False, a.steps)
    launches = L.launch_count()
    prof, eng.profile = eng.profile, None
    clocks = sampler.stop() if sampler else None
    step(True)
    ms_e2e = timed(True, a.steps)

    # third arm: uint8 1024x1024 images in pinned host memory -> GPU preprocessing (smart_resize -> 980x980) -> scoring
    from llava_reward_b200.processing import Qwen2VLImageProcessorB200
    from llava_reward_b200.synth import hash_randint
    proc = Qwen2VLImageProcessorB200(device=dev)
    ORIG = (1024, 1024)
    assert proc.grid(ORIG) == GRID
    u8 = [hash_randint(f"img.{rank}.{i}", ORIG[0] * ORIG[1] * 3, 0, 256, 7).to(torch.uint8)
          .view(ORIG[0], ORIG[1], 3).pin_memory() for i in range(B)]
    pix_slot = torch.empty_like(resident["pixel_values"])

    def step_u8():
        ib = {k: host[k].to(dev, non_blocking=True) for k in ("input_ids", "attention_mask")}
        pp = proc.preprocess(u8, return_tensors="pt", out=pix_slot)
        ib["pixel_values"], ib["image_grid_thw"] = pp["pixel_values"], pp["image_grid_thw"]
        r, _ = model.custom_forward(inputs_batch=ib)
        rc, rr = r[0::2].contiguous(), r[1::2].contiguous()
        res = torch.cat([rc.float(), rr.float(), eng.preference(rc, rr)[:, None]], dim=1)
        if world > 1:
            dist.all_gather_into_tensor(gather, res)
            res = gather
        return res.cpu()

    step_u8()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step_u8()
    e1.record()
    torch.cuda.synchronize()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_u8 = ms_t.item()
    h2d_u8 = B * ORIG[0] * ORIG[1] * 3 + sum(host[k].numel() * host[k].element_size()
                                             for k in ("input_ids", "attention_mask"))

    if rank == 0:
        peaks = load_peaks()
        n = P * world * a.steps
        value, e2e_value = n / (ms_dev / 1e3), n / (ms_e2e / 1e3)
        lens = host["attention_mask"].sum(1).tolist()
        tf = 2 * sum(tflop_per_sample(cfg, GRID, int(s)) for s in lens) / B
        M = B * S
        K = cfg.hidden_size + (2 * cfg.lora_rank if cfg.use_lora else 0)
        flops = 2.0 * M * (2 * cfg.intermediate_size) * K
        durs = [s.elapsed_time(e) for s, e, _ in prof["gate_up"]]
        ach = flops / (sum(durs) / len(durs) * 1e-3) / 1e12 if durs else None
        line = {
            "metric": "text-image pairs scored/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[3]: Qwen2.5-VL-7B reward (BT head, LoRA r128 on "
                                   f"q/k/v/o/gate/up/down), {P} pairs = {B} samples per step per GPU, every image a "
                                   f"{GRID[0]}x{GRID[1]} patch grid ({GRID[0] * GRID[1]} patches, {n_img_tok} image tokens), "
                                   f"S={S} left-padded, random-init weights",
                       "pairs_per_step_per_gpu": P, "seq_len": S,
                       "parallelism": f"dp{world} (pairs sharded, full bf16 replica per GPU)",
                       "l2_policy": "inputs larger than L2 (1.5 GB of fp32 patches + >1 GB activations per step)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4 * out_w * P * world},
            "e2e_uint8": {"value": n / (ms_u8 / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_u8,
                          "d2h_bytes_per_step": 4 * out_w * P * world,
                          "note": "uint8 1024x1024 images from pinned host memory, Qwen2-VL preprocessing on the GPU "
                                  "(smart_resize to 980x980, lr_resample_u8 bicubic + lr_qwen_patchify_f32), then the "
                                  "same scoring step"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "pair::gemm_pair_kernel<256,SWIGLU> (decoder gate|up + 2 LoRA-B blocks)",
                         "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (ach / peaks["bf16_sustained"]) if ach else None, "traffic": None,
                         "launches_timed": len(durs), "flops_per_launch": flops, "peak_source": peaks["source"]},
            "step_roofline": {"tflop_per_pair": tf, "achieved_tflops_per_gpu": value / world * tf,
                              "frac_of_sustained": value / world * tf / peaks["bf16_sustained"],
                              "frac_of_burst": value / world * tf / peaks["bf16_burst"],
                              "note": "the reference's redundant second vision pass (rw_model_general_preference.py:356) "
                                      "and discarded lm_head GEMM are neither run nor credited"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
