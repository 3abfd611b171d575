#!/usr/bin/env python
"""Benchmark of the LLaVA-v1.6 branch (BASELINE.json configs[4]: "LLaVA-v1.6-7B anyres-672 reward, 64 candidates/prompt
inference-time-scaling batch"). Same JSON-line schema as bench.py; bench.py itself stays on configs[1].

    python tools/bench_llava.py [--steps K] [--warmup W] [--size 7b|13b]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_llava.py --gpus N

One step = one prompt with 64 candidate images: custom_forward(inputs_batch) on 64 samples (672x672 originals ->
5 patches, 2928 image tokens, text U[40,127] tokens first, S = 3057 left-padded), BT head, LoRA r128 on all seven
decoder linears; the 64 rewards are read back. Weak scaling: every rank scores its own prompt.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CANDIDATES = 64
ORIG_HW = (672, 672)
TEXT_LEN_RANGE = (40, 128)


def tflop_per_sample(cfg, n_patches: int, n_img_tokens: int, S: int) -> float:
    """Algorithmic FLOPs (SURVEY.md 8(d) conventions: 2MNK, causal attention at half cost, LoRA unmerged, only real
    patches, the reference's discarded lm_head GEMM NOT credited)."""
    D, DI, T = cfg.clip_hidden, cfg.clip_intermediate, cfg.clip_tokens
    H, I, Lyr, r = cfg.hidden_size, cfg.intermediate_size, cfg.num_layers, (cfg.lora_rank if cfg.use_lora else 0)
    clip = cfg.clip_layers * (2 * T * (4 * D * D + 2 * D * DI) + 4 * T * T * D) + 2 * (T - 1) * 588 * D
    proj = 2 * (T - 1) * (D * H + H * H)
    lin = Lyr * 2 * (4 * H * H + 3 * H * I) * S
    lora = Lyr * 2 * r * (4 * 2 * H + 2 * (H + I) + (I + H)) * S
    attn = Lyr * 2 * S * S * H
    return (n_patches * (clip + proj) + lin + lora + attn + 2 * S * H * cfg.vhd) / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", default="7b", choices=["7b", "13b"])
    ap.add_argument("--candidates", type=int, default=CANDIDATES)
    ap.add_argument("--profile-run", action="store_true")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import yaml

    sys.path.insert(0, ROOT)
    from bench import ClockSampler, load_peaks
    from llava_reward_b200 import _lib as L
    from llava_reward_b200.config import anyres_geometry
    from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor
    from llava_reward_b200.synth import synth_batch_llava

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    ypath = f"/tmp/llava_reward_b200_bench_llava_{rank}.yaml"
    with open(ypath, "w") as f:
        yaml.safe_dump({"is_general_preference": False, "add_cross_attention": False, "value_head_dim": 1,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:1234" + (":13b" if a.size == "13b" else ""), pm_path=None,
                                 cache_dir=None, ft_projector=False, config_overrides={})
    args, model = load_reward_adaptor(args, "llava", ypath)
    model = model.to(dev).eval()
    eng, cfg = model.engine, model.config
    B = a.candidates
    geo = anyres_geometry(ORIG_HW, cfg.image_grid_pinpoints)
    S = 1 + (TEXT_LEN_RANGE[1] - 1) + geo["n_tokens"] + 1
    batch = synth_batch_llava(cfg, B, [ORIG_HW] * B, S, seed=7 + rank, tag="cand", device=dev,
                              text_len_range=TEXT_LEN_RANGE)
    host = {k: v.cpu().pin_memory() for k, v in batch.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    resident["image_sizes"] = host["image_sizes"]  # read on the host by the planner
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("input_ids", "attention_mask", "pixel_values"))
    gather = torch.empty(world * B, dtype=torch.float32, device=dev) if world > 1 else None

    def step(from_host: bool):
        if from_host:
            ib = {k: host[k].to(dev, non_blocking=True) for k in ("input_ids", "attention_mask", "pixel_values")}
            ib["image_sizes"] = host["image_sizes"]
        else:
            ib = resident
        r, _ = model.custom_forward(inputs_batch=ib)
        r = r.float().view(-1)
        if world > 1:
            dist.all_gather_into_tensor(gather, r)
            r = gather
        return r.cpu() if from_host else r

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

    # third arm: uint8 672x672 candidates in pinned host memory -> GPU anyres preprocessing -> scoring
    from llava_reward_b200.processing import LlavaNextImageProcessorB200
    from llava_reward_b200.synth import hash_randint
    proc = LlavaNextImageProcessorB200(cfg.image_grid_pinpoints, device=dev)
    u8 = [hash_randint(f"cand.{rank}.{i}", ORIG_HW[0] * ORIG_HW[1] * 3, 0, 256, 7).to(torch.uint8)
          .view(ORIG_HW[0], ORIG_HW[1], 3).pin_memory() for i in range(B)]
    pix_slot = torch.empty_like(resident["pixel_values"])

    def step_u8():
        ib = {k: host[k].to(dev, non_blocking=True) for k in ("input_ids", "attention_mask")}
        pp = proc.preprocess(u8, return_tensors="pt", out=pix_slot)
        ib["pixel_values"], ib["image_sizes"] = pp["pixel_values"], pp["image_sizes"]
        r, _ = model.custom_forward(inputs_batch=ib)
        r = r.float().view(-1)
        if world > 1:
            dist.all_gather_into_tensor(gather, r)
            r = gather
        return r.cpu()

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
    h2d_u8 = B * ORIG_HW[0] * ORIG_HW[1] * 3 + sum(host[k].numel() * host[k].element_size()
                                                   for k in ("input_ids", "attention_mask"))

    if rank == 0:
        peaks = load_peaks()
        n = B * world * a.steps
        value, e2e_value = n / (ms_dev / 1e3), n / (ms_e2e / 1e3)
        lens = host["attention_mask"].sum(1).tolist()
        tf = sum(tflop_per_sample(cfg, geo["n_patches"], geo["n_tokens"], int(s)) for s in lens) / B
        M = B * S
        K = cfg.hidden_size + (2 * cfg.lora_rank if cfg.use_lora else 0)
        flops = 2.0 * M * (2 * cfg.intermediate_size) * K
        durs = [s.elapsed_time(e) for s, e, _ in prof["gate_up"]]
        ach = flops / (sum(durs) / len(durs) * 1e-3) / 1e12 if durs else None
        line = {
            "metric": "text-image pairs scored/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE.json configs[4]: LLaVA-v1.6-vicuna-{a.size} reward (BT head, LoRA r128 on "
                                   f"q/k/v/o/gate/up/down), anyres-672: {B} candidate images (672x672 -> 5 patches, "
                                   f"{geo['n_tokens']} image tokens) for one prompt per step per GPU, S={S}, "
                                   "random-init weights; a 'pair' = one (prompt, candidate image) sample",
                       "candidates_per_step_per_gpu": B, "seq_len": S,
                       "parallelism": f"dp{world} (prompts sharded, full bf16 replica per GPU)",
                       "l2_policy": "inputs larger than L2 (433 MB of pixels + >1 GB activations per step)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * B * world},
            "e2e_uint8": {"value": n / (ms_u8 / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": h2d_u8,
                          "d2h_bytes_per_step": 4 * B * world,
                          "note": "uint8 672x672 candidates from pinned host memory, anyres preprocessing on the GPU "
                                  "(lr_resample_u8 bicubic + lr_patch_pack_f32), then the same scoring step"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "pair::gemm_pair_kernel<256,SWIGLU> (decoder gate|up + 2 LoRA-B blocks)",
                         "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (ach / peaks["bf16_sustained"]) if ach else None, "traffic": None,
                         "launches_timed": len(durs), "flops_per_launch": flops, "peak_source": peaks["source"]},
            "step_roofline": {"tflop_per_pair": tf, "achieved_tflops_per_gpu": value / world * tf,
                              "frac_of_sustained": value / world * tf / peaks["bf16_sustained"],
                              "frac_of_burst": value / world * tf / peaks["bf16_burst"],
                              "note": "the reference's discarded lm_head GEMM (0.8 TFLOP/sample) is neither run nor credited"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
