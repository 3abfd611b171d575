"""Write synthetic weights to disk in the directory layouts the reference consumes, for checkpoint round-trip tests
 and for users who want to try the loader without network access:

  <base_dir>/   a HF checkpoint directory: config.json + sharded model-0000i-of-0000n.safetensors (bf16) under the
                state_dict names of the backbone (Phi-3.5-vision / llava-v1.6 / Qwen2.5-VL)
  <pm_dir>/     the reference's `save_model_lora` output (llava_reward/utils/deepspeed.py:333-417):
                reward_config.yaml, pytorch_model.bin (value_head, W_q/W_k/W_v, ca_layernorm + the fine-tuned
                projector: img_projection / multi_modal_projector / merger), lora/adapter_model.bin with PEFT key
                names (`base_model.model.<module>.lora_A.weight`), lora/adapter_config.json

    python tools/write_checkpoint.py phi3v /tmp/base /tmp/pm [--seed 1234] [--layers 2 --clip-layers 2]

`load_reward_adaptor(args(pretrain=base_dir, pm_path=pm_dir, ft_projector=True), model_type, pm_dir/reward_config.yaml)`
then yields rewards bit-identical to `pretrain="synthetic:<seed>"` (tests/test_checkpoint_gpu.py).
With `decoy_projector` the projector stored in the BASE checkpoint is negated, so the round trip only succeeds if the
`ft_projector` override from pytorch_model.bin is honoured (reference eval/reward_adaptor_loader.py:57-59, 93-103, 146-148).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HEAD_MODULES = ("value_head", "W_q", "W_k", "W_v", "ca_layernorm")
PROJECTOR = {"phi3v": "img_projection", "llava": "multi_modal_projector", "qwen": "merger"}


def hf_config(cfg, model_type: str) -> dict:
    if model_type == "phi3v":
        return {"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size,
                "num_hidden_layers": cfg.num_layers, "num_attention_heads": cfg.num_heads, "rms_norm_eps": cfg.rms_eps,
                "rope_theta": cfg.rope_theta, "max_position_embeddings": cfg.max_position_embeddings,
                "original_max_position_embeddings": cfg.original_max_position_embeddings,
                "rope_scaling": {"type": "su", "short_factor": cfg.short_factor, "long_factor": cfg.long_factor}}
    if model_type == "llava":
        return {"text_config": {"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                                "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers,
                                "num_attention_heads": cfg.num_heads, "num_key_value_heads": cfg.num_heads,
                                "rms_norm_eps": cfg.rms_eps, "rope_theta": cfg.rope_theta},
                "image_token_index": cfg.image_token_id, "image_grid_pinpoints": cfg.image_grid_pinpoints}
    if model_type == "qwen":
        return {"text_config": {"vocab_size": cfg.vocab_size, "hidden_size": cfg.hidden_size,
                                "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_layers,
                                "num_attention_heads": cfg.num_heads, "num_key_value_heads": cfg.num_kv_heads,
                                "rms_norm_eps": cfg.rms_eps,
                                "rope_parameters": {"rope_theta": cfg.rope_theta, "mrope_section": cfg.mrope_section}},
                "vision_config": {"depth": cfg.vit_depth, "hidden_size": cfg.vit_hidden,
                                  "intermediate_size": cfg.vit_intermediate, "num_heads": cfg.vit_heads,
                                  "window_size": cfg.vit_window, "fullatt_block_indexes": cfg.vit_fullatt},
                "image_token_id": cfg.image_token_id}
    raise ValueError(model_type)


def base_name(name: str, model_type: str) -> str:
    """state_dict name as the backbone's HF checkpoint stores it: llava is written with the transformers-5.x prefixes
    (`model.language_model.*`, `model.vision_tower.*`, ...) so that the loader's era mapping is exercised; phi3v and
    qwen with the 4.50 names the reference was written against."""
    if model_type == "llava":
        if name.startswith("language_model.model."):
            return "model.language_model." + name[len("language_model.model."):]
        if name.startswith(("vision_tower.", "multi_modal_projector.", "image_newline")):
            return "model." + name
    return name


def write_reference_layout(cfg, model_type: str, seed: int, base_dir: str, pm_dir: str, device="cpu",
                           decoy_projector: bool = True, n_shards: int = 2) -> None:
    from safetensors.torch import save_file

    from llava_reward_b200.synth import SynthProvider

    prov = SynthProvider(cfg, seed=seed, device=device)
    os.makedirs(base_dir, exist_ok=True)
    os.makedirs(os.path.join(pm_dir, "lora"), exist_ok=True)
    with open(os.path.join(base_dir, "config.json"), "w") as f:
        json.dump(hf_config(cfg, model_type), f, indent=1)
    with open(os.path.join(pm_dir, "reward_config.yaml"), "w") as f:
        yaml.safe_dump({"is_general_preference": bool(cfg.is_general_preference),
                        "add_cross_attention": bool(cfg.add_cross_attention), "value_head_dim": int(cfg.value_head_dim),
                        "general_preference_tau": float(cfg.general_preference_tau)}, f)
    base, heads, lora = {}, {}, {}
    for name in prov.names():
        t = prov(name).to(torch.bfloat16).cpu().contiguous()
        if ".lora_" in name:
            lora["base_model.model." + base_name(name, model_type)] = t
        elif name.split(".")[0] in HEAD_MODULES:
            heads[name] = t
        else:
            key = base_name(name, model_type)
            if PROJECTOR[model_type] in name:
                heads[key] = t
                base[key] = (-t).contiguous() if decoy_projector else t
            else:
                base[key] = t
            if model_type == "phi3v" and name == "model.embed_tokens.weight":
                base["model.vision_embed_tokens.wte.weight"] = t.clone()   # the HF checkpoint stores both (untied)
    keys = sorted(base)
    per = (len(keys) + n_shards - 1) // n_shards
    for i in range(n_shards):
        part = {k: base[k] for k in keys[i * per:(i + 1) * per]}
        if part:
            save_file(part, os.path.join(base_dir, f"model-{i + 1:05d}-of-{n_shards:05d}.safetensors"))
    torch.save(heads, os.path.join(pm_dir, "pytorch_model.bin"))
    if cfg.use_lora:
        torch.save(lora, os.path.join(pm_dir, "lora", "adapter_model.bin"))
        with open(os.path.join(pm_dir, "lora", "adapter_config.json"), "w") as f:
            json.dump({"peft_type": "LORA", "r": cfg.lora_rank, "lora_alpha": cfg.lora_alpha, "lora_dropout": 0.0,
                       "init_lora_weights": "gaussian", "bias": "none"}, f, indent=1)


def main():
    from llava_reward_b200.config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig

    ap = argparse.ArgumentParser()
    ap.add_argument("model_type", choices=["phi3v", "llava", "qwen"])
    ap.add_argument("base_dir")
    ap.add_argument("pm_dir")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--clip-layers", type=int, default=None)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    over = {}
    if a.layers is not None:
        over["num_layers"] = a.layers
    if a.clip_layers is not None:
        over["vit_depth" if a.model_type == "qwen" else "clip_layers"] = a.clip_layers
    cfg = {"phi3v": RewardConfig, "llava": LlavaNextRewardConfig, "qwen": QwenVLRewardConfig}[a.model_type](**over)
    write_reference_layout(cfg, a.model_type, a.seed, a.base_dir, a.pm_dir, device=a.device)
    print(f"wrote {a.base_dir} and {a.pm_dir}")


if __name__ == "__main__":
    main()
