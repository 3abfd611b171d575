"""Checkpoint ingestion: the reference's `save_model_lora` directory layout + a HF-style base
checkpoint are mapped to the parameter names `pack_weights` consumes (reference eval/reward_adaptor_loader.py:43-60,
llava_reward/utils/deepspeed.py:333-417). Uses tiny fake tensors; only names / routing are checked."""
import json
import os

import torch

from llava_reward_b200.checkpoint import checkpoint_provider
from llava_reward_b200.config import RewardConfig


def test_checkpoint_directory_layout(tmp_path):
    base, pm = tmp_path / "base", tmp_path / "pm"
    os.makedirs(base)
    os.makedirs(pm / "lora")
    with open(base / "config.json", "w") as f:
        json.dump({"vocab_size": 32064, "hidden_size": 3072, "intermediate_size": 8192, "num_hidden_layers": 32,
                   "num_attention_heads": 32, "rms_norm_eps": 1e-5, "rope_theta": 10000.0,
                   "max_position_embeddings": 131072, "original_max_position_embeddings": 4096,
                   "rope_scaling": {"type": "su", "short_factor": [1.0] * 48, "long_factor": [2.0] * 48}}, f)
    t = lambda v: torch.full((2, 2), float(v))  # noqa: E731
    torch.save({"model.embed_tokens.weight": t(1), "model.vision_embed_tokens.wte.weight": t(1),
                "model.layers.0.self_attn.qkv_proj.weight": t(2),
                "model.vision_embed_tokens.img_projection.0.weight": t(3)}, base / "pytorch_model.bin")
    torch.save({"base_model.model.model.layers.0.self_attn.qkv_proj.lora_A.default.weight": t(4),
                "base_model.model.model.layers.0.self_attn.qkv_proj.lora_B.weight": t(5)}, pm / "lora" / "adapter_model.bin")
    with open(pm / "lora" / "adapter_config.json", "w") as f:
        json.dump({"r": 64, "lora_alpha": 128}, f)
    torch.save({"value_head.weight": t(6), "W_q.weight": t(7), "W_k.weight": t(8), "W_v.weight": t(9),
                "ca_layernorm.weight": t(10), "model.vision_embed_tokens.img_projection.0.weight": t(11)},
               pm / "pytorch_model.bin")
    cfg, get = checkpoint_provider(RewardConfig(), str(base), str(pm), ft_projector=True)
    assert cfg.use_lora and cfg.lora_rank == 64 and cfg.lora_alpha == 128 and cfg.long_factor == [2.0] * 48
    assert get("model.embed_tokens.weight")[0, 0] == 1
    assert get("model.layers.0.self_attn.qkv_proj.weight")[0, 0] == 2
    assert get("model.layers.0.self_attn.qkv_proj.lora_A.weight")[0, 0] == 4
    assert get("model.layers.0.self_attn.qkv_proj.lora_B.weight")[0, 0] == 5
    assert get("value_head.weight")[0, 0] == 6 and get("W_q.weight")[0, 0] == 7 and get("W_v.weight")[0, 0] == 9
    assert get("ca_layernorm.weight")[0, 0] == 10
    assert get("model.vision_embed_tokens.img_projection.0.weight")[0, 0] == 11   # ft_projector overrides the base
    cfg2, get2 = checkpoint_provider(RewardConfig(), str(base), str(pm), ft_projector=False)
    assert get2("model.vision_embed_tokens.img_projection.0.weight")[0, 0] == 3
    try:
        get("model.layers.1.self_attn.qkv_proj.weight")
        assert False
    except KeyError:
        pass
