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


def test_qwen_checkpoint_directory_layout(tmp_path):
    """Qwen2.5-VL: both transformers-era parameter prefixes map to the 4.50 names the reference was written against;
    the reward heads / merger keys are selected as reference eval/reward_adaptor_loader.py:80-105 does."""
    from llava_reward_b200.checkpoint import qwen_checkpoint_provider
    from llava_reward_b200.config import QwenVLRewardConfig
    base, pm = tmp_path / "qbase", tmp_path / "qpm"
    os.makedirs(base)
    os.makedirs(pm / "lora")
    with open(base / "config.json", "w") as f:
        json.dump({"text_config": {"vocab_size": 152064, "hidden_size": 3584, "intermediate_size": 18944,
                                   "num_hidden_layers": 28, "num_attention_heads": 28, "num_key_value_heads": 4,
                                   "rms_norm_eps": 1e-6,
                                   "rope_parameters": {"rope_theta": 1000000.0, "mrope_section": [16, 24, 24]}},
                   "vision_config": {"depth": 32, "hidden_size": 1280, "intermediate_size": 3420, "num_heads": 16,
                                     "window_size": 112, "fullatt_block_indexes": [7, 15, 23, 31]},
                   "image_token_id": 151655}, f)
    t = lambda v: torch.full((2, 2), float(v))  # noqa: E731
    torch.save({"model.language_model.embed_tokens.weight": t(1),                 # transformers 5.x names
                "model.language_model.layers.0.self_attn.q_proj.weight": t(2),
                "model.visual.blocks.0.attn.qkv.weight": t(3),
                "visual.merger.mlp.0.weight": t(4),                                # 4.50 names pass through
                "model.layers.1.mlp.up_proj.weight": t(5)}, base / "pytorch_model.bin")
    torch.save({"base_model.model.model.layers.0.self_attn.q_proj.lora_A.default.weight": t(6),
                "base_model.model.model.language_model.layers.0.self_attn.q_proj.lora_B.weight": t(7)},
               pm / "lora" / "adapter_model.bin")
    with open(pm / "lora" / "adapter_config.json", "w") as f:
        json.dump({"r": 128, "lora_alpha": 256}, f)
    torch.save({"value_head.weight": t(8), "W_q.weight": t(9), "ca_layernorm.weight": t(10),
                "visual.merger.ln_q.weight": t(11), "visual.merger.mlp.0.weight": t(12),
                "visual.merger.mlp.2.bias": t(13)}, pm / "pytorch_model.bin")
    cfg, get = qwen_checkpoint_provider(QwenVLRewardConfig(), str(base), str(pm), ft_projector=True)
    assert cfg.use_lora and cfg.lora_rank == 128 and cfg.num_kv_heads == 4 and cfg.vit_fullatt == [7, 15, 23, 31]
    assert get("model.embed_tokens.weight")[0, 0] == 1
    assert get("model.layers.0.self_attn.q_proj.weight")[0, 0] == 2
    assert get("visual.blocks.0.attn.qkv.weight")[0, 0] == 3
    assert get("model.layers.1.mlp.up_proj.weight")[0, 0] == 5
    assert get("model.layers.0.self_attn.q_proj.lora_A.weight")[0, 0] == 6
    assert get("model.layers.0.self_attn.q_proj.lora_B.weight")[0, 0] == 7
    assert get("value_head.weight")[0, 0] == 8 and get("W_q.weight")[0, 0] == 9 and get("ca_layernorm.weight")[0, 0] == 10
    assert get("visual.merger.ln_q.weight")[0, 0] == 11 and get("visual.merger.mlp.0.weight")[0, 0] == 12
    assert get("visual.merger.mlp.2.bias")[0, 0] == 13
    _, get2 = qwen_checkpoint_provider(QwenVLRewardConfig(), str(base), str(pm), ft_projector=False)
    assert get2("visual.merger.mlp.0.weight")[0, 0] == 4


def test_qwen_loader_and_model_surface(tmp_path):
    """load_reward_adaptor(..., 'qwen', ...) mutates args like the reference and returns an un-placed model that refuses
    to run without a GPU / with the phi3v calling convention."""
    import types

    import pytest
    import yaml
    from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor
    y = tmp_path / "reward_config.yaml"
    with open(y, "w") as f:
        yaml.safe_dump({"is_general_preference": True, "add_cross_attention": True, "value_head_dim": 2,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:5", pm_path=None, cache_dir=None, ft_projector=False,
                                 config_overrides=dict(num_layers=1, vit_depth=1, vit_fullatt=[0]))
    args, model = load_reward_adaptor(args, "qwen", str(y))
    assert args.is_general_preference and args.add_cross_attention and args.value_head_dim == 2
    assert model.model_type == "qwen" and model.config.vhd == 2 and model.config.num_kv_heads == 4
    with pytest.raises(RuntimeError):
        model.custom_forward(inputs_batch={})
    with pytest.raises(NotImplementedError):
        load_reward_adaptor(args, "blip", str(y))
