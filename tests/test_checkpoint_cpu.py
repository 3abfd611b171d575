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
        json.dump({"r": 2, "lora_alpha": 128}, f)
    torch.save({"value_head.weight": t(6), "W_q.weight": t(7), "W_k.weight": t(8), "W_v.weight": t(9),
                "ca_layernorm.weight": t(10), "model.vision_embed_tokens.img_projection.0.weight": t(11)},
               pm / "pytorch_model.bin")
    cfg, get = checkpoint_provider(RewardConfig(), str(base), str(pm), ft_projector=True)
    assert cfg.use_lora and cfg.lora_rank == 2 and cfg.lora_alpha == 128 and cfg.long_factor == [2.0] * 48
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
        json.dump({"r": 2, "lora_alpha": 256}, f)
    torch.save({"value_head.weight": t(8), "W_q.weight": t(9), "ca_layernorm.weight": t(10),
                "visual.merger.ln_q.weight": t(11), "visual.merger.mlp.0.weight": t(12),
                "visual.merger.mlp.2.bias": t(13)}, pm / "pytorch_model.bin")
    cfg, get = qwen_checkpoint_provider(QwenVLRewardConfig(), str(base), str(pm), ft_projector=True)
    assert cfg.use_lora and cfg.lora_rank == 2 and cfg.num_kv_heads == 4 and cfg.vit_fullatt == [7, 15, 23, 31]
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


def _phi_dirs(tmp_path, adapter, adapter_cfg, base_extra=None):
    base, pm = tmp_path / "base", tmp_path / "pm"
    os.makedirs(base)
    os.makedirs(pm / "lora")
    with open(base / "config.json", "w") as f:
        json.dump({"hidden_size": 3072}, f)
    g = torch.Generator().manual_seed(0)
    sd = {"model.layers.0.self_attn.qkv_proj.weight": torch.randn(6, 4, generator=g),
          "model.vision_embed_tokens.img_processor.vision_model.encoder.layers.0.self_attn.q_proj.weight":
              torch.randn(4, 4, generator=g),
          "model.vision_embed_tokens.img_projection.0.weight": torch.randn(4, 4, generator=g)}
    sd.update(base_extra or {})
    torch.save(sd, base / "pytorch_model.bin")
    if adapter is not None:
        torch.save(adapter, pm / "lora" / "adapter_model.bin")
    if adapter_cfg is not None:
        with open(pm / "lora" / "adapter_config.json", "w") as f:
            json.dump(adapter_cfg, f)
    torch.save({"value_head.weight": torch.ones(2, 4)}, pm / "pytorch_model.bin")
    return str(base), str(pm), sd


def test_vision_and_projector_lora_is_folded_not_dropped(tmp_path):
    """create_lora_config's default (freeze_vision_model=False) also adapts the CLIP linears and img_projection.0/.2
    (reference llava_reward/utils/utils.py:194-222): those deltas are merged into the dense weights, the decoder's stay
    un-merged for the K-extension GEMM, and an adapter key without a home raises instead of vanishing."""
    import pytest
    g = torch.Generator().manual_seed(1)
    r = 2
    clip = "model.vision_embed_tokens.img_processor.vision_model.encoder.layers.0.self_attn.q_proj"
    proj = "model.vision_embed_tokens.img_projection.0"
    dec = "model.layers.0.self_attn.qkv_proj"
    ad = {}
    for m, (o, i) in ((clip, (4, 4)), (proj, (4, 4)), (dec, (6, 4))):
        ad[f"base_model.model.{m}.lora_A.weight"] = torch.randn(r, i, generator=g)
        ad[f"base_model.model.{m}.lora_B.weight"] = torch.randn(o, r, generator=g)
    base, pm, sd = _phi_dirs(tmp_path, ad, {"r": r, "lora_alpha": 4})
    cfg, get = checkpoint_provider(RewardConfig(), base, pm)
    assert cfg.lora_scale == 2.0
    for m in (clip, proj):
        want = sd[m + ".weight"] + 2.0 * ad[f"base_model.model.{m}.lora_B.weight"] @ ad[f"base_model.model.{m}.lora_A.weight"]
        assert torch.allclose(get(m + ".weight"), want, atol=1e-6)
        with pytest.raises(KeyError):
            get(m + ".lora_A.weight")
    assert torch.equal(get(dec + ".weight"), sd[dec + ".weight"])
    assert torch.equal(get(dec + ".lora_B.weight"), ad[f"base_model.model.{dec}.lora_B.weight"])
    # an adapted module the base checkpoint does not have
    ad2 = dict(ad)
    ad2["base_model.model.model.layers.7.mlp.down_proj.lora_A.weight"] = torch.zeros(r, 4)
    ad2["base_model.model.model.layers.7.mlp.down_proj.lora_B.weight"] = torch.zeros(4, r)
    torch.save(ad2, os.path.join(pm, "lora", "adapter_model.bin"))
    with pytest.raises(KeyError, match="layers.7.mlp.down_proj"):
        checkpoint_provider(RewardConfig(), base, pm)


def test_adapter_errors_and_options(tmp_path):
    import pytest
    ad = {"base_model.model.model.layers.0.self_attn.qkv_proj.lora_A.weight": torch.ones(2, 4),
          "base_model.model.model.layers.0.self_attn.qkv_proj.lora_B.weight": torch.ones(6, 2)}
    base, pm, _ = _phi_dirs(tmp_path, None, None)
    with pytest.raises(FileNotFoundError, match="adapter"):      # pm_path without lora/adapter_model.*: load_adapter raises
        checkpoint_provider(RewardConfig(), base, pm)
    torch.save(ad, os.path.join(pm, "lora", "adapter_model.bin"))
    for bad in ({"r": 2, "lora_alpha": 4, "rank_pattern": {"qkv_proj": 4}}, {"r": 2, "lora_alpha": 4, "use_dora": True},
                {"r": 2, "lora_alpha": 4, "alpha_pattern": {"o_proj": 1}}):
        with open(os.path.join(pm, "lora", "adapter_config.json"), "w") as f:
            json.dump(bad, f)
        with pytest.raises(NotImplementedError):
            checkpoint_provider(RewardConfig(), base, pm)
    with open(os.path.join(pm, "lora", "adapter_config.json"), "w") as f:
        json.dump({"r": 4, "lora_alpha": 4}, f)
    with pytest.raises(ValueError, match="rank"):
        checkpoint_provider(RewardConfig(), base, pm)
    with open(os.path.join(pm, "lora", "adapter_config.json"), "w") as f:
        json.dump({"r": 2, "lora_alpha": 4, "use_rslora": True}, f)
    cfg, _ = checkpoint_provider(RewardConfig(), base, pm)
    assert abs(cfg.lora_scale - 4 / 2 ** 0.5) < 1e-12          # peft: lora_alpha / sqrt(r)


def test_lora_rank_is_padded_to_gemm_granularity():
    """adapters of any rank run: A / B are zero-padded to a multiple of 128 at pack time (lr_gemm_bf16 needs N % 128)"""
    from llava_reward_b200.weights import _pad_rank
    A, B = torch.randn(16, 64), torch.randn(32, 16)
    Ap, Bp = _pad_rank(A, B)
    assert Ap.shape == (128, 64) and Bp.shape == (32, 128)
    assert torch.equal(Ap[:16], A) and not Ap[16:].any() and torch.equal(Bp[:, :16], B) and not Bp[:, 16:].any()
    assert torch.allclose(Bp @ Ap, B @ A)
    A2, B2 = _pad_rank(torch.randn(128, 8), torch.randn(8, 128))
    assert A2.shape == (128, 8) and B2.shape == (8, 128)
