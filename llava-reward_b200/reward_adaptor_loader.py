"""Drop-in for the reference's scoring API (reference eval/reward_adaptor_loader.py:24-181):
`load_reward_adaptor`, `inference_process_phi3v`, `preference_compute` with the same signatures,
argument meaning and error behaviour, backed by the B200 engine.
"""
from __future__ import annotations

import os
from typing import Callable

import numpy as np
import torch
import yaml

from .checkpoint import checkpoint_provider, llava_checkpoint_provider, qwen_checkpoint_provider
from .config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig
from .model import B200LlavaNextRewardModel, B200QwenRewardModel, B200RewardModel
from .synth import SynthProvider


def load_reward_adaptor(args, model_type, reward_config_path, load_tokenizer=False):
    """Same contract as reference eval/reward_adaptor_loader.py:24-156: reads the 4 yaml keys into `args`
    (mutating it), builds the reward model for `args.pretrain` + `args.pm_path`, returns (args, model) or
    (args, model, processor, tokenizer). The model is returned un-placed; the caller does `.to('cuda').eval()`.

    `args.pretrain` is either a local directory holding the Phi-3.5-vision HF checkpoint
    (config.json + *.safetensors) or ``synthetic[:seed]`` for random-init weights of the named architecture
    (the only option without network access)."""
    with open(reward_config_path) as f:
        reward_cfg = yaml.safe_load(f)
    args.is_general_preference = reward_cfg['is_general_preference']
    args.add_cross_attention = reward_cfg['add_cross_attention']
    args.value_head_dim = reward_cfg['value_head_dim']
    args.general_preference_tau = reward_cfg['general_preference_tau']
    if model_type not in ('phi3v', 'llava', 'qwen'):
        raise NotImplementedError(f"model_type={model_type!r}: the reference knows 'phi3v', 'qwen' and 'llava'")
    overrides = dict(getattr(args, "config_overrides", None) or {})
    pretrain = str(args.pretrain)
    synthetic = pretrain.startswith("synthetic")
    if model_type == 'qwen':
        # reference :64-109 (Qwen2.5-VL-7B-Instruct + LoRA on the seven decoder linears, optional SkipCA and merger)
        qcfg = QwenVLRewardConfig(is_general_preference=bool(args.is_general_preference),
                                  add_cross_attention=bool(args.add_cross_attention),
                                  value_head_dim=int(args.value_head_dim),
                                  general_preference_tau=float(args.general_preference_tau), **overrides)
        if synthetic:
            parts = [x for x in pretrain.split(":")[1:] if x.isdigit()]
            qprov: Callable[[str], torch.Tensor] = SynthProvider(qcfg, seed=int(parts[0]) if parts else 1234)
        else:
            qcfg, qprov = qwen_checkpoint_provider(qcfg, pretrain, getattr(args, "pm_path", None),
                                                   ft_projector=bool(getattr(args, "ft_projector", False)))
        qmodel = B200QwenRewardModel(qcfg, qprov)
        if load_tokenizer:
            # get_tokenizer_qwen (llava_reward/utils/utils.py:34-44): tokenizer with left padding + the image processor
            # with min_pixels = 256*28*28, max_pixels = 1280*28*28 (here: GPU preprocessing, bit-identical output)
            from .processing import load_processor_qwen
            processor, tokenizer = load_processor_qwen(pretrain, cache_dir=getattr(args, "cache_dir", None),
                                                       use_fast=not getattr(args, "disable_fast_tokenizer", False))
            tokenizer.truncation_side = "right"
            return args, qmodel, processor, tokenizer
        return args, qmodel
    if model_type == 'llava':
        # reference :110-151. add_cross_attention is read from the yaml but the llava branch of custom_forward never
        # applies SkipCA (rw_model_general_preference.py:376-397), so it does not change the forward.
        if synthetic and "13b" in pretrain:
            overrides = {**dict(hidden_size=5120, intermediate_size=13824, num_layers=40, num_heads=40), **overrides}
        lcfg = LlavaNextRewardConfig(is_general_preference=bool(args.is_general_preference),
                                     value_head_dim=int(args.value_head_dim),
                                     general_preference_tau=float(args.general_preference_tau), **overrides)
        if synthetic:
            parts = [x for x in pretrain.split(":")[1:] if x.isdigit()]
            lprov: Callable[[str], torch.Tensor] = SynthProvider(lcfg, seed=int(parts[0]) if parts else 1234)
        else:
            lcfg, lprov = llava_checkpoint_provider(lcfg, pretrain, getattr(args, "pm_path", None),
                                                    ft_projector=bool(getattr(args, "ft_projector", False)))
        lmodel = B200LlavaNextRewardModel(lcfg, lprov)
        if load_tokenizer:
            from .processing import load_processor_llava
            processor, tokenizer = load_processor_llava(pretrain, lcfg, cache_dir=getattr(args, "cache_dir", None),
                                                        use_fast=not getattr(args, "disable_fast_tokenizer", False))
            tokenizer.truncation_side = "right"
            return args, lmodel, processor, tokenizer
        return args, lmodel
    cfg = RewardConfig(is_general_preference=bool(args.is_general_preference),
                       add_cross_attention=bool(args.add_cross_attention),
                       value_head_dim=int(args.value_head_dim),
                       general_preference_tau=float(args.general_preference_tau), **overrides)
    if synthetic:
        seed = int(pretrain.split(":", 1)[1]) if ":" in pretrain else 1234
        provider: Callable[[str], torch.Tensor] = SynthProvider(cfg, seed=seed)
    else:
        cfg, provider = checkpoint_provider(cfg, pretrain, getattr(args, "pm_path", None),
                                            ft_projector=bool(getattr(args, "ft_projector", False)))
    model = B200RewardModel(cfg, provider)
    model.precision = str(getattr(args, "precision", "bf16"))   # "fp32" = verification path (tests only)
    if load_tokenizer:
        from .processing import load_processor
        processor, tokenizer = load_processor(pretrain, cfg, cache_dir=getattr(args, "cache_dir", None),
                                              use_fast=not getattr(args, "disable_fast_tokenizer", False))
        tokenizer.truncation_side = "right"
        return args, model, processor, tokenizer
    return args, model


def inference_process_phi3v(args, processor, tokenizer, img_dir_list, caption, device='cuda'):
    """Same contract as reference eval/reward_adaptor_loader.py:158-173: one BatchFeature-like dict per image
    with input_ids / attention_mask / pixel_values / image_sizes on `device`."""
    from PIL import Image
    img_list = [Image.open(d).convert("RGB") for d in img_dir_list]
    prompt_messages = {'role': 'user', 'content': f'<|image_1|>\n{caption}'}
    prompt = tokenizer.apply_chat_template([prompt_messages], tokenize=False, add_generation_prompt=True)[:-22] \
        + tokenizer.eos_token
    img_inputs = []
    for img in img_list:
        img_input = processor(text=prompt, images=[img], return_tensors="pt", padding=True, truncation=True)
        for k in img_input:
            img_input[k] = img_input[k].to(device)
        img_inputs.append(img_input)
    return img_inputs


def preference_compute(args, chosen_rewards, reject_rewards):
    """Same contract as reference eval/reward_adaptor_loader.py:174-181 -> np.float32 [B] (synchronises)."""
    from . import ops
    if not (torch.is_tensor(chosen_rewards) and chosen_rewards.is_cuda):
        raise RuntimeError("preference_compute expects the CUDA reward tensors returned by custom_forward")
    n, vhd = chosen_rewards.shape
    prob = torch.empty(n, dtype=torch.float32, device=chosen_rewards.device)
    with torch.cuda.device(chosen_rewards.device):
        # rewards keep the dtype custom_forward produced them in: bf16 (product) or fp32 (verification path)
        dt = torch.float32 if chosen_rewards.dtype == torch.float32 else torch.bfloat16
        ops.preference(chosen_rewards.to(dt).contiguous(), reject_rewards.to(dt).contiguous(),
                       prob, n, vhd, bool(args.is_general_preference) and int(args.value_head_dim) == 2,
                       float(args.general_preference_tau))
    return prob.cpu().numpy()
