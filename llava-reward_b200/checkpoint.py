"""Checkpoint ingestion: HF Phi-3.5-vision safetensors + the reference's `save_model_lora` directory layout
(reference llava_reward/utils/deepspeed.py:333-417: pytorch_model.bin with value_head / W_q / W_k / W_v /
ca_layernorm / img_projection keys, lora/adapter_model.{bin,safetensors}, reward_config.yaml) -> a
``name -> tensor`` provider with the reference state_dict names that `pack_weights` consumes.
"""
from __future__ import annotations

import glob
import json
import os
import re
from typing import Callable, Dict, Optional, Tuple

import torch

from .config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig


_PHI3_DECODER_LORA = re.compile(r"^model\.layers\.\d+\.(self_attn\.(qkv_proj|o_proj)|mlp\.(gate_up_proj|down_proj))$")
_LLAMA_DECODER_LORA = re.compile(r"^(language_model\.)?model\.layers\.\d+\.(self_attn\.[qkvo]_proj|mlp\.(gate|up|down)_proj)$")


def _load_adapter(cfg, pm_path: str, canonical: Callable[[str], str]) -> Dict[str, torch.Tensor]:
    """PEFT adapter of the reference's `save_model_lora` layout (<pm_path>/lora/adapter_model.{safetensors,bin} +
    adapter_config.json, reference deepspeed.py:391-398) -> {canonical module name + '.lora_A.weight' / '.lora_B.weight'}.
    Sets cfg.use_lora / lora_rank / lora_alpha. The reference's `model.load_adapter(...)`
    (eval/reward_adaptor_loader.py:44) raises when the adapter is missing - so does this; adapter options whose
    arithmetic the engine does not implement are rejected instead of being silently ignored."""
    found = None
    for cand in ("adapter_model.safetensors", "adapter_model.bin"):
        pth = os.path.join(pm_path, "lora", cand)
        if os.path.exists(pth):
            found = pth
            break
    if found is None:
        raise FileNotFoundError(f"no LoRA adapter under {os.path.join(pm_path, 'lora')!r} (adapter_model.safetensors / "
                                "adapter_model.bin): the reference's load_adapter fails here as well")
    out = {}
    for k, v in _load_file(found).items():
        out[canonical(k.replace("base_model.model.", "", 1).replace(".default", ""))] = v
    cfg.use_lora = True
    acfg = os.path.join(pm_path, "lora", "adapter_config.json")
    if os.path.exists(acfg):
        with open(acfg) as f:
            a = json.load(f)
        for opt in ("rank_pattern", "alpha_pattern"):
            if a.get(opt):
                raise NotImplementedError(f"adapter_config.json sets {opt}: per-module LoRA ranks / alphas are not "
                                          "supported (the reference's create_lora_config* never sets them)")
        for opt in ("use_dora", "bias"):
            if a.get(opt) not in (None, False, "none"):
                raise NotImplementedError(f"adapter_config.json sets {opt}={a.get(opt)!r}: not supported")
        cfg.lora_rank = int(a.get("r", cfg.lora_rank))
        cfg.lora_alpha = float(a.get("lora_alpha", cfg.lora_alpha))
        if a.get("use_rslora"):
            # peft: scaling = lora_alpha / sqrt(r); expressed through alpha so that cfg.lora_scale = alpha / r holds
            cfg.lora_alpha = cfg.lora_alpha * (cfg.lora_rank ** 0.5)
    ranks = {v.shape[0] for k, v in out.items() if k.endswith(".lora_A.weight")}
    if len(ranks) > 1:
        raise NotImplementedError(f"adapter holds LoRA matrices of different ranks {sorted(ranks)}")
    if ranks and ranks != {cfg.lora_rank}:
        raise ValueError(f"adapter rank {ranks.pop()} does not match adapter_config.json r={cfg.lora_rank}")
    return out


def _merge_adapter(tensors: Dict[str, torch.Tensor], adapter: Dict[str, torch.Tensor], scale: float, unmerged) -> None:
    """Put the adapter into `tensors`. Modules matched by `unmerged` (the decoder linears) keep their lora_A / lora_B
    matrices: the engine applies them un-merged through the K-extension of the GEMM, op for op like peft's
    lora.Linear.forward. EVERY OTHER adapted module - the reference's default create_lora_config also targets the CLIP
    q/k/v/out_proj/fc1/fc2 and img_projection.0/.2 (llava_reward/utils/utils.py:194-222; freeze_vision_model=False) - is
    folded into its dense weight, W += (alpha/r) B A in fp32 (the merged form of the same linear map; it differs from
    the un-merged bf16 evaluation only by rounding). A lora key without a base weight raises: nothing is dropped."""
    mods = sorted({k[: -len(".lora_A.weight")] for k in adapter if k.endswith(".lora_A.weight")})
    seen = set()
    for m in mods:
        ka, kb = m + ".lora_A.weight", m + ".lora_B.weight"
        if kb not in adapter:
            raise KeyError(f"adapter has {ka} but no {kb}")
        seen.update((ka, kb))
        base = m + ".weight"
        if base not in tensors and m + ".base_layer.weight" in tensors:
            tensors[base] = tensors.pop(m + ".base_layer.weight")
        if base not in tensors:
            raise KeyError(f"LoRA adapter targets {m!r} but the checkpoint has no {base!r}")
        if unmerged.match(m):
            tensors[ka], tensors[kb] = adapter[ka], adapter[kb]
        else:
            tensors[base] = tensors[base].float() + scale * (adapter[kb].float() @ adapter[ka].float())
    extra = sorted(k for k in adapter if k not in seen)
    if extra:
        raise KeyError(f"adapter keys the loader does not understand: {extra[:5]}{' ...' if len(extra) > 5 else ''}")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def checkpoint_provider(cfg: RewardConfig, pretrain_dir: str, pm_path: Optional[str],
                        ft_projector: bool = False) -> Tuple[RewardConfig, Callable[[str], torch.Tensor]]:
    if not os.path.isdir(pretrain_dir):
        raise FileNotFoundError(f"args.pretrain={pretrain_dir!r} is not a local directory (no network access: "
                                "hub ids cannot be resolved) - use 'synthetic' for random-init weights")
    with open(os.path.join(pretrain_dir, "config.json")) as f:
        hf = json.load(f)
    rs = hf.get("rope_scaling") or {}
    cfg.vocab_size = hf.get("vocab_size", cfg.vocab_size)
    cfg.hidden_size = hf.get("hidden_size", cfg.hidden_size)
    cfg.intermediate_size = hf.get("intermediate_size", cfg.intermediate_size)
    cfg.num_layers = hf.get("num_hidden_layers", cfg.num_layers)
    cfg.num_heads = hf.get("num_attention_heads", cfg.num_heads)
    cfg.rms_eps = hf.get("rms_norm_eps", cfg.rms_eps)
    cfg.rope_theta = hf.get("rope_theta", cfg.rope_theta)
    cfg.max_position_embeddings = hf.get("max_position_embeddings", cfg.max_position_embeddings)
    cfg.original_max_position_embeddings = hf.get("original_max_position_embeddings",
                                                   cfg.original_max_position_embeddings)
    if rs:
        cfg.short_factor, cfg.long_factor = list(rs["short_factor"]), list(rs["long_factor"])
    tensors: Dict[str, torch.Tensor] = {}
    files = sorted(glob.glob(os.path.join(pretrain_dir, "*.safetensors"))) or \
        sorted(glob.glob(os.path.join(pretrain_dir, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {pretrain_dir}")
    for fpath in files:
        tensors.update(_load_file(fpath))
    tensors = {k.replace("model.vision_embed_tokens.wte.", "model.embed_tokens."): v for k, v in tensors.items()}
    cfg.use_lora = False
    if pm_path:
        # LoRA adapter saved by PEFT: keys 'base_model.model.<module>.lora_A.weight' (maybe with '.default')
        adapter = _load_adapter(cfg, pm_path, lambda k: k.replace("model.vision_embed_tokens.wte.", "model.embed_tokens."))
        heads = os.path.join(pm_path, "pytorch_model.bin")
        if os.path.exists(heads):
            sd = torch.load(heads, map_location="cpu", weights_only=True)
            # key selection mirrors reference eval/reward_adaptor_loader.py:46-60
            for k, v in sd.items():
                leaf = k.split(".")[-1]
                for mod in ("value_head", "W_q", "W_k", "W_v", "ca_layernorm"):
                    if mod in k:
                        tensors[f"{mod}.{leaf}"] = v
                if ft_projector and "img_projection" in k and ".lora_" not in k:
                    # reference :57-59 keeps the last two name components ('0.weight'); a peft-wrapped projector saves
                    # '<i>.base_layer.weight' (its lora_A/B live in the adapter file)
                    parts = [x for x in k.split(".") if x != "base_layer"]
                    tensors["model.vision_embed_tokens.img_projection." + ".".join(parts[-2:])] = v
        _merge_adapter(tensors, adapter, cfg.lora_scale, _PHI3_DECODER_LORA)

    def get(name: str) -> torch.Tensor:
        if name not in tensors:
            raise KeyError(f"checkpoint is missing parameter {name!r}")
        return tensors[name]

    return cfg, get


def _canonical_llava_name(k: str) -> str:
    """Any era of transformers LlavaNext state_dict name -> the 4.50 names the reference was written against
    (`language_model.model.*`, `vision_tower.*`, `multi_modal_projector.*`, `image_newline`)."""
    if k.startswith("model.language_model."):
        return "language_model.model." + k[len("model.language_model."):]
    if k.startswith(("model.vision_tower.", "model.multi_modal_projector.", "model.image_newline")):
        return k[len("model."):]
    return k


def llava_checkpoint_provider(cfg: LlavaNextRewardConfig, pretrain_dir: str, pm_path: Optional[str],
                              ft_projector: bool = False):
    """HF llava-v1.6-vicuna checkpoint directory + the reference's save_model_lora layout for the llava branch
    (key selection of reference eval/reward_adaptor_loader.py:124-148) -> provider with synth.llava_param_specs names."""
    if not os.path.isdir(pretrain_dir):
        raise FileNotFoundError(f"args.pretrain={pretrain_dir!r} is not a local directory (no network access: "
                                "hub ids cannot be resolved) - use 'synthetic' for random-init weights")
    with open(os.path.join(pretrain_dir, "config.json")) as f:
        hf = json.load(f)
    t = hf.get("text_config", {})
    cfg.vocab_size = t.get("vocab_size", cfg.vocab_size)
    cfg.hidden_size = t.get("hidden_size", cfg.hidden_size)
    cfg.intermediate_size = t.get("intermediate_size", cfg.intermediate_size)
    cfg.num_layers = t.get("num_hidden_layers", cfg.num_layers)
    cfg.num_heads = t.get("num_attention_heads", cfg.num_heads)
    if t.get("num_key_value_heads", cfg.num_heads) != cfg.num_heads:
        raise NotImplementedError("grouped-query attention (e.g. llava-v1.6-mistral): the reference's LoRA config and "
                                  "training script target the Vicuna (MHA) checkpoints only")
    cfg.rms_eps = t.get("rms_norm_eps", cfg.rms_eps)
    cfg.rope_theta = t.get("rope_theta", cfg.rope_theta)
    cfg.image_token_id = hf.get("image_token_index", hf.get("image_token_id", cfg.image_token_id))
    cfg.image_grid_pinpoints = hf.get("image_grid_pinpoints", cfg.image_grid_pinpoints)
    tensors: Dict[str, torch.Tensor] = {}
    files = sorted(glob.glob(os.path.join(pretrain_dir, "*.safetensors"))) or \
        sorted(glob.glob(os.path.join(pretrain_dir, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {pretrain_dir}")
    for fpath in files:
        tensors.update({_canonical_llava_name(k): v for k, v in _load_file(fpath).items()})
    cfg.use_lora = False
    if pm_path:
        adapter = _load_adapter(cfg, pm_path, _canonical_llava_name)
        heads = os.path.join(pm_path, "pytorch_model.bin")
        if os.path.exists(heads):
            sd = torch.load(heads, map_location="cpu", weights_only=True)
            for k, v in sd.items():
                if "value_head" in k:
                    tensors["value_head." + k.split(".")[-1]] = v
                if ft_projector and "multi_modal_projector" in k:
                    tensors["multi_modal_projector." + ".".join(k.split(".")[-2:])] = v
        _merge_adapter(tensors, adapter, cfg.lora_scale, _LLAMA_DECODER_LORA)

    def get(name: str) -> torch.Tensor:
        if name not in tensors:
            raise KeyError(f"checkpoint is missing parameter {name!r}")
        return tensors[name]

    return cfg, get


def _canonical_qwen_name(k: str) -> str:
    """Any era of transformers Qwen2.5-VL state_dict name -> the 4.50 names the reference was written against
    (`visual.*`, `model.*` = text decoder)."""
    if k.startswith("model.visual."):
        return k[len("model."):]
    if k.startswith("model.language_model."):
        return "model." + k[len("model.language_model."):]
    return k


def qwen_checkpoint_provider(cfg: QwenVLRewardConfig, pretrain_dir: str, pm_path: Optional[str],
                             ft_projector: bool = False):
    """HF Qwen2.5-VL checkpoint directory + the reference's save_model_lora layout for the qwen branch (key selection
    of reference eval/reward_adaptor_loader.py:80-105, incl. the `merger` remap of :93-103) -> provider with
    synth.qwen_param_specs names."""
    if not os.path.isdir(pretrain_dir):
        raise FileNotFoundError(f"args.pretrain={pretrain_dir!r} is not a local directory (no network access: "
                                "hub ids cannot be resolved) - use 'synthetic' for random-init weights")
    with open(os.path.join(pretrain_dir, "config.json")) as f:
        hf = json.load(f)
    t = hf.get("text_config", hf)      # 4.50-era config.json keeps the text fields at top level
    v = hf.get("vision_config", {})
    cfg.vocab_size = t.get("vocab_size", cfg.vocab_size)
    cfg.hidden_size = t.get("hidden_size", cfg.hidden_size)
    cfg.intermediate_size = t.get("intermediate_size", cfg.intermediate_size)
    cfg.num_layers = t.get("num_hidden_layers", cfg.num_layers)
    cfg.num_heads = t.get("num_attention_heads", cfg.num_heads)
    cfg.num_kv_heads = t.get("num_key_value_heads", cfg.num_kv_heads)
    cfg.rms_eps = t.get("rms_norm_eps", cfg.rms_eps)
    rp = t.get("rope_parameters") or t.get("rope_scaling") or {}
    cfg.rope_theta = rp.get("rope_theta", t.get("rope_theta", cfg.rope_theta))
    cfg.mrope_section = list(rp.get("mrope_section", cfg.mrope_section))
    cfg.image_token_id = hf.get("image_token_id", cfg.image_token_id)
    cfg.vit_depth = v.get("depth", cfg.vit_depth)
    cfg.vit_hidden = v.get("hidden_size", cfg.vit_hidden)
    cfg.vit_intermediate = v.get("intermediate_size", cfg.vit_intermediate)
    cfg.vit_heads = v.get("num_heads", cfg.vit_heads)
    cfg.vit_window = v.get("window_size", cfg.vit_window)
    cfg.vit_fullatt = list(v.get("fullatt_block_indexes", cfg.vit_fullatt))
    tensors: Dict[str, torch.Tensor] = {}
    files = sorted(glob.glob(os.path.join(pretrain_dir, "*.safetensors"))) or \
        sorted(glob.glob(os.path.join(pretrain_dir, "pytorch_model*.bin")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors / pytorch_model*.bin under {pretrain_dir}")
    for fpath in files:
        tensors.update({_canonical_qwen_name(k): val for k, val in _load_file(fpath).items()})
    cfg.use_lora = False
    if pm_path:
        adapter = _load_adapter(cfg, pm_path, _canonical_qwen_name)
        heads = os.path.join(pm_path, "pytorch_model.bin")
        if os.path.exists(heads):
            sd = torch.load(heads, map_location="cpu", weights_only=True)
            for k, val in sd.items():
                leaf = k.split(".")[-1]
                for mod in ("value_head", "W_q", "W_k", "W_v", "ca_layernorm"):
                    if mod in k:
                        tensors[f"{mod}.{leaf}"] = val
                if ft_projector and "merger" in k:
                    # reference :93-103: keys are cut to their last two components ('ln_q.weight', '0.weight', ...)
                    tail = ".".join(k.split(".")[-2:])
                    tensors["visual.merger." + (tail if tail.startswith("ln_q") else "mlp." + tail)] = val
        _merge_adapter(tensors, adapter, cfg.lora_scale, _LLAMA_DECODER_LORA)

    def get(name: str) -> torch.Tensor:
        if name not in tensors:
            raise KeyError(f"checkpoint is missing parameter {name!r}")
        return tensors[name]

    return cfg, get
