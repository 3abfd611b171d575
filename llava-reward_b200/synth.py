"""Deterministic synthetic weights and inputs for the reward-scoring path.

There is no network in the build/bench environment, so weights are random-init of the
named architecture (BASELINE.json `configs`). The generator is a counter-based integer
hash followed by ONE fp32 multiply, so the same (name, seed) gives bit-identical values
from torch on CPU (golden generation, oracle) and on CUDA (tests, bench) without moving
16 GB of fp32 weights around.

Tensor names follow the reference state_dict (``Phi3VForCausalLM`` +
``CustomRewardModel`` heads, reference rw_model_general_preference.py:306-333,
modeling_phi3_v.py:118-205,1332-1368) plus PEFT-style ``lora_A`` / ``lora_B`` entries.
"""
from __future__ import annotations

import zlib
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

from .config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig, anyres_geometry, num_image_tokens

_M32 = 0xFFFFFFFF
_IH_STD = 65536.0 / (3.0 ** 0.5)  # std of the sum of four U{0..65535} (Irwin-Hall, n=4)


def _fmix32(x: torch.Tensor) -> torch.Tensor:
    """murmur3 finaliser on int64 lanes holding 32-bit values (only the low 32 bits matter)."""
    x = x ^ (x >> 16)
    x = (x * 0x85EBCA6B) & _M32
    x = x ^ (x >> 13)
    x = (x * 0xC2B2AE35) & _M32
    x = x ^ (x >> 16)
    return x


def hash_normal(name: str, shape, std: float, seed: int, device="cpu", mean: float = 0.0,
                dtype=torch.float32, chunk: int = 1 << 24) -> torch.Tensor:
    """Pseudo-normal tensor: sum of four 16-bit uniforms, centred, times one fp32 scale."""
    n = 1
    for s in shape:
        n *= int(s)
    key = (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & _M32
    if torch.device(device).type == "cuda" and dtype == torch.float32:
        # the same integer hash as one launch of the library's generator kernel (bit-identical, tested)
        from . import _lib as L
        out = torch.empty(n, dtype=torch.float32, device=device)
        with torch.cuda.device(out.device):
            L.call("lr_synth_normal_f32", out.data_ptr(), n, key, std / _IH_STD, mean, int(mean != 0.0),
                   torch.cuda.current_stream().cuda_stream)
        return out.view(*shape)
    out = torch.empty(n, dtype=dtype, device=device)
    scale = torch.tensor(std / _IH_STD, dtype=torch.float32, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
        a = _fmix32(((idx * 2) * 0x9E3779B1 + key) & _M32)
        b = _fmix32(((idx * 2 + 1) * 0x9E3779B1 + key) & _M32)
        s = (a & 0xFFFF) + (a >> 16) + (b & 0xFFFF) + (b >> 16) - 131070
        v = s.to(torch.float32) * scale
        if mean != 0.0:
            v = v + mean
        out[lo:hi] = v.to(dtype)
    return out.view(*shape)


def hash_randint(name: str, n: int, lo: int, hi: int, seed: int, device="cpu") -> torch.Tensor:
    """Deterministic integers in [lo, hi)."""
    key = (zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & _M32
    idx = torch.arange(n, dtype=torch.int64, device=device)
    h = _fmix32((idx * 0x9E3779B1 + key) & _M32)
    return lo + h % (hi - lo)


# --------------------------------------------------------------------------------------
# parameter inventory
# --------------------------------------------------------------------------------------
CLIP_PREFIX = "model.vision_embed_tokens.img_processor.vision_model."


def param_specs(cfg: RewardConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (name, shape, kind) for every parameter the scoring path reads.

    kind: 'w' = N(0, std), 'n' = norm weight 1 + N(0, std), all biases are 'w' as well so
    that a dropped bias / separator / LoRA branch is visible to the parity tests
    (default init leaves those at zero).
    """
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    D, DI = cfg.clip_hidden, cfg.clip_intermediate
    yield "model.embed_tokens.weight", (V, H), "w"
    ve = "model.vision_embed_tokens."
    yield ve + "glb_GN", (1, 1, 4 * D), "w"
    yield ve + "sub_GN", (1, 1, 1, 4 * D), "w"
    yield ve + "img_projection.0.weight", (H, 4 * D), "w"
    yield ve + "img_projection.0.bias", (H,), "w"
    yield ve + "img_projection.2.weight", (H, H), "w"
    yield ve + "img_projection.2.bias", (H,), "w"
    c = CLIP_PREFIX
    yield c + "embeddings.class_embedding", (D,), "w"
    yield c + "embeddings.patch_embedding.weight", (D, 3, cfg.patch, cfg.patch), "w"
    yield c + "embeddings.position_embedding.weight", (cfg.clip_tokens, D), "w"
    yield c + "pre_layrnorm.weight", (D,), "n"
    yield c + "pre_layrnorm.bias", (D,), "w"
    for i in range(cfg.clip_layers):
        p = f"{c}encoder.layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{proj}.weight", (D, D), "w"
            yield p + f"self_attn.{proj}.bias", (D,), "w"
        yield p + "layer_norm1.weight", (D,), "n"
        yield p + "layer_norm1.bias", (D,), "w"
        yield p + "mlp.fc1.weight", (DI, D), "w"
        yield p + "mlp.fc1.bias", (DI,), "w"
        yield p + "mlp.fc2.weight", (D, DI), "w"
        yield p + "mlp.fc2.bias", (D,), "w"
        yield p + "layer_norm2.weight", (D,), "n"
        yield p + "layer_norm2.bias", (D,), "w"
    r = cfg.lora_rank
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        lin = (("self_attn.qkv_proj", 3 * H, H), ("self_attn.o_proj", H, H),
               ("mlp.gate_up_proj", 2 * I, H), ("mlp.down_proj", H, I))
        for nm, o, k in lin:
            yield p + nm + ".weight", (o, k), "w"
            if cfg.use_lora:
                yield p + nm + ".lora_A.weight", (r, k), "w"
                yield p + nm + ".lora_B.weight", (o, r), "w"
        yield p + "input_layernorm.weight", (H,), "n"
        yield p + "post_attention_layernorm.weight", (H,), "n"
    yield "model.norm.weight", (H,), "n"
    yield "value_head.weight", (cfg.vhd, H), "w"
    if cfg.add_cross_attention:
        yield "W_q.weight", (H, H), "w"
        yield "W_k.weight", (H, H), "w"
        yield "W_v.weight", (H, H), "w"
        yield "ca_layernorm.weight", (H,), "n"


LLAVA_CLIP_PREFIX = "vision_tower.vision_model."
LLAVA_LM_PREFIX = "language_model.model."


def llava_param_specs(cfg: LlavaNextRewardConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Parameters of the LLaVA-v1.6 reward model under the state_dict names of the transformers release the
    reference pins (4.50: `vision_tower.*`, `multi_modal_projector.*`, `image_newline`, `language_model.model.*` -
    the names create_lora_config_llava16_vicuna targets, llava_reward/utils/utils.py:243-251) + `value_head`."""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    D, DI = cfg.clip_hidden, cfg.clip_intermediate
    lm, c = LLAVA_LM_PREFIX, LLAVA_CLIP_PREFIX
    yield lm + "embed_tokens.weight", (V, H), "w"
    yield "image_newline", (H,), "w"
    yield "multi_modal_projector.linear_1.weight", (H, D), "w"
    yield "multi_modal_projector.linear_1.bias", (H,), "w"
    yield "multi_modal_projector.linear_2.weight", (H, H), "w"
    yield "multi_modal_projector.linear_2.bias", (H,), "w"
    yield c + "embeddings.class_embedding", (D,), "w"
    yield c + "embeddings.patch_embedding.weight", (D, 3, cfg.patch, cfg.patch), "w"
    yield c + "embeddings.position_embedding.weight", (cfg.clip_tokens, D), "w"
    yield c + "pre_layrnorm.weight", (D,), "n"
    yield c + "pre_layrnorm.bias", (D,), "w"
    for i in range(cfg.clip_layers):
        p = f"{c}encoder.layers.{i}."
        for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
            yield p + f"self_attn.{proj}.weight", (D, D), "w"
            yield p + f"self_attn.{proj}.bias", (D,), "w"
        yield p + "layer_norm1.weight", (D,), "n"
        yield p + "layer_norm1.bias", (D,), "w"
        yield p + "mlp.fc1.weight", (DI, D), "w"
        yield p + "mlp.fc1.bias", (DI,), "w"
        yield p + "mlp.fc2.weight", (D, DI), "w"
        yield p + "mlp.fc2.bias", (D,), "w"
        yield p + "layer_norm2.weight", (D,), "n"
        yield p + "layer_norm2.bias", (D,), "w"
    r = cfg.lora_rank
    for i in range(cfg.num_layers):
        p = f"{lm}layers.{i}."
        lin = [(f"self_attn.{n}_proj", H, H) for n in "qkvo"] + [("mlp.gate_proj", I, H), ("mlp.up_proj", I, H),
                                                                ("mlp.down_proj", H, I)]
        for nm, o, k in lin:
            yield p + nm + ".weight", (o, k), "w"
            if cfg.use_lora:
                yield p + nm + ".lora_A.weight", (r, k), "w"
                yield p + nm + ".lora_B.weight", (o, r), "w"
        yield p + "input_layernorm.weight", (H,), "n"
        yield p + "post_attention_layernorm.weight", (H,), "n"
    yield lm + "norm.weight", (H,), "n"
    yield "value_head.weight", (cfg.vhd, H), "w"


def qwen_param_specs(cfg: QwenVLRewardConfig) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Parameters of the Qwen2.5-VL reward model under the state_dict names of the transformers release the reference
    pins (4.50: `visual.*`, `model.*` = the text decoder - the names create_lora_config_qwen targets,
    llava_reward/utils/utils.py:223-231) + the reward heads (rw_model_general_preference.py:314-326)."""
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    D, DI = cfg.vit_hidden, cfg.vit_intermediate
    kv = cfg.num_kv_heads * cfg.head_dim
    yield "model.embed_tokens.weight", (V, H), "w"
    yield "visual.patch_embed.proj.weight", (D, 3, cfg.vit_temporal_patch, cfg.vit_patch, cfg.vit_patch), "w"
    for i in range(cfg.vit_depth):
        p = f"visual.blocks.{i}."
        yield p + "norm1.weight", (D,), "n"
        yield p + "norm2.weight", (D,), "n"
        yield p + "attn.qkv.weight", (3 * D, D), "w"
        yield p + "attn.qkv.bias", (3 * D,), "w"
        yield p + "attn.proj.weight", (D, D), "w"
        yield p + "attn.proj.bias", (D,), "w"
        for nm, o, k in (("gate_proj", DI, D), ("up_proj", DI, D), ("down_proj", D, DI)):
            yield p + f"mlp.{nm}.weight", (o, k), "w"
            yield p + f"mlp.{nm}.bias", (o,), "w"
    M4 = D * cfg.vit_merge ** 2
    yield "visual.merger.ln_q.weight", (D,), "n"
    yield "visual.merger.mlp.0.weight", (M4, M4), "w"
    yield "visual.merger.mlp.0.bias", (M4,), "w"
    yield "visual.merger.mlp.2.weight", (H, M4), "w"
    yield "visual.merger.mlp.2.bias", (H,), "w"
    r = cfg.lora_rank
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        lin = [("self_attn.q_proj", H, H, True), ("self_attn.k_proj", kv, H, True), ("self_attn.v_proj", kv, H, True),
               ("self_attn.o_proj", H, H, False), ("mlp.gate_proj", I, H, False), ("mlp.up_proj", I, H, False),
               ("mlp.down_proj", H, I, False)]
        for nm, o, k, bias in lin:
            yield p + nm + ".weight", (o, k), "w"
            if bias:
                yield p + nm + ".bias", (o,), "w"
            if cfg.use_lora:
                yield p + nm + ".lora_A.weight", (r, k), "w"
                yield p + nm + ".lora_B.weight", (o, r), "w"
        yield p + "input_layernorm.weight", (H,), "n"
        yield p + "post_attention_layernorm.weight", (H,), "n"
    yield "model.norm.weight", (H,), "n"
    yield "value_head.weight", (cfg.vhd, H), "w"
    if cfg.add_cross_attention:
        yield "W_q.weight", (H, H), "w"
        yield "W_k.weight", (H, H), "w"
        yield "W_v.weight", (H, H), "w"
        yield "ca_layernorm.weight", (H,), "n"


class SynthProvider:
    """Callable ``name -> tensor`` producing synthetic parameters on demand."""

    def __init__(self, cfg: RewardConfig, seed: int = 1234, std: float = 0.02, device="cpu",
                 dtype=torch.float32):
        self.cfg, self.seed, self.std, self.device, self.dtype = cfg, seed, std, device, dtype
        gen = (llava_param_specs if isinstance(cfg, LlavaNextRewardConfig) else
               qwen_param_specs if isinstance(cfg, QwenVLRewardConfig) else param_specs)
        self.specs: Dict[str, Tuple[Tuple[int, ...], str]] = {n: (s, k) for n, s, k in gen(cfg)}

    def names(self) -> List[str]:
        return list(self.specs)

    def __contains__(self, name: str) -> bool:
        return name in self.specs

    def __call__(self, name: str) -> torch.Tensor:
        shape, kind = self.specs[name]
        mean = 1.0 if kind == "n" else 0.0
        return hash_normal(name, shape, self.std, self.seed, device=self.device, mean=mean, dtype=self.dtype)


# --------------------------------------------------------------------------------------
# synthetic inputs (the batch layout of reference reward_dataset.py:137-180 after squeeze(1))
# --------------------------------------------------------------------------------------
BOS, USER, NL, EOS, PAD = 1, 32010, 13, 32000, 32000


def synth_batch(cfg: RewardConfig, batch: int, image_hw: Tuple[int, int], seq_len: Optional[int],
                seed: int = 7, device="cpu", text_len_range: Tuple[int, int] = (40, 128),
                tag: str = "c", image_hw_list=None):
    """Build (input_ids, attention_mask, pixel_values, image_sizes) for `batch` samples.

    Every sample is  [pad..., BOS, <|user|>, \\n, -1 x N_v, \\n, text..., EOS]  left-padded to
    `seq_len` (None -> longest sample). Pixel values are N(0,1) in the real crop slots and zero
    in the padded slots (what ``pad_to_max_num_crops_tensor`` produces,
    reference processing_phi3_v.py:128-136).
    """
    ids_rows, lens = [], []
    C = cfg.num_crops + 1
    pix = torch.zeros(batch, C, 3, cfg.image_size, cfg.image_size, dtype=torch.float32, device=device)
    sizes = torch.zeros(batch, 2, dtype=torch.int64)
    tl = hash_randint(f"textlen.{tag}", batch, text_len_range[0], text_len_range[1], seed).tolist()
    for b in range(batch):
        h, w = image_hw_list[b] if image_hw_list is not None else image_hw
        sizes[b, 0], sizes[b, 1] = h, w
        nv = num_image_tokens(h, w)
        ncrop = (h // 336) * (w // 336) + 1
        pix[b, :ncrop] = hash_normal(f"pixels.{tag}.{b}", (ncrop, 3, cfg.image_size, cfg.image_size), 1.0, seed,
                                     device=device)
        text = hash_randint(f"text.{tag}.{b}", tl[b], 3, 31999, seed).tolist()
        row = [BOS, USER, NL] + [-1] * nv + [NL] + text + [EOS]
        ids_rows.append(row)
        lens.append(len(row))
    S = max(lens) if seq_len is None else seq_len
    ids = torch.full((batch, S), PAD, dtype=torch.int64)
    mask = torch.zeros((batch, S), dtype=torch.int64)
    for b, row in enumerate(ids_rows):
        if len(row) > S:
            raise ValueError(f"sample {b} needs {len(row)} tokens > seq_len {S}")
        ids[b, S - len(row):] = torch.tensor(row, dtype=torch.int64)
        mask[b, S - len(row):] = 1
    return ids.to(device), mask.to(device), pix, sizes.to(device)


def synth_batch_llava(cfg: LlavaNextRewardConfig, batch: int, orig_hw_list, seq_len: Optional[int], seed: int = 7,
                      device="cpu", text_len_range: Tuple[int, int] = (40, 128), tag: str = "c",
                      padding_side: str = "left"):
    """The `inputs_batch` LlavaNextProcessor(images, text, padding=True) hands to custom_forward
    (reference reward_dataset.py:334-346): input_ids [B,S] with `image_token_id` repeated N_v times, text FIRST then
    the image (:267-277), attention_mask, pixel_values [B, max_patches, 3, 336, 336] zero-padded, image_sizes [B,2]
    = ORIGINAL (h, w). Rows are [pad..., BOS, text..., IMG x N_v, EOS]."""
    geos = [anyres_geometry(hw, cfg.image_grid_pinpoints, cfg.image_size, cfg.patch) for hw in orig_hw_list]
    P = max(g["n_patches"] for g in geos)
    pix = torch.zeros(batch, P, 3, cfg.image_size, cfg.image_size, dtype=torch.float32, device=device)
    sizes = torch.tensor([list(hw) for hw in orig_hw_list], dtype=torch.int64)
    tl = hash_randint(f"textlen.{tag}", batch, text_len_range[0], text_len_range[1], seed).tolist()
    rows = []
    for b in range(batch):
        g = geos[b]
        pix[b, : g["n_patches"]] = hash_normal(f"pixels.{tag}.{b}", (g["n_patches"], 3, cfg.image_size, cfg.image_size),
                                               1.0, seed, device=device)
        text = hash_randint(f"text.{tag}.{b}", tl[b], 3, 31999, seed).tolist()
        rows.append([BOS] + text + [cfg.image_token_id] * g["n_tokens"] + [2])
    S = max(len(r) for r in rows) if seq_len is None else seq_len
    ids = torch.full((batch, S), 0, dtype=torch.int64)
    mask = torch.zeros((batch, S), dtype=torch.int64)
    for b, row in enumerate(rows):
        if len(row) > S:
            raise ValueError(f"sample {b} needs {len(row)} tokens > seq_len {S}")
        if padding_side == "left":
            ids[b, S - len(row):] = torch.tensor(row, dtype=torch.int64)
            mask[b, S - len(row):] = 1
        else:
            ids[b, : len(row)] = torch.tensor(row, dtype=torch.int64)
            mask[b, : len(row)] = 1
    return {"input_ids": ids.to(device), "attention_mask": mask.to(device), "pixel_values": pix,
            "image_sizes": sizes.to(device)}


def synth_batch_qwen(cfg: QwenVLRewardConfig, grid_hw_list, seq_len: Optional[int], seed: int = 7, device="cpu",
                     text_len_range: Tuple[int, int] = (40, 128), tag: str = "c", padding_side: str = "left"):
    """The `inputs_batch` Qwen2_5_VLProcessor(text, images, padding=True) hands to custom_forward
    (reference reward_dataset.py:466-489): input_ids [B,S] = [pad..., <|im_start|>, user, \n, <|vision_start|>,
    <|image_pad|> x (h/2 * w/2), <|vision_end|>, text..., <|im_end|>] (the chat template cut at :417), attention_mask,
    pixel_values [sum h*w, 1176] fp32 (flattened (c, t, ph, pw) patches in 2x2-merge order), image_grid_thw [B,3],
    mm_token_type_ids (1 at image positions; what the installed processor adds). grid_hw_list: (h, w) in patches."""
    B = len(grid_hw_list)
    tl = hash_randint(f"textlen.{tag}", B, text_len_range[0], text_len_range[1], seed).tolist()
    rows, pix = [], []
    for b, (h, w) in enumerate(grid_hw_list):
        assert h % cfg.vit_merge == 0 and w % cfg.vit_merge == 0
        pix.append(hash_normal(f"pixels.{tag}.{b}", (h * w, cfg.patch_dim), 1.0, seed, device=device))
        text = hash_randint(f"text.{tag}.{b}", tl[b], 3, 151000, seed).tolist()
        n_img = (h // cfg.vit_merge) * (w // cfg.vit_merge)
        rows.append([151644, 872, 198, cfg.vision_start_token_id] + [cfg.image_token_id] * n_img +
                    [cfg.vision_end_token_id] + text + [151645])
    S = max(len(r) for r in rows) if seq_len is None else seq_len
    ids = torch.full((B, S), cfg.pad_token_id, dtype=torch.int64)
    mask = torch.zeros((B, S), dtype=torch.int64)
    for b, row in enumerate(rows):
        if len(row) > S:
            raise ValueError(f"sample {b} needs {len(row)} tokens > seq_len {S}")
        if padding_side == "left":
            ids[b, S - len(row):] = torch.tensor(row, dtype=torch.int64)
            mask[b, S - len(row):] = 1
        else:
            ids[b, : len(row)] = torch.tensor(row, dtype=torch.int64)
            mask[b, : len(row)] = 1
    grid = torch.tensor([[1, h, w] for h, w in grid_hw_list], dtype=torch.int64)
    return {"input_ids": ids.to(device), "attention_mask": mask.to(device), "pixel_values": torch.cat(pix, 0),
            "image_grid_thw": grid.to(device), "mm_token_type_ids": (ids == cfg.image_token_id).int().to(device)}
