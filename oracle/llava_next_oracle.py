"""ORACLE (test infrastructure, not product code) - LLaVA-v1.6 branch.

Plain-torch restatement of the reference's `custom_forward` for model_type == 'llava'
(llava_reward/models/rw_model_general_preference.py:372-375 -> `self.forward(**inputs_batch,
output_hidden_states=True)`, value head + last-valid-token gather :407-448). `self.forward` is the third-party
`transformers` `LlavaNextForConditionalGeneration` (reference pins transformers==4.50.0, requirements.txt:9; installed
here: 5.5.0, file transformers/models/llava_next/modeling_llava_next.py - line numbers below are of the installed
file) on a Llama decoder (transformers/models/llama/modeling_llama.py). The unused lm_head GEMM the reference computes
is not restated.

PARITY PINNING: pinned against the reference executed in the build container
(`tests/golden/make_golden_llava.py` -> tests/golden/llava_*.pt); re-checked by tests/test_llava_oracle_golden.py.
LoRA restates peft 0.13.2 lora.Linear.forward (un-vendored; parity unpinned at the PEFT boundary).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn.functional as F

from .reward_oracle import Params, _rot_half, causal_padding_mask, clip_features, eos_gather, head_tail, lora_linear, rmsnorm

CLIP = "vision_tower.vision_model."
LM = "language_model.model."


def select_best_resolution(original_hw, pinpoints) -> Tuple[int, int]:
    """transformers image_processing_utils.select_best_resolution (called at modeling_llava_next.py:68, 97)."""
    oh, ow = original_hw
    best, best_eff, best_waste = None, 0, float("inf")
    for h, w in pinpoints:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (h, w), eff, waste
    return best


def unpad_image(t: torch.Tensor, original_hw) -> torch.Tensor:
    """modeling_llava_next.py:109-146 on a [C, H, W] feature map."""
    oh, ow = original_hw
    ch, cw = t.shape[1:]
    if ow / oh > cw / ch:
        new_h = int(round(oh * (cw / ow), 7))
        pad = (ch - new_h) // 2
        return t[:, pad: ch - pad, :]
    new_w = int(round(ow * (ch / oh), 7))
    pad = (cw - new_w) // 2
    return t[:, :, pad: cw - pad]


def image_features(P: Params, cfg, pixel_values: torch.Tensor, image_sizes, taps=None) -> List[torch.Tensor]:
    """get_image_features + pack_image_features (modeling_llava_next.py:349-417, 277-343): CLIP hidden_states[-2]
    without CLS -> multi_modal_projector (linear_1, GELU, linear_2 :215-219) -> per image [base 576 rows |
    unpadded grid rows, each grid row followed by image_newline]."""
    sizes = [tuple(int(v) for v in s) for s in (image_sizes.tolist() if torch.is_tensor(image_sizes) else image_sizes)]
    side = cfg.image_size // cfg.patch
    n_patches = []
    for hw in sizes:
        bh, bw = select_best_resolution(hw, cfg.image_grid_pinpoints)
        n_patches.append((bh // cfg.image_size) * (bw // cfg.image_size) + 1)
    pix = torch.cat([pixel_values[b, :n] for b, n in enumerate(n_patches)], dim=0)
    feats = clip_features(P, cfg, pix, taps, prefix=CLIP)  # [sum patches, 576, D]
    h = F.linear(feats, P("multi_modal_projector.linear_1.weight"), P("multi_modal_projector.linear_1.bias"))
    h = F.gelu(h)
    h = F.linear(h, P("multi_modal_projector.linear_2.weight"), P("multi_modal_projector.linear_2.bias"))
    if taps is not None:
        taps["projector_out"] = h
    newline = P("image_newline")
    out = []
    for b, feat in enumerate(torch.split(h, n_patches, dim=0)):
        base, rest = feat[0], feat[1:]
        bh, bw = select_best_resolution(sizes[b], cfg.image_grid_pinpoints)
        gh, gw = bh // cfg.image_size, bw // cfg.image_size
        g = rest.view(gh, gw, side, side, -1).permute(4, 0, 2, 1, 3).contiguous()
        g = g.flatten(1, 2).flatten(2, 3)                      # [H, gh*24, gw*24]
        g = unpad_image(g, sizes[b])
        g = torch.cat((g, newline[:, None, None].expand(*g.shape[:-1], 1)), dim=-1)
        g = g.flatten(1, 2).transpose(0, 1)
        out.append(torch.cat((base, g), dim=0))
    return out


def rope_cos_sin(cfg, position_ids: torch.Tensor, dtype):
    """LlamaRotaryEmbedding.forward (default rope, attention_scaling 1; modeling_llama.py:124-135)."""
    hd = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :].to(position_ids.device)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def decoder_layer(P: Params, cfg, i: int, x: torch.Tensor, mask4d, cos, sin) -> torch.Tensor:
    """LlamaDecoderLayer / LlamaAttention (eager) / LlamaMLP (modeling_llama.py) with LoRA on all seven linears."""
    p = f"{LM}layers.{i}."
    B, S, H = x.shape
    nh, hd = cfg.num_heads, cfg.head_dim
    h = rmsnorm(x, P(p + "input_layernorm.weight"), cfg.rms_eps)
    q = lora_linear(P, cfg, p + "self_attn.q_proj", h).view(B, S, nh, hd).transpose(1, 2)
    k = lora_linear(P, cfg, p + "self_attn.k_proj", h).view(B, S, nh, hd).transpose(1, 2)
    v = lora_linear(P, cfg, p + "self_attn.v_proj", h).view(B, S, nh, hd).transpose(1, 2)
    c, s = cos[:, None], sin[:, None]
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = w + mask4d
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    a = torch.matmul(w, v).transpose(1, 2).reshape(B, S, H)
    x = x + lora_linear(P, cfg, p + "self_attn.o_proj", a)
    h = rmsnorm(x, P(p + "post_attention_layernorm.weight"), cfg.rms_eps)
    act = F.silu(lora_linear(P, cfg, p + "mlp.gate_proj", h)) * lora_linear(P, cfg, p + "mlp.up_proj", h)
    return x + lora_linear(P, cfg, p + "mlp.down_proj", act)


def custom_forward(P: Params, cfg, inputs_batch, taps=None, training: bool = False,
                   mean_hidden_state: bool = False) -> torch.Tensor:
    """-> reward [B, vhd] (GPM) or [B, 1] (BT) for the `inputs_batch` dict of the LlavaNext processor."""
    ids, mask = inputs_batch["input_ids"], inputs_batch["attention_mask"]
    B, S = ids.shape
    x = F.embedding(ids, P(LM + "embed_tokens.weight"))
    feats = torch.cat(image_features(P, cfg, inputs_batch["pixel_values"], inputs_batch["image_sizes"], taps), dim=0)
    sel = ids == cfg.image_token_id
    if int(sel.sum()) != feats.shape[0]:
        raise ValueError(f"Image features and image tokens do not match, tokens: {int(sel.sum())}, "
                         f"features: {feats.shape[0]}")  # modeling_llava_next.py:437-441
    x = x.masked_scatter(sel[..., None].expand_as(x), feats.to(x.dtype))
    if taps is not None:
        taps["image_features"], taps["inputs_embeds"] = feats, x
    # position_ids=None -> LlamaModel uses arange(S) for every row, padded or not
    pos = torch.arange(S, device=ids.device)[None].expand(B, S)
    cos, sin = rope_cos_sin(cfg, pos, P.dtype)
    mask4d = causal_padding_mask(mask, P.dtype)
    for i in range(cfg.num_layers):
        x = decoder_layer(P, cfg, i, x, mask4d, cos, sin)
        if taps is not None:
            taps[f"hidden_{i}"] = x
    x = rmsnorm(x, P(LM + "norm.weight"), cfg.rms_eps)
    if taps is not None:
        taps["last_hidden"] = x
    return head_tail(P, cfg, x, mask, training, mean_hidden_state)
