"""ORACLE (test infrastructure, not product code) - Qwen2.5-VL branch.

Plain-torch restatement of the reference's `custom_forward` for model_type == 'qwen'
(llava_reward/models/rw_model_general_preference.py:354-371 forward + vision rows, :387-397 SkipCA arm, :407-448 value
head + last-valid-token gather). `self.forward` / `self.visual` are the third-party `transformers`
`Qwen2_5_VLForConditionalGeneration` (reference pins transformers==4.50.0, requirements.txt:9; installed here: 5.5.0,
file transformers/models/qwen2_5_vl/modeling_qwen2_5_vl.py - line numbers below are of the installed file). The unused
lm_head GEMM and the redundant extra `self.visual(...)` pass the reference computes (:356) are not restated.

PARITY PINNING: pinned against the reference class executed in the build container on the installed transformers
(`tests/golden/make_golden_qwen.py` -> tests/golden/qwen_*.pt; two harness shims, documented there, make the class
construct under transformers 5.x without touching its arithmetic); re-checked by tests/test_qwen_oracle_golden.py.
LoRA restates peft 0.13.2 lora.Linear.forward (un-vendored; parity unpinned at the PEFT boundary).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .reward_oracle import Params, _rot_half, causal_padding_mask, eos_gather, head_tail, rmsnorm

V = "visual."
LM = "model."


# ----------------------------------------------------------------------------------------
# vision tower: Qwen2_5_VisionTransformerPretrainedModel.forward (modeling_qwen2_5_vl.py:455-523)
# ----------------------------------------------------------------------------------------
def window_plan(cfg, grid_thw):
    """rot_pos_emb (:382-409) + get_window_index (:411-453) + the two cu_seqlens (:470-500), as python lists."""
    m, unit = cfg.vit_merge, cfg.vit_merge ** 2
    wsz = cfg.vit_window // m // cfg.vit_patch
    pos_ids, window_index, cu_win, cu_img, base = [], [], [0], [0], 0
    for t, h, w in [tuple(int(v) for v in g) for g in grid_thw]:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // m, m, w // m, m).permute(0, 2, 1, 3).flatten()
        pos_ids.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
        lh, lw = h // m, w // m
        index = torch.arange(t * lh * lw).reshape(t, lh, lw)
        ph, pw = wsz - lh % wsz, wsz - lw % wsz
        nh, nw = (lh + ph) // wsz, (lw + pw) // wsz
        padded = F.pad(index, (0, pw, 0, ph), "constant", -100)
        padded = padded.reshape(t, nh, wsz, nw, wsz).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, wsz, wsz)
        seqlens = (padded != -100).sum([2, 3]).reshape(-1)
        flat = padded.reshape(-1)
        window_index.append(flat[flat != -100] + base)
        cu_win.extend((seqlens.cumsum(0) * unit + cu_win[-1]).tolist())
        cu_img.extend([cu_img[-1] + h * w * (i + 1) for i in range(t)])
        base += t * lh * lw
    cu_win = torch.unique_consecutive(torch.tensor(cu_win)).tolist()
    return torch.cat(pos_ids, 0), torch.cat(window_index, 0), cu_win, cu_img


def vision_block(P: Params, cfg, i: int, x: torch.Tensor, cu, cos, sin) -> torch.Tensor:
    """Qwen2_5_VLVisionBlock (:290-322): RMSNorm -> qkv(+bias) -> fp32 rotary (apply_rotary_pos_emb_vision :156-167)
    -> per-chunk eager attention -> proj(+bias); RMSNorm -> SwiGLU MLP with biases (:77-88)."""
    p = f"{V}blocks.{i}."
    T, D = x.shape
    nh, hd = cfg.vit_heads, cfg.vit_head_dim
    h = rmsnorm(x, P(p + "norm1.weight"), cfg.vit_eps)
    qkv = F.linear(h, P(p + "attn.qkv.weight"), P(p + "attn.qkv.bias")).reshape(T, 3, nh, hd)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    c, s = cos[:, None, :].float(), sin[:, None, :].float()
    q = (q.float() * c + _rot_half(q.float()) * s).to(x.dtype)
    k = (k.float() * c + _rot_half(k.float()) * s).to(x.dtype)
    outs = []
    for a, b in zip(cu[:-1], cu[1:]):
        qc, kc, vc = (t[a:b].transpose(0, 1) for t in (q, k, v))          # [nh, n, hd]
        w = torch.matmul(qc, kc.transpose(1, 2)) * (hd ** -0.5)
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(qc.dtype)
        outs.append(torch.matmul(w, vc).transpose(0, 1))
    a = torch.cat(outs, 0).reshape(T, D)
    x = x + F.linear(a, P(p + "attn.proj.weight"), P(p + "attn.proj.bias"))
    h = rmsnorm(x, P(p + "norm2.weight"), cfg.vit_eps)
    g = F.silu(F.linear(h, P(p + "mlp.gate_proj.weight"), P(p + "mlp.gate_proj.bias")))
    u = F.linear(h, P(p + "mlp.up_proj.weight"), P(p + "mlp.up_proj.bias"))
    return x + F.linear(g * u, P(p + "mlp.down_proj.weight"), P(p + "mlp.down_proj.bias"))


def vision_tower(P: Params, cfg, pixel_values: torch.Tensor, grid_thw, taps=None) -> torch.Tensor:
    """-> merged image embeddings [sum t*h*w/4, H] in the ORIGINAL (un-windowed) order (`pooler_output`, :515-523)."""
    D, unit = cfg.vit_hidden, cfg.vit_merge ** 2
    x = F.linear(pixel_values.to(P.dtype), P(V + "patch_embed.proj.weight").reshape(D, -1))  # Conv3d, stride = kernel
    pos_ids, window_index, cu_win, cu_img = window_plan(cfg, grid_thw)
    window_index = window_index.to(x.device)
    dim = cfg.vit_head_dim // 2
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
    freqs = torch.outer(torch.arange(int(pos_ids.max()) + 1, dtype=torch.float), inv_freq)
    rot = freqs[pos_ids].flatten(1).to(x.device)                                               # [T, hd/2]
    T = x.shape[0]
    x = x.reshape(T // unit, unit, -1)[window_index].reshape(T, -1)
    rot = rot.reshape(T // unit, unit, -1)[window_index].reshape(T, -1)
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos(), emb.sin()
    if taps is not None:
        taps["vit_embed"] = x
    for i in range(cfg.vit_depth):
        x = vision_block(P, cfg, i, x, cu_img if i in cfg.vit_fullatt else cu_win, cos, sin)
        if taps is not None and i == 0:
            taps["vit_layer0"] = x
    if taps is not None:
        taps["vit_out"] = x
    h = rmsnorm(x, P(V + "merger.ln_q.weight"), 1e-6).view(-1, D * unit)
    h = F.gelu(F.linear(h, P(V + "merger.mlp.0.weight"), P(V + "merger.mlp.0.bias")))
    h = F.linear(h, P(V + "merger.mlp.2.weight"), P(V + "merger.mlp.2.bias"))
    return h[torch.argsort(window_index)]


# ----------------------------------------------------------------------------------------
# M-RoPE: Qwen2_5_VLModel.get_rope_index (:1024-1135) + Qwen2_5_VLRotaryEmbedding.forward (:595-608) +
# apply_multimodal_rotary_pos_emb (:627-669)
# ----------------------------------------------------------------------------------------
def rope_index(cfg, input_ids: torch.Tensor, attention_mask: torch.Tensor, grid_thw) -> torch.Tensor:
    """position_ids [3, B, S] (images only, t = 1): text runs count up from the running position, an image's tokens
    get (start, start + row, start + col) on its merged grid and advance the running position by max(h, w) / merge;
    padded positions stay 0."""
    B, S = input_ids.shape
    pos = torch.zeros(3, B, S, dtype=torch.int64)
    grids = iter([tuple(int(v) for v in g) for g in grid_thw])
    for b in range(B):
        keep = attention_mask[b].bool().cpu()
        ids = input_ids[b].cpu()[keep].tolist()
        out, cur, i = [], 0, 0
        while i < len(ids):
            j = i
            is_img = ids[i] == cfg.image_token_id
            while j < len(ids) and (ids[j] == cfg.image_token_id) == is_img:
                j += 1
            if not is_img:
                out.append(torch.arange(j - i).view(1, -1).expand(3, -1) + cur)
                cur += j - i
            else:
                t, h, w = next(grids)
                lh, lw = h // cfg.vit_merge, w // cfg.vit_merge
                if j - i != t * lh * lw:
                    raise ValueError(f"sample {b}: {j - i} image tokens but image_grid_thw implies {t * lh * lw}")
                pw = torch.arange(cur, cur + lw).repeat(lh * t)
                ph = torch.arange(cur, cur + lh).repeat_interleave(lw * t)
                pt = torch.full((t * lh * lw,), cur, dtype=torch.long)
                out.append(torch.stack([pt, ph, pw], 0))
                cur += max(h, w) // cfg.vit_merge
            i = j
        pos[:, b, keep] = torch.cat(out, 1)
    return pos.to(input_ids.device)


def mrope_cos_sin(cfg, position_ids: torch.Tensor, dtype):
    """cos/sin [B, S, head_dim] with the frequency sections taken from the t / h / w positions."""
    hd = cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
    freqs = position_ids[..., None].float() * inv_freq.to(position_ids.device)      # [3, B, S, hd/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(dtype), emb.sin().to(dtype)
    sec = list(cfg.mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1)
    return cos, sin


def _lora_linear(P: Params, cfg, name: str, x: torch.Tensor, bias: bool) -> torch.Tensor:
    y = F.linear(x, P(name + ".weight"), P(name + ".bias") if bias else None)
    if cfg.use_lora:
        y = y + F.linear(F.linear(x, P(name + ".lora_A.weight")), P(name + ".lora_B.weight")) * cfg.lora_scale
    return y


def decoder_layer(P: Params, cfg, i: int, x: torch.Tensor, mask4d, cos, sin) -> torch.Tensor:
    """Qwen2_5_VLDecoderLayer (:762-828) / Qwen2_5_VLAttention eager (:672-759, GQA via repeat_kv) / Qwen2MLP."""
    p = f"{LM}layers.{i}."
    B, S, H = x.shape
    nh, nkv, hd = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    h = rmsnorm(x, P(p + "input_layernorm.weight"), cfg.rms_eps)
    q = _lora_linear(P, cfg, p + "self_attn.q_proj", h, True).view(B, S, nh, hd).transpose(1, 2)
    k = _lora_linear(P, cfg, p + "self_attn.k_proj", h, True).view(B, S, nkv, hd).transpose(1, 2)
    v = _lora_linear(P, cfg, p + "self_attn.v_proj", h, True).view(B, S, nkv, hd).transpose(1, 2)
    c, s = cos[:, None], sin[:, None]
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    k = k.repeat_interleave(nh // nkv, dim=1)
    v = v.repeat_interleave(nh // nkv, dim=1)
    w = torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5)
    w = w + mask4d
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    a = torch.matmul(w, v).transpose(1, 2).reshape(B, S, H)
    x = x + _lora_linear(P, cfg, p + "self_attn.o_proj", a, False)
    h = rmsnorm(x, P(p + "post_attention_layernorm.weight"), cfg.rms_eps)
    act = F.silu(_lora_linear(P, cfg, p + "mlp.gate_proj", h, False)) * _lora_linear(P, cfg, p + "mlp.up_proj", h, False)
    return x + _lora_linear(P, cfg, p + "mlp.down_proj", act, False)


def skipca_qwen(P: Params, cfg, last_hidden: torch.Tensor, hidden0: torch.Tensor, input_ids: torch.Tensor):
    """rw_model_general_preference.py:358-371, 387-395: the 'vision' rows are hidden_states[0] at the positions whose
    token id is 151643 (the pad / <|endoftext|> id, NOT <|image_pad|>), zero-padded to the batch maximum; padded keys
    are masked with -1e4; single-head attention, residual, ca_layernorm."""
    B, L, H = last_hidden.shape
    image_mask = input_ids == cfg.pad_token_id
    vis_lens = image_mask.sum(dim=1)
    max_v = int(vis_lens.max())
    vision_pad = last_hidden.new_zeros(B, max_v, H)
    pad_mask = torch.ones(B, max_v, dtype=torch.bool, device=last_hidden.device)
    for i in range(B):
        n = int(vis_lens[i])
        vision_pad[i, :n] = hidden0[i, image_mask[i]]
        pad_mask[i, :n] = False
    q = F.linear(last_hidden, P("W_q.weight"))
    k = F.linear(vision_pad, P("W_k.weight"))
    v = F.linear(vision_pad, P("W_v.weight"))
    sc = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(H)
    sc = sc.masked_fill(pad_mask.unsqueeze(1), -1e4)
    o = torch.bmm(F.softmax(sc, dim=-1), v)
    return rmsnorm(last_hidden + o, P("ca_layernorm.weight"), cfg.rms_eps)


def custom_forward(P: Params, cfg, inputs_batch, taps=None, training: bool = False,
                   mean_hidden_state: bool = False) -> torch.Tensor:
    """-> reward [B, vhd] (GPM) or [B, 1] (BT) for the `inputs_batch` dict of the Qwen2.5-VL processor."""
    ids, mask = inputs_batch["input_ids"], inputs_batch["attention_mask"]
    grid = inputs_batch["image_grid_thw"].tolist()
    x = F.embedding(ids, P(LM + "embed_tokens.weight"))
    feats = vision_tower(P, cfg, inputs_batch["pixel_values"].to(x.device), grid, taps)
    sel = ids == cfg.image_token_id
    if int(sel.sum()) != feats.shape[0]:
        raise ValueError(f"Image features and image tokens do not match, tokens: {int(sel.sum())}, "
                         f"features: {feats.shape[0]}")  # get_placeholder_mask :1204-1208
    x = x.masked_scatter(sel[..., None].expand_as(x), feats.to(x.dtype))
    hidden0 = x
    if taps is not None:
        taps["image_embeds"], taps["inputs_embeds"] = feats, x
    pos = rope_index(cfg, ids, mask, grid)
    cos, sin = mrope_cos_sin(cfg, pos, P.dtype)
    mask4d = causal_padding_mask(mask, P.dtype)
    for i in range(cfg.num_layers):
        x = decoder_layer(P, cfg, i, x, mask4d, cos, sin)
        if taps is not None:
            taps[f"hidden_{i}"] = x
    x = rmsnorm(x, P(LM + "norm.weight"), cfg.rms_eps)
    if taps is not None:
        taps["last_hidden"] = x
    if cfg.add_cross_attention:
        x = skipca_qwen(P, cfg, x, hidden0, ids)
        if taps is not None:
            taps["after_skipca"] = x
    return head_tail(P, cfg, x, mask, training, mean_hidden_state)
