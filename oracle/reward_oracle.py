"""ORACLE (test infrastructure, not product code).

A plain-torch functional restatement of the reference's reward-scoring forward
(`CustomRewardModel.custom_forward` + `preference_compute`) for the Phi-3.5-vision backbone.
It runs in whatever dtype/device the parameter provider hands it (fp32 on CPU = the
reference's CPU path; bf16 on CUDA = the reference's GPU arithmetic with eager attention),
using the same sequence of torch ops the reference's modules issue, so bf16 rounding points
coincide with the reference's.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module; the product package (`llava-reward_b200/`) never does.

PARITY PINNING: the reference ships no tests or golden vectors for this path,
so this restatement is pinned against outputs of the reference itself, executed in the build
container by `tests/golden/make_golden.py` (imports /root/reference with stub modules for the
absent deepspeed/peft/accelerate/loralib) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` re-checks it on every CPU test run.
The LoRA branch has no runnable upstream implementation here (peft is absent and un-vendored):
it restates peft 0.13.2 `lora.Linear.forward` (base(x) + lora_B(lora_A(x)) * alpha/r) and is
therefore "parity unpinned" at the PEFT boundary.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

CLIP = "model.vision_embed_tokens.img_processor.vision_model."
VE = "model.vision_embed_tokens."


class Params:
    """name -> tensor accessor that casts to the oracle's compute dtype/device."""

    def __init__(self, provider: Callable[[str], torch.Tensor], dtype=torch.float32, device="cpu", cache=True):
        self.provider, self.dtype, self.device, self.cache = provider, dtype, device, cache
        self._c: Dict[str, torch.Tensor] = {}

    def __call__(self, name: str) -> torch.Tensor:
        if name in self._c:
            return self._c[name]
        t = self.provider(name).to(device=self.device, dtype=self.dtype)
        if self.cache:
            self._c[name] = t
        return t


# ----------------------------------------------------------------------------------------
# vision tower: HF CLIPVisionModel as called from modeling_phi3_v.py:208-219
# (hidden_states[-2] of the 24-layer tower == output of encoder layer 23, CLS dropped, no post_layernorm)
# arithmetic: transformers/models/clip/modeling_clip.py (third-party; CLIPVisionEmbeddings.forward,
# CLIPEncoderLayer.forward, eager_attention_forward, CLIPMLP.forward)
# ----------------------------------------------------------------------------------------
def clip_features(P: Params, cfg, pixel_values: torch.Tensor, taps=None, prefix: Optional[str] = None) -> torch.Tensor:
    CLIP = prefix or globals()["CLIP"]
    n = pixel_values.shape[0]
    D, heads, hd = cfg.clip_hidden, cfg.clip_heads, cfg.clip_head_dim
    x = F.conv2d(pixel_values.to(P.dtype), P(CLIP + "embeddings.patch_embedding.weight"), stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)  # [n, 576, D]
    cls = P(CLIP + "embeddings.class_embedding").expand(n, 1, D)
    x = torch.cat([cls, x], dim=1) + P(CLIP + "embeddings.position_embedding.weight")[None]
    x = F.layer_norm(x, (D,), P(CLIP + "pre_layrnorm.weight"), P(CLIP + "pre_layrnorm.bias"), cfg.clip_eps)
    if taps is not None:
        taps["clip_embed"] = x
    T = x.shape[1]
    for i in range(cfg.clip_layers):
        p = f"{CLIP}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), P(p + "layer_norm1.weight"), P(p + "layer_norm1.bias"), cfg.clip_eps)
        q = F.linear(h, P(p + "self_attn.q_proj.weight"), P(p + "self_attn.q_proj.bias"))
        k = F.linear(h, P(p + "self_attn.k_proj.weight"), P(p + "self_attn.k_proj.bias"))
        v = F.linear(h, P(p + "self_attn.v_proj.weight"), P(p + "self_attn.v_proj.bias"))
        q = q.view(n, T, heads, hd).transpose(1, 2)
        k = k.view(n, T, heads, hd).transpose(1, 2)
        v = v.view(n, T, heads, hd).transpose(1, 2)
        w = torch.matmul(q, k.transpose(-1, -2)) * (hd ** -0.5)
        w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
        a = torch.matmul(w, v).transpose(1, 2).reshape(n, T, D)
        x = x + F.linear(a, P(p + "self_attn.out_proj.weight"), P(p + "self_attn.out_proj.bias"))
        h = F.layer_norm(x, (D,), P(p + "layer_norm2.weight"), P(p + "layer_norm2.bias"), cfg.clip_eps)
        h = F.linear(h, P(p + "mlp.fc1.weight"), P(p + "mlp.fc1.bias"))
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu (modeling_phi3_v.py:71)
        x = x + F.linear(h, P(p + "mlp.fc2.weight"), P(p + "mlp.fc2.bias"))
        if taps is not None and i == 0:
            taps["clip_layer0"] = x
    return x[:, 1:]


# ----------------------------------------------------------------------------------------
# HD feature transform: modeling_phi3_v.py:254-362
# ----------------------------------------------------------------------------------------
def hd_row_index(h_crop: int, w_crop: int) -> torch.Tensor:
    """[rows, 4] source-token index into a sample's flattened [17*576] CLIP tokens for the
    'sub_glb' order (modeling_phi3_v.py:259, 283-295): sub image rows with a trailing sub_GN
    per 12-px row (-1), then glb_GN (-2), then the global crop with its own newlines.
    2x2 merge channel order (dy,dx) = (0,0),(0,1),(1,0),(1,1) (modeling_phi3_v.py:315-318)."""
    rows: List[List[int]] = []

    def merged(crop: int, py: int, px: int) -> List[int]:
        return [crop * 576 + (2 * py + dy) * 24 + (2 * px + dx) for dy in (0, 1) for dx in (0, 1)]

    for y in range(h_crop * 12):
        for x in range(w_crop * 12):
            crop = 1 + (y // 12) * w_crop + (x // 12)
            rows.append(merged(crop, y % 12, x % 12))
        rows.append([-1] * 4)
    rows.append([-2] * 4)
    for y in range(12):
        for x in range(12):
            rows.append(merged(0, y, x))
        rows.append([-1] * 4)
    return torch.tensor(rows, dtype=torch.int64)


def hd_feature_rows(P: Params, cfg, feats: torch.Tensor, image_sizes) -> Tuple[torch.Tensor, List[int]]:
    """feats [B, 17, 576, 1024] -> rows [sum N_v, 4096] (input of img_projection), per-sample counts."""
    B = feats.shape[0]
    D = cfg.clip_hidden
    sub_gn = P(VE + "sub_GN").reshape(4 * D)
    glb_gn = P(VE + "glb_GN").reshape(4 * D)
    out, counts = [], []
    for b in range(B):
        h, w = int(image_sizes[b][0]), int(image_sizes[b][1])
        idx = hd_row_index(h // 336, w // 336).to(feats.device)
        flat = feats[b].reshape(-1, D)
        rows = flat[idx.clamp_min(0)].reshape(idx.shape[0], 4 * D)
        rows = torch.where((idx[:, :1] == -1), sub_gn[None], rows)
        rows = torch.where((idx[:, :1] == -2), glb_gn[None], rows)
        out.append(rows)
        counts.append(rows.shape[0])
    return torch.cat(out, dim=0), counts


def img_projection(P: Params, rows: torch.Tensor) -> torch.Tensor:
    """Linear(4096->3072) -> exact GELU -> Linear(3072->3072): modeling_phi3_v.py:172-179, 299-301."""
    h = F.linear(rows, P(VE + "img_projection.0.weight"), P(VE + "img_projection.0.bias"))
    h = F.gelu(h)
    return F.linear(h, P(VE + "img_projection.2.weight"), P(VE + "img_projection.2.bias"))


def embed_tokens(P: Params, cfg, input_ids: torch.Tensor, proj: torch.Tensor):
    """wte gather + scatter of image rows at negative-id positions + zero-padded vision_embeds:
    modeling_phi3_v.py:228-252."""
    is_img = (input_ids < 0) & (input_ids > -int(1e9))
    hidden = F.embedding(input_ids.clamp(0, cfg.vocab_size), P("model.embed_tokens.weight"))
    hidden = hidden.clone()
    hidden[is_img] = proj  # row-major order of positions == order of proj rows
    per_row = is_img.sum(dim=1).tolist()
    chunks = torch.split(proj, per_row)
    mx = max(per_row)
    vis = torch.stack([F.pad(c, (0, 0, 0, mx - c.shape[0])) for c in chunks])
    return hidden, vis


# ----------------------------------------------------------------------------------------
# Phi-3 decoder: modeling_phi3_v.py:377-391 (RMSNorm), 438-476 (su RoPE), 529-553, 556-572 (MLP),
# 588-720 (eager attention), 1130-1205 (layer)
# ----------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def su_rope_cos_sin(cfg, position_ids: torch.Tensor, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    hd = cfg.head_dim
    seq_len = int(position_ids.max()) + 1
    fac = cfg.long_factor if seq_len > cfg.original_max_position_embeddings else cfg.short_factor
    ext = torch.tensor(fac, dtype=torch.float32, device=position_ids.device)
    expo = torch.arange(0, hd, 2, dtype=torch.int64, device=position_ids.device).float() / hd
    inv_freq = 1.0 / (ext * cfg.rope_theta ** expo)
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    s = cfg.rope_scaling_factor
    return (emb.cos() * s).to(dtype), (emb.sin() * s).to(dtype)


def _rot_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def lora_linear(P: Params, cfg, name: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Linear (no bias) wrapped by PEFT LoRA (peft 0.13.2 lora.Linear.forward; un-vendored)."""
    y = F.linear(x, P(name + ".weight"))
    if cfg.use_lora:
        y = y + F.linear(F.linear(x, P(name + ".lora_A.weight")), P(name + ".lora_B.weight")) * cfg.lora_scale
    return y


def causal_padding_mask(attention_mask: torch.Tensor, dtype) -> torch.Tensor:
    """Additive [B,1,S,S] mask: causal AND key-not-padded, dtype-min elsewhere
    (transformers `_prepare_4d_causal_attention_mask`, called at modeling_phi3_v.py:1453-1459)."""
    B, S = attention_mask.shape
    causal = torch.ones(S, S, dtype=torch.bool, device=attention_mask.device).tril()
    ok = causal[None, None] & attention_mask[:, None, None, :].bool()
    m = torch.zeros(B, 1, S, S, dtype=dtype, device=attention_mask.device)
    return m.masked_fill(~ok, torch.finfo(dtype).min)


def decoder_layer(P: Params, cfg, i: int, x: torch.Tensor, mask4d, cos, sin) -> torch.Tensor:
    p = f"model.layers.{i}."
    B, S, H = x.shape
    nh, hd = cfg.num_heads, cfg.head_dim
    h = rmsnorm(x, P(p + "input_layernorm.weight"), cfg.rms_eps)
    qkv = lora_linear(P, cfg, p + "self_attn.qkv_proj", h)
    q = qkv[..., :H].view(B, S, nh, hd).transpose(1, 2)
    k = qkv[..., H:2 * H].view(B, S, nh, hd).transpose(1, 2)
    v = qkv[..., 2 * H:].view(B, S, nh, hd).transpose(1, 2)
    c, s = cos[:, None], sin[:, None]
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd)
    w = w + mask4d
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(v.dtype)
    a = torch.matmul(w, v).transpose(1, 2).reshape(B, S, H)
    x = x + lora_linear(P, cfg, p + "self_attn.o_proj", a)
    h = rmsnorm(x, P(p + "post_attention_layernorm.weight"), cfg.rms_eps)
    gu = lora_linear(P, cfg, p + "mlp.gate_up_proj", h)
    gate, up = gu.chunk(2, dim=-1)
    x = x + lora_linear(P, cfg, p + "mlp.down_proj", up * F.silu(gate))
    return x


# ----------------------------------------------------------------------------------------
# reward head: rw_model_general_preference.py:334-448
# ----------------------------------------------------------------------------------------
def skipca(P: Params, cfg, last_hidden: torch.Tensor, vision_embeds: torch.Tensor) -> torch.Tensor:
    """Single-head cross attention text->image, no mask over zero-padded vision rows, residual,
    RMSNorm (rw_model_general_preference.py:376-386)."""
    q = F.linear(last_hidden, P("W_q.weight"))
    k = F.linear(vision_embeds, P("W_k.weight"))
    v = F.linear(vision_embeds, P("W_v.weight"))
    sc = torch.bmm(q, k.transpose(1, 2)) / math.sqrt(vision_embeds.shape[2])
    w = F.softmax(sc, dim=-1)
    o = torch.bmm(w, v)
    return rmsnorm(last_hidden + o, P("ca_layernorm.weight"), cfg.rms_eps)


def eos_gather(values: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """Index of the last 1 in each mask row; gather that row of `values` [B,S,vhd]
    (rw_model_general_preference.py:420-421, 439-444)."""
    S = attention_mask.shape[1]
    eos = S - 1 - attention_mask.long().flip(1).argmax(dim=1)
    return values[torch.arange(values.shape[0], device=values.device), eos]


def masked_mean(last_hidden: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """mean_hidden_state pooling (rw_model_general_preference.py:398-406): every op in the model dtype."""
    mask = attention_mask.to(dtype=last_hidden.dtype).unsqueeze(-1)
    total = (last_hidden * mask).sum(dim=1)
    lens = mask.sum(dim=1).clamp(min=1e-8)
    return total / lens


def head_tail(P: Params, cfg, x: torch.Tensor, attention_mask: torch.Tensor, training: bool = False,
              mean_hidden_state: bool = False) -> torch.Tensor:
    """value head + the gather rule of rw_model_general_preference.py:398-448 on the final hidden states x [B,S,H]"""
    if mean_hidden_state:
        return F.linear(masked_mean(x, attention_mask), P("value_head.weight"))
    values = F.linear(x, P("value_head.weight"))
    if training:
        # BT: `values.squeeze(-1)[:, -1]` -> [B]; GPM: `values[:, -1, :]` -> [B, vhd] (:413-418, 432-436)
        return values[:, -1] if cfg.is_general_preference else values.squeeze(-1)[:, -1]
    return eos_gather(values, attention_mask)


def custom_forward(P: Params, cfg, input_ids, attention_mask, pixel_values, image_sizes, taps=None, layer_id: int = 32,
                   training: bool = False, mean_hidden_state: bool = False):
    """-> reward [B, vhd] (GPM) or [B, 1] (BT).
    layer_id / training / mean_hidden_state are the attributes the reference's custom_forward reads
    (rw_model_general_preference.py:327-333): `layer_id == 32` -> last_hidden_state (after the final norm), otherwise
    `hidden_states[layer_id]` of (inputs_embeds, h_1 .. h_{L-1}, norm(h_L), vision_embeds) (:349-352,
    modeling_phi3_v.py:1463-1505); training -> the LAST position instead of the last valid one (:413-418, 432-436);
    mean_hidden_state -> masked mean over the sequence before the value head (:398-406)."""
    position_ids = attention_mask.long().cumsum(-1) - 1
    position_ids = position_ids.masked_fill(attention_mask == 0, 1)
    B, C = pixel_values.shape[:2]
    feats = clip_features(P, cfg, pixel_values.flatten(0, 1), taps).reshape(B, C, -1, cfg.clip_hidden)
    rows, _ = hd_feature_rows(P, cfg, feats, image_sizes)
    proj = img_projection(P, rows)
    x, vis = embed_tokens(P, cfg, input_ids, proj)
    if taps is not None:
        taps["clip_features"], taps["img_proj"], taps["inputs_embeds"] = feats, proj, x
    mask4d = causal_padding_mask(attention_mask, P.dtype)
    cos, sin = su_rope_cos_sin(cfg, position_ids, P.dtype)
    if layer_id == 32 or layer_id == cfg.num_layers:
        n_run, final_norm = cfg.num_layers, True
    elif 0 <= layer_id < cfg.num_layers:
        n_run, final_norm = layer_id, False
    else:
        raise ValueError(f"layer_id {layer_id}")
    for i in range(n_run):
        x = decoder_layer(P, cfg, i, x, mask4d, cos, sin)
        if taps is not None:
            taps[f"hidden_{i}"] = x
    if final_norm:
        x = rmsnorm(x, P("model.norm.weight"), cfg.rms_eps)
    if taps is not None:
        taps["last_hidden"] = x
    if cfg.add_cross_attention:
        x = skipca(P, cfg, x, vis)
        if taps is not None:
            taps["skipca_out"] = x
    return head_tail(P, cfg, x, attention_mask, training, mean_hidden_state)


def preference_compute(cfg, chosen: torch.Tensor, reject: torch.Tensor) -> torch.Tensor:
    """eval/reward_adaptor_loader.py:174-181 (returns the fp32 torch tensor, caller does .numpy())."""
    if cfg.is_general_preference and cfg.value_head_dim == 2:
        g = chosen[:, 0] * reject[:, 1] - chosen[:, 1] * reject[:, 0]
        prob = torch.sigmoid(g / cfg.general_preference_tau)
    else:
        prob = torch.sigmoid((chosen - reject) / cfg.general_preference_tau).squeeze(-1)
    return prob.float()
