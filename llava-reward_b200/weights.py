"""Pack reference-named parameters into the device layouts the kernels consume.

Input: a provider ``name -> tensor`` using the reference's state_dict names
(``Phi3VForCausalLM`` + reward heads + PEFT ``lora_A/lora_B``; see synth.param_specs) - a real
checkpoint loaded by `load_reward_adaptor` yields the same names.
Output: bf16 CUDA tensors
  * CLIP q/k/v fused into one [3072,1024] weight (the reference issues 3 GEMMs, modeling_phi3_v.py:101-103)
  * patch-embedding conv weight flattened to [1024, 640] (588 + zero pad so K is a multiple of 64)
  * LoRA folded by K-extension: W_ext = [W | (alpha/r) B]  ([out, in + r]); lora_A kept as its own [r, in] GEMM
  * gate_up_proj rows interleaved in blocks of 128 ([gate_b | up_b]) so silu(gate)*up is a GEMM epilogue
  * W_k and W_v stacked into one [2H, H] weight
  * qkv_proj rows of the q and k thirds re-ordered inside every head so that the two halves are interleaved
    (new row 2i = old i, 2i+1 = old i + head_dim/2): a RoPE rotation pair becomes two adjacent output columns and
    the rotation runs in the GEMM epilogue (lr_gemm_rope_bf16). q.k^T is invariant under this common permutation.
"""
from __future__ import annotations

from typing import Callable, Dict, List

import torch

from .config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig
from .synth import CLIP_PREFIX, LLAVA_CLIP_PREFIX, LLAVA_LM_PREFIX

VE = "model.vision_embed_tokens."


class PackedWeights:
    def __init__(self):
        self.clip: Dict[str, torch.Tensor] = {}
        self.clip_layers: List[Dict[str, torch.Tensor]] = []
        self.proj: Dict[str, torch.Tensor] = {}
        self.layers: List[Dict[str, torch.Tensor]] = []
        self.head: Dict[str, torch.Tensor] = {}
        self.embed: torch.Tensor = None
        self.vit: Dict[str, torch.Tensor] = {}                 # Qwen2.5-VL vision tower
        self.vit_layers: List[Dict[str, torch.Tensor]] = []

    def nbytes(self) -> int:
        tot = 0
        for d in [self.clip, self.proj, self.head, self.vit, *self.clip_layers, *self.layers, *self.vit_layers]:
            tot += sum(t.numel() * t.element_size() for t in d.values())
        return tot + self.embed.numel() * self.embed.element_size()


def _pack_clip(pw: PackedWeights, cfg, g, c: str, device, bf=torch.bfloat16) -> None:
    """CLIP ViT-L/14-336 tower (shared by the Phi-3-V and LLaVA-v1.6 branches; `c` = state_dict prefix).
    `bf` = storage dtype (bf16; fp32 for the verification path)."""
    D = cfg.clip_hidden
    pe = g(c + "embeddings.patch_embedding.weight").reshape(D, -1)  # [1024, 588], (ch, ky, kx) order
    patch_w = torch.zeros(D, 640, dtype=bf, device=device)
    patch_w[:, : pe.shape[1]] = pe
    pw.clip = {
        "patch_w": patch_w,
        "cls": g(c + "embeddings.class_embedding"),
        "pos": g(c + "embeddings.position_embedding.weight"),
        "pre_w": g(c + "pre_layrnorm.weight"),
        "pre_b": g(c + "pre_layrnorm.bias"),
    }
    for i in range(cfg.clip_layers):
        p = f"{c}encoder.layers.{i}."
        a = p + "self_attn."
        pw.clip_layers.append({
            "ln1_w": g(p + "layer_norm1.weight"), "ln1_b": g(p + "layer_norm1.bias"),
            "qkv_w": torch.cat([g(a + "q_proj.weight"), g(a + "k_proj.weight"), g(a + "v_proj.weight")], 0).contiguous(),
            "qkv_b": torch.cat([g(a + "q_proj.bias"), g(a + "k_proj.bias"), g(a + "v_proj.bias")], 0).contiguous(),
            "out_w": g(a + "out_proj.weight"), "out_b": g(a + "out_proj.bias"),
            "ln2_w": g(p + "layer_norm2.weight"), "ln2_b": g(p + "layer_norm2.bias"),
            "fc1_w": g(p + "mlp.fc1.weight"), "fc1_b": g(p + "mlp.fc1.bias"),
            "fc2_w": g(p + "mlp.fc2.weight"), "fc2_b": g(p + "mlp.fc2.bias"),
        })


def _pad128(n: int) -> int:
    return (n + 127) // 128 * 128


def _pad_rank(A: torch.Tensor, B: torch.Tensor):
    """Zero-pad a LoRA pair (A [r, in], B [out, r]) to a rank that is a multiple of 128: the `x A^T` GEMM needs
    N % 128 == 0 and the K-extension K % 64 == 0 (lr_gemm_bf16); zero rows of A times zero columns of B add exactly 0,
    so adapters of any rank (peft r = 8, 16, 64, ...) run unchanged."""
    r = A.shape[0]
    rp = _pad128(r)
    if rp == r:
        return A, B
    A = torch.cat([A, torch.zeros(rp - r, A.shape[1], dtype=A.dtype, device=A.device)], 0).contiguous()
    B = torch.cat([B, torch.zeros(B.shape[0], rp - r, dtype=B.dtype, device=B.device)], 1).contiguous()
    return A, B


def _qk_interleave_perm(n_heads: int, head_dim: int, device) -> torch.Tensor:
    """Row permutation of a [n_heads*head_dim, *] q or k projection: inside every head new row 2i = old i,
    2i+1 = old i + head_dim/2 (RoPE pairs adjacent, see lr_gemm_rope_bf16)."""
    inter = torch.stack([torch.arange(head_dim // 2), torch.arange(head_dim // 2) + head_dim // 2], dim=1).reshape(-1)
    return (torch.arange(n_heads)[:, None] * head_dim + inter[None, :]).reshape(-1).to(device)


def pack_weights(cfg: RewardConfig, get: Callable[[str], torch.Tensor], device="cuda",
                 dtype=torch.bfloat16) -> PackedWeights:
    """`dtype` = torch.float32 packs the SAME layouts in fp32 for the verification path (RewardEngine precision="fp32")."""
    bf = dtype

    def g(name):
        return get(name).to(device=device, dtype=bf).contiguous()

    pw = PackedWeights()
    _pack_clip(pw, cfg, g, CLIP_PREFIX, device, bf)
    pw.proj = {
        "p0_w": g(VE + "img_projection.0.weight"), "p0_b": g(VE + "img_projection.0.bias"),
        "p2_w": g(VE + "img_projection.2.weight"), "p2_b": g(VE + "img_projection.2.bias"),
        "sub_gn": g(VE + "sub_GN").reshape(-1), "glb_gn": g(VE + "glb_GN").reshape(-1),
    }
    pw.embed = g("model.embed_tokens.weight")
    H, I, r = cfg.hidden_size, cfg.intermediate_size, cfg.lora_rank
    assert I % 128 == 0
    nb = I // 128
    perm = torch.arange(2 * I, device=device).view(2, nb, 128).permute(1, 0, 2).reshape(-1)  # [gate_b | up_b] blocks

    def ext(name):
        W = g(name + ".weight")
        if not cfg.use_lora:
            return W, None
        A = g(name + ".lora_A.weight")
        B = (get(name + ".lora_B.weight").to(device=device, dtype=torch.float32) * cfg.lora_scale).to(bf)
        A, B = _pad_rank(A, B)
        return torch.cat([W, B], dim=1).contiguous(), A

    qkv_perm = torch.cat([_qk_interleave_perm(2 * cfg.num_heads, cfg.head_dim, device),  # q and k heads: 0,48,1,49,..
                          torch.arange(2 * H, 3 * H, device=device)])
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        qkv_w, qkv_a = ext(p + "self_attn.qkv_proj")
        qkv_w = qkv_w[qkv_perm].contiguous()
        o_w, o_a = ext(p + "self_attn.o_proj")
        gu_w, gu_a = ext(p + "mlp.gate_up_proj")
        gu_w = gu_w[perm].contiguous()
        dn_w, dn_a = ext(p + "mlp.down_proj")
        d = {"in_ln": g(p + "input_layernorm.weight"), "post_ln": g(p + "post_attention_layernorm.weight"),
             "qkv_w": qkv_w, "o_w": o_w, "gu_w": gu_w, "dn_w": dn_w}
        if cfg.use_lora:
            d.update({"qkv_a": qkv_a, "o_a": o_a, "gu_a": gu_a, "dn_a": dn_a})
        pw.layers.append(d)
    pw.head = {"norm": g("model.norm.weight"), "vh": g("value_head.weight")}
    if cfg.add_cross_attention:
        pw.head.update({
            "wq": g("W_q.weight"),
            "wkv": torch.cat([g("W_k.weight"), g("W_v.weight")], 0).contiguous(),
            "ca_ln": g("ca_layernorm.weight"),
        })
    return pw


def pack_weights_llava(cfg: LlavaNextRewardConfig, get: Callable[[str], torch.Tensor], device="cuda") -> PackedWeights:
    """LLaVA-v1.6 (reference names of transformers 4.50, see synth.llava_param_specs) -> the SAME per-layer dict the
    Phi-3 decoder loop consumes, so one engine code path serves both backbones:
      * q/k/v stacked into one [3H, H + 3r] weight: columns [W | 2B_q 0 0 ; 0 2B_k 0 ; 0 0 2B_v] against the stacked
        A = [A_q; A_k; A_v] ([3r, H]) - three LoRA branches in one K-extension
      * gate/up stacked the same way ([2I, H + 2r]), then interleaved in blocks of 128 rows for the SwiGLU epilogue
      * q/k rows head-interleaved for the fused RoPE epilogue (head_dim 128)."""
    bf = torch.bfloat16

    def g(name):
        return get(name).to(device=device, dtype=bf).contiguous()

    pw = PackedWeights()
    _pack_clip(pw, cfg, g, LLAVA_CLIP_PREFIX, device)
    pw.proj = {
        "p0_w": g("multi_modal_projector.linear_1.weight"), "p0_b": g("multi_modal_projector.linear_1.bias"),
        "p2_w": g("multi_modal_projector.linear_2.weight"), "p2_b": g("multi_modal_projector.linear_2.bias"),
        "newline": g("image_newline").reshape(-1),
    }
    lm = LLAVA_LM_PREFIX
    pw.embed = g(lm + "embed_tokens.weight")
    H, I, r = cfg.hidden_size, cfg.intermediate_size, cfg.lora_rank
    assert I % 128 == 0 and H % 256 == 0
    nb = I // 128
    gu_perm = torch.arange(2 * I, device=device).view(2, nb, 128).permute(1, 0, 2).reshape(-1)
    qk = _qk_interleave_perm(cfg.num_heads, cfg.head_dim, device)
    qkv_perm = torch.cat([qk, H + qk, torch.arange(2 * H, 3 * H, device=device)])

    def stacked(names):
        """rows = concatenation of the named linears; LoRA B blocks on the block diagonal of the K-extension"""
        Ws = [g(n + ".weight") for n in names]
        if not cfg.use_lora:
            return torch.cat(Ws, 0).contiguous(), None
        k = len(names)
        As = [g(n + ".lora_A.weight") for n in names]
        r = As[0].shape[0]                      # the adapter's own rank (any value: the stack is zero-padded to 128)
        kr = _pad128(k * r)
        blocks = []
        for j, n in enumerate(names):
            B = (get(n + ".lora_B.weight").to(device=device, dtype=torch.float32) * cfg.lora_scale).to(bf)
            ext = torch.zeros(B.shape[0], kr, dtype=bf, device=device)
            ext[:, j * r:(j + 1) * r] = B
            blocks.append(torch.cat([Ws[j], ext], dim=1))
        A = torch.cat(As, 0)
        if kr != k * r:
            A = torch.cat([A, torch.zeros(kr - k * r, A.shape[1], dtype=bf, device=device)], 0)
        return torch.cat(blocks, 0).contiguous(), A.contiguous()

    for i in range(cfg.num_layers):
        p = f"{lm}layers.{i}."
        qkv_w, qkv_a = stacked([p + f"self_attn.{n}_proj" for n in "qkv"])
        qkv_w = qkv_w[qkv_perm].contiguous()
        o_w, o_a = stacked([p + "self_attn.o_proj"])
        gu_w, gu_a = stacked([p + "mlp.gate_proj", p + "mlp.up_proj"])
        gu_w = gu_w[gu_perm].contiguous()
        dn_w, dn_a = stacked([p + "mlp.down_proj"])
        d = {"in_ln": g(p + "input_layernorm.weight"), "post_ln": g(p + "post_attention_layernorm.weight"),
             "qkv_w": qkv_w, "o_w": o_w, "gu_w": gu_w, "dn_w": dn_w}
        if cfg.use_lora:
            d.update({"qkv_a": qkv_a, "o_a": o_a, "gu_a": gu_a, "dn_a": dn_a})
        pw.layers.append(d)
    pw.head = {"norm": g(lm + "norm.weight"), "vh": g("value_head.weight")}
    return pw


def qwen_vit_padded_head_dim(head_dim: int) -> int:
    """head_dim of the vision tower as the attention kernel sees it: 80 is zero-padded to 96 (three 32-column TMA atoms)."""
    assert head_dim % 2 == 0 and head_dim <= 96
    return 64 if head_dim <= 64 else 96


def pack_weights_qwen(cfg: QwenVLRewardConfig, get: Callable[[str], torch.Tensor], device="cuda") -> PackedWeights:
    """Qwen2.5-VL (reference-era names, see synth.qwen_param_specs) -> kernel layouts.
    Vision tower (transformers modeling_qwen2_5_vl.py:207-322):
      * patch_embed Conv3d weight flattened to [D, 1176 -> 1216] (K padded to a multiple of 64)
      * attn.qkv [3D, D] (+bias): every head padded from head_dim 80 to 96 rows (zero rows, zero bias), q/k rows inside a
        head interleaved so that rotation pair (i, i + 40) is two adjacent output columns (lr_gemm_rope_ex_bf16,
        LR_EPI_BIAS_ROPE_F32); attn.proj [D, D] gets the matching zero COLUMNS
      * mlp gate/up (+biases) zero-padded from 3420 to 3456 rows each and interleaved in blocks of 128 for the
        SwiGLU epilogue (LR_EPI_BIAS_SWIGLU); down_proj gets the matching zero columns
    Decoder (same per-layer dict as the Phi-3 / LLaVA loop): q/k/v stacked into one [H + 2 kv, H + 3r] weight with the
    three LoRA-B blocks on the block diagonal of the K-extension, q/k rows head-interleaved, `qkv_b` the stacked bias;
    gate/up stacked [2I, H + 2r] and interleaved in blocks of 128."""
    bf = torch.bfloat16

    def g(name):
        return get(name).to(device=device, dtype=bf).contiguous()

    def rows(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """t[idx] with idx == -1 -> zero row"""
        out = t[idx.clamp(min=0)]
        out[idx < 0] = 0
        return out.contiguous()

    pw = PackedWeights()
    D, DI, nh, hd = cfg.vit_hidden, cfg.vit_intermediate, cfg.vit_heads, cfg.vit_head_dim
    hdp = qwen_vit_padded_head_dim(hd)
    K0 = cfg.patch_dim
    K0p = (K0 + 63) // 64 * 64
    pe = g("visual.patch_embed.proj.weight").reshape(D, -1)
    patch_w = torch.zeros(D, K0p, dtype=bf, device=device)
    patch_w[:, :K0] = pe
    pw.vit = {"patch_w": patch_w}
    # row index maps
    half = hd // 2
    inter = torch.stack([torch.arange(half), torch.arange(half) + half], 1).reshape(-1)          # 0,40,1,41,..
    qk_head = torch.cat([inter, torch.full((hdp - hd,), -1, dtype=torch.long)])                  # [hdp]
    v_head = torch.cat([torch.arange(hd), torch.full((hdp - hd,), -1, dtype=torch.long)])
    heads = torch.arange(nh)[:, None] * hd

    def head_map(per_head, base):
        m = heads + per_head[None, :]
        m = torch.where(per_head[None, :] < 0, torch.full_like(m, -1), m + base)
        return m.reshape(-1)

    qkv_idx = torch.cat([head_map(qk_head, 0), head_map(qk_head, D), head_map(v_head, 2 * D)]).to(device)
    proj_cols = head_map(v_head, 0).to(device)                                                    # [nh*hdp]
    Ip = (DI + 127) // 128 * 128
    pad_i = torch.cat([torch.arange(DI), torch.full((Ip - DI,), -1, dtype=torch.long)])
    gu_idx = torch.cat([pad_i, torch.where(pad_i < 0, pad_i, pad_i + DI)])
    gu_idx = gu_idx.view(2, Ip // 128, 128).permute(1, 0, 2).reshape(-1).to(device)
    dn_cols = pad_i.to(device)
    for i in range(cfg.vit_depth):
        p = f"visual.blocks.{i}."
        gu_w = torch.cat([g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")], 0)
        gu_b = torch.cat([g(p + "mlp.gate_proj.bias"), g(p + "mlp.up_proj.bias")], 0)
        pw.vit_layers.append({
            "n1": g(p + "norm1.weight"), "n2": g(p + "norm2.weight"),
            "qkv_w": rows(g(p + "attn.qkv.weight"), qkv_idx), "qkv_b": rows(g(p + "attn.qkv.bias"), qkv_idx),
            "proj_w": rows(g(p + "attn.proj.weight").t().contiguous(), proj_cols).t().contiguous(),
            "proj_b": g(p + "attn.proj.bias"),
            "gu_w": rows(gu_w, gu_idx), "gu_b": rows(gu_b, gu_idx),
            "dn_w": rows(g(p + "mlp.down_proj.weight").t().contiguous(), dn_cols).t().contiguous(),
            "dn_b": g(p + "mlp.down_proj.bias"),
        })
    pw.proj = {"ln_q": g("visual.merger.ln_q.weight"),
               "m0_w": g("visual.merger.mlp.0.weight"), "m0_b": g("visual.merger.mlp.0.bias"),
               "m2_w": g("visual.merger.mlp.2.weight"), "m2_b": g("visual.merger.mlp.2.bias")}
    pw.embed = g("model.embed_tokens.weight")
    H, I, r = cfg.hidden_size, cfg.intermediate_size, cfg.lora_rank
    kvw = cfg.num_kv_heads * cfg.head_dim
    assert I % 128 == 0 and (H + 2 * kvw) % 256 == 0 and (H + kvw) % 256 == 0
    gu_perm = torch.arange(2 * I, device=device).view(2, I // 128, 128).permute(1, 0, 2).reshape(-1)
    qkv_perm = torch.cat([_qk_interleave_perm(cfg.num_heads, cfg.head_dim, device),
                          H + _qk_interleave_perm(cfg.num_kv_heads, cfg.head_dim, device),
                          torch.arange(H + kvw, H + 2 * kvw, device=device)])

    def stacked(names):
        Ws = [g(n + ".weight") for n in names]
        if not cfg.use_lora:
            return torch.cat(Ws, 0).contiguous(), None
        k = len(names)
        As = [g(n + ".lora_A.weight") for n in names]
        r = As[0].shape[0]                      # the adapter's own rank (any value: the stack is zero-padded to 128)
        kr = _pad128(k * r)
        blocks = []
        for j, n in enumerate(names):
            B = (get(n + ".lora_B.weight").to(device=device, dtype=torch.float32) * cfg.lora_scale).to(bf)
            ext = torch.zeros(B.shape[0], kr, dtype=bf, device=device)
            ext[:, j * r:(j + 1) * r] = B
            blocks.append(torch.cat([Ws[j], ext], dim=1))
        A = torch.cat(As, 0)
        if kr != k * r:
            A = torch.cat([A, torch.zeros(kr - k * r, A.shape[1], dtype=bf, device=device)], 0)
        return torch.cat(blocks, 0).contiguous(), A.contiguous()

    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        names = [p + f"self_attn.{n}_proj" for n in "qkv"]
        qkv_w, qkv_a = stacked(names)
        qkv_b = torch.cat([g(n + ".bias") for n in names], 0)
        o_w, o_a = stacked([p + "self_attn.o_proj"])
        gu_w, gu_a = stacked([p + "mlp.gate_proj", p + "mlp.up_proj"])
        dn_w, dn_a = stacked([p + "mlp.down_proj"])
        d = {"in_ln": g(p + "input_layernorm.weight"), "post_ln": g(p + "post_attention_layernorm.weight"),
             "qkv_w": qkv_w[qkv_perm].contiguous(), "qkv_b": qkv_b[qkv_perm].contiguous(), "o_w": o_w,
             "gu_w": gu_w[gu_perm].contiguous(), "dn_w": dn_w}
        if cfg.use_lora:
            d.update({"qkv_a": qkv_a, "o_a": o_a, "gu_a": gu_a, "dn_a": dn_a})
        pw.layers.append(d)
    pw.head = {"norm": g("model.norm.weight"), "vh": g("value_head.weight")}
    if cfg.add_cross_attention:
        pw.head.update({"wq": g("W_q.weight"),
                        "wkv": torch.cat([g("W_k.weight"), g("W_v.weight")], 0).contiguous(),
                        "ca_ln": g("ca_layernorm.weight")})
    return pw
