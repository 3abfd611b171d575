"""Architecture constants of the reward-scoring path (Phi-3.5-vision LLaVA-Reward).

Mirrors what the reference reads from ``Phi3VConfig`` (reference
``llava_reward/models/base_mllm/phi3_v/configuration_phi3_v.py:107-217``) and the
fixed CLIP ViT-L/14-336 config (``modeling_phi3_v.py:68-83``), reduced to the
fields the forward path uses.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List


def _default_short_factor() -> List[float]:
    # The hub config.json of Phi-3.5-vision carries 48 su/longrope factors that are not in
    # the reference repo; without network we use a documented deterministic
    # list. Any 48 positive floats exercise the same code path.
    return [round(1.0 + 0.04 * i, 2) for i in range(48)]


def _default_long_factor() -> List[float]:
    return [round(1.0 + 1.25 * i, 2) for i in range(48)]


@dataclasses.dataclass
class RewardConfig:
    # Phi-3 decoder
    vocab_size: int = 32064
    hidden_size: int = 3072
    intermediate_size: int = 8192
    num_layers: int = 32
    num_heads: int = 32
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 131072
    original_max_position_embeddings: int = 4096
    short_factor: List[float] = dataclasses.field(default_factory=_default_short_factor)
    long_factor: List[float] = dataclasses.field(default_factory=_default_long_factor)
    # CLIP ViT-L/14-336 vision tower; the path uses the output of encoder layer
    # ``clip_layers`` (= hidden_states[-2] of a 24-layer tower, modeling_phi3_v.py:208-219)
    clip_hidden: int = 1024
    clip_intermediate: int = 4096
    clip_heads: int = 16
    clip_layers: int = 23
    clip_eps: float = 1e-5
    image_size: int = 336
    patch: int = 14
    num_crops: int = 16
    # LoRA (peft 0.13.2 semantics: y = Wx + (alpha/r) B A x), targets qkv/o/gate_up/down
    lora_rank: int = 128
    lora_alpha: float = 256.0
    use_lora: bool = True
    # reward heads (reference rw_model_general_preference.py:306-333)
    is_general_preference: bool = True
    add_cross_attention: bool = True
    value_head_dim: int = 2
    general_preference_tau: float = 0.1

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def clip_head_dim(self) -> int:
        return self.clip_hidden // self.clip_heads

    @property
    def clip_tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_rank

    @property
    def vhd(self) -> int:
        return self.value_head_dim if self.is_general_preference else 1

    @property
    def rope_scaling_factor(self) -> float:
        scale = self.max_position_embeddings / self.original_max_position_embeddings
        if scale <= 1.0:
            return 1.0
        return math.sqrt(1 + math.log(scale) / math.log(self.original_max_position_embeddings))


def num_image_tokens(h: int, w: int) -> int:
    """Image-token count for a padded HD size (reference processing_phi3_v.py:269)."""
    hc, wc = h // 336, w // 336
    return (hc * wc + 1) * 144 + 1 + (hc + 1) * 12


# --------------------------------------------------------------------------------------
# LLaVA-v1.6 (LlavaNext) backbone: reference branch model_type == 'llava'
# (rw_model_general_preference.py:372-375, eval/reward_adaptor_loader.py:110-151)
# --------------------------------------------------------------------------------------
def _default_pinpoints() -> List[List[int]]:
    # llava-hf/llava-v1.6-vicuna-*-hf config.json `image_grid_pinpoints` ("anyres-672")
    return [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]


@dataclasses.dataclass
class LlavaNextRewardConfig:
    """LLaVA-v1.6 Vicuna reward model: CLIP ViT-L/14-336 (hidden_states[-2], CLS dropped) -> 2-layer GELU projector ->
    anyres 'spatial_unpad' packing with image_newline -> Llama decoder (MHA, plain RoPE) with LoRA on q/k/v/o/gate/up/down
    (create_lora_config_llava16_vicuna, llava_reward/utils/utils.py:243-262) -> final norm -> value head on the last
    valid token. Defaults = llava-v1.6-vicuna-7b (BASELINE.json configs[4]); `vicuna_13b()` gives the model the
    reference's training script names (scripts/run_train_rm_single_lora_llava.sh:9)."""
    vocab_size: int = 32064
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_layers: int = 32
    num_heads: int = 32
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    image_token_id: int = 32000
    image_grid_pinpoints: List[List[int]] = dataclasses.field(default_factory=_default_pinpoints)
    clip_hidden: int = 1024
    clip_intermediate: int = 4096
    clip_heads: int = 16
    clip_layers: int = 23          # vision_feature_layer = -2 of a 24-layer tower
    clip_eps: float = 1e-5
    image_size: int = 336
    patch: int = 14
    lora_rank: int = 128
    lora_alpha: float = 256.0
    use_lora: bool = True
    is_general_preference: bool = False
    value_head_dim: int = 2
    general_preference_tau: float = 0.1
    add_cross_attention: bool = False   # the reference's llava branch never applies SkipCA (:376-397 has no llava arm)

    @classmethod
    def vicuna_13b(cls, **kw) -> "LlavaNextRewardConfig":
        return cls(hidden_size=5120, intermediate_size=13824, num_layers=40, num_heads=40, **kw)

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def clip_head_dim(self) -> int:
        return self.clip_hidden // self.clip_heads

    @property
    def clip_tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_rank

    @property
    def vhd(self) -> int:
        return self.value_head_dim if self.is_general_preference else 1

    # fields RewardEngine's shared decoder / CLIP code reads
    short_factor = None
    long_factor = None
    rope_scaling_factor = 1.0
    original_max_position_embeddings = 1 << 30


def select_best_resolution(original_hw, pinpoints) -> tuple:
    """(height, width) of the grid pinpoint that keeps most of the image at least downscaling
    (transformers image_processing_utils.select_best_resolution, called by modeling_llava_next.py:41-69)."""
    oh, ow = int(original_hw[0]), int(original_hw[1])
    best, best_eff, best_waste = None, 0, float("inf")
    for h, w in pinpoints:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > best_eff or (eff == best_eff and waste < best_waste):
            best, best_eff, best_waste = (h, w), eff, waste
    return best


def anyres_geometry(original_hw, pinpoints, image_size: int = 336, patch: int = 14):
    """Per-image packing geometry of LlavaNextModel.pack_image_features (modeling_llava_next.py:277-343):
    returns dict(grid_h, grid_w, n_patches, top, left, keep_h, keep_w, n_tokens)."""
    side = image_size // patch  # 24
    bh, bw = select_best_resolution(original_hw, pinpoints)
    gh, gw = bh // image_size, bw // image_size
    cur_h, cur_w = gh * side, gw * side
    oh, ow = int(original_hw[0]), int(original_hw[1])
    top = left = 0
    if ow / oh > cur_w / cur_h:
        new_h = int(round(oh * (cur_w / ow), 7))
        top = (cur_h - new_h) // 2
    else:
        new_w = int(round(ow * (cur_h / oh), 7))
        left = (cur_w - new_w) // 2
    keep_h, keep_w = cur_h - 2 * top, cur_w - 2 * left
    return dict(grid_h=gh, grid_w=gw, n_patches=gh * gw + 1, top=top, left=left, keep_h=keep_h, keep_w=keep_w,
                n_tokens=side * side + keep_h * (keep_w + 1))


# --------------------------------------------------------------------------------------
# Qwen2.5-VL backbone: reference branch model_type == 'qwen'
# (rw_model_general_preference.py:354-371, 387-397; eval/reward_adaptor_loader.py:64-109)
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class QwenVLRewardConfig:
    """Qwen2.5-VL reward model (BASELINE.json configs[3]): window-attention ViT (RMSNorm, SwiGLU MLP with biases, 2D
    rotary, 2x2 patch merger) -> Qwen2 decoder (GQA, q/k/v biases, M-RoPE sections [16,24,24]) with LoRA on
    q/k/v/o/gate/up/down (create_lora_config_qwen, llava_reward/utils/utils.py:223-242) -> final norm -> optional SkipCA
    (qwen arm, :387-397) -> value head on the last valid token. Defaults = Qwen/Qwen2.5-VL-7B-Instruct."""
    vocab_size: int = 152064
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_layers: int = 28
    num_heads: int = 28
    num_kv_heads: int = 4
    rms_eps: float = 1e-6
    rope_theta: float = 1000000.0
    mrope_section: List[int] = dataclasses.field(default_factory=lambda: [16, 24, 24])
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653
    pad_token_id: int = 151643     # also the token id the reference's SkipCA arm keys its "vision" rows on (:358)
    # vision tower
    vit_depth: int = 32
    vit_hidden: int = 1280
    vit_intermediate: int = 3420
    vit_heads: int = 16
    vit_patch: int = 14
    vit_temporal_patch: int = 2
    vit_merge: int = 2
    vit_window: int = 112
    vit_fullatt: List[int] = dataclasses.field(default_factory=lambda: [7, 15, 23, 31])
    vit_eps: float = 1e-6
    # LoRA + heads
    lora_rank: int = 128
    lora_alpha: float = 256.0
    use_lora: bool = True
    is_general_preference: bool = False
    value_head_dim: int = 2
    general_preference_tau: float = 0.1
    add_cross_attention: bool = False

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_heads

    @property
    def vit_head_dim(self) -> int:
        return self.vit_hidden // self.vit_heads

    @property
    def patch_dim(self) -> int:
        return 3 * self.vit_temporal_patch * self.vit_patch * self.vit_patch

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_rank

    @property
    def vhd(self) -> int:
        return self.value_head_dim if self.is_general_preference else 1


def qwen_window_plan(grid_thw, merge: int = 2, window: int = 112, patch: int = 14):
    """Host-side index plan of the Qwen2.5-VL vision tower for a list of (t, h, w) patch grids
    (transformers modeling_qwen2_5_vl.py: rot_pos_emb :382-409, get_window_index :411-453, forward :470-500):
      window_index [T/4]  merged-unit order after the window permutation (new unit j = old unit window_index[j])
      src_row [T]         patch row of window-ordered token i
      pos_hw [T, 2]       (h, w) patch coordinates of window-ordered token i (rotary)
      win_cu              cumulative token counts of the non-empty windows (cu_window_seqlens after unique_consecutive)
      img_cu              cumulative token counts per frame (cu_seqlens of the full-attention blocks)."""
    import numpy as np

    unit = merge * merge
    wsz = window // merge // patch
    widx, win_cu, img_cu, pos = [], [0], [0], []
    base = 0
    for t, h, w in grid_thw:
        t, h, w = int(t), int(h), int(w)
        lh, lw = h // merge, w // merge
        hp = np.broadcast_to(np.arange(h)[:, None], (h, w)).reshape(lh, merge, lw, merge).transpose(0, 2, 1, 3).reshape(-1)
        wp = np.broadcast_to(np.arange(w)[None, :], (h, w)).reshape(lh, merge, lw, merge).transpose(0, 2, 1, 3).reshape(-1)
        pos.append(np.tile(np.stack([hp, wp], -1), (t, 1)))
        idx = np.arange(t * lh * lw).reshape(t, lh, lw)
        ph, pw_ = wsz - lh % wsz, wsz - lw % wsz
        nh, nw = (lh + ph) // wsz, (lw + pw_) // wsz
        pad = np.full((t, lh + ph, lw + pw_), -100, dtype=np.int64)
        pad[:, :lh, :lw] = idx
        pad = pad.reshape(t, nh, wsz, nw, wsz).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, wsz, wsz)
        lens = (pad != -100).sum((2, 3)).reshape(-1)
        flat = pad.reshape(-1)
        widx.append(flat[flat != -100] + base)
        for n in lens:
            if n > 0:
                win_cu.append(win_cu[-1] + int(n) * unit)
        for _ in range(t):
            img_cu.append(img_cu[-1] + h * w)
        base += t * lh * lw
    widx = np.concatenate(widx)
    pos = np.concatenate(pos, 0)
    src_row = (widx[:, None] * unit + np.arange(unit)[None, :]).reshape(-1)
    return dict(window_index=widx.astype(np.int32), src_row=src_row.astype(np.int32),
                pos_hw=pos[src_row].astype(np.int32), win_cu=np.asarray(win_cu, dtype=np.int32),
                img_cu=np.asarray(img_cu, dtype=np.int32))


def packed_row_plan(seq_start, seq_len, S: int, pos_from_zero: bool):
    """Host-side index plan of the packed decoder layout (engine.pack_rows): sample b's valid run
    [seq_start[b], +seq_len[b]) of the slot layout [B, S] moves to rows [base[b], +seq_len[b]).
    -> (row_index [rows] into the slot layout, position [rows], base [B], last_row [B]) as int32 numpy arrays.
    Positions count from the first valid token when position_ids = cumsum(mask) - 1 (phi3v, and the order in which the
    per-token M-RoPE rows of the qwen branch are gathered) or keep the slot index when position_ids = arange(S) (llava)."""
    import numpy as np
    start = np.asarray(seq_start, dtype=np.int64)
    length = np.asarray(seq_len, dtype=np.int64)
    B = start.shape[0]
    if (length < 0).any() or (start < 0).any() or (start + length > S).any():
        raise ValueError("valid runs must lie inside [0, S)")
    base = np.concatenate([[0], np.cumsum(length)[:-1]]) if B else np.zeros(0, dtype=np.int64)
    idx = np.concatenate([b * S + start[b] + np.arange(length[b]) for b in range(B)]) if B else np.zeros(0, np.int64)
    pos = np.concatenate([(0 if pos_from_zero else start[b]) + np.arange(length[b]) for b in range(B)]) if B else idx
    return idx.astype(np.int32), pos.astype(np.int32), base.astype(np.int32), (base + length - 1).astype(np.int32)
