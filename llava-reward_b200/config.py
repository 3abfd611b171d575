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
