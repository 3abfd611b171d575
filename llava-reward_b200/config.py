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
