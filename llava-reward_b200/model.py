"""Model object with the reference's scoring surface (`custom_forward`, `.to()`, `.eval()`).

Mirrors `CustomRewardModel` (reference llava_reward/models/rw_model_general_preference.py:303-448) for
model_type 'phi3v'. It is NOT an nn.Module wrapper around torch layers: `.to('cuda')` packs the
parameters into kernel layouts and `custom_forward` runs `RewardEngine.forward`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib as L
from .config import LlavaNextRewardConfig, QwenVLRewardConfig, RewardConfig
from .engine import LlavaNextRewardEngine, QwenVLRewardEngine, RewardEngine
from .weights import pack_weights, pack_weights_llava, pack_weights_qwen


class B200RewardModel:
    model_type = "phi3v"

    def __init__(self, cfg: RewardConfig, provider: Callable[[str], torch.Tensor]):
        self.config = cfg
        self._provider = provider
        self.engine: Optional[RewardEngine] = None
        self.training = False
        # attributes the reference's custom_forward honours (rw_model_general_preference.py:327-333)
        self.mean_hidden_state = None
        self.layer_id = 32
        self.vision_layer_id = -1
        self.value_head_dim = cfg.vhd
        self.add_cross_attention = cfg.add_cross_attention
        self.is_general_preference = cfg.is_general_preference
        self.device = torch.device("cpu")
        self.dtype = torch.bfloat16
        # "fp32" selects the verification path (engine.py / csrc/f32_verify.cu); set before .to('cuda')
        self.precision = "bf16"

    # --- nn.Module-like plumbing used by the reference's callers (eval/simple_inference.py:17-18)
    def to(self, device=None, *args, **kwargs):
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise RuntimeError("llava-reward-b200 runs on sm_100a CUDA devices only; there is no CPU fallback")
        L.load()
        L.check(L.load().lr_device_check(), "lr_device_check")
        if hasattr(self._provider, "device"):
            self._provider.device = device  # synthetic weights are generated directly on the GPU
        with torch.cuda.device(device):
            self.engine = self._build_engine(device)
        self.device = device
        return self

    def _build_engine(self, device):
        dt = torch.float32 if self.precision == "fp32" else torch.bfloat16
        self.dtype = dt
        return RewardEngine(self.config, pack_weights(self.config, self._provider, device=device, dtype=dt), device=device,
                            precision=self.precision)

    def cuda(self, device=None):
        return self.to("cuda" if device is None else device)

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        """Sets the `training` attribute custom_forward reads. There is no backward and no dropout on this path (the
        reference's scoring configs have every dropout probability at 0), so the only effect is the reference's
        training-mode gather: the reward is read at the LAST position instead of the last valid token
        (rw_model_general_preference.py:413-418, 432-436) - identical for left-padded batches."""
        self.training = bool(mode)
        return self

    def parameters(self):
        return iter(())

    # --- the hot path
    def custom_forward(self, input_ids: torch.LongTensor = None, attention_mask: Optional[torch.Tensor] = None,
                       pixel_values: Optional[torch.FloatTensor] = None, image_sizes: Optional[torch.LongTensor] = None,
                       return_output=False, inputs_batch=None):
        """-> (reward [B, vhd] (GPM) or [B, 1] (BT), bf16 on device, None).
        Same signature/positional order as the reference (rw_model_general_preference.py:334-342)."""
        if self.engine is None:
            raise RuntimeError("call .to('cuda') before custom_forward (weights are packed on the device)")
        if inputs_batch is not None:
            raise NotImplementedError("inputs_batch is the qwen/llava calling convention; this build covers phi3v")
        if input_ids is None or attention_mask is None:
            raise ValueError("input_ids and attention_mask are required")
        if pixel_values is None or image_sizes is None:
            # the reference path is image-only by construction (UnboundLocalError at modeling_phi3_v.py:252)
            raise ValueError("pixel_values and image_sizes are required (the scoring path is image-only)")
        kw = dict(layer_id=self.layer_id, last_position=bool(self.training) and not self.mean_hidden_state,
                  mean_pool=bool(self.mean_hidden_state), vision_layer_id=int(self.vision_layer_id))
        with torch.cuda.device(self.device):
            if not return_output:
                reward = self.engine.forward(input_ids, attention_mask, pixel_values, image_sizes, **kw)
                return self._shape_like_reference(reward), None
            # return_output: the trainer-side callers read outputs["hidden_states"] / ["last_hidden_state"]
            # (rw_model_general_preference.py:422-425, 445-448). Every layer's hidden state is copied out of the
            # in-place residual buffer - a debugging / analysis path, not the scoring path.
            if self.engine._resolve_layer_id(self.layer_id)[0] != self.config.num_layers:
                raise NotImplementedError("return_output=True together with an early layer_id")
            self.engine.taps = {}
            try:
                reward = self.engine.forward(input_ids, attention_mask, pixel_values, image_sizes, **kw)
                outputs = self.engine.model_outputs(self.engine.taps, *input_ids.shape)
            finally:
                self.engine.taps = None
        return self._shape_like_reference(reward), outputs

    def _shape_like_reference(self, reward):
        """BT + training-mode gather returns `values.squeeze(-1)[:, -1]`, shape [B] (rw_model_general_preference.py
        :413-414); every other combination is [B, vhd]."""
        if self.training and not self.mean_hidden_state and not self.is_general_preference:
            return reward[:, 0]
        return reward

    __call__ = custom_forward


class B200LlavaNextRewardModel(B200RewardModel):
    """model_type 'llava' (LLaVA-v1.6 Vicuna): the caller passes the processor's BatchFeature as `inputs_batch`
    (reference eval/batch_inference_rm_llava.py:86-87, rw_model_general_preference.py:372-375)."""
    model_type = "llava"

    def __init__(self, cfg: LlavaNextRewardConfig, provider: Callable[[str], torch.Tensor]):
        super().__init__(cfg, provider)
        self.layer_id = cfg.num_layers

    def _build_engine(self, device):
        if self.precision != "bf16":
            raise NotImplementedError("the fp32 verification path covers the phi3v backbone")
        return LlavaNextRewardEngine(self.config, pack_weights_llava(self.config, self._provider, device=device),
                                     device=device)

    def custom_forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_sizes=None,
                       return_output=False, inputs_batch=None):
        if self.engine is None:
            raise RuntimeError("call .to('cuda') before custom_forward (weights are packed on the device)")
        if inputs_batch is None:
            # the reference reads inputs_batch['attention_mask'] unconditionally in this branch (:373)
            raise TypeError("model_type 'llava' is called as custom_forward(inputs_batch=processor_output)")
        for k in ("input_ids", "attention_mask", "pixel_values", "image_sizes"):
            if k not in inputs_batch:
                raise KeyError(k)
        outputs = None
        with torch.cuda.device(self.device):   # `layer_id` is not read by the reference in this branch (:372-375)
            if return_output:
                self.engine.taps = {}
            try:
                reward = self.engine.forward(inputs_batch["input_ids"], inputs_batch["attention_mask"],
                                             inputs_batch["pixel_values"], inputs_batch["image_sizes"],
                                             last_position=bool(self.training) and not self.mean_hidden_state,
                                             mean_pool=bool(self.mean_hidden_state))
                if return_output:   # hidden_states of every layer; logits stay None (lm_head is not executed)
                    outputs = self.engine.lm_outputs(self.engine.taps, *inputs_batch["input_ids"].shape)
            finally:
                if return_output:
                    self.engine.taps = None
        return self._shape_like_reference(reward), outputs

    __call__ = custom_forward


class B200QwenRewardModel(B200RewardModel):
    """model_type 'qwen' (Qwen2.5-VL): the caller passes the processor's BatchFeature as `inputs_batch`
    (reference eval/batch_inference_rm_qwen.py:91-92, rw_model_general_preference.py:354-371, 387-397)."""
    model_type = "qwen"

    def __init__(self, cfg: QwenVLRewardConfig, provider: Callable[[str], torch.Tensor]):
        super().__init__(cfg, provider)
        self.layer_id = cfg.num_layers

    def _build_engine(self, device):
        if self.precision != "bf16":
            raise NotImplementedError("the fp32 verification path covers the phi3v backbone")
        return QwenVLRewardEngine(self.config, pack_weights_qwen(self.config, self._provider, device=device),
                                  device=device)

    def custom_forward(self, input_ids=None, attention_mask=None, pixel_values=None, image_sizes=None,
                       return_output=False, inputs_batch=None):
        if self.engine is None:
            raise RuntimeError("call .to('cuda') before custom_forward (weights are packed on the device)")
        if inputs_batch is None:
            # the reference reads inputs_batch['attention_mask'] unconditionally in this branch (:355)
            raise TypeError("model_type 'qwen' is called as custom_forward(inputs_batch=processor_output)")
        for k in ("input_ids", "attention_mask", "pixel_values", "image_grid_thw"):
            if k not in inputs_batch:
                raise KeyError(k)
        if inputs_batch.get("pixel_values_videos") is not None:
            raise NotImplementedError("video inputs: the reference's reward datasets are image-only")
        outputs = None
        with torch.cuda.device(self.device):
            if return_output:
                self.engine.taps = {}
            try:
                reward = self.engine.forward(inputs_batch["input_ids"], inputs_batch["attention_mask"],
                                             inputs_batch["pixel_values"], inputs_batch["image_grid_thw"],
                                             last_position=bool(self.training) and not self.mean_hidden_state,
                                             mean_pool=bool(self.mean_hidden_state))
                if return_output:   # hidden_states of every layer; logits stay None (lm_head is not executed)
                    outputs = self.engine.lm_outputs(self.engine.taps, *inputs_batch["input_ids"].shape)
            finally:
                if return_output:
                    self.engine.taps = None
        return self._shape_like_reference(reward), outputs

    __call__ = custom_forward
