"""TEST INFRASTRUCTURE - harness that runs the UNMODIFIED reference (sjz5202/LLaVA-Reward) for the scoring path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s reference / cpu_baseline / gpu_reference legs import this
module; the product package never does.

The reference is Python only. It is imported from a byte-for-byte copy under `baseline/_ref/` (git-ignored, made by
`tools/make_baseline_ref.sh`, shipped to the GPU box by gpurun) or, in the build container, straight from
`/root/reference`. Its arithmetic is untouched; what the harness adds is what a user of the reference would have to
provide anyway in this offline image:

* stub modules for third-party packages the reference imports at module top but never executes on the scoring path
  (deepspeed, peft, accelerate, loralib - llava_reward/utils/deepspeed.py:10-19, rw_model_general_preference.py:4-8,
  eval/reward_adaptor_loader.py:11,15) - recipe of SURVEY.md 8(c);
* a hand-written `Phi3VConfig` (the hub config.json is unreachable) with `use_cache=False`;
* deterministic synthetic weights (`llava_reward_b200.synth.SynthProvider`) copied into the reference's parameters;
* LoRA: peft is neither installed nor vendored, so the adapters are attached with `LoraWrapped`, a restatement of peft
  0.13.2 `lora.Linear.forward` (base(x) + lora_B(lora_A(x)) * alpha/r). PARITY UNPINNED at that boundary;
* flash-attention 2 on the GPU: the reference's own `Phi3FlashAttention2` / `CLIPAttentionFA2` classes
  (modeling_phi3_v.py:723-1029, 85-115) are selected by re-classing the attention modules of an eager-built model
  (transformers 5.5 no longer honours the 4.x `_supports_flash_attn_2` flag the vendored class sets).
  The weights are the same objects; only `forward` changes, to the reference's FA2 forward.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def reference_root() -> str:
    for p in _CANDIDATES:
        if os.path.isdir(os.path.join(p, "llava_reward")):
            return p
    raise FileNotFoundError("reference sources not found: run tools/make_baseline_ref.sh in the build container "
                            "(copies /root/reference -> baseline/_ref)")


def available() -> bool:
    try:
        reference_root()
        return True
    except FileNotFoundError:
        return False


_REFMODS = None


def import_reference():
    """-> (_get_reward_model, modeling_phi3_v module, reference eval/reward_adaptor_loader module)"""
    global _REFMODS
    if _REFMODS is not None:
        return _REFMODS
    import transformers  # noqa: F401  (must be imported before the stubs are registered)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    if "deepspeed" not in sys.modules:
        stub("accelerate", Accelerator=_Dummy)
        ds = stub("deepspeed")
        ds.zero = stub("deepspeed.zero", GatheredParameters=_Dummy)
        ds.ops = stub("deepspeed.ops")
        ds.ops.adam = stub("deepspeed.ops.adam", DeepSpeedCPUAdam=_Dummy, FusedAdam=_Dummy)
        ds.runtime = stub("deepspeed.runtime")
        ds.runtime.zero = stub("deepspeed.runtime.zero")
        ds.runtime.zero.partition_parameters = stub("deepspeed.runtime.zero.partition_parameters",
                                                    ZeroParamStatus=_Dummy)
        pf = stub("peft", LoraConfig=_Dummy, get_peft_model=_Dummy, PeftModel=_Dummy,
                  get_peft_model_state_dict=_Dummy)
        pf.tuners = stub("peft.tuners")
        pf.tuners.lora = stub("peft.tuners.lora", LoraLayer=_Dummy)
        stub("loralib")
    ref = reference_root()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from llava_reward.models import _get_reward_model  # noqa
    from llava_reward.models.base_mllm.phi3_v import modeling_phi3_v as mp
    # /root/repo/eval is a regular package and would shadow the reference's namespace package `eval`:
    # load the reference file by path
    spec = importlib.util.spec_from_file_location("ref_reward_adaptor_loader",
                                                  os.path.join(ref, "eval", "reward_adaptor_loader.py"))
    ral = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ral)
    _REFMODS = (_get_reward_model, mp, ral)
    return _REFMODS


class LoraWrapped(torch.nn.Module):
    """peft 0.13.2 lora.Linear.forward restated (dropout is identity in eval)."""

    def __init__(self, base, A, B, scale):
        super().__init__()
        self.base, self.scale = base, scale
        dev, dt = base.weight.device, base.weight.dtype
        self.lora_A = torch.nn.Linear(A.shape[1], A.shape[0], bias=False, device=dev, dtype=dt)
        self.lora_B = torch.nn.Linear(B.shape[1], B.shape[0], bias=False, device=dev, dtype=dt)
        self.lora_A.weight.data.copy_(A)
        self.lora_B.weight.data.copy_(B)

    def forward(self, x):
        return self.base(x) + self.lora_B(self.lora_A(x)) * self.scale


@contextlib.contextmanager
def _skip_random_init():
    """Every parameter is overwritten by the synthetic generator right after construction, so the minutes torch spends
    on kaiming/normal initialisation of 4 G fp32 parameters on the CPU are skipped (harness only: no arithmetic of the
    scoring path is involved)."""
    import torch.nn.init as init

    names = ("kaiming_uniform_", "uniform_", "normal_", "trunc_normal_", "xavier_uniform_", "zeros_", "ones_")
    saved = {n: getattr(init, n) for n in names}
    saved_t = {n: getattr(torch.Tensor, n) for n in ("normal_", "uniform_")}
    try:
        for n in names:
            if n not in ("zeros_", "ones_"):
                setattr(init, n, lambda t, *a, **k: t)
        for n in saved_t:
            setattr(torch.Tensor, n, lambda t, *a, **k: t)
        yield
    finally:
        for n, f in saved.items():
            setattr(init, n, f)
        for n, f in saved_t.items():
            setattr(torch.Tensor, n, f)


def phi3v_config(cfg, mp):
    from llava_reward.models.base_mllm.phi3_v.configuration_phi3_v import Phi3VConfig

    mp.CLIP_VIT_LARGE_PATCH14_336_CONFIG.num_hidden_layers = cfg.clip_layers + 1
    rcfg = Phi3VConfig(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_heads,
        max_position_embeddings=cfg.max_position_embeddings,
        original_max_position_embeddings=cfg.original_max_position_embeddings,
        rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta,
        rope_scaling={"type": "su", "short_factor": cfg.short_factor, "long_factor": cfg.long_factor},
        sliding_window=262144,
        embd_layer={"embedding_cls": "image", "hd_transform_order": "sub_glb", "projection_cls": "mlp",
                    "use_hd_transform": True, "with_learnable_separator": True},
        img_processor={"name": "clip_vision_model", "model_name": "openai/clip-vit-large-patch14-336",
                       "image_dim_out": 1024, "num_img_tokens": 144},
    )
    rcfg.use_cache = False
    rcfg._attn_implementation = "eager"
    return rcfg


def build_reference_model(cfg, seed, refmods=None, device="cpu", dtype=torch.float32, fast_init=True, verbose=True,
                          gen_device=None):
    """The reference's CustomRewardModel (phi3v) with the synthetic weights of `SynthProvider(cfg, seed)`, eager
    attention, eval mode, on `device` in `dtype` (weights are generated in fp32 and rounded once, like the engine's).
    `gen_device`: where the counter-hash generator runs (bit-identical on CPU and CUDA; a GPU makes 4.4 G values in
    seconds instead of minutes) - the values are then copied to `device`."""
    from llava_reward_b200.synth import SynthProvider

    _get_reward_model, mp, _ = refmods or import_reference()
    rcfg = phi3v_config(cfg, mp)
    cls = _get_reward_model(mp.Phi3VForCausalLM, mp.Phi3VModel, RMSNorm_class=mp.Phi3RMSNorm,
                            RMSNorm_class_eps=1e-5, is_general_preference=cfg.is_general_preference,
                            add_cross_attention=cfg.add_cross_attention, value_head_dim=cfg.value_head_dim)
    t0 = time.time()
    device = torch.device(device)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ctx = _skip_random_init() if fast_init else contextlib.nullcontext()
        with ctx, torch.device(device):
            model = cls(rcfg)
    finally:
        torch.set_default_dtype(old)
    model.model_type = "phi3v"
    model.eval()
    prov = SynthProvider(cfg, seed=seed, device=gen_device if gen_device is not None else device)
    sd = model.state_dict()
    used = set()
    with torch.no_grad():
        for name, t in sd.items():
            key = name.replace("model.vision_embed_tokens.wte.", "model.embed_tokens.")
            if key in prov:
                t.copy_(prov(key))
                used.add(key)
            elif f"encoder.layers.{cfg.clip_layers}." in name or "post_layernorm" in name or name.startswith("lm_head."):
                # never read by custom_forward: CLIP features = hidden_states[-2] without pooling, no logits
                t.zero_()
            elif t.dtype.is_floating_point and "inv_freq" not in name and "position_ids" not in name:
                raise AssertionError(f"reference parameter {name} has no synthetic value")
    missing = [n for n in prov.names() if n not in used and ".lora_" not in n]
    assert not missing, missing
    if cfg.use_lora:
        for i, layer in enumerate(model.model.layers):
            for holder, attr, nm in ((layer.self_attn, "qkv_proj", "self_attn.qkv_proj"),
                                     (layer.self_attn, "o_proj", "self_attn.o_proj"),
                                     (layer.mlp, "gate_up_proj", "mlp.gate_up_proj"),
                                     (layer.mlp, "down_proj", "mlp.down_proj")):
                p = f"model.layers.{i}.{nm}"
                setattr(holder, attr, LoraWrapped(getattr(holder, attr), prov(p + ".lora_A.weight"),
                                                  prov(p + ".lora_B.weight"), cfg.lora_scale))
    if verbose:
        print(f"  reference model built in {time.time() - t0:.1f}s on {device} ({dtype}), "
              f"{sum(p.numel() for p in model.parameters()) / 1e6:.1f} M params", flush=True)
    return model


def set_attention(model, impl: str, refmods=None):
    """Switch an eager-built reference model between its two shipped attention paths by re-classing the attention
    modules: 'flash_attention_2' -> Phi3FlashAttention2 + CLIPAttentionFA2, 'eager' -> Phi3Attention +
    transformers' CLIPAttention. Weights are untouched."""
    _, mp, _ = refmods or import_reference()
    from transformers.models.clip.modeling_clip import CLIPAttention

    fa2 = impl == "flash_attention_2"
    if fa2 and not hasattr(mp, "flash_attn_func"):
        raise ImportError("flash_attn is not importable: the reference's FA2 path cannot run")
    if not hasattr(mp, "_CLIPAttentionFA2Compat"):
        class _CLIPAttentionFA2Compat(mp.CLIPAttentionFA2):
            """transformers 5.5's CLIPEncoderLayer passes keyword arguments the 4.50-era signature of the reference's
            class does not list; drop them and run the reference's forward unchanged."""

            def forward(self, hidden_states, attention_mask=None, **kwargs):
                return mp.CLIPAttentionFA2.forward(self, hidden_states, attention_mask)

        mp._CLIPAttentionFA2Compat = _CLIPAttentionFA2Compat
    for layer in model.model.layers:
        layer.self_attn.__class__ = mp.Phi3FlashAttention2 if fa2 else mp.Phi3Attention
        if fa2:
            layer.self_attn._flash_attn_uses_top_left_mask = False
    for layer in model.model.vision_embed_tokens.img_processor.vision_model.encoder.layers:
        layer.self_attn.__class__ = mp._CLIPAttentionFA2Compat if fa2 else CLIPAttention
    model.model._attn_implementation = impl
    model.config._attn_implementation = impl
    return model


def preference_args(cfg):
    return types.SimpleNamespace(is_general_preference=cfg.is_general_preference, value_head_dim=cfg.value_head_dim,
                                 general_preference_tau=cfg.general_preference_tau)
