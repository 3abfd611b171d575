"""Generate golden vectors by executing the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py slim_gpm slim_bt            # seconds .. a minute each
    python tests/golden/make_golden.py full_bt full_gpm           # minutes each, ~20 GB RAM

The reference (/root/reference, Python only) is imported with stub modules for the third-party
packages that are absent here and never executed on the scoring path (deepspeed, peft,
accelerate, loralib) - recipe of SURVEY.md 8(c). Weights/inputs come from the deterministic
generator in `llava_reward_b200.synth`, so a fixture only stores the config, seeds and the
reference's outputs (rewards, probabilities, strided samples of intermediate tensors).

LoRA: peft is not installed and not vendored by the reference, so the adapters are attached
with a 10-line wrapper that restates peft 0.13.2 `lora.Linear.forward`
(base(x) + lora_B(lora_A(x)) * alpha/r). Everything else is the reference's own code.

/root/reference does not exist on the GPU box; this script only runs here.
"""
from __future__ import annotations

import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
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

    stub("accelerate", Accelerator=_Dummy)
    ds = stub("deepspeed")
    ds.zero = stub("deepspeed.zero", GatheredParameters=_Dummy)
    ds.ops = stub("deepspeed.ops")
    ds.ops.adam = stub("deepspeed.ops.adam", DeepSpeedCPUAdam=_Dummy, FusedAdam=_Dummy)
    ds.runtime = stub("deepspeed.runtime")
    ds.runtime.zero = stub("deepspeed.runtime.zero")
    ds.runtime.zero.partition_parameters = stub("deepspeed.runtime.zero.partition_parameters", ZeroParamStatus=_Dummy)
    pf = stub("peft", LoraConfig=_Dummy, get_peft_model=_Dummy, PeftModel=_Dummy, get_peft_model_state_dict=_Dummy)
    pf.tuners = stub("peft.tuners")
    pf.tuners.lora = stub("peft.tuners.lora", LoraLayer=_Dummy)
    stub("loralib")
    sys.path.insert(0, REF)
    from llava_reward.models import _get_reward_model  # noqa
    from llava_reward.models.base_mllm.phi3_v import modeling_phi3_v as mp
    # /root/repo/eval is a regular package and would shadow the reference's namespace package `eval`:
    # load the reference file by path
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_reward_adaptor_loader", os.path.join(REF, "eval", "reward_adaptor_loader.py"))
    ral = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ral)
    return _get_reward_model, mp, ral


class LoraWrapped(torch.nn.Module):
    """peft 0.13.2 lora.Linear.forward restated (dropout is identity in eval)."""

    def __init__(self, base, A, B, scale):
        super().__init__()
        self.base, self.scale = base, scale
        self.lora_A = torch.nn.Linear(A.shape[1], A.shape[0], bias=False)
        self.lora_B = torch.nn.Linear(B.shape[1], B.shape[0], bias=False)
        self.lora_A.weight.data.copy_(A)
        self.lora_B.weight.data.copy_(B)

    def forward(self, x):
        return self.base(x) + self.lora_B(self.lora_A(x)) * self.scale


def build_reference_model(cfg, seed, refmods):
    from llava_reward_b200.synth import SynthProvider

    _get_reward_model, mp, _ = refmods
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
    cls = _get_reward_model(mp.Phi3VForCausalLM, mp.Phi3VModel, RMSNorm_class=mp.Phi3RMSNorm,
                            RMSNorm_class_eps=1e-5, is_general_preference=cfg.is_general_preference,
                            add_cross_attention=cfg.add_cross_attention, value_head_dim=cfg.value_head_dim)
    t0 = time.time()
    model = cls(rcfg)
    model.model_type = "phi3v"
    model.eval()
    prov = SynthProvider(cfg, seed=seed)
    sd = model.state_dict()
    used = set()
    with torch.no_grad():
        for name, t in sd.items():
            key = name.replace("model.vision_embed_tokens.wte.", "model.embed_tokens.")
            if key in prov:
                t.copy_(prov(key))
                used.add(key)
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
    print(f"  reference model built in {time.time() - t0:.1f}s, "
          f"{sum(p.numel() for p in model.parameters()) / 1e6:.1f} M params", flush=True)
    return model


def sample(t: torch.Tensor, n: int = 2048) -> dict:
    f = t.detach().float().flatten()
    stride = max(1, f.numel() // n)
    return {"shape": list(t.shape), "stride": stride, "mean": f.mean().item(), "abs_mean": f.abs().mean().item(),
            "vals": f[::stride][:n].clone()}


CASES = {
    # name: (cfg overrides, batches [(tag, batch, image sizes list, seq_len)])
    "slim_gpm": (dict(num_layers=2, clip_layers=2), [("c", 2, [(336, 672), (672, 336)], None),
                                                     ("r", 2, [(336, 672), (672, 336)], None)]),
    "slim_bt": (dict(num_layers=2, clip_layers=2, is_general_preference=False, add_cross_attention=False,
                     use_lora=False), [("c", 2, [(336, 672), (672, 336)], None),
                                       ("r", 2, [(336, 672), (672, 336)], None)]),
    "full_bt": (dict(is_general_preference=False, add_cross_attention=False, use_lora=False),
                [("c", 1, [(1344, 1344)], None), ("r", 1, [(1344, 1344)], None)]),
    "full_gpm": (dict(), [("c", 1, [(1008, 1344)], 2048), ("r", 1, [(1008, 1344)], 2048)]),
    # sequences beyond original_max_position_embeddings (4096): the su-RoPE switches to the long factors for the WHOLE
    # batch (seq_len = max(position_ids) + 1 over the batch, modeling_phi3_v.py:446-451), incl. the short sample
    "slim_bt_long": (dict(num_layers=2, clip_layers=2, is_general_preference=False, add_cross_attention=False,
                          use_lora=False), [("c", 2, [(1344, 1344), (336, 336)], None, (1600, 1700)),
                                            ("r", 2, [(1344, 1344), (336, 336)], None, (1600, 1700))]),
}
SEED_W, SEED_X = 1234, 7
ATTR_CASES = ("slim_gpm", "slim_bt")
ATTR_VARIANTS = {"layer_id_1": {"layer_id": 1}, "layer_id_0": {"layer_id": 0}, "training": {"training": True},
                 "mean": {"mean_hidden_state": True}, "mean_layer_id_1": {"mean_hidden_state": True, "layer_id": 1}}


def run_case(name: str, refmods):
    from llava_reward_b200.config import RewardConfig
    from llava_reward_b200.synth import synth_batch

    over, batches = CASES[name]
    cfg = RewardConfig(**over)
    print(f"[{name}] building", flush=True)
    model = build_reference_model(cfg, SEED_W, refmods)
    _, _, ral = refmods
    args = types.SimpleNamespace(is_general_preference=cfg.is_general_preference,
                                 value_head_dim=cfg.value_head_dim,
                                 general_preference_tau=cfg.general_preference_tau)
    fixture = {"case": name, "cfg_overrides": over, "seed_w": SEED_W, "seed_x": SEED_X, "batches": [],
               "torch": torch.__version__}
    rewards = {}
    for tag, B, hw_list, seq_len, *rest in batches:
        tlr = rest[0] if rest else (40, 128)
        ids, mask, pix, sizes = synth_batch(cfg, B, hw_list[0], seq_len, seed=SEED_X, tag=tag, image_hw_list=hw_list,
                                            text_len_range=tlr)
        t0 = time.time()
        with torch.no_grad():
            reward, out = model.custom_forward(ids, mask, pix, sizes, return_output=True)
        dt = time.time() - t0
        print(f"  batch {tag}: S={ids.shape[1]} reward={reward.flatten().tolist()} ({dt:.1f}s)", flush=True)
        hs = out["hidden_states"]
        entry = {"tag": tag, "batch": B, "image_hw": hw_list, "seq_len": seq_len, "S": ids.shape[1],
                 "text_len_range": list(tlr),
                 "seconds": dt, "reward": reward.float().clone(),
                 "taps": {"inputs_embeds": sample(hs[0]), "hidden_0": sample(hs[1]),
                          "last_hidden": sample(out["last_hidden_state"]), "vision_embeds": sample(hs[-1])}}
        # exact small slices (last valid row of first/last sample) for tighter checks
        entry["last_hidden_eos"] = out["last_hidden_state"][:, -1, :64].float().clone()
        if name in ATTR_CASES:
            # the attributes custom_forward honours (rw_model_general_preference.py:327-333, 349-352, 398-448), set on
            # the reference model object exactly as a caller would; dropout probabilities are 0 in these configs, so
            # the top-level `training` flag only switches the gather rule
            entry["attrs"] = {}
            for key, attrs in ATTR_VARIANTS.items():
                saved = {k: getattr(model, k) for k in attrs}
                for k, v in attrs.items():
                    setattr(model, k, v)
                with torch.no_grad():
                    r2, _ = model.custom_forward(ids, mask, pix, sizes)
                for k, v in saved.items():
                    setattr(model, k, v)
                entry["attrs"][key] = r2.float().clone()
                print(f"    {key}: {r2.flatten().tolist()}", flush=True)
        fixture["batches"].append(entry)
        rewards[tag] = reward
    prob = ral.preference_compute(args, rewards["c"], rewards["r"])
    fixture["prob"] = torch.from_numpy(prob).clone()
    print(f"  prob={prob.tolist()}", flush=True)
    path = os.path.join(OUT, f"{name}.pt")
    torch.save(fixture, path)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)", flush=True)
    meta = {k: v for k, v in fixture.items() if k in ("case", "cfg_overrides", "seed_w", "seed_x", "torch")}
    meta["rewards"] = {t: rewards[t].flatten().tolist() for t in rewards}
    meta["prob"] = prob.tolist()
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    mods = import_reference()
    for case in sys.argv[1:]:
        run_case(case, mods)
