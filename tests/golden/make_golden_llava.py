"""Golden vectors for the LLaVA-v1.6 branch: the UNMODIFIED reference `CustomRewardModel` (model_type='llava',
rw_model_general_preference.py:304-448) on top of the installed transformers `LlavaNextForConditionalGeneration`.

    python tests/golden/make_golden_llava.py llava_slim_bt llava_slim_gpm      # ~1 min each
    python tests/golden/make_golden_llava.py llava_wide_bt                     # 7B-width decoder layers, full CLIP

Same recipe as make_golden.py (stub modules for deepspeed/peft/accelerate, deterministic hash weights, LoRA attached
by the 10-line peft restatement). transformers here is 5.5 (the reference pins 4.50): parameter names moved
(`language_model.model.*` -> `model.language_model.*` etc.); the arithmetic of the path is the same - the name map
below is the only adaptation. This script only runs in the build container.
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
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.dirname(os.path.abspath(__file__))

from make_golden import LoraWrapped, import_reference, sample  # noqa: E402

SEED_W, SEED_X = 1234, 7

ATTR_VARIANTS = {"training": {"training": True}, "mean": {"mean_hidden_state": True}}
ATTR_CASES = {"llava_slim_bt": ("training", "mean"), "llava_slim_gpm": ("training", "mean")}
CASES = {
    # name: (cfg overrides, batches [(tag, original image sizes, seq_len, padding_side)])
    "llava_slim_bt": (dict(hidden_size=512, intermediate_size=1024, num_heads=4, num_layers=2, clip_layers=2),
                      [("c", [(480, 640), (500, 333)], None, "left"), ("r", [(300, 900), (672, 672)], None, "left")]),
    "llava_slim_gpm": (dict(hidden_size=512, intermediate_size=1024, num_heads=4, num_layers=2, clip_layers=2,
                            is_general_preference=True),
                       [("c", [(900, 300), (336, 336)], None, "left"), ("r", [(640, 480), (1000, 700)], None, "right")]),
    "llava_wide_bt": (dict(num_layers=2), [("c", [(480, 640)], None, "left"), ("r", [(672, 672)], None, "left")]),
}


def to_hf_name(name: str) -> str:
    """reference-era (transformers 4.50) parameter name -> installed transformers 5.x name"""
    if name.startswith("language_model.model."):
        return "model.language_model." + name[len("language_model.model."):]
    if name.startswith(("vision_tower.", "multi_modal_projector.", "image_newline")):
        return "model." + name
    return name


def build_reference_model(cfg, seed, refmods):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaNextConfig, LlavaNextForConditionalGeneration
    from llava_reward_b200.synth import SynthProvider

    _get_reward_model = refmods[0]
    from llava_reward.utils import LlamaRMSNorm

    vcfg = CLIPVisionConfig(hidden_size=cfg.clip_hidden, intermediate_size=cfg.clip_intermediate,
                            num_hidden_layers=cfg.clip_layers + 1, num_attention_heads=cfg.clip_heads,
                            image_size=cfg.image_size, patch_size=cfg.patch, projection_dim=768,
                            hidden_act="quick_gelu", layer_norm_eps=cfg.clip_eps)
    tcfg = LlamaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                       num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                       num_key_value_heads=cfg.num_heads, rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta,
                       max_position_embeddings=4096, hidden_act="silu", attention_bias=False, mlp_bias=False)
    rcfg = LlavaNextConfig(vision_config=vcfg, text_config=tcfg, image_token_index=cfg.image_token_id,
                           image_grid_pinpoints=cfg.image_grid_pinpoints, vision_feature_layer=-2,
                           vision_feature_select_strategy="default", projector_hidden_act="gelu")
    rcfg.use_cache = False
    rcfg._attn_implementation = "eager"
    cls = _get_reward_model(LlavaNextForConditionalGeneration, LlavaNextForConditionalGeneration,
                            is_general_preference=cfg.is_general_preference, add_cross_attention=False,
                            value_head_dim=cfg.value_head_dim, RMSNorm_class=LlamaRMSNorm, RMSNorm_class_eps=1e-5,
                            model_type="llava")
    t0 = time.time()
    model = cls(rcfg)
    model.eval()
    prov = SynthProvider(cfg, seed=seed)
    sd = model.state_dict()
    used = set()
    with torch.no_grad():
        for name in prov.names():
            if ".lora_" in name:
                continue
            sd[to_hf_name(name)].copy_(prov(name))
            used.add(name)
    if cfg.use_lora:
        layers = model.model.language_model.layers
        for i, layer in enumerate(layers):
            for holder, attrs in ((layer.self_attn, ("q_proj", "k_proj", "v_proj", "o_proj")),
                                  (layer.mlp, ("gate_proj", "up_proj", "down_proj"))):
                sub = "self_attn" if holder is layer.self_attn else "mlp"
                for attr in attrs:
                    p = f"language_model.model.layers.{i}.{sub}.{attr}"
                    setattr(holder, attr, LoraWrapped(getattr(holder, attr), prov(p + ".lora_A.weight"),
                                                      prov(p + ".lora_B.weight"), cfg.lora_scale))
    print(f"  reference model built in {time.time() - t0:.1f}s, "
          f"{sum(p.numel() for p in model.parameters()) / 1e6:.1f} M params", flush=True)
    return model


def run_case(name, refmods):
    from llava_reward_b200.config import LlavaNextRewardConfig
    from llava_reward_b200.synth import synth_batch_llava

    over, batches = CASES[name]
    cfg = LlavaNextRewardConfig(**over)
    print(f"[{name}] building", flush=True)
    model = build_reference_model(cfg, SEED_W, refmods)
    ral = refmods[2]
    args = types.SimpleNamespace(is_general_preference=cfg.is_general_preference, value_head_dim=cfg.value_head_dim,
                                 general_preference_tau=cfg.general_preference_tau)
    fixture = {"case": name, "cfg_overrides": over, "seed_w": SEED_W, "seed_x": SEED_X, "batches": [],
               "torch": torch.__version__}
    rewards = {}
    for tag, hw_list, seq_len, side in batches:
        batch = synth_batch_llava(cfg, len(hw_list), hw_list, seq_len, seed=SEED_X, tag=tag, padding_side=side)
        t0 = time.time()
        with torch.no_grad():
            reward, out = model.custom_forward(inputs_batch=batch, return_output=True)
        dt = time.time() - t0
        print(f"  batch {tag}: S={batch['input_ids'].shape[1]} reward={reward.flatten().tolist()} ({dt:.1f}s)", flush=True)
        hs = out["hidden_states"]
        eos = batch["attention_mask"].shape[1] - 1 - batch["attention_mask"].flip(1).argmax(1)
        entry = {"tag": tag, "image_hw": hw_list, "seq_len": seq_len, "padding_side": side,
                 "S": batch["input_ids"].shape[1], "seconds": dt, "reward": reward.float().clone(),
                 "taps": {"inputs_embeds": sample(hs[0]), "hidden_0": sample(hs[1]), "last_hidden": sample(hs[-1])},
                 "last_hidden_eos": hs[-1][torch.arange(len(hw_list)), eos, :64].float().clone()}
        # attributes custom_forward reads (rw_model_general_preference.py:327-333, 398-448), set on the reference model
        entry["attrs"] = {}
        for key, attrs in ATTR_VARIANTS.items():
            if key not in ATTR_CASES.get(name, ()):
                continue
            saved = {k: getattr(model, k) for k in attrs}
            for k, v in attrs.items():
                setattr(model, k, v)
            with torch.no_grad():
                r2, _ = model.custom_forward(inputs_batch=batch)
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
