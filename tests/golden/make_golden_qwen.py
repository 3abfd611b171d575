"""Golden vectors for the Qwen2.5-VL branch: the reference `CustomRewardModel` (model_type='qwen',
rw_model_general_preference.py:304-448) on top of the installed transformers `Qwen2_5_VLForConditionalGeneration`.

    python tests/golden/make_golden_qwen.py qwen_slim_bt qwen_slim_gpm       # ~1 min each
    python tests/golden/make_golden_qwen.py qwen_wide_bt                     # 7B-width layers (2 ViT + 2 decoder)

Same recipe as make_golden.py / make_golden_llava.py (stub modules for deepspeed/peft/accelerate, deterministic hash
weights, LoRA attached by the 10-line peft restatement). transformers here is 5.5 (the reference pins 4.50), under which
the reference class does not construct as written. Two HARNESS shims, neither touching arithmetic:
  1. `config.hidden_size` (read at rw_model_general_preference.py:313) is set from `config.text_config.hidden_size`
     (4.50 kept the text fields on the top-level config);
  2. `CustomRewardModel.visual` (read at :356 for the redundant extra vision pass) is aliased to `self.model.visual`
     (4.50 kept the tower at top level).
  3. M-RoPE positions: 4.50's `get_rope_index` gives an image's tokens the temporal position `start` (its
     second_per_grid_t is 0 for images); the installed 5.5 multiplies that start position by tokens_per_second
     (`position_temporal * time_interval`, modeling_qwen2_5_vl.py:1017), i.e. different numbers for the same input. The
     reference pins 4.50, so `rope_index_v450` below restates 4.50's algorithm (vision_start scan) and the positions
     are handed to the unmodified forward as `inputs_batch["position_ids"]`, which both releases accept verbatim.
Parameter names moved too (`visual.*` -> `model.visual.*`, `model.*` -> `model.language_model.*`); `to_hf_name` is that
map. The installed processor also emits `mm_token_type_ids`, which 5.x needs to compute the M-RoPE positions that 4.50
derived from input_ids alone (get_rope_index); synth_batch_qwen provides it. This script only runs in the build container.
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

SLIM = dict(hidden_size=512, intermediate_size=1024, num_heads=4, num_kv_heads=2, num_layers=2, vit_depth=4,
            vit_hidden=640, vit_intermediate=856, vit_heads=8, vit_fullatt=[1, 3], vocab_size=152064)
ATTR_VARIANTS = {"training": {"training": True}, "mean": {"mean_hidden_state": True}}
ATTR_CASES = {"qwen_slim_bt": ("training", "mean"), "qwen_slim_gpm": ("training", "mean")}
CASES = {
    # name: (cfg overrides, batches [(tag, (h, w) patch grids, seq_len, padding_side)])
    "qwen_slim_bt": (dict(SLIM), [("c", [(16, 24), (22, 10)], None, "left"), ("r", [(8, 8), (34, 18)], None, "left")]),
    "qwen_slim_gpm": (dict(SLIM, is_general_preference=True, add_cross_attention=True),
                      [("c", [(20, 12), (16, 16)], None, "left"), ("r", [(12, 26), (10, 10)], None, "right")]),
    "qwen_wide_bt": (dict(num_layers=2, vit_depth=2, vit_fullatt=[1]),
                     [("c", [(32, 32)], None, "left"), ("r", [(22, 34)], None, "left")]),
    "qwen_wide_gpm": (dict(num_layers=2, vit_depth=2, vit_fullatt=[1], is_general_preference=True,
                           add_cross_attention=True),
                      [("c", [(32, 32), (16, 20)], None, "left"), ("r", [(22, 34), (32, 32)], None, "left")]),
}


def to_hf_name(name: str) -> str:
    """reference-era (transformers 4.50) parameter name -> installed transformers 5.x name"""
    if name.startswith("visual."):
        return "model." + name
    if name.startswith("model."):
        return "model.language_model." + name[len("model."):]
    return name


def rope_index_v450(cfg, input_ids, attention_mask, image_grid_thw):
    """Qwen2_5_VLForConditionalGeneration.get_rope_index of transformers 4.50 (images only), restated as written there:
    scan for <|vision_start|>, text run, then (t, h, w) indices + text_len + st_idx; padded positions stay 1."""
    m = cfg.vit_merge
    position_ids = torch.ones(3, input_ids.shape[0], input_ids.shape[1], dtype=input_ids.dtype)
    image_index = 0
    for i, row in enumerate(input_ids):
        row = row[attention_mask[i] == 1]
        starts = torch.argwhere(row == cfg.vision_start_token_id).squeeze(1)
        image_nums = int((row[starts + 1] == cfg.image_token_id).sum())
        tokens = row.tolist()
        lst, st = [], 0
        for _ in range(image_nums):
            ed = tokens.index(cfg.image_token_id, st)
            t, h, w = (int(v) for v in image_grid_thw[image_index])
            image_index += 1
            gt, gh, gw = t, h // m, w // m
            text_len = ed - st
            st_idx = int(lst[-1].max()) + 1 if lst else 0
            lst.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
            t_index = (torch.arange(gt).view(-1, 1).expand(-1, gh * gw) * 0 * 2).long().flatten()
            h_index = torch.arange(gh).view(1, -1, 1).expand(gt, -1, gw).flatten()
            w_index = torch.arange(gw).view(1, 1, -1).expand(gt, gh, -1).flatten()
            lst.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
            st = ed + gt * gh * gw
        if st < len(tokens):
            st_idx = int(lst[-1].max()) + 1 if lst else 0
            lst.append(torch.arange(len(tokens) - st).view(1, -1).expand(3, -1) + st_idx)
        position_ids[:, i, attention_mask[i] == 1] = torch.cat(lst, dim=1).reshape(3, -1)
    return position_ids


def build_reference_model(cfg, seed, refmods):
    from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration, Qwen2_5_VLModel
    from llava_reward_b200.synth import SynthProvider

    _get_reward_model = refmods[0]
    from llava_reward.utils import Qwen2RMSNorm

    rcfg = Qwen2_5_VLConfig(
        text_config=dict(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                         num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
                         num_key_value_heads=cfg.num_kv_heads, rms_norm_eps=cfg.rms_eps, max_position_embeddings=128000,
                         rope_parameters={"rope_type": "default", "rope_theta": cfg.rope_theta,
                                          "mrope_section": list(cfg.mrope_section)},
                         use_sliding_window=False, tie_word_embeddings=False, hidden_act="silu"),
        vision_config=dict(depth=cfg.vit_depth, hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate,
                           num_heads=cfg.vit_heads, patch_size=cfg.vit_patch, spatial_merge_size=cfg.vit_merge,
                           temporal_patch_size=cfg.vit_temporal_patch, window_size=cfg.vit_window,
                           fullatt_block_indexes=list(cfg.vit_fullatt), out_hidden_size=cfg.hidden_size,
                           hidden_act="silu", in_channels=3, tokens_per_second=2),
        image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
        vision_start_token_id=cfg.vision_start_token_id, vision_end_token_id=cfg.vision_end_token_id)
    rcfg.use_cache = False
    rcfg._attn_implementation = "eager"
    rcfg.text_config._attn_implementation = "eager"
    rcfg.vision_config._attn_implementation = "eager"
    rcfg.hidden_size = rcfg.text_config.hidden_size                       # harness shim 1
    cls = _get_reward_model(Qwen2_5_VLForConditionalGeneration, Qwen2_5_VLModel,
                            is_general_preference=cfg.is_general_preference,
                            add_cross_attention=cfg.add_cross_attention, value_head_dim=cfg.value_head_dim,
                            RMSNorm_class=Qwen2RMSNorm, RMSNorm_class_eps=1e-6)
    cls.visual = property(lambda self: self.model.visual)                  # harness shim 2
    t0 = time.time()
    model = cls(rcfg)
    model.model_type = "qwen"
    model.eval()
    prov = SynthProvider(cfg, seed=seed)
    sd = model.state_dict()
    with torch.no_grad():
        for name in prov.names():
            if ".lora_" in name:
                continue
            sd[to_hf_name(name)].copy_(prov(name))
    if cfg.use_lora:
        for i, layer in enumerate(model.model.language_model.layers):
            for holder, sub, attrs in ((layer.self_attn, "self_attn", ("q_proj", "k_proj", "v_proj", "o_proj")),
                                       (layer.mlp, "mlp", ("gate_proj", "up_proj", "down_proj"))):
                for attr in attrs:
                    p = f"model.layers.{i}.{sub}.{attr}"
                    setattr(holder, attr, LoraWrapped(getattr(holder, attr), prov(p + ".lora_A.weight"),
                                                      prov(p + ".lora_B.weight"), cfg.lora_scale))
    print(f"  reference model built in {time.time() - t0:.1f}s, "
          f"{sum(p.numel() for p in model.parameters()) / 1e6:.1f} M params", flush=True)
    return model


def run_case(name, refmods):
    from transformers.feature_extraction_utils import BatchFeature
    from llava_reward_b200.config import QwenVLRewardConfig
    from llava_reward_b200.synth import synth_batch_qwen

    over, batches = CASES[name]
    cfg = QwenVLRewardConfig(**over)
    print(f"[{name}] building", flush=True)
    model = build_reference_model(cfg, SEED_W, refmods)
    ral = refmods[2]
    args = types.SimpleNamespace(is_general_preference=cfg.is_general_preference, value_head_dim=cfg.value_head_dim,
                                 general_preference_tau=cfg.general_preference_tau)
    fixture = {"case": name, "cfg_overrides": over, "seed_w": SEED_W, "seed_x": SEED_X, "batches": [],
               "torch": torch.__version__}
    rewards = {}
    for tag, grids, seq_len, side in batches:
        batch = BatchFeature(synth_batch_qwen(cfg, grids, seq_len, seed=SEED_X, tag=tag, padding_side=side))
        batch["position_ids"] = rope_index_v450(cfg, batch["input_ids"], batch["attention_mask"], batch["image_grid_thw"])
        t0 = time.time()
        with torch.no_grad():
            reward, out = model.custom_forward(inputs_batch=batch, return_output=True)
            vis = model.model.visual(batch["pixel_values"], grid_thw=batch["image_grid_thw"]).pooler_output
        dt = time.time() - t0
        print(f"  batch {tag}: S={batch['input_ids'].shape[1]} reward={reward.flatten().tolist()} ({dt:.1f}s)", flush=True)
        hs = out["hidden_states"]
        eos = batch["attention_mask"].shape[1] - 1 - batch["attention_mask"].flip(1).argmax(1)
        entry = {"tag": tag, "grids": grids, "seq_len": seq_len, "padding_side": side,
                 "S": batch["input_ids"].shape[1], "seconds": dt, "reward": reward.float().clone(),
                 "taps": {"image_embeds": sample(vis), "inputs_embeds": sample(hs[0]), "hidden_0": sample(hs[1]),
                          "last_hidden": sample(hs[-1])},
                 "last_hidden_eos": hs[-1][torch.arange(len(grids)), eos, :64].float().clone(), "n_hidden": len(hs)}
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
