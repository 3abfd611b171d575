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
OUT = os.path.dirname(os.path.abspath(__file__))


from oracle.ref_harness import LoraWrapped, build_reference_model, import_reference  # noqa: E402,F401  (the harness shared with
# bench.py's reference arm and tests/test_reference_gpu.py)


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
CASES["slim_gpm_b1"] = (dict(num_layers=2, clip_layers=2), [("c", 1, [(336, 672)], None), ("r", 1, [(672, 336)], None)])
SEED_W, SEED_X = 1234, 7
ATTR_CASES = ("slim_gpm", "slim_bt", "slim_gpm_b1")
# `vision_layer_id` (rw_model_general_preference.py:353): SkipCA keys/values = hidden_states[id][:, :N_v_max] with
# hidden_states = (inputs_embeds, h_1, norm(h_2), vision_embeds) at 2 layers. Batches of ONE sample: no padded positions,
# whose hidden states differ between the reference's eager and flash-attention paths.
VL_VARIANTS = {"vision_layer_id_0": {"vision_layer_id": 0}, "vision_layer_id_1": {"vision_layer_id": 1},
               "vision_layer_id_2": {"vision_layer_id": 2}, "vision_layer_id_-2": {"vision_layer_id": -2},
               "vision_layer_id_-3": {"vision_layer_id": -3}}
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
            for key, attrs in (VL_VARIANTS if name == "slim_gpm_b1" else ATTR_VARIANTS).items():
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
