"""Golden vectors for BASELINE.json configs[0] AS WRITTEN: the two JPEGs of reference eval/simple_inference.py:22
(copied with the rest of data/sample_test to tests/golden/sample_test/) decoded by PIL, through the reference's own
`Phi3VImageProcessor(num_crops=16)` (processing_phi3_v.py:208-288) and the UNMODIFIED reference model
(`custom_forward`, BT head, no SkipCA, no LoRA, fp32 on the CPU, eager attention) + `preference_compute`.

    python tests/golden/make_golden_real.py real_slim_bt        # 2 CLIP + 2 decoder layers, a minute
    python tests/golden/make_golden_real.py real_full_bt        # configs[0]: 23 + 32 layers, 4146.6 M params, ~10 min, 20 GB

The tokenizer of Phi-3.5-vision is not available offline, so the caption is 80 seeded token ids
(torch.randint(3, 31999, (80,), Generator().manual_seed(7))) inside the reference's prompt frame
[bos, <|user|>, \\n, -1 x N_v, \\n, caption..., eos]. Runs only in the build container.
"""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
from oracle.ref_harness import build_reference_model, import_reference  # noqa: E402

IMAGES = ["sample_test/sample_img/0_1_id_000904-0035.jpg", "sample_test/sample_img/4_3_id_000904-0035.jpg"]
BT = dict(is_general_preference=False, add_cross_attention=False, use_lora=False)
CASES = {"real_slim_bt": dict(num_layers=2, clip_layers=2, **BT), "real_full_bt": dict(**BT)}
SEED_W = 1234


def caption_ids():
    return torch.randint(3, 31999, (80,), generator=torch.Generator().manual_seed(7)).tolist()


def prompt_ids(n_img_tokens: int):
    from llava_reward_b200.synth import BOS, EOS, NL, USER
    return [BOS, USER, NL] + [-1] * n_img_tokens + [NL] + caption_ids() + [EOS]


def run_case(name):
    from PIL import Image

    from llava_reward_b200.config import RewardConfig
    refmods = import_reference()
    from llava_reward.models.base_mllm.phi3_v.processing_phi3_v import Phi3VImageProcessor
    over = CASES[name]
    cfg = RewardConfig(**over)
    proc = Phi3VImageProcessor(num_crops=16)
    model = build_reference_model(cfg, SEED_W, refmods)
    ral = refmods[2]
    args = types.SimpleNamespace(is_general_preference=False, value_head_dim=1, general_preference_tau=cfg.general_preference_tau)
    fx = {"case": name, "cfg_overrides": over, "seed_w": SEED_W, "images": IMAGES, "samples": [], "torch": torch.__version__}
    rewards = []
    for rel in IMAGES:
        img = Image.open(os.path.join(OUT, rel)).convert("RGB")
        out = proc.preprocess(img, return_tensors="pt")
        pv, sizes, ntok = out["pixel_values"], out["image_sizes"], int(out["num_img_tokens"][0])
        ids = torch.tensor([prompt_ids(ntok)], dtype=torch.int64)
        mask = torch.ones_like(ids)
        with torch.no_grad():
            r, _ = model.custom_forward(ids, mask, pv, sizes)
        print(f"  {rel}: size {img.size} -> image_sizes {sizes.tolist()} N_v {ntok} S {ids.shape[1]} reward {r.flatten().tolist()}",
              flush=True)
        fx["samples"].append({"image": rel, "pil_size": list(img.size), "image_sizes": sizes[0].tolist(), "num_img_tokens": ntok,
                              "S": ids.shape[1], "reward": r.float().clone(), "pixel_sum": pv.double().sum().item(),
                              "pixel_abs_sum": pv.double().abs().sum().item(), "pixel_sample": pv.flatten()[::997].clone()})
        rewards.append(r)
    prob = ral.preference_compute(args, rewards[0], rewards[1])
    fx["prob"] = torch.from_numpy(prob).clone()
    print(f"  prob {prob.tolist()}")
    torch.save(fx, os.path.join(OUT, f"{name}.pt"))
    with open(os.path.join(OUT, f"{name}.json"), "w") as f:
        json.dump({"case": name, "cfg_overrides": over, "images": IMAGES, "rewards": [float(r) for r in rewards],
                   "prob": prob.tolist(), "S": [s["S"] for s in fx["samples"]]}, f, indent=1)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for c in sys.argv[1:]:
        print(f"[{c}]", flush=True)
        run_case(c)
