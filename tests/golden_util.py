"""Helpers shared by the golden-vector tests."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    path = os.path.join(GOLDEN_DIR, f"{name}.pt")
    return torch.load(path, weights_only=False)


def fixture_cfg(fx):
    from llava_reward_b200.config import RewardConfig
    return RewardConfig(**fx["cfg_overrides"])


def fixture_batch(fx, entry, cfg, device="cpu"):
    from llava_reward_b200.synth import synth_batch
    hw = [tuple(x) for x in entry["image_hw"]]
    return synth_batch(cfg, entry["batch"], hw[0], entry["seq_len"], seed=fx["seed_x"], tag=entry["tag"],
                       image_hw_list=hw, device=device, text_len_range=tuple(entry.get("text_len_range", (40, 128))))


def strided(t, stride, n=2048):
    return t.detach().float().flatten()[::stride][:n].cpu()


def llava_fixture_cfg(fx):
    from llava_reward_b200.config import LlavaNextRewardConfig
    return LlavaNextRewardConfig(**fx["cfg_overrides"])


def llava_fixture_batch(fx, entry, cfg, device="cpu"):
    from llava_reward_b200.synth import synth_batch_llava
    hw = [tuple(x) for x in entry["image_hw"]]
    return synth_batch_llava(cfg, len(hw), hw, entry["seq_len"], seed=fx["seed_x"], tag=entry["tag"],
                             padding_side=entry["padding_side"], device=device)


def qwen_fixture_cfg(fx):
    from llava_reward_b200.config import QwenVLRewardConfig
    return QwenVLRewardConfig(**fx["cfg_overrides"])


def qwen_fixture_batch(fx, entry, cfg, device="cpu"):
    from llava_reward_b200.synth import synth_batch_qwen
    grids = [tuple(x) for x in entry["grids"]]
    return synth_batch_qwen(cfg, grids, entry["seq_len"], seed=fx["seed_x"], tag=entry["tag"],
                            padding_side=entry["padding_side"], device=device)
