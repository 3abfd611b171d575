"""Checkpoint round trip on the GPU: the synthetic weights are written to disk in the layouts the
reference consumes - a HF checkpoint directory + the `save_model_lora` output with the reference's key names
(llava_reward/utils/deepspeed.py:344-347, 391-398; eval/reward_adaptor_loader.py:46-60, 80-105, 124-148) - loaded back
through `load_reward_adaptor(pretrain=dir, pm_path=dir, ft_projector=True)` and scored: rewards must be BIT-identical
to the `synthetic:` path, for all three backbones. The base checkpoint holds a negated projector, so the round trip
only succeeds if the fine-tuned projector of pytorch_model.bin overrides it as the reference does."""
import os
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from golden_util import (fixture_batch, fixture_cfg, llava_fixture_batch, llava_fixture_cfg, load_fixture,  # noqa: E402
                         qwen_fixture_batch, qwen_fixture_cfg)
from write_checkpoint import write_reference_layout  # noqa: E402

from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor  # noqa: E402

CASES = {
    "phi3v": ("slim_gpm", fixture_cfg, fixture_batch),
    "llava": ("llava_slim_gpm", llava_fixture_cfg, llava_fixture_batch),
    "qwen": ("qwen_slim_gpm", qwen_fixture_cfg, qwen_fixture_batch),
}
# depth of the vision towers is not part of a HF config.json the loaders read for phi3v / llava (fixed 24-layer CLIP)
DEPTH_KEYS = ("clip_layers",)


def forward(model, model_type, batch):
    if model_type == "phi3v":
        return model.custom_forward(*batch)[0]
    return model.custom_forward(inputs_batch={k: v.to("cuda") for k, v in batch.items()})[0]


@pytest.mark.parametrize("model_type", ["phi3v", "llava", "qwen"])
def test_checkpoint_round_trip_is_bit_identical(model_type, tmp_path):
    case, mk_cfg, mk_batch = CASES[model_type]
    fx = load_fixture(case)
    cfg = mk_cfg(fx)
    assert cfg.use_lora
    base, pm = str(tmp_path / "base"), str(tmp_path / "pm")
    write_reference_layout(cfg, model_type, fx["seed_w"], base, pm, device="cuda")
    ypath = os.path.join(pm, "reward_config.yaml")
    over_all = {k: v for k, v in fx["cfg_overrides"].items() if k not in ("is_general_preference", "add_cross_attention")}
    a_syn = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None, ft_projector=False,
                                  config_overrides=over_all)
    _, m_syn = load_reward_adaptor(a_syn, model_type, ypath)
    a_ckpt = types.SimpleNamespace(pretrain=base, pm_path=pm, cache_dir=None, ft_projector=True,
                                   config_overrides={k: v for k, v in over_all.items() if k in DEPTH_KEYS})
    a_ckpt, m_ckpt = load_reward_adaptor(a_ckpt, model_type, ypath)
    assert a_ckpt.is_general_preference == cfg.is_general_preference and a_ckpt.value_head_dim == cfg.value_head_dim
    c = m_ckpt.config
    assert (c.hidden_size, c.num_layers, c.use_lora, c.lora_rank) == (cfg.hidden_size, cfg.num_layers, True, cfg.lora_rank)
    m_syn, m_ckpt = m_syn.to("cuda").eval(), m_ckpt.to("cuda").eval()
    for entry in fx["batches"]:
        batch = mk_batch(fx, entry, cfg, device="cuda")
        r_syn = forward(m_syn, model_type, batch)
        r_ckpt = forward(m_ckpt, model_type, batch)
        assert torch.equal(r_syn, r_ckpt), (r_syn, r_ckpt)
        assert (r_ckpt.float().cpu() - entry["reward"]).abs().max().item() < 0.1   # and it is the golden model
    # without ft_projector the (negated) base projector is used: the reward must change
    a_np = types.SimpleNamespace(pretrain=base, pm_path=pm, cache_dir=None, ft_projector=False,
                                 config_overrides=a_ckpt.config_overrides)
    _, m_np = load_reward_adaptor(a_np, model_type, ypath)
    batch = mk_batch(fx, fx["batches"][0], cfg, device="cuda")
    assert not torch.equal(forward(m_np.to("cuda").eval(), model_type, batch), forward(m_syn, model_type, batch))


def test_lora_rank_16_adapter_runs_and_matches_padded_math(tmp_path):
    """an adapter of rank 16 (not a multiple of the GEMM's 128 granularity): zero-padded at pack time, same rewards as the
    same adapter zero-padded to rank 128 on disk with alpha scaled to keep alpha/r"""
    fx = load_fixture("slim_gpm")
    cfg16 = fixture_cfg(fx)
    cfg16.lora_rank, cfg16.lora_alpha = 16, 32.0
    base, pm = str(tmp_path / "base"), str(tmp_path / "pm")
    write_reference_layout(cfg16, "phi3v", fx["seed_w"], base, pm, device="cuda", decoy_projector=False)
    ypath = os.path.join(pm, "reward_config.yaml")
    over = {"clip_layers": cfg16.clip_layers}
    a16 = types.SimpleNamespace(pretrain=base, pm_path=pm, cache_dir=None, ft_projector=False, config_overrides=over)
    _, m16 = load_reward_adaptor(a16, "phi3v", ypath)
    assert m16.config.lora_rank == 16
    # the same adapter padded on disk
    ad = torch.load(os.path.join(pm, "lora", "adapter_model.bin"))
    pad = {}
    for k, v in ad.items():
        if ".lora_A." in k:
            pad[k] = torch.cat([v, torch.zeros(112, v.shape[1], dtype=v.dtype)], 0)
        else:
            pad[k] = torch.cat([v, torch.zeros(v.shape[0], 112, dtype=v.dtype)], 1)
    pm2 = str(tmp_path / "pm128")
    os.makedirs(os.path.join(pm2, "lora"))
    torch.save(pad, os.path.join(pm2, "lora", "adapter_model.bin"))
    import json
    import shutil
    with open(os.path.join(pm2, "lora", "adapter_config.json"), "w") as f:
        json.dump({"r": 128, "lora_alpha": 256.0}, f)
    shutil.copy(os.path.join(pm, "pytorch_model.bin"), os.path.join(pm2, "pytorch_model.bin"))
    a128 = types.SimpleNamespace(pretrain=base, pm_path=pm2, cache_dir=None, ft_projector=False, config_overrides=over)
    _, m128 = load_reward_adaptor(a128, "phi3v", ypath)
    batch = fixture_batch(fx, fx["batches"][0], fixture_cfg(fx), device="cuda")
    r16 = m16.to("cuda").eval().custom_forward(*batch)[0]
    r128 = m128.to("cuda").eval().custom_forward(*batch)[0]
    assert torch.equal(r16, r128)
