"""The fp32 verification path (north_star: rewards within 1e-4 of the reference's fp32): the engine's dataflow - token
plan, packed layouts (LoRA K-extension, head-interleaved q/k rows, [gate|up] blocks), HD gather, embedding scatter,
packed-row decoder, last-layer row shortcut, EOS-row SkipCA head - with every floating-point kernel in plain fp32
(csrc/f32_verify.cu), against the goldens the UNMODIFIED reference produced in fp32 on the CPU. No bf16 rounding is
involved on either side, so the gate is the plain 1e-4 and decisions must be identical."""
import os
import types

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

from golden_util import fixture_batch, fixture_cfg, load_fixture  # noqa: E402

from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402

TOL = 1e-4
_models = {}


def build_f32(fx, tmp_path_factory):
    key = fx["case"]
    if key not in _models:
        cfg = fixture_cfg(fx)
        d = tmp_path_factory.mktemp(key + "_f32")
        ypath = os.path.join(d, "reward_config.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"is_general_preference": cfg.is_general_preference,
                            "add_cross_attention": cfg.add_cross_attention, "value_head_dim": cfg.value_head_dim,
                            "general_preference_tau": cfg.general_preference_tau}, f)
        args = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None,
                                     ft_projector=False, precision="fp32",
                                     config_overrides={k: v for k, v in fx["cfg_overrides"].items()
                                                       if k in ("num_layers", "clip_layers", "use_lora")})
        args, model = load_reward_adaptor(args, "phi3v", ypath)
        _models.clear()
        _models[key] = (args, model.to("cuda").eval(), cfg)
    return _models[key]


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt", "slim_bt_long", "slim_gpm_b1"])
def test_fp32_path_matches_reference_fp32(case, tmp_path_factory):
    fx = load_fixture(case)
    args, model, cfg = build_f32(fx, tmp_path_factory)
    assert model.engine.precision == "fp32" and model.engine.w.embed.dtype == torch.float32
    rewards, worst = {}, 0.0
    for entry in fx["batches"]:
        ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
        r, _ = model.custom_forward(ids, mask, pix, sizes)
        assert r.dtype == torch.float32 and tuple(r.shape) == tuple(entry["reward"].shape)
        err = (r.cpu() - entry["reward"]).abs().max().item()
        print(f"{case}/{entry['tag']}: fp32 path {r.flatten().tolist()} reference {entry['reward'].flatten().tolist()} "
              f"|d| {err:.3g}")
        worst = max(worst, err)
        rewards[entry["tag"]] = r
    assert worst <= TOL, f"{case}: {worst:.3g}"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    ref = fx["prob"].numpy()
    assert abs(prob - ref).max() <= 5e-4
    assert ((prob > 0.5) == (ref > 0.5)).all()


def test_fp32_path_layout_switches_are_exact(tmp_path_factory):
    """packed valid rows on / off and the last-layer row shortcut on / off: the same fp32 arithmetic per row"""
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_f32(fx, tmp_path_factory)
    eng = model.engine
    ids, mask, pix, sizes = fixture_batch(fx, fx["batches"][0], cfg, device="cuda")
    base, _ = model.custom_forward(ids, mask, pix, sizes)
    try:
        eng.pack_rows = False
        a, _ = model.custom_forward(ids, mask, pix, sizes)
        eng.last_layer_rows = False
        b, _ = model.custom_forward(ids, mask, pix, sizes)
    finally:
        eng.pack_rows = eng.last_layer_rows = True
    assert (a - base).abs().max().item() <= 2e-6 and (b - base).abs().max().item() <= 2e-6


@pytest.mark.parametrize("case", ["full_gpm", "full_bt"])
def test_fp32_path_full_depth(case, tmp_path_factory):
    """23 CLIP + 32 decoder layers (4.1 - 4.4 G parameters in fp32 = 17.5 GB): BASELINE configs[1] / configs[0] shapes"""
    fx = load_fixture(case)
    args, model, cfg = build_f32(fx, tmp_path_factory)
    worst = 0.0
    rewards = {}
    for entry in fx["batches"]:
        ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
        r, _ = model.custom_forward(ids, mask, pix, sizes)
        err = (r.cpu() - entry["reward"]).abs().max().item()
        print(f"{case}/{entry['tag']}: fp32 path {r.flatten().tolist()} reference {entry['reward'].flatten().tolist()} "
              f"|d| {err:.3g}")
        worst = max(worst, err)
        rewards[entry["tag"]] = r
    assert worst <= TOL, f"{case}: {worst:.3g}"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    assert ((prob > 0.5) == (fx["prob"].numpy() > 0.5)).all()
    _models.clear()
    torch.cuda.empty_cache()
