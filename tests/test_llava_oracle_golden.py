"""LLaVA-v1.6 branch: the oracle restatement must reproduce the reference's own outputs (fixtures made by
tests/golden/make_golden_llava.py, which executes /root/reference on the installed transformers LlavaNext).
fp32 CPU, tolerance 1e-4 on rewards."""
import pytest
import torch

from golden_util import llava_fixture_batch, llava_fixture_cfg, load_fixture, strided
from oracle import llava_next_oracle as O
from oracle.reward_oracle import Params, preference_compute
from llava_reward_b200.config import anyres_geometry
from llava_reward_b200.synth import SynthProvider

TOL = 1e-4


@pytest.mark.parametrize("case", ["llava_slim_bt", "llava_slim_gpm"])
def test_llava_oracle_matches_reference(case):
    fx = load_fixture(case)
    cfg = llava_fixture_cfg(fx)
    P = Params(SynthProvider(cfg, seed=fx["seed_w"]), dtype=torch.float32)
    rewards = {}
    for entry in fx["batches"]:
        batch = llava_fixture_batch(fx, entry, cfg)
        mask = batch["attention_mask"]
        assert batch["input_ids"].shape[1] == entry["S"]
        taps = {}
        with torch.no_grad():
            r = O.custom_forward(P, cfg, batch, taps)
        rewards[entry["tag"]] = r
        assert r.shape == entry["reward"].shape
        assert (r - entry["reward"]).abs().max().item() < TOL
        # `training` / `mean_hidden_state` set on the reference model object (rw_model_general_preference.py:327-333)
        for key, g in entry.get("attrs", {}).items():
            if key == "training" and entry["padding_side"] == "right":
                continue  # position S-1 is a padded row there: not defined behaviour (differs between attention paths)
            kw = {"training": dict(training=True), "mean": dict(mean_hidden_state=True)}[key]
            with torch.no_grad():
                r2 = O.custom_forward(P, cfg, batch, **kw)
            assert r2.shape == g.shape and (r2 - g).abs().max().item() < TOL, (key, r2, g)
        for k in ("inputs_embeds", "hidden_0", "last_hidden"):
            t, g = taps[k], entry["taps"][k]
            assert list(t.shape) == g["shape"], k
            valid = mask.bool()[:, :, None].expand_as(t)
            a = strided(torch.where(valid, t, torch.zeros_like(t)), g["stride"])
            shape_mask = strided(valid.float(), g["stride"])
            assert ((a - g["vals"] * shape_mask).abs().max().item()) < 2e-4, k
        eos = mask.shape[1] - 1 - mask.flip(1).argmax(1)
        mine = taps["last_hidden"][torch.arange(mask.shape[0]), eos, :64]
        assert (mine - entry["last_hidden_eos"]).abs().max().item() < 2e-4
    p = preference_compute(cfg, rewards["c"], rewards["r"])
    assert (p - fx["prob"]).abs().max().item() < 1e-3
    assert ((p > 0.5) == (fx["prob"] > 0.5)).all()


def test_anyres_geometry_matches_oracle_packing():
    """config.anyres_geometry (host planning of the product path) against the oracle's pack on random sizes."""
    from llava_reward_b200.config import LlavaNextRewardConfig
    cfg = LlavaNextRewardConfig()
    g = torch.Generator().manual_seed(3)
    for _ in range(200):
        h, w = (int(v) for v in torch.randint(40, 1400, (2,), generator=g))
        geo = anyres_geometry((h, w), cfg.image_grid_pinpoints)
        bh, bw = O.select_best_resolution((h, w), cfg.image_grid_pinpoints)
        assert (geo["grid_h"], geo["grid_w"]) == (bh // 336, bw // 336)
        fm = torch.zeros(1, geo["grid_h"] * 24, geo["grid_w"] * 24)
        un = O.unpad_image(fm, (h, w))
        assert (un.shape[1], un.shape[2]) == (geo["keep_h"], geo["keep_w"]), (h, w)
        assert geo["n_tokens"] == 576 + un.shape[1] * (un.shape[2] + 1)
