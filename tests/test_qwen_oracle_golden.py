"""Qwen2.5-VL branch: the oracle restatement must reproduce the reference's own outputs (fixtures made by
tests/golden/make_golden_qwen.py, which executes the reference class on the installed transformers Qwen2.5-VL).
fp32 CPU, tolerance 1e-4 on rewards."""
import numpy as np
import pytest
import torch

from golden_util import load_fixture, qwen_fixture_batch, qwen_fixture_cfg, strided
from oracle import qwen_vl_oracle as O
from oracle.reward_oracle import Params, preference_compute
from llava_reward_b200.config import QwenVLRewardConfig, qwen_window_plan
from llava_reward_b200.synth import SynthProvider

TOL = 1e-4


@pytest.mark.parametrize("case", ["qwen_slim_bt", "qwen_slim_gpm"])
def test_qwen_oracle_matches_reference(case):
    fx = load_fixture(case)
    cfg = qwen_fixture_cfg(fx)
    P = Params(SynthProvider(cfg, seed=fx["seed_w"]), dtype=torch.float32)
    rewards = {}
    for entry in fx["batches"]:
        batch = qwen_fixture_batch(fx, entry, cfg)
        mask = batch["attention_mask"]
        assert batch["input_ids"].shape[1] == entry["S"]
        assert entry["n_hidden"] == cfg.num_layers + 1
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
        g = entry["taps"]["image_embeds"]
        assert list(taps["image_embeds"].shape) == g["shape"]
        assert (strided(taps["image_embeds"], g["stride"]) - g["vals"]).abs().max().item() < 2e-4
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


def test_window_plan_matches_oracle():
    """config.qwen_window_plan (host planning of the product path, numpy) against the oracle's restatement of
    rot_pos_emb / get_window_index on assorted grids, incl. grids that are / are not multiples of the 8-patch window."""
    cfg = QwenVLRewardConfig()
    for grids in ([(1, 32, 32)], [(1, 22, 34), (1, 8, 8)], [(1, 16, 24), (1, 2, 2), (1, 36, 6)], [(2, 12, 20)]):
        pos, widx, cu_win, cu_img = O.window_plan(cfg, grids)
        plan = qwen_window_plan(grids, cfg.vit_merge, cfg.vit_window, cfg.vit_patch)
        assert np.array_equal(plan["window_index"], widx.numpy())
        assert plan["win_cu"].tolist() == cu_win and plan["img_cu"].tolist() == cu_img
        T = pos.shape[0]
        want = pos.reshape(T // 4, 4, 2)[widx].reshape(T, 2).numpy()
        assert np.array_equal(plan["pos_hw"], want)
        src = torch.arange(T).reshape(T // 4, 4)[widx].reshape(-1).numpy()
        assert np.array_equal(plan["src_row"], src)
        assert max(np.diff(plan["win_cu"])) <= 64
