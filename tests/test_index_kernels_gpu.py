"""Bit-exact tests of the index / gather kernels of the Phi-3.5-V headline path and per-kernel tests of the
SkipCA head, each against the oracle restatement (oracle/reward_oracle.py) on the SAME bf16 tensors.

  lr_hd_gather_bf16      vs oracle.hd_feature_rows   (modeling_phi3_v.py:254-362)      torch.equal
  lr_embed_scatter_bf16  vs oracle.embed_tokens      (modeling_phi3_v.py:228-252)      torch.equal
  lr_skipca_scores/_head vs oracle.skipca + value head in fp32 (rw_model_general_preference.py:376-386, 407-448)
  lr_synth_normal_f32    vs the torch-on-CPU counter hash (synth.hash_normal)          torch.equal
"""
import math
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402
from llava_reward_b200.config import RewardConfig, num_image_tokens  # noqa: E402
from llava_reward_b200.synth import hash_normal  # noqa: E402
from oracle import reward_oracle as O  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16
# every crop geometry from 1x1 to 4x4 that fits 16 crops, tall, wide and square, in ragged batches
SIZE_SETS = [
    [(336, 336), (1344, 1344), (336, 1344), (1344, 336)],
    [(672, 1008), (1008, 672), (672, 672)],
    [(1008, 1344), (336, 672), (1344, 1008), (1008, 1008), (672, 336)],
    [(336, 5376 // 16 * 16)],   # 1 x 16 crops: the widest legal layout
]


def rnd(*shape, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(bf).to(DEV)


def make_plan(sizes):
    """[B, PLAN_STRIDE] int32 plan rows (hc, wc, crop_base, row_base, nv) exactly as RewardEngine.forward builds them"""
    plan = torch.zeros(len(sizes), L.PLAN_STRIDE, dtype=torch.int32)
    crop_base = row_base = 0
    for b, (h, w) in enumerate(sizes):
        hc, wc = h // 336, w // 336
        nv = num_image_tokens(h, w)
        plan[b, :5] = torch.tensor([hc, wc, crop_base, row_base, nv], dtype=torch.int32)
        crop_base += hc * wc + 1
        row_base += nv
    return plan.to(DEV), crop_base, row_base


def fake_params(tensors):
    return O.Params(lambda name: tensors[name], dtype=bf, device=DEV)


@pytest.mark.parametrize("sizes", SIZE_SETS)
def test_hd_gather_is_bit_exact(sizes):
    cfg = RewardConfig()
    D = cfg.clip_hidden
    B = len(sizes)
    plan, n_crops, sum_nv = make_plan(sizes)
    clip = rnd(n_crops * 577, D, seed=len(sizes))                     # CLIP tokens incl. the CLS row of every crop
    sub_gn, glb_gn = rnd(4 * D, seed=11), rnd(4 * D, seed=12)
    rows = torch.full((sum_nv, 4 * D), float("nan"), dtype=bf, device=DEV)
    max_nv = int(plan[:, L.PLAN_NV].max())
    ops.hd_gather(clip, plan.view(-1), sub_gn, glb_gn, rows, B, max_nv)
    # oracle input layout: [B, 17, 576, 1024] with the global crop in slot 0 and zero-padded slots
    feats = torch.zeros(B, 17, 576, D, dtype=bf, device=DEV)
    c3 = clip.view(n_crops, 577, D)
    for b in range(B):
        hc, wc, cb = int(plan[b, 0]), int(plan[b, 1]), int(plan[b, 2])
        feats[b, : hc * wc + 1] = c3[cb: cb + hc * wc + 1, 1:]
    P = fake_params({O.VE + "sub_GN": sub_gn.view(1, 1, 1, -1), O.VE + "glb_GN": glb_gn.view(1, 1, -1)})
    ref, counts = O.hd_feature_rows(P, cfg, feats, sizes)
    assert counts == [num_image_tokens(h, w) for h, w in sizes]
    assert ref.shape == rows.shape and torch.equal(rows, ref)


@pytest.mark.parametrize("padding", ["left", "right"])
@pytest.mark.parametrize("sizes", SIZE_SETS[:3])
def test_embed_scatter_is_bit_exact(sizes, padding):
    cfg = RewardConfig()
    H, V, B = cfg.hidden_size, cfg.vocab_size, len(sizes)
    nvs = [num_image_tokens(h, w) for h, w in sizes]
    text = [17 + 29 * b for b in range(B)]
    lens = [3 + nv + 1 + t + 1 for nv, t in zip(nvs, text)]
    S = max(lens) + 5
    g = torch.Generator().manual_seed(3)
    ids = torch.full((B, S), 32000, dtype=torch.int64)
    mask = torch.zeros(B, S, dtype=torch.int64)
    for b in range(B):
        row = [1, 32010, 13] + [-1] * nvs[b] + [13] + torch.randint(3, V, (text[b],), generator=g).tolist() + [32000]
        lo = S - len(row) if padding == "left" else 0
        ids[b, lo: lo + len(row)] = torch.tensor(row)
        mask[b, lo: lo + len(row)] = 1
    ids, mask = ids.to(DEV), mask.to(DEV)
    i32 = dict(dtype=torch.int32, device=DEV)
    pos, ordn = torch.empty(B * S, **i32), torch.empty(B * S, **i32)
    ss, sl, er, ni = (torch.zeros(B, **i32) for _ in range(4))
    fl = torch.zeros(1, **i32)
    ops.token_plan(ids, mask, B, S, pos, ordn, ss, sl, er, ni, fl)
    assert ni.tolist() == nvs and fl.item() == 0
    plan, _, sum_nv = make_plan(sizes)
    wte = rnd(V, H, std=0.02, seed=5)
    img = rnd(sum_nv, H, seed=6)
    hid = torch.full((B * S, H), float("nan"), dtype=bf, device=DEV)
    ops.embed_scatter(ids, ordn, plan.view(-1), wte, img, hid, B, S, H, V)
    P = fake_params({"model.embed_tokens.weight": wte})
    ref_hidden, ref_vis = O.embed_tokens(P, cfg, ids, img)
    assert torch.equal(hid.view(B, S, H), ref_hidden)
    # vision_embeds (zero-padded to the batch maximum, modeling_phi3_v.py:242-246) = the rows SkipCA reads through the plan
    for b in range(B):
        rb = int(plan[b, L.PLAN_ROW_BASE])
        assert torch.equal(img[rb: rb + nvs[b]], ref_vis[b, : nvs[b]]) and not ref_vis[b, nvs[b]:].any()


@pytest.mark.parametrize("sizes", [[(1008, 1344), (1008, 1344)], [(336, 336), (1344, 1344), (672, 1008)]])
@pytest.mark.parametrize("vhd", [1, 2])
def test_skipca_scores_and_head_vs_oracle_fp32(sizes, vhd):
    """K and V rows are given (bf16, as the W_k|W_v GEMM leaves them); with identity W_q/W_k/W_v the oracle's skipca()
    computes the same cross attention in fp32 from the same values - incl. the reference's zero-padded vision rows of
    the shorter samples, which take part in the softmax (maxN_v > N_v)."""
    cfg = RewardConfig(value_head_dim=vhd, is_general_preference=vhd > 1)
    H, B = cfg.hidden_size, len(sizes)
    plan, _, sum_nv = make_plan(sizes)
    nvs = [num_image_tokens(h, w) for h, w in sizes]
    max_nv = max(nvs)
    x = rnd(B, H, std=1.0, seed=21)                 # final-norm output of the EOS rows
    q = rnd(B, H, std=2.0, seed=22)
    kv = rnd(sum_nv, 2 * H, std=1.0, seed=23)
    ca_ln = (1 + 0.02 * torch.randn(H, generator=torch.Generator().manual_seed(24))).to(bf).to(DEV)
    vh = rnd(vhd, H, std=0.02, seed=25)
    scores = torch.full((B, max_nv), float("nan"), dtype=torch.float32, device=DEV)
    ops.skipca_scores(q, kv, plan.view(-1), scores, B, H, max_nv)
    reward = torch.empty(B, vhd, dtype=bf, device=DEV)
    ops.skipca_head(scores, kv, plan.view(-1), x, ca_ln, vh, reward, B, H, max_nv, vhd, cfg.rms_eps)
    # fp32 reference on the same values
    K = torch.zeros(B, max_nv, H, device=DEV)
    Vv = torch.zeros(B, max_nv, H, device=DEV)
    for b in range(B):
        rb = int(plan[b, L.PLAN_ROW_BASE])
        K[b, : nvs[b]] = kv[rb: rb + nvs[b], :H].float()
        Vv[b, : nvs[b]] = kv[rb: rb + nvs[b], H:].float()
    sc_ref = torch.einsum("bh,bjh->bj", q.float(), K) / math.sqrt(H)
    for b in range(B):
        assert (scores[b, nvs[b]:] == 0).all()       # zero K rows -> score exactly 0, still inside the softmax
    assert (scores - sc_ref).abs().max().item() <= 2 ** -7 * sc_ref.abs().max().item() + 1e-6   # two bf16 roundings
    w = torch.softmax(sc_ref, dim=-1)
    y = x.float() + torch.einsum("bj,bjh->bh", w, Vv)
    y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + cfg.rms_eps) * ca_ln.float()
    ref = y @ vh.float().t()
    err = (reward.float() - ref).abs().max().item()
    print(f"skipca head vs fp32: max err {err:.4g} (reward absmax {ref.abs().max().item():.3g})")
    assert err <= 1.5e-2 * max(1.0, ref.abs().max().item())
    # the same numbers through the oracle's own skipca() with identity projections (ties the formula to the cited lines)
    eye = torch.eye(H, device=DEV)
    tensors = {"W_q.weight": eye, "W_k.weight": eye, "W_v.weight": eye, "ca_layernorm.weight": ca_ln.float(),
               "value_head.weight": vh.float()}
    P = O.Params(lambda name: tensors[name], dtype=torch.float32, device=DEV)
    # oracle.skipca takes q from last_hidden itself; feed it the pair (last_hidden = x, K = V = vis) in a form where
    # q == x: run the kernels again with q := x, K := V := vis
    vis = torch.zeros(B, max_nv, H, dtype=bf, device=DEV)
    kv2 = torch.empty(sum_nv, 2 * H, dtype=bf, device=DEV)
    for b in range(B):
        rb = int(plan[b, L.PLAN_ROW_BASE])
        vis[b, : nvs[b]] = kv[rb: rb + nvs[b], :H]
        kv2[rb: rb + nvs[b], :H] = kv[rb: rb + nvs[b], :H]
        kv2[rb: rb + nvs[b], H:] = kv[rb: rb + nvs[b], :H]
    ops.skipca_scores(x, kv2, plan.view(-1), scores, B, H, max_nv)
    ops.skipca_head(scores, kv2, plan.view(-1), x, ca_ln, vh, reward, B, H, max_nv, vhd, cfg.rms_eps)
    ref2 = O.skipca(P, cfg, x.float()[:, None, :], vis.float())[:, 0] @ vh.float().t()
    err2 = (reward.float() - ref2).abs().max().item()
    print(f"skipca head vs oracle.skipca fp32: max err {err2:.4g}")
    assert err2 <= 1.5e-2 * max(1.0, ref2.abs().max().item())


def test_value_head_only_form():
    """add_cross_attention=False: reward = x . value_head^T (rw_model_general_preference.py:407-448), H up to 8192"""
    for H in (3072, 4096, 5120):
        x, vh = rnd(7, H, seed=H), rnd(2, H, std=0.02, seed=H + 1)
        reward = torch.empty(7, 2, dtype=bf, device=DEV)
        ops.skipca_head(None, None, None, x, None, vh, reward, 7, H, 0, 2, 1e-5)
        ref = x.float() @ vh.float().t()
        assert (reward.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("shape,mean", [((1000,), 0.0), ((3, 5, 7), 1.0), ((3072, 3072), 0.0), ((1 << 24) + 13, 1.0)])
def test_synth_generator_kernel_matches_cpu_hash(shape, mean):
    shape = shape if isinstance(shape, tuple) else (shape,)
    a = hash_normal("model.layers.3.mlp.down_proj.weight", shape, 0.02, 1234, device="cpu", mean=mean)
    b = hash_normal("model.layers.3.mlp.down_proj.weight", shape, 0.02, 1234, device=DEV, mean=mean)
    assert b.is_cuda and torch.equal(a, b.cpu())
