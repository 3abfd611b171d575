"""Qwen2.5-VL branch on the B200: kernels added for it (packed varlen / grouped-query tcgen05 attention, bias + RoPE
GEMM epilogues with bf16 and fp32 tables, bias + SwiGLU epilogue, patch-row packing, M-RoPE plan, row compaction, SkipCA
with masked pad scores) and the engine end to end against
 (a) the reference's own fp32 outputs (tests/golden/qwen_*.pt, made by tests/golden/make_golden_qwen.py),
 (b) the oracle restatement in bf16 on the same GPU, stage by stage.
Tolerance (north_star): rewards within 2e-2 absolute in bf16 (+ the measured bf16 noise of the reference arithmetic)."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F
import yaml

pytestmark = pytest.mark.gpu

from golden_util import load_fixture, qwen_fixture_batch, qwen_fixture_cfg  # noqa: E402
from test_kernels_gpu import attn_ref, bfr, check_close, rnd  # noqa: E402
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402
from llava_reward_b200.config import QwenVLRewardConfig, qwen_window_plan  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from llava_reward_b200.synth import SynthProvider, synth_batch_qwen  # noqa: E402
from oracle import qwen_vl_oracle as O  # noqa: E402
from oracle.reward_oracle import Params  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16
REWARD_TOL = 2e-2
_models = {}


# ----------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("lens", [[64, 64, 16, 64, 8, 4, 48], [1024], [700, 64, 257, 128, 129], [4], [64] * 40])
def test_attention_packed_hd96_noncausal(lens):
    """packed variable-length sequences (windows / whole images of the vision tower), head_dim 80 padded to 96:
    rows of a sequence see only that sequence; nothing outside [base, base+len) of a sequence's tile is written"""
    heads, hd, hdp = 3, 80, 96
    T = sum(lens)
    AW = heads * hdp
    qkv = rnd(T, 3 * AW, seed=7)
    qkv.view(T, 3, heads, hdp)[..., hd:] = 0           # the zero padding the packed weights produce
    o = torch.full((T + 64, AW), float("nan"), dtype=bf, device=DEV)
    base = torch.tensor(np.concatenate([[0], np.cumsum(lens)[:-1]]), dtype=torch.int32, device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    scale = hd ** -0.5
    ops.attention_ex(qkv, qkv[:, AW:], qkv[:, 2 * AW:], o, 3 * AW, AW, T, len(lens), max(lens), base, None, ln, heads,
                     heads, hdp, False, scale)
    torch.cuda.synchronize()
    assert torch.isnan(o[T:].float()).all()            # rows past the last sequence untouched
    f = qkv.float().view(T, 3, heads, hdp)
    b0 = 0
    for n in lens:
        ref = attn_ref(f[b0:b0 + n, 0], f[b0:b0 + n, 1], f[b0:b0 + n, 2], False, scale).reshape(n, AW)
        check_close(o[b0:b0 + n], ref, f"packed attention seq at {b0}", atol=1e-2, rtol=2e-2)
        b0 += n
    assert (o[:T].view(T, heads, hdp)[..., hd:] == 0).all()   # padded head columns stay exactly zero


@pytest.mark.parametrize("lens", [[64, 64, 16, 64, 8, 4, 48], [64] * 9 + [48] * 3 + [36], [4], [128, 64, 1, 127, 64],
                                  [48, 36] * 40 + [64] * 33])
def test_attention_segments_hd96(lens):
    """segment layout: 128-row tiles spanning several windows, every row attends to its own window only - same
    results as the one-sequence-per-window packed launch (bit-identical inputs, same per-row arithmetic up to the
    order of the masked-out zero terms) and as the fp32 reference"""
    heads, hd, hdp = 3, 80, 96
    T = sum(lens)
    AW = heads * hdp
    qkv = rnd(T, 3 * AW, seed=9)
    qkv.view(T, 3, heads, hdp)[..., hd:] = 0
    cu = np.concatenate([[0], np.cumsum(lens)])
    lo = torch.tensor(np.repeat(cu[:-1], lens), dtype=torch.int32, device=DEV)
    hi = torch.tensor(np.repeat(cu[1:], lens), dtype=torch.int32, device=DEV)
    o = torch.full((T + 128, AW), float("nan"), dtype=bf, device=DEV)
    scale = hd ** -0.5
    ops.attention_seg(qkv, qkv[:, AW:], qkv[:, 2 * AW:], o, 3 * AW, AW, T, lo, hi, heads, hdp, scale)
    o2 = torch.full((T, AW), float("nan"), dtype=bf, device=DEV)
    base = torch.tensor(cu[:-1], dtype=torch.int32, device=DEV)
    ln = torch.tensor(lens, dtype=torch.int32, device=DEV)
    ops.attention_ex(qkv, qkv[:, AW:], qkv[:, 2 * AW:], o2, 3 * AW, AW, T, len(lens), max(lens), base, None, ln, heads,
                     heads, hdp, False, scale)
    torch.cuda.synchronize()
    assert torch.isnan(o[T:].float()).all()
    f = qkv.float().view(T, 3, heads, hdp)
    b0 = 0
    for n in lens:
        ref = attn_ref(f[b0:b0 + n, 0], f[b0:b0 + n, 1], f[b0:b0 + n, 2], False, scale).reshape(n, AW)
        check_close(o[b0:b0 + n], ref, f"segment attention at {b0}", atol=1e-2, rtol=2e-2)
        b0 += n
    assert (o[:T].float() - o2.float()).abs().max().item() < 8e-3


@pytest.mark.parametrize("heads,kvh,T,nseq", [(4, 2, 130, 3), (28, 4, 700, 2), (8, 1, 257, 2), (4, 4, 300, 1)])
def test_attention_gqa_hd128(heads, kvh, T, nseq):
    hd = 128
    H, kvw = heads * hd, kvh * hd
    QW = H + 2 * kvw
    qkv = rnd(nseq * T, QW, seed=5)
    o = torch.full((nseq * T, H + 128), float("nan"), dtype=bf, device=DEV)
    lens = [T, T - 37, 5][:nseq]
    starts = [T - n for n in lens]
    if nseq > 1:
        starts[1] = 0
    ss = torch.tensor(starts, dtype=torch.int32, device=DEV)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    scale = hd ** -0.5
    ops.attention_ex(qkv, qkv[:, H:], qkv[:, H + kvw:], o, QW, H + 128, nseq * T, nseq, T, None, ss, sl, heads, kvh, hd,
                     True, scale)
    torch.cuda.synchronize()
    f = qkv.float().view(nseq, T, QW)
    g = heads // kvh
    for s in range(nseq):
        q = f[s, :, :H].view(T, heads, hd)
        k = f[s, :, H:H + kvw].view(T, kvh, hd).repeat_interleave(g, dim=1)
        v = f[s, :, H + kvw:].view(T, kvh, hd).repeat_interleave(g, dim=1)
        ref = attn_ref(q, k, v, True, scale, starts[s], lens[s]).reshape(T, H)
        check_close(o[s * T:(s + 1) * T, :H], ref, f"gqa attention seq {s}", atol=1e-2, rtol=2e-2)


@pytest.mark.parametrize("M", [600, 64])
def test_gemm_bias_swiglu(M):
    N, K = 512, 320                                      # N = 2 * padded intermediate, packed [gate128 | up128]
    x = rnd(M, K, seed=1)
    W = rnd(N, K, std=K ** -0.5, seed=2)
    b = rnd(N, std=0.5, seed=3)
    out = torch.full((M, N // 2), float("nan"), dtype=bf, device=DEV)
    ops.gemm(x, W, out, M, N, K, L.EPI_BIAS_SWIGLU, b)
    torch.cuda.synchronize()
    acc = (x.float() @ W.float().t() + b.float()).view(M, N // 256, 2, 128)
    gate, up = bfr(acc[:, :, 0]), bfr(acc[:, :, 1])
    check_close(out, (up * bfr(F.silu(gate))).reshape(M, N // 2), "bias swiglu")


def test_gemm_bias_rope_bf16_tables_per_token():
    """q/k/v projection with bias + rotary (bf16 per-token tables, position_ids NULL): equals bias GEMM followed by
    lr_rope_su_bf16 with an identity position list, bit for bit (GQA layout: q 4 heads, k/v 2 heads)"""
    M, heads, kvh, hd, K = 520, 4, 2, 128, 576
    H, kvw = heads * hd, kvh * hd
    N = H + 2 * kvw
    x = rnd(M, K, seed=31)
    W = rnd(N, K, std=K ** -0.5, seed=32)
    b = rnd(N, std=0.3, seed=33)
    ang = torch.rand(M, hd // 2, device=DEV) * 50
    cos, sin = ang.cos().to(bf).contiguous(), ang.sin().to(bf).contiguous()
    # reference: plain bias GEMM, then rope on the q heads and (separately) the k heads
    ref = torch.empty(M, N, dtype=bf, device=DEV)
    ops.gemm(x, W, ref, M, N, K, L.EPI_BIAS, b)
    c, s = cos.float()[:, None, :], sin.float()[:, None, :]

    def rope(t, n):
        t = t.float().view(M, n, hd)
        x1, x2 = t[..., : hd // 2], t[..., hd // 2:]
        o1 = bfr(bfr(x1 * c) + bfr(-x2 * s))
        o2 = bfr(bfr(x2 * c) + bfr(x1 * s))
        return torch.cat([o1, o2], -1).reshape(M, n * hd)

    want = torch.cat([rope(ref[:, :H], heads), rope(ref[:, H:H + kvw], kvh), ref[:, H + kvw:].float()], 1)
    inter = torch.stack([torch.arange(hd // 2), torch.arange(hd // 2) + hd // 2], dim=1).reshape(-1)
    qk_perm = (torch.arange(heads + kvh)[:, None] * hd + inter[None, :]).reshape(-1)
    perm = torch.cat([qk_perm, torch.arange(H + kvw, N)]).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    ops.gemm_rope_ex(x, W[perm].contiguous(), out, M, N, K, b[perm].contiguous(), None, cos, sin, H + kvw, hd,
                     L.EPI_BIAS_ROPE)
    torch.cuda.synchronize()
    unperm = torch.empty_like(out)
    unperm[:, perm] = out
    assert torch.equal(unperm.float(), want)


def test_gemm_bias_rope_f32_tables_padded_heads():
    """vision qkv: head_dim 80 padded to 96, fp32 tables, single rounding (apply_rotary_pos_emb_vision)"""
    from llava_reward_b200.weights import qwen_vit_padded_head_dim
    M, heads, hd, K = 777, 8, 80, 640
    hdp = qwen_vit_padded_head_dim(hd)
    D = heads * hd
    x = rnd(M, K, seed=41)
    W = rnd(3 * D, K, std=K ** -0.5, seed=42)
    b = rnd(3 * D, std=0.3, seed=43)
    ang = torch.rand(M, hd // 2, device=DEV) * 30
    lin = bfr(x.float() @ W.float().t() + b.float()).view(M, 3, heads, hd)
    c = torch.cat([ang.cos(), ang.cos()], -1)[:, None, :]
    s = torch.cat([ang.sin(), ang.sin()], -1)[:, None, :]

    def rot(t):
        return torch.cat([-t[..., hd // 2:], t[..., : hd // 2]], -1)

    q = bfr(lin[:, 0] * c + rot(lin[:, 0]) * s)
    k = bfr(lin[:, 1] * c + rot(lin[:, 1]) * s)
    v = lin[:, 2]
    # packed layout
    half = hd // 2
    inter = torch.stack([torch.arange(half), torch.arange(half) + half], 1).reshape(-1)
    qk_head = torch.cat([inter, torch.full((hdp - hd,), -1, dtype=torch.long)])
    v_head = torch.cat([torch.arange(hd), torch.full((hdp - hd,), -1, dtype=torch.long)])

    def hm(per, base):
        m = torch.arange(heads)[:, None] * hd + per[None, :]
        return torch.where(per[None, :] < 0, torch.full_like(m, -1), m + base).reshape(-1)

    idx = torch.cat([hm(qk_head, 0), hm(qk_head, D), hm(v_head, 2 * D)]).to(DEV)
    Wp = W[idx.clamp(min=0)].clone()
    Wp[idx < 0] = 0
    bp = b[idx.clamp(min=0)].clone()
    bp[idx < 0] = 0
    cos = torch.ones(M, hdp // 2, device=DEV)
    sin = torch.zeros(M, hdp // 2, device=DEV)
    cos[:, :half], sin[:, :half] = ang.cos(), ang.sin()
    N = 3 * heads * hdp
    out = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    ops.gemm_rope_ex(x, Wp.contiguous(), out, M, N, K, bp.contiguous(), None, cos.contiguous(), sin.contiguous(),
                     2 * heads * hdp, hdp, L.EPI_BIAS_ROPE_F32)
    torch.cuda.synchronize()
    got = out.float().view(M, 3, heads, hdp)
    assert (got[..., hd:] == 0).all()
    # un-interleave q/k
    gq = torch.empty(M, heads, hd, device=DEV)
    gk = torch.empty(M, heads, hd, device=DEV)
    gq[..., inter.to(DEV)] = got[:, 0, :, :hd]
    gk[..., inter.to(DEV)] = got[:, 1, :, :hd]
    check_close(gq, q, "vision rope q", atol=2e-2, rtol=1e-2)
    check_close(gk, k, "vision rope k", atol=2e-2, rtol=1e-2)
    check_close(got[:, 2, :, :hd], v, "vision v", atol=2e-2, rtol=1e-2)
    # the rotation itself is exact given the same rounded linear output: compare against the kernel's own v-style
    # path by feeding identity tables
    one, zero = torch.ones_like(cos), torch.zeros_like(sin)
    out2 = torch.empty_like(out)
    ops.gemm_rope_ex(x, Wp.contiguous(), out2, M, N, K, bp.contiguous(), None, one, zero, 2 * heads * hdp, hdp,
                     L.EPI_BIAS_ROPE_F32)
    lin2 = out2.float().view(M, 3, heads, hdp)
    x1, x2 = lin2[:, 0, :, 0:hd:2], lin2[:, 0, :, 1:hd:2]
    cc, ss = ang.cos()[:, None, :], ang.sin()[:, None, :]
    assert torch.equal(got[:, 0, :, 0:hd:2], bfr(x1 * cc + (-x2) * ss))
    assert torch.equal(got[:, 0, :, 1:hd:2], bfr(x2 * cc + x1 * ss))


def test_patch_rows_gather_and_pad():
    T, K, Kp = 300, 1176, 1216
    pix = torch.randn(T, K, device=DEV)
    src = torch.randperm(T, device=DEV).to(torch.int32)
    out = torch.full((T, Kp), float("nan"), dtype=bf, device=DEV)
    ops.patch_rows(pix, src, out, T, K, Kp)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :K], pix[src.long()].to(bf))
    assert (out[:, K:] == 0).all()


@pytest.mark.parametrize("side", ["left", "right"])
def test_mrope_plan_matches_oracle(side):
    """positions bit-exact vs the oracle's get_rope_index restatement; per-token tables = rows of the by-position tables"""
    cfg = QwenVLRewardConfig()
    grids = [(16, 24), (22, 10), (2, 2), (34, 18)]
    batch = synth_batch_qwen(cfg, grids, None, seed=5, padding_side=side)
    ids, mask = batch["input_ids"].to(DEV), batch["attention_mask"].to(DEV)
    B, S = ids.shape
    grid = batch["image_grid_thw"].to(DEV, torch.int32).contiguous()
    half, n_pos = cfg.head_dim // 2, S + 8
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, device=DEV).float() / cfg.head_dim))
    ang = torch.arange(n_pos, device=DEV).float()[:, None] * inv[None]
    cos, sin = ang.cos().to(bf).contiguous(), ang.sin().to(bf).contiguous()
    i32 = dict(dtype=torch.int32, device=DEV)
    rc, fl = torch.zeros(B, **i32), torch.zeros(1, **i32)
    pos3 = torch.full((3, B * S), -7, **i32)
    ct = torch.empty(B * S, half, dtype=bf, device=DEV)
    st = torch.empty(B * S, half, dtype=bf, device=DEV)
    ops.mrope_plan(ids, mask, B, S, cfg.image_token_id, grid, B, cfg.vit_merge, rc, cos, sin, n_pos, half,
                   cfg.mrope_section[0], cfg.mrope_section[1], pos3, ct, st, fl)
    torch.cuda.synchronize()
    assert fl.item() == 0 and rc.cpu().tolist() == [1] * B
    ref = O.rope_index(cfg, batch["input_ids"], batch["attention_mask"], batch["image_grid_thw"].tolist())
    assert torch.equal(pos3.view(3, B, S).cpu().long(), ref)
    rc_, rs_ = O.mrope_cos_sin(cfg, ref.to(DEV), bf)
    assert torch.equal(ct.view(B, S, half), rc_[..., :half]) and torch.equal(st.view(B, S, half), rs_[..., :half])
    # a grid that disagrees with the token run is flagged
    bad = grid.clone()
    bad[1, 1] += 2
    ops.mrope_plan(ids, mask, B, S, cfg.image_token_id, bad, B, cfg.vit_merge, rc, cos, sin, n_pos, half,
                   cfg.mrope_section[0], cfg.mrope_section[1], pos3, ct, st, fl)
    assert fl.item() & 2


def test_compact_rows_and_masked_skipca():
    """lr_compact_rows_bf16 + lr_skipca_scores_ex(pad = bf16(-1e4)) + lr_skipca_head == the reference's qwen SkipCA arm
    on the last-valid-token row (incl. a sample without any pad row)."""
    B, S, H, vhd = 3, 40, 512, 2
    hid0 = rnd(B * S, H, seed=3)
    ids = torch.randint(5, 1000, (B, S))
    npad = [6, 0, 11]
    for b in range(B):
        ids[b, :npad[b]] = 151643
    ids_d = ids.to(DEV)
    mask = (ids_d != 151643).long()
    i32 = dict(dtype=torch.int32, device=DEV)
    pos, ordn = torch.empty(B * S, **i32), torch.empty(B * S, **i32)
    ss, sl, er, ni = (torch.zeros(B, **i32) for _ in range(4))
    fl = torch.zeros(1, **i32)
    ops.token_plan_ex(ids_d, mask, B, S, 151643, L.POS_ARANGE, pos, ordn, ss, sl, er, ni, fl)
    assert ni.cpu().tolist() == npad
    plan = torch.zeros(B, L.PLAN_STRIDE, dtype=torch.int32)
    plan[:, L.PLAN_NV] = torch.tensor(npad)
    plan[:, L.PLAN_ROW_BASE] = torch.tensor([0, 6, 6])
    plan_d = plan.to(DEV).view(-1)
    src = torch.full((sum(npad), H), float("nan"), dtype=bf, device=DEV)
    ops.compact_rows(hid0, ordn, plan_d, src, B, S, H)
    torch.cuda.synchronize()
    want = torch.cat([hid0.view(B, S, H)[b, :npad[b]] for b in range(B)])
    assert torch.equal(src, want)
    # SkipCA on one row per sample
    xe = rnd(B, H, seed=4)
    wq, wk, wv = (rnd(H, H, std=H ** -0.5, seed=10 + i) for i in range(3))
    ln_w = (1 + 0.02 * torch.randn(H)).to(bf).to(DEV)
    vh = rnd(vhd, H, std=H ** -0.5, seed=20)
    q = torch.empty(B, H, dtype=bf, device=DEV)
    ops.gemm(xe, wq, q, B, H, H)
    kv = torch.empty(sum(npad), 2 * H, dtype=bf, device=DEV)
    ops.gemm(src, torch.cat([wk, wv], 0).contiguous(), kv, sum(npad), 2 * H, H)
    mx = max(npad)
    scores = torch.empty(B, mx, dtype=torch.float32, device=DEV)
    ops.skipca_scores_ex(q, kv, plan_d, scores, B, H, mx, -9984.0)
    reward = torch.empty(B, vhd, dtype=bf, device=DEV)
    ops.skipca_head(scores, kv, plan_d, xe, ln_w, vh, reward, B, H, mx, vhd, 1e-6)
    torch.cuda.synchronize()
    # reference arm in bf16 torch ops
    vp = torch.zeros(B, mx, H, dtype=bf, device=DEV)
    pm = torch.ones(B, mx, dtype=torch.bool, device=DEV)
    for b in range(B):
        vp[b, :npad[b]] = hid0.view(B, S, H)[b, :npad[b]]
        pm[b, :npad[b]] = False
    Q = F.linear(xe[:, None], wq)
    K_, V_ = F.linear(vp, wk), F.linear(vp, wv)
    sc = (torch.bmm(Q, K_.transpose(1, 2)) / (H ** 0.5)).masked_fill(pm[:, None], -1e4)
    o = torch.bmm(F.softmax(sc, dim=-1), V_)
    y = (xe[:, None] + o).float()
    y = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + 1e-6)).to(bf) * ln_w
    ref = F.linear(y, vh)[:, 0]
    check_close(reward, ref, "qwen skipca head", atol=2e-2, rtol=2e-2)


# ----------------------------------------------------------------------------------------------- engine
def build_model(fx, tmp_path_factory):
    key = fx["case"]
    if key not in _models:
        cfg = qwen_fixture_cfg(fx)
        d = tmp_path_factory.mktemp(key)
        ypath = os.path.join(d, "reward_config.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"is_general_preference": cfg.is_general_preference,
                            "add_cross_attention": cfg.add_cross_attention, "value_head_dim": cfg.value_head_dim,
                            "general_preference_tau": cfg.general_preference_tau}, f)
        over = {k: v for k, v in fx["cfg_overrides"].items() if k not in ("is_general_preference", "add_cross_attention")}
        args = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None,
                                     ft_projector=False, disable_fast_tokenizer=False, config_overrides=over)
        args, model = load_reward_adaptor(args, "qwen", ypath)
        _models[key] = (args, model.to("cuda").eval(), cfg)
    return _models[key]


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def bf16_floor(fx, cfg, batch, ref_fp32):
    P = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV, cache=False)
    outs = []
    with torch.no_grad():
        outs.append(O.custom_forward(P, cfg, batch).float().cpu())
        old = torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction
        torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = not old
        try:
            outs.append(O.custom_forward(P, cfg, batch).float().cpu())
        finally:
            torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = old
    d = torch.stack([(r - ref_fp32).abs() for r in outs])
    return d.max().item(), d.pow(2).mean().sqrt().item()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("case", ["qwen_slim_bt", "qwen_slim_gpm", "qwen_wide_bt", "qwen_wide_gpm"])
def test_qwen_vs_reference_golden(case, tmp_path_factory):
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    rewards, errs, floors = {}, [], []
    for entry in fx["batches"]:
        batch = to_dev(qwen_fixture_batch(fx, entry, cfg))
        r, _ = model.custom_forward(inputs_batch=batch)
        assert r.dtype == bf and r.is_cuda and tuple(r.shape) == tuple(entry["reward"].shape)
        errs.append((r.float().cpu() - entry["reward"]).abs().max().item())
        floors.append(bf16_floor(fx, cfg, batch, entry["reward"]))
        print(f"{case}/{entry['tag']}: engine {r.flatten().tolist()} ref {entry['reward'].flatten().tolist()} | "
              f"engine-vs-fp32 {errs[-1]:.4g}, reference bf16-vs-fp32 max {floors[-1][0]:.4g} rms {floors[-1][1]:.4g} "
              f"launches {model.engine.launches}")
        rewards[entry["tag"]] = r
    mx = max(f[0] for f in floors)
    rms = (sum(f[1] ** 2 for f in floors) / len(floors)) ** 0.5
    assert max(errs) < REWARD_TOL + max(mx, 3 * rms), f"{case}: reward err {max(errs):.4g} vs reference fp32"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    ref = fx["prob"].numpy()
    decided = abs(ref - 0.5) > 0.05
    assert ((prob > 0.5) == (ref > 0.5))[decided].all()
    print(f"{case}: prob engine {prob.tolist()} reference {ref.tolist()}")


@pytest.mark.parametrize("case", ["qwen_slim_bt", "qwen_slim_gpm"])
def test_qwen_stages_vs_oracle(case, tmp_path_factory):
    """Stage by stage: engine (bf16 kernels) and the oracle in bf16 on this GPU, both against the oracle in fp32 -
    the engine may not be further from fp32 than 1.5x the reference arithmetic in bf16 (+2e-3)."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][1]
    batch = to_dev(qwen_fixture_batch(fx, entry, cfg))
    ids, mask = batch["input_ids"], batch["attention_mask"]
    B, S = ids.shape
    P32 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=torch.float32, device=DEV)
    P16 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV)
    t32, t16 = {}, {}
    with torch.no_grad():
        O.custom_forward(P32, cfg, batch, t32)
        O.custom_forward(P16, cfg, batch, t16)
    model.engine.taps = {}
    model.custom_forward(inputs_batch=batch)
    te, model.engine.taps = model.engine.taps, None
    # M-RoPE positions: integers, bit-exact
    ref_pos = O.rope_index(cfg, ids.cpu(), mask.cpu(), batch["image_grid_thw"].tolist())
    assert torch.equal(te["pos3"].view(3, B, S).cpu().long(), ref_pos)
    valid = mask.bool()
    rows = [("vit_embed", te["vit_embed"], t16["vit_embed"], t32["vit_embed"]),
            ("vit_layer0", te["vit_layer0"], t16["vit_layer0"], t32["vit_layer0"]),
            ("vit_out", te["vit_out"], t16["vit_out"], t32["vit_out"]),
            ("image_embeds", te["image_embeds"], t16["image_embeds"], t32["image_embeds"]),
            ("inputs_embeds", te["inputs_embeds"].view(B, S, -1)[valid], t16["inputs_embeds"][valid],
             t32["inputs_embeds"][valid])]
    for i in range(cfg.num_layers):
        rows.append((f"hidden_{i}", te[f"hidden_{i}"].view(B, S, -1)[valid], t16[f"hidden_{i}"][valid],
                     t32[f"hidden_{i}"][valid]))
    eos = S - 1 - mask.flip(1).argmax(1)
    ar = torch.arange(B, device=DEV)
    rows.append(("last_hidden_eos", te["last_hidden_eos"][:B], t16["last_hidden"][ar, eos], t32["last_hidden"][ar, eos]))
    worst = 0.0
    for name, e, o16, o32 in rows:
        ee, eo = rel_err(e, o32), rel_err(o16, o32)
        print(f"  {name}: rel L2 err vs fp32: engine {ee:.4g}, reference bf16 {eo:.4g}")
        assert ee < 1.5 * eo + 2e-3, name
        worst = max(worst, ee / max(eo, 1e-9))
    print(f"{case}: worst engine/reference error ratio {worst:.3f}")
    # image rows of inputs_embeds are bit-exact copies of the merger output (original order)
    sel = ids == cfg.image_token_id
    assert torch.equal(te["inputs_embeds"].view(B, S, -1)[sel], te["image_embeds"])


def test_qwen_window_attention_layouts_agree(tmp_path_factory):
    """engine with the segment-layout window attention (product) vs one packed sequence per window"""
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(qwen_fixture_batch(fx, fx["batches"][1], cfg))
    assert model.engine.window_attn == "seg"
    r_seg, _ = model.custom_forward(inputs_batch=batch)
    model.engine.window_attn = "packed"
    try:
        r_pk, _ = model.custom_forward(inputs_batch=batch)
    finally:
        model.engine.window_attn = "seg"
    print(f"window attention: seg {r_seg.flatten().tolist()} packed {r_pk.flatten().tolist()}")
    assert (r_seg.float() - r_pk.float()).abs().max().item() < 1e-2


def test_qwen_last_layer_row_shortcut(tmp_path_factory):
    fx = load_fixture("qwen_slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(qwen_fixture_batch(fx, fx["batches"][0], cfg))
    r_short, _ = model.custom_forward(inputs_batch=batch)
    model.engine.last_layer_rows = False
    try:
        r_full, _ = model.custom_forward(inputs_batch=batch)
    finally:
        model.engine.last_layer_rows = True
    d = (r_short.float() - r_full.float()).abs().max().item()
    print(f"last-layer rows: shortcut {r_short.flatten().tolist()} full {r_full.flatten().tolist()} |d| {d:.3g}")
    assert d <= 4e-3


def test_qwen_two_images_in_one_sample(tmp_path_factory):
    """a sample with TWO images (two <|vision_start|> ... <|vision_end|> runs): M-RoPE restarts the grid positions at
    the running offset, image rows are scattered in order - engine vs the oracle (positions bit-exact, rewards within the
    bf16 noise). The reference's datasets put one image per sample; the path itself is general."""
    from llava_reward_b200.synth import hash_normal, hash_randint
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    grids = [(8, 12), (14, 6), (10, 10)]                      # sample 0: images 0 and 1; sample 1: image 2
    pix = torch.cat([hash_normal(f"mi.{i}", (h * w, cfg.patch_dim), 1.0, 3) for i, (h, w) in enumerate(grids)], 0)

    def img(i):
        h, w = grids[i]
        return [cfg.vision_start_token_id] + [cfg.image_token_id] * (h * w // 4) + [cfg.vision_end_token_id]

    def txt(tag, n):
        return hash_randint(tag, n, 3, 151000, 3).tolist()

    rows = [[151644, 872, 198] + img(0) + txt("a", 9) + img(1) + txt("b", 17) + [151645],
            [151644, 872, 198] + img(2) + txt("c", 30) + [151645]]
    S = max(len(r) for r in rows)
    ids = torch.tensor([[cfg.pad_token_id] * (S - len(r)) + r for r in rows])
    mask = torch.tensor([[0] * (S - len(r)) + [1] * len(r) for r in rows])
    batch = to_dev({"input_ids": ids, "attention_mask": mask, "pixel_values": pix,
                    "image_grid_thw": torch.tensor([[1, h, w] for h, w in grids])})
    model.engine.taps = {}
    r, _ = model.custom_forward(inputs_batch=batch)
    te, model.engine.taps = model.engine.taps, None
    ref_pos = O.rope_index(cfg, ids, mask, batch["image_grid_thw"].tolist())
    assert torch.equal(te["pos3"].view(3, 2, S).cpu().long(), ref_pos)
    P16 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV)
    P32 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=torch.float32, device=DEV)
    with torch.no_grad():
        o16, o32 = O.custom_forward(P16, cfg, batch).float(), O.custom_forward(P32, cfg, batch).float()
    floor, err = (o16 - o32).abs().max().item(), (r.float() - o32).abs().max().item()
    print(f"two images: engine {r.flatten().tolist()} fp32 {o32.flatten().tolist()} err {err:.4g} floor {floor:.4g}")
    assert err < REWARD_TOL + 3 * floor


def test_qwen_validation_errors(tmp_path_factory):
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(qwen_fixture_batch(fx, fx["batches"][0], cfg))
    with pytest.raises(TypeError):
        model.custom_forward(batch["input_ids"], batch["attention_mask"], batch["pixel_values"], None)
    bad = dict(batch)
    bad["image_grid_thw"] = batch["image_grid_thw"].flip(0)       # per-image token runs no longer match
    with pytest.raises(ValueError, match="Image features and image tokens do not match"):
        model.custom_forward(inputs_batch=bad)
    bad = dict(batch)
    bad["pixel_values"] = batch["pixel_values"][:-4]
    with pytest.raises(ValueError, match="patches"):
        model.custom_forward(inputs_batch=bad)
    with pytest.raises(KeyError):
        model.custom_forward(inputs_batch={k: v for k, v in batch.items() if k != "image_grid_thw"})


def test_qwen_batch_invariance_without_skipca(tmp_path_factory):
    """BT / no SkipCA: a sample scored alone equals the same sample in a left-padded batch (M-RoPE positions count
    valid tokens only, padded keys are masked)."""
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    grids = [(10, 14), (24, 24)]
    both = to_dev(synth_batch_qwen(cfg, grids, None, seed=11, tag="inv"))
    r2, _ = model.custom_forward(inputs_batch=both)
    n0 = int(both["attention_mask"][0].sum())
    n_patch0 = grids[0][0] * grids[0][1]
    alone = {"input_ids": both["input_ids"][:1, -n0:], "attention_mask": both["attention_mask"][:1, -n0:],
             "pixel_values": both["pixel_values"][:n_patch0], "image_grid_thw": both["image_grid_thw"][:1]}
    r1, _ = model.custom_forward(inputs_batch=alone)
    print(f"alone {r1.flatten().tolist()} batched {r2[:1].flatten().tolist()}")
    assert (r1.float() - r2[:1].float()).abs().max().item() < 2e-2


def test_qwen_skipca_without_any_pad_token(tmp_path_factory):
    """SkipCA model, batch with no token-id-151643 position: the reference's vision_pad is empty, attn_o = 0 and the
    head is value_head(ca_layernorm(last_hidden)) - checked against the oracle in bf16."""
    fx = load_fixture("qwen_slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(synth_batch_qwen(cfg, [(12, 12)], None, seed=13, tag="nopad"))
    assert int((batch["input_ids"] == cfg.pad_token_id).sum()) == 0
    r, _ = model.custom_forward(inputs_batch=batch)
    P16 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV)
    P32 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=torch.float32, device=DEV)
    with torch.no_grad():
        o16 = O.custom_forward(P16, cfg, batch).float()
        o32 = O.custom_forward(P32, cfg, batch).float()
    floor = (o16 - o32).abs().max().item()
    err = (r.float() - o32).abs().max().item()
    print(f"no-pad SkipCA: engine {r.flatten().tolist()} fp32 {o32.flatten().tolist()} err {err:.4g} floor {floor:.4g}")
    assert err < REWARD_TOL + 3 * floor


# ----------------------------------------------------------------------------------------------- preprocessing + callers
def test_qwen_gpu_preprocess_bit_exact_vs_transformers():
    """GPU preprocessing (smart_resize + Pillow-exact bicubic taps + lr_qwen_patchify_f32) against transformers' own PIL
    processor (tests/golden/qwen_preprocess.pt): SHA-1 of the float32 bytes."""
    import hashlib
    from preprocess_util import QWEN_CASES, synth_image
    from llava_reward_b200.processing import Qwen2VLImageProcessorB200
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qwen_preprocess.pt"),
                    weights_only=False)
    proc = Qwen2VLImageProcessorB200()
    for e in fx["cases"]:
        h, w = e["hw"]
        out = proc.preprocess([synth_image(e["name"], h, w)], return_tensors="pt")
        pv = out["pixel_values"].cpu().contiguous()
        assert list(pv.shape) == e["shape"] and out["image_grid_thw"][0].tolist() == e["grid"], e["name"]
        assert torch.equal(pv[5:9], e["rows"]), e["name"]
        assert hashlib.sha1(pv.numpy().tobytes()).hexdigest() == e["sha1"], e["name"]
    b = fx["batch"]
    out = proc.preprocess([synth_image(n, *QWEN_CASES[n]) for n in b["names"]], return_tensors="pt")
    assert list(out["pixel_values"].shape) == b["shape"] and out["image_grid_thw"].tolist() == b["grid"]
    assert hashlib.sha1(out["pixel_values"].cpu().contiguous().numpy().tobytes()).hexdigest() == b["sha1"]


def test_qwen_eval_loops_from_uint8_images(tmp_path_factory):
    """the pairwise / single-image loops of eval/batch_inference_rm_qwen.py on uint8 images preprocessed on the GPU by
    the processor wrapper; rewards agree with the oracle preprocessing + oracle forward in fp32"""
    from preprocess_util import synth_image
    from test_qwen_processor_cpu import FakeTok
    from llava_reward_b200.batch_eval import score_pairs, score_single
    from llava_reward_b200.processing import Qwen2_5_VLProcessorB200, Qwen2VLImageProcessorB200
    from oracle import qwen_preprocess_oracle as PO
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    # a small pixel budget keeps the test fast; same code path as the reference's 256..1280 token budget
    proc = Qwen2_5_VLProcessorB200(Qwen2VLImageProcessorB200(min_pixels=16 * 28 * 28, max_pixels=96 * 28 * 28), FakeTok())
    imgs_c = [synth_image("qc0", 300, 420), synth_image("qc1", 500, 260)]
    imgs_r = [synth_image("qr0", 333, 200), synth_image("qr1", 64, 48)]
    text = "<|vision_start|><|image_pad|><|vision_end|>a photo"
    bc = proc(text=[text, text], images=imgs_c, padding=True, return_tensors="pt").to(DEV)
    br = proc(text=[text, text + " of a dog"], images=imgs_r, padding=True, return_tensors="pt").to(DEV)
    res = score_pairs(model, args, [(bc, br)])
    assert res["probs"].shape == (2,) and len(res["chosen_rewards"]) == 2
    single = score_single(model, args, [(bc, torch.tensor([1, 0]))], cls_based=True)
    assert len(single["rewards"]) == 2 and 0.0 <= single["accuracy"] <= 1.0
    # oracle: numpy preprocessing (bit-identical pixels) + fp32 forward
    pix = np.concatenate([PO.preprocess(im, min_pixels=16 * 28 * 28, max_pixels=96 * 28 * 28)[0] for im in imgs_c], 0)
    assert torch.equal(torch.from_numpy(pix), bc["pixel_values"].cpu())
    P32 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=torch.float32, device=DEV)
    with torch.no_grad():
        r32 = O.custom_forward(P32, cfg, dict(bc)).flatten().cpu()
    err = (torch.tensor(res["chosen_rewards"]) - r32).abs().max().item()
    print(f"eval loop: engine {res['chosen_rewards']} oracle fp32 {r32.tolist()} err {err:.4g}")
    assert err < 2e-2


@pytest.mark.parametrize("case", ["qwen_slim_bt", "qwen_slim_gpm"])
def test_qwen_attribute_variants_vs_reference_golden(case, tmp_path_factory):
    """`training` (reward read at position S-1) and `mean_hidden_state` (masked mean before the value head; with the qwen
    SkipCA arm: the S x N_pad cross attention of every row, padded keys masked with -1e4, :387-397) set on the
    model object as on the reference's (rw_model_general_preference.py:327-333, 398-448); goldens made by setting them
    on the reference model (tests/golden/make_golden_qwen.py). Right-padded batches are skipped for `training`: position S-1
    is a padded row there, which is not defined behaviour."""
    from oracle import qwen_vl_oracle as OO
    from oracle.reward_oracle import Params
    from llava_reward_b200.synth import SynthProvider
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    P = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV, cache=False)
    kws = {"training": dict(training=True), "mean": dict(mean_hidden_state=True)}
    checked = 0
    for entry in fx["batches"]:
        batch = to_dev(qwen_fixture_batch(fx, entry, cfg))
        for key, g in entry["attrs"].items():
            if key == "training" and entry["padding_side"] == "right":
                continue
            saved = (model.training, model.mean_hidden_state)
            model.training, model.mean_hidden_state = key == "training", key == "mean"
            try:
                r, _ = model.custom_forward(inputs_batch=batch)
            finally:
                model.training, model.mean_hidden_state = saved
            assert tuple(r.shape) == tuple(g.shape), key
            with torch.no_grad():
                ro = OO.custom_forward(P, cfg, batch, **kws[key]).float().cpu()
            err, floor = (r.float().cpu() - g).abs().max().item(), (ro - g).abs().max().item()
            print(f"{case}/{entry['tag']}/{key}: engine-vs-fp32 {err:.4g}, reference bf16-vs-fp32 {floor:.4g}")
            assert err < REWARD_TOL + 3.0 * floor, key
            checked += 1
    assert checked >= 1
    assert model.mean_hidden_state in (None, False)


@pytest.mark.parametrize("case", ["qwen_slim_bt", "qwen_slim_gpm", "qwen_wide_gpm"])
def test_qwen_packed_valid_rows_are_output_identical(case, tmp_path_factory):
    """engine.pack_rows (the decoder on the valid rows only, packed-sequence attention) against the slot layout:
    bit-identical rewards on the golden batches (left and right padding, mixed lengths)."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    eng = model.engine
    try:
        for entry in fx["batches"]:
            batch = to_dev(qwen_fixture_batch(fx, entry, cfg))
            eng.pack_rows = True
            rp = model.custom_forward(inputs_batch=batch)[0].clone()
            n_packed = eng.launches
            eng.pack_rows = False
            rs = model.custom_forward(inputs_batch=batch)[0].clone()
            m = batch["attention_mask"]
            print(f"{case}/{entry['tag']} ({entry['padding_side']} padding): valid rows {int(m.sum())} of {m.numel()}, "
                  f"|d| {(rp.float() - rs.float()).abs().max().item():.3g}, launches {n_packed} / {eng.launches}")
            assert torch.equal(rp, rs)
    finally:
        eng.pack_rows = True


def test_return_output_hidden_states_inputs_batch_branch(tmp_path_factory):
    """custom_forward(inputs_batch=..., return_output=True): the reference returns `self.forward(..., output_hidden_states=
    True)` (rw_model_general_preference.py:357, 374); the engine returns the same hidden_states tuple (inputs_embeds, h_1
    .. h_{L-1}, norm(h_L)) with logits None (the lm_head GEMM custom_forward never reads is not executed)."""
    fx = load_fixture("qwen_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    batch = {k: v.to(DEV) for k, v in qwen_fixture_batch(fx, entry, cfg).items()}
    r0, none = model.custom_forward(inputs_batch=batch)
    r, out = model.custom_forward(inputs_batch=batch, return_output=True)
    assert none is None and model.engine.taps is None and out.logits is None
    assert (r.float() - r0.float()).abs().max().item() < 2e-2
    hs = out["hidden_states"]
    assert len(hs) == cfg.num_layers + 1
    valid = batch["attention_mask"].bool()[:, :, None]
    for name, t in (("inputs_embeds", hs[0]), ("hidden_0", hs[1]), ("last_hidden", hs[-1])):
        g = entry["taps"][name]
        assert list(t.shape) == g["shape"], name
        tt = torch.where(valid.expand_as(t), t, torch.zeros_like(t))
        a = tt.float().flatten()[:: g["stride"]][:2048].cpu()
        w = valid.expand_as(t).float().flatten()[:: g["stride"]][:2048].cpu()
        rel = ((a - g["vals"] * w).norm() / (g["vals"] * w).norm()).item()
        print(f"return_output {name}: rel L2 err vs reference fp32 {rel:.4g}")
        assert rel < 3e-2, name
