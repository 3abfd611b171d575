"""LLaVA-v1.6 branch on the B200: kernels added for it (token plan with an image token id / arange positions, anyres
pack + embedding gather, head_dim-128 attention and RoPE epilogue) and the engine end to end against
 (a) the reference's own fp32 outputs (tests/golden/llava_*.pt, made by tests/golden/make_golden_llava.py),
 (b) the oracle restatement in bf16 on the same GPU, stage by stage.
Tolerance (north_star): rewards within 2e-2 absolute in bf16 (+ the measured bf16 noise of the reference arithmetic)."""
import os
import types

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

from golden_util import llava_fixture_batch, llava_fixture_cfg, load_fixture  # noqa: E402
from test_kernels_gpu import attn_ref, check_close, rnd  # noqa: E402
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402
from llava_reward_b200.config import LlavaNextRewardConfig, anyres_geometry  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from llava_reward_b200.synth import SynthProvider, synth_batch_llava  # noqa: E402
from oracle import llava_next_oracle as O  # noqa: E402
from oracle.reward_oracle import Params  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16
REWARD_TOL = 2e-2
_models = {}


# ----------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("impl", [L.ATTN_TCGEN05, L.ATTN_MMA_SYNC, L.ATTN_TCGEN05_2TILE, L.ATTN_TCGEN05_1TILE])
@pytest.mark.parametrize("heads,T,nseq", [(4, 130, 3), (2, 2395, 2), (32, 700, 2), (1, 128, 1), (3, 257, 2)])
def test_attention_hd128(heads, T, nseq, impl):
    hd = 128
    D = heads * hd
    qkv = rnd(nseq * T, 3 * D, seed=5)
    o = torch.full((nseq * T, D + 128), float("nan"), dtype=bf, device=DEV)
    lens = [T, T - 37, 5][:nseq]
    starts = [T - n for n in lens]
    if nseq > 1:
        starts[1] = 0  # right padding on one sequence
    ss = torch.tensor(starts, dtype=torch.int32, device=DEV)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    scale = hd ** -0.5
    ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D + 128, nseq, T, ss, sl, heads, hd, True, scale, impl)
    torch.cuda.synchronize()
    f = qkv.float().view(nseq, T, 3, heads, hd)
    for s in range(nseq):
        ref = attn_ref(f[s, :, 0], f[s, :, 1], f[s, :, 2], True, scale, starts[s], lens[s]).reshape(T, D)
        check_close(o[s * T:(s + 1) * T, :D], ref, f"attention hd128 seq {s}", atol=1e-2, rtol=2e-2)


def test_attention_hd128_rejects_split_softmax_variant():
    qkv = rnd(256, 3 * 128, seed=1)
    o = torch.empty(256, 128, dtype=bf, device=DEV)
    with pytest.raises(RuntimeError, match="LR_ERR_BAD_ARG"):
        ops.attention(qkv, qkv[:, 128:], qkv[:, 256:], o, 384, 128, 1, 256, None, None, 1, 128, True, 0.1,
                      L.ATTN_TCGEN05_SPLIT)


@pytest.mark.parametrize("impl", [L.GEMM_TCGEN05, L.GEMM_SIMT])
def test_gemm_rope_hd128(impl):
    """fused q/k/v projection + plain RoPE at head_dim 128 == GEMM then lr_rope_su_bf16, bit for bit"""
    M, heads, hd, K = 600, 4, 128, 896
    H = heads * hd
    N = 3 * H
    x = rnd(M, K, seed=31)
    W = rnd(N, K, std=K ** -0.5, seed=32)
    pos = torch.randint(0, 3000, (M,), device=DEV, dtype=torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=DEV).float() / hd))
    ang = torch.arange(3000, device=DEV).float()[:, None] * inv[None]
    cos, sin = ang.cos().to(bf).contiguous(), ang.sin().to(bf).contiguous()
    ref = torch.empty(M, N, dtype=bf, device=DEV)
    ops.gemm(x, W, ref, M, N, K, impl=L.GEMM_TCGEN05_PAIR if impl == L.GEMM_TCGEN05 else impl)
    ops.rope_su(ref, pos, cos, sin, M, heads, hd)
    inter = torch.stack([torch.arange(hd // 2), torch.arange(hd // 2) + hd // 2], dim=1).reshape(-1)
    qk_perm = (torch.arange(2 * heads)[:, None] * hd + inter[None, :]).reshape(-1)
    perm = torch.cat([qk_perm, torch.arange(2 * H, 3 * H)]).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    ops.gemm_rope(x, W[perm].contiguous(), out, M, N, K, pos, cos, sin, 2 * H, hd, impl)
    torch.cuda.synchronize()
    unperm = torch.empty_like(out)
    unperm[:, perm] = out
    assert torch.equal(unperm, ref)


def test_token_plan_ex_image_token_and_arange():
    B, S, tok = 4, 900, 32000
    ids = torch.randint(3, 31999, (B, S), dtype=torch.int64)
    mask = torch.ones(B, S, dtype=torch.int64)
    pads, nimg = [0, 100, 450, 0], [300, 250, 7, 0]
    mask[3, 700:] = 0  # right padding
    for b in range(B):
        mask[b, :pads[b]] = 0
        ids[b, pads[b] + 5: pads[b] + 5 + nimg[b]] = tok
    ids[0, 2] = -1  # a negative id is NOT an image position in this mode
    ids_d, mask_d = ids.to(DEV), mask.to(DEV)
    i32 = dict(dtype=torch.int32, device=DEV)
    pos, ordn = torch.empty(B * S, **i32), torch.empty(B * S, **i32)
    ss, sl, er, ni = (torch.zeros(B, **i32) for _ in range(4))
    fl = torch.zeros(1, **i32)
    ops.token_plan_ex(ids_d, mask_d, B, S, tok, L.POS_ARANGE, pos, ordn, ss, sl, er, ni, fl)
    torch.cuda.synchronize()
    assert torch.equal(pos.view(B, S).cpu(), torch.arange(S, dtype=torch.int32)[None].expand(B, S))
    is_img = ids == tok
    ref_ord = torch.where(is_img, is_img.long().cumsum(1) - 1, torch.full_like(ids, -1)).to(torch.int32)
    assert torch.equal(ordn.view(B, S).cpu(), ref_ord)
    assert ni.cpu().tolist() == nimg
    assert ss.cpu().tolist() == [0, 100, 450, 0] and sl.cpu().tolist() == [900, 800, 450, 700]
    eos = S - 1 - mask.flip(1).argmax(1)
    assert er.cpu().tolist() == [b * S + int(eos[b]) for b in range(B)]
    assert fl.item() == 0
    # mask-derived positions stay available with an image token id
    ops.token_plan_ex(ids_d, mask_d, B, S, tok, L.POS_FROM_MASK, pos, ordn, ss, sl, er, ni, fl)
    ref_pos = (mask.cumsum(1) - 1).masked_fill(mask == 0, 1).to(torch.int32)
    assert torch.equal(pos.view(B, S).cpu(), ref_pos)


@pytest.mark.parametrize("sizes", [[(480, 640), (500, 333)], [(300, 900), (672, 672), (1000, 700)], [(336, 336)]])
def test_anyres_embed_scatter_matches_pack_image_features(sizes):
    """one kernel == embedding lookup + pack_image_features (unpad, image_newline) + masked_scatter of the oracle"""
    cfg = LlavaNextRewardConfig(hidden_size=256, intermediate_size=512, num_heads=2, num_layers=1, clip_layers=1)
    H, V, T = cfg.hidden_size, 1000, 577
    batch = synth_batch_llava(cfg, len(sizes), sizes, None, seed=3)
    ids = batch["input_ids"].clamp(max=V - 1)
    ids[batch["input_ids"] == cfg.image_token_id] = cfg.image_token_id  # keep the placeholders
    B, S = ids.shape
    geos = [anyres_geometry(hw, cfg.image_grid_pinpoints) for hw in sizes]
    n_patches = sum(g["n_patches"] for g in geos)
    feat = rnd(n_patches * T, H, seed=11)
    newline = rnd(H, seed=12)
    wte = rnd(V, H, seed=13)
    # oracle packing on the projector output without the CLS rows
    f3 = feat.view(n_patches, T, H)[:, 1:]
    side, packed = 24, []
    base = 0
    for g, hw in zip(geos, sizes):
        fm = f3[base: base + g["n_patches"]]
        base += g["n_patches"]
        grid = fm[1:].view(g["grid_h"], g["grid_w"], side, side, H).permute(4, 0, 2, 1, 3).contiguous()
        grid = grid.flatten(1, 2).flatten(2, 3)
        grid = O.unpad_image(grid, hw)
        grid = torch.cat((grid, newline[:, None, None].expand(*grid.shape[:-1], 1)), dim=-1)
        packed.append(torch.cat((fm[0], grid.flatten(1, 2).transpose(0, 1)), dim=0))
    packed = torch.cat(packed, 0)
    ids_d = ids.to(DEV)
    sel = ids_d == cfg.image_token_id
    ref = wte[ids_d.clamp(max=V - 1)]
    ref = ref.masked_scatter(sel[..., None].expand_as(ref), packed)
    # product path
    i32 = dict(dtype=torch.int32, device=DEV)
    pos, ordn = torch.empty(B * S, **i32), torch.empty(B * S, **i32)
    ss, sl, er, ni = (torch.zeros(B, **i32) for _ in range(4))
    fl = torch.zeros(1, **i32)
    ops.token_plan_ex(ids_d, batch["attention_mask"].to(DEV), B, S, cfg.image_token_id, L.POS_ARANGE, pos, ordn, ss, sl,
                      er, ni, fl)
    plan = torch.zeros(B, L.PLAN_STRIDE, dtype=torch.int32)
    pb = 0
    for b, g in enumerate(geos):
        plan[b, :7] = torch.tensor([g["grid_h"], g["grid_w"], pb, 0, g["n_tokens"], g["top"], g["left"]])
        pb += g["n_patches"]
    assert ni.cpu().tolist() == [g["n_tokens"] for g in geos]
    hidden = torch.full((B * S, H), float("nan"), dtype=bf, device=DEV)
    ops.anyres_embed_scatter(ids_d, ordn, plan.to(DEV).view(-1), wte, feat, newline, hidden, B, S, H, V)
    torch.cuda.synchronize()
    assert torch.equal(hidden.view(B, S, H), ref)


# ----------------------------------------------------------------------------------------------- engine
def build_model(fx, tmp_path_factory):
    key = fx["case"]
    if key not in _models:
        cfg = llava_fixture_cfg(fx)
        d = tmp_path_factory.mktemp(key)
        ypath = os.path.join(d, "reward_config.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"is_general_preference": cfg.is_general_preference, "add_cross_attention": False,
                            "value_head_dim": cfg.value_head_dim,
                            "general_preference_tau": cfg.general_preference_tau}, f)
        over = {k: v for k, v in fx["cfg_overrides"].items() if k != "is_general_preference"}
        args = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None,
                                     ft_projector=False, disable_fast_tokenizer=False, config_overrides=over)
        args, model = load_reward_adaptor(args, "llava", ypath)
        _models[key] = (args, model.to("cuda").eval(), cfg)
    return _models[key]


def to_dev(batch):
    return {k: v.to(DEV) for k, v in batch.items()}


def bf16_floor(fx, cfg, batch, ref_fp32):
    """|reference arithmetic in bf16 - reference fp32| on this GPU for two summation orders (max, rms)."""
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


@pytest.mark.parametrize("case", ["llava_slim_bt", "llava_slim_gpm", "llava_wide_bt"])
def test_llava_vs_reference_golden(case, tmp_path_factory):
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    rewards, errs, floors = {}, [], []
    for entry in fx["batches"]:
        batch = to_dev(llava_fixture_batch(fx, entry, cfg))
        r, _ = model.custom_forward(inputs_batch=batch)
        assert r.dtype == bf and r.is_cuda and tuple(r.shape) == tuple(entry["reward"].shape)
        errs.append((r.float().cpu() - entry["reward"]).abs().max().item())
        floors.append(bf16_floor(fx, cfg, batch, entry["reward"]))
        print(f"{case}/{entry['tag']}: engine {r.flatten().tolist()} ref {entry['reward'].flatten().tolist()} | "
              f"engine-vs-fp32 {errs[-1]:.4g}, reference bf16-vs-fp32 max {floors[-1][0]:.4g} rms {floors[-1][1]:.4g}")
        rewards[entry["tag"]] = r
    mx = max(f[0] for f in floors)
    rms = (sum(f[1] ** 2 for f in floors) / len(floors)) ** 0.5
    assert max(errs) < REWARD_TOL + max(mx, 3 * rms), f"{case}: reward err {max(errs):.4g} vs reference fp32"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    ref = fx["prob"].numpy()
    decided = abs(ref - 0.5) > 0.05
    assert ((prob > 0.5) == (ref > 0.5))[decided].all()
    print(f"{case}: prob engine {prob.tolist()} reference {ref.tolist()}")


@pytest.mark.parametrize("case", ["llava_slim_bt", "llava_slim_gpm"])
def test_llava_stages_vs_oracle(case, tmp_path_factory):
    """Stage by stage: engine (bf16 kernels) and the oracle in bf16 on this GPU, both against the oracle in fp32 -
    the engine may not be further from fp32 than 1.5x the reference arithmetic in bf16 (+2e-3)."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][1]
    batch = to_dev(llava_fixture_batch(fx, entry, cfg))
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
    valid = mask.bool()
    T = cfg.clip_tokens
    rows = [("projector_out", te["projector_out"].view(-1, T, cfg.hidden_size)[:, 1:], t16["projector_out"],
             t32["projector_out"]),
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


def test_llava_image_rows_are_bit_exact_gathers(tmp_path_factory):
    """inputs_embeds rows at image positions are pure copies of projector rows / image_newline: compare the engine's
    own projector output packed by the ORACLE's pack with the engine's inputs_embeds, bit for bit."""
    fx = load_fixture("llava_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(llava_fixture_batch(fx, fx["batches"][0], cfg))
    model.engine.taps = {}
    model.custom_forward(inputs_batch=batch)
    te, model.engine.taps = model.engine.taps, None
    H, T = cfg.hidden_size, cfg.clip_tokens
    feats = te["projector_out"].view(-1, T, H)[:, 1:]
    sizes = [tuple(int(v) for v in s) for s in batch["image_sizes"].tolist()]
    newline = SynthProvider(cfg, seed=fx["seed_w"], device=DEV)("image_newline").to(bf)
    packed, base = [], 0
    for hw in sizes:
        g = anyres_geometry(hw, cfg.image_grid_pinpoints)
        fm = feats[base: base + g["n_patches"]]
        base += g["n_patches"]
        grid = fm[1:].view(g["grid_h"], g["grid_w"], 24, 24, H).permute(4, 0, 2, 1, 3).contiguous().flatten(1, 2).flatten(2, 3)
        grid = O.unpad_image(grid, hw)
        grid = torch.cat((grid, newline[:, None, None].expand(*grid.shape[:-1], 1)), dim=-1)
        packed.append(torch.cat((fm[0], grid.flatten(1, 2).transpose(0, 1)), dim=0))
    packed = torch.cat(packed, 0)
    sel = batch["input_ids"] == cfg.image_token_id
    got = te["inputs_embeds"].view(*sel.shape, H)[sel]
    assert torch.equal(got, packed)


def test_llava_validation_errors(tmp_path_factory):
    fx = load_fixture("llava_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    batch = to_dev(llava_fixture_batch(fx, fx["batches"][0], cfg))
    with pytest.raises(TypeError):
        model.custom_forward(batch["input_ids"], batch["attention_mask"], batch["pixel_values"], batch["image_sizes"])
    bad = dict(batch)
    bad["image_sizes"] = torch.tensor([[672, 672], [500, 333]], device=DEV)  # token count no longer matches
    with pytest.raises(ValueError, match="Image features and image tokens do not match"):
        model.custom_forward(inputs_batch=bad)
    bad = dict(batch)
    bad["pixel_values"] = batch["pixel_values"][:, :2]
    with pytest.raises(ValueError, match="patches"):
        model.custom_forward(inputs_batch=bad)
    with pytest.raises(KeyError):
        model.custom_forward(inputs_batch={k: v for k, v in batch.items() if k != "image_sizes"})


def test_llava_batch_invariance(tmp_path_factory):
    """A sample scored alone equals the same sample scored in a left-padded batch (no SkipCA in this branch, arange
    positions: RoPE is relative, so padding only shifts absolute positions)."""
    fx = load_fixture("llava_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    sizes = [(500, 333), (672, 672)]
    both = to_dev(synth_batch_llava(cfg, 2, sizes, None, seed=11, tag="inv"))
    r2, _ = model.custom_forward(inputs_batch=both)
    n0 = int(both["attention_mask"][0].sum())
    alone = {"input_ids": both["input_ids"][:1, -n0:], "attention_mask": both["attention_mask"][:1, -n0:],
             "pixel_values": both["pixel_values"][:1], "image_sizes": both["image_sizes"][:1]}
    r1, _ = model.custom_forward(inputs_batch=alone)
    print(f"alone {r1.flatten().tolist()} batched {r2[:1].flatten().tolist()}")
    assert (r1.float() - r2[:1].float()).abs().max().item() < 2e-2


@pytest.mark.parametrize("H", [4096, 5120])
def test_value_head_wide_hidden(H):
    """value-head-only form of lr_skipca_head at the Vicuna-7B / 13B hidden sizes"""
    B, vhd = 3, 2
    x = rnd(B, H, seed=41)
    w = rnd(vhd, H, std=H ** -0.5, seed=42)
    out = torch.empty(B, vhd, dtype=bf, device=DEV)
    ops.skipca_head(None, None, None, x, None, w, out, B, H, 0, vhd, 1e-5)
    torch.cuda.synchronize()
    check_close(out, x.float() @ w.float().t(), "value head", atol=1e-2, rtol=1e-2)


# ----------------------------------------------------------------------------------------------- preprocessing + callers
def test_llava_gpu_preprocess_bit_exact_vs_transformers():
    """GPU anyres preprocessing (Pillow-exact bicubic taps + lr_patch_pack_f32) against transformers' own PIL processor
    (tests/golden/llava_preprocess.pt): SHA-1 of the float32 bytes."""
    import hashlib
    from preprocess_util import synth_image
    from llava_reward_b200.processing import LlavaNextImageProcessorB200
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "llava_preprocess.pt"),
                    weights_only=False)
    proc = LlavaNextImageProcessorB200()
    for e in fx["cases"]:
        h, w = e["hw"]
        out = proc.preprocess([synth_image(e["name"], h, w)], return_tensors="pt")
        pv = out["pixel_values"][0].cpu().contiguous()
        assert list(pv.shape) == e["shape"] and out["image_sizes"][0].tolist() == e["image_sizes"]
        assert torch.equal(pv[0, :, 100:104, :], e["base_rows"]), e["name"]
        assert torch.equal(pv[1, :, 100:104, :], e["patch1_rows"]), e["name"]
        assert hashlib.sha1(pv.numpy().tobytes()).hexdigest() == e["sha1"], e["name"]
    b = fx["batch"]
    from preprocess_util import LLAVA_CASES
    out = proc.preprocess([synth_image(n, *LLAVA_CASES[n]) for n in b["names"]], return_tensors="pt")
    assert list(out["pixel_values"].shape) == b["shape"] and out["image_sizes"].tolist() == b["image_sizes"]
    assert hashlib.sha1(out["pixel_values"].cpu().contiguous().numpy().tobytes()).hexdigest() == b["sha1"]


def test_llava_eval_loops_and_best_of_n(tmp_path_factory):
    """the pairwise / single-image loops of eval/batch_inference_rm_llava.py and best-of-N over uint8 images
    preprocessed on the GPU"""
    from preprocess_util import synth_image
    from llava_reward_b200.batch_eval import best_of_n, score_pairs, score_single
    from llava_reward_b200.processing import LlavaNextImageProcessorB200
    fx = load_fixture("llava_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    bc = to_dev(llava_fixture_batch(fx, fx["batches"][0], cfg))
    br = to_dev(llava_fixture_batch(fx, fx["batches"][1], cfg))
    res = score_pairs(model, args, [(bc, br)])
    ref = fx["prob"].numpy()
    assert res["probs"].shape == (2,) and abs(res["probs"] - ref).max() < 0.05
    single = score_single(model, args, [(bc, torch.tensor([1, 0]))], cls_based=True)
    assert len(single["rewards"]) == 2 and 0.0 <= single["accuracy"] <= 1.0
    # best-of-N from uint8 images: GPU preprocessing -> scoring
    proc = LlavaNextImageProcessorB200(cfg.image_grid_pinpoints)
    imgs = [synth_image(f"cand{i}", 400 + 40 * i, 520) for i in range(4)]
    pp = proc.preprocess(imgs, return_tensors="pt")
    from llava_reward_b200.config import anyres_geometry as geo
    rows = [[1, 5, 6, 7] + [cfg.image_token_id] * geo(im.shape[:2], cfg.image_grid_pinpoints)["n_tokens"] + [2] for im in imgs]
    S = max(len(r) for r in rows)
    ids = torch.tensor([[0] * (S - len(r)) + r for r in rows])
    mask = torch.tensor([[0] * (S - len(r)) + [1] * len(r) for r in rows])
    batch = {"input_ids": ids.to(DEV), "attention_mask": mask.to(DEV), "pixel_values": pp["pixel_values"],
             "image_sizes": pp["image_sizes"]}
    out = best_of_n(model, [batch])
    assert len(out["rewards"]) == 4 and 0 <= out["best"] < 4
    # the same candidates through the oracle preprocessing + oracle forward (fp32): rewards agree within bf16 noise
    from oracle import llava_preprocess_oracle as PO
    pix = torch.zeros_like(pp["pixel_values"], device="cpu")
    for i, im in enumerate(imgs):
        pv, _ = PO.preprocess(im)
        pix[i, : pv.shape[0]] = torch.from_numpy(pv)
    assert torch.equal(pix, pp["pixel_values"].cpu())
    P32 = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=torch.float32, device=DEV)
    with torch.no_grad():
        r32 = O.custom_forward(P32, cfg, {**batch, "pixel_values": pix.to(DEV)}).flatten().cpu()
    err = (torch.tensor(out["rewards"]) - r32).abs().max().item()
    print(f"best-of-4: engine {out['rewards']} oracle fp32 {r32.tolist()} err {err:.4g}")
    assert err < 2e-2


@pytest.mark.parametrize("case", ["llava_slim_bt", "llava_slim_gpm"])
def test_llava_attribute_variants_vs_reference_golden(case, tmp_path_factory):
    """`training` (reward read at position S-1) and `mean_hidden_state` (masked mean before the value head) set on the
    model object as on the reference's (rw_model_general_preference.py:327-333, 398-448); goldens made by setting them
    on the reference model (tests/golden/make_golden_llava.py). Right-padded batches are skipped for `training`: position S-1
    is a padded row there, which is not defined behaviour."""
    from oracle import llava_next_oracle as OO
    from oracle.reward_oracle import Params
    from llava_reward_b200.synth import SynthProvider
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    P = Params(SynthProvider(cfg, seed=fx["seed_w"], device=DEV), dtype=bf, device=DEV, cache=False)
    kws = {"training": dict(training=True), "mean": dict(mean_hidden_state=True)}
    checked = 0
    for entry in fx["batches"]:
        batch = to_dev(llava_fixture_batch(fx, entry, cfg))
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


@pytest.mark.parametrize("case", ["llava_slim_bt", "llava_slim_gpm"])
def test_llava_packed_valid_rows_are_output_identical(case, tmp_path_factory):
    """engine.pack_rows (the decoder on the valid rows only, packed-sequence attention) against the slot layout:
    bit-identical rewards on the golden batches (left and right padding, mixed lengths)."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    eng = model.engine
    try:
        for entry in fx["batches"]:
            batch = to_dev(llava_fixture_batch(fx, entry, cfg))
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
    fx = load_fixture("llava_slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    batch = {k: v.to(DEV) for k, v in llava_fixture_batch(fx, entry, cfg).items()}
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
