"""End-to-end parity of the CUDA engine (through the reference-shaped API) against
 (a) the reference's own fp32 outputs stored in tests/golden/*.pt (made by tests/golden/make_golden.py),
 (b) the oracle restatement run in bf16 on the same GPU (= the reference's GPU arithmetic, eager attention).
Tolerance (north_star): per-sample rewards within 2e-2 absolute in bf16; identical preference decisions."""
import os
import types

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

from golden_util import GOLDEN_DIR, fixture_batch, fixture_cfg, load_fixture  # noqa: E402
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from llava_reward_b200.synth import SynthProvider  # noqa: E402
from oracle import reward_oracle as O  # noqa: E402

REWARD_TOL = 2e-2
_models = {}


def build_model(fx, tmp_path_factory):
    key = fx["case"]
    if key not in _models:
        cfg = fixture_cfg(fx)
        d = tmp_path_factory.mktemp(key)
        ypath = os.path.join(d, "reward_config.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"is_general_preference": cfg.is_general_preference,
                            "add_cross_attention": cfg.add_cross_attention,
                            "value_head_dim": cfg.value_head_dim,
                            "general_preference_tau": cfg.general_preference_tau}, f)
        args = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None,
                                     ft_projector=False, disable_fast_tokenizer=False,
                                     config_overrides={k: v for k, v in fx["cfg_overrides"].items()
                                                       if k in ("num_layers", "clip_layers", "use_lora")})
        args, model = load_reward_adaptor(args, "phi3v", ypath)
        _models[key] = (args, model.to("cuda").eval(), cfg)
    return _models[key]


def reference_bf16_realizations(fx, cfg, ids, mask, pix, sizes, n=3):
    """Rewards of the reference arithmetic in bf16 on this GPU under n mathematically equivalent evaluations that only
    differ in fp32 summation order inside cuBLAS (plain / reduced-precision reductions off / batch duplicated)."""
    P = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda", cache=False)
    outs = []
    with torch.no_grad():
        outs.append(O.custom_forward(P, cfg, ids, mask, pix, sizes).float().cpu())
        if n > 1:
            old = torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction
            torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = not old
            try:
                outs.append(O.custom_forward(P, cfg, ids, mask, pix, sizes).float().cpu())
            finally:
                torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = old
        if n > 2:
            B = ids.shape[0]
            dup = lambda t: torch.cat([t, t], 0)  # noqa: E731
            outs.append(O.custom_forward(P, cfg, dup(ids), dup(mask), dup(pix), dup(sizes))[:B].float().cpu())
    return outs


def bf16_noise_floor(fx, cfg, ids, mask, pix, sizes, ref_fp32, n=3):
    """(max, rms) of |reference-in-bf16 - reference-in-fp32| over n bf16 realizations: the bf16 rounding noise of the
    random-init network itself (hard part 3). A from-scratch bf16 implementation cannot be closer to
    the reference than the reference's own bf16 evaluations are to its fp32 run."""
    d = torch.stack([(r - ref_fp32).abs() for r in reference_bf16_realizations(fx, cfg, ids, mask, pix, sizes, n)])
    return d.max().item(), d.pow(2).mean().sqrt().item()


def reward_gate(err, floors):
    """north_star tolerance 2e-2, widened by the measured bf16 noise of the reference arithmetic on the same inputs:
    3 x RMS over realizations (a one-sided 3-sigma bound), never below the largest single realization."""
    mx = max(f[0] for f in floors)
    rms = (sum(f[1] ** 2 for f in floors) / len(floors)) ** 0.5
    return REWARD_TOL + max(mx, 3.0 * rms)


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt", "slim_bt_long"])
def test_slim_vs_reference_golden(case, tmp_path_factory):
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    rewards, errs, floors = {}, [], []
    for entry in fx["batches"]:
        ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
        r, _ = model.custom_forward(ids, mask, pix, sizes)
        assert r.dtype == torch.bfloat16 and r.is_cuda and tuple(r.shape) == tuple(entry["reward"].shape)
        errs.append((r.float().cpu() - entry["reward"]).abs().max().item())
        # the reference's OWN bf16 error against its fp32 run on these inputs (oracle in bf16 on this GPU)
        floors.append(bf16_noise_floor(fx, cfg, ids, mask, pix, sizes, entry["reward"]))
        print(f"{case}/{entry['tag']}: engine-vs-fp32 {errs[-1]:.4g}, reference bf16-vs-fp32 max {floors[-1][0]:.4g} "
              f"rms {floors[-1][1]:.4g}")
        rewards[entry["tag"]] = r
    assert max(errs) < reward_gate(max(errs), floors), f"{case}: reward err {max(errs):.4g} vs reference fp32"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    assert prob.dtype.name == "float32" and prob.shape == (fx["prob"].shape[0],)
    ref = fx["prob"].numpy()
    decided = abs(ref - 0.5) > 0.05  # pairs whose reference margin is above bf16 noise
    assert ((prob > 0.5) == (ref > 0.5))[decided].all()


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt"])
def test_slim_stages_vs_oracle_bf16(case, tmp_path_factory):
    """Stage-by-stage against the oracle in bf16 on the GPU (relative L2 error per stage)."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
    P = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda")
    taps_o = {}
    with torch.no_grad():
        r_o = O.custom_forward(P, cfg, ids, mask, pix, sizes, taps_o)
    model.engine.taps = {}
    r_e, _ = model.custom_forward(ids, mask, pix, sizes)
    taps_e, model.engine.taps = model.engine.taps, None
    B, S = ids.shape
    valid = mask.bool()
    # CLIP: the engine runs only the real crops, compacted; the oracle runs all 17 slots like the reference
    n_slots = pix.shape[1]
    real = []
    for b in range(B):
        nc = (int(sizes[b, 0]) // 336) * (int(sizes[b, 1]) // 336) + 1
        real += [b * n_slots + i for i in range(nc)]
    real = torch.tensor(real, device="cuda")
    T = cfg.clip_tokens
    errs = {
        "clip_embed": rel_err(taps_e["clip_embed"].view(-1, T, 1024), taps_o["clip_embed"][real]),
        "clip_layer0": rel_err(taps_e["clip_layer0"].view(-1, T, 1024), taps_o["clip_layer0"][real]),
        "clip_out": rel_err(taps_e["clip_out"].view(-1, T, 1024)[:, 1:], taps_o["clip_features"].flatten(0, 1)[real]),
        "img_proj": rel_err(taps_e["img_proj"], taps_o["img_proj"]),
        "inputs_embeds": rel_err(taps_e["inputs_embeds"].view(B, S, -1)[valid], taps_o["inputs_embeds"][valid]),
    }
    for i in range(cfg.num_layers):
        errs[f"hidden_{i}"] = rel_err(taps_e[f"hidden_{i}"].view(B, S, -1)[valid], taps_o[f"hidden_{i}"][valid])
    errs["last_hidden_eos"] = rel_err(taps_e["last_hidden_eos"], taps_o["last_hidden"][:, -1])
    print({k: round(v, 5) for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 3e-2, (k, v)
    floor = bf16_noise_floor(fx, cfg, ids, mask, pix, sizes, entry["reward"])
    diff = (r_e.float() - r_o.float()).abs().max().item()
    print(f"engine-vs-oracle(bf16) {diff:.4g}, oracle bf16-vs-fp32 max {floor[0]:.4g} rms {floor[1]:.4g}")
    assert diff < reward_gate(diff, [floor])


def test_batch_composition_quirk(tmp_path_factory):
    """SkipCA softmax includes the zero-padded vision rows of the longest image in the batch
    (SURVEY.md 3.1): scoring sample 0 alone or next to a larger image follows the oracle in both cases."""
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
    r_batched, _ = model.custom_forward(ids, mask, pix, sizes)
    first = int(mask[0].nonzero()[0])
    r_alone, _ = model.custom_forward(ids[:1, first:], mask[:1, first:], pix[:1], sizes[:1])
    P = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.float32, device="cuda")
    P16 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda")
    with torch.no_grad():
        o_b = O.custom_forward(P, cfg, ids, mask, pix, sizes)
        o_a = O.custom_forward(P, cfg, ids[:1, first:], mask[:1, first:], pix[:1], sizes[:1])
        h_b = O.custom_forward(P16, cfg, ids, mask, pix, sizes)
        h_a = O.custom_forward(P16, cfg, ids[:1, first:], mask[:1, first:], pix[:1], sizes[:1])
    floor = 3.0 * max((h_b.float() - o_b).abs().max().item(), (h_a.float() - o_a).abs().max().item())
    e_b = (r_batched[0].float() - o_b[0]).abs().max().item()
    e_a = (r_alone[0].float() - o_a[0]).abs().max().item()
    print(f"quirk: fp32 oracle alone {o_a[0].tolist()} batched {o_b[0].tolist()} | engine err batched {e_b:.4g} "
          f"alone {e_a:.4g} | reference bf16 noise {floor:.4g}")
    assert e_b < REWARD_TOL + floor and e_a < REWARD_TOL + floor


def test_tcgen05_and_simt_engines_agree(tmp_path_factory):
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    ids, mask, pix, sizes = fixture_batch(fx, fx["batches"][0], cfg, device="cuda")
    r1, _ = model.custom_forward(ids, mask, pix, sizes)
    model.engine.gemm_impl = L.GEMM_SIMT
    try:
        r2, _ = model.custom_forward(ids, mask, pix, sizes)
    finally:
        model.engine.gemm_impl = L.GEMM_TCGEN05
    # same arithmetic, different fp32 summation order: differences are pure bf16-noise amplification
    assert (r1.float() - r2.float()).abs().max().item() < REWARD_TOL + 1e-2


def test_input_validation(tmp_path_factory):
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    ids, mask, pix, sizes = fixture_batch(fx, fx["batches"][0], cfg, device="cuda")
    with pytest.raises(ValueError):
        model.custom_forward(ids, mask, None, None)
    bad = sizes.clone()
    bad[0, 0] = 672
    with pytest.raises(ValueError):
        model.custom_forward(ids, mask, pix, bad)
    with pytest.raises(AssertionError):
        model.custom_forward(ids, mask, pix[:, :, :, :100], sizes)


@pytest.mark.parametrize("case", ["full_bt", "full_gpm"])
def test_full_depth_vs_reference_golden(case, tmp_path_factory):
    """BASELINE.json configs[0] / configs[1] shapes at full depth (23-layer CLIP path, 32-layer decoder).
    (1) layer by layer, the engine is as close to the reference's fp32 hidden states as the reference's own bf16
        arithmetic is (relative L2 error of every decoder layer output);
    (2) rewards vs the reference's fp32 CPU goldens within 2e-2 + the measured bf16 noise of this network;
    (3) same preference decision."""
    if not os.path.exists(os.path.join(GOLDEN_DIR, f"{case}.pt")):
        pytest.skip("fixture not generated")
    for k in list(_models):
        del _models[k]
    torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = False
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    rewards, errs, floors = {}, [], []
    for bi, entry in enumerate(fx["batches"]):
        ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
        if bi == 0:
            model.engine.taps = {}
        r, _ = model.custom_forward(ids, mask, pix, sizes)
        taps_e, model.engine.taps = model.engine.taps, None
        errs.append((r.float().cpu() - entry["reward"]).abs().max().item())
        floors.append(bf16_noise_floor(fx, cfg, ids, mask, pix, sizes, entry["reward"]))
        print(f"{case}/{entry['tag']}: engine {r.float().flatten().tolist()} ref {entry['reward'].flatten().tolist()}"
              f" | engine-vs-fp32 {errs[-1]:.4g}, reference bf16-vs-fp32 max {floors[-1][0]:.4g} rms {floors[-1][1]:.4g}")
        rewards[entry["tag"]] = r
        if bi == 0:
            valid = mask.bool()
            B, S = ids.shape
            t32, t16 = {}, {}
            with torch.no_grad():
                P32 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.float32, device="cuda",
                               cache=False)
                r32 = O.custom_forward(P32, cfg, ids, mask, pix, sizes, t32)
                t32 = {k: v[valid].clone() for k, v in t32.items() if k.startswith("hidden_")}
                P16 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16,
                               device="cuda", cache=False)
                O.custom_forward(P16, cfg, ids, mask, pix, sizes, t16)
            assert (r32.cpu() - entry["reward"]).abs().max().item() < 2e-3  # fp32 on GPU reproduces the CPU golden
            worst = 0.0
            for li in range(cfg.num_layers):
                e_eng = rel_err(taps_e[f"hidden_{li}"].view(B, S, -1)[valid], t32[f"hidden_{li}"])
                e_ref = rel_err(t16[f"hidden_{li}"][valid], t32[f"hidden_{li}"])
                worst = max(worst, e_eng / e_ref)
                if li in (0, 7, 15, 23, 31):
                    print(f"  layer {li}: rel L2 err vs fp32: engine {e_eng:.4g}, reference bf16 {e_ref:.4g}")
                assert e_eng < 1.5 * e_ref + 2e-3, (li, e_eng, e_ref)
            print(f"  worst engine/reference error ratio over layers: {worst:.3f}")
            del t32, t16, taps_e
            torch.cuda.empty_cache()
    assert max(errs) < reward_gate(max(errs), floors), f"{case}: reward err {max(errs):.4g} vs reference fp32"
    prob = preference_compute(args, rewards["c"], rewards["r"])
    ref = fx["prob"].numpy()
    decided = abs(ref - 0.5) > 0.1
    assert ((prob > 0.5) == (ref > 0.5))[decided].all()
    _models.pop(case, None)
    torch.cuda.empty_cache()


def test_decision_agreement_many_pairs(tmp_path_factory):
    """Pairwise decisions (prob > 0.5) of the engine vs the reference arithmetic in fp32 on 48 synthetic pairs,
    reported next to the reference's own bf16-vs-fp32 flip rate (north_star: >= 99.9% identical decisions;
    pairs whose fp32 margin is inside the bf16 noise band are the only ones allowed to differ)."""
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    from llava_reward_b200.synth import synth_batch
    P32 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.float32, device="cuda")
    P16 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda")
    n_batches, B = 6, 8
    pe, p32, p16 = [], [], []
    for i in range(n_batches):
        rs = {}
        for tag in ("c", "r"):
            ids, mask, pix, sizes = synth_batch(cfg, B, (336, 672), None, seed=100 + i, tag=tag, device="cuda")
            re_, _ = model.custom_forward(ids, mask, pix, sizes)
            with torch.no_grad():
                rs[tag] = (re_, O.custom_forward(P32, cfg, ids, mask, pix, sizes),
                           O.custom_forward(P16, cfg, ids, mask, pix, sizes))
        pe.append(torch.from_numpy(preference_compute(args, rs["c"][0], rs["r"][0])))
        p32.append(O.preference_compute(cfg, rs["c"][1], rs["r"][1]).cpu())
        p16.append(O.preference_compute(cfg, rs["c"][2], rs["r"][2]).cpu())
    pe, p32, p16 = torch.cat(pe), torch.cat(p32), torch.cat(p16)
    agree_engine = ((pe > 0.5) == (p32 > 0.5)).float().mean().item()
    agree_ref16 = ((p16 > 0.5) == (p32 > 0.5)).float().mean().item()
    print(f"decision agreement over {pe.numel()} pairs: engine-vs-fp32 {agree_engine:.4f}, "
          f"reference bf16-vs-fp32 {agree_ref16:.4f}; max |dprob| engine {float((pe - p32).abs().max()):.3g} "
          f"ref16 {float((p16 - p32).abs().max()):.3g}")
    clear = (p32 - 0.5).abs() > 0.25     # margin well outside the bf16 noise band
    assert ((pe > 0.5) == (p32 > 0.5))[clear].all()
    assert agree_engine >= agree_ref16 - 0.05


def test_right_padding_single_sample_and_max_crops(tmp_path_factory):
    """Edge cases of the batch layout: right-padded mask (EOS is not the last column), B=1, and the 4x4-crop maximum
    image (17 crop slots all real, N_v=2509), against the oracle in fp32."""
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    from llava_reward_b200.synth import synth_batch
    P32 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.float32, device="cuda")
    P16 = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda")
    # (a) right padding: roll every row so the valid run starts at column 0
    ids, mask, pix, sizes = synth_batch(cfg, 3, (336, 672), None, seed=31, tag="rp", device="cuda",
                                        image_hw_list=[(336, 672), (672, 336), (336, 336)])
    ids_r, mask_r = ids.clone(), mask.clone()
    for b in range(ids.shape[0]):
        n_pad = int((mask[b] == 0).sum())
        ids_r[b] = torch.roll(ids[b], -n_pad)
        mask_r[b] = torch.roll(mask[b], -n_pad)
    assert int(mask_r[:, -1].sum()) < ids.shape[0]  # at least one row really is right-padded
    r_e, _ = model.custom_forward(ids_r, mask_r, pix, sizes)
    with torch.no_grad():
        r32 = O.custom_forward(P32, cfg, ids_r, mask_r, pix, sizes)
        r16 = O.custom_forward(P16, cfg, ids_r, mask_r, pix, sizes)
    floor = (r16.float() - r32).abs().max().item()
    err = (r_e.float() - r32).abs().max().item()
    print(f"right padding: engine-vs-fp32 {err:.4g}, reference bf16-vs-fp32 {floor:.4g}")
    assert err < REWARD_TOL + 3 * floor
    # (b) B = 1 with the maximum HD size
    ids, mask, pix, sizes = synth_batch(cfg, 1, (1344, 1344), None, seed=32, tag="mx", device="cuda")
    assert int((ids < 0).sum()) == 2509
    r_e, _ = model.custom_forward(ids, mask, pix, sizes)
    with torch.no_grad():
        r32 = O.custom_forward(P32, cfg, ids, mask, pix, sizes)
        r16 = O.custom_forward(P16, cfg, ids, mask, pix, sizes)
    floor = (r16.float() - r32).abs().max().item()
    err = (r_e.float() - r32).abs().max().item()
    print(f"B=1, 17 crops: engine-vs-fp32 {err:.4g}, reference bf16-vs-fp32 {floor:.4g}")
    assert err < REWARD_TOL + 3 * floor


def test_gpu_preprocessing_feeds_the_engine(tmp_path_factory):
    """uint8 image -> GPU preprocessing -> custom_forward: same reward as feeding the oracle-preprocessed pixels."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from preprocess_util import synth_image
    from oracle import preprocess_oracle as PO
    from llava_reward_b200.processing import Phi3VImageProcessorB200
    from llava_reward_b200.synth import BOS, USER, NL, EOS
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    img = synth_image("small_200x333", 200, 333)
    out = Phi3VImageProcessorB200(num_crops=16).preprocess([img], return_tensors="pt")
    ntok = int(out["num_img_tokens"][0])
    ids = torch.tensor([[BOS, USER, NL] + [-1] * ntok + [NL, 100, 200, 300, EOS]], device="cuda")
    mask = torch.ones_like(ids)
    r_gpu, _ = model.custom_forward(ids, mask, out["pixel_values"], out["image_sizes"])
    ref_pix, (h, w), ntok2 = PO.preprocess(img)
    assert ntok2 == ntok and [h, w] == out["image_sizes"][0].tolist()
    r_ref, _ = model.custom_forward(ids, mask, torch.from_numpy(ref_pix)[None].cuda(), torch.tensor([[h, w]]))
    # the two pixel tensors differ by <= 1e-5 in the bicubic global view only; after the bf16 conversion a few values
    # land on the other side of a rounding boundary, i.e. the two forwards are two bf16 noise realizations of the same
    # function (rms 0.02-0.03 on these weights), not bit-equal inputs
    assert (r_gpu.float() - r_ref.float()).abs().max().item() < 3e-2


def test_batch_eval_loops(tmp_path_factory):
    """Pairwise and single-image eval loops (reference eval/batch_inference_rm_phi.py:70-152) over ragged samples:
    collated left-padded batches give the same rewards as scoring each sample alone (BT model: batch-invariant)."""
    from llava_reward_b200.batch_eval import collate_samples, score_pairs, score_single
    from llava_reward_b200.synth import synth_batch, PAD
    fx = load_fixture("slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)

    def samples(tag, n):
        out = []
        for i in range(n):
            ids, mask, pix, sizes = synth_batch(cfg, 1, (336, 336 * (1 + i % 2)), None, seed=50 + i, tag=f"{tag}{i}",
                                                device="cuda", text_len_range=(5, 40))
            out.append({"input_ids": ids, "attention_mask": mask, "pixel_values": pix, "image_sizes": sizes})
        return out

    ch, rj = samples("c", 4), samples("r", 4)
    batches = [(collate_samples(ch[i:i + 2], PAD), collate_samples(rj[i:i + 2], PAD)) for i in (0, 2)]
    res = score_pairs(model, args, batches)
    assert res["probs"].shape == (4,) and 0.0 <= res["proportion"] <= 1.0
    alone_c = [model.custom_forward(**s)[0].float().item() for s in ch]
    assert max(abs(a - b) for a, b in zip(alone_c, res["chosen_rewards"])) < REWARD_TOL
    single = score_single(model, args, [(collate_samples(ch, PAD), torch.tensor([1, 0, 1, 0]))], cls_based=True)
    assert len(single["rewards"]) == 4 and 0.0 <= single["accuracy"] <= 1.0
    gfx = load_fixture("slim_gpm")
    gargs, gmodel, _ = build_model(gfx, tmp_path_factory)
    with pytest.raises(ValueError):
        score_single(gmodel, gargs, [])


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt"])
def test_last_layer_row_shortcut_is_output_identical(case, tmp_path_factory):
    """o_proj / MLP of the LAST decoder layer on the last-valid-token rows only (engine.last_layer_rows, the product
    default) against the same layer run on every row: the value the head consumes must not change."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    ids, mask, pix, sizes = fixture_batch(fx, fx["batches"][0], cfg, device="cuda")
    eng = model.engine
    assert eng.last_layer_rows
    r_short, _ = model.custom_forward(ids, mask, pix, sizes)
    n_short = eng.launches
    eng.last_layer_rows = False
    try:
        r_full, _ = model.custom_forward(ids, mask, pix, sizes)
    finally:
        eng.last_layer_rows = True
    d = (r_short.float() - r_full.float()).abs().max().item()
    print(f"{case}: shortcut {r_short.flatten().tolist()} full {r_full.flatten().tolist()} |d| {d:.3g}; "
          f"launches {n_short} vs {eng.launches}")
    assert d <= 4e-3   # same arithmetic per row; only the GEMM tile shape (small-M kernel) may differ in fp32 summation


ATTR_VARIANTS = {"layer_id_1": dict(layer_id=1), "layer_id_0": dict(layer_id=0), "training": dict(training=True),
                 "mean": dict(mean_hidden_state=True), "mean_layer_id_1": dict(mean_hidden_state=True, layer_id=1)}


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt"])
def test_attribute_variants_vs_reference_golden(case, tmp_path_factory):
    """`layer_id`, `training` and `mean_hidden_state` - the attributes the reference's custom_forward reads
    (rw_model_general_preference.py:327-333) - set on the model object exactly as on the reference's; goldens were
    made by setting them on the reference model (tests/golden/make_golden.py). mean_hidden_state runs the all-rows
    SkipCA (per-sample GEMMs + lr_softmax_rows_bf16) and lr_masked_mean_rows_bf16."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
    P = O.Params(SynthProvider(cfg, seed=fx["seed_w"], device="cuda"), dtype=torch.bfloat16, device="cuda", cache=False)
    saved = {k: getattr(model, k) for k in ("layer_id", "training", "mean_hidden_state")}
    try:
        for key, attrs in ATTR_VARIANTS.items():
            for k, v in saved.items():
                setattr(model, k, v)
            for k, v in attrs.items():
                setattr(model, k, v)
            r, _ = model.custom_forward(ids, mask, pix, sizes)
            g = entry["attrs"][key]
            assert tuple(r.shape) == tuple(g.shape), key
            with torch.no_grad():
                ro = O.custom_forward(P, cfg, ids, mask, pix, sizes, **attrs).float().cpu()
            err, floor = (r.float().cpu() - g).abs().max().item(), (ro - g).abs().max().item()
            print(f"{case}/{key}: engine {r.float().flatten().tolist()} ref {g.flatten().tolist()} | engine-vs-fp32 "
                  f"{err:.4g}, reference bf16-vs-fp32 {floor:.4g}")
            assert err < REWARD_TOL + 3.0 * floor, key
    finally:
        for k, v in saved.items():
            setattr(model, k, v)
    # training-mode gather = eval gather on a left-padded batch (position S-1 is the last valid token)
    assert bool((mask[:, -1] == 1).all())


def test_gather_rows_negative_index_is_zero_row():
    from llava_reward_b200 import ops
    src = torch.randn(5, 64, device="cuda").to(torch.bfloat16)
    idx = torch.tensor([3, -1, 0, -1], dtype=torch.int32, device="cuda")
    dst = torch.full((4, 64), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gather_rows(src, idx, dst, 4, 64)
    torch.cuda.synchronize()
    assert torch.equal(dst[0], src[3]) and torch.equal(dst[2], src[0])
    assert (dst[1] == 0).all() and (dst[3] == 0).all()


@pytest.mark.parametrize("rows,n_valid,n_total", [(7, 100, 256), (33, 1921, 2048), (5, 8, 8)])
def test_softmax_rows_and_masked_mean_kernels(rows, n_valid, n_total):
    from llava_reward_b200 import ops
    torch.manual_seed(rows)
    s = (torch.randn(rows, n_total, device="cuda") * 30).to(torch.bfloat16)
    ref = torch.softmax((s[:, :n_valid] / 55.42562584220407).float().to(torch.bfloat16), dim=-1)  # sqrt(3072)
    out = s.clone()
    ops.softmax_rows(out, rows, n_valid, n_total, 1.0 / 55.42562584220407)
    torch.cuda.synchronize()
    assert (out[:, n_valid:] == 0).all()
    assert (out[:, :n_valid].float() - ref.float()).abs().max().item() <= 2 ** -8 * ref.float().max().item() + 1e-6
    B, S, H = 3, 37, 512
    x = torch.randn(B * S, H, device="cuda").to(torch.bfloat16)
    m = torch.zeros(B, S, dtype=torch.int64, device="cuda")
    m[0, 5:] = 1
    m[1, :] = 1
    m[2, 30:] = 1
    pooled = torch.empty(B, H, device="cuda", dtype=torch.bfloat16)
    ops.masked_mean_rows(x, m, pooled, B, S, H)
    xb = x.view(B, S, H)
    mb = m.to(torch.bfloat16).unsqueeze(-1)
    want = (xb * mb).sum(dim=1) / mb.sum(dim=1).clamp(min=1e-8)
    torch.cuda.synchronize()
    assert (pooled.float() - want.float()).abs().max().item() <= 2 ** -7 * want.float().abs().max().item()


def test_return_output_hidden_states(tmp_path_factory):
    """custom_forward(return_output=True): the reference returns its BaseModelOutputWithPast; the golden fixture holds
    strided samples of hidden_states[0], [1], last_hidden_state and the zero-padded vision_embeds ([-1])."""
    fx = load_fixture("slim_gpm")
    args, model, cfg = build_model(fx, tmp_path_factory)
    entry = fx["batches"][0]
    ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
    r0, none = model.custom_forward(ids, mask, pix, sizes)
    r, out = model.custom_forward(ids, mask, pix, sizes, return_output=True)
    assert none is None and model.engine.taps is None
    assert (r.float() - r0.float()).abs().max().item() < 2e-2   # the capture path runs the un-shortcut last layer
    hs = out["hidden_states"]
    assert len(hs) == cfg.num_layers + 2 and out["last_hidden_state"] is hs[-2]
    valid = mask.bool()[:, :, None]
    for name, t in (("inputs_embeds", hs[0]), ("hidden_0", hs[1]), ("last_hidden", out.last_hidden_state),
                    ("vision_embeds", hs[-1])):
        g = entry["taps"][name]
        assert list(t.shape) == g["shape"], name
        tt = t if name == "vision_embeds" else torch.where(valid.expand_as(t), t, torch.zeros_like(t))
        a = tt.float().flatten()[:: g["stride"]][:2048].cpu()
        w = torch.ones_like(a) if name == "vision_embeds" else \
            valid.expand_as(t).float().flatten()[:: g["stride"]][:2048].cpu()
        rel = ((a - g["vals"] * w).norm() / (g["vals"] * w).norm()).item()
        print(f"return_output {name}: rel L2 err vs reference fp32 {rel:.4g}")
        assert rel < 3e-2, name


def test_device_prefetcher_feeds_the_eval_loop(tmp_path_factory):
    """feed.DevicePrefetcher: CPU batches of any nesting arrive on the device unchanged and in order (pinned staging
    sets reused round-robin, copies on a side stream), and score_pairs over the prefetcher gives exactly the rewards of
    the blocking loop."""
    from llava_reward_b200.batch_eval import collate_samples, score_pairs
    from llava_reward_b200.feed import DevicePrefetcher
    from llava_reward_b200.synth import synth_batch, PAD
    raw = [{"a": torch.full((3, 5), float(i)), "nest": (torch.arange(4) + i, [torch.ones(2, 2) * i]), "tag": f"b{i}"}
           for i in range(5)]
    pf = DevicePrefetcher(raw, device="cuda")
    seen = list(pf)
    assert len(seen) == 5 and pf.h2d_bytes == 5 * (15 * 4 + 4 * 8 + 4 * 4)
    for i, b in enumerate(seen):
        assert b["tag"] == f"b{i}" and b["a"].is_cuda and b["nest"][1][0].is_cuda
        assert torch.equal(b["a"].cpu(), raw[i]["a"]) and torch.equal(b["nest"][0].cpu(), raw[i]["nest"][0])
    assert list(DevicePrefetcher([], device="cuda")) == []

    fx = load_fixture("slim_bt")
    args, model, cfg = build_model(fx, tmp_path_factory)

    def samples(tag, n):
        out = []
        for i in range(n):
            ids, mask, pix, sizes = synth_batch(cfg, 1, (336, 336 * (1 + i % 2)), None, seed=50 + i, tag=f"{tag}{i}",
                                                text_len_range=(5, 40))
            out.append({"input_ids": ids, "attention_mask": mask, "pixel_values": pix, "image_sizes": sizes})
        return out

    ch, rj = samples("c", 6), samples("r", 6)
    cpu_batches = [(collate_samples(ch[i:i + 2], PAD), collate_samples(rj[i:i + 2], PAD)) for i in (0, 2, 4)]
    dev_batches = [tuple({k: v.cuda() for k, v in b.items()} for b in pair) for pair in cpu_batches]
    blocking = score_pairs(model, args, dev_batches)
    prefetched = score_pairs(model, args, DevicePrefetcher(cpu_batches, device="cuda"))
    assert blocking["chosen_rewards"] == prefetched["chosen_rewards"]
    assert blocking["reject_rewards"] == prefetched["reject_rewards"]
    assert (blocking["probs"] == prefetched["probs"]).all()


@pytest.mark.parametrize("case", ["slim_gpm", "slim_bt"])
def test_packed_valid_rows_are_output_identical(case, tmp_path_factory):
    """engine.pack_rows (product default): the decoder runs on the valid rows only, packed back to back, with the
    packed-sequence attention layout. Every decoder kernel is row-wise and the attention walks the same K/V blocks from
    the start of each sample in both layouts, so the rewards must be BIT-identical to the slot layout - for left and
    right padding, mixed lengths and a batch without any padding (where packing switches itself off)."""
    from llava_reward_b200.synth import synth_batch
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    eng = model.engine
    batches = [fixture_batch(fx, e, cfg, device="cuda") for e in fx["batches"]]
    ids, mask, pix, sizes = synth_batch(cfg, 3, (336, 672), 900, seed=91, tag="pk", device="cuda",
                                        image_hw_list=[(336, 672), (672, 672), (336, 336)], text_len_range=(3, 90))
    batches.append((ids, mask, pix, sizes))
    # the same samples right-padded: roll every row so that its valid run starts at column 0
    S = ids.shape[1]
    lens = mask.sum(1)
    ids_r, mask_r = ids.clone(), mask.clone()
    for b in range(ids.shape[0]):
        n = int(lens[b])
        ids_r[b] = torch.cat([ids[b, S - n:], ids[b, :S - n]])
        mask_r[b] = torch.cat([mask[b, S - n:], mask[b, :S - n]])
    batches.append((ids_r, mask_r, pix, sizes))
    one = synth_batch(cfg, 1, (336, 336), None, seed=92, tag="np", device="cuda")   # no padding at all
    batches.append(one)
    try:
        for i, (ids, mask, pix, sizes) in enumerate(batches):
            eng.pack_rows = True
            rp = model.custom_forward(ids, mask, pix, sizes)[0].clone()
            n_packed = eng.launches
            eng.pack_rows = False
            rs = model.custom_forward(ids, mask, pix, sizes)[0].clone()
            d = (rp.float() - rs.float()).abs().max().item()
            print(f"{case} batch {i}: valid rows {int(mask.sum())} of {mask.numel()}, packed {rp.flatten().tolist()} "
                  f"slot {rs.flatten().tolist()} |d| {d:.3g}, launches {n_packed} / {eng.launches}")
            assert torch.equal(rp, rs)
    finally:
        eng.pack_rows = True
    # left- and right-padded forms of the same samples agree as well (positions count from the first valid token)


def test_vision_layer_id_vs_reference_golden(tmp_path_factory):
    """`vision_layer_id` (rw_model_general_preference.py:353): SkipCA keys / values taken from
    hidden_states[vision_layer_id][:, :N_v_max] instead of the projected image tokens. Goldens: the attribute set on the
    reference model (tests/golden/make_golden.py: slim_gpm_b1, batches of one sample = no padded positions)."""
    fx = load_fixture("slim_gpm_b1")
    args, model, cfg = build_model(fx, tmp_path_factory)
    assert model.vision_layer_id == -1
    try:
        for entry in fx["batches"]:
            ids, mask, pix, sizes = fixture_batch(fx, entry, cfg, device="cuda")
            model.vision_layer_id = -1
            r0, _ = model.custom_forward(ids, mask, pix, sizes)
            base_err = (r0.float().cpu() - entry["reward"]).abs().max().item()
            for key, g in entry["attrs"].items():
                model.vision_layer_id = int(key.rsplit("_", 1)[1])
                r, _ = model.custom_forward(ids, mask, pix, sizes)
                err = (r.float().cpu() - g).abs().max().item()
                print(f"{entry['tag']}/{key}: engine {r.float().flatten().tolist()} ref {g.flatten().tolist()} err {err:.4g} "
                      f"(default path err {base_err:.4g})")
                assert tuple(r.shape) == tuple(g.shape)
                assert err < REWARD_TOL + 2.0 * base_err, key
            assert model.engine.taps is None
            # index cfg.num_layers + 1 is the vision_embeds entry itself (= -1)
            model.vision_layer_id = cfg.num_layers + 1
            r1, _ = model.custom_forward(ids, mask, pix, sizes)
            assert torch.equal(r1, r0)
    finally:
        model.vision_layer_id = -1


def test_device_prefetcher_pinned_memory_is_bounded_for_ragged_batches():
    """ragged inputs (a new shape almost every batch) must not grow the pinned staging memory without
    bound - one byte arena per staging set, grown to the largest batch seen."""
    from llava_reward_b200.feed import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    sizes = [int(x) for x in torch.randint(1000, 200000, (40,), generator=g)]
    batches = [{"a": torch.randn(n, generator=g), "b": (torch.arange(n // 7, dtype=torch.int64), "meta")} for n in sizes]
    pf = DevicePrefetcher(batches, device="cuda", depth=2)
    seen = 0
    for ref, out in zip(batches, pf):
        assert out["a"].is_cuda and torch.equal(out["a"].cpu(), ref["a"]) and torch.equal(out["b"][0].cpu(), ref["b"][0])
        assert out["b"][1] == "meta"
        seen += 1
    assert seen == len(batches)
    biggest = max(n * 4 + (n // 7) * 8 for n in sizes)
    assert pf.pinned_bytes <= 2 * (1.25 * biggest + 1024), (pf.pinned_bytes, biggest)
    assert pf.h2d_bytes == sum(n * 4 + (n // 7) * 8 for n in sizes)


@pytest.mark.parametrize("case", ["slim_bt", "slim_gpm"])
def test_all_padding_sample_does_not_disturb_the_batch(case, tmp_path_factory):
    """Degenerate input: one sample of a ragged batch has an all-zero attention mask (the reference then reads row
    S - 1 of that sample, rw_model_general_preference.py:420-421; its value depends on the attention implementation -
    uniform softmax in the eager path, an empty sequence under flash-attention - so only finiteness is asserted for
    it). The other samples must come out bit-identical to the same batch with a normal middle sample: the slot
    layout the engine falls back to for such a batch (no rows to pack for that sample) is output-identical."""
    fx = load_fixture(case)
    args, model, cfg = build_model(fx, tmp_path_factory)
    from llava_reward_b200.synth import synth_batch
    ids, mask, pix, sizes = synth_batch(cfg, 3, (336, 672), None, seed=41, tag="dg", device="cuda",
                                        image_hw_list=[(336, 672), (672, 336), (336, 672)])
    r_ok, _ = model.custom_forward(ids, mask, pix, sizes)
    mask_d = mask.clone()
    mask_d[1] = 0
    r_d, _ = model.custom_forward(ids, mask_d, pix, sizes)
    assert torch.isfinite(r_d.float()).all()
    assert torch.equal(r_d[0], r_ok[0]) and torch.equal(r_d[2], r_ok[2]), (r_d, r_ok)
    # and the whole call is repeatable (no state left behind by the degenerate batch)
    r_d2, _ = model.custom_forward(ids, mask_d, pix, sizes)
    assert torch.equal(r_d2, r_d)
