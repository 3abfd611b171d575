"""BASELINE.json configs[0] as written, on the GPU: the two JPEGs of reference eval/simple_inference.py:22
(tests/golden/sample_test = the reference's data/sample_test) decoded by PIL -> pinned uint8 ->
`Phi3VImageProcessorB200` (GPU resize / pad / normalise / crop) -> engine (BT head, no SkipCA, no LoRA) -> reward and
preference probability, against the golden made by the reference's own processor + `custom_forward` in fp32
(tests/golden/make_golden_real.py) and against the reference model in bf16 on the same GPU. Then the f2 callers: the
reference's sample manifests through `GeneralRewardDataset` -> `manifest_batches` -> `DevicePrefetcher` ->
`score_pairs` / `score_single`."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)

from golden_util import load_fixture  # noqa: E402
from make_golden_real import prompt_ids  # noqa: E402
from stub_tokenizer import StubPhi3Tokenizer  # noqa: E402

from llava_reward_b200.batch_eval import score_pairs, score_single  # noqa: E402
from llava_reward_b200.config import RewardConfig  # noqa: E402
from llava_reward_b200.datasets import GeneralRewardDataset, decode_rgb, load_manifest, manifest_batches  # noqa: E402
from llava_reward_b200.feed import DevicePrefetcher  # noqa: E402
from llava_reward_b200.processing import Phi3VImageProcessorB200, Phi3VProcessorB200  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

_models = {}


def bt_model(case, tmp_path_factory):
    if case not in _models:
        fx = load_fixture(case)
        d = tmp_path_factory.mktemp(case)
        ypath = os.path.join(d, "reward_config.yaml")
        with open(ypath, "w") as f:
            yaml.safe_dump({"is_general_preference": False, "add_cross_attention": False, "value_head_dim": 1,
                            "general_preference_tau": 0.1}, f)
        over = {k: v for k, v in fx["cfg_overrides"].items() if k in ("num_layers", "clip_layers", "use_lora")}
        args = types.SimpleNamespace(pretrain=f"synthetic:{fx['seed_w']}", pm_path=None, cache_dir=None,
                                     ft_projector=False, config_overrides=over)
        args, model = load_reward_adaptor(args, "phi3v", ypath)
        _models.clear()
        _models[case] = (args, model.to("cuda").eval(), fx)
    return _models[case]


@pytest.mark.parametrize("case", ["real_slim_bt", "real_full_bt"])
def test_config0_real_jpegs_end_to_end(case, tmp_path_factory):
    if not os.path.exists(os.path.join(GOLD, f"{case}.pt")):
        pytest.skip(f"{case}.pt not generated")
    args, model, fx = bt_model(case, tmp_path_factory)
    cfg = RewardConfig(**fx["cfg_overrides"])
    proc = Phi3VImageProcessorB200(num_crops=16)
    ref_model = ref_proc = None
    if RH.available():
        RH.import_reference()
        from llava_reward.models.base_mllm.phi3_v.processing_phi3_v import Phi3VImageProcessor
        from PIL import Image
        ref_proc = Phi3VImageProcessor(num_crops=16)
        ref_model = RH.build_reference_model(cfg, fx["seed_w"], device="cuda", dtype=torch.bfloat16, verbose=False)
        RH.set_attention(ref_model, "flash_attention_2")
    rewards, ref_rewards = [], []
    for smp in fx["samples"]:
        path = os.path.join(GOLD, smp["image"])
        u8 = torch.from_numpy(decode_rgb(path)).pin_memory()              # PIL decode -> pinned uint8 HWC
        assert list(u8.shape[1::-1]) == smp["pil_size"]
        out = proc.preprocess([u8], return_tensors="pt")
        pv = out["pixel_values"]
        assert out["image_sizes"][0].tolist() == smp["image_sizes"] and int(out["num_img_tokens"][0]) == smp["num_img_tokens"]
        assert (pv.flatten()[::997].cpu() - smp["pixel_sample"]).abs().max().item() < 1e-5
        assert abs(pv.double().sum().item() - smp["pixel_sum"]) < 1e-6 * smp["pixel_abs_sum"]
        ids = torch.tensor([prompt_ids(smp["num_img_tokens"])], dtype=torch.int64, device="cuda")
        assert ids.shape[1] == smp["S"]
        mask = torch.ones_like(ids)
        r, _ = model.custom_forward(ids, mask, pv, out["image_sizes"])
        assert tuple(r.shape) == (1, 1)
        rewards.append(r)
        if ref_model is not None:
            rp = ref_proc.preprocess(Image.open(path).convert("RGB"), return_tensors="pt")
            pr = rp["pixel_values"]
            assert torch.equal(pv[0, 1:].cpu(), pr[0, 1:])                 # every crop bit-exact on a REAL decoded JPEG
            assert (pv[0, 0].cpu() - pr[0, 0]).abs().max().item() < 1e-5    # bicubic global view
            with torch.no_grad():
                rr, _ = ref_model.custom_forward(ids, mask, pr.cuda(), rp["image_sizes"].cuda())
            ref_rewards.append(rr)
    errs = [abs(float(r) - float(s["reward"])) for r, s in zip(rewards, fx["samples"])]
    ref_errs = [abs(float(r) - float(s["reward"])) for r, s in zip(ref_rewards, fx["samples"])]
    pair = [abs(float(a) - float(b)) for a, b in zip(rewards, ref_rewards)]
    print(f"{case}: engine {[float(r) for r in rewards]} reference fp32 {[float(s['reward']) for s in fx['samples']]} | "
          f"|engine-fp32| {errs} | |reference_bf16-fp32| {ref_errs} | |engine-reference_bf16| {pair}")
    floor = max(ref_errs) if ref_errs else 0.0
    assert max(errs) <= max(2e-2, 1.5 * floor)
    prob = preference_compute(args, rewards[0], rewards[1])
    print(f"{case}: prob {prob.tolist()} reference fp32 {fx['prob'].tolist()}")
    assert prob.shape == (1,)
    # the decision is only defined above the bf16 noise of the two rewards it is made of (full depth: reference fp32
    # rewards 0.9002 / 0.8606, i.e. a margin of 0.04 against a per-reward bf16 error of ~0.02 for the reference itself)
    gap = abs(float(fx["samples"][0]["reward"]) - float(fx["samples"][1]["reward"]))
    if gap > 2.0 * max(errs + ref_errs) + 1e-3:
        assert (prob[0] > 0.5) == (float(fx["prob"][0]) > 0.5)
    # d prob / d (r_c - r_r) <= 1 / (4 tau) = 2.5; two rewards; + one bf16 rounding of the probability itself
    assert abs(float(prob[0]) - float(fx["prob"][0])) <= 0.01 + 5.0 * max(errs)


def test_manifest_reader_feeds_score_pairs_and_score_single(tmp_path_factory):
    """pairwise_sample.json / non_pairwise_sample.json -> dataset -> threaded decode -> prefetcher -> eval loops; the
    batched results equal scoring every image alone (no SkipCA: a sample's reward does not depend on its batch)."""
    args, model, fx = bt_model("real_slim_bt", tmp_path_factory)
    tok = StubPhi3Tokenizer()
    processor = Phi3VProcessorB200(Phi3VImageProcessorB200(num_crops=16), tok)
    root = os.path.join(GOLD)                       # manifests say data/sample_test/...: map 'data' -> tests/golden
    link = tmp_path_factory.mktemp("root")
    os.symlink(os.path.join(GOLD, "sample_test"), os.path.join(link, "sample_test"))
    os.makedirs(os.path.join(link, "data"))
    os.symlink(os.path.join(GOLD, "sample_test"), os.path.join(link, "data", "sample_test"))
    del root
    rows = load_manifest(os.path.join(GOLD, "sample_test", "pairwise_sample.json"))
    ds = GeneralRewardDataset(rows, processor=processor, tokenizer=tok, image_root=str(link))
    res = score_pairs(model, args, DevicePrefetcher(manifest_batches(ds, 3, decode_threads=2)))
    assert res["probs"].shape == (len(rows),) and len(res["chosen_rewards"]) == len(rows)
    assert 0.0 <= res["proportion"] <= 1.0
    # every sample alone
    for i in range(len(rows)):
        it = ds[i]
        rc, _ = model.custom_forward(it[0], it[1], it[2], it[3])
        rr, _ = model.custom_forward(it[4], it[5], it[6], it[7])
        assert abs(float(rc) - res["chosen_rewards"][i]) <= 2e-2 and abs(float(rr) - res["reject_rewards"][i]) <= 2e-2
        p = preference_compute(args, rc, rr)
        assert abs(float(p[0]) - float(res["probs"][i])) <= 0.1
    # the first manifest row is the pair of eval/simple_inference.py:22 with the real caption instead of the seeded ids
    assert rows[0]["chosen_path"].endswith("0_1_id_000904-0035.jpg") and rows[0]["reject_path"].endswith("4_3_id_000904-0035.jpg")
    rows1 = load_manifest(os.path.join(GOLD, "sample_test", "non_pairwise_sample.json"))
    ds1 = GeneralRewardDataset(rows1, processor=processor, tokenizer=tok, cls_based=True, image_root=str(link))
    res1 = score_single(model, args, DevicePrefetcher(manifest_batches(ds1, 2)), cls_based=True)
    assert len(res1["rewards"]) == len(rows1) and res1["labels"] == [r["label"] for r in rows1]
    assert set(res1) >= {"accuracy", "f1", "recall"}
    assert np.isfinite(res1["rewards"]).all()
