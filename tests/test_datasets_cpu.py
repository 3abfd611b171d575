"""`llava_reward_b200.datasets.GeneralRewardDataset` against the reference's own class
(baseline/_ref/llava_reward/datasets/reward_dataset.py:25-202) on the reference's sample manifests
(tests/golden/sample_test = data/sample_test), both driven by the same stub processor / tokenizer: identical item
tuples, identical collated batches (left padding, stacking), identical micro-batch order from `manifest_batches`."""
import os
import types

import pytest
import torch

from stub_tokenizer import StubPhi3Tokenizer

from llava_reward_b200.datasets import (GeneralRewardDataset, is_non_pairwise, load_manifest, manifest_batches,
                                        squeeze_batch)
from oracle import ref_harness as RH

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class StubProcessor:
    """text -> ids through the tokenizer, image -> a [1,17,3,2,2] tensor that depends on the pixels (so a swapped or
    mis-ordered image is visible), image_sizes from the real shape"""

    def __init__(self, tok):
        self.tok = tok

    def __call__(self, text, images, return_tensors="pt"):
        import numpy as np
        ids = torch.tensor([self.tok(text).input_ids])
        img = np.asarray(images[0])
        pv = torch.full((1, 17, 3, 2, 2), float(img.astype(np.float64).mean()))
        return {"input_ids": ids, "attention_mask": torch.ones_like(ids), "pixel_values": pv,
                "image_sizes": torch.tensor([[img.shape[0], img.shape[1]]])}


def _cwd_data():
    return os.path.join(DATA)   # manifest paths are 'data/sample_test/...'; tests/golden/sample_test mirrors that tree


@pytest.fixture()
def in_data_root(monkeypatch, tmp_path):
    # the reference resolves manifest paths against the working directory: give it one where data/sample_test exists
    os.symlink(os.path.join(DATA, "sample_test"), tmp_path / "sample_test_link")
    os.makedirs(tmp_path / "data")
    os.symlink(os.path.join(DATA, "sample_test"), tmp_path / "data" / "sample_test")
    monkeypatch.chdir(tmp_path)
    return str(tmp_path)


def _eq(a, b):
    if torch.is_tensor(a):
        return torch.is_tensor(b) and a.shape == b.shape and torch.equal(a, b)
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_eq(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_eq(x, y) for x, y in zip(a, b))
    return a == b


@pytest.mark.skipif(not RH.available(), reason="baseline/_ref absent")
@pytest.mark.parametrize("manifest", ["pairwise_sample.json", "non_pairwise_sample.json"])
def test_dataset_matches_reference_class(manifest, in_data_root):
    RH.import_reference()
    from llava_reward.datasets.reward_dataset import GeneralRewardDataset as RefDataset
    rows = load_manifest(os.path.join(DATA, "sample_test", manifest))
    cls_based = is_non_pairwise(rows)
    assert cls_based == (manifest.startswith("non_"))
    tok = StubPhi3Tokenizer()
    proc = StubProcessor(tok)
    strategy = types.SimpleNamespace(is_rank_0=lambda: False)
    ref = RefDataset(rows, processor=proc, tokenizer=tok, strategy=strategy, cls_based=cls_based)
    mine = GeneralRewardDataset(rows, processor=proc, tokenizer=tok, strategy=None, cls_based=cls_based,
                                image_root=in_data_root)
    assert len(ref) == len(mine) == len(rows)
    items_r = [ref[i] for i in range(len(ref))]
    items_m = [mine[i] for i in range(len(mine))]
    assert _eq(items_r, items_m)
    # ragged prompts so that the left padding is exercised
    for bs in (1, 2, 3):
        for s in range(0, len(rows), bs):
            assert _eq(ref.collate_fn(items_r[s:s + bs]), mine.collate_fn(items_m[s:s + bs]))
    # the threaded feed yields the same micro-batches, squeezed as the eval loop does
    got = list(manifest_batches(mine, 3, decode_threads=2))
    for k, s in enumerate(range(0, len(rows), 3)):
        want = ref.collate_fn(items_r[s:s + 3])
        if cls_based:
            assert _eq(got[k], (squeeze_batch(want[0]), want[1]))
        else:
            assert _eq(got[k], (squeeze_batch(want[0]), squeeze_batch(want[1])))


def test_prompt_frame_and_padding(in_data_root):
    rows = load_manifest(os.path.join(DATA, "sample_test", "pairwise_sample.json"), max_samples=3)
    assert len(rows) == 3 and not is_non_pairwise(rows)
    rows[1] = dict(rows[1], prompt=["a short one", "a considerably longer prompt for the rejected image here"])
    tok = StubPhi3Tokenizer()
    ds = GeneralRewardDataset(rows, processor=StubProcessor(tok), tokenizer=tok, image_root=in_data_root)
    text = ds.prompt_text("hello world")
    assert text == "<|user|>\n<|image_1|>\nhello world<|endoftext|>"     # generation prompt cut, eos appended (:82-88)
    bc, br, c_rates, r_rates = ds.collate_fn([ds[i] for i in range(3)])
    assert bc["input_ids"].shape[:2] == (3, 1) and bc["pixel_values"].shape == (3, 1, 17, 3, 2, 2)
    assert c_rates == [r["c_rate"] for r in rows]
    short = bc["input_ids"][1, 0]
    n_pad = int((bc["attention_mask"][1, 0] == 0).sum())
    assert n_pad > 0 and (short[:n_pad] == tok.pad_token_id).all() and (bc["attention_mask"][1, 0, n_pad:] == 1).all()
    n_pad_r = int((br["attention_mask"][1, 0] == 0).sum())
    assert n_pad_r < n_pad                                               # per-image prompts are tokenised separately
