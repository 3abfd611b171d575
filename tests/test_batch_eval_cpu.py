"""Host-side batch-eval helpers vs the reference's own `zero_pad_sequences` semantics (left padding, pad id / 0) and
sklearn metrics (what eval/batch_inference_rm_phi.py:142-152 prints)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from llava_reward_b200.batch_eval import binary_metrics, collate_samples, zero_pad_sequences


def test_left_padding_matches_reference_semantics():
    seqs = [torch.arange(1, 6)[None], torch.arange(1, 3)[None], torch.arange(1, 9)[None]]
    out = zero_pad_sequences(seqs, value=32000)
    assert out.shape == (3, 1, 8)
    assert out[1, 0].tolist() == [32000] * 6 + [1, 2] and out[2, 0].tolist() == list(range(1, 9))
    right = zero_pad_sequences(seqs, side="right")
    assert right[1, 0].tolist() == [1, 2] + [0] * 6


def test_collate_samples_layout():
    items = []
    for n in (4, 7):
        items.append({"input_ids": torch.arange(n)[None], "attention_mask": torch.ones(1, n, dtype=torch.long),
                      "pixel_values": torch.zeros(1, 17, 3, 4, 4), "image_sizes": torch.tensor([[336, 672]])})
    b = collate_samples(items, pad_token_id=9)
    assert b["input_ids"].shape == (2, 7) and b["attention_mask"].shape == (2, 7)
    assert b["input_ids"][0].tolist() == [9, 9, 9, 0, 1, 2, 3] and b["attention_mask"][0].tolist() == [0, 0, 0, 1, 1, 1, 1]
    assert b["pixel_values"].shape == (2, 17, 3, 4, 4) and b["image_sizes"].tolist() == [[336, 672], [336, 672]]


def test_binary_metrics_match_sklearn():
    from sklearn.metrics import f1_score, recall_score
    rng = np.random.default_rng(0)
    label, pred = rng.integers(0, 2, 200), rng.integers(0, 2, 200)
    m = binary_metrics(pred, label)
    assert abs(m["f1"] - f1_score(label, pred, average="binary")) < 1e-12
    assert abs(m["recall"] - recall_score(label, pred)) < 1e-12
    assert abs(m["accuracy"] - (pred == label).mean()) < 1e-12


def test_feed_map_tensors_keeps_structure():
    """feed._map_tensors walks dict / BatchFeature / tuple / list batches and leaves non-tensors alone; the prefetcher
    itself refuses a CPU target (no CPU path in this package)."""
    import pytest
    import torch
    from transformers import BatchFeature
    from llava_reward_b200.feed import DevicePrefetcher, _map_tensors
    b = BatchFeature({"input_ids": torch.arange(6).view(2, 3), "pixel_values": torch.ones(2, 2)})
    nested = ({"x": torch.zeros(2), "n": 3, "s": "keep"}, [torch.ones(1), (torch.ones(2), None)], b)
    out = _map_tensors(nested, lambda t: t + 1)
    assert out[0]["n"] == 3 and out[0]["s"] == "keep" and torch.equal(out[0]["x"], torch.ones(2))
    assert torch.equal(out[1][0], torch.full((1,), 2.0)) and out[1][1][1] is None
    assert isinstance(out[2], BatchFeature) and torch.equal(out[2]["input_ids"], b["input_ids"] + 1)
    with pytest.raises(RuntimeError):
        DevicePrefetcher([], device="cpu")


def test_packed_row_plan():
    """config.packed_row_plan: the host index plan of the packed decoder layout (left / right padding, full rows)."""
    import numpy as np
    import pytest
    from llava_reward_b200.config import packed_row_plan
    S = 8
    start, length = [3, 0, 0, 6], [5, 8, 2, 2]          # left-padded, full, right-padded, left-padded
    idx, pos, base, last = packed_row_plan(start, length, S, True)
    assert idx.dtype == np.int32 and idx.tolist() == [3, 4, 5, 6, 7] + list(range(8, 16)) + [16, 17] + [30, 31]
    assert pos.tolist() == [0, 1, 2, 3, 4] + list(range(8)) + [0, 1] + [0, 1]
    assert base.tolist() == [0, 5, 13, 15] and last.tolist() == [4, 12, 14, 16]
    _, pos_slot, _, _ = packed_row_plan(start, length, S, False)   # position_ids = arange(S): keep the slot index
    assert pos_slot.tolist() == [3, 4, 5, 6, 7] + list(range(8)) + [0, 1] + [6, 7]
    mask = np.zeros((4, S), dtype=np.int64)
    for b in range(4):
        mask[b, start[b]:start[b] + length[b]] = 1
    assert (np.flatnonzero(mask.reshape(-1)) == idx).all()            # exactly the valid positions, in order
    assert ((np.cumsum(mask, 1) - 1)[mask == 1] == pos).all()         # position_ids = cumsum(mask) - 1 on them
    with pytest.raises(ValueError):
        packed_row_plan([5], [4], S, True)
