"""Qwen2.5-VL image preprocessing: the numpy oracle must reproduce transformers' own PIL processor bit for bit
(tests/golden/qwen_preprocess.pt, made by tests/golden/make_qwen_preprocess_golden.py)."""
import hashlib
import os

import numpy as np
import torch

from preprocess_util import QWEN_CASES, synth_image
from oracle import qwen_preprocess_oracle as PO

FX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qwen_preprocess.pt")


def test_qwen_preprocess_oracle_bit_exact():
    fx = torch.load(FX, weights_only=False)
    for e in fx["cases"]:
        h, w = e["hw"]
        pv, grid = PO.preprocess(synth_image(e["name"], h, w))
        assert list(pv.shape) == e["shape"] and list(grid) == e["grid"], e["name"]
        assert np.array_equal(pv[5:9], e["rows"].numpy()), e["name"]
        assert hashlib.sha1(pv.tobytes()).hexdigest() == e["sha1"], e["name"]
    b = fx["batch"]
    outs = [PO.preprocess(synth_image(n, *QWEN_CASES[n])) for n in b["names"]]
    pv = np.concatenate([o[0] for o in outs], 0)
    assert list(pv.shape) == b["shape"] and [list(o[1]) for o in outs] == b["grid"]
    assert hashlib.sha1(pv.tobytes()).hexdigest() == b["sha1"]


def test_smart_resize_host_matches_oracle():
    """the product's host-side smart_resize (processing.py) against the oracle's on assorted sizes"""
    from llava_reward_b200.processing import smart_resize
    g = torch.Generator().manual_seed(5)
    for _ in range(500):
        h, w = (int(v) for v in torch.randint(20, 3000, (2,), generator=g))
        if max(h, w) / min(h, w) > 200:
            continue
        assert smart_resize(h, w) == PO.smart_resize(h, w), (h, w)
        rh, rw = smart_resize(h, w)
        assert rh % 28 == 0 and rw % 28 == 0 and rh * rw <= 1280 * 28 * 28
