"""LLaVA-v1.6 image preprocessing: the numpy oracle against transformers' own PIL processor (fixtures from
tests/golden/make_llava_preprocess_golden.py). uint8 resampling and the float32 normalisation are bit-exact (SHA-1)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import llava_preprocess_oracle as PO
from preprocess_util import synth_image

FX = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "llava_preprocess.pt"),
                weights_only=False)


@pytest.mark.parametrize("entry", FX["cases"], ids=[e["name"] for e in FX["cases"]])
def test_llava_preprocess_oracle_matches_transformers(entry):
    h, w = entry["hw"]
    pv, hw = PO.preprocess(synth_image(entry["name"], h, w))
    assert list(pv.shape) == entry["shape"] and list(hw) == entry["image_sizes"]
    t = torch.from_numpy(pv)
    assert torch.equal(t[0, :, 100:104, :], entry["base_rows"])
    assert torch.equal(t[1, :, 100:104, :], entry["patch1_rows"])
    assert torch.equal(t.flatten()[::997], entry["sample"])
    assert hashlib.sha1(np.ascontiguousarray(pv).tobytes()).hexdigest() == entry["sha1"]
