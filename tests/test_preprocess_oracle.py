"""The numpy preprocessing oracle vs the reference's own `Phi3VImageProcessor` outputs (fixture made by
tests/golden/make_preprocess_golden.py): crops bit-exact (uint8 resample + IEEE normalise), bicubic global view 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as PO
from preprocess_util import CASES, synth_image

FX = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.pt"), weights_only=False)


@pytest.mark.parametrize("entry", FX["cases"], ids=[e["name"] for e in FX["cases"]])
def test_preprocess_oracle_matches_reference(entry):
    h, w = entry["hw"]
    out, (ph, pw), ntok = PO.preprocess(synth_image(entry["name"], h, w))
    assert [ph, pw] == entry["image_sizes"] and ntok == entry["num_img_tokens"]
    assert list(out.shape) == entry["shape"]
    pv = torch.from_numpy(out)
    assert torch.equal(pv[1, :, 100:104, :], entry["crop1_rows"])          # crops: bit-exact
    assert (pv[0].flatten()[::101] - entry["global_sample"]).abs().max().item() < 1e-5
    assert (pv.flatten()[::997] - entry["sample"]).abs().max().item() < 1e-5
    assert abs(pv.double().sum().item() - entry["sum"]) < 1e-3 * max(1.0, abs(entry["sum"])) * 1e-3 + 0.5
