"""GPU image preprocessing (lr_resample_u8 + lr_hd_pack_f32) vs the reference's own processor outputs
(tests/golden/preprocess.pt) and vs the numpy oracle on the full tensor: crops bit-exact, bicubic global view 1e-5."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import preprocess_oracle as PO  # noqa: E402
from preprocess_util import synth_image  # noqa: E402
from llava_reward_b200.processing import Phi3VImageProcessorB200  # noqa: E402

FX = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.pt"), weights_only=False)


@pytest.mark.parametrize("entry", FX["cases"], ids=[e["name"] for e in FX["cases"]])
def test_gpu_preprocess_matches_reference(entry):
    h, w = entry["hw"]
    img = synth_image(entry["name"], h, w)
    proc = Phi3VImageProcessorB200(num_crops=16)
    out = proc.preprocess([img], return_tensors="pt")
    pv = out["pixel_values"][0].cpu()
    assert out["image_sizes"][0].tolist() == entry["image_sizes"]
    assert int(out["num_img_tokens"][0]) == entry["num_img_tokens"]
    assert list(pv.shape) == entry["shape"]
    assert torch.equal(pv[1, :, 100:104, :], entry["crop1_rows"])
    assert (pv[0].flatten()[::101] - entry["global_sample"]).abs().max().item() < 1e-5
    assert (pv.flatten()[::997] - entry["sample"]).abs().max().item() < 1e-5
    ref, _, _ = PO.preprocess(img)
    ref = torch.from_numpy(ref)
    assert torch.equal(pv[1:], ref[1:])                       # every crop + zero slots: bit-exact
    assert (pv[0] - ref[0]).abs().max().item() < 1e-5


def test_processor_token_counts_match_reference_formula():
    proc = Phi3VImageProcessorB200(num_crops=16)
    for e in FX["cases"]:
        h, w = e["hw"]
        assert proc.calc_num_image_tokens_from_image_size(w, h) == e["num_img_tokens"]
