"""Prompt/image-token assembly of `Phi3VProcessorB200.__call__` vs the reference's
`_convert_images_texts_to_inputs` (processing_phi3_v.py:407-454), with a fake tokenizer and a stubbed image half."""
import sys
import types

import pytest
import torch

from llava_reward_b200.processing import Phi3VProcessorB200, calc_hd_transform_size


class FakeTok:
    bos_token_id = 1

    def __call__(self, text, **kw):
        return types.SimpleNamespace(input_ids=[1] + [10 + (ord(c) % 50) for c in text])


class StubProc(Phi3VProcessorB200):
    def __init__(self, ntoks):
        self.tokenizer, self._n = FakeTok(), ntoks

    def _image_inputs(self, images):
        return {"pixel_values": torch.zeros(len(images), 17, 3, 2, 2), "image_sizes": torch.tensor([[336, 336]] * len(images)),
                "num_img_tokens": torch.tensor(self._n)}


def reference_assembly(tok, text, ntoks):
    import re
    pattern = r"<\|image_\d+\|>"
    chunks = [tok(c).input_ids for c in re.split(pattern, text)]
    ids_ = [int(s.split("|")[1].split("_")[-1]) for s in re.findall(pattern, text)]
    pads = [[-i] * ntoks[i - 1] for i in ids_]
    if len(chunks) > len(pads):
        pads.append([])
    out = []
    for a, b in zip(chunks, pads):
        out.extend(a)
        out.extend(b)
    return out


@pytest.mark.parametrize("text,ntoks", [("<|user|>\\n<|image_1|>\\na caption<|end|>", [7]),
                                         ("x<|image_1|>y<|image_2|>z", [3, 5]), ("<|image_1|>", [4])])
def test_prompt_assembly_matches_reference(text, ntoks):
    proc = StubProc(ntoks)
    out = proc(text, images=[object()] * len(ntoks))
    ref = reference_assembly(FakeTok(), text, ntoks)
    assert out["input_ids"].tolist() == [ref]
    assert out["attention_mask"].tolist() == [[1] * len(ref)]
    assert (out["input_ids"] < 0).sum().item() == sum(ntoks)


def test_bad_image_tags_raise():
    with pytest.raises(AssertionError):
        StubProc([3, 3])("a<|image_1|>b<|image_3|>", images=[object(), object()])
    with pytest.raises(AssertionError):
        StubProc([3])("a<|image_1|>b", images=[object(), object()])


def test_hd_size_table():
    # (width, height) -> padded HD (width, height), reference calc_hd_transform_size (processing_phi3_v.py:106-126)
    assert calc_hd_transform_size(640, 512) == (1344, 1344)
    assert calc_hd_transform_size(300, 1000) == (672, 2016)
    assert calc_hd_transform_size(1920, 1080) == (1680, 1008)
    assert calc_hd_transform_size(800, 600) == (1344, 1008)
