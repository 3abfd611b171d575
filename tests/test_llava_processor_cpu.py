"""Host logic of the LLaVA-v1.6 processor / loader (no GPU): <image> expansion vs transformers' own
LlavaNextProcessor token arithmetic, bicubic taps vs the oracle's Pillow restatement, checkpoint-name mapping."""
import types

import numpy as np
import pytest
import torch

from llava_reward_b200.config import LlavaNextRewardConfig, anyres_geometry
from llava_reward_b200.processing import (LlavaNextProcessorB200, _bicubic_taps, _patch_output_size)
from oracle import llava_preprocess_oracle as PO
from oracle.preprocess_oracle import resample_coeffs


class FakeTok:
    pad_token_id, padding_side = 0, "left"

    def __call__(self, texts, padding=False, return_tensors="pt", **kw):
        rows = []
        for t in texts:
            ids, i = [1], 0
            while i < len(t):
                if t.startswith("<image>", i):
                    ids.append(32000)
                    i += 7
                else:
                    ids.append(10 + ord(t[i]) % 50)
                    i += 1
            rows.append(ids)
        S = max(len(r) for r in rows)
        ids = torch.tensor([[0] * (S - len(r)) + r for r in rows])
        mask = torch.tensor([[0] * (S - len(r)) + [1] * len(r) for r in rows])
        return {"input_ids": ids, "attention_mask": mask}


class StubImageProc:
    image_grid_pinpoints = LlavaNextRewardConfig().image_grid_pinpoints
    size = 336

    def __call__(self, images, return_tensors="pt"):
        sizes = [im.shape[:2] for im in images]
        return {"pixel_values": torch.zeros(len(images), 5, 3, 2, 2), "image_sizes": torch.tensor(sizes)}


def hf_num_image_tokens(h, w):
    """transformers LlavaNextProcessor._get_number_of_features with the 'default' select strategy"""
    from transformers.models.llava_next.processing_llava_next import LlavaNextProcessor
    p = LlavaNextProcessor.__new__(LlavaNextProcessor)
    p.image_processor = types.SimpleNamespace(image_grid_pinpoints=StubImageProc.image_grid_pinpoints)
    p.patch_size, p.num_additional_image_tokens = 14, 1
    return p._get_number_of_features(h, w, 336, 336) - 1


def test_image_token_count_matches_transformers():
    g = torch.Generator().manual_seed(5)
    for _ in range(300):
        h, w = (int(v) for v in torch.randint(30, 1500, (2,), generator=g))
        assert anyres_geometry((h, w), StubImageProc.image_grid_pinpoints)["n_tokens"] == hf_num_image_tokens(h, w), (h, w)


def test_image_placeholder_expansion_and_padding():
    proc = LlavaNextProcessorB200(StubImageProc(), FakeTok())
    imgs = [np.zeros((480, 640, 3), np.uint8), np.zeros((200, 333, 3), np.uint8)]
    out = proc(images=imgs, text=["USER: <image>\nab ASSISTANT:", "x <image> y"], padding=True, return_tensors="pt")
    n = [anyres_geometry(im.shape[:2], StubImageProc.image_grid_pinpoints)["n_tokens"] for im in imgs]
    assert (out["input_ids"] == 32000).sum(1).tolist() == n
    assert out["attention_mask"].sum(1).tolist() == [1 + 6 + n[0] + 14, 1 + 2 + n[1] + 2]
    assert out["input_ids"].shape == out["attention_mask"].shape and out["pixel_values"].shape[0] == 2
    assert hasattr(out, "to")  # BatchFeature: the reference caller does inputs_batch.to(device)
    with pytest.raises(ValueError):
        proc(images=imgs, text=["<image>", "no placeholder"], padding=True)
    with pytest.raises(ValueError):
        proc(images=imgs[:1], text=["<image> <image>"], padding=True)


@pytest.mark.parametrize("n_in,n_out", [(640, 336), (333, 672), (1000, 1008), (48, 336), (700, 436), (336, 336)])
def test_bicubic_taps_match_oracle(n_in, n_out):
    b, k, ksize = _bicubic_taps(n_in, n_out)
    ob, ok = resample_coeffs(n_in, n_out, "bicubic")
    assert np.array_equal(b, ob) and ok.shape[1] == ksize and np.array_equal(k, ok)


def test_patch_output_size_matches_oracle():
    g = torch.Generator().manual_seed(9)
    for _ in range(200):
        h, w = (int(v) for v in torch.randint(30, 1500, (2,), generator=g))
        for t in StubImageProc.image_grid_pinpoints:
            assert _patch_output_size((h, w), t) == PO.patch_output_size((h, w), t)


def test_llava_checkpoint_name_mapping_and_loader_errors(tmp_path):
    from llava_reward_b200.checkpoint import _canonical_llava_name
    from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor
    assert _canonical_llava_name("model.language_model.layers.3.mlp.up_proj.weight") == \
        "language_model.model.layers.3.mlp.up_proj.weight"
    assert _canonical_llava_name("model.vision_tower.vision_model.pre_layrnorm.bias") == "vision_tower.vision_model.pre_layrnorm.bias"
    assert _canonical_llava_name("model.image_newline") == "image_newline"
    assert _canonical_llava_name("language_model.model.norm.weight") == "language_model.model.norm.weight"
    y = tmp_path / "reward_config.yaml"
    y.write_text("is_general_preference: false\nadd_cross_attention: false\nvalue_head_dim: 1\ngeneral_preference_tau: 0.1\n")
    args = types.SimpleNamespace(pretrain="synthetic:1:13b", pm_path=None, cache_dir=None, ft_projector=False)
    args, model = load_reward_adaptor(args, "llava", str(y))
    assert model.model_type == "llava" and model.config.hidden_size == 5120 and model.config.num_layers == 40
    assert args.value_head_dim == 1 and args.is_general_preference is False
    with pytest.raises(RuntimeError):
        model.custom_forward(inputs_batch={})          # not placed on a GPU yet
    with pytest.raises(RuntimeError):
        model.to("cpu")                                # no CPU fallback
    with pytest.raises(NotImplementedError):
        load_reward_adaptor(args, "internvl", str(y))   # the reference knows phi3v / qwen / llava only
    args.pretrain = str(tmp_path / "missing")
    with pytest.raises(FileNotFoundError):
        load_reward_adaptor(args, "llava", str(y))
