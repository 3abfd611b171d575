"""ORACLE (test infrastructure, not product code): numpy restatement of the image half of the LlavaNext processor the
reference's llava branch uses (`get_tokenizer_llava` -> AutoProcessor, llava_reward/utils/utils.py; called from
llava_reward/datasets/reward_dataset.py:334-346). The arithmetic is third-party `transformers`
(pinned 4.50.0, requirements.txt:9; installed 5.5.0): image_processing_pil_llava_next.py `get_image_patches` /
`_preprocess`, image_processing_utils.py `select_best_resolution` / `get_patch_output_size`, image_transforms.py
`rescale` / `normalize`; the resize is Pillow's 8-bit antialiased BICUBIC `Image.resize`
(Resample.c precompute_coeffs / normalize_coeffs_8bpc, horizontal pass then vertical pass, uint8 after each).

Pinned against `LlavaNextImageProcessorPil` itself by tests/golden/make_llava_preprocess_golden.py
(tests/golden/llava_preprocess.pt)."""
from __future__ import annotations

import math

import numpy as np

from .llava_next_oracle import select_best_resolution
from .preprocess_oracle import CLIP_MEAN, CLIP_STD, resample_axis

PINPOINTS = [[336, 672], [672, 336], [672, 672], [1008, 336], [336, 1008]]


def pil_resize_bicubic(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), BICUBIC) on HxWx3 uint8: horizontal pass first, a pass is skipped when the
    size along it is unchanged (Pillow ImagingResample)."""
    if img.shape[1] != out_w:
        img = resample_axis(img, out_w, 1, "bicubic")
    if img.shape[0] != out_h:
        img = resample_axis(img, out_h, 0, "bicubic")
    return img


def patch_output_size(hw, target_hw):
    """get_patch_output_size (image_processing_utils.py:671-688)"""
    oh, ow = hw
    th, tw = target_hw
    sw, sh = tw / ow, th / oh
    if sw < sh:
        return min(math.ceil(oh * sw), th), tw
    return th, min(math.ceil(ow * sh), tw)


def normalise(u8_chw: np.ndarray) -> np.ndarray:
    """rescale (float64 multiply by 1/255, cast to float32) then (x - mean) / std in float32"""
    x = (u8_chw.astype(np.float64) * (1 / 255)).astype(np.float32)
    mean = np.array(CLIP_MEAN, dtype=np.float32)[:, None, None]
    std = np.array(CLIP_STD, dtype=np.float32)[:, None, None]
    return (x - mean) / std


def preprocess(img_u8: np.ndarray, pinpoints=PINPOINTS, side: int = 336):
    """HxWx3 uint8 -> (patches [n_patches, 3, 336, 336] float32 (base view first), (h, w))."""
    h, w = img_u8.shape[:2]
    bh, bw = select_best_resolution((h, w), pinpoints)
    nh, nw = patch_output_size((h, w), (bh, bw))
    resized = pil_resize_bicubic(img_u8, nh, nw)
    canvas = np.zeros((bh, bw, 3), dtype=np.uint8)
    top, left = (bh - nh) // 2, (bw - nw) // 2
    canvas[top: top + nh, left: left + nw] = resized
    views = [pil_resize_bicubic(img_u8, side, side)]
    for i in range(0, bh, side):
        for j in range(0, bw, side):
            views.append(canvas[i: i + side, j: j + side])
    return np.stack([normalise(v.transpose(2, 0, 1)) for v in views]), (h, w)
