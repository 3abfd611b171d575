"""ORACLE (test infrastructure, not product code): numpy restatement of the image half of the Qwen2.5-VL processor the
reference's qwen branch uses (`get_tokenizer_qwen` -> AutoProcessor(min_pixels=256*28*28, max_pixels=1280*28*28),
llava_reward/utils/utils.py:34-44; called from llava_reward/datasets/reward_dataset.py:472-487). The arithmetic is
third-party `transformers` (pinned 4.50.0, requirements.txt:9; installed 5.5.0):
models/qwen2_vl/image_processing_pil_qwen2_vl.py `smart_resize` (:56-84) and `_preprocess` (:148-227),
image_transforms.py `rescale` / `normalize`; the resize is Pillow's 8-bit antialiased BICUBIC `Image.resize`.

Pinned against `Qwen2VLImageProcessorPil` itself by tests/golden/make_qwen_preprocess_golden.py
(tests/golden/qwen_preprocess.pt)."""
from __future__ import annotations

import math

import numpy as np

from .llava_preprocess_oracle import normalise, pil_resize_bicubic

MIN_PIXELS, MAX_PIXELS = 256 * 28 * 28, 1280 * 28 * 28


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = MIN_PIXELS, max_pixels: int = MAX_PIXELS):
    if max(height, width) / min(height, width) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def preprocess(img_u8: np.ndarray, patch: int = 14, merge: int = 2, temporal: int = 2,
               min_pixels: int = MIN_PIXELS, max_pixels: int = MAX_PIXELS):
    """HxWx3 uint8 -> (flattened patches [gh*gw, 3*temporal*patch*patch] float32, (1, gh, gw))."""
    h, w = img_u8.shape[:2]
    rh, rw = smart_resize(h, w, patch * merge, min_pixels, max_pixels)
    x = normalise(pil_resize_bicubic(img_u8, rh, rw).transpose(2, 0, 1))          # [3, rh, rw]
    patches = np.repeat(x[None, None], temporal, axis=1)                          # [1, T, 3, rh, rw]
    gh, gw = rh // patch, rw // patch
    patches = patches.reshape(1, 1, temporal, 3, gh // merge, merge, patch, gw // merge, merge, patch)
    patches = patches.transpose(0, 1, 4, 7, 5, 8, 3, 2, 6, 9)
    return np.ascontiguousarray(patches.reshape(gh * gw, 3 * temporal * patch * patch)), (1, gh, gw)
