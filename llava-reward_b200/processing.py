"""GPU image preprocessing + prompt assembly with the reference's processor interface.

`Phi3VImageProcessorB200.preprocess` mirrors `Phi3VImageProcessor.preprocess`
(reference llava_reward/models/base_mllm/phi3_v/processing_phi3_v.py:208-288): same arguments, same returned keys
(`pixel_values`, `image_sizes`, `num_img_tokens`), but the pixels never exist on the host as fp32: the uint8 image
is copied to the device once and resized / padded / normalised / cropped by `lr_resample_u8` + `lr_hd_pack_f32`.
The resample taps are computed on the host exactly as Pillow does (precompute_coeffs / normalize_coeffs_8bpc for the
triangle filter), so the uint8 result is bit-identical to torchvision-on-PIL.
"""
from __future__ import annotations

import functools
import math
import re
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from .config import RewardConfig

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_PRECISION_BITS = 22


def calc_hd_transform_size(width: int, height: int, hd_num: int = 16) -> Tuple[int, int]:
    """Padded HD (width, height) for an input size (reference processing_phi3_v.py:106-126)."""
    trans, new_w, new_h, tar = _hd_geometry(width, height, hd_num)
    return (tar, new_w) if trans else (new_w, tar)


def _hd_geometry(width: int, height: int, hd_num: int):
    trans = width < height
    if trans:
        width, height = height, width
    ratio = width / height
    scale = 1
    while scale * math.ceil(scale / ratio) <= hd_num:
        scale += 1
    scale -= 1
    new_w = int(scale * 336)
    new_h = int(new_w / ratio)
    tar = int(math.ceil(new_h / 336) * 336)
    return trans, new_w, new_h, tar


@functools.lru_cache(maxsize=256)
def _triangle_taps(in_size: int, out_size: int):
    """Pillow's bilinear resample taps (22-bit fixed point) for a full-image box: bounds [out,2], coeffs [out,ksize]."""
    scale = in_size / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    centers = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((centers - support + 0.5).astype(np.int64), 0)          # C (int) cast truncates toward zero
    xmax = np.minimum((centers + support + 0.5).astype(np.int64), in_size)
    n = (xmax - xmin).astype(np.int64)
    t = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((t + xmin[:, None] - centers[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where((arg < 1.0) & (t < n[:, None]), 1.0 - arg, 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):          # sequential accumulation, as the C loop does
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = (0.5 + w * float(1 << _PRECISION_BITS)).astype(np.int64).astype(np.int32)  # weights are >= 0 for this filter
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return bounds, kk, ksize


def _to_hwc_u8(image) -> np.ndarray:
    if isinstance(image, np.ndarray):
        arr = image
    elif torch.is_tensor(image):
        arr = image.cpu().numpy()
    else:  # PIL
        arr = np.asarray(image.convert("RGB"))
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError("images must be RGB uint8 HxWx3 (PIL.Image, numpy.ndarray or torch.Tensor)")
    return np.ascontiguousarray(arr)


class Phi3VImageProcessorB200:
    model_input_names = ["pixel_values"]

    def __init__(self, num_crops: int = 1, image_mean=None, image_std=None, do_convert_rgb: bool = True,
                 device="cuda", **kwargs):
        self.num_crops = num_crops
        self.image_mean = tuple(image_mean) if image_mean is not None else OPENAI_CLIP_MEAN
        self.image_std = tuple(image_std) if image_std is not None else OPENAI_CLIP_STD
        self.do_convert_rgb = do_convert_rgb
        self.device = torch.device(device)
        self._taps = {}

    # -- host arithmetic (reference :157-206)
    def calc_num_image_tokens_from_image_size(self, width: int, height: int) -> int:
        w, h = calc_hd_transform_size(width, height, hd_num=self.num_crops)
        return int((h // 336 * w // 336 + 1) * 144 + 1 + (h // 336 + 1) * 12)

    def calc_num_image_tokens(self, images) -> List[int]:
        images = images if isinstance(images, (list, tuple)) else [images]
        return [self.calc_num_image_tokens_from_image_size(*(_to_hwc_u8(im).shape[1::-1])) for im in images]

    def _dev_taps(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        if key not in self._taps:
            b, k, ksize = _triangle_taps(in_size, out_size)
            self._taps[key] = (torch.from_numpy(b).to(self.device), torch.from_numpy(k).to(self.device), ksize)
        return self._taps[key]

    def _resample(self, src: torch.Tensor, out_size: int, axis: int) -> torch.Tensor:
        h, w = src.shape[:2]
        dh, dw = (h, out_size) if axis == 1 else (out_size, w)
        dst = torch.empty(dh, dw, 3, dtype=torch.uint8, device=self.device)
        b, k, ksize = self._dev_taps(w if axis == 1 else h, out_size)
        L.call("lr_resample_u8", src.data_ptr(), h, w, dst.data_ptr(), dh, dw, axis, b.data_ptr(), k.data_ptr(), ksize,
               torch.cuda.current_stream().cuda_stream)
        return dst

    def preprocess(self, images, image_mean=None, image_std=None, do_convert_rgb=None, return_tensors=None, out=None):
        if self.device.type != "cuda":
            raise RuntimeError("Phi3VImageProcessorB200 runs on CUDA only (no CPU fallback)")
        import ctypes
        mean = tuple(image_mean) if image_mean is not None else self.image_mean
        std = tuple(image_std) if image_std is not None else self.image_std
        mean_c, std_c = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
        images = list(images) if isinstance(images, (list, tuple)) else [images]
        n_slots = self.num_crops + 1
        out = out if out is not None else torch.empty(len(images), n_slots, 3, 336, 336, dtype=torch.float32,
                                                      device=self.device)
        shapes, ntoks = [], []
        with torch.cuda.device(self.device):
            for i, image in enumerate(images):
                if torch.is_tensor(image):  # uint8 HWC tensor, already on the device or in (pinned) host memory
                    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
                        raise ValueError("image tensors must be uint8 HxWx3")
                    x = image.to(self.device, non_blocking=True).contiguous()
                else:
                    x = torch.from_numpy(_to_hwc_u8(image)).to(self.device, non_blocking=True)
                H0, W0 = int(x.shape[0]), int(x.shape[1])
                trans, new_w, new_h, tar = _hd_geometry(W0, H0, self.num_crops)
                # Pillow resizes the (possibly transposed) image horizontally first, then vertically; in the
                # original orientation a transposed image is therefore resampled along axis 0 first.
                if trans:
                    tgt_h, tgt_w, order = new_w, new_h, (0, 1)
                else:
                    tgt_h, tgt_w, order = new_h, new_w, (1, 0)
                if (H0, W0) != (tgt_h, tgt_w):
                    for axis in order:
                        tgt = tgt_w if axis == 1 else tgt_h
                        if x.shape[axis] != tgt:
                            x = self._resample(x, tgt, axis)
                pad = int((tar - new_h) / 2)
                if trans:
                    Hh, Ww, pad_top, pad_left = new_w, tar, 0, pad
                else:
                    Hh, Ww, pad_top, pad_left = tar, new_w, pad, 0
                L.call("lr_hd_pack_f32", x.data_ptr(), x.shape[0], x.shape[1], pad_top, pad_left, Hh, Ww, mean_c, std_c,
                       out[i].data_ptr(), n_slots, torch.cuda.current_stream().cuda_stream)
                shapes.append([Hh, Ww])
                ntoks.append(int(((Hh // 336) * (Ww // 336) + 1) * 144 + 1 + (Hh // 336 + 1) * 12))
        data = {"pixel_values": out, "image_sizes": shapes, "num_img_tokens": ntoks}
        if return_tensors == "pt":
            data["image_sizes"] = torch.tensor(shapes, dtype=torch.int64)
            data["num_img_tokens"] = torch.tensor(ntoks, dtype=torch.int64)
        return data

    __call__ = preprocess


class Phi3VProcessorB200:
    """Prompt + image assembly with the call signature of the reference's `Phi3VProcessor`
    (processing_phi3_v.py:291-477): text split on `<|image_N|>`, each image slot filled with N_v copies of -N,
    `attention_mask = input_ids > -1000000` (:407-454)."""

    def __init__(self, image_processor: Phi3VImageProcessorB200, tokenizer):
        self.image_processor, self.tokenizer = image_processor, tokenizer
        self.num_img_tokens = 144  # per 336x336 crop after the 2x2 merge (config img_processor.num_img_tokens)

    def _image_inputs(self, images):
        return self.image_processor(images, return_tensors="pt")

    def __call__(self, text, images=None, padding=False, truncation=None, max_length=None, return_tensors="pt"):
        if images is None:
            return self.tokenizer(text, return_tensors=return_tensors, padding=padding, truncation=truncation,
                                  max_length=max_length)
        image_inputs = self._image_inputs(images)
        pattern = r"<\|image_\d+\|>"
        chunks = [self.tokenizer(c).input_ids for c in re.split(pattern, text)]
        tags = re.findall(pattern, text)
        image_ids = [int(s.split("|")[1].split("_")[-1]) for s in tags]
        unique = sorted(set(image_ids))
        if unique != list(range(1, len(unique) + 1)) or len(unique) != len(images if isinstance(images, (list, tuple)) else [images]):
            raise AssertionError("image tags must be 1..n and match the number of images")  # reference :429-432
        ntok = image_inputs["num_img_tokens"].tolist()
        pads = [[-iid] * ntok[iid - 1] for iid in image_ids]
        # chunks and image-token runs alternate; like the reference (offset = 0) nothing is stripped from the chunks
        ids: List[int] = []
        for j, c in enumerate(chunks):
            ids.extend(c)
            if j < len(pads):
                ids.extend(pads[j])
        input_ids = torch.tensor(ids, dtype=torch.long).unsqueeze(0)
        return {"input_ids": input_ids, "attention_mask": (input_ids > -1000000).to(torch.long),
                "pixel_values": image_inputs["pixel_values"], "image_sizes": image_inputs["image_sizes"]}


def load_processor(pretrain_dir: str, cfg: RewardConfig, cache_dir=None, use_fast=True, device="cuda"):
    """(processor, tokenizer) like reference llava_reward/utils/utils.py:19-32, from a LOCAL checkpoint directory
    (no hub access here): tokenizer via transformers, image half on the GPU with num_crops=16."""
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(pretrain_dir, use_fast=use_fast, cache_dir=cache_dir, padding_side="left")
    if tokenizer.pad_token is None:
        tokenizer.pad_token = tokenizer.eos_token
    proc = Phi3VProcessorB200(Phi3VImageProcessorB200(num_crops=cfg.num_crops, device=device), tokenizer)
    return proc, tokenizer
