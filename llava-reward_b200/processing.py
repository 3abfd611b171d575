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
from .config import LlavaNextRewardConfig, RewardConfig, anyres_geometry, select_best_resolution

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


def _pil_bicubic_weight(x: np.ndarray) -> np.ndarray:
    """Pillow bicubic_filter (a = -0.5) on |x| (Resample.c)."""
    a = -0.5
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@functools.lru_cache(maxsize=256)
def _bicubic_taps(in_size: int, out_size: int):
    """Pillow's BICUBIC resample taps (22-bit fixed point, negative lobes rounded away from zero like
    normalize_coeffs_8bpc) for a full-image box: bounds [out,2], coeffs [out,ksize]."""
    scale = in_size / out_size
    filterscale = scale if scale > 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    centers = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((centers - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum((centers + support + 0.5).astype(np.int64), in_size)
    n = (xmax - xmin).astype(np.int64)
    t = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((t + xmin[:, None] - centers[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(t < n[:, None], _pil_bicubic_weight(arg), 0.0)
    ww = np.zeros(out_size, dtype=np.float64)
    for j in range(ksize):          # sequential accumulation, as the C loop does
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    v = w * float(1 << _PRECISION_BITS)
    kk = np.where(w < 0, (-0.5 + v), (0.5 + v)).astype(np.int64).astype(np.int32)  # C (int) cast truncates toward zero
    bounds = np.stack([xmin, n], axis=1).astype(np.int32)
    return bounds, kk, ksize


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


# --------------------------------------------------------------------------------------------------
# LLaVA-v1.6 (LlavaNext) processor: reference llava branch, reward_dataset.py:334-346 via AutoProcessor
# --------------------------------------------------------------------------------------------------
def _patch_output_size(hw, target_hw):
    """transformers get_patch_output_size (image_processing_utils.py:671-688)."""
    oh, ow = hw
    th, tw = target_hw
    sw, sh = tw / ow, th / oh
    if sw < sh:
        return min(math.ceil(oh * sw), th), tw
    return th, min(math.ceil(ow * sh), tw)


class LlavaNextImageProcessorB200:
    """GPU counterpart of transformers' LlavaNextImageProcessor (PIL/numpy backend) with the llava-v1.6-vicuna
    preprocessor_config (shortest_edge 336, crop 336, BICUBIC, CLIP mean/std, anyres pinpoints): the uint8 image goes
    to the device once; Pillow-exact bicubic resampling (`lr_resample_u8` with host-computed 22-bit taps), zero
    padding, patch split and normalisation (`lr_patch_pack_f32`) run there. Output is bit-identical to the PIL path."""
    model_input_names = ["pixel_values", "image_sizes"]

    def __init__(self, image_grid_pinpoints=None, image_mean=None, image_std=None, size: int = 336, device="cuda",
                 **kwargs):
        self.image_grid_pinpoints = [list(p) for p in (image_grid_pinpoints or LlavaNextRewardConfig().image_grid_pinpoints)]
        self.image_mean = tuple(image_mean) if image_mean is not None else OPENAI_CLIP_MEAN
        self.image_std = tuple(image_std) if image_std is not None else OPENAI_CLIP_STD
        self.size = int(size)
        if self.size != 336:
            raise ValueError("the CLIP tower of this build is ViT-L/14-336: size must be 336")
        self.device = torch.device(device)
        self._taps = {}
        # uint8 -> float table with numpy's arithmetic of transformers rescale() + normalize()
        x = (np.arange(256, dtype=np.uint8).astype(np.float64) * (1 / 255)).astype(np.float32)
        mean = np.array(self.image_mean, dtype=np.float32)[:, None]
        std = np.array(self.image_std, dtype=np.float32)[:, None]
        self._lut = np.ascontiguousarray(((x[None, :] - mean) / std).astype(np.float32))

    def _dev_taps(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        if key not in self._taps:
            b, k, ksize = _bicubic_taps(in_size, out_size)
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

    def _resize(self, x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
        """PIL Image.resize order: horizontal pass, then vertical; unchanged axes are skipped."""
        if x.shape[1] != out_w:
            x = self._resample(x, out_w, 1)
        if x.shape[0] != out_h:
            x = self._resample(x, out_h, 0)
        return x

    def num_patches(self, hw) -> int:
        return anyres_geometry(hw, self.image_grid_pinpoints, self.size)["n_patches"]

    def preprocess(self, images, return_tensors=None, out=None, **kwargs):
        if self.device.type != "cuda":
            raise RuntimeError("LlavaNextImageProcessorB200 runs on CUDA only (no CPU fallback)")
        images = list(images) if isinstance(images, (list, tuple)) else [images]
        arrs = [im if torch.is_tensor(im) else _to_hwc_u8(im) for im in images]
        sizes = [(int(a.shape[0]), int(a.shape[1])) for a in arrs]
        n_pat = [self.num_patches(hw) for hw in sizes]
        P = max(n_pat)
        side = self.size
        if out is None:
            out = torch.empty(len(arrs), P, 3, side, side, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (len(arrs), P, 3, side, side) or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError(f"out must be a CUDA float32 tensor of shape {(len(arrs), P, 3, side, side)}")
        import ctypes
        lut = self._lut.ctypes.data_as(ctypes.c_void_p)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            for i, a in enumerate(arrs):
                if torch.is_tensor(a):
                    if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[2] != 3:
                        raise ValueError("image tensors must be uint8 HxWx3")
                    x = a.to(self.device, non_blocking=True).contiguous()
                else:
                    x = torch.from_numpy(a).to(self.device, non_blocking=True)
                h, w = sizes[i]
                bh, bw = select_best_resolution((h, w), self.image_grid_pinpoints)
                nh, nw = _patch_output_size((h, w), (bh, bw))
                base = self._resize(x, side, side)
                L.call("lr_patch_pack_f32", base.data_ptr(), side, side, 0, 0, 1, 1, lut, out[i, 0].data_ptr(), stream)
                grid = self._resize(x, nh, nw)
                L.call("lr_patch_pack_f32", grid.data_ptr(), nh, nw, (bh - nh) // 2, (bw - nw) // 2, bh // side,
                       bw // side, lut, out[i, 1].data_ptr(), stream)
                if n_pat[i] < P:
                    out[i, n_pat[i]:].zero_()   # _pad_for_batching: zero patches up to the batch maximum
        data = {"pixel_values": out, "image_sizes": [list(s) for s in sizes]}
        if return_tensors == "pt":
            data["image_sizes"] = torch.tensor(data["image_sizes"], dtype=torch.int64)
        return data

    __call__ = preprocess


class LlavaNextProcessorB200:
    """`processor(images=..., text=..., padding=True, return_tensors="pt")` of transformers' LlavaNextProcessor
    (processing_llava_next.py) as the reference's collate_fn calls it (reward_dataset.py:334-346): every `<image>` in
    a prompt is expanded to the image's token count (base 576 + unpadded grid + one newline per grid row; the CLS
    token is dropped by the 'default' feature-select strategy), then the tokenizer pads the batch. Returns a
    transformers BatchFeature (so `.to(device)` works like the reference's caller expects)."""
    image_token = "<image>"

    def __init__(self, image_processor: LlavaNextImageProcessorB200, tokenizer):
        self.image_processor, self.tokenizer = image_processor, tokenizer

    def apply_chat_template(self, conversation, tokenize=False, add_generation_prompt=True, **kw):
        if hasattr(self.tokenizer, "apply_chat_template") and getattr(self.tokenizer, "chat_template", None):
            return self.tokenizer.apply_chat_template(conversation, tokenize=tokenize,
                                                      add_generation_prompt=add_generation_prompt, **kw)
        # Fallback when the checkpoint directory ships no chat template (restated from the llava-v1.6-vicuna
        # chat_template.json from memory, unverified offline): "USER: <image>\n<text> ASSISTANT:", images first.
        out = ""
        for msg in conversation:
            out += msg["role"].upper() + ": "
            out += "".join("<image>\n" for c in msg["content"] if c["type"] == "image")
            out += "".join(c["text"] + " " for c in msg["content"] if c["type"] == "text")
        return out + ("ASSISTANT:" if add_generation_prompt else "")

    def num_image_tokens(self, hw) -> int:
        return anyres_geometry(hw, self.image_processor.image_grid_pinpoints, self.image_processor.size)["n_tokens"]

    def _image_inputs(self, images):
        return self.image_processor(images, return_tensors="pt")

    def __call__(self, images=None, text=None, padding=False, truncation=None, max_length=None, return_tensors="pt",
                 **kwargs):
        from transformers import BatchFeature
        if text is None:
            raise ValueError("You have to specify at least `text`.")
        texts = [text] if isinstance(text, str) else list(text)
        data = {}
        if images is not None:
            images = list(images) if isinstance(images, (list, tuple)) else [images]
            image_inputs = self._image_inputs(images)
            sizes = image_inputs["image_sizes"].tolist()
            it = iter(sizes)
            expanded = []
            for t in texts:
                parts = t.split(self.image_token)
                s = parts[0]
                for tail in parts[1:]:
                    try:
                        hw = next(it)
                    except StopIteration:
                        raise ValueError("more <image> placeholders than images") from None
                    s += self.image_token * self.num_image_tokens(hw) + tail
                expanded.append(s)
            if next(it, None) is not None:
                raise ValueError("fewer <image> placeholders than images")
            texts = expanded
            data.update(image_inputs)
        tok = self.tokenizer(texts, padding=padding, truncation=truncation, max_length=max_length,
                             return_tensors=return_tensors)
        data.update({"input_ids": tok["input_ids"], "attention_mask": tok["attention_mask"]})
        return BatchFeature(data=data)


def load_processor_llava(pretrain_dir: str, cfg: LlavaNextRewardConfig, cache_dir=None, use_fast=True, device="cuda"):
    """(processor, tokenizer) like reference get_tokenizer_llava (llava_reward/utils/utils.py), from a LOCAL
    llava-v1.6-vicuna checkpoint directory: tokenizer via transformers (left padding), image half on the GPU."""
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(pretrain_dir, use_fast=use_fast, cache_dir=cache_dir, padding_side="left")
    if tokenizer.pad_token is None:   # reference utils.py:50-52
        tokenizer.pad_token = tokenizer.eos_token
        tokenizer.pad_token_id = tokenizer.eos_token_id
    if not getattr(tokenizer, "chat_template", None):
        import json
        import os
        for name in ("chat_template.jinja", "chat_template.json"):
            path = os.path.join(pretrain_dir, name)
            if os.path.exists(path):
                with open(path) as f:
                    raw = f.read()
                tokenizer.chat_template = json.loads(raw)["chat_template"] if name.endswith(".json") else raw
                break
    proc = LlavaNextProcessorB200(LlavaNextImageProcessorB200(cfg.image_grid_pinpoints, device=device), tokenizer)
    return proc, tokenizer


# --------------------------------------------------------------------------------------
# Qwen2.5-VL: reference get_tokenizer_qwen (llava_reward/utils/utils.py:34-44) -> AutoProcessor with
# min_pixels = 256*28*28, max_pixels = 1280*28*28; image side of transformers' Qwen2VLImageProcessor
# --------------------------------------------------------------------------------------
def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 256 * 28 * 28,
                 max_pixels: int = 1280 * 28 * 28) -> Tuple[int, int]:
    """transformers image_processing_pil_qwen2_vl.smart_resize (:56-84): both sides multiples of `factor`, pixel count
    inside [min_pixels, max_pixels], aspect ratio kept as closely as possible."""
    import math
    if max(height, width) / min(height, width) > 200:
        raise ValueError(f"absolute aspect ratio must be smaller than 200, got {max(height, width) / min(height, width)}")
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


class Qwen2VLImageProcessorB200:
    """GPU counterpart of transformers' Qwen2VLImageProcessor (PIL backend) as the reference configures it
    (BICUBIC, CLIP mean/std, patch 14, temporal 2, merge 2, min/max_pixels of get_tokenizer_qwen): the uint8 image goes
    to the device once; Pillow-exact bicubic resampling (`lr_resample_u8`), rescale + normalise + patch flattening
    (`lr_qwen_patchify_f32`) run there. Output is bit-identical to the PIL path (tests/golden/qwen_preprocess.pt)."""
    model_input_names = ["pixel_values", "image_grid_thw"]

    def __init__(self, min_pixels: int = 256 * 28 * 28, max_pixels: int = 1280 * 28 * 28, patch_size: int = 14,
                 merge_size: int = 2, temporal_patch_size: int = 2, image_mean=None, image_std=None, device="cuda",
                 **kwargs):
        if temporal_patch_size != 2:
            raise ValueError("temporal_patch_size must be 2 (Qwen2-VL / Qwen2.5-VL)")
        self.min_pixels, self.max_pixels = int(min_pixels), int(max_pixels)
        self.patch_size, self.merge_size, self.temporal_patch_size = int(patch_size), int(merge_size), 2
        self.image_mean = tuple(image_mean) if image_mean is not None else OPENAI_CLIP_MEAN
        self.image_std = tuple(image_std) if image_std is not None else OPENAI_CLIP_STD
        self.device = torch.device(device)
        self._taps = {}
        x = (np.arange(256, dtype=np.uint8).astype(np.float64) * (1 / 255)).astype(np.float32)
        mean = np.array(self.image_mean, dtype=np.float32)[:, None]
        std = np.array(self.image_std, dtype=np.float32)[:, None]
        self._lut = np.ascontiguousarray(((x[None, :] - mean) / std).astype(np.float32))

    _dev_taps = LlavaNextImageProcessorB200._dev_taps
    _resample = LlavaNextImageProcessorB200._resample
    _resize = LlavaNextImageProcessorB200._resize

    def grid(self, hw) -> Tuple[int, int]:
        rh, rw = smart_resize(int(hw[0]), int(hw[1]), self.patch_size * self.merge_size, self.min_pixels, self.max_pixels)
        return rh // self.patch_size, rw // self.patch_size

    def get_number_of_image_patches(self, height: int, width: int, images_kwargs=None) -> int:
        gh, gw = self.grid((height, width))
        return gh * gw

    def preprocess(self, images, return_tensors=None, out=None, **kwargs):
        if self.device.type != "cuda":
            raise RuntimeError("Qwen2VLImageProcessorB200 runs on CUDA only (no CPU fallback)")
        images = list(images) if isinstance(images, (list, tuple)) else [images]
        arrs = [im if torch.is_tensor(im) else _to_hwc_u8(im) for im in images]
        grids = [self.grid(a.shape[:2]) for a in arrs]
        K = 3 * 2 * self.patch_size * self.patch_size
        T = sum(gh * gw for gh, gw in grids)
        if out is None:
            out = torch.empty(T, K, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (T, K) or out.dtype != torch.float32 or not out.is_cuda:
            raise ValueError(f"out must be a CUDA float32 tensor of shape {(T, K)}")
        import ctypes
        lut = self._lut.ctypes.data_as(ctypes.c_void_p)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            row = 0
            for a, (gh, gw) in zip(arrs, grids):
                if torch.is_tensor(a):
                    if a.dtype != torch.uint8 or a.dim() != 3 or a.shape[2] != 3:
                        raise ValueError("image tensors must be uint8 HxWx3")
                    x = a.to(self.device, non_blocking=True).contiguous()
                else:
                    x = torch.from_numpy(a).to(self.device, non_blocking=True)
                x = self._resize(x, gh * self.patch_size, gw * self.patch_size)
                L.call("lr_qwen_patchify_f32", x.data_ptr(), gh * self.patch_size, gw * self.patch_size, self.patch_size,
                       self.merge_size, lut, out[row].data_ptr(), stream)
                row += gh * gw
        thw = [[1, gh, gw] for gh, gw in grids]
        data = {"pixel_values": out, "image_grid_thw": thw}
        if return_tensors == "pt":
            data["image_grid_thw"] = torch.tensor(thw, dtype=torch.int64)
        return data

    __call__ = preprocess


class Qwen2_5_VLProcessorB200:
    """`processor(text=..., images=..., padding=True, return_tensors="pt")` of transformers' Qwen2_5_VLProcessor
    (processing_qwen2_5_vl.py) as the reference's collate_fn calls it (reward_dataset.py:472-487): every
    `<|image_pad|>` in a prompt is expanded to the image's merged-token count (grid_h * grid_w / merge^2), then the
    tokenizer pads the batch. Image preprocessing runs on the GPU. Returns a transformers BatchFeature."""
    image_token = "<|image_pad|>"

    def __init__(self, image_processor: Qwen2VLImageProcessorB200, tokenizer):
        self.image_processor, self.tokenizer = image_processor, tokenizer

    def apply_chat_template(self, conversation, tokenize=False, add_generation_prompt=True, **kw):
        if hasattr(self.tokenizer, "apply_chat_template") and getattr(self.tokenizer, "chat_template", None):
            return self.tokenizer.apply_chat_template(conversation, tokenize=tokenize,
                                                      add_generation_prompt=add_generation_prompt, **kw)
        # Fallback when the checkpoint directory ships no chat template: the Qwen2.5-VL-Instruct template restated
        # (default system prompt; an image is <|vision_start|><|image_pad|><|vision_end|>). The reference slices this
        # string with [58:-23] (reward_dataset.py:417), which removes exactly the system turn and the generation prompt.
        out = "<|im_start|>system\nYou are a helpful assistant.<|im_end|>\n"
        for msg in conversation:
            out += f"<|im_start|>{msg['role']}\n"
            content = msg["content"] if isinstance(msg["content"], list) else [{"type": "text", "text": msg["content"]}]
            for c in content:
                out += "<|vision_start|><|image_pad|><|vision_end|>" if c["type"] == "image" else c.get("text", "")
            out += "<|im_end|>\n"
        return out + ("<|im_start|>assistant\n" if add_generation_prompt else "")

    def __call__(self, text=None, images=None, videos=None, padding=False, truncation=None, max_length=None,
                 return_tensors="pt", **kwargs):
        from transformers import BatchFeature
        if text is None:
            raise ValueError("You have to specify at least `text`.")
        if videos:
            raise NotImplementedError("video inputs: the reference's reward datasets are image-only")
        texts = [text] if isinstance(text, str) else list(text)
        data = {}
        if images is not None:
            images = list(images) if isinstance(images, (list, tuple)) else [images]
            image_inputs = self.image_processor(images, return_tensors="pt")
            m2 = self.image_processor.merge_size ** 2
            it = iter(image_inputs["image_grid_thw"].tolist())
            expanded = []
            for t in texts:
                parts = t.split(self.image_token)
                s = parts[0]
                for tail in parts[1:]:
                    try:
                        g = next(it)
                    except StopIteration:
                        raise ValueError("more <|image_pad|> placeholders than images") from None
                    s += self.image_token * (g[0] * g[1] * g[2] // m2) + tail
                expanded.append(s)
            if next(it, None) is not None:
                raise ValueError("fewer <|image_pad|> placeholders than images")
            texts = expanded
            data.update(image_inputs)
        tok = self.tokenizer(texts, padding=padding, truncation=truncation, max_length=max_length,
                             return_tensors=return_tensors)
        data.update({"input_ids": tok["input_ids"], "attention_mask": tok["attention_mask"]})
        return BatchFeature(data=data)


def load_processor_qwen(pretrain_dir: str, cache_dir=None, use_fast=True, device="cuda"):
    """(processor, tokenizer) like reference get_tokenizer_qwen (llava_reward/utils/utils.py:34-44), from a LOCAL
    Qwen2.5-VL checkpoint directory: tokenizer via transformers (left padding), image half on the GPU with the
    reference's pixel budget."""
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(pretrain_dir, use_fast=use_fast, cache_dir=cache_dir, padding_side="left")
    if tokenizer.pad_token is None:   # reference utils.py:40-42
        tokenizer.pad_token = tokenizer.eos_token
        tokenizer.pad_token_id = tokenizer.eos_token_id
    proc = Qwen2_5_VLProcessorB200(Qwen2VLImageProcessorB200(256 * 28 * 28, 1280 * 28 * 28, device=device), tokenizer)
    return proc, tokenizer
