"""ORACLE (test infrastructure, not product code): numpy restatement of the reference's image preprocessing
`Phi3VImageProcessor.preprocess` (reference llava_reward/models/base_mllm/phi3_v/processing_phi3_v.py:208-288):
HD_transform (:83-104) -> torchvision resize on a PIL image = Pillow's antialiased bilinear `ImagingResample`
(third-party, Pillow 12.2 src/libImaging/Resample.c: precompute_coeffs / normalize_coeffs_8bpc /
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc, uint8 in, 22-bit fixed-point taps, uint8 rounding
after EACH pass) -> white padding to a multiple of 336 (:62-71) -> ToTensor + Normalize (:252-255) -> bicubic global
view of the NORMALISED image (torch F.interpolate, A=-0.75, align_corners=False, no antialias, :265) -> crop split
(:272) -> zero pad to num_crops+1 slots (:277).

Pinned against the reference itself by tests/golden/make_preprocess_golden.py (fixtures tests/golden/preprocess_*.pt).
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2


def hd_geometry(width: int, height: int, hd_num: int = 16):
    """-> (transposed, new_w, new_h, padded_h) in the (possibly transposed) frame of HD_transform (:83-104)."""
    trans = width < height
    if trans:
        width, height = height, width
    ratio = width / height
    scale = 1
    while scale * np.ceil(scale / ratio) <= hd_num:
        scale += 1
    scale -= 1
    new_w = int(scale * 336)
    new_h = int(new_w / ratio)
    tar = int(np.ceil(new_h / 336) * 336)
    return trans, new_w, new_h, tar


def _triangle(a: float) -> float:
    return 1.0 - a if a < 1.0 else 0.0


def _pil_bicubic(x: float) -> float:
    """Pillow bicubic_filter (Resample.c, a = -0.5), argument already |x|."""
    a = -0.5
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int, filt: str = "bilinear"):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc, full-image box; filt = 'bilinear' (triangle, support 1)
    or 'bicubic' (a = -0.5, support 2)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    fsupport, ffun = (1.0, _triangle) if filt == "bilinear" else (2.0, _pil_bicubic)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        n = xmax - xmin
        w = np.zeros(n, dtype=np.float64)
        for x in range(n):
            w[x] = ffun(abs((x + xmin - center + 0.5) * ss))
        ww = w.sum()  # sequential double accumulation in C; numpy pairwise sum differs only below 1e-16 relative
        ww = 0.0
        for x in range(n):
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, n)
    return bounds, kk


def resample_axis(img: np.ndarray, out_size: int, axis: int, filt: str = "bilinear") -> np.ndarray:
    """One Pillow 8bpc resample pass along `axis` (0 = vertical, 1 = horizontal) of an HxWx3 uint8 image."""
    in_size = img.shape[axis]
    bounds, kk = resample_coeffs(in_size, out_size, filt)
    src = np.moveaxis(img, axis, 0).astype(np.int64)  # [in, other, 3]
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc += np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def _cubic_coeffs(t: np.ndarray, A: float = -0.75):
    t = t.astype(np.float32)
    A = np.float32(A)

    def c1(x):  # |x| <= 1
        return ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + np.float32(1)

    def c2(x):  # 1 < |x| < 2
        return ((A * x - np.float32(5) * A) * x + np.float32(8) * A) * x - np.float32(4) * A

    return [c2(t + np.float32(1)), c1(t), c1(np.float32(1) - t), c2(np.float32(2) - t)]


def bicubic_resize(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """torch upsample_bicubic2d (align_corners=False, no antialias) on a [C,H,W] fp32 array."""
    C, H, W = img.shape

    def src_index(out_size, in_size):
        scale = np.float32(in_size / out_size)
        real = scale * (np.arange(out_size, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5)
        i0 = np.floor(real).astype(np.int64)
        return i0, (real - i0.astype(np.float32)).astype(np.float32)

    iy, ty = src_index(out_h, H)
    ix, tx = src_index(out_w, W)
    wy, wx = _cubic_coeffs(ty), _cubic_coeffs(tx)
    out = np.zeros((C, out_h, out_w), dtype=np.float32)
    for i in range(4):
        yy = np.clip(iy - 1 + i, 0, H - 1)
        row = np.zeros((C, out_h, out_w), dtype=np.float32)
        for j in range(4):
            xx = np.clip(ix - 1 + j, 0, W - 1)
            row += img[:, yy][:, :, xx] * wx[j][None, None, :]
        out += row * wy[i][None, :, None]
    return out


def preprocess(img_u8: np.ndarray, num_crops: int = 16):
    """img_u8 [H,W,3] uint8 RGB -> (pixel_values [num_crops+1,3,336,336] fp32, (h, w) padded HD size, num_img_tokens)."""
    H0, W0 = img_u8.shape[:2]
    trans, new_w, new_h, tar = hd_geometry(W0, H0, num_crops)
    x = img_u8.transpose(1, 0, 2) if trans else img_u8
    if (x.shape[1], x.shape[0]) != (new_w, new_h):
        if x.shape[1] != new_w:
            x = resample_axis(x, new_w, axis=1)      # Pillow: horizontal pass first ...
        if x.shape[0] != new_h:
            x = resample_axis(x, new_h, axis=0)      # ... then vertical, on the uint8 intermediate
    top = int((tar - new_h) / 2)
    padded = np.full((tar, new_w, 3), 255, dtype=np.uint8)
    padded[top:top + new_h] = x
    if trans:
        padded = padded.transpose(1, 0, 2)
    h, w = padded.shape[:2]
    t = padded.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    mean = np.asarray(CLIP_MEAN, dtype=np.float32)[:, None, None]
    std = np.asarray(CLIP_STD, dtype=np.float32)[:, None, None]
    hd = (t - mean) / std
    glb = bicubic_resize(hd, 336, 336)
    crops = hd.reshape(3, h // 336, 336, w // 336, 336).transpose(1, 3, 0, 2, 4).reshape(-1, 3, 336, 336)
    out = np.zeros((num_crops + 1, 3, 336, 336), dtype=np.float32)
    out[0] = glb
    out[1:1 + crops.shape[0]] = crops
    ntok = ((h // 336) * (w // 336) + 1) * 144 + 1 + (h // 336 + 1) * 12
    return out, (h, w), ntok
