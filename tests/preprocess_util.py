"""Deterministic synthetic uint8 RGB test images (smooth gradients + hash noise) shared by the golden generator and tests."""
import numpy as np
import zlib

CASES = {"landscape_512x640": (512, 640), "square_768": (768, 768), "portrait_1000x300": (1000, 300),
         "wide_1080x1920": (1080, 1920), "small_200x333": (200, 333), "exact_672x1008": (672, 1008)}


def synth_image(name: str, h: int, w: int) -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    key = zlib.crc32(name.encode())
    noise = ((xx * 73856093) ^ (yy * 19349663) ^ key) & 0xFFFFFFFF
    noise = ((noise * 2654435761) & 0xFFFFFFFF) >> 24
    img = np.empty((h, w, 3), dtype=np.uint8)
    img[..., 0] = ((xx * 255) // max(w - 1, 1) * 3 // 4 + noise // 4) & 0xFF
    img[..., 1] = ((yy * 255) // max(h - 1, 1) * 3 // 4 + (noise * 7 % 256) // 4) & 0xFF
    img[..., 2] = (((xx + yy) * 255) // max(h + w - 2, 1) // 2 + noise // 2) & 0xFF
    return img

# LLaVA-v1.6 anyres cases: every grid pinpoint, up- and down-scaling, odd sizes, exact fits
LLAVA_CASES = {"landscape_512x640": (512, 640), "square_768": (768, 768), "portrait_1000x300": (1000, 300),
               "wide_300x900": (300, 900), "small_200x333": (200, 333), "exact_672x672": (672, 672),
               "tall_1080x700": (1080, 700), "tiny_64x48": (64, 48)}

# Qwen2.5-VL cases: inside the pixel budget, above max_pixels (downscale), below min_pixels (upscale), odd sizes
QWEN_CASES = {"landscape_512x640": (512, 640), "square_1024": (1024, 1024), "portrait_1000x300": (1000, 300),
              "wide_1080x1920": (1080, 1920), "small_200x333": (200, 333), "tiny_64x48": (64, 48),
              "exact_448x448": (448, 448)}
