"""Golden vectors for the LLaVA-v1.6 image preprocessing: runs transformers' own PIL/numpy `LlavaNextImageProcessorPil`
(size shortest_edge 336, crop 336 = the llava-v1.6-vicuna preprocessor_config.json) on the deterministic synthetic
images of tests/preprocess_util.py; stores shapes, checksums, SHA-1 of the float32 bytes and strided samples.
Runs only in the build container."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.dirname(os.path.abspath(__file__))
from preprocess_util import LLAVA_CASES, synth_image  # noqa: E402


def main():
    from PIL import Image
    from transformers.models.llava_next.image_processing_pil_llava_next import LlavaNextImageProcessorPil
    proc = LlavaNextImageProcessorPil(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336})
    fx = {"cases": []}
    for name, (h, w) in LLAVA_CASES.items():
        img = synth_image(name, h, w)
        out = proc.preprocess(Image.fromarray(img), return_tensors="pt")
        pv = out["pixel_values"][0].contiguous()
        entry = {"name": name, "hw": (h, w), "image_sizes": [int(v) for v in out["image_sizes"][0]],
                 "shape": list(pv.shape), "sum": pv.double().sum().item(), "abs_sum": pv.double().abs().sum().item(),
                 "sha1": hashlib.sha1(pv.numpy().tobytes()).hexdigest(),
                 "sample": pv.flatten()[::997].clone(), "base_rows": pv[0, :, 100:104, :].clone(),
                 "patch1_rows": pv[1, :, 100:104, :].clone()}
        print(name, entry["shape"], entry["image_sizes"], entry["sum"], entry["sha1"])
        fx["cases"].append(entry)
    # batch padding to the largest patch count (zeros)
    imgs = [Image.fromarray(synth_image(n, *LLAVA_CASES[n])) for n in ("square_768", "small_200x333")]
    out = proc.preprocess(imgs, return_tensors="pt")
    fx["batch"] = {"names": ["square_768", "small_200x333"], "shape": list(out["pixel_values"].shape),
                   "image_sizes": out["image_sizes"].tolist(),
                   "sha1": hashlib.sha1(out["pixel_values"].contiguous().numpy().tobytes()).hexdigest()}
    print("batch", fx["batch"])
    torch.save(fx, os.path.join(OUT, "llava_preprocess.pt"))


if __name__ == "__main__":
    main()
