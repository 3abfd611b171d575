"""Golden vectors for the Qwen2.5-VL image preprocessing: runs transformers' own PIL/numpy `Qwen2VLImageProcessorPil`
with the reference's pixel budget (get_tokenizer_qwen, llava_reward/utils/utils.py:35-37) on the deterministic
synthetic images of tests/preprocess_util.py; stores shapes, grids, SHA-1 of the float32 bytes and samples.
Runs only in the build container."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.dirname(os.path.abspath(__file__))
from preprocess_util import QWEN_CASES, synth_image  # noqa: E402


def main():
    from PIL import Image
    from transformers.models.qwen2_vl.image_processing_pil_qwen2_vl import Qwen2VLImageProcessorPil
    proc = Qwen2VLImageProcessorPil(min_pixels=256 * 28 * 28, max_pixels=1280 * 28 * 28)
    fx = {"cases": []}
    for name, (h, w) in QWEN_CASES.items():
        out = proc.preprocess(Image.fromarray(synth_image(name, h, w)), return_tensors="pt")
        pv = out["pixel_values"].contiguous()
        entry = {"name": name, "hw": (h, w), "grid": out["image_grid_thw"][0].tolist(), "shape": list(pv.shape),
                 "sum": pv.double().sum().item(), "sha1": hashlib.sha1(pv.numpy().tobytes()).hexdigest(),
                 "sample": pv.flatten()[::997].clone(), "rows": pv[5:9].clone()}
        print(name, entry["shape"], entry["grid"], entry["sum"], entry["sha1"])
        fx["cases"].append(entry)
    names = ["small_200x333", "square_1024"]
    out = proc.preprocess([Image.fromarray(synth_image(n, *QWEN_CASES[n])) for n in names], return_tensors="pt")
    fx["batch"] = {"names": names, "shape": list(out["pixel_values"].shape), "grid": out["image_grid_thw"].tolist(),
                   "sha1": hashlib.sha1(out["pixel_values"].contiguous().numpy().tobytes()).hexdigest()}
    print("batch", fx["batch"])
    torch.save(fx, os.path.join(OUT, "qwen_preprocess.pt"))


if __name__ == "__main__":
    main()
