"""Golden vectors for image preprocessing: runs the reference's own vendored `Phi3VImageProcessor(num_crops=16)`
(reference processing_phi3_v.py:208-288; PIL + torchvision + torch) on deterministic synthetic uint8 images and
stores strided samples + checksums (the full tensors are 23 MB each). Runs only in the build container."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.dirname(os.path.abspath(__file__))
from preprocess_util import CASES, synth_image  # noqa: E402


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    import_reference()
    from PIL import Image
    from llava_reward.models.base_mllm.phi3_v.processing_phi3_v import Phi3VImageProcessor
    proc = Phi3VImageProcessor(num_crops=16)
    fx = {"cases": []}
    for name, (h, w) in CASES.items():
        img = synth_image(name, h, w)
        out = proc.preprocess(Image.fromarray(img), return_tensors="pt")
        pv = out["pixel_values"][0]
        entry = {"name": name, "hw": (h, w), "image_sizes": [int(v) for v in out["image_sizes"][0]],
                 "num_img_tokens": int(out["num_img_tokens"][0]), "shape": list(pv.shape),
                 "sum": pv.double().sum().item(), "abs_sum": pv.double().abs().sum().item(),
                 "sample": pv.flatten()[::997].clone(), "global_sample": pv[0].flatten()[::101].clone(),
                 "crop1_rows": pv[1, :, 100:104, :].clone()}
        print(name, entry["image_sizes"], entry["num_img_tokens"], entry["sum"])
        fx["cases"].append(entry)
    torch.save(fx, os.path.join(OUT, "preprocess.pt"))


if __name__ == "__main__":
    main()
