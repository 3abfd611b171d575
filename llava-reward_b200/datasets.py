"""Manifest reader + dataset for the batch-eval callers: the host-side mirror of the reference's
`GeneralRewardDataset` (llava_reward/datasets/reward_dataset.py:25-202) for the Phi-3.5-V path - same constructor
arguments, same `__getitem__` tuples (pairwise: 10 fields, cls: 5 fields), same `collate_fn` dictionaries - and a
batched feed that turns a manifest into what `batch_eval.score_pairs / score_single` consume.

What differs from the reference is WHERE the pixels are made: `processor` is this package's `Phi3VProcessorB200`, whose
image half runs on the GPU (lr_resample_u8 + lr_hd_pack_f32, bit-exact crops), so an item carries device tensors and
the 1.47 GB of fp32 pixels per 64 samples never cross PCIe; `manifest_batches` additionally decodes the JPEGs of the
next micro-batches on worker threads while the current one is being scored. No arithmetic of the scoring path
happens here.
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from .batch_eval import zero_pad_sequences


def load_manifest(path: str, max_samples: Optional[int] = None) -> List[dict]:
    """The json manifests the reference's eval scripts read through `blending_datasets` (llava_reward/utils/utils.py:
    104-192 -> datasets.load_dataset('json')): a list of {prompt, chosen_path, reject_path, c_rate, r_rate} (pairwise,
    data/sample_test/pairwise_sample.json) or {path, label, prompt} (single image, non_pairwise_sample.json); .jsonl
    with one object per line is accepted as well. `max_samples` = `dataset.select(range(min(max_samples, len)))`
    (eval/batch_inference_rm_phi.py:44)."""
    with open(path) as f:
        text = f.read()
    try:
        rows = json.loads(text)
    except json.JSONDecodeError:
        rows = [json.loads(line) for line in text.splitlines() if line.strip()]
    if isinstance(rows, dict):
        rows = rows.get("data", rows.get("train", [rows]))
    if not isinstance(rows, list) or not rows or not isinstance(rows[0], dict):
        raise ValueError(f"{path}: expected a JSON list of objects")
    return rows[: max_samples] if max_samples is not None else rows


def is_non_pairwise(rows: Sequence[dict]) -> bool:
    """the reference's test `len(dataset[0]) == 3` (eval/batch_inference_rm_phi.py:45-48)"""
    return len(rows[0]) == 3


def decode_rgb(path: str) -> np.ndarray:
    """`Image.open(path).convert("RGB")` (reward_dataset.py:79-80) as a uint8 HxWx3 array; truncated files load like in
    the reference (ImageFile.LOAD_TRUNCATED_IMAGES = True, :10)."""
    from PIL import Image, ImageFile
    ImageFile.LOAD_TRUNCATED_IMAGES = True
    with Image.open(path) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB")))


class GeneralRewardDataset(Dataset):
    """Reference llava_reward/datasets/reward_dataset.py:25-202 for the Phi-3.5-V processor.

    `image_root`: directory relative paths of the manifest are resolved against (the reference resolves them against
    the process's working directory, i.e. the repo root)."""

    def __init__(self, dataset, processor: Callable, tokenizer: Callable, strategy=None, is_custom=False,
                 return_prompt_length=False, cls_based=False, image_root: Optional[str] = None) -> None:
        super().__init__()
        self.tokenizer, self.processor, self.strategy = tokenizer, processor, strategy
        self.is_custom, self.return_prompt_length, self.cls_based = is_custom, return_prompt_length, cls_based
        self.image_root = image_root
        self.prompts: List = []
        if cls_based:
            self.path_list, self.label_list = [], []
            for data in dataset:
                self.prompts.append(data["prompt"])
                self.path_list.append(data["path"])
                self.label_list.append(data["label"])
        else:
            self.chosens, self.rejects, self.c_rates, self.r_rates = [], [], [], []
            for data in dataset:
                self.prompts.append(data["prompt"])
                self.chosens.append(data["chosen_path"])
                self.rejects.append(data["reject_path"])
                self.c_rates.append(data["c_rate"])
                self.r_rates.append(data["r_rate"])

    def __len__(self):
        return len(self.prompts)

    # ---- pieces of __getitem__ (reused by the threaded feed) ----------------------------------------------------
    def resolve(self, path: str) -> str:
        return path if os.path.isabs(path) or self.image_root is None else os.path.join(self.image_root, path)

    def prompt_text(self, prompt: str) -> str:
        """reward_dataset.py:82-88: chat template of one user turn '<|image_1|>\\n{prompt}', generation prompt cut off
        (the last 22 characters = '<|end|>\\n<|assistant|>\\n'), eos appended"""
        msg = {"role": "user", "content": f"<|image_1|>\n{prompt}"}
        text = self.tokenizer.apply_chat_template([msg], tokenize=False, add_generation_prompt=True)[:-22]
        return text + self.tokenizer.eos_token

    def prompts_of(self, idx: int):
        p = self.prompts[idx]
        return (p[0], p[1]) if isinstance(p, list) else (p, p)     # per-image prompts or one shared prompt (:81-104)

    def encode(self, prompt: str, image) -> Dict[str, torch.Tensor]:
        return self.processor(self.prompt_text(prompt), [image], return_tensors="pt")

    def __getitem__(self, idx):
        if not self.cls_based:
            pc, pr = self.prompts_of(idx)
            c = self.encode(pc, decode_rgb(self.resolve(self.chosens[idx])))
            r = self.encode(pr, decode_rgb(self.resolve(self.rejects[idx])))
            return (c["input_ids"], c["attention_mask"], c["pixel_values"], c["image_sizes"],
                    r["input_ids"], r["attention_mask"], r["pixel_values"], r["image_sizes"],
                    self.c_rates[idx], self.r_rates[idx])
        t = self.encode(self.prompts[idx], decode_rgb(self.resolve(self.path_list[idx])))
        return (t["input_ids"], t["attention_mask"], t["pixel_values"], t["image_sizes"], self.label_list[idx])

    def _batch(self, ids, masks, pixels, sizes) -> Dict[str, torch.Tensor]:
        return {"input_ids": zero_pad_sequences(ids, value=self.tokenizer.pad_token_id),
                "attention_mask": zero_pad_sequences(masks),
                "pixel_values": torch.stack(pixels, dim=0), "image_sizes": torch.stack(sizes, dim=0)}

    def collate_fn(self, item_list):
        """reward_dataset.py:137-202: left padding with pad_token_id / 0, pixels and sizes stacked; tensors keep the
        [B, 1, ...] shape the eval loop squeezes (eval/batch_inference_rm_phi.py:82-90)."""
        if not self.cls_based:
            cols = list(zip(*item_list))
            return (self._batch(cols[0], cols[1], cols[2], cols[3]), self._batch(cols[4], cols[5], cols[6], cols[7]),
                    list(cols[8]), list(cols[9]))
        cols = list(zip(*item_list))
        return self._batch(cols[0], cols[1], cols[2], cols[3]), torch.tensor(list(cols[4]), dtype=torch.long)


def squeeze_batch(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """the `.squeeze(1)` of the reference's eval loop (eval/batch_inference_rm_phi.py:82-90)"""
    return {k: v.squeeze(1) for k, v in batch.items()}


def manifest_batches(dataset: GeneralRewardDataset, micro_batch_size: int, decode_threads: int = 4,
                     lookahead: int = 2) -> Iterator:
    """Micro-batches of `dataset` in order (DistributedSampler(shuffle=False, num_replicas=1) + DataLoader(drop_last=
    False) of eval/batch_inference_rm_phi.py:50-66), already squeezed, in the form the loops of `batch_eval` take:
    pairwise -> (batch_chosen, batch_rejected); cls -> (batch, labels).

    JPEG decoding (the only CPU-heavy step left: PIL, releases the GIL) of the next `lookahead` micro-batches runs on
    `decode_threads` worker threads while the GPU scores the current one; tokenisation and the GPU preprocessing
    launches stay on the calling thread (one CUDA stream, no cross-thread ordering to reason about)."""
    n = len(dataset)
    starts = list(range(0, n, micro_batch_size))

    def paths_of(i):
        return (dataset.path_list[i],) if dataset.cls_based else (dataset.chosens[i], dataset.rejects[i])

    with ThreadPoolExecutor(max_workers=max(1, decode_threads)) as pool:
        def submit(s):
            return [[pool.submit(decode_rgb, dataset.resolve(p)) for p in paths_of(i)]
                    for i in range(s, min(s + micro_batch_size, n))]

        pending = {k: submit(starts[k]) for k in range(min(lookahead, len(starts)))}
        for k, s in enumerate(starts):
            if k + lookahead < len(starts):
                pending[k + lookahead] = submit(starts[k + lookahead])
            futs = pending.pop(k)
            items = []
            for j, i in enumerate(range(s, min(s + micro_batch_size, n))):
                imgs = [f.result() for f in futs[j]]
                if dataset.cls_based:
                    t = dataset.encode(dataset.prompts[i], imgs[0])
                    items.append((t["input_ids"], t["attention_mask"], t["pixel_values"], t["image_sizes"],
                                  dataset.label_list[i]))
                else:
                    pc, pr = dataset.prompts_of(i)
                    c, r = dataset.encode(pc, imgs[0]), dataset.encode(pr, imgs[1])
                    items.append((c["input_ids"], c["attention_mask"], c["pixel_values"], c["image_sizes"],
                                  r["input_ids"], r["attention_mask"], r["pixel_values"], r["image_sizes"],
                                  dataset.c_rates[i], dataset.r_rates[i]))
            out = dataset.collate_fn(items)
            if dataset.cls_based:
                yield squeeze_batch(out[0]), out[1]
            else:
                yield squeeze_batch(out[0]), squeeze_batch(out[1])
