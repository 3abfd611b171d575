"""Batch evaluation callers around the hot path: the collate step of
`GeneralRewardDataset.collate_fn` (reference llava_reward/datasets/reward_dataset.py:137-202, left padding via
`zero_pad_sequences`, datasets/utils.py:5-13) and the two loops of `batch_rm_inference`
(reference eval/batch_inference_rm_phi.py:70-152, eval/batch_inference_rm_llava.py:70-152 and
eval/batch_inference_rm_qwen.py:75-146 - the same loops with the `inputs_batch=` calling convention): pairwise preference accuracy and single-image (BT / cls) scoring.
Inputs are per-sample dicts as produced by `inference_process_phi3v` / the processor; everything stays on the GPU
until the final numpy conversion.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def zero_pad_sequences(sequences: Sequence[torch.Tensor], side: str = "left", value=0) -> torch.Tensor:
    """Pad 1-D/2-D tensors along the last dim to the longest and stack (reference datasets/utils.py:5-13)."""
    assert side in ("left", "right")
    max_len = max(s.size(-1) for s in sequences)
    out = []
    for s in sequences:
        pad = max_len - s.size(-1)
        out.append(F.pad(s, (pad, 0) if side == "left" else (0, pad), value=value))
    return torch.stack(out, dim=0)


def collate_samples(items: Sequence[Dict[str, torch.Tensor]], pad_token_id: int) -> Dict[str, torch.Tensor]:
    """List of per-sample inputs ([1,S_i] ids/mask, [1,17,3,336,336] pixels, [1,2] sizes) -> one left-padded batch
    with the `squeeze(1)` of the reference's eval loop already applied (eval/batch_inference_rm_phi.py:82-90)."""
    ids = zero_pad_sequences([it["input_ids"] for it in items], value=pad_token_id)
    mask = zero_pad_sequences([it["attention_mask"] for it in items])
    pix = torch.stack([it["pixel_values"] for it in items], dim=0)
    sizes = torch.stack([torch.as_tensor(it["image_sizes"]) for it in items], dim=0)
    return {"input_ids": ids.squeeze(1), "attention_mask": mask.squeeze(1), "pixel_values": pix.squeeze(1),
            "image_sizes": sizes.squeeze(1)}


def _forward(model, batch):
    """One scoring call in the backbone's own convention: positional tensors for phi3v
    (eval/batch_inference_rm_phi.py:93), the processor's BatchFeature as `inputs_batch` for llava
    (eval/batch_inference_rm_llava.py:86-87) and qwen (eval/batch_inference_rm_qwen.py:91-92)."""
    if getattr(model, "model_type", "phi3v") in ("llava", "qwen"):
        return model.custom_forward(inputs_batch=batch)[0]
    return model.custom_forward(batch["input_ids"], batch["attention_mask"], batch["pixel_values"],
                                batch["image_sizes"])[0]


@torch.no_grad()
def score_pairs(model, args, batches: Iterable) -> Dict[str, object]:
    """Pairwise mode (eval/batch_inference_rm_phi.py:70-121). `batches` yields (batch_chosen, batch_rejected) dicts.
    Returns the per-pair probabilities and the summary numbers the reference prints."""
    from .reward_adaptor_loader import preference_compute
    probs: List[float] = []
    chosen_rewards, reject_rewards = [], []
    for bc, br in batches:
        rc, rr = _forward(model, bc), _forward(model, br)
        if not args.is_general_preference:
            chosen_rewards.extend(rc.squeeze(-1).tolist())
            reject_rewards.extend(rr.squeeze(-1).tolist())
        probs.extend(preference_compute(args, rc, rr).tolist())
    total = len(probs)
    wins = sum(p > 0.5 for p in probs)
    ties = sum(p == 0.5 for p in probs)
    return {"probs": np.asarray(probs, dtype=np.float32), "prob_mean": float(np.mean(probs)) if total else float("nan"),
            "proportion": wins / total if total else float("nan"),
            "proportion_wo_tie": wins / (total - ties) if total - ties else float("nan"),
            "chosen_rewards": chosen_rewards, "reject_rewards": reject_rewards}


def binary_metrics(pred: Sequence[int], label: Sequence[int]) -> Dict[str, float]:
    pred, label = np.asarray(pred).astype(int), np.asarray(label).astype(int)
    tp = int(((pred == 1) & (label == 1)).sum())
    fp = int(((pred == 1) & (label == 0)).sum())
    fn = int(((pred == 0) & (label == 1)).sum())
    precision = tp / (tp + fp) if tp + fp else 0.0
    recall = tp / (tp + fn) if tp + fn else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall else 0.0
    return {"accuracy": float((pred == label).mean()) if len(label) else float("nan"), "f1": f1, "recall": recall}


@torch.no_grad()
def score_single(model, args, batches: Iterable, cls_based: bool = False) -> Dict[str, object]:
    """Single-image mode (eval/batch_inference_rm_phi.py:123-152). `batches` yields (batch, labels)."""
    if args.is_general_preference:
        raise ValueError("General preference loss-based model is not supported for single image evaluation. "
                         "Please use BT model instead.")
    rewards, preds, labels = [], [], []
    for batch, lab in batches:
        r = _forward(model, batch)
        rewards.extend(r.squeeze(-1).tolist())
        labels.extend(torch.as_tensor(lab).tolist())
        if cls_based:
            preds.extend((torch.sigmoid(r.float()).squeeze(-1) >= 0.5).long().tolist())
    out = {"rewards": rewards, "labels": labels}
    if cls_based:
        out.update(binary_metrics(preds, labels))
    return out


@torch.no_grad()
def best_of_n(model, batches: Iterable) -> Dict[str, object]:
    """Inference-time scaling (BASELINE.json configs[4]): score N candidate images of one prompt with a BT model and
    return the rewards and the arg-max candidate. `batches` yields processor outputs covering the N candidates."""
    rewards: List[float] = []
    for batch in batches:
        r = _forward(model, batch)
        if r.shape[1] != 1:
            raise ValueError("best-of-N selection needs a scalar (BT) reward head")
        rewards.extend(r.float().squeeze(-1).tolist())
    return {"rewards": rewards, "best": int(np.argmax(rewards)) if rewards else -1}
