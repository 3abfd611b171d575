"""Full-depth reward / decision parity on the GPU box against the UNMODIFIED reference (baseline/_ref):

    python tools/parity_study.py [n_pairs=512] [--batch 8] [--eager-pairs 32] [--out gpurun_out/parity_study.json]

For N synthetic BASELINE configs[1] pairs (Phi-3.5-V + SkipCA + LoRA r128 + GPM, (1008,1344) -> 13 crops, S=2048,
23 CLIP + 32 decoder layers, random-init seed 1234) it scores every sample with

  engine      this repo, bf16, through load_reward_adaptor / custom_forward
  ref_fp32    the reference model in fp32 on the GPU (eager attention, TF32 off)  = ground truth
  ref_fa2     the reference model in bf16 with its flash-attention-2 path          = what the reference ships on a GPU
  ref_eager   the reference model in bf16, eager attention (first --eager-pairs pairs only; it is slow)

and reports reward errors and pairwise-decision agreement (north_star: rewards within 2e-2 in bf16, >= 99.9 %
identical decisions) next to the reference's own bf16-vs-fp32 numbers. Decisions are reported on the N (c_i, r_i)
pairs and, as a tighter estimate of the same flip rate, on all N x N cross pairs (c_i, r_j) of the same rewards.
Results are bit-reproducible inputs: synth_batch(seed=1000 + first pair index of the micro-batch).
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor  # noqa: E402
from llava_reward_b200.synth import synth_batch  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402


def gpm_prob(c, r, tau):
    """preference_compute (eval/reward_adaptor_loader.py:174-181) for value_head_dim 2 in fp64 on given rewards"""
    c, r = c.double(), r.double()
    return torch.sigmoid((c[..., 0] * r[..., 1] - c[..., 1] * r[..., 0]) / tau)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("n_pairs", type=int, nargs="?", default=512)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--eager-pairs", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_study.json"))
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    ypath = "/tmp/parity_study.yaml"
    with open(ypath, "w") as f:
        yaml.safe_dump({"is_general_preference": True, "add_cross_attention": True, "value_head_dim": 2,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:1234", pm_path=None, cache_dir=None, ft_projector=False)
    args, model = load_reward_adaptor(args, "phi3v", ypath)
    model = model.to(dev).eval()
    cfg = model.config
    t0 = time.time()
    ref32 = RH.build_reference_model(cfg, 1234, device=dev, dtype=torch.float32)
    ref16 = RH.build_reference_model(cfg, 1234, device=dev, dtype=torch.bfloat16)
    print(f"models ready in {time.time() - t0:.0f}s, {torch.cuda.memory_allocated() / 2**30:.1f} GiB allocated", flush=True)
    B = a.batch
    out = {k: {"c": [], "r": []} for k in ("engine", "ref_fp32", "ref_fa2", "ref_eager")}
    secs = {k: 0.0 for k in out}

    def timed(key, fn):
        torch.cuda.synchronize()
        t = time.time()
        with torch.no_grad():
            r = fn()
        torch.cuda.synchronize()
        secs[key] += time.time() - t
        return r.float().cpu()

    for i in range(0, a.n_pairs, B):
        for tag in ("c", "r"):
            ids, mask, pix, sizes = synth_batch(cfg, B, (1008, 1344), 2048, seed=1000 + i, tag=tag, device=dev,
                                                text_len_range=(35, 123))
            out["engine"][tag].append(timed("engine", lambda: model.custom_forward(ids, mask, pix, sizes)[0]))
            # the slow reference passes (fp32, eager) are given the 13 real crop slots only: the 4 zero-padded slots are never read by
            # hd_feature_transform (modeling_phi3_v.py:276-282), so the rewards are identical and the fp32 pass is 25 % cheaper
            pr = pix[:, :13].contiguous()
            RH.set_attention(ref16, "flash_attention_2")
            out["ref_fa2"][tag].append(timed("ref_fa2", lambda: ref16.custom_forward(ids, mask, pix, sizes)[0]))
            if i < a.eager_pairs:
                RH.set_attention(ref16, "eager")
                out["ref_eager"][tag].append(timed("ref_eager", lambda: ref16.custom_forward(ids, mask, pr, sizes)[0]))
            out["ref_fp32"][tag].append(timed("ref_fp32", lambda: ref32.custom_forward(ids, mask, pr, sizes)[0]))
        print(f"pairs {i + B}/{a.n_pairs}  seconds so far {json.dumps({k: round(v, 1) for k, v in secs.items()})}", flush=True)
    R = {k: {t: torch.cat(v[t]) for t in v if v[t]} for k, v in out.items()}
    tau = cfg.general_preference_tau
    res = {"n_pairs": a.n_pairs, "config": "BASELINE configs[1] shape, full depth (23 CLIP + 32 decoder layers), seed 1234",
           "seconds": secs, "reward": {}, "decisions_pairs": {}, "decisions_cross": {}}

    def err(x, y, n=None):
        d = torch.cat([(R[x][t][:n] - R[y][t][:n]).abs().flatten() for t in ("c", "r")])
        return {"max": d.max().item(), "rms": d.pow(2).mean().sqrt().item(), "p99": d.quantile(0.99).item(),
                "frac_within_2e-2": (d <= 2e-2).float().mean().item(), "n_values": d.numel()}

    ne = R["ref_eager"]["c"].shape[0] if R["ref_eager"] else 0
    res["reward"]["engine_vs_ref_fa2"] = err("engine", "ref_fa2")
    res["reward"]["engine_vs_ref_fp32"] = err("engine", "ref_fp32")
    res["reward"]["ref_fa2_vs_ref_fp32"] = err("ref_fa2", "ref_fp32")
    if ne:
        res["reward"]["ref_eager_vs_ref_fa2"] = err("ref_eager", "ref_fa2", ne)
        res["reward"]["engine_vs_ref_eager"] = err("engine", "ref_eager", ne)
        res["reward"]["engine_vs_ref_fa2_same_subset"] = err("engine", "ref_fa2", ne)
        res["reward"]["ref_eager_vs_ref_fp32"] = err("ref_eager", "ref_fp32", ne)
    P = {k: gpm_prob(R[k]["c"], R[k]["r"], tau) for k in ("engine", "ref_fp32", "ref_fa2")}
    X = {k: gpm_prob(R[k]["c"][:, None, :], R[k]["r"][None, :, :], tau) for k in ("engine", "ref_fp32", "ref_fa2")}

    def agree(p, q, m=None):
        s = (p > 0.5) == (q > 0.5)
        if m is not None:
            s = s[m]
        return {"agree": int(s.sum()), "of": int(s.numel()), "rate": float(s.float().mean()) if s.numel() else None}

    for name, D in (("decisions_pairs", P), ("decisions_cross", X)):
        margin = (D["ref_fp32"] - 0.5).abs()
        # noise band: the largest fp32 margin at which EITHER bf16 implementation flips a decision
        fl = ((D["engine"] > 0.5) != (D["ref_fp32"] > 0.5)) | ((D["ref_fa2"] > 0.5) != (D["ref_fp32"] > 0.5))
        band = float(margin[fl].max()) if fl.any() else 0.0
        res[name] = {"engine_vs_ref_fp32": agree(D["engine"], D["ref_fp32"]),
                     "ref_fa2_vs_ref_fp32": agree(D["ref_fa2"], D["ref_fp32"]),
                     "engine_vs_ref_fa2": agree(D["engine"], D["ref_fa2"]),
                     "noise_band_abs_margin": band,
                     "engine_flip_max_margin": float(margin[(D["engine"] > 0.5) != (D["ref_fp32"] > 0.5)].max())
                     if ((D["engine"] > 0.5) != (D["ref_fp32"] > 0.5)).any() else 0.0,
                     "ref_fa2_flip_max_margin": float(margin[(D["ref_fa2"] > 0.5) != (D["ref_fp32"] > 0.5)].max())
                     if ((D["ref_fa2"] > 0.5) != (D["ref_fp32"] > 0.5)).any() else 0.0}
        for thr in (0.01, 0.02, 0.05, 0.1):
            m = margin > thr
            res[name][f"margin>{thr}"] = {"engine_vs_ref_fp32": agree(D["engine"], D["ref_fp32"], m),
                                          "ref_fa2_vs_ref_fp32": agree(D["ref_fa2"], D["ref_fp32"], m)}
    res["rewards_head"] = {k: R[k]["c"][:4].tolist() for k in R if R[k]}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    np.savez(a.out.replace(".json", ".npz"), **{f"{k}_{t}": R[k][t].numpy() for k in R for t in R[k]})
    print(json.dumps(res, indent=1))
    # the bar of VERDICT r01 item 2
    e, r = res["decisions_pairs"]["engine_vs_ref_fp32"]["rate"], res["decisions_pairs"]["ref_fa2_vs_ref_fp32"]["rate"]
    ok = e >= r - 0.005
    print(f"ENGINE decision agreement with fp32 {e:.4f} vs the reference's own bf16 {r:.4f}: {'OK' if ok else 'BELOW'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
