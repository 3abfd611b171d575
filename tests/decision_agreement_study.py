"""Full-depth decision agreement (north_star: identical pairwise decisions on >= 99.9 % of pairs), run on the GPU box:
engine (bf16, this repo) vs the oracle = the reference's arithmetic in fp32 and in bf16 (eager attention) on N
synthetic config-2 pairs (Phi-3.5-V + SkipCA + LoRA + GPM, (1008,1344), S=2048, 23+32 layers).
Prints agreement rates next to the reference's own bf16-vs-fp32 flip rate (noise floor).
usage: python tests/decision_agreement_study.py [n_pairs]"""
import os
import sys
import types

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200.config import RewardConfig  # noqa: E402
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor, preference_compute  # noqa: E402
from llava_reward_b200.synth import SynthProvider, synth_batch  # noqa: E402
from oracle import reward_oracle as O  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.backends.cudnn.allow_tf32 = False
    ypath = "/tmp/da.yaml"
    with open(ypath, "w") as f:
        yaml.safe_dump({"is_general_preference": True, "add_cross_attention": True, "value_head_dim": 2,
                        "general_preference_tau": 0.1}, f)
    args = types.SimpleNamespace(pretrain="synthetic:1234", pm_path=None, cache_dir=None, ft_projector=False)
    args, model = load_reward_adaptor(args, "phi3v", ypath)
    model = model.to("cuda").eval()
    cfg = model.config
    P32 = O.Params(SynthProvider(cfg, seed=1234, device="cuda"), dtype=torch.float32, device="cuda", cache=False)
    P16 = O.Params(SynthProvider(cfg, seed=1234, device="cuda"), dtype=torch.bfloat16, device="cuda", cache=False)
    B = 4
    pe, p32, p16, re_, r32_, r16_ = [], [], [], [], [], []
    for i in range(0, n_pairs, B):
        rs = {}
        for tag in ("c", "r"):
            ids, mask, pix, sizes = synth_batch(cfg, B, (1008, 1344), 2048, seed=1000 + i, tag=tag, device="cuda",
                                                text_len_range=(35, 123))
            e, _ = model.custom_forward(ids, mask, pix, sizes)
            a, b = [], []
            with torch.no_grad():
                for k in range(B):  # oracle one sample at a time (eager attention memory); BT/GPM rewards of equal-size
                    sl = slice(k, k + 1)  # images are batch-invariant, so this equals the batched reference
                    a.append(O.custom_forward(P32, cfg, ids[sl], mask[sl], pix[sl], sizes[sl]))
                    b.append(O.custom_forward(P16, cfg, ids[sl], mask[sl], pix[sl], sizes[sl]))
            rs[tag] = (e, torch.cat(a), torch.cat(b))
        pe.append(torch.from_numpy(preference_compute(args, rs["c"][0], rs["r"][0])))
        p32.append(O.preference_compute(cfg, rs["c"][1], rs["r"][1]).cpu())
        p16.append(O.preference_compute(cfg, rs["c"][2], rs["r"][2]).cpu())
        for tag in ("c", "r"):
            re_.append(rs[tag][0].float().cpu()); r32_.append(rs[tag][1].cpu()); r16_.append(rs[tag][2].float().cpu())
        print(f"pairs {i + B}/{n_pairs}", flush=True)
    pe, p32, p16 = torch.cat(pe), torch.cat(p32), torch.cat(p16)
    re_, r32_, r16_ = torch.cat(re_), torch.cat(r32_), torch.cat(r16_)
    d = lambda a, b: ((a > 0.5) == (b > 0.5)).float().mean().item()  # noqa: E731
    print(f"pairs: {pe.numel()}  (config 2 shape, full depth, random-init weights seed 1234)")
    print(f"reward |engine - ref_fp32|: max {float((re_ - r32_).abs().max()):.4f} rms {float((re_ - r32_).pow(2).mean().sqrt()):.4f}")
    print(f"reward |ref_bf16 - ref_fp32|: max {float((r16_ - r32_).abs().max()):.4f} rms {float((r16_ - r32_).pow(2).mean().sqrt()):.4f}")
    print(f"reward |engine - ref_bf16|: max {float((re_ - r16_).abs().max()):.4f} rms {float((re_ - r16_).pow(2).mean().sqrt()):.4f}")
    print(f"decision agreement engine vs ref_fp32 : {d(pe, p32):.4f}")
    print(f"decision agreement ref_bf16 vs ref_fp32: {d(p16, p32):.4f}   (the reference's own precision flip rate)")
    print(f"decision agreement engine vs ref_bf16 : {d(pe, p16):.4f}")
    margin = (p32 - 0.5).abs()
    for thr in (0.0, 0.05, 0.1, 0.25):
        m = margin > thr
        if m.any():
            print(f"  pairs with |p_fp32 - 0.5| > {thr}: {int(m.sum())}  engine agreement {d(pe[m], p32[m]):.4f}  "
                  f"ref_bf16 agreement {d(p16[m], p32[m]):.4f}")
    flips = [(float(p32[i]), float(pe[i]), float(p16[i])) for i in range(pe.numel()) if (pe[i] > 0.5) != (p32[i] > 0.5)]
    print("engine flips (p_fp32, p_engine, p_ref_bf16):", flips)


if __name__ == "__main__":
    main()
