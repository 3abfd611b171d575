"""A/B on ONE box (box-to-box clock spread is ~4 %): the config-2 scoring step with engine.pack_rows on / off,
alternating, device-resident inputs, CUDA events. usage (GPU box): python tools/ab_pack.py"""
import os, sys, types
import torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from llava_reward_b200.reward_adaptor_loader import load_reward_adaptor
from llava_reward_b200.synth import synth_batch

ypath = "/tmp/llava_reward_b200_ab.yaml"
with open(ypath, "w") as f:
    yaml.safe_dump({"is_general_preference": True, "add_cross_attention": True, "value_head_dim": 2,
                    "general_preference_tau": 0.1}, f)
args = types.SimpleNamespace(pretrain="synthetic:1234", pm_path=None, cache_dir=None, ft_projector=False)
args, model = load_reward_adaptor(args, "phi3v", ypath)
model = model.to("cuda").eval()
eng, cfg = model.engine, model.config
data = {t: synth_batch(cfg, 32, (1008, 1344), 2048, seed=7, tag=t, device="cuda", text_len_range=(35, 123)) for t in "cr"}


def step():
    rs = {t: model.custom_forward(*data[t])[0] for t in "cr"}
    return eng.preference(rs["c"], rs["r"])


for _ in range(3):
    step()
res = {True: [], False: []}
for rnd in range(4):
    for flag in (True, False):
        eng.pack_rows = flag
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record()
        torch.cuda.synchronize()
        res[flag].append(e0.elapsed_time(e1) / 3)
for flag in (True, False):
    ms = sorted(res[flag])
    print(f"pack_rows={flag}: ms/step {['%.1f' % m for m in res[flag]]} median {ms[len(ms) // 2]:.1f} -> "
          f"{32e3 / ms[len(ms) // 2]:.2f} pairs/s")
print(f"valid rows per forward: {int(data['c'][1].sum())} of {data['c'][1].numel()}")
