"""Interleaved A/B timing of compile-time variants of the tcgen05 attention kernel.

tools/attn_variants.py times variant after variant; on a power-capped B200 the FIRST variant then always wins (the same
object code measured 1.49 ms first and 1.68 ms sixth, profiles/r02_attention_softmax_variants.txt, hd-128 lines), so
its verdicts on +-5 % effects are not reliable. This tool loads every variant, heats the GPU for a few seconds, then
walks ROUNDS rounds; in each round every (variant, shape) gets 10 launches between two CUDA events, the variant order
rotating from round to round. Reported: median and best over the rounds, the SM clock sampled right after each
measurement (NVML), and the relative L2 error against an fp32 reference.

usage: python tools/attn_ab.py --build-only      (build container: nvcc cross-compiles, no GPU)
       python tools/attn_ab.py [--rounds N]       (GPU box)
"""
import ctypes as C
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llava-reward_b200", "csrc")
OUTDIR = os.path.join(ROOT, "llava-reward_b200", "lib", "variants")

# name -> -D defines. Every entry names ALL the hand-off switches (P_HALF, EXP_FIRST, DESC32, MMA_WARP) explicitly, so
# a line of an old log keeps its meaning when the product defaults move; "base" = the product defaults of the source.
def _v(p_half=0, exp_first=0, desc32=0, mma_warp=0, **extra):
    d = {"LR_ATTN_P_HALF": p_half, "LR_ATTN_EXP_FIRST": exp_first, "LR_ATTN_DESC32": desc32, "LR_ATTN_MMA_WARP": mma_warp}
    d.update(extra)
    return d


VARIANTS = {
    "base": {},
    "old": None,   # the product object code from before the hand-off experiments were added (kept .so, never rebuilt)
    "r02a": _v(),                                   # the hand-off switches of the first half of round 2 (other defaults)
    "r02a_full": _v(LR_ATTN_CHUNK_MASK=0, LR_ATTN_FFMA2=0),   # = the product of the first half of round 2, every later switch off
    "hoist40": _v(LR_ATTN_HOIST_DESC=1),
    "aux48": _v(LR_ATTN_AUX_REGS=48),
    "aux56": _v(LR_ATTN_AUX_REGS=56),
    "desc32": _v(desc32=1),
    "desc32_aux48": _v(desc32=1, LR_ATTN_AUX_REGS=48),
    "phalf": _v(p_half=1),
    "phalf_desc32": _v(p_half=1, desc32=1),
    "expfirst": _v(exp_first=1),
    "epd": _v(1, 1, 1),
    "expfirst_phalf_desc32": _v(1, 1, 1),
    "kv2": _v(LR_ATTN_KV2=1),
    "kv2_phalf_desc32": _v(p_half=1, desc32=1, LR_ATTN_KV2=1),
    "max3": _v(LR_ATTN_MAX3=1),
    "pipeld": _v(LR_ATTN_PIPE_LD=1),
    "esfree": _v(LR_ATTN_EARLY_SFREE=1),
    "poly1": _v(LR_ATTN_POLY_NUM=1),
    **{f"stag1_{d}": _v(LR_ATTN_STAGGER=d) for d in (400, 700, 1000, 1300, 1600, 1900, 2500)},
    **{f"stag2_{d}": _v(LR_ATTN_STAGGER=d, LR_ATTN_STAGGER_MODE=2) for d in (700, 1300, 1900)},
    **{f"epd_stag1_{d}": _v(1, 1, 1, LR_ATTN_STAGGER=d) for d in (1000, 1300, 1900)},
    "epd_esfree": _v(1, 1, 1, LR_ATTN_EARLY_SFREE=1),
    "pd_esfree": _v(1, 0, 1, LR_ATTN_EARLY_SFREE=1),
    **{f"epd_nt2stag{d}": _v(1, 1, 1, LR_ATTN_NT2_STAGGER=d) for d in (1, 500, 1700, 2300)},
    "mw": _v(mma_warp=1),
    "mw_d": _v(desc32=1, mma_warp=1),
    "mw_pd": _v(1, 0, 1, 1),
    "mw_epd": _v(1, 1, 1, 1),
    "mw2_epd": _v(1, 1, 1, 2),                      # = the product after the hand-off work
    "cm": _v(1, 1, 1, 2, LR_ATTN_CHUNK_MASK=1),
    "cm0": _v(1, 1, 1, 2, LR_ATTN_CHUNK_MASK=0),
    "ffma2": _v(1, 1, 1, 2, LR_ATTN_CHUNK_MASK=1, LR_ATTN_FFMA2=1),
    "es": {"LR_ATTN_EPI_STAGE": 1},
    "es0": {"LR_ATTN_EPI_STAGE": 0},
    **{f"sleep{d}": {"LR_ATTN_POLL_SLEEP": d} for d in (20, 50, 100, 200)},
    "l2pf": {"LR_ATTN_L2_PREFETCH": 1},
    "p_poly1": {"LR_ATTN_POLY_NUM": 1},
    "p_poly2": {"LR_ATTN_POLY_NUM": 2},
    "base@split2": {"_so": "base", "_env": {"LR_ATTN_VARIANT": "6"}},
    "split104": {"LR_ATTN_SPLIT_REGS": 104},
    "split104@split2": {"_so": "split104", "_env": {"LR_ATTN_VARIANT": "6"}},
    "preload": {"LR_ATTN_PRELOAD": 1},
    "preload0": {"LR_ATTN_PRELOAD": 0},
    "one_cta_per_sm": {"LR_ATTN_PAD_SMEM": 102400},
    "no_ones": {"LR_ATTN_NO_ONES": 1},
    "phyb": {"LR_ATTN_P_HYBRID": 1},
    "specmax": {"LR_ATTN_SPEC_MAX": 1},
    "cm_spin": _v(1, 1, 1, 2, LR_ATTN_CHUNK_MASK=1, LR_ATTN_SPIN_WAIT=1),
    "mw_epd_aux48": _v(1, 1, 1, 1, LR_ATTN_AUX_REGS=48),
    "mw_epd_esfree": _v(1, 1, 1, 1, LR_ATTN_EARLY_SFREE=1),
    # "@nt2": the same object code as the named variant, launched as two query tiles per CTA / one CTA per SM for
    # every head_dim (LR_ATTN_VARIANT=1 in the environment)
    "r02a@nt2": {"_so": "r02a", "_env": {"LR_ATTN_VARIANT": "1"}},
    "epd@nt2": {"_so": "epd", "_env": {"LR_ATTN_VARIANT": "1"}},
    "epd_esfree@nt2": {"_so": "epd_esfree", "_env": {"LR_ATTN_VARIANT": "1"}},
    "mw_epd@nt2": {"_so": "mw_epd", "_env": {"LR_ATTN_VARIANT": "1"}},
}
for a in sys.argv:
    if a.startswith("--only="):
        keep = a[len("--only="):].split(",")
        VARIANTS = {k: VARIANTS[k] for k in keep}


def so_path(name):
    defs = VARIANTS.get(name) or {}
    return os.path.join(OUTDIR, f"libattn_ab_{defs.get('_so', name)}.so")


def build():
    os.makedirs(OUTDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("attention_tc.cu", "attention.cu")]
    procs = []
    for name, defs in VARIANTS.items():
        if defs is None or "_so" in defs:
            continue
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "--use_fast_math", "--prec-div=true", "--prec-sqrt=true", "--fmad=true",
               *[f"-D{k}={v}" for k, v in defs.items()], "-shared", "-o", so_path(name), *srcs, "-lcudart"]
        procs.append((name, subprocess.Popen(cmd)))
        if len(procs) % 8 == 0:
            for _, p in procs[-8:]:
                p.wait()
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"nvcc failed on {name}")


def main():
    if "--build-only" in sys.argv or not all(os.path.exists(so_path(v)) for v in VARIANTS):
        build()
        if "--build-only" in sys.argv:
            return
    rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 7
    import torch
    import torch.nn.functional as F
    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(0)
        clock = lambda: pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)
    except Exception:
        clock = lambda: 0

    p, i32, f32 = C.c_void_p, C.c_int, C.c_float
    bf = torch.bfloat16
    shapes = {"clip hd64": (416, 577, 16, 64, False), "decoder hd96 causal": (32, 2048, 32, 96, True),
              "llava hd128 causal": (16, 3057, 32, 128, True)}
    if "--no-llava" in sys.argv:
        shapes.pop("llava hd128 causal")
    data = {}
    torch.manual_seed(0)
    for name, (nseq, T, heads, hd, causal) in shapes.items():
        D = heads * hd
        qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
        q, k, v = (qkv[:T, i * D:(i + 1) * D].float().reshape(T, heads, hd).transpose(0, 1)[None] for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal)[0].transpose(0, 1).reshape(T, D)
        # the LAST sequence as well (multi-tile CTAs, tail handling)
        ql, kl, vl = (qkv[-T:, i * D:(i + 1) * D].float().reshape(T, heads, hd).transpose(0, 1)[None] for i in range(3))
        refl = F.scaled_dot_product_attention(ql, kl, vl, is_causal=causal)[0].transpose(0, 1).reshape(T, D)
        data[name] = (qkv, torch.empty(nseq * T, D, device="cuda", dtype=bf), ref, refl)
    libs = {}
    for vn in VARIANTS:
        lib = C.CDLL(so_path(vn))
        lib.lr_attention_bf16.argtypes = [p, p, p, p, i32, i32, i32, i32, p, p, i32, i32, i32, f32, i32, p]
        lib._lr_env = (VARIANTS[vn] or {}).get("_env", {})
        libs[vn] = lib

    def run(lib, name):
        os.environ["LR_ATTN_VARIANT"] = lib._lr_env.get("LR_ATTN_VARIANT", "0")   # read by the library on every call
        nseq, T, heads, hd, causal = shapes[name]
        D = heads * hd
        qkv, o = data[name][:2]
        st = lib.lr_attention_bf16(qkv.data_ptr(), qkv[:, D:].data_ptr(), qkv[:, 2 * D:].data_ptr(), o.data_ptr(),
                                   3 * D, D, nseq, T, None, None, heads, hd, int(causal), hd ** -0.5, 0,
                                   torch.cuda.current_stream().cuda_stream)
        return st

    # correctness of every variant first (also the warm-up of every kernel)
    errs = {}
    base_out, same = {}, {}
    unsupported = set()
    for vn, lib in libs.items():
        for name in shapes:
            T = shapes[name][1]
            o = data[name][1]
            o.zero_()
            if run(lib, name) != 0:      # this variant does not support this shape
                errs[(vn, name)] = 0.0
                unsupported.add((vn, name))
                continue
            torch.cuda.synchronize()
            ref, refl = data[name][2:]
            e0 = ((o[:T].float() - ref).norm() / ref.norm()).item()
            e1 = ((o[-T:].float() - refl).norm() / refl.norm()).item()
            errs[(vn, name)] = max(e0, e1)
            if vn == "base":
                base_out[name] = o.clone()
            elif name in base_out:
                same[(vn, name)] = bool(torch.equal(o, base_out[name]))
            if not (errs[(vn, name)] < 5e-3):
                print(f"!! {vn} | {name}: rel L2 err {e0:.3e} / {e1:.3e}", flush=True)
        ok = all(errs[(vn, name)] < 5e-3 for name in shapes)
        print(("CHECK_OK " if ok else "CHECK_BAD ") + vn, flush=True)
    if "--check-only" in sys.argv:
        return
    # heat: ~4 s of the decoder shape with the base variant
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first = next(iter(libs.values()))
    for _ in range(3):
        for _ in range(1000):
            run(first, "decoder hd96 causal")
        torch.cuda.synchronize()
    times = {(vn, name): [] for vn in libs for name in shapes}
    clocks = {(vn, name): [] for vn in libs for name in shapes}
    names = list(libs)
    for r in range(rounds):
        order = names[r % len(names):] + names[:r % len(names)]
        if r % 2:
            order = order[::-1]
        for vn in order:
            for name in shapes:
                if (vn, name) in unsupported:
                    times[(vn, name)].append(float("nan"))
                    clocks[(vn, name)].append(0)
                    continue
                run(libs[vn], name)
                e0.record()
                for _ in range(10):
                    run(libs[vn], name)
                e1.record()
                torch.cuda.synchronize()
                times[(vn, name)].append(e0.elapsed_time(e1) / 10)
                clocks[(vn, name)].append(clock())
    print(f"# interleaved A/B, {rounds} rounds x 10 launches per (variant, shape); ms = median over rounds [best]; "
          f"rel = median / base median")
    for name in shapes:
        nseq, T, heads, hd, causal = shapes[name]
        fl = 4.0 * nseq * heads * T * T * hd * (0.5 if causal else 1.0)
        base = statistics.median(times[("base", name)]) if "base" in libs else None
        for vn in names:
            t = times[(vn, name)]
            med = statistics.median(t)
            print(f"{name:22s} | {vn:24s} | {med:.4f} ms [{min(t):.4f}] = {fl / med / 1e9:5.0f} TF/s | "
                  f"rel {med / base if base else 0:.3f} | sm clock {statistics.median(clocks[(vn, name)]):.0f} MHz | "
                  f"err {errs[(vn, name)]:.2e}" + (f" | bits == base: {same[(vn, name)]}" if (vn, name) in same else ""),
                  flush=True)


if __name__ == "__main__":
    main()
