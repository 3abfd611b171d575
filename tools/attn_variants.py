"""Compare compile-time variants of the tcgen05 attention kernel (softmax exponentials split between the MUFU and the
FMA pipe, early release of the S buffer): builds one private .so per variant (here, no GPU needed: --build-only) and,
on the GPU box, times every variant on the three product shapes and checks it against an fp32 reference.
usage: python tools/attn_variants.py --build-only ; (GPU box) python tools/attn_variants.py

SUPERSEDED by tools/attn_ab.py: this script times one variant after the other, and on a power-capped B200 the first
variant then always wins (identical object code measured 1.49 ms first and 1.68 ms sixth). Kept because the r01 / early
r02 logs under profiles/ were produced with it; do not use it for effects below ~15 %."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llava-reward_b200", "csrc")
OUTDIR = os.path.join(ROOT, "llava-reward_b200", "lib", "variants")
# (LR_ATTN_POLY_NUM, LR_ATTN_EARLY_SFREE, LR_ATTN_HOIST_DESC, LR_ATTN_AUX_REGS, LR_ATTN_PIPE_LD, LR_ATTN_MAX3, extra -D
# string); r01 measured the first six (all slower than (0, 0, 0, 40)); then the MMA-issue-path experiments of DESIGN.md
# section 11; --r02 = the softmax-side experiments of round 2
VARIANTS = [(0, 0, 0, 40, 0, 0), (0, 1, 0, 40, 0, 0), (1, 1, 0, 40, 0, 0), (2, 1, 0, 40, 0, 0), (1, 0, 0, 40, 0, 0),
            (3, 1, 0, 40, 0, 0), (0, 0, 1, 40, 0, 0), (0, 0, 0, 48, 0, 0), (0, 0, 1, 48, 0, 0)]
if "--mma-only" in sys.argv:
    VARIANTS = [v for v in VARIANTS if v[0] == 0 and v[1] == 0]
if "--r02" in sys.argv:
    VARIANTS = [(0, 0, 0, 40, 0, 0), (0, 0, 0, 40, 1, 0), (0, 0, 0, 40, 0, 1), (0, 0, 0, 40, 1, 1), (0, 0, 1, 48, 1, 1),
                (0, 0, 0, 48, 1, 1), (1, 0, 0, 48, 1, 1)]
for a in sys.argv:
    if a.startswith("--only="):   # --only=p,e,h,r,l,m[;p,e,h,r,l,m...]
        VARIANTS = [tuple(int(x) for x in t.split(",")) for t in a[len("--only="):].split(";")]
EXTRA = [a[len("--define="):] for a in sys.argv if a.startswith("--define=")]


def so_path(v):
    tag = "".join("_" + e.replace("=", "") for e in EXTRA)
    return os.path.join(OUTDIR, f"libattn_p{v[0]}e{v[1]}h{v[2]}r{v[3]}l{v[4]}m{v[5]}{tag}.so")


def build():
    os.makedirs(OUTDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("attention_tc.cu", "attention.cu")]
    procs = []
    for v in VARIANTS:
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "--use_fast_math", "--prec-div=true", "--prec-sqrt=true", "--fmad=true",
               f"-DLR_ATTN_POLY_NUM={v[0]}", f"-DLR_ATTN_EARLY_SFREE={v[1]}", f"-DLR_ATTN_HOIST_DESC={v[2]}",
               f"-DLR_ATTN_AUX_REGS={v[3]}", f"-DLR_ATTN_PIPE_LD={v[4]}", f"-DLR_ATTN_MAX3={v[5]}",
               *[f"-D{e}" for e in EXTRA], "-shared", "-o", so_path(v), *srcs,
               "-lcudart"]
        procs.append(subprocess.Popen(cmd))
    for p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed")


def main():
    if "--build-only" in sys.argv or not all(os.path.exists(so_path(v)) for v in VARIANTS):
        build()
        if "--build-only" in sys.argv:
            return
    import torch
    import torch.nn.functional as F

    p, i32, f32 = C.c_void_p, C.c_int, C.c_float
    bf = torch.bfloat16
    ATTN_TCGEN05 = 0
    shapes = {"clip hd64": (416, 577, 16, 64, False), "decoder hd96 causal": (32, 2048, 32, 96, True),
              "llava hd128 causal": (16, 3057, 32, 128, True)}
    data = {}
    torch.manual_seed(0)
    for name, (nseq, T, heads, hd, causal) in shapes.items():
        D = heads * hd
        qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
        q, k, v = (qkv[:T, i * D:(i + 1) * D].float().reshape(T, heads, hd).transpose(0, 1)[None] for i in range(3))
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal)[0].transpose(0, 1).reshape(T, D)
        data[name] = (qkv, torch.empty(nseq * T, D, device="cuda", dtype=bf), ref)
    for v in VARIANTS:
        lib = C.CDLL(so_path(v))
        lib.lr_attention_bf16.argtypes = [p, p, p, p, i32, i32, i32, i32, p, p, i32, i32, i32, f32, i32, p]
        for name, (nseq, T, heads, hd, causal) in shapes.items():
            D = heads * hd
            qkv, o, ref = data[name]
            fl = 4.0 * nseq * heads * T * T * hd * (0.5 if causal else 1.0)

            def run():
                st = lib.lr_attention_bf16(qkv.data_ptr(), qkv[:, D:].data_ptr(), qkv[:, 2 * D:].data_ptr(), o.data_ptr(),
                                           3 * D, D, nseq, T, None, None, heads, hd, int(causal), hd ** -0.5,
                                           ATTN_TCGEN05, torch.cuda.current_stream().cuda_stream)
                assert st == 0, st

            for _ in range(3):
                run()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    run()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 10)
            err = ((o[:T].float() - ref).norm() / ref.norm()).item()
            print(f"poly {v[0]}/4 early_sfree {v[1]} hoist_desc {v[2]} aux_regs {v[3]} pipe_ld {v[4]} max3 {v[5]} {' '.join(EXTRA)} | {name}: {best:.3f} ms = {fl / best / 1e9:.0f} TF/s | rel L2 err vs fp32 "
                  f"{err:.3e}", flush=True)


if __name__ == "__main__":
    main()
