"""Timeline of one CTA of the tcgen05 attention kernel (debug tool). Builds a private .so with -DLR_ATTN_TRACE.
usage (GPU box): python tools/attn_trace.py"""
import ctypes as C
import glob
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llava-reward_b200", "csrc")
EXTRA = [a for a in os.environ.get("LR_TRACE_DEFS", "").split() if a]   # e.g. "-DLR_ATTN_P_HALF=1 -DLR_ATTN_DESC32=1"
TAG = "".join(c for c in "".join(EXTRA) if c.isalnum())
OUT = os.path.join(ROOT, "llava-reward_b200", "lib", f"libllavareward_trace{TAG}.so")


def build():
    srcs = [os.path.join(CSRC, f) for f in ("attention_tc.cu", "attention.cu")]
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler",
           "-fPIC", "--use_fast_math", "--prec-div=true", "--prec-sqrt=true", "--fmad=true", "-DLR_ATTN_TRACE", *EXTRA, "-shared", "-o", OUT, *srcs, "-lcudart"]
    subprocess.run(cmd, check=True)


IMPL = int(sys.argv[sys.argv.index("--impl") + 1]) if "--impl" in sys.argv else 3  # 3 = two tiles per CTA, 4 = one


def main():
    if not os.path.exists(OUT) or "--build" in sys.argv or "--build-only" in sys.argv:
        build()
    if "--build-only" in sys.argv:
        return
    lib = C.CDLL(OUT)
    p, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.lr_attention_bf16.argtypes = [p, p, p, p, i32, i32, i32, i32, p, p, i32, i32, i32, f32, i32, p]
    lib.lr_attn_trace_set.argtypes = [p]
    bf = torch.bfloat16
    cases = {"dec": (32, 2048, 32, 96, True), "clip": (416, 577, 16, 64, False)}
    if "--solo" in sys.argv:   # one sequence, one head: every CTA has an SM to itself (no co-resident CTA)
        cases = {"dec solo": (1, 2048, 1, 96, True), "clip solo": (1, 577, 1, 64, False)}
    for name, (nseq, T, heads, hd, causal) in cases.items():
        D = heads * hd
        qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
        o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
        trace = torch.zeros(3 * 8 * 32, dtype=torch.int64, device="cuda")
        assert lib.lr_attn_trace_set(trace.data_ptr()) == 0
        for _ in range(2):
            st = lib.lr_attention_bf16(qkv.data_ptr(), qkv[:, D:].data_ptr(), qkv[:, 2 * D:].data_ptr(), o.data_ptr(), 3 * D,
                                       D, nseq, T, None, None, heads, hd, int(causal), hd ** -0.5, IMPL,
                                       torch.cuda.current_stream().cuda_stream)
            assert st == 0, st
        torch.cuda.synchronize()
        t = trace.view(3, 8, 32).cpu()
        t0 = int(t[t > 0].min())
        nb = int((t[1, 0] > 0).sum())
        print(f"== {name}: blocks traced {nb}; cycles relative to first event")
        names = {0: ["A:S(j+1) issue", "A:p_full seen", "A:PV issued", "-", "B:S(j+1) issue", "B:p_full seen", "B:PV issued", "-"],
                 1: ["wait s_full", "s_full seen", "S in regs", "max done", "pv_done+rescale", "exp+store+fence", "p_full arrived", "-"]}
        for j in range(min(nb, 8)):
            row = [f"j={j}"]
            for ev in range(7):
                v = int(t[1, ev, j])
                row.append(f"{names[1][ev]}={v - t0 if v else -1}")
            print("  softmaxA " + " | ".join(row))
            row = []
            for ev in range(7):
                v = int(t[2, ev, j])
                row.append(f"{v - t0 if v else -1}")
            print("  softmaxB " + " ".join(row))
            row = []
            for ev in (0, 1, 2, 4, 5, 6):
                v = int(t[0, ev, j])
                row.append(f"{names[0][ev]}={v - t0 if v else -1}")
            print("  mma      " + " | ".join(row))
        d = [int(t[1, 7, k]) for k in range(3)]
        if all(d):
            last_p = int(t[1, 6, nb - 1])
            print(f"  tile drain: last p_full -> drain start {d[0] - last_p}, wait last P.V (o_final) {d[1] - d[0]}, "
                  f"O -> bf16 -> global {d[2] - d[1]} cycles; first event -> first s_full seen {int(t[1, 1, 0]) - t0}")
        # per-iteration deltas for softmax A
        import statistics
        its = [int(t[1, 6, j] - t[1, 6, j - 1]) for j in range(1, nb) if t[1, 6, j] and t[1, 6, j - 1]]
        if its:
            print(f"  softmax A iteration period: median {statistics.median(its)} cycles, list {its[:12]}")
        seg = lambda a, b: [int(t[1, b, j] - t[1, a, j]) for j in range(1, nb)]
        print("  segments (median over j>=1): wait_s_full", statistics.median(seg(0, 1)), "ld", statistics.median(seg(1, 2)),
              "mask+max", statistics.median(seg(2, 3)), "wait_pv_done(+rescale)", statistics.median(seg(3, 4)),
              "exp+store+fence", statistics.median(seg(4, 5)), "arrive", statistics.median(seg(5, 6)))


if __name__ == "__main__":
    main()
