"""GPU-box diagnostics: structured tcgen05 GEMM error maps + quick kernel timings (not a benchmark).
usage: python tools/gpu_diag.py [gemm] [perf]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402

bf = torch.bfloat16
out = {}


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def gemm_diag():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (128, 128, 64), (256, 512, 256), (1000, 1024, 640)]:
        A = torch.randn(M, K, device="cuda").to(bf)
        W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(bf)
        C = torch.full((M, N), float("nan"), dtype=bf, device="cuda")
        ops.gemm(A, W, C, M, N, K)
        torch.cuda.synchronize()
        ref = A.float() @ W.float().t()
        err = (C.float() - ref).abs()
        nan = torch.isnan(C.float()).sum().item()
        print(f"gemm {M}x{N}x{K}: max err {err.nan_to_num(9e9).max().item():.4g}, nan {nan}, "
              f"ref absmax {ref.abs().max().item():.3g}")
        if err.nan_to_num(9e9).max().item() > 0.05:
            blk = err.nan_to_num(9e9)[: (M // 32) * 32, : (N // 32) * 32].view(M // 32, 32, N // 32, 32).amax((1, 3))
            print("  32x32 block max err:\n", blk[:8, :8])
            # structured probe: which (row, col) does each output come from?
            A2 = torch.zeros(M, K, device="cuda", dtype=bf)
            A2[:, 0] = torch.arange(M, device="cuda").to(bf)           # row id in column 0
            W2 = torch.zeros(N, K, device="cuda", dtype=bf)
            W2[:, 0] = 1.0
            C2 = torch.full((M, N), float("nan"), dtype=bf, device="cuda")
            ops.gemm(A2, W2, C2, M, N, K)
            torch.cuda.synchronize()
            print("  row-id probe C[:16,0]:", C2[:16, 0].float().tolist())
            print("  row-id probe C[0,:16]:", C2[0, :16].float().tolist())
            W3 = torch.zeros(N, K, device="cuda", dtype=bf)
            W3[:, 0] = (torch.arange(N, device="cuda") % 128).to(bf)
            A3 = torch.zeros(M, K, device="cuda", dtype=bf)
            A3[:, 0] = 1.0
            ops.gemm(A3, W3, C2, M, N, K)
            torch.cuda.synchronize()
            print("  col-id probe C[0,:24]:", C2[0, :24].float().tolist())
            # k probe: A has ones only at column kk, W[n, kk] = kk -> output = kk if k-slices line up
            for kk in (1, 8, 16, 17, 32, 63):
                if kk >= K:
                    continue
                A4 = torch.zeros(M, K, device="cuda", dtype=bf)
                A4[:, kk] = 1.0
                W4 = torch.arange(K, device="cuda").to(bf)[None].expand(N, K).contiguous()
                ops.gemm(A4, W4, C2, M, N, K)
                torch.cuda.synchronize()
                print(f"  k probe kk={kk}: C[0,0]={C2[0,0].item()} C[5,9]={C2[5,9].item()}")
            break


def perf(gemms=True):
    res = {}
    for name, (M, N, K, epi) in ({} if not gemms else {
        "qkv": (65536, 9216, 3200, L.EPI_NONE), "o": (65536, 3072, 3200, L.EPI_RESIDUAL),
        "gate_up": (65536, 16384, 3200, L.EPI_SWIGLU), "down": (65536, 3072, 8320, L.EPI_RESIDUAL),
        "lora_a": (65536, 128, 3072, L.EPI_NONE), "clip_qkv": (240032, 3072, 1024, L.EPI_BIAS),
        "clip_fc1": (240032, 4096, 1024, L.EPI_BIAS_QUICKGELU), "clip_fc2": (240032, 1024, 4096, L.EPI_BIAS_RESIDUAL),
        "clip_out": (240032, 1024, 1024, L.EPI_BIAS_RESIDUAL),
    }).items():
        A = torch.randn(M, K, device="cuda", dtype=bf)
        W = torch.randn(N, K, device="cuda", dtype=bf) * 0.02
        n_out = N // 2 if epi == L.EPI_SWIGLU else N
        C = torch.empty(M, n_out, device="cuda", dtype=bf)
        bias = torch.zeros(N, device="cuda", dtype=bf)
        R = torch.zeros(M, n_out, device="cuda", dtype=bf) if epi in (L.EPI_RESIDUAL, L.EPI_BIAS_RESIDUAL) else None
        ms = timeit(lambda: ops.gemm(A, W, C, M, N, K, epi, bias, R, impl=L.GEMM_TCGEN05_SINGLE))
        tf = 2.0 * M * N * K / ms / 1e9
        if N % 256 == 0 and os.environ.get("LR_DIAG_PAIR", "1") == "1":
            ms2 = timeit(lambda: ops.gemm(A, W, C, M, N, K, epi, bias, R, impl=L.GEMM_TCGEN05_PAIR))
            print(f"gemm {name:9s} pair kernel: {ms2:.3f} ms {2.0 * M * N * K / ms2 / 1e9:.0f} TF/s", flush=True)
        ms_t = timeit(lambda: torch.matmul(A, W.t()))
        res[name] = {"ms": ms, "tflops": tf, "torch_ms": ms_t, "torch_tflops": 2.0 * M * N * K / ms_t / 1e9}
        print(f"gemm {name:9s} {M}x{N}x{K}: {ms:.3f} ms {tf:.0f} TF/s | torch.matmul {ms_t:.3f} ms "
              f"{2.0 * M * N * K / ms_t / 1e9:.0f} TF/s", flush=True)
        del A, W, C, R
    # attention
    for name, (nseq, T, heads, hd, causal) in {"clip": (416, 577, 16, 64, False), "dec": (32, 2048, 32, 96, True)}.items():
        D = heads * hd
        qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
        o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
        fl = 4.0 * nseq * heads * T * T * hd * (0.5 if causal else 1.0)
        for impl, iname in ((L.ATTN_TCGEN05, "tcgen05"), (L.ATTN_TCGEN05_SPLIT, "tcgen05-split"), (L.ATTN_TCGEN05_2TILE, "tcgen05-2tile"), (L.ATTN_MMA_SYNC, "mma.sync")):
            ms = timeit(lambda: ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, None, None, heads,
                                              hd, causal, hd ** -0.5, impl))
            res[f"attn_{name}_{iname}"] = {"ms": ms, "tflops": fl / ms / 1e9}
            print(f"attention {name} {iname}: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s", flush=True)
    # norms
    x = torch.randn(65536, 3072, device="cuda", dtype=bf)
    w = torch.ones(3072, device="cuda", dtype=bf)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.rmsnorm(x, w, y, 65536, 3072, 1e-5))
    res["rmsnorm"] = {"ms": ms, "gbs": 2 * x.numel() * 2 / ms / 1e6}
    print(f"rmsnorm 65536x3072: {ms:.3f} ms {2 * x.numel() * 2 / ms / 1e6:.0f} GB/s")
    out["perf"] = res


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "perf"]
    if "gemm" in what:
        gemm_diag()
    if "perf" in what:
        perf()
    if "attn" in what:
        perf(gemms=False)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(out, f, indent=1)


def accuracy():
    """relative L2 error of kernels against float64 references (is any kernel systematically less accurate?)"""
    import torch.nn.functional as F
    torch.manual_seed(1)
    for name, (nseq, T, heads, hd, causal) in {"clip": (2, 577, 16, 64, False), "dec": (2, 2048, 8, 96, True)}.items():
        D = heads * hd
        qkv = (torch.randn(nseq * T, 3 * D, device="cuda") * 1.0).to(bf)
        f = qkv.double().view(nseq, T, 3, heads, hd)
        q, k, v = (f[:, :, i].transpose(1, 2) for i in range(3))
        s = q @ k.transpose(-1, -2) * hd ** -0.5
        if causal:
            s = s.masked_fill(~torch.ones(T, T, dtype=torch.bool, device="cuda").tril(), float("-inf"))
        ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(nseq * T, D)
        for impl, iname in ((L.ATTN_TCGEN05, "tcgen05"), (L.ATTN_TCGEN05_SPLIT, "tcgen05-split"), (L.ATTN_MMA_SYNC, "mma.sync")):
            o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, None, None, heads, hd, causal, hd ** -0.5, impl)
            torch.cuda.synchronize()
            err = (o.double() - ref).norm() / ref.norm()
            bias = ((o.double() - ref).mean() / ref.abs().mean()).item()
            print(f"attention {name} {iname}: rel L2 err {err.item():.3e}, mean signed err/|ref| {bias:.2e}")
        # torch SDPA bf16 for comparison
        o2 = F.scaled_dot_product_attention(q.to(bf), k.to(bf), v.to(bf), is_causal=causal).transpose(1, 2).reshape(nseq * T, D)
        print(f"attention {name} torch sdpa bf16: rel L2 err {((o2.double() - ref).norm() / ref.norm()).item():.3e}")
    M, N, K = 4096, 1024, 3200
    A = torch.randn(M, K, device="cuda").to(bf)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(bf)
    acc = A.double() @ W.double().t()
    a = acc.view(M, N // 256, 2, 128)
    ref = (a[:, :, 1] * F.silu(a[:, :, 0])).reshape(M, N // 2)
    for impl, iname in ((L.GEMM_TCGEN05, "tcgen05"), (L.GEMM_SIMT, "simt")):
        C = torch.empty(M, N // 2, device="cuda", dtype=bf)
        ops.gemm(A, W, C, M, N, K, L.EPI_SWIGLU, impl=impl)
        torch.cuda.synchronize()
        print(f"gemm swiglu {iname}: rel L2 err {((C.double() - ref).norm() / ref.norm()).item():.3e}")
    gate, up = (a[:, :, 0].float().to(bf), a[:, :, 1].float().to(bf))
    t = (up * F.silu(gate)).reshape(M, N // 2)
    print(f"torch bf16 silu*up: rel L2 err {((t.double() - ref).norm() / ref.norm()).item():.3e}")
    bias = torch.randn(N, device="cuda").to(bf)
    x = (acc + bias.double())
    ref = x * torch.sigmoid(1.702 * x)
    C = torch.empty(M, N, device="cuda", dtype=bf)
    ops.gemm(A, W, C, M, N, K, L.EPI_BIAS_QUICKGELU, bias)
    torch.cuda.synchronize()
    print(f"gemm quickgelu tcgen05: rel L2 err {((C.double() - ref).norm() / ref.norm()).item():.3e}")
    xb = x.float().to(bf)
    t = xb * torch.sigmoid(1.702 * xb)
    print(f"torch bf16 quickgelu: rel L2 err {((t.double() - ref).norm() / ref.norm()).item():.3e}")


if __name__ == "__main__" and "acc" in sys.argv[1:]:
    accuracy()
