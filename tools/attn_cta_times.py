"""Occupancy timeline of the tcgen05 attention kernel: every CTA records {SM, start, end (globaltimer ns), K/V blocks}
(private .so built with -DLR_ATTN_TRACE). Prints, per shape: duration vs block count (linear fit = time per block and
fixed cost per CTA), how long each SM had 0 / 1 / 2 CTAs resident, and the gap between a CTA's exit and the start of the
CTA that takes its place. usage (GPU box): python tools/attn_cta_times.py [--impl 4]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import attn_trace  # noqa: E402  (build() of the -DLR_ATTN_TRACE library)

IMPL = int(sys.argv[sys.argv.index("--impl") + 1]) if "--impl" in sys.argv else 4


def main():
    if not os.path.exists(attn_trace.OUT) or "--build" in sys.argv or "--build-only" in sys.argv:
        attn_trace.build()
    if "--build-only" in sys.argv:
        return
    lib = C.CDLL(attn_trace.OUT)
    p, i32, f32 = C.c_void_p, C.c_int, C.c_float
    lib.lr_attention_bf16.argtypes = [p, p, p, p, i32, i32, i32, i32, p, p, i32, i32, i32, f32, i32, p]
    lib.lr_attn_cta_set.argtypes = [p]
    bf = torch.bfloat16
    for name, (nseq, T, heads, hd, causal) in {"dec": (32, 2048, 32, 96, True), "clip": (416, 577, 16, 64, False)}.items():
        D = heads * hd
        qkv = torch.randn(nseq * T, 3 * D, device="cuda", dtype=bf)
        o = torch.empty(nseq * T, D, device="cuda", dtype=bf)
        tiles = (T + 127) // 128
        n_cta = tiles * heads * nseq
        rec = torch.zeros(n_cta * 4, dtype=torch.int64, device="cuda")
        assert lib.lr_attn_cta_set(rec.data_ptr()) == 0
        for _ in range(3):   # the last launch's records stay
            st = lib.lr_attention_bf16(qkv.data_ptr(), qkv[:, D:].data_ptr(), qkv[:, 2 * D:].data_ptr(), o.data_ptr(),
                                       3 * D, D, nseq, T, None, None, heads, hd, int(causal), hd ** -0.5, IMPL,
                                       torch.cuda.current_stream().cuda_stream)
            assert st == 0, st
        torch.cuda.synchronize()
        r = rec.view(n_cta, 4).cpu().numpy()
        r = r[r[:, 2] > 0]   # the multi-tile grid has fewer CTAs than tiles: keep the records that were written
        n_cta = len(r)       # (for a multi-tile CTA the block count is that of its LAST tile)
        sm, t0, t1, nb = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
        dur = (t1 - t0).astype(np.float64)
        span = float(t1.max() - t0.min())
        print(f"== {name}: {n_cta} CTAs, kernel span {span / 1e3:.1f} us, {len(np.unique(sm))} SMs")
        A = np.stack([nb.astype(np.float64), np.ones_like(dur)], 1)
        (per_blk, fixed), *_ = np.linalg.lstsq(A, dur, rcond=None)
        print(f"  CTA duration = {fixed:.0f} ns + {per_blk:.0f} ns per K/V block (least squares over all CTAs)")
        for k in sorted(set(nb.tolist())):
            d = dur[nb == k]
            print(f"    {k:2d} blocks: {len(d):6d} CTAs, mean {d.mean():8.0f} ns, p10 {np.percentile(d, 10):8.0f}, p90 {np.percentile(d, 90):8.0f}")
        # per-SM residency
        res = np.zeros(4)
        gaps = []
        for s_ in np.unique(sm):
            m = sm == s_
            ev = sorted([(t, +1) for t in t0[m]] + [(t, -1) for t in t1[m]])
            cur, last = 0, t0.min()
            for t, d_ in ev:
                res[min(cur, 3)] += t - last
                cur += d_
                last = t
            res[0] += t1.max() - last
            ends = np.sort(t1[m])
            starts = np.sort(t0[m])
            # a CTA that starts after >= 2 earlier CTAs of this SM have ended replaced one of them: gap to the most
            # recent exit before its start
            for st_ in starts[2:]:
                prev = ends[ends <= st_]
                if len(prev):
                    gaps.append(st_ - prev[-1])
        tot = res.sum()
        print(f"  SM time with 0 / 1 / 2 CTAs resident: {res[0] / tot:.1%} / {res[1] / tot:.1%} / {res[2] / tot:.1%}")
        g = np.asarray(gaps, dtype=np.float64)
        print(f"  exit -> next CTA start on the same SM: median {np.median(g):.0f} ns, mean {g.mean():.0f}, p90 {np.percentile(g, 90):.0f}")
        print(f"  sum of CTA durations / (SMs x span x 2 slots) = {dur.sum() / (len(np.unique(sm)) * span * 2):.1%}")


if __name__ == "__main__":
    main()
