"""Per-kernel parity on the B200: every C-ABI entry against a plain PyTorch fp32 statement of the same op
(tolerances = bf16 output rounding; integer/index outputs bit-exact)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from llava_reward_b200 import _lib as L  # noqa: E402
from llava_reward_b200 import ops  # noqa: E402

DEV = "cuda"
bf = torch.bfloat16


def rnd(*shape, std=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * std).to(bf).to(DEV)


def bfr(x):
    return x.to(bf).float()


def gemm_ref(A, W, epi, bias=None, R=None):
    acc = A.float() @ W.float().t()
    if epi == L.EPI_NONE:
        return acc
    if epi == L.EPI_BIAS:
        return acc + bias.float()
    if epi == L.EPI_BIAS_QUICKGELU:
        x = bfr(acc + bias.float())
        return x * torch.sigmoid(1.702 * x)
    if epi == L.EPI_BIAS_GELU:
        return F.gelu(bfr(acc + bias.float()))
    if epi == L.EPI_RESIDUAL:
        return bfr(acc) + R.float()
    if epi == L.EPI_BIAS_RESIDUAL:
        return bfr(acc + bias.float()) + R.float()
    if epi == L.EPI_SWIGLU:
        N = W.shape[0]
        a = acc.view(acc.shape[0], N // 256, 2, 128)
        gate, up = bfr(a[:, :, 0]), bfr(a[:, :, 1])
        return (up * bfr(F.silu(gate))).reshape(acc.shape[0], N // 2)
    raise ValueError(epi)


def check_close(out, ref, what, atol=2e-2, rtol=2e-2):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} elements off, max err {err.max().item():.4g} " \
                     f"(ref absmax {ref.abs().max().item():.4g})"


GEMM_SHAPES = [
    # M, N, K
    (128, 256, 64), (128, 256, 256), (300, 256, 128), (1000, 1024, 640), (2, 3072, 3072), (2500, 128, 3072),
    (4096, 9216, 3200), (577 * 3, 3072, 1024),
]


@pytest.mark.parametrize("impl", [L.GEMM_TCGEN05, L.GEMM_SIMT, L.GEMM_TCGEN05_PAIR, L.GEMM_TCGEN05_SINGLE])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(M, N, K, impl):
    if impl == L.GEMM_SIMT and M * N * K > 2e10:
        pytest.skip("SIMT cross-check kernel only on small shapes")
    if impl == L.GEMM_TCGEN05_PAIR and N % 256:
        pytest.skip("CTA-pair kernel needs N % 256 == 0")
    A, W = rnd(M, K, seed=1), rnd(N, K, std=K ** -0.5, seed=2)
    C = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    ops.gemm(A, W, C, M, N, K, L.EPI_NONE, impl=impl)
    torch.cuda.synchronize()
    check_close(C, gemm_ref(A, W, L.EPI_NONE), f"gemm {M}x{N}x{K} impl={impl}")


@pytest.mark.parametrize("impl", [L.GEMM_TCGEN05, L.GEMM_SIMT, L.GEMM_TCGEN05_PAIR, L.GEMM_TCGEN05_SINGLE])
@pytest.mark.parametrize("epi", [L.EPI_BIAS, L.EPI_BIAS_QUICKGELU, L.EPI_BIAS_GELU, L.EPI_RESIDUAL,
                                 L.EPI_BIAS_RESIDUAL, L.EPI_SWIGLU])
def test_gemm_epilogues(epi, impl):
    M, N, K = 777, 512, 320
    A, W = rnd(M, K, seed=3), rnd(N, K, std=K ** -0.5, seed=4)
    bias, R = rnd(N, seed=5), rnd(M, N, seed=6)
    n_out = N // 2 if epi == L.EPI_SWIGLU else N
    C = torch.full((M, n_out), float("nan"), dtype=bf, device=DEV)
    ops.gemm(A, W, C, M, N, K, epi, bias, R if epi in (L.EPI_RESIDUAL, L.EPI_BIAS_RESIDUAL) else None, impl=impl)
    torch.cuda.synchronize()
    check_close(C, gemm_ref(A, W, epi, bias, R), f"epilogue {epi} impl={impl}")


def test_gemm_strided_lora_extension():
    """LoRA K-extension: t = x A^T written into columns [K, K+r) of the same buffer, then y = [x|t] [W|2B]^T."""
    M, K, r, N = 1500, 3072, 128, 1024
    x = rnd(M, K, seed=7)
    Aw, Bw, W = rnd(r, K, std=0.02, seed=8), rnd(N, r, std=0.02, seed=9), rnd(N, K, std=0.02, seed=10)
    xe = torch.zeros(M, K + r, dtype=bf, device=DEV)
    xe[:, :K] = x
    ops.gemm(xe, Aw, xe[:, K:], M, r, K)
    Wext = torch.cat([W, (Bw.float() * 2).to(bf)], 1).contiguous()
    y = torch.empty(M, N, dtype=bf, device=DEV)
    ops.gemm(xe, Wext, y, M, N, K + r)
    torch.cuda.synchronize()
    t = bfr(x.float() @ Aw.float().t())
    check_close(xe[:, K:], t, "lora_A", atol=1e-2)
    ref = x.float() @ W.float().t() + 2 * (t @ Bw.float().t())
    check_close(y, ref, "lora ext", atol=2e-2)


@pytest.mark.parametrize("impl", [L.GEMM_TCGEN05, L.GEMM_SIMT])
def test_gemm_rope_fused_equals_gemm_then_rope(impl):
    """qkv projection with RoPE in the epilogue (head-interleaved q/k rows) == plain GEMM followed by lr_rope_su_bf16,
    bit for bit after undoing the interleave (same accumulations, same op-by-op bf16 rounding)."""
    M, heads, hd, K = 700, 8, 96, 3200
    H = heads * hd
    N = 3 * H  # 2304 = 9 * 256
    x = rnd(M, K, seed=21)
    W = rnd(N, K, std=K ** -0.5, seed=22)
    pos = torch.randint(0, 2048, (M,), device=DEV, dtype=torch.int32)
    ang = torch.rand(2048, hd // 2, device=DEV) * 6.0
    cos, sin = (ang.cos() * 1.19).to(bf).contiguous(), (ang.sin() * 1.19).to(bf).contiguous()
    ref = torch.empty(M, N, dtype=bf, device=DEV)
    ops.gemm(x, W, ref, M, N, K, impl=L.GEMM_TCGEN05_PAIR if impl == L.GEMM_TCGEN05 else impl)
    ops.rope_su(ref, pos, cos, sin, M, heads, hd)
    inter = torch.stack([torch.arange(hd // 2), torch.arange(hd // 2) + hd // 2], dim=1).reshape(-1)
    qk_perm = (torch.arange(2 * heads)[:, None] * hd + inter[None, :]).reshape(-1)
    perm = torch.cat([qk_perm, torch.arange(2 * H, 3 * H)]).to(DEV)
    out = torch.full((M, N), float("nan"), dtype=bf, device=DEV)
    ops.gemm_rope(x, W[perm].contiguous(), out, M, N, K, pos, cos, sin, 2 * H, hd, impl)
    torch.cuda.synchronize()
    unperm = torch.empty_like(out)
    unperm[:, perm] = out
    assert torch.equal(unperm, ref)


def test_gemm_inplace_residual():
    M, N, K = 2048, 1024, 4096
    A, W, b = rnd(M, K, seed=11), rnd(N, K, std=K ** -0.5, seed=12), rnd(N, seed=13)
    X = rnd(M, N, seed=14)
    ref = gemm_ref(A, W, L.EPI_BIAS_RESIDUAL, b, X.clone())
    ops.gemm(A, W, X, M, N, K, L.EPI_BIAS_RESIDUAL, b, X)
    torch.cuda.synchronize()
    check_close(X, ref, "in-place residual")


def test_gemm_bad_args():
    A, W, C = rnd(128, 64), rnd(256, 64), rnd(128, 256)
    with pytest.raises(RuntimeError, match="BAD_ARG"):
        ops.gemm(A, W, C, 128, 256, 63)
    with pytest.raises(RuntimeError, match="BAD_ARG"):
        ops.gemm(A, W, C, 128, 256, 64, L.EPI_BIAS, None)


@pytest.mark.parametrize("rows,cols", [(1, 3072), (1000, 3072), (333, 1024), (17, 8192), (5003, 1280), (2050, 640),
                                       (4100, 2048), (1024, 512)])
def test_rmsnorm(rows, cols):
    x, w = rnd(rows, cols, std=2.0, seed=1), (1 + rnd(cols, std=0.1, seed=2).float()).to(bf)
    y = torch.empty_like(x)
    ops.rmsnorm(x, w, y, rows, cols, 1e-5)
    xf = x.float()
    ref = w.float() * bfr(xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    check_close(y, ref, "rmsnorm", atol=1e-2, rtol=1e-2)
    idx = torch.randperm(rows, device=DEV)[: max(1, rows // 2)].to(torch.int32)
    y2 = torch.empty(idx.numel(), cols, dtype=bf, device=DEV)
    ops.rmsnorm(x, w, y2, idx.numel(), cols, 1e-5, row_index=idx)
    check_close(y2, ref[idx.long()], "rmsnorm gather", atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("rows,cols", [(577 * 2, 1024), (5, 4096)])
def test_layernorm(rows, cols):
    x = rnd(rows, cols, std=3.0, seed=3) + 0.5
    w, b = (1 + rnd(cols, std=0.1, seed=4).float()).to(bf), rnd(cols, std=0.1, seed=5)
    y = torch.empty_like(x)
    ops.layernorm(x, w, b, y, rows, cols, 1e-5)
    ref = F.layer_norm(x.float(), (cols,), w.float(), b.float(), 1e-5)
    check_close(y, ref, "layernorm", atol=1e-2, rtol=1e-2)


def test_clip_front():
    n_slots, crops = 5, [3, 0, 4]
    g = torch.Generator().manual_seed(0)
    pix = torch.randn(n_slots, 3, 336, 336, generator=g).to(DEV)
    crop_src = torch.tensor(crops, dtype=torch.int32, device=DEV)
    A = torch.empty(len(crops) * 576, 640, dtype=bf, device=DEV)
    ops.clip_im2col(pix, crop_src, A, len(crops))
    ref = F.unfold(pix[crop_src.long()].to(bf).float(), kernel_size=14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(A[:, :588].float(), ref)
    assert (A[:, 588:] == 0).all()
    patch, cls, pos = rnd(len(crops) * 576, 1024, seed=1), rnd(1024, seed=2), rnd(577, 1024, seed=3)
    w, b = (1 + rnd(1024, std=0.1, seed=4).float()).to(bf), rnd(1024, std=0.1, seed=5)
    tok = torch.empty(len(crops) * 577, 1024, dtype=bf, device=DEV)
    ops.clip_embed_ln(patch, cls, pos, w, b, tok, len(crops), 1e-5)
    e = torch.cat([cls.expand(len(crops), 1, 1024), patch.view(len(crops), 576, 1024)], 1) + pos[None]
    ref = F.layer_norm(e.float(), (1024,), w.float(), b.float(), 1e-5).reshape(-1, 1024)
    check_close(tok, ref, "clip_embed_ln", atol=1e-2, rtol=1e-2)


def attn_ref(q, k, v, causal, scale, start=0, length=None):
    """q,k,v [T, heads, hd] fp32; valid rows [start, start+length)."""
    T = q.shape[0]
    length = T if length is None else length
    out = torch.zeros_like(q)
    sl = slice(start, start + length)
    qq, kk, vv = (t[sl].transpose(0, 1) for t in (q, k, v))
    s = qq @ kk.transpose(1, 2) * scale
    if causal:
        s = s.masked_fill(~torch.ones(length, length, dtype=torch.bool, device=q.device).tril(), float("-inf"))
    out[sl] = (torch.softmax(s, -1) @ vv).transpose(0, 1)
    return out


@pytest.mark.parametrize("impl", [L.ATTN_TCGEN05, L.ATTN_MMA_SYNC, L.ATTN_TCGEN05_SPLIT, L.ATTN_TCGEN05_2TILE,
                                  L.ATTN_TCGEN05_1TILE, L.ATTN_TCGEN05_MULTITILE])
@pytest.mark.parametrize("hd,heads,T,nseq,causal", [(64, 16, 577, 3, False), (96, 32, 700, 2, True),
                                                    (96, 4, 130, 3, True), (64, 2, 64, 1, False),
                                                    (96, 2, 2048, 3, True), (64, 3, 1000, 2, False)])
def test_attention(hd, heads, T, nseq, causal, impl):
    D = heads * hd
    qkv = rnd(nseq * T, 3 * D, seed=1)
    o = torch.full((nseq * T, D + 128), float("nan"), dtype=bf, device=DEV)
    if causal:
        lens = [T, T - 37, 5][:nseq] if T < 2000 else [T, T - 87, 1300][:nseq]
        starts = [T - n for n in lens]
        ss = torch.tensor(starts, dtype=torch.int32, device=DEV)
        sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    else:
        lens, starts, ss, sl = [T] * nseq, [0] * nseq, None, None
    scale = hd ** -0.5
    ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D + 128, nseq, T, ss, sl, heads, hd, causal, scale, impl)
    torch.cuda.synchronize()
    f = qkv.float().view(nseq, T, 3, heads, hd)
    for s in range(nseq):
        ref = attn_ref(f[s, :, 0], f[s, :, 1], f[s, :, 2], causal, scale, starts[s], lens[s]).reshape(T, D)
        check_close(o[s * T:(s + 1) * T, :D], ref, f"attention seq {s}", atol=1e-2, rtol=2e-2)


def test_rope():
    rows, heads, hd = 300, 32, 96
    qkv = rnd(rows, 3 * heads * hd, seed=1)
    pos = torch.randint(0, 500, (rows,), device=DEV, dtype=torch.int32)
    ang = torch.rand(500, hd // 2, device=DEV) * 6.0
    cos, sin = (ang.cos() * 1.19).to(bf), (ang.sin() * 1.19).to(bf)
    ref = qkv.clone()
    x = qkv[:, : 2 * heads * hd].view(rows, 2 * heads, hd)
    c = torch.cat([cos, cos], -1)[pos.long()][:, None]
    s = torch.cat([sin, sin], -1)[pos.long()][:, None]
    rot = torch.cat([-x[..., hd // 2:], x[..., : hd // 2]], -1)
    ref[:, : 2 * heads * hd] = ((x * c) + (rot * s)).reshape(rows, -1)  # bf16 op by op, like the reference
    ops.rope_su(qkv, pos, cos, sin, rows, heads, hd)
    assert torch.equal(qkv, ref)


def test_token_plan():
    B, S = 5, 700
    ids = torch.randint(3, 32000, (B, S), dtype=torch.int64)
    mask = torch.ones(B, S, dtype=torch.int64)
    pads = [0, 13, 300, 699, 256]
    for b, p in enumerate(pads):
        mask[b, :p] = 0
        n = min(S - p - 2, 100 + 31 * b)
        if n > 0:
            ids[b, p + 1: p + 1 + n] = -1
    ids, mask = ids.to(DEV), mask.to(DEV)
    i32 = dict(dtype=torch.int32, device=DEV)
    pos, ordn = torch.empty(B * S, **i32), torch.empty(B * S, **i32)
    ss, sl, er, ni = (torch.zeros(B, **i32) for _ in range(4))
    fl = torch.zeros(1, **i32)
    ops.token_plan(ids, mask, B, S, pos, ordn, ss, sl, er, ni, fl)
    ref_pos = (mask.cumsum(-1) - 1).masked_fill(mask == 0, 1)
    assert torch.equal(pos.view(B, S).long(), ref_pos)
    is_img = ids < 0
    ref_ord = torch.where(is_img, is_img.long().cumsum(-1) - 1, torch.full_like(ids, -1))
    assert torch.equal(ordn.view(B, S).long(), ref_ord)
    assert ss.tolist() == pads and sl.tolist() == [S - p for p in pads]
    eos = S - 1 - mask.flip(1).argmax(1)
    assert er.tolist() == [b * S + int(eos[b]) for b in range(B)]
    assert ni.tolist() == is_img.sum(1).tolist() and fl.item() == 0
    mask[2, 500] = 0
    ops.token_plan(ids, mask, B, S, pos, ordn, ss, sl, er, ni, fl)
    assert fl.item() == 1


def test_preference():
    n = 1000
    c, r = rnd(n, 2, seed=1), rnd(n, 2, seed=2)
    prob = torch.empty(n, device=DEV)
    ops.preference(c, r, prob, n, 2, True, 0.1)
    ref = torch.sigmoid((c[:, 0] * r[:, 1] - c[:, 1] * r[:, 0]) / 0.1).float()
    assert (prob - ref).abs().max().item() <= 4e-3
    c1, r1 = rnd(n, 1, seed=3), rnd(n, 1, seed=4)
    ops.preference(c1, r1, prob, n, 1, False, 0.1)
    ref = torch.sigmoid((c1 - r1) / 0.1).squeeze(-1).float()
    assert (prob - ref).abs().max().item() <= 4e-3


@pytest.mark.parametrize("hd,heads,T,nseq,causal", [(64, 16, 577, 5, False), (96, 8, 2048, 3, True), (96, 4, 900, 2, True),
                                                    (64, 4, 1400, 2, False)])
def test_attention_multitile_is_bit_identical_to_one_tile_per_cta(hd, heads, T, nseq, causal):
    """LR_ATTN_TCGEN05_MULTITILE walks several query tiles per CTA with the SAME per-tile pipeline (barrier phases carried
    across tiles): every output element must equal the one-tile-per-CTA kernel's bit for bit, incl. ragged valid runs."""
    D = heads * hd
    qkv = rnd(nseq * T, 3 * D, seed=3)
    if causal:
        lens = [T, T - 87, 300][:nseq]
        ss = torch.tensor([T - n for n in lens], dtype=torch.int32, device=DEV)
        sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    else:
        ss, sl = None, None
    outs = []
    for impl in (L.ATTN_TCGEN05_1TILE, L.ATTN_TCGEN05_MULTITILE):
        o = torch.full((nseq * T, D), float("nan"), dtype=bf, device=DEV)
        for _ in range(2):   # twice: a second launch must not depend on leftovers of the first
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, ss, sl, heads, hd, causal, hd ** -0.5, impl)
        torch.cuda.synchronize()
        outs.append(o)
    assert not torch.isnan(outs[1].float()).any()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("hd,causal", [(96, True), (64, False), (128, True), (96, False)])
@pytest.mark.parametrize("left_pad", [True, False])
def test_attention_mask_boundaries(hd, causal, left_pad):
    """Valid lengths on every side of the 32-column chunk and 128-row tile boundaries (the chunk-level masking of the
    tcgen05 kernel classifies chunks from the warp-wide min / max of the valid prefix): slot layout, left- or
    right-aligned valid runs, causal and non-causal, against an fp32 reference per sequence."""
    heads, T = 2, 300
    lens = [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 159, 160, 161, 255, 256, 257, 289, 300]
    nseq = len(lens)
    D = heads * hd
    qkv = rnd(nseq * T, 3 * D, seed=11)
    starts = [T - n if left_pad else 0 for n in lens]
    ss = torch.tensor(starts, dtype=torch.int32, device=DEV)
    sl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    o = torch.full((nseq * T, D), float("nan"), dtype=bf, device=DEV)
    scale = hd ** -0.5
    ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, 3 * D, D, nseq, T, ss, sl, heads, hd, causal, scale, L.ATTN_TCGEN05)
    torch.cuda.synchronize()
    f = qkv.float().view(nseq, T, 3, heads, hd)
    for s in range(nseq):
        ref = attn_ref(f[s, :, 0], f[s, :, 1], f[s, :, 2], causal, scale, starts[s], lens[s]).reshape(T, D)
        got = o[s * T:(s + 1) * T]
        assert not torch.isnan(got.float()).any(), f"len {lens[s]}: rows left unwritten"
        check_close(got, ref, f"attention len {lens[s]} start {starts[s]}", atol=1e-2, rtol=2e-2)
