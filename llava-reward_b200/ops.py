"""Tensor-level wrappers over the C ABI (pointer extraction + shape checks only; no math happens here)."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("llava-reward-b200 kernels need CUDA tensors (there is no CPU fallback)")


def gemm(A: torch.Tensor, W: torch.Tensor, C: torch.Tensor, M: int, N: int, K: int, epilogue: int = L.EPI_NONE,
         bias: Optional[torch.Tensor] = None, R: Optional[torch.Tensor] = None, impl: int = L.GEMM_TCGEN05,
         lda: Optional[int] = None, ldw: Optional[int] = None, ldc: Optional[int] = None, ldr: Optional[int] = None):
    """C[M,N'] = epi(A[M,K] W[N,K]^T). A/W/C/R may be column-sliced views: leading dims default to stride(0).
    fp32 tensors (the verification path) run csrc/f32_verify.cu: plain fp32 GEMM, SwiGLU as a second kernel."""
    _need_cuda(A, W, C, bias, R)
    if A.dtype == torch.float32:
        st = _stream()
        if epilogue == L.EPI_SWIGLU:
            raw = torch.empty(M, N, dtype=torch.float32, device=A.device)
            L.call("lr_f32_gemm", _ptr(A), lda or A.stride(0), _ptr(W), ldw or W.stride(0), _ptr(raw), N, M, N, K,
                   L.EPI_NONE, None, None, 0, st)
            L.call("lr_f32_swiglu", _ptr(raw), N, _ptr(C), ldc or C.stride(0), M, N, st)
            return
        L.call("lr_f32_gemm", _ptr(A), lda or A.stride(0), _ptr(W), ldw or W.stride(0), _ptr(C), ldc or C.stride(0), M, N,
               K, epilogue, _ptr(bias), _ptr(R), (ldr or (R.stride(0) if R is not None else 0)), st)
        return
    L.call("lr_gemm_bf16", _ptr(A), lda or A.stride(0), _ptr(W), ldw or W.stride(0), _ptr(C), ldc or C.stride(0),
           M, N, K, epilogue, _ptr(bias), _ptr(R), (ldr or (R.stride(0) if R is not None else 0)), impl, _stream())


def gemm_rope(A, W, C, M, N, K, position_ids, cos_tab, sin_tab, rope_cols, head_dim, impl: int = L.GEMM_TCGEN05):
    """C = A W^T with su-RoPE fused on columns [0, rope_cols) (W's q/k rows head-interleaved, see weights.py)."""
    _need_cuda(A, W, C, position_ids, cos_tab, sin_tab)
    if A.dtype == torch.float32:   # verification path: GEMM, then the rotation in place on the same interleaved layout
        st = _stream()
        L.call("lr_f32_gemm", _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(C), C.stride(0), M, N, K, L.EPI_NONE, None,
               None, 0, st)
        L.call("lr_f32_rope", _ptr(C), C.stride(0), _ptr(position_ids), _ptr(cos_tab), _ptr(sin_tab), M, rope_cols,
               head_dim, st)
        return
    L.call("lr_gemm_rope_bf16", _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(C), C.stride(0), M, N, K,
           _ptr(position_ids), _ptr(cos_tab), _ptr(sin_tab), rope_cols, head_dim, impl, _stream())


def rmsnorm(x, w, y, rows, cols, eps, row_index=None):
    _need_cuda(x, w, y, row_index)
    if x.dtype == torch.float32:
        L.call("lr_f32_rmsnorm", _ptr(x), x.stride(0), _ptr(row_index), _ptr(w), _ptr(y), y.stride(0), rows, cols, eps,
               _stream())
        return
    L.call("lr_rmsnorm_bf16", _ptr(x), x.stride(0), _ptr(row_index), _ptr(w), _ptr(y), y.stride(0), rows, cols, eps,
           _stream())


def layernorm(x, w, b, y, rows, cols, eps):
    _need_cuda(x, w, b, y)
    if x.dtype == torch.float32:
        L.call("lr_f32_layernorm", _ptr(x), x.stride(0), _ptr(w), _ptr(b), _ptr(y), y.stride(0), rows, cols, eps, _stream())
        return
    L.call("lr_layernorm_bf16", _ptr(x), x.stride(0), _ptr(w), _ptr(b), _ptr(y), y.stride(0), rows, cols, eps, _stream())


def clip_im2col(pixels, crop_src, A, n_crops):
    _need_cuda(pixels, crop_src, A)
    if A.dtype == torch.float32:
        L.call("lr_f32_clip_im2col", _ptr(pixels), _ptr(crop_src), _ptr(A), n_crops, _stream())
        return
    L.call("lr_clip_im2col", _ptr(pixels), _ptr(crop_src), _ptr(A), n_crops, _stream())


def clip_embed_ln(patch, cls, pos, w, b, tokens, n_crops, eps):
    _need_cuda(patch, cls, pos, w, b, tokens)
    if tokens.dtype == torch.float32:
        L.call("lr_f32_clip_embed_ln", _ptr(patch), _ptr(cls), _ptr(pos), _ptr(w), _ptr(b), _ptr(tokens), n_crops, eps,
               _stream())
        return
    L.call("lr_clip_embed_ln", _ptr(patch), _ptr(cls), _ptr(pos), _ptr(w), _ptr(b), _ptr(tokens), n_crops, eps, _stream())


def attention(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, head_dim, causal, scale,
              impl: int = L.ATTN_TCGEN05):
    _need_cuda(q, k, v, o, seq_start, seq_len)
    if q.dtype == torch.float32:
        L.call("lr_f32_attention", _ptr(q), _ptr(k), _ptr(v), _ptr(o), ld_qkv, ld_o, n_seq, rows_per_seq, None,
               _ptr(seq_start), _ptr(seq_len), n_heads, n_heads, head_dim, int(causal), scale, _stream())
        return
    L.call("lr_attention_bf16", _ptr(q), _ptr(k), _ptr(v), _ptr(o), ld_qkv, ld_o, n_seq, rows_per_seq, _ptr(seq_start),
           _ptr(seq_len), n_heads, head_dim, int(causal), scale, impl, _stream())


def rope_su(qkv, position_ids, cos_tab, sin_tab, rows, n_heads, head_dim):
    _need_cuda(qkv, position_ids, cos_tab, sin_tab)
    L.call("lr_rope_su_bf16", _ptr(qkv), qkv.stride(0), _ptr(position_ids), _ptr(cos_tab), _ptr(sin_tab), rows, n_heads,
           head_dim, _stream())


def token_plan(ids, mask, B, S, pos, img_ord, seq_start, seq_len, eos_row, n_img, flags):
    _need_cuda(ids, mask, pos, img_ord, seq_start, seq_len, eos_row, n_img, flags)
    L.call("lr_token_plan", _ptr(ids), _ptr(mask), B, S, _ptr(pos), _ptr(img_ord), _ptr(seq_start), _ptr(seq_len),
           _ptr(eos_row), _ptr(n_img), _ptr(flags), _stream())


def token_plan_ex(ids, mask, B, S, image_token_id, position_mode, pos, img_ord, seq_start, seq_len, eos_row, n_img,
                  flags):
    _need_cuda(ids, mask, pos, img_ord, seq_start, seq_len, eos_row, n_img, flags)
    L.call("lr_token_plan_ex", _ptr(ids), _ptr(mask), B, S, int(image_token_id), int(position_mode), _ptr(pos),
           _ptr(img_ord), _ptr(seq_start), _ptr(seq_len), _ptr(eos_row), _ptr(n_img), _ptr(flags), _stream())


def anyres_embed_scatter(ids, img_ord, plan, wte, feat, newline, hidden, B, S, H, V):
    _need_cuda(ids, img_ord, plan, wte, feat, newline, hidden)
    L.call("lr_anyres_embed_scatter_bf16", _ptr(ids), _ptr(img_ord), _ptr(plan), _ptr(wte), _ptr(feat), feat.stride(0),
           _ptr(newline), _ptr(hidden), hidden.stride(0), B, S, H, V, _stream())


def hd_gather(clip_tokens, plan, sub_gn, glb_gn, rows, B, max_nv):
    _need_cuda(clip_tokens, plan, sub_gn, glb_gn, rows)
    if rows.dtype == torch.float32:   # the product's index kernel on 2048 two-byte units per CLIP token row
        L.call("lr_f32_hd_gather", _ptr(clip_tokens), _ptr(plan), _ptr(sub_gn), _ptr(glb_gn), _ptr(rows), B, max_nv,
               _stream())
        return
    L.call("lr_hd_gather_bf16", _ptr(clip_tokens), _ptr(plan), _ptr(sub_gn), _ptr(glb_gn), _ptr(rows), B, max_nv, _stream())


def embed_scatter(ids, img_ord, plan, wte, img_proj, hidden, B, S, H, V):
    _need_cuda(ids, img_ord, plan, wte, img_proj, hidden)
    if hidden.dtype == torch.float32:   # byte-wise: an fp32 row of H values is a row of 2H two-byte units
        L.call("lr_embed_scatter_bf16", _ptr(ids), _ptr(img_ord), _ptr(plan), _ptr(wte), _ptr(img_proj), _ptr(hidden),
               2 * hidden.stride(0), B, S, 2 * H, V, _stream())
        return
    L.call("lr_embed_scatter_bf16", _ptr(ids), _ptr(img_ord), _ptr(plan), _ptr(wte), _ptr(img_proj), _ptr(hidden),
           hidden.stride(0), B, S, H, V, _stream())


def skipca_scores(q, kv, plan, scores, B, H, max_nv):
    _need_cuda(q, kv, plan, scores)
    if q.dtype == torch.float32:
        L.call("lr_f32_skipca_scores", _ptr(q), q.stride(0), _ptr(kv), kv.stride(0), _ptr(plan), _ptr(scores), B, H,
               max_nv, _stream())
        return
    L.call("lr_skipca_scores", _ptr(q), q.stride(0), _ptr(kv), kv.stride(0), _ptr(plan), _ptr(scores), B, H, max_nv,
           _stream())


def skipca_head(scores, kv, plan, x, ca_ln_w, vh_w, reward, B, H, max_nv, vhd, eps):
    _need_cuda(scores, kv, plan, x, ca_ln_w, vh_w, reward)
    if x.dtype == torch.float32:
        L.call("lr_f32_skipca_head", _ptr(scores), _ptr(kv), (kv.stride(0) if kv is not None else 0), _ptr(plan), _ptr(x),
               x.stride(0), _ptr(ca_ln_w), _ptr(vh_w), _ptr(reward), B, H, max_nv, vhd, eps, _stream())
        return
    L.call("lr_skipca_head", _ptr(scores), _ptr(kv), (kv.stride(0) if kv is not None else 0), _ptr(plan), _ptr(x),
           x.stride(0), _ptr(ca_ln_w), _ptr(vh_w), _ptr(reward), B, H, max_nv, vhd, eps, _stream())


def preference(chosen, reject, prob, n, vhd, is_gpm, tau):
    _need_cuda(chosen, reject, prob)
    if chosen.dtype == torch.float32:
        L.call("lr_f32_preference", _ptr(chosen), _ptr(reject), _ptr(prob), n, vhd, int(is_gpm), float(tau), _stream())
        return
    L.call("lr_preference", _ptr(chosen), _ptr(reject), _ptr(prob), n, vhd, int(is_gpm), float(tau), _stream())


# ---- Qwen2.5-VL branch -------------------------------------------------------------------------------------------
def gemm_rope_ex(A, W, C, M, N, K, bias, position_ids, cos_tab, sin_tab, rope_cols, head_dim, epilogue):
    """C = A W^T + bias with the rotary embedding fused on columns [0, rope_cols); position_ids None = per-token tables."""
    _need_cuda(A, W, C, bias, position_ids, cos_tab, sin_tab)
    L.call("lr_gemm_rope_ex_bf16", _ptr(A), A.stride(0), _ptr(W), W.stride(0), _ptr(C), C.stride(0), M, N, K, _ptr(bias),
           _ptr(position_ids), _ptr(cos_tab), _ptr(sin_tab), rope_cols, head_dim, epilogue, _stream())


def attention_ex(q, k, v, o, ld_qkv, ld_o, total_rows, n_seq, max_len, seq_base, seq_start, seq_len, n_heads,
                 n_kv_heads, head_dim, causal, scale, impl: int = L.ATTN_TCGEN05):
    _need_cuda(q, k, v, o, seq_base, seq_start, seq_len)
    if q.dtype == torch.float32:
        L.call("lr_f32_attention", _ptr(q), _ptr(k), _ptr(v), _ptr(o), ld_qkv, ld_o, n_seq, max_len, _ptr(seq_base),
               _ptr(seq_start), _ptr(seq_len), n_heads, n_kv_heads, head_dim, int(causal), scale, _stream())
        return
    L.call("lr_attention_ex_bf16", _ptr(q), _ptr(k), _ptr(v), _ptr(o), ld_qkv, ld_o, total_rows, n_seq, max_len,
           _ptr(seq_base), _ptr(seq_start), _ptr(seq_len), n_heads, n_kv_heads, head_dim, int(causal), scale, impl,
           _stream())


def attention_seg(q, k, v, o, ld_qkv, ld_o, total_rows, row_lo, row_hi, n_heads, head_dim, scale):
    _need_cuda(q, k, v, o, row_lo, row_hi)
    L.call("lr_attention_seg_bf16", _ptr(q), _ptr(k), _ptr(v), _ptr(o), ld_qkv, ld_o, total_rows, _ptr(row_lo),
           _ptr(row_hi), n_heads, head_dim, scale, _stream())


def skipca_scores_ex(q, kv, plan, scores, B, H, max_nv, pad_score):
    _need_cuda(q, kv, plan, scores)
    L.call("lr_skipca_scores_ex", _ptr(q), q.stride(0), _ptr(kv), kv.stride(0), _ptr(plan), _ptr(scores), B, H, max_nv,
           float(pad_score), _stream())


def patch_rows(pixels, src_row, out, rows, K, Kpad):
    _need_cuda(pixels, src_row, out)
    L.call("lr_patch_rows_bf16", _ptr(pixels), _ptr(src_row), _ptr(out), out.stride(0), rows, K, Kpad, _stream())


def mrope_plan(ids, mask, B, S, image_token_id, grid_thw, n_images, merge, run_count, cos_tab, sin_tab, max_pos, half,
               sec0, sec1, pos3, cos_out, sin_out, flags):
    _need_cuda(ids, mask, grid_thw, run_count, cos_tab, sin_tab, pos3, cos_out, sin_out, flags)
    L.call("lr_mrope_plan", _ptr(ids), _ptr(mask), B, S, int(image_token_id), _ptr(grid_thw), n_images, merge,
           _ptr(run_count), _ptr(cos_tab), _ptr(sin_tab), max_pos, half, sec0, sec1, _ptr(pos3), _ptr(cos_out),
           _ptr(sin_out), _ptr(flags), _stream())


def compact_rows(src, ord_, plan, dst, B, S, cols):
    _need_cuda(src, ord_, plan, dst)
    L.call("lr_compact_rows_bf16", _ptr(src), src.stride(0), _ptr(ord_), _ptr(plan), _ptr(dst), dst.stride(0), B, S,
           cols, _stream())


def gather_rows(src, row_index, dst, rows, cols):
    _need_cuda(src, row_index, dst)
    if src.dtype == torch.float32:   # byte-wise on two-byte units (row_index -1 still yields a zero row)
        L.call("lr_gather_rows_bf16", _ptr(src), 2 * src.stride(0), _ptr(row_index), _ptr(dst), 2 * dst.stride(0), rows,
               2 * cols, _stream())
        return
    L.call("lr_gather_rows_bf16", _ptr(src), src.stride(0), _ptr(row_index), _ptr(dst), dst.stride(0), rows, cols,
           _stream())


# ---- all-rows head (mean_hidden_state) ------------------------------------------------------------------------------
def softmax_rows(scores, rows, n_valid, n_total, inv_sqrt_d):
    """In place: bf16 softmax over columns [0, n_valid) of bf16(score * inv_sqrt_d), zeros in [n_valid, n_total)."""
    _need_cuda(scores)
    L.call("lr_softmax_rows_bf16", _ptr(scores), scores.stride(0), rows, n_valid, n_total, float(inv_sqrt_d), _stream())


def masked_mean_rows(x, mask, out, B, S, H):
    _need_cuda(x, mask, out)
    L.call("lr_masked_mean_rows_bf16", _ptr(x), x.stride(0), _ptr(mask), _ptr(out), out.stride(0), B, S, H, _stream())
