"""ctypes binding of libllavareward.so (C ABI declared in include/llava_reward_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry
fails loudly, and every entry returns an error code on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libllavareward.so")

LR_OK = 0
EPI_NONE, EPI_BIAS, EPI_BIAS_QUICKGELU, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_BIAS_RESIDUAL, EPI_SWIGLU = range(7)
EPI_BIAS_SWIGLU, EPI_BIAS_ROPE, EPI_BIAS_ROPE_F32 = 8, 9, 10
GEMM_TCGEN05, GEMM_SIMT, GEMM_TCGEN05_PAIR, GEMM_TCGEN05_SINGLE = 0, 1, 2, 3
ATTN_TCGEN05, ATTN_MMA_SYNC, ATTN_TCGEN05_SPLIT, ATTN_TCGEN05_2TILE, ATTN_TCGEN05_1TILE = 0, 1, 2, 3, 4
ATTN_TCGEN05_MULTITILE = 5
PLAN_STRIDE = 8
PLAN_HCROP, PLAN_WCROP, PLAN_CROP_BASE, PLAN_ROW_BASE, PLAN_NV, PLAN_TOP, PLAN_LEFT = 0, 1, 2, 3, 4, 5, 6
POS_FROM_MASK, POS_ARANGE = 0, 1

_ERR = {-1: "LR_ERR_BAD_ARG", -2: "LR_ERR_ALIGN", -3: "LR_ERR_NO_DRIVER", -4: "LR_ERR_UNSUPPORTED"}

p, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

SIGNATURES = {
    "lr_version": ([], i32),
    "lr_device_check": ([], i32),
    "lr_gemm_bf16": ([p, i32, p, i32, p, i32, i32, i32, i32, i32, p, p, i32, i32, p], i32),
    "lr_gemm_rope_bf16": ([p, i32, p, i32, p, i32, i32, i32, i32, p, p, p, i32, i32, i32, p], i32),
    "lr_rmsnorm_bf16": ([p, i32, p, p, p, i32, i32, i32, f32, p], i32),
    "lr_layernorm_bf16": ([p, i32, p, p, p, i32, i32, i32, f32, p], i32),
    "lr_clip_im2col": ([p, p, p, i32, p], i32),
    "lr_clip_embed_ln": ([p, p, p, p, p, p, i32, f32, p], i32),
    "lr_attention_bf16": ([p, p, p, p, i32, i32, i32, i32, p, p, i32, i32, i32, f32, i32, p], i32),
    "lr_rope_su_bf16": ([p, i32, p, p, p, i32, i32, i32, p], i32),
    "lr_token_plan": ([p, p, i32, i32, p, p, p, p, p, p, p, p], i32),
    "lr_token_plan_ex": ([p, p, i32, i32, i64, i32, p, p, p, p, p, p, p, p], i32),
    "lr_anyres_embed_scatter_bf16": ([p, p, p, p, p, i32, p, p, i32, i32, i32, i32, i32, p], i32),
    "lr_hd_gather_bf16": ([p, p, p, p, p, i32, i32, p], i32),
    "lr_embed_scatter_bf16": ([p, p, p, p, p, p, i32, i32, i32, i32, i32, p], i32),
    "lr_skipca_scores": ([p, i32, p, i32, p, p, i32, i32, i32, p], i32),
    "lr_skipca_head": ([p, p, i32, p, p, i32, p, p, p, i32, i32, i32, i32, f32, p], i32),
    "lr_preference": ([p, p, p, i32, i32, i32, f32, p], i32),
    "lr_resample_u8": ([p, i32, i32, p, i32, i32, i32, p, p, i32, p], i32),
    "lr_patch_pack_f32": ([p, i32, i32, i32, i32, i32, i32, p, p, p], i32),
    "lr_hd_pack_f32": ([p, i32, i32, i32, i32, i32, i32, p, p, p, i32, p], i32),
    "lr_gemm_rope_ex_bf16": ([p, i32, p, i32, p, i32, i32, i32, i32, p, p, p, p, i32, i32, i32, p], i32),
    "lr_attention_ex_bf16": ([p, p, p, p, i32, i32, i32, i32, i32, p, p, p, i32, i32, i32, i32, f32, i32, p], i32),
    "lr_attention_seg_bf16": ([p, p, p, p, i32, i32, i32, p, p, i32, i32, f32, p], i32),
    "lr_skipca_scores_ex": ([p, i32, p, i32, p, p, i32, i32, i32, f32, p], i32),
    "lr_patch_rows_bf16": ([p, p, p, i32, i32, i32, i32, p], i32),
    "lr_qwen_patchify_f32": ([p, i32, i32, i32, i32, p, p, p], i32),
    "lr_mrope_plan": ([p, p, i32, i32, i64, p, i32, i32, p, p, p, i32, i32, i32, i32, p, p, p, p, p], i32),
    "lr_gather_rows_bf16": ([p, i32, p, p, i32, i32, i32, p], i32),
    "lr_compact_rows_bf16": ([p, i32, p, p, p, i32, i32, i32, i32, p], i32),
    "lr_softmax_rows_bf16": ([p, i32, i32, i32, i32, f32, p], i32),
    "lr_masked_mean_rows_bf16": ([p, i32, p, p, i32, i32, i32, i32, p], i32),
    "lr_synth_normal_f32": ([p, i64, C.c_uint32, f32, f32, i32, p], i32),
    # fp32 verification path (csrc/f32_verify.cu)
    "lr_f32_gemm": ([p, i32, p, i32, p, i32, i32, i32, i32, i32, p, p, i32, p], i32),
    "lr_f32_swiglu": ([p, i32, p, i32, i32, i32, p], i32),
    "lr_f32_rope": ([p, i32, p, p, p, i32, i32, i32, p], i32),
    "lr_f32_rmsnorm": ([p, i32, p, p, p, i32, i32, i32, f32, p], i32),
    "lr_f32_layernorm": ([p, i32, p, p, p, i32, i32, i32, f32, p], i32),
    "lr_f32_clip_im2col": ([p, p, p, i32, p], i32),
    "lr_f32_clip_embed_ln": ([p, p, p, p, p, p, i32, f32, p], i32),
    "lr_f32_attention": ([p, p, p, p, i32, i32, i32, i32, p, p, p, i32, i32, i32, i32, f32, p], i32),
    "lr_f32_hd_gather": ([p, p, p, p, p, i32, i32, p], i32),
    "lr_f32_skipca_scores": ([p, i32, p, i32, p, p, i32, i32, i32, p], i32),
    "lr_f32_skipca_head": ([p, p, i32, p, p, i32, p, p, p, i32, i32, i32, i32, f32, p], i32),
    "lr_f32_preference": ([p, p, p, i32, i32, i32, f32, p], i32),
}


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (raises LibraryMissing with build instructions if absent)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). llava-reward-b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
        _lib = lib
    return _lib


def check(status: int, what: str):
    if status != LR_OK:
        if status < 0:
            raise RuntimeError(f"{what}: {_ERR.get(status, status)}")
        raise RuntimeError(f"{what}: CUDA error {status}")


_launches = 0


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count() -> int:
    """Kernels launched through this binding since the last reset (every compute entry = one launch)."""
    return _launches


def call(name: str, *args):
    global _launches
    check(getattr(load(), name)(*args), name)
    _launches += 1
