/*
 * llava_reward_b200.h - C ABI of libllavareward.so (hand-written sm_100a CUDA).
 *
 * The reference (sjz5202/LLaVA-Reward) has no FFI/plugin layer of its own: its hot path is
 * Python calling torch/cuBLAS/cuDNN/flash-attn. This header is therefore the
 * boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub) to replace, call site
 * by call site, the library launches issued from
 *   llava_reward/models/rw_model_general_preference.py:334-448   (custom_forward)
 *   llava_reward/models/base_mllm/phi3_v/modeling_phi3_v.py      (Phi3VModel.forward and below)
 *   llava_reward/models/base_mllm/phi3_v/processing_phi3_v.py    (Phi3VImageProcessor.preprocess)
 *   eval/reward_adaptor_loader.py:174-181                        (preference_compute)
 *
 * Conventions: every entry takes raw DEVICE pointers, explicit sizes and leading dimensions
 * (in elements), and a cudaStream_t passed as void* LAST. Nothing allocates, nothing keeps global
 * state except a cached driver entry point; calls are re-entrant per stream. Return value: 0 = ok,
 * negative = LR_ERR_*, positive = cudaError_t passthrough from the launch. bf16 tensors are
 * row-major; "ld" = elements between consecutive rows. There is no CPU fallback: on a machine
 * without an sm_100 device every compute entry returns an error.
 */
#ifndef LLAVA_REWARD_B200_H
#define LLAVA_REWARD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LR_OK 0
#define LR_ERR_BAD_ARG (-1)     /* null pointer, non-positive size, unsupported shape */
#define LR_ERR_ALIGN (-2)       /* pointer / leading dimension not 16-byte aligned */
#define LR_ERR_NO_DRIVER (-3)   /* cuTensorMapEncodeTiled not obtainable (no CUDA driver) */
#define LR_ERR_UNSUPPORTED (-4) /* device is not sm_100 */

/* GEMM epilogues (what follows the nn.Linear in the reference) */
#define LR_EPI_NONE 0           /* C = bf16(acc)                                   W_q/W_k/W_v, lora_A */
#define LR_EPI_BIAS 1           /* C = bf16(acc + bias)                            CLIP q|k|v, projector.2 */
#define LR_EPI_BIAS_QUICKGELU 2 /* x=bf16(acc+bias); C = x*sigmoid(1.702x)         CLIP fc1 (modeling_phi3_v.py:71) */
#define LR_EPI_BIAS_GELU 3      /* x=bf16(acc+bias); C = gelu_erf(x)               projector.0 + nn.GELU (:172-179) */
#define LR_EPI_RESIDUAL 4       /* C = bf16(bf16(acc) + R)                         o_proj / down_proj (:1189,1194) */
#define LR_EPI_BIAS_RESIDUAL 5  /* C = bf16(bf16(acc+bias) + R)                    CLIP out_proj / fc2 */
#define LR_EPI_SWIGLU 6         /* W rows packed [gate128|up128] per 256; C[:,N/2] = up*silu(gate)  Phi3MLP (:566-572) */
#define LR_EPI_ROPE 7           /* su-RoPE on columns [0, rope_cols) (only through lr_gemm_rope_bf16)                  */
#define LR_EPI_BIAS_SWIGLU 8    /* as LR_EPI_SWIGLU with gate/up biases (packed like the W rows): C = bf16(up+b_u)*silu(bf16(gate+b_g))
                                   Qwen2_5_VLMLP(bias=True) of the vision blocks (transformers modeling_qwen2_5_vl.py:77-88) */
#define LR_EPI_BIAS_ROPE 9      /* x = bf16(acc + bias), then LR_EPI_ROPE on columns [0, rope_cols) (only lr_gemm_rope_ex_bf16) */
#define LR_EPI_BIAS_ROPE_F32 10 /* x = bf16(acc + bias); out = bf16(x1*cos - x2*sin), bf16(x2*cos + x1*sin) in fp32 with fp32
                                   tables: apply_rotary_pos_emb_vision (modeling_qwen2_5_vl.py:156-167) (only lr_gemm_rope_ex_bf16) */

#define LR_GEMM_TCGEN05 0 /* tcgen05.mma + TMEM + TMA pipeline (product path): CTA-pair kernel when N % 256 == 0 and
                             M > 256, else the single-CTA kernel */
#define LR_GEMM_SIMT 1    /* plain CUDA-core kernel, used only to cross-check the tcgen05 path in tests */
#define LR_GEMM_TCGEN05_PAIR 2   /* force the CTA-pair kernel (cta_group::2, 256x256 tile per cluster of 2); N % 256 == 0 */
#define LR_GEMM_TCGEN05_SINGLE 3 /* force the single-CTA kernel (128 x BN tile, BN = 256 or 128) */

int lr_version(void);
/* 0 when the current CUDA device is compute capability 10.x, else LR_ERR_UNSUPPORTED / cudaError. */
int lr_device_check(void);

/* C[M, N'] = epilogue(A[M,K] . W[N,K]^T); A, W, C, R bf16; bias bf16 [N]; fp32 accumulation.
 * N' = N/2 for LR_EPI_SWIGLU else N. K % 64 == 0, N % 128 == 0; M arbitrary (TMA zero-fills).
 * LoRA is applied by K-extension: A = [x | x.A^T], W = [W0 | (alpha/r).B]  (peft lora.Linear.forward).
 * Replaces: every nn.Linear / F.linear on the path (modeling_phi3_v.py:101-114,172-179,567-572,769,880;
 * rw_model_general_preference.py:378-380; HF modeling_clip.py CLIPMLP/CLIPAttention). */
int lr_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                 int epilogue, const void* bias, const void* R, int ldr, int impl, void* stream);

/* Fused qkv projection + rotary embedding: C = A . W^T with su-RoPE applied in the epilogue to columns
 * [0, rope_cols) (the q and k thirds). W's q/k rows must be packed with the two halves of every head interleaved
 * (new row 2i = old i, 2i+1 = old i + head_dim/2), so a rotation pair is two adjacent output columns; q.k^T is
 * invariant under this common permutation of q and k, v is untouched. Arithmetic per pair (x1, x2) = bf16(acc):
 * out1 = bf16(bf16(x1*cos) + bf16(-x2*sin)), out2 = bf16(bf16(x2*cos) + bf16(x1*sin)) with bf16 tables
 * cos/sin[pos, head_dim/2] - the same op-by-op rounding as lr_rope_su_bf16 / apply_rotary_pos_emb
 * (modeling_phi3_v.py:529-553). N % 256 == 0, rope_cols % 256 == 0, head_dim % 32 == 0.
 * impl: LR_GEMM_TCGEN05 (CTA-pair kernel) or LR_GEMM_SIMT (cross-check). */
int lr_gemm_rope_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                      const int* position_ids, const void* cos_tab, const void* sin_tab, int rope_cols, int head_dim,
                      int impl, void* stream);

/* lr_gemm_rope_bf16 with a linear bias (bf16 [N], packed like the W rows, added to every column before the rounding /
 * rotation) and a table mode: epilogue = LR_EPI_BIAS_ROPE (bf16 tables, op-by-op bf16 rounding: the Qwen2 decoder's
 * q/k/v projection + apply_multimodal_rotary_pos_emb, transformers modeling_qwen2_5_vl.py:627-669, 725-739) or
 * LR_EPI_BIAS_ROPE_F32 (fp32 tables cos/sin[pos, head_dim/2], one rounding: the vision blocks' qkv +
 * apply_rotary_pos_emb_vision, :156-167, 231-238). position_ids == NULL means pos = row (per-token tables). */
int lr_gemm_rope_ex_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                         const void* bias, const int* position_ids, const void* cos_tab, const void* sin_tab,
                         int rope_cols, int head_dim, int epilogue, void* stream);

/* y[i,:] = w * bf16(x[r,:] * rsqrt(mean(x[r,:]^2) + eps)),  r = row_index ? row_index[i] : i.
 * Replaces Phi3RMSNorm.forward (modeling_phi3_v.py:386-391). cols % 8 == 0, cols <= 8192. */
int lr_rmsnorm_bf16(const void* x, int ldx, const int* row_index, const void* w, void* y, int ldy, int rows,
                    int cols, float eps, void* stream);

/* y = LayerNorm(x) * w + b over the last dim (fp32 statistics, one rounding). Replaces nn.LayerNorm in
 * HF CLIPEncoderLayer.layer_norm1 / layer_norm2 and CLIPVisionTransformer.pre_layrnorm (installed transformers 5.5
 * modeling_clip.py:371, :380, :677; reached from the reference at modeling_phi3_v.py:208-219). cols % 8 == 0, cols <= 8192. */
int lr_layernorm_bf16(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows, int cols,
                      float eps, void* stream);

/* Patch-embedding operand: pixels fp32 [*,3,336,336] (crop c read at slot crop_src[c]) ->
 * A[n_crops*576, 640] bf16, column = ch*196 + ky*14 + kx (Conv2d weight order), columns 588..639 zero.
 * Replaces the im2col inside cuDNN conv (modeling_clip.py CLIPVisionEmbeddings.forward). */
int lr_clip_im2col(const float* pixels, const int* crop_src, void* A, int n_crops, void* stream);

/* tokens[c*577+t,:] = LayerNorm( (t==0 ? class_emb : patch[c*576+t-1,:]) + pos_emb[t,:] ) : CLS concat,
 * position add (bf16) and pre_layrnorm fused (modeling_clip.py CLIPVisionEmbeddings.forward + pre_layrnorm). */
int lr_clip_embed_ln(const void* patch, const void* class_emb, const void* pos_emb, const void* ln_w,
                     const void* ln_b, void* tokens, int n_crops, float eps, void* stream);

/* Flash-style prefill attention, bf16 in/out, fp32 softmax. q/k/v point at column 0 of the first head inside a
 * fused projection buffer with row stride ld_qkv; head h occupies columns [h*head_dim, (h+1)*head_dim).
 * Sequence s owns rows [s*rows_per_seq, (s+1)*rows_per_seq); only rows [start, start+len) are valid
 * (seq_start/seq_len may be NULL = whole slot). Invalid rows of o are zero-filled.
 * head_dim 64 (CLIP, non-causal; replaces CLIPAttentionFA2, modeling_phi3_v.py:85-115) or
 * 96 (Phi-3, causal varlen; replaces Phi3FlashAttention2._flash_attention_forward, :888-986).
 * For LR_ATTN_TCGEN05 q, k, v must be column offsets into one row-major buffer (the fused qkv projection). */
#define LR_ATTN_TCGEN05 0 /* tcgen05.mma + TMEM + TMA, the product configuration for the head_dim: the one-query-tile
                             pipeline with two CTAs per SM, several tiles per CTA where that pays (= LR_ATTN_TCGEN05_MULTITILE,
                             head_dim 64 / 96); see LR_ATTN_TCGEN05_2TILE for head_dim 128 */
#define LR_ATTN_MMA_SYNC 1 /* mma.sync kernel, kept to cross-check the tcgen05 path in tests */
#define LR_ATTN_TCGEN05_2TILE 3 /* tcgen05 kernel with two query tiles per CTA sharing K/V, one CTA per SM (slower on CLIP;
                                   at head_dim 128 the row sums live in registers and K/V are single-staged) */
#define LR_ATTN_TCGEN05_1TILE 4 /* tcgen05 kernel, strictly one query tile per CTA (the reference point of the multi-tile form) */
#define LR_ATTN_TCGEN05_MULTITILE 5 /* the one-tile pipeline walking several query tiles per CTA (causal: the pair
                                       {nt-1-x, x}, every CTA nt+1 K/V blocks; otherwise up to 5 consecutive tiles): the
                                       6.4 us per-CTA fixed cost is paid once per CTA instead of once per tile */
#define LR_ATTN_TCGEN05_SPLIT 2 /* tcgen05 kernel with two softmax threads per query row (4 softmax warpgroups);
                                   measured slower than the default (the kernel is shared-memory-bandwidth bound) */
int lr_attention_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int n_seq,
                      int rows_per_seq, const int* seq_start, const int* seq_len, int n_heads, int head_dim,
                      int causal, float scale, int impl, void* stream);

/* lr_attention_bf16 (tcgen05 kernels only) with grouped-query attention (query head h reads K/V head
 * h / (n_heads / n_kv_heads); k and v are n_kv_heads*head_dim wide) and, when seq_base != NULL, PACKED variable-length
 * sequences: sequence s owns exactly rows [seq_base[s], seq_base[s] + seq_len[s]) of the total_rows-row buffer and
 * max_len bounds every seq_len (it sizes the grid); rows of other sequences are never written. seq_base == NULL: the
 * slot layout of lr_attention_bf16 (total_rows = n_seq * max_len). head_dim 96 non-causal (the Qwen2.5-VL vision tower's
 * head_dim 80 zero-padded to 96: window attention = packed sequences of <= 64 tokens, full attention = one sequence per
 * image; replaces Qwen2_5_VLVisionAttention's varlen flash-attention call, transformers modeling_qwen2_5_vl.py:244-262)
 * and head_dim 128 causal GQA (Qwen2_5_VLAttention, :739-753) in addition to the lr_attention_bf16 shapes. */
int lr_attention_ex_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows,
                         int n_seq, int max_len, const int* seq_base, const int* seq_start, const int* seq_len,
                         int n_heads, int n_kv_heads, int head_dim, int causal, float scale, int impl, void* stream);

/* Non-causal attention over SHORT SEGMENTS of one packed buffer (tcgen05 kernel): row r attends to the key rows
 * [row_lo[r], row_hi[r]) (int32 per row, device; segments are contiguous, row_lo[r] <= r < row_hi[r], and at most 128
 * rows long so that a 128-row query tile needs at most ~256 key rows). One CTA per 128 consecutive rows and head:
 * full tiles for the <= 64-token windows of the Qwen2.5-VL vision blocks (window attention through cu_window_seqlens,
 * transformers modeling_qwen2_5_vl.py:500-505, 244-262) instead of one half-empty tile per window. head_dim 64 or 96. */
int lr_attention_seg_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows,
                          const int* row_lo, const int* row_hi, int n_heads, int head_dim, float scale, void* stream);

/* In-place su/longrope rotary embedding on the q and k thirds of a fused qkv buffer [rows, 3*n_heads*head_dim]:
 * x = bf16(bf16(x*cos) + bf16(rot_half(x)*sin)) with bf16 tables cos/sin[pos, head_dim/2].
 * Replaces Phi3SuScaledRotaryEmbedding + apply_rotary_pos_emb (modeling_phi3_v.py:438-476, 529-553). */
int lr_rope_su_bf16(void* qkv, int ld, const int* position_ids, const void* cos_tab, const void* sin_tab,
                    int rows, int n_heads, int head_dim, void* stream);

/* One pass over input_ids / attention_mask [B,S] (int64):
 *   position_ids = cumsum(mask)-1, 1 where mask==0     (rw_model_general_preference.py:344-345)
 *   img_ord[b,s] = ordinal of s among the image positions (-1e9 < id < 0) of row b, else -1 (modeling_phi3_v.py:228)
 *   seq_start/seq_len = first valid column / number of valid columns; eos_row = b*S + last valid column (:420,439)
 *   n_img[b] = number of image positions; flags[0] |= 1 if some mask row is not one contiguous run. */
int lr_token_plan(const int64_t* input_ids, const int64_t* attention_mask, int B, int S, int* position_ids,
                  int* img_ord, int* seq_start, int* seq_len, int* eos_row, int* n_img, int* flags, void* stream);

/* lr_token_plan with the image placeholder and position rule as parameters (LLaVA-v1.6 branch):
 *   image_token_id >= 0: image positions are ids == image_token_id (LlavaNextModel.get_placeholder_mask,
 *   transformers modeling_llava_next.py:419-441); < 0: negative ids as lr_token_plan.
 *   position_mode LR_POS_FROM_MASK: as lr_token_plan; LR_POS_ARANGE: position_ids[b,s] = s - what LlamaModel uses
 *   when the caller passes no position_ids, which is how the reference's llava branch calls it
 *   (rw_model_general_preference.py:372-374). */
#define LR_POS_FROM_MASK 0
#define LR_POS_ARANGE 1
int lr_token_plan_ex(const int64_t* input_ids, const int64_t* attention_mask, int B, int S, int64_t image_token_id,
                     int position_mode, int* position_ids, int* img_ord, int* seq_start, int* seq_len, int* eos_row,
                     int* n_img, int* flags, void* stream);

/* per-sample plan record (int32 x 8) shared by lr_hd_gather_bf16 / lr_embed_scatter_bf16 / lr_skipca_* */
#define LR_PLAN_STRIDE 8
#define LR_PLAN_HCROP 0     /* image_sizes[b][0] / 336 */
#define LR_PLAN_WCROP 1     /* image_sizes[b][1] / 336 */
#define LR_PLAN_CROP_BASE 2 /* index of the sample's global crop in the compacted crop list */
#define LR_PLAN_ROW_BASE 3  /* first row of the sample in the concatenated image-token matrix */
#define LR_PLAN_NV 4        /* number of image tokens of the sample */
#define LR_PLAN_TOP 5       /* anyres only: feature rows removed at the top (= bottom) by unpad_image */
#define LR_PLAN_LEFT 6      /* anyres only: feature columns removed at the left (= right) */

/* HD feature transform: CLIP tokens [n_crops*577, 1024] (row 0 of each crop = CLS, skipped) ->
 * rows [sum N_v, 4096] in 'sub_glb' order with sub_GN newlines and the glb_GN separator.
 * Replaces hd_feature_transform / reshape_hd_patches_2x2merge / add_image_newline (modeling_phi3_v.py:254-362). */
int lr_hd_gather_bf16(const void* clip_tokens, const int* plan, const void* sub_gn, const void* glb_gn, void* rows,
                      int B, int max_nv, void* stream);

/* hidden[b,s,:] = img_ord[b,s] >= 0 ? img_proj[row_base[b] + img_ord[b,s], :] : wte[clamp(id,0,V), :].
 * Replaces wte + index_put (modeling_phi3_v.py:230-231, 247-249). */
int lr_embed_scatter_bf16(const int64_t* input_ids, const int* img_ord, const int* plan, const void* wte,
                          const void* img_proj, void* hidden, int ldh, int B, int S, int H, int V, void* stream);

/* LLaVA-v1.6: hidden[b,s,:] = image position ? anyres-packed image feature : wte[id]. The packed order is
 * [576 base-patch tokens | unpadded (grid_h*24 - 2*top) x (grid_w*24 - 2*left) grid, row-major, image_newline after each
 * grid row]; plan record: HCROP/WCROP = grid_h/grid_w, CROP_BASE = first patch of the sample in `feat`, TOP, LEFT.
 * feat = projector output for all 577 CLIP tokens of every patch (row stride ldf; the CLS rows are never read).
 * Replaces pack_image_features + masked_scatter (transformers modeling_llava_next.py:277-343, 484-494) as called by
 * the reference's llava branch (rw_model_general_preference.py:372-375). */
int lr_anyres_embed_scatter_bf16(const int64_t* input_ids, const int* img_ord, const int* plan, const void* wte,
                                 const void* feat, int ldf, const void* image_newline, void* hidden, int ldh, int B,
                                 int S, int H, int V, void* stream);

/* SkipCA, last-valid-token row only (the only row the reference consumes in eval, :420-421/439-444):
 * scores[b,j] = bf16(bf16(q_b . K_bj) / sqrt(H)), j < N_v(b).  kv = [K | V] rows of the concatenated image tokens. */
int lr_skipca_scores(const void* q, int ldq, const void* kv, int ldkv, const int* plan, float* scores, int B, int H,
                     int max_nv, void* stream);

/* lr_skipca_scores with the score written for the rows N_v(b) <= j < max_nv as a parameter: 0 reproduces the Phi-3
 * arm (zero-padded vision rows: q.0 = 0), bf16(-1e4) = -9984 the qwen arm's masked_fill (rw_model_general_preference.py
 * :387-392; K/V rows = hidden_states[0] at the token-id-151643 positions, :358-371, gathered by lr_compact_rows_bf16). */
int lr_skipca_scores_ex(const void* q, int ldq, const void* kv, int ldkv, const int* plan, float* scores, int B, int H,
                        int max_nv, float pad_score, void* stream);

/* reward[b,:] = value_head( ca_ln( x_b + sum_j softmax_j(scores_b over max_nv incl. zero-padded rows) V_bj ) ).
 * scores == NULL skips the cross-attention (BT / no-SkipCA models: reward = value_head(x_b)).
 * Replaces rw_model_general_preference.py:381-386 (softmax, bmm, residual, ca_layernorm) and :407-448. */
int lr_skipca_head(const float* scores, const void* kv, int ldkv, const int* plan, const void* x, int ldx,
                   const void* ca_ln_w, const void* value_head_w, void* reward, int B, int H, int max_nv, int vhd,
                   float eps, void* stream);

/* prob[i] = sigmoid((c0*r1 - c1*r0)/tau) (GPM, vhd==2) or sigmoid((c - r)/tau), bf16 arithmetic like the
 * reference, fp32 output. Replaces preference_compute (eval/reward_adaptor_loader.py:174-181). */
int lr_preference(const void* chosen, const void* reject, float* prob, int n, int vhd, int is_gpm, float tau,
                  void* stream);

/* One pass of Pillow's 8-bit antialiased resample (ImagingResampleHorizontal/Vertical_8bpc) along `axis`
 * (1 = horizontal, 0 = vertical) of an H x W x 3 uint8 image on the device. bounds[o] = {first source index, taps},
 * coeffs[o][ksize] = 22-bit fixed-point taps, both precomputed on the host exactly as Pillow's precompute_coeffs +
 * normalize_coeffs_8bpc (device pointers). Replaces torchvision.transforms.functional.resize on a PIL image
 * (processing_phi3_v.py:98). Bit-exact. */
int lr_resample_u8(const uint8_t* src, int src_h, int src_w, uint8_t* dst, int dst_h, int dst_w, int axis,
                   const int* bounds, const int* coeffs, int ksize, void* stream);

/* Resized uint8 image [rh, rw, 3] placed at (pad_top, pad_left) inside a white H x W canvas (H, W multiples of 336)
 * -> out fp32 [n_slots, 3, 336, 336]: slot 0 = bicubic (A=-0.75, align_corners=False) 336x336 view of the normalised
 * canvas, slots 1..(H/336*W/336) = its 336x336 crops in row-major order, remaining slots zero. Normalisation =
 * (u8/255 - mean)/std in IEEE fp32 (mean3/std3 are HOST pointers to 3 floats). Replaces padding_336, ToTensor,
 * Normalize, F.interpolate(bicubic), the crop reshape/permute and pad_to_max_num_crops_tensor
 * (processing_phi3_v.py:62-71, 128-136, 252-277). Crops bit-exact, global view within 1e-5. */
int lr_hd_pack_f32(const uint8_t* img, int rh, int rw, int pad_top, int pad_left, int H, int W, const float* mean3,
                   const float* std3, float* out, int n_slots, void* stream);

/* LLaVA-v1.6 anyres patches: uint8 HWC image [rh, rw, 3] (device) placed at (pad_top, pad_left) on a zero canvas of
 * (grid_h*336) x (grid_w*336), split row-major into grid_h*grid_w patches out[p][3][336][336] fp32.
 * lut768 (HOST pointer, [3][256] floats) maps a uint8 channel value to its rescaled + normalised float; the caller
 * builds it with the processor's own arithmetic so the result is bit-identical. Replaces _pad_for_patching /
 * divide_to_patches / rescale / normalize of transformers' LlavaNextImageProcessor (image_processing_pil_llava_next.py
 * :105-146, 205-225) used by the reference through AutoProcessor (llava_reward/datasets/reward_dataset.py:334-346).
 * The base view is the same call with grid 1x1 on the image resized to 336x336. */
int lr_patch_pack_f32(const uint8_t* img, int rh, int rw, int pad_top, int pad_left, int grid_h, int grid_w,
                      const float* lut768, float* out, void* stream);

/* ---- Qwen2.5-VL branch (reference model_type == 'qwen', rw_model_general_preference.py:354-371) ------------------ */

/* out[i, 0:K] = bf16(pixels[src_row[i], 0:K]), out[i, K:Kpad] = 0 (src_row NULL = identity): the A operand of the
 * patch-embedding GEMM, rows already in the vision tower's window order. pixels = the processor's flattened patches
 * fp32 [T, 3*2*14*14]. Replaces the .to(bf16) + Conv3d im2col of Qwen2_5_VisionPatchEmbed.forward and the
 * hidden_states[window_index] gather (transformers modeling_qwen2_5_vl.py:108-114, 484-486). K % 4 == 0, Kpad % 8 == 0. */
int lr_patch_rows_bf16(const float* pixels, const int* src_row, void* out, int ldo, int rows, int K, int Kpad,
                       void* stream);

/* Qwen2-VL image -> flattened patches: uint8 HWC image [H, W, 3] (device; already resized by smart_resize, H and W
 * multiples of patch*merge) -> out fp32 [(H/patch)*(W/patch), 3*2*patch*patch]: rescale + normalise through lut768
 * (HOST pointer, [3][256] floats, built by the caller with the processor's own arithmetic so the result is
 * bit-identical), the still image repeated over the temporal_patch_size = 2 axis, rows in 2x2-merge order and columns
 * in (channel, t, py, px) order. Replaces rescale / normalize / the reshape-transpose of transformers'
 * Qwen2VLImageProcessor._preprocess (image_processing_pil_qwen2_vl.py:186-217) used by the reference through
 * AutoProcessor (llava_reward/datasets/reward_dataset.py:472-487). patch even. */
int lr_qwen_patchify_f32(const uint8_t* img, int H, int W, int patch, int merge, const float* lut768, float* out,
                         void* stream);

/* M-RoPE plan, images only: pos3[c, b*S+s] (c = temporal, height, width; int32, comp stride B*S) as
 * Qwen2_5_VLForConditionalGeneration.get_rope_index of transformers 4.50 (the release the reference pins; called from
 * the forward the reference invokes at rw_model_general_preference.py:357) computes them from input_ids,
 * image_grid_thw (int32 [n_images,3], device) and the attention mask - padded positions get 0 - and the per-token rotary
 * rows cos_out/sin_out[b*S+s, i] = cos_tab/sin_tab[pos3[sec(i)], i] (bf16 [max_pos, half] tables; sec = 0 for
 * i < sec0, 1 for i < sec0+sec1, else 2: apply_multimodal_rotary_pos_emb, modeling_qwen2_5_vl.py:627-669).
 * run_count: int32 [B] scratch (image runs per sample). flags |= 2 when a run of image tokens disagrees with
 * image_grid_thw, |= 4 when a position reaches max_pos. S <= 12800. Two launches. */
int lr_mrope_plan(const int64_t* input_ids, const int64_t* attention_mask, int B, int S, int64_t image_token_id,
                  const int* grid_thw, int n_images, int merge, int* run_count, const void* cos_tab,
                  const void* sin_tab, int max_pos, int half, int sec0, int sec1, int* pos3, void* cos_out,
                  void* sin_out, int* flags, void* stream);

/* dst[plan[b].ROW_BASE + ord[b,s], :] = src[b*S+s, :] for every position with ord[b,s] >= 0 (ord = the img_ord output
 * of lr_token_plan_ex for some token id). Replaces the vision_pad gather loop of the reference's qwen SkipCA arm
 * (rw_model_general_preference.py:362-371). cols % 8 == 0. */
int lr_compact_rows_bf16(const void* src, int lds, const int* ord, const int* plan, void* dst, int ldd, int B, int S,
                         int cols, void* stream);

/* dst[i, :] = src[row_index[i], :], zeros where row_index[i] < 0 (bf16 rows, cols % 8 == 0). Used to continue the LAST decoder layer on the
 * last-valid-token rows only - the only rows of hidden_states[-1] the reference's eval-mode head reads
 * (rw_model_general_preference.py:420-421, 439-444): o_proj / MLP of that layer run on B rows instead of B*S. */
int lr_gather_rows_bf16(const void* src, int lds, const int* row_index, void* dst, int ldd, int rows, int cols,
                        void* stream);

/* ---- all-rows head: `mean_hidden_state` pooling (rw_model_general_preference.py:398-406) ---------------------------
 * With that attribute set the reference pools EVERY row of the (SkipCA'd) last hidden state, so the S x N_v cross
 * attention of :376-386 is needed for all rows: Q K^T and P V run as lr_gemm_bf16 per sample on the zero-padded
 * vision rows (lr_gather_rows_bf16 with -1 indices builds the reference's zero padding), and these two entries are the
 * pieces between the GEMMs.
 *
 * In place on bf16 scores [rows, lds]: p_j = bf16(softmax_j(bf16(s_j * inv_sqrt_d))) over columns j < n_valid,
 * 0 for n_valid <= j < n_total (n_total % 8 == 0). Replaces `scores / sqrt(d_k)` + F.softmax (:381-383). */
int lr_softmax_rows_bf16(void* scores, int lds, int rows, int n_valid, int n_total, float inv_sqrt_d, void* stream);

/* out[b,:] = bf16(bf16(sum_s x[b*S+s,:] * mask[b,s]) / bf16(sum_s mask[b,s])) (H % 256 == 0). Replaces the masked mean
 * of :398-406 (each torch op rounds to bf16 once; mask.sum() is itself a bf16 tensor). */
int lr_masked_mean_rows_bf16(const void* x, int ldx, const int64_t* attention_mask, void* out, int ldo, int B, int S,
                             int H, void* stream);

/* Synthetic weights: out[i] = fp32(s_i) * scale (+ mean), s_i = centred sum of the four 16-bit halves of two murmur3-
 * finalised counters (2i, 2i+1) * 0x9E3779B1 + key - the counter-hash generator of llava_reward_b200/synth.py, bit-
 * identical to its torch-on-CPU form. The reference has no counterpart (it loads hub checkpoints,
 * eval/reward_adaptor_loader.py:31-42); BASELINE.json asks for random-init weights of the named architecture, and
 * 4.4 G values are generated in place on the device instead of being shipped. out fp32 [n]. */
int lr_synth_normal_f32(void* out, int64_t n, uint32_t key, float scale, float mean, int add_mean, void* stream);

/* ---- fp32 verification path (csrc/f32_verify.cu) -------------------------------------------------------------------
 * Plain fp32 CUDA-core kernels WITHOUT any bf16 rounding point, used by RewardEngine(precision="fp32") to show the
 * engine's dataflow, layouts and index work against the reference's fp32 outputs at 1e-4 (north_star; SURVEY 8d
 * parity gates). All tensors fp32, same operand layouts as the bf16 entries they mirror (W is [N, K] row-major, the
 * LoRA K-extension, head-interleaved q/k rows, [gate128|up128] blocks). Debug only: not a scoring path.
 * The byte-moving product kernels serve fp32 rows unchanged: lr_gather_rows_bf16 / lr_embed_scatter_bf16 are called
 * with every width and leading dimension doubled; lr_f32_hd_gather is hd_gather_kernel on 2048 two-byte units. */
int lr_f32_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epilogue,
                const void* bias, const void* R, int ldr, void* stream); /* NONE, BIAS, BIAS_QUICKGELU, BIAS_GELU,
                                                                            RESIDUAL, BIAS_RESIDUAL */
int lr_f32_swiglu(const void* raw, int ldr, void* out, int ldo, int M, int N, void* stream); /* out[:, N/2] from the
                                                   packed [gate128|up128] columns of raw (Phi3MLP, modeling_phi3_v.py:566-572) */
int lr_f32_rope(void* x, int ld, const int* position_ids, const void* cos_tab, const void* sin_tab, int rows,
                int rope_cols, int head_dim, void* stream); /* in place, head-interleaved pairs, fp32 tables [pos, hd/2]
                                                               (apply_rotary_pos_emb, :529-553) */
int lr_f32_rmsnorm(const void* x, int ldx, const int* row_index, const void* w, void* y, int ldy, int rows, int cols,
                   float eps, void* stream);
int lr_f32_layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows, int cols, float eps,
                     void* stream);
int lr_f32_clip_im2col(const float* pixels, const int* crop_src, void* A, int n_crops, void* stream);
int lr_f32_clip_embed_ln(const void* patch, const void* class_emb, const void* pos_emb, const void* ln_w, const void* ln_b,
                         void* tokens, int n_crops, float eps, void* stream);
int lr_f32_attention(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int n_seq,
                     int rows_per_seq, const int* seq_base, const int* seq_start, const int* seq_len, int n_heads,
                     int n_kv_heads, int head_dim, int causal, float scale, void* stream); /* exact softmax; slot layout
                                                   (seq_base NULL) or packed sequences, like lr_attention_ex_bf16 */
int lr_f32_hd_gather(const void* clip_tokens, const int* plan, const void* sub_gn, const void* glb_gn, void* rows, int B,
                     int max_nv, void* stream);
int lr_f32_skipca_scores(const void* q, int ldq, const void* kv, int ldkv, const int* plan, void* scores, int B, int H,
                         int max_nv, void* stream);
int lr_f32_skipca_head(const void* scores, const void* kv, int ldkv, const int* plan, const void* x, int ldx,
                       const void* ca_ln_w, const void* value_head_w, void* reward, int B, int H, int max_nv, int vhd,
                       float eps, void* stream);
int lr_f32_preference(const void* chosen, const void* reject, void* prob, int n, int vhd, int is_gpm, float tau,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LLAVA_REWARD_B200_H */
