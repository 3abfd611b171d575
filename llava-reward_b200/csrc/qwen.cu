// Qwen2.5-VL branch: index / gather kernels around the GEMMs (HBM-bound, 16-byte accesses).
//   patch_rows_kernel   fp32 flattened patches -> bf16 GEMM operand in window order (patch-embed Conv3d = GEMM)
//   mrope_*_kernel      M-RoPE positions (get_rope_index) + per-token cos/sin rows
//   compact_rows_kernel rows of a [B*S, H] matrix at flagged positions -> dense row list (SkipCA qwen arm)
#include "common.cuh"

namespace lr {

// out[i, 0:K] = bf16(pix[src_row[i], 0:K]), out[i, K:Kpad] = 0. One warp per row; K % 4 == 0, Kpad % 8 == 0.
__global__ void __launch_bounds__(256)
patch_rows_kernel(const float* __restrict__ pix, const int* __restrict__ src_row, bf16* __restrict__ out, int ldo,
                  int rows, int K, int Kpad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* src = pix + size_t(src_row ? src_row[warp] : warp) * K;
  bf16* dst = out + size_t(warp) * ldo;
  for (int c = lane * 8; c < Kpad; c += 256) {
    float f[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cc = c + 4 * h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc + 3 < K) v = __ldg(reinterpret_cast<const float4*>(src + cc));
      f[4 * h] = v.x, f[4 * h + 1] = v.y, f[4 * h + 2] = v.z, f[4 * h + 3] = v.w;
    }
    stg128(dst + c, make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                               pack_bf16x2(f[6], f[7])));
  }
}

// number of image runs (maximal runs of image_token_id among the valid tokens) per sample
__global__ void __launch_bounds__(256)
mrope_count_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask, int S, int64_t image_token_id,
                   int* __restrict__ run_count) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int64_t* id = ids + size_t(b) * S;
  const int64_t* mk = mask + size_t(b) * S;
  float n = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const bool img = mk[s] != 0 && id[s] == image_token_id;
    const bool prev = s > 0 && mk[s - 1] != 0 && id[s - 1] == image_token_id;
    n += (img && !prev) ? 1.f : 0.f;
  }
  n = block_sum(n, red);
  if (threadIdx.x == 0) run_count[b] = int(n + 0.5f);
}

// One CTA per sample. smem: tok[S] (0 = padded, 1 = text, 2 = image) | pos[3][S].
// Thread 0 walks the valid tokens in order (Qwen2_5_VLForConditionalGeneration.get_rope_index of transformers 4.50,
// images only): a text token takes (cur, cur, cur) and advances cur; an image run of t*lh*lw tokens takes
// (cur, cur + row, cur + col) on its merged grid and advances cur by max(lh, lw). Padded positions get 0.
// Then all threads write pos3 [3, B*S] and gather the per-token table rows:
//   cos_out[b*S+s, i] = cos_tab[pos[sec(i)][s], i],  sec(i) = 0 (i < sec0), 1 (i < sec0+sec1), 2 otherwise.
__global__ void __launch_bounds__(256)
mrope_pos_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mask, int S, int64_t image_token_id,
                 const int* __restrict__ grid_thw, int n_images, int merge, const int* __restrict__ run_count,
                 const bf16* __restrict__ cos_tab, const bf16* __restrict__ sin_tab, int max_pos, int half, int sec0,
                 int sec1, int* __restrict__ pos3, size_t comp_stride, bf16* __restrict__ cos_out,
                 bf16* __restrict__ sin_out, int* __restrict__ flags) {
  extern __shared__ int mp_smem[];
  int* tok = mp_smem;
  int* pos = mp_smem + S;  // [3][S]
  const int b = blockIdx.x;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const size_t g = size_t(b) * S + s;
    tok[s] = mask[g] == 0 ? 0 : (ids[g] == image_token_id ? 2 : 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int img = 0, bad = 0;
    for (int i = 0; i < b; ++i) img += run_count[i];
    int cur = 0, s = 0;
    while (s < S) {
      const int t = tok[s];
      if (t == 0) {
        pos[s] = pos[S + s] = pos[2 * S + s] = 0;
        ++s;
      } else if (t == 1) {
        pos[s] = pos[S + s] = pos[2 * S + s] = cur++;
        ++s;
      } else {
        int run = 0;
        while (s + run < S && tok[s + run] == 2) ++run;
        int gt = 1, lh = 1, lw = run;
        if (img < n_images) {
          gt = grid_thw[3 * img], lh = grid_thw[3 * img + 1] / merge, lw = grid_thw[3 * img + 2] / merge;
        } else {
          bad |= 2;
        }
        ++img;
        if (gt * lh * lw != run || lw <= 0) {  // tokens and image_grid_thw disagree: flag, keep going as one text-like row
          bad |= 2;
          gt = 1, lh = 1, lw = run;
        }
        for (int k = 0; k < run; ++k) {
          pos[s + k] = cur;
          pos[S + s + k] = cur + (k / lw) % lh;
          pos[2 * S + s + k] = cur + k % lw;
        }
        cur += max(lh, lw);
        s += run;
      }
    }
    if (bad) atomicOr(flags, bad);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * S; i += blockDim.x)
    pos3[size_t(i / S) * comp_stride + size_t(b) * S + (i % S)] = pos[i];
  const int cpr = half >> 3;  // 16-byte chunks per table row
  int over = 0;
  for (int i = threadIdx.x; i < S * cpr; i += blockDim.x) {
    const int s = i / cpr, c = (i % cpr) * 8;
    const int comp = c < sec0 ? 0 : (c < sec0 + sec1 ? 1 : 2);
    int p = pos[comp * S + s];
    if (p >= max_pos) {
      over = 1;
      p = max_pos - 1;
    }
    const size_t dst = (size_t(b) * S + s) * half + c;
    stg128(cos_out + dst, ldg128(cos_tab + size_t(p) * half + c));
    stg128(sin_out + dst, ldg128(sin_tab + size_t(p) * half + c));
  }
  if (over) atomicOr(flags, 4);
}

// dst[row_base[b] + ord[b,s], :] = src[b*S + s, :] for every (b, s) with ord[b,s] >= 0. One warp per (b, s).
__global__ void __launch_bounds__(256)
compact_rows_kernel(const bf16* __restrict__ src, int lds, const int* __restrict__ ord, const int* __restrict__ plan,
                    bf16* __restrict__ dst, int ldd, int B, int S, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * S) return;
  const int o = ord[warp];
  if (o < 0) return;
  const int b = warp / S;
  const bf16* s = src + size_t(warp) * lds;
  bf16* d = dst + size_t(plan[b * LR_PLAN_STRIDE + LR_PLAN_ROW_BASE] + o) * ldd;
  for (int c = lane * 8; c < cols; c += 256) stg128(d + c, ldg128(s + c));
}

// dst[i, :] = src[row_index[i], :], or zeros where row_index[i] < 0. One warp per row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const bf16* __restrict__ src, int lds, const int* __restrict__ row_index, bf16* __restrict__ dst,
                   int ldd, int rows, int cols) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int ri = row_index[warp];
  const bf16* s = src + size_t(ri < 0 ? 0 : ri) * lds;
  bf16* d = dst + size_t(warp) * ldd;
  for (int c = lane * 8; c < cols; c += 256) stg128(d + c, ri < 0 ? make_uint4(0, 0, 0, 0) : ldg128(s + c));
}

}  // namespace lr

using namespace lr;
static inline bool q_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int lr_patch_rows_bf16(const float* pixels, const int* src_row, void* out, int ldo, int rows, int K,
                                  int Kpad, void* stream) {
  LR_CHECK_ARG(pixels && out && rows > 0 && K > 0 && K % 4 == 0 && Kpad >= K && Kpad % 8 == 0 && ldo >= Kpad);
  if (!q_aligned16(pixels) || !q_aligned16(out) || (ldo % 8)) return LR_ERR_ALIGN;
  patch_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      pixels, src_row, reinterpret_cast<bf16*>(out), ldo, rows, K, Kpad);
  return lr_launch_status();
}

extern "C" int lr_mrope_plan(const int64_t* input_ids, const int64_t* attention_mask, int B, int S,
                             int64_t image_token_id, const int* grid_thw, int n_images, int merge, int* run_count,
                             const void* cos_tab, const void* sin_tab, int max_pos, int half, int sec0, int sec1,
                             int* pos3, void* cos_out, void* sin_out, int* flags, void* stream) {
  LR_CHECK_ARG(input_ids && attention_mask && grid_thw && run_count && cos_tab && sin_tab && pos3 && cos_out &&
               sin_out && flags && B > 0 && S > 0 && n_images >= 0 && merge > 0 && max_pos > 0);
  LR_CHECK_ARG(half > 0 && half % 8 == 0 && sec0 % 8 == 0 && sec1 % 8 == 0 && sec0 + sec1 <= half);
  const size_t smem = size_t(S) * 16;
  LR_CHECK_ARG(smem <= 200 * 1024);
  if (!q_aligned16(cos_tab) || !q_aligned16(sin_tab) || !q_aligned16(cos_out) || !q_aligned16(sin_out)) return LR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  mrope_count_kernel<<<B, 256, 0, s>>>(input_ids, attention_mask, S, image_token_id, run_count);
  int st = lr_launch_status();
  if (st != LR_OK) return st;
  cudaError_t e = cudaFuncSetAttribute(mrope_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  mrope_pos_kernel<<<B, 256, smem, s>>>(input_ids, attention_mask, S, image_token_id, grid_thw, n_images, merge,
                                        run_count, reinterpret_cast<const bf16*>(cos_tab),
                                        reinterpret_cast<const bf16*>(sin_tab), max_pos, half, sec0, sec1, pos3,
                                        size_t(B) * S, reinterpret_cast<bf16*>(cos_out),
                                        reinterpret_cast<bf16*>(sin_out), flags);
  return lr_launch_status();
}

extern "C" int lr_compact_rows_bf16(const void* src, int lds, const int* ord, const int* plan, void* dst, int ldd,
                                    int B, int S, int cols, void* stream) {
  LR_CHECK_ARG(src && ord && plan && dst && B > 0 && S > 0 && cols > 0 && cols % 8 == 0 && lds >= cols && ldd >= cols);
  if (!q_aligned16(src) || !q_aligned16(dst) || (lds % 8) || (ldd % 8)) return LR_ERR_ALIGN;
  const long long warps = (long long)B * S;
  compact_rows_kernel<<<unsigned((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(src), lds, ord, plan, reinterpret_cast<bf16*>(dst), ldd, B, S, cols);
  return lr_launch_status();
}


extern "C" int lr_gather_rows_bf16(const void* src, int lds, const int* row_index, void* dst, int ldd, int rows,
                                   int cols, void* stream) {
  LR_CHECK_ARG(src && row_index && dst && rows > 0 && cols > 0 && cols % 8 == 0 && lds >= cols && ldd >= cols);
  if (!q_aligned16(src) || !q_aligned16(dst) || (lds % 8) || (ldd % 8)) return LR_ERR_ALIGN;
  gather_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(src), lds, row_index, reinterpret_cast<bf16*>(dst), ldd, rows, cols);
  return lr_launch_status();
}
