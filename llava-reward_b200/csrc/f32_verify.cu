// fp32 VERIFICATION path (north_star: "1e-4 against the reference's fp32"): the same engine dataflow - token plan,
// packed weight layouts (LoRA K-extension, interleaved q/k rows for the fused RoPE, [gate|up] blocks for SwiGLU),
// HD gather, embedding scatter, packed-row decoder, EOS-row SkipCA head - with every floating-point kernel replaced
// by a plain fp32 CUDA-core kernel that has NO bf16 rounding point. It exists so that the engine's fidelity can be
// shown free of bf16 noise (tests/test_f32_verify_gpu.py: rewards within 1e-4 of the reference's fp32 goldens); it is
// a debug configuration (`RewardEngine(..., precision="fp32")`), orders of magnitude slower than the tcgen05 path and
// never used for scoring. The index / gather kernels are the PRODUCT kernels run byte-wise on fp32 rows
// (lr_gather_rows_bf16 / lr_embed_scatter_bf16 with doubled widths, hd_gather_kernel<2048>).
// This file is compiled WITHOUT --use_fast_math (build.py): expf / erff / division are the IEEE-accurate versions.
#include "common.cuh"

namespace lr {
namespace f32v {

__device__ __forceinline__ float act_quick_gelu(float x) { return x / (1.f + expf(-1.702f * x)); }
__device__ __forceinline__ float act_gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_silu(float x) { return x / (1.f + expf(-x)); }

// C[M,N] = epi(A[M,K] . W[N,K]^T), all fp32, 64x64 tile, 16-deep k steps, 4x4 outputs per thread.
template <int EPI>
__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, float* __restrict__ C, int ldc,
            int M, int N, int K, const float* __restrict__ bias, const float* __restrict__ R, int ldr) {
  __shared__ float sA[16][64 + 1], sB[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      sA[c][r] = (m0 + r < M && k0 + c < K) ? A[size_t(m0 + r) * lda + k0 + c] : 0.f;
      sB[c][r] = (n0 + r < N && k0 + c < K) ? W[size_t(n0 + r) * ldw + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i], b[i] = sB[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if constexpr (epi_has_bias(EPI)) v += bias[n];
      if constexpr (EPI == LR_EPI_BIAS_QUICKGELU) v = act_quick_gelu(v);
      if constexpr (EPI == LR_EPI_BIAS_GELU) v = act_gelu_erf(v);
      if constexpr (epi_has_res(EPI)) v += R[size_t(m) * ldr + n];
      C[size_t(m) * ldc + n] = v;
    }
  }
}

// out[m, blk*128 + i] = up * silu(gate) from raw[m, blk*256 + i] (gate) and raw[m, blk*256 + 128 + i] (up): the packed
// [gate128 | up128] row order of the product's SwiGLU epilogue (weights.py)
__global__ void swiglu_kernel(const float* __restrict__ raw, int ldr, float* __restrict__ out, int ldo, int M, int N2) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(M) * N2) return;
  const int m = int(i / N2), c = int(i % N2), blk = c >> 7, j = c & 127;
  const float g = raw[size_t(m) * ldr + blk * 256 + j], u = raw[size_t(m) * ldr + blk * 256 + 128 + j];
  out[size_t(m) * ldo + c] = u * act_silu(g);
}

// rotary embedding in place on columns [0, rope_cols) of the head-interleaved layout (pair k of a head = columns 2k,
// 2k+1 of that head, see lr_gemm_rope_bf16): x1' = x1 cos - x2 sin, x2' = x2 cos + x1 sin, tables fp32 [n_pos, hd/2]
__global__ void rope_kernel(float* __restrict__ x, int ld, const int* __restrict__ pos, const float* __restrict__ cs,
                            const float* __restrict__ sn, int rows, int rope_cols, int head_dim) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int pairs = rope_cols >> 1;
  if (i >= size_t(rows) * pairs) return;
  const int row = int(i / pairs), pr = int(i % pairs), col = pr * 2, k = (col % head_dim) >> 1, half = head_dim >> 1;
  const int p = pos ? pos[row] : row;
  float* q = x + size_t(row) * ld + col;
  const float x1 = q[0], x2 = q[1], c = cs[size_t(p) * half + k], s = sn[size_t(p) * half + k];
  q[0] = x1 * c - x2 * s;
  q[1] = x2 * c + x1 * s;
}

// one warp per row
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ row_index, const float* __restrict__ w,
               float* __restrict__ y, int ldy, int rows, int cols, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + size_t(row_index ? row_index[row] : row) * ldx;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) ss += xr[c] * xr[c];
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / float(cols) + eps);
  for (int c = lane; c < cols; c += 32) y[size_t(row) * ldy + c] = w[c] * (xr[c] * rstd);
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ y, int ldy, int rows, int cols, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + size_t(row) * ldx;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
  const float mean = warp_sum(s) / float(cols);
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) v += (xr[c] - mean) * (xr[c] - mean);
  const float rstd = rsqrtf(warp_sum(v) / float(cols) + eps);
  for (int c = lane; c < cols; c += 32) y[size_t(row) * ldy + c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

// im2col of Conv2d(3, 1024, k=14, s=14) in the (ch, ky, kx) order of the flattened conv weight, K 588 -> 640 zero pad
__global__ void clip_im2col_kernel(const float* __restrict__ pixels, const int* __restrict__ crop_src,
                                   float* __restrict__ A, int n_rows) {
  constexpr int IMG = 336, P = 14, G = 24, KP = 640, KV = 588;
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(n_rows) * KP) return;
  const int row = int(i / KP), kk = int(i % KP);
  float v = 0.f;
  if (kk < KV) {
    const int crop = row / (G * G), py = (row / G) % G, px = row % G;
    const int ch = kk / (P * P), ky = (kk / P) % P, kx = kk % P;
    v = pixels[(size_t(crop_src[crop]) * 3 + ch) * IMG * IMG + size_t(py * P + ky) * IMG + px * P + kx];
  }
  A[i] = v;
}

// tokens[crop, t] = pre_layernorm(t ? patch[crop, t-1] : cls) + pos[t]); one warp per token row, D = 1024
__global__ void __launch_bounds__(256)
clip_embed_ln_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                     const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ tokens, int rows,
                     float eps) {
  constexpr int D = 1024, T = 577;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int crop = row / T, t = row % T;
  const float* src = t ? patch + (size_t(crop) * (T - 1) + (t - 1)) * D : cls;
  float xv[D / 32];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < D / 32; ++j) {
    xv[j] = src[lane + 32 * j] + pos[size_t(t) * D + lane + 32 * j];
    s += xv[j];
  }
  const float mean = warp_sum(s) / float(D);
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < D / 32; ++j) v += (xv[j] - mean) * (xv[j] - mean);
  const float rstd = rsqrtf(warp_sum(v) / float(D) + eps);
#pragma unroll
  for (int j = 0; j < D / 32; ++j)
    tokens[size_t(row) * D + lane + 32 * j] = (xv[j] - mean) * rstd * w[lane + 32 * j] + b[lane + 32 * j];
}

// softmax attention, one warp per (query row, head): exact two-pass softmax in fp32 over the valid keys.
// slot layout (seq_base == NULL): sequence s owns rows [s * rows_per_seq, +rows_per_seq), valid run [start, start+len);
// packed layout: sequence s owns rows [seq_base[s], + seq_len[s]). Rows outside the valid run are written as zeros.
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                 float* __restrict__ o, int ld_qkv, int ld_o, int rows_per_seq, const int* __restrict__ seq_base,
                 const int* __restrict__ seq_start, const int* __restrict__ seq_len, int n_heads, int kv_group,
                 int causal, float scale) {
  extern __shared__ float sc[];   // [4 warps][rows_per_seq]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + warp, head = blockIdx.y, seq = blockIdx.z;
  if (r >= rows_per_seq) return;
  const int start = (seq_start && !seq_base) ? seq_start[seq] : 0;
  const int len = seq_len ? seq_len[seq] : rows_per_seq;
  const int row0 = seq_base ? seq_base[seq] : seq * rows_per_seq;
  if (seq_base && r >= len) return;           // packed: rows beyond this sequence belong to the next one
  float* orow = o + size_t(row0 + r) * ld_o + head * HD;
  if (r < start || r >= start + len) {
    for (int d = lane; d < HD; d += 32) orow[d] = 0.f;
    return;
  }
  const int kend = causal ? r + 1 : start + len;
  const float* qr = q + size_t(row0 + r) * ld_qkv + head * HD;
  const int kvh = head / kv_group;
  float* s = sc + size_t(warp) * rows_per_seq;
  float mx = -INFINITY;
  for (int j = start + lane; j < kend; j += 32) {
    const float* kr = k + size_t(row0 + j) * ld_qkv + kvh * HD;
    float dot = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) dot = fmaf(qr[d], kr[d], dot);
    dot *= scale;
    s[j] = dot;
    mx = fmaxf(mx, dot);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = start + lane; j < kend; j += 32) {
    const float e = expf(s[j] - mx);
    s[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int d = lane; d < HD; d += 32) {
    float acc = 0.f;
    for (int j = start; j < kend; ++j) acc = fmaf(s[j], v[size_t(row0 + j) * ld_qkv + kvh * HD + d], acc);
    orow[d] = acc * inv;
  }
}

// SkipCA on the EOS row: scores over max_nv rows (rows >= N_v(b) are the reference's zero-padded rows: score 0)
__global__ void __launch_bounds__(256)
skipca_scores_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kv, int ldkv,
                     const int* __restrict__ plan, float* __restrict__ scores, int H, int max_nv, float inv_sqrt_d) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= max_nv) return;
  const int row_base = plan[b * LR_PLAN_STRIDE + LR_PLAN_ROW_BASE], nv = plan[b * LR_PLAN_STRIDE + LR_PLAN_NV];
  float acc = 0.f;
  if (j < nv)
    for (int c = lane; c < H; c += 32) acc = fmaf(q[size_t(b) * ldq + c], kv[size_t(row_base + j) * ldkv + c], acc);
  acc = warp_sum(acc);
  if (lane == 0) scores[size_t(b) * max_nv + j] = (j < nv) ? acc * inv_sqrt_d : 0.f;
}

// one CTA per sample: softmax over max_nv scores, out = P V, y = x + out, RMSNorm (ln_w), value head; without
// scores: reward = x . value_head^T
__global__ void __launch_bounds__(1024)
skipca_head_kernel(const float* __restrict__ scores, const float* __restrict__ kv, int ldkv,
                   const int* __restrict__ plan, const float* __restrict__ x, int ldx, const float* __restrict__ ln_w,
                   const float* __restrict__ vh_w, float* __restrict__ reward, int H, int max_nv, int vhd, float eps) {
  __shared__ float red[32];
  extern __shared__ float prob[];   // [max_nv] then [H]
  float* y = prob + max_nv;
  const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
  for (int c = tid; c < H; c += nthr) y[c] = x[size_t(b) * ldx + c];
  if (scores) {
    const int row_base = plan[b * LR_PLAN_STRIDE + LR_PLAN_ROW_BASE], nv = plan[b * LR_PLAN_STRIDE + LR_PLAN_NV];
    const float* sc = scores + size_t(b) * max_nv;
    float mx = -INFINITY;
    for (int j = tid; j < max_nv; j += nthr) mx = fmaxf(mx, sc[j]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = tid; j < max_nv; j += nthr) {
      const float e = expf(sc[j] - mx);
      prob[j] = e;
      sum += e;
    }
    sum = block_sum(sum, red);
    __syncthreads();
    const float inv = 1.f / sum;
    for (int c = tid; c < H; c += nthr) {
      float acc = 0.f;
      for (int j = 0; j < nv; ++j) acc = fmaf(prob[j], kv[size_t(row_base + j) * ldkv + H + c], acc);
      y[c] += acc * inv;
    }
    __syncthreads();
    float ss = 0.f;
    for (int c = tid; c < H; c += nthr) ss += y[c] * y[c];
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / float(H) + eps);
    __syncthreads();
    for (int c = tid; c < H; c += nthr) y[c] = ln_w[c] * (y[c] * rstd);
  }
  __syncthreads();
  for (int d = 0; d < vhd; ++d) {
    float dot = 0.f;
    for (int c = tid; c < H; c += nthr) dot = fmaf(y[c], vh_w[size_t(d) * H + c], dot);
    dot = block_sum(dot, red);
    if (tid == 0) reward[size_t(b) * vhd + d] = dot;
  }
}

__global__ void preference_kernel(const float* __restrict__ c, const float* __restrict__ r, float* __restrict__ prob,
                                  int n, int vhd, int is_gpm, float tau) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z;
  if (is_gpm && vhd == 2) z = c[2 * i] * r[2 * i + 1] - c[2 * i + 1] * r[2 * i];
  else z = c[size_t(i) * vhd] - r[size_t(i) * vhd];
  prob[i] = 1.f / (1.f + expf(-z / tau));
}

}  // namespace f32v
}  // namespace lr

using namespace lr;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define F(p) reinterpret_cast<const float*>(p)
#define FM(p) reinterpret_cast<float*>(p)

extern "C" int lr_f32_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                           int epilogue, const void* bias, const void* R, int ldr, void* stream) {
  LR_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K && ldc >= N);
  if (epi_has_bias(epilogue)) LR_CHECK_ARG(bias != nullptr);
  if (epi_has_res(epilogue)) LR_CHECK_ARG(R != nullptr && ldr >= N);
  const dim3 grid((N + 63) / 64, (M + 63) / 64);
#define LR_F32_GEMM_CASE(E)                                                                                     \
  case E:                                                                                                       \
    f32v::gemm_kernel<E><<<grid, 256, 0, ST(stream)>>>(F(A), lda, F(W), ldw, FM(C), ldc, M, N, K, F(bias), F(R), ldr); \
    break;
  switch (epilogue) {
    LR_F32_GEMM_CASE(LR_EPI_NONE)
    LR_F32_GEMM_CASE(LR_EPI_BIAS)
    LR_F32_GEMM_CASE(LR_EPI_BIAS_QUICKGELU)
    LR_F32_GEMM_CASE(LR_EPI_BIAS_GELU)
    LR_F32_GEMM_CASE(LR_EPI_RESIDUAL)
    LR_F32_GEMM_CASE(LR_EPI_BIAS_RESIDUAL)
    default: return LR_ERR_BAD_ARG;
  }
#undef LR_F32_GEMM_CASE
  return lr_launch_status();
}

extern "C" int lr_f32_swiglu(const void* raw, int ldr, void* out, int ldo, int M, int N, void* stream) {
  LR_CHECK_ARG(raw && out && M > 0 && N > 0 && N % 256 == 0 && ldr >= N && ldo >= N / 2);
  const size_t n = size_t(M) * (N / 2);
  f32v::swiglu_kernel<<<unsigned((n + 255) / 256), 256, 0, ST(stream)>>>(F(raw), ldr, FM(out), ldo, M, N / 2);
  return lr_launch_status();
}

extern "C" int lr_f32_rope(void* x, int ld, const int* position_ids, const void* cos_tab, const void* sin_tab, int rows,
                           int rope_cols, int head_dim, void* stream) {
  LR_CHECK_ARG(x && cos_tab && sin_tab && rows > 0 && rope_cols > 0 && head_dim > 0 && head_dim % 2 == 0 &&
               rope_cols % head_dim == 0 && ld >= rope_cols);
  const size_t n = size_t(rows) * (rope_cols / 2);
  f32v::rope_kernel<<<unsigned((n + 255) / 256), 256, 0, ST(stream)>>>(FM(x), ld, position_ids, F(cos_tab), F(sin_tab),
                                                                      rows, rope_cols, head_dim);
  return lr_launch_status();
}

extern "C" int lr_f32_rmsnorm(const void* x, int ldx, const int* row_index, const void* w, void* y, int ldy, int rows,
                              int cols, float eps, void* stream) {
  LR_CHECK_ARG(x && w && y && rows > 0 && cols > 0);
  f32v::rmsnorm_kernel<<<(rows + 7) / 8, 256, 0, ST(stream)>>>(F(x), ldx, row_index, F(w), FM(y), ldy, rows, cols, eps);
  return lr_launch_status();
}

extern "C" int lr_f32_layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows,
                                int cols, float eps, void* stream) {
  LR_CHECK_ARG(x && w && b && y && rows > 0 && cols > 0);
  f32v::layernorm_kernel<<<(rows + 7) / 8, 256, 0, ST(stream)>>>(F(x), ldx, F(w), F(b), FM(y), ldy, rows, cols, eps);
  return lr_launch_status();
}

extern "C" int lr_f32_clip_im2col(const float* pixels, const int* crop_src, void* A, int n_crops, void* stream) {
  LR_CHECK_ARG(pixels && crop_src && A && n_crops > 0);
  const int rows = n_crops * 576;
  const size_t n = size_t(rows) * 640;
  f32v::clip_im2col_kernel<<<unsigned((n + 255) / 256), 256, 0, ST(stream)>>>(pixels, crop_src, FM(A), rows);
  return lr_launch_status();
}

extern "C" int lr_f32_clip_embed_ln(const void* patch, const void* class_emb, const void* pos_emb, const void* ln_w,
                                    const void* ln_b, void* tokens, int n_crops, float eps, void* stream) {
  LR_CHECK_ARG(patch && class_emb && pos_emb && ln_w && ln_b && tokens && n_crops > 0);
  const int rows = n_crops * 577;
  f32v::clip_embed_ln_kernel<<<(rows + 7) / 8, 256, 0, ST(stream)>>>(F(patch), F(class_emb), F(pos_emb), F(ln_w), F(ln_b),
                                                                    FM(tokens), rows, eps);
  return lr_launch_status();
}

extern "C" int lr_f32_attention(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int n_seq,
                                int rows_per_seq, const int* seq_base, const int* seq_start, const int* seq_len,
                                int n_heads, int n_kv_heads, int head_dim, int causal, float scale, void* stream) {
  LR_CHECK_ARG(q && k && v && o && n_seq > 0 && rows_per_seq > 0 && n_heads > 0 && n_kv_heads > 0 &&
               n_heads % n_kv_heads == 0 && rows_per_seq <= 12000);
  if (seq_base) LR_CHECK_ARG(seq_len != nullptr);
  const dim3 grid((rows_per_seq + 3) / 4, n_heads, n_seq);
  const size_t smem = size_t(4) * rows_per_seq * sizeof(float);
#define LR_F32_ATTN_CASE(HD)                                                                                         \
  if (head_dim == HD) {                                                                                              \
    auto kern = f32v::attention_kernel<HD>;                                                                          \
    if (smem > 48 * 1024) {                                                                                          \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));            \
      if (e != cudaSuccess) return static_cast<int>(e);                                                              \
    }                                                                                                                \
    kern<<<grid, 128, smem, ST(stream)>>>(F(q), F(k), F(v), FM(o), ld_qkv, ld_o, rows_per_seq, seq_base, seq_start,  \
                                          seq_len, n_heads, n_heads / n_kv_heads, causal, scale);                    \
    return lr_launch_status();                                                                                       \
  }
  LR_F32_ATTN_CASE(64)
  LR_F32_ATTN_CASE(96)
  LR_F32_ATTN_CASE(128)
#undef LR_F32_ATTN_CASE
  return LR_ERR_BAD_ARG;
}

extern "C" int lr_f32_skipca_scores(const void* q, int ldq, const void* kv, int ldkv, const int* plan, void* scores,
                                    int B, int H, int max_nv, void* stream) {
  LR_CHECK_ARG(q && kv && plan && scores && B > 0 && H > 0 && max_nv > 0);
  f32v::skipca_scores_kernel<<<dim3((max_nv + 7) / 8, B), 256, 0, ST(stream)>>>(F(q), ldq, F(kv), ldkv, plan, FM(scores), H,
                                                                                max_nv, 1.0f / sqrtf(float(H)));
  return lr_launch_status();
}

extern "C" int lr_f32_skipca_head(const void* scores, const void* kv, int ldkv, const int* plan, const void* x, int ldx,
                                  const void* ca_ln_w, const void* value_head_w, void* reward, int B, int H, int max_nv,
                                  int vhd, float eps, void* stream) {
  LR_CHECK_ARG(x && value_head_w && reward && B > 0 && H > 0 && vhd > 0);
  if (scores) LR_CHECK_ARG(kv && plan && ca_ln_w && max_nv > 0);
  const size_t smem = (size_t(scores ? max_nv : 0) + H) * sizeof(float);
  auto kern = f32v::skipca_head_kernel;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  kern<<<B, 1024, smem, ST(stream)>>>(F(scores), F(kv), ldkv, plan, F(x), ldx, F(ca_ln_w), F(value_head_w), FM(reward), H,
                                      scores ? max_nv : 0, vhd, eps);
  return lr_launch_status();
}

extern "C" int lr_f32_preference(const void* chosen, const void* reject, void* prob, int n, int vhd, int is_gpm, float tau,
                                 void* stream) {
  LR_CHECK_ARG(chosen && reject && prob && n > 0 && vhd > 0 && tau != 0.f);
  f32v::preference_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(F(chosen), F(reject), FM(prob), n, vhd, is_gpm, tau);
  return lr_launch_status();
}
