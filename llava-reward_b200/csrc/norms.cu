// Memory-bound row kernels: RMSNorm, LayerNorm, CLIP embedding assembly (+pre-LayerNorm), im2col for the
// patch-embedding GEMM. One 128-thread CTA per row, 128-bit loads, the row is held in registers between the
// statistics pass and the write pass (each element is read from HBM exactly once).
#include "common.cuh"

namespace lr {

constexpr int kRowThreads = 128;
constexpr int kMaxChunks = 8;  // 8 x 128 threads x 8 elements = 8192 columns max

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// y = w * bf16(x * rsqrt(mean(x^2) + eps))      (Phi3RMSNorm: normalise in fp32, round, then scale in bf16)
__global__ void __launch_bounds__(kRowThreads)
rmsnorm_kernel(const bf16* __restrict__ x, int ldx, const int* __restrict__ row_index, const bf16* __restrict__ w,
               bf16* __restrict__ y, int ldy, int cols, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const size_t src_row = row_index ? size_t(row_index[row]) : size_t(row);
  const bf16* xr = x + src_row * ldx;
  bf16* yr = y + size_t(row) * ldy;
  const int nchunk = cols >> 3;
  uint4 v[kMaxChunks];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) {
      v[i] = ldg128(xr + c * 8);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / float(cols) + eps);
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) {
      float f[8], g[8];
      unpack8(v[i], f);
      unpack8(ldg128(w + c * 8), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
      stg128(yr + c * 8, pack8(f));
    }
  }
}

// Warp-per-row RMSNorm for narrow rows (cols <= 2048: the Qwen2.5-VL vision tower's 1280): a 2.5 KB row does not fill a
// CTA, and one CTA per row made launch/scheduling overhead the bound (3.1 TB/s measured). NCH x 16 B per lane, shuffle
// reduction only, 8 rows per CTA.
template <int NCH>
__global__ void __launch_bounds__(256)
rmsnorm_warp_kernel(const bf16* __restrict__ x, int ldx, const int* __restrict__ row_index, const bf16* __restrict__ w,
                    bf16* __restrict__ y, int ldy, int rows, int cols, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const size_t src_row = row_index ? size_t(row_index[row]) : size_t(row);
  const bf16* xr = x + src_row * ldx;
  bf16* yr = y + size_t(row) * ldy;
  const int nchunk = cols >> 3;
  uint4 v[NCH];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      v[i] = ldg128(xr + c * 8);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / float(cols) + eps);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + i * 32;
    if (c < nchunk) {
      float f[8], g[8];
      unpack8(v[i], f);
      unpack8(ldg128(w + c * 8), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
      stg128(yr + c * 8, pack8(f));
    }
  }
}

// Shared LayerNorm body: v[] holds this thread's chunks of the (already assembled) row.
__device__ __forceinline__ void layernorm_row(uint4 (&v)[kMaxChunks], int nchunk, int cols, const bf16* w,
                                              const bf16* b, bf16* yr, float eps, float* red) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
  }
  const float mean = block_sum(s, red) / float(cols);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) sq += (f[j] - mean) * (f[j] - mean);
    }
  }
  const float rstd = rsqrtf(block_sum(sq, red) / float(cols) + eps);
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) {
      float f[8], g[8], h[8];
      unpack8(v[i], f);
      unpack8(ldg128(w + c * 8), g);
      unpack8(ldg128(b + c * 8), h);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * g[j] + h[j];
      stg128(yr + c * 8, pack8(f));
    }
  }
}

__global__ void __launch_bounds__(kRowThreads)
layernorm_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w, const bf16* __restrict__ b,
                 bf16* __restrict__ y, int ldy, int cols, float eps) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const int nchunk = cols >> 3;
  uint4 v[kMaxChunks];
#pragma unroll
  for (int i = 0; i < kMaxChunks; ++i) {
    const int c = threadIdx.x + i * kRowThreads;
    if (c < nchunk) v[i] = ldg128(x + row * ldx + c * 8);
  }
  layernorm_row(v, nchunk, cols, w, b, y + row * ldy, eps, red);
}

// Warp-per-row LayerNorm for 1024-column rows (CLIP): 4 x 16 B per lane, shuffle reductions only, 8 rows per CTA.
__global__ void __launch_bounds__(256)
layernorm1024_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w, const bf16* __restrict__ b,
                     bf16* __restrict__ y, int ldy, int rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bf16* xr = x + size_t(row) * ldx;
  float f[4][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unpack8(ldg128(xr + (lane + 32 * i) * 8), f[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += f[i][j];
  }
  const float mean = warp_sum(s) * (1.f / 1024.f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) sq += (f[i][j] - mean) * (f[i][j] - mean);
  const float rstd = rsqrtf(warp_sum(sq) * (1.f / 1024.f) + eps);
  bf16* yr = y + size_t(row) * ldy;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (lane + 32 * i) * 8;
    float g[8], h[8];
    unpack8(ldg128(w + c), g);
    unpack8(ldg128(b + c), h);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[i][j] = (f[i][j] - mean) * rstd * g[j] + h[j];
    stg128(yr + c, pack8(f[i]));
  }
}

// CLIP: tokens[c*577+t] = LN( bf16( (t ? patch[c*576+t-1] : class_emb) + pos[t] ) ), 1024 columns.
__global__ void __launch_bounds__(kRowThreads)
clip_embed_ln_kernel(const bf16* __restrict__ patch, const bf16* __restrict__ cls, const bf16* __restrict__ pos,
                     const bf16* __restrict__ w, const bf16* __restrict__ b, bf16* __restrict__ tokens, float eps) {
  constexpr int D = 1024, T = 577;
  __shared__ float red[32];
  const int crop = blockIdx.x / T, t = blockIdx.x % T;
  const bf16* src = t ? patch + (size_t(crop) * (T - 1) + (t - 1)) * D : cls;
  uint4 v[kMaxChunks];
  {
    float f[8], p[8];
    unpack8(ldg128(src + threadIdx.x * 8), f);
    unpack8(ldg128(pos + size_t(t) * D + threadIdx.x * 8), p);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += p[j];
    v[0] = pack8(f);  // the reference materialises the sum in bf16 before pre_layrnorm
  }
  layernorm_row(v, D / 8, D, w, b, tokens + size_t(blockIdx.x) * D, eps, red);
}

// im2col for Conv2d(3,1024,k=14,s=14): one CTA per (crop, patch-row): stage 3x14x336 pixels in smem as bf16,
// emit 24 rows of 640 (588 + zero pad) bf16 with 16-byte stores.
__global__ void __launch_bounds__(256)
clip_im2col_kernel(const float* __restrict__ pixels, const int* __restrict__ crop_src, bf16* __restrict__ A) {
  constexpr int IMG = 336, P = 14, G = 24, KP = 640, KV = 588;
  __shared__ bf16 tile[3 * P * IMG];
  const int crop = blockIdx.x / G, py = blockIdx.x % G;
  const float* base = pixels + size_t(crop_src[crop]) * 3 * IMG * IMG;
  for (int i = threadIdx.x; i < 3 * P * IMG / 4; i += blockDim.x) {
    const int e = i * 4;
    const int ch = e / (P * IMG), rem = e % (P * IMG), ky = rem / IMG, xx = rem % IMG;
    const float4 f = __ldg(reinterpret_cast<const float4*>(base + (size_t(ch) * IMG + py * P + ky) * IMG + xx));
    bf16* d = tile + e;
    d[0] = __float2bfloat16_rn(f.x);
    d[1] = __float2bfloat16_rn(f.y);
    d[2] = __float2bfloat16_rn(f.z);
    d[3] = __float2bfloat16_rn(f.w);
  }
  __syncthreads();
  bf16* out = A + (size_t(crop) * G * G + size_t(py) * G) * KP;
  for (int i = threadIdx.x; i < G * (KP / 8); i += blockDim.x) {
    const int px = i / (KP / 8), c8 = (i % (KP / 8)) * 8;
    __align__(16) bf16 vals[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = c8 + j;
      bf16 val = __float2bfloat16_rn(0.f);
      if (kk < KV) {
        const int ch = kk / (P * P), r = kk % (P * P), ky = r / P, kx = r % P;
        val = tile[(ch * P + ky) * IMG + px * P + kx];
      }
      vals[j] = val;
    }
    stg128(out + size_t(px) * KP + c8, *reinterpret_cast<uint4*>(vals));
  }
}

}  // namespace lr

using namespace lr;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int lr_rmsnorm_bf16(const void* x, int ldx, const int* row_index, const void* w, void* y, int ldy,
                               int rows, int cols, float eps, void* stream) {
  LR_CHECK_ARG(x && w && y && rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kRowThreads * kMaxChunks * 8);
  if ((ldx % 8) || (ldy % 8) || !aligned16(x) || !aligned16(w) || !aligned16(y)) return LR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cols <= 2048 && rows >= 1024) {  // narrow rows, many of them: one warp per row
    const bf16 *xb = reinterpret_cast<const bf16*>(x), *wb = reinterpret_cast<const bf16*>(w);
    bf16* yb = reinterpret_cast<bf16*>(y);
    if (cols <= 1024)
      rmsnorm_warp_kernel<4><<<(rows + 7) / 8, 256, 0, s>>>(xb, ldx, row_index, wb, yb, ldy, rows, cols, eps);
    else
      rmsnorm_warp_kernel<8><<<(rows + 7) / 8, 256, 0, s>>>(xb, ldx, row_index, wb, yb, ldy, rows, cols, eps);
    return lr_launch_status();
  }
  rmsnorm_kernel<<<rows, kRowThreads, 0, s>>>(
      reinterpret_cast<const bf16*>(x), ldx, row_index, reinterpret_cast<const bf16*>(w), reinterpret_cast<bf16*>(y),
      ldy, cols, eps);
  return lr_launch_status();
}

extern "C" int lr_layernorm_bf16(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int rows,
                                 int cols, float eps, void* stream) {
  LR_CHECK_ARG(x && w && b && y && rows > 0 && cols > 0 && cols % 8 == 0 && cols <= kRowThreads * kMaxChunks * 8);
  if ((ldx % 8) || (ldy % 8) || !aligned16(x) || !aligned16(w) || !aligned16(b) || !aligned16(y)) return LR_ERR_ALIGN;
  if (cols == 1024) {
    layernorm1024_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(w), reinterpret_cast<const bf16*>(b),
        reinterpret_cast<bf16*>(y), ldy, rows, eps);
    return lr_launch_status();
  }
  layernorm_kernel<<<rows, kRowThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(w), reinterpret_cast<const bf16*>(b),
      reinterpret_cast<bf16*>(y), ldy, cols, eps);
  return lr_launch_status();
}

extern "C" int lr_clip_im2col(const float* pixels, const int* crop_src, void* A, int n_crops, void* stream) {
  LR_CHECK_ARG(pixels && crop_src && A && n_crops > 0);
  if (!aligned16(pixels) || !aligned16(A)) return LR_ERR_ALIGN;
  clip_im2col_kernel<<<n_crops * 24, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pixels, crop_src,
                                                                                      reinterpret_cast<bf16*>(A));
  return lr_launch_status();
}

extern "C" int lr_clip_embed_ln(const void* patch, const void* class_emb, const void* pos_emb, const void* ln_w,
                                const void* ln_b, void* tokens, int n_crops, float eps, void* stream) {
  LR_CHECK_ARG(patch && class_emb && pos_emb && ln_w && ln_b && tokens && n_crops > 0);
  if (!aligned16(patch) || !aligned16(class_emb) || !aligned16(pos_emb) || !aligned16(ln_w) || !aligned16(ln_b) ||
      !aligned16(tokens))
    return LR_ERR_ALIGN;
  clip_embed_ln_kernel<<<n_crops * 577, kRowThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(patch), reinterpret_cast<const bf16*>(class_emb),
      reinterpret_cast<const bf16*>(pos_emb), reinterpret_cast<const bf16*>(ln_w), reinterpret_cast<const bf16*>(ln_b),
      reinterpret_cast<bf16*>(tokens), eps);
  return lr_launch_status();
}
