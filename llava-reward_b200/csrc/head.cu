// Reward head: SkipCA on the last-valid-token row, residual + RMSNorm, value head, preference probability.
// The reference computes W_q/W_k/W_v, the S x N_v score matrix and the value head for all S rows and then
// gathers one row per sample (rw_model_general_preference.py:376-386, 407-448); only that row is computed here.
#include <cooperative_groups.h>

#include "common.cuh"

namespace lr {

// scores[b, j] = bf16( bf16(q_b . K_bj) / sqrt(H) ) for j < N_v(b), 0 for N_v(b) <= j < max_nv
// (rows the reference zero-pads: K row = 0 -> score 0). grid (ceil(max_nv/8), B), one warp per K row.
__global__ void __launch_bounds__(256)
skipca_scores_kernel(const bf16* __restrict__ q, int ldq, const bf16* __restrict__ kv, int ldkv,
                     const int* __restrict__ plan, float* __restrict__ scores, int H, int max_nv, float inv_sqrt_d,
                     float pad_score) {
  extern __shared__ __align__(16) uint8_t sc_smem[];
  bf16* sq = reinterpret_cast<bf16*>(sc_smem);
  const int b = blockIdx.y;
  const int row_base = plan[b * LR_PLAN_STRIDE + LR_PLAN_ROW_BASE], nv = plan[b * LR_PLAN_STRIDE + LR_PLAN_NV];
  for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8) stg128(sq + c, ldg128(q + size_t(b) * ldq + c));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= max_nv) return;
  float acc = 0.f;
  if (j < nv) {
    const bf16* kr = kv + size_t(row_base + j) * ldkv;
    for (int c = lane * 8; c < H; c += 256) {
      const uint4 a = *reinterpret_cast<const uint4*>(sq + c), k4 = ldg128(kr + c);
      float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      float2 k0 = unpack_bf16x2(k4.x), k1 = unpack_bf16x2(k4.y), k2 = unpack_bf16x2(k4.z), k3 = unpack_bf16x2(k4.w);
      acc += a0.x * k0.x + a0.y * k0.y + a1.x * k1.x + a1.y * k1.y + a2.x * k2.x + a2.y * k2.y + a3.x * k3.x +
             a3.y * k3.y;
    }
    acc = warp_sum(acc);
  }
  if (lane == 0) scores[size_t(b) * max_nv + j] = (j < nv) ? bf16_round(bf16_round(acc) * inv_sqrt_d) : pad_score;
}

// A CLUSTER of kHeadCluster CTAs per sample (one CTA without cross attention), H/8 threads each (one thread per
// 16-byte chunk of a row): every CTA recomputes the softmax statistics of the sample's max_nv scores (7.7 KB, L2),
// takes a contiguous slice of the V rows and accumulates P.V for it with 8 rows = 128 B per thread in flight; the
// partial sums meet in cluster rank 0 through DSMEM in a fixed order (bit-reproducible), which then does
// y = x + out, RMSNorm, value head. 32 samples -> 256 CTAs of 384 threads (several per SM, one wave) instead of 32
// CTAs: the 378 MB of V rows of a config-2 forward stream at HBM speed instead of from 32 SMs.
constexpr int kHeadMaxThreads = 1024;  // H <= 8192
constexpr int kHeadMaxNv = 4096;
constexpr int kHeadCluster = 8;
constexpr int kHeadSlice = kHeadMaxNv / kHeadCluster + 2;
constexpr int kHeadRows = 8;           // V rows in flight per thread

__global__ void __launch_bounds__(kHeadMaxThreads)
skipca_head_kernel(const float* __restrict__ scores, const bf16* __restrict__ kv, int ldkv,
                   const int* __restrict__ plan, const bf16* __restrict__ x, int ldx, const bf16* __restrict__ ln_w,
                   const bf16* __restrict__ vh_w, bf16* __restrict__ reward, int H, int max_nv, int vhd, float eps) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[32];
  __shared__ float prob[kHeadSlice];
  extern __shared__ __align__(16) uint8_t hd_smem[];
  float* part = reinterpret_cast<float*>(hd_smem);  // [H] partial PV sums of this CTA
  const int ncta = cluster.num_blocks(), rank = cluster.block_rank();
  const int b = blockIdx.x / ncta, tid = threadIdx.x, nthr = blockDim.x;
  const bool own = tid < (H >> 3);     // this thread owns columns [8 tid, 8 tid + 8)
  if (scores) {
    const int row_base = plan[b * LR_PLAN_STRIDE + LR_PLAN_ROW_BASE], nv = plan[b * LR_PLAN_STRIDE + LR_PLAN_NV];
    const float* sc = scores + size_t(b) * max_nv;
    float mx = -INFINITY;
    for (int j = tid; j < max_nv; j += nthr) mx = fmaxf(mx, sc[j]);
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = tid; j < max_nv; j += nthr) sum += __expf(sc[j] - mx);
    sum = block_sum(sum, red);
    const float inv = 1.f / sum;
    // this CTA's slice of the real rows (the zero-padded rows nv <= j < max_nv only feed the denominator)
    const int per = (nv + ncta - 1) / ncta;
    const int j0 = min(rank * per, nv), j1 = min(j0 + per, nv);
    for (int j = j0 + tid; j < j1; j += nthr)
      prob[j - j0] = bf16_round(__expf(sc[j] - mx) * inv);  // softmax output is bf16
    __syncthreads();
    if (own) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const bf16* vbase = kv + size_t(row_base) * ldkv + H + tid * 8;  // V = second half of the [K|V] row
      int j = j0;
      for (; j + kHeadRows <= j1; j += kHeadRows) {
        uint4 u[kHeadRows];
#pragma unroll
        for (int t = 0; t < kHeadRows; ++t) u[t] = ldg128(vbase + size_t(j + t) * ldkv);
#pragma unroll
        for (int t = 0; t < kHeadRows; ++t) {
          const float p = prob[j + t - j0];
          float2 f0 = unpack_bf16x2(u[t].x), f1 = unpack_bf16x2(u[t].y), f2 = unpack_bf16x2(u[t].z),
                 f3 = unpack_bf16x2(u[t].w);
          acc[0] += p * f0.x, acc[1] += p * f0.y, acc[2] += p * f1.x, acc[3] += p * f1.y;
          acc[4] += p * f2.x, acc[5] += p * f2.y, acc[6] += p * f3.x, acc[7] += p * f3.y;
        }
      }
      for (; j < j1; ++j) {
        const uint4 u = ldg128(vbase + size_t(j) * ldkv);
        const float p = prob[j - j0];
        float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        acc[0] += p * f0.x, acc[1] += p * f0.y, acc[2] += p * f1.x, acc[3] += p * f1.y;
        acc[4] += p * f2.x, acc[5] += p * f2.y, acc[6] += p * f3.x, acc[7] += p * f3.y;
      }
      *reinterpret_cast<float4*>(part + tid * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(part + tid * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    cluster.sync();  // every CTA's partial sums are visible cluster-wide
  }
  if (rank == 0) {
    // y = bf16(x + bf16(attn_out)); RMSNorm; value head. Owning threads hold 8 columns each.
    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float ss = 0.f;
    if (own) {
      const uint4 u = ldg128(x + size_t(b) * ldx + tid * 8);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      y[0] = f0.x, y[1] = f0.y, y[2] = f1.x, y[3] = f1.y, y[4] = f2.x, y[5] = f2.y, y[6] = f3.x, y[7] = f3.y;
      if (scores) {
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < ncta; ++r) {  // fixed order: rank 0 .. ncta-1
          const float* rp = cluster.map_shared_rank(part, r);
          const float4 a0 = *reinterpret_cast<const float4*>(rp + tid * 8);
          const float4 a1 = *reinterpret_cast<const float4*>(rp + tid * 8 + 4);
          o[0] += a0.x, o[1] += a0.y, o[2] += a0.z, o[3] += a0.w;
          o[4] += a1.x, o[5] += a1.y, o[6] += a1.z, o[7] += a1.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          y[e] = bf16_round(y[e] + bf16_round(o[e]));
          ss += y[e] * y[e];
        }
      }
    }
    if (scores) {
      ss = block_sum(ss, red);
      const float rstd = rsqrtf(ss / float(H) + eps);
      if (own) {
        const uint4 u = ldg128(ln_w + tid * 8);
        float2 g0 = unpack_bf16x2(u.x), g1 = unpack_bf16x2(u.y), g2 = unpack_bf16x2(u.z), g3 = unpack_bf16x2(u.w);
        const float g[8] = {g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, g3.x, g3.y};
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = bf16_round(g[e] * bf16_round(y[e] * rstd));
      }
    }
    for (int d = 0; d < vhd; ++d) {
      float dot = 0.f;
      if (own) {
        const uint4 u = ldg128(vh_w + size_t(d) * H + tid * 8);
        float2 w0 = unpack_bf16x2(u.x), w1 = unpack_bf16x2(u.y), w2 = unpack_bf16x2(u.z), w3 = unpack_bf16x2(u.w);
        dot = y[0] * w0.x + y[1] * w0.y + y[2] * w1.x + y[3] * w1.y + y[4] * w2.x + y[5] * w2.y + y[6] * w3.x +
              y[7] * w3.y;
      }
      dot = block_sum(dot, red);
      if (tid == 0) reward[size_t(b) * vhd + d] = __float2bfloat16_rn(dot);
    }
  }
  if (scores) cluster.sync();  // rank 0 has finished reading the other CTAs' shared memory
}

// bf16 arithmetic, one rounding per torch op of preference_compute
__global__ void preference_kernel(const bf16* __restrict__ c, const bf16* __restrict__ r, float* __restrict__ prob,
                                  int n, int vhd, int is_gpm, float tau) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float z;
  if (is_gpm && vhd == 2) {
    const float c0 = __bfloat162float(c[2 * i]), c1 = __bfloat162float(c[2 * i + 1]);
    const float r0 = __bfloat162float(r[2 * i]), r1 = __bfloat162float(r[2 * i + 1]);
    z = bf16_round(bf16_round(c0 * r1) - bf16_round(c1 * r0));
  } else {
    z = bf16_round(__bfloat162float(c[size_t(i) * vhd]) - __bfloat162float(r[size_t(i) * vhd]));
  }
  z = bf16_round(z / tau);
  prob[i] = bf16_round(1.f / (1.f + expf(-z)));
}

// ------------------------------------------------------------------------------------------------------------------
// All-rows form of the head (mean_hidden_state pooling, rw_model_general_preference.py:376-386 + 398-406): the S x N_v
// score matrix of every sample comes out of the GEMM kernel in bf16 (torch.bmm output); this kernel turns it into the
// bf16 softmax in place. One warp per row: v_j = bf16(s_j / sqrt(d_k)) for j < n_valid (the batch-wide vision length
// incl. the reference's zero-padded rows, whose scores are exact zeros), p_j = bf16(exp(v_j - max) / sum);
// columns n_valid <= j < n_total (alignment padding of the K dimension of the P.V GEMM) are set to 0.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(bf16* __restrict__ s, int lds, int rows, int n_valid, int n_total, float inv_sqrt_d) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  bf16* r = s + size_t(row) * lds;
  auto load8 = [&](int c, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(r + c);
    const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    const float t[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (c + e < n_valid) ? bf16_round(t[e] * inv_sqrt_d) : -INFINITY;
  };
  float mx = -INFINITY;
  for (int c = lane * 8; c < n_valid; c += 256) {
    float v[8];
    load8(c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[e]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane * 8; c < n_valid; c += 256) {
    float v[8];
    load8(c, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += __expf(v[e] - mx);  // exp(-inf) = 0 for the columns beyond n_valid
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  for (int c = lane * 8; c < n_total; c += 256) {
    float v[8];
    load8(c, v);
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pk[e] = pack_bf16x2(__expf(v[2 * e] - mx) * inv, __expf(v[2 * e + 1] - mx) * inv);
    *reinterpret_cast<uint4*>(r + c) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// out[b, :] = bf16( bf16(sum_s x[b,s,:] * mask[b,s]) / bf16(sum_s mask[b,s]) ), each torch op of the reference's pooling
// rounded to bf16 once (fp32 accumulation inside the sums). grid (H/256, B), 8 warps split the rows of a sample.
__global__ void __launch_bounds__(256)
masked_mean_kernel(const bf16* __restrict__ x, int ldx, const int64_t* __restrict__ mask, bf16* __restrict__ out,
                   int ldo, int S) {
  __shared__ float part[8][256];
  __shared__ int cnt[8];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int n = 0;
  for (int srow = warp; srow < S; srow += 8) {
    if (mask[size_t(b) * S + srow] == 0) continue;
    ++n;
    const uint4 u = ldg128(x + (size_t(b) * S + srow) * ldx + c0);
    const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    acc[0] += f0.x, acc[1] += f0.y, acc[2] += f1.x, acc[3] += f1.y;
    acc[4] += f2.x, acc[5] += f2.y, acc[6] += f3.x, acc[7] += f3.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = acc[e];
  if (lane == 0) cnt[warp] = n;
  __syncthreads();
  const int t = threadIdx.x;
  float tot = 0.f;
  int len = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    tot += part[w][t];
    len += cnt[w];
  }
  const float lens = fmaxf(bf16_round(float(len)), 1e-8f);  // mask.sum(dim=1) is a bf16 tensor, then .clamp(min=1e-8)
  out[size_t(b) * ldo + blockIdx.x * 256 + t] = __float2bfloat16_rn(bf16_round(tot) / lens);
}

}  // namespace lr

using namespace lr;
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int lr_skipca_scores_ex(const void* q, int ldq, const void* kv, int ldkv, const int* plan, float* scores,
                                   int B, int H, int max_nv, float pad_score, void* stream) {
  LR_CHECK_ARG(q && kv && plan && scores && B > 0 && H > 0 && H % 256 == 0 && max_nv > 0);
  if ((ldq % 8) || (ldkv % 8) || !aligned16(q) || !aligned16(kv)) return LR_ERR_ALIGN;
  skipca_scores_kernel<<<dim3((max_nv + 7) / 8, B), 256, H * 2, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(kv), ldkv, plan, scores, H, max_nv,
      1.0f / sqrtf(float(H)), pad_score);
  return lr_launch_status();
}

extern "C" int lr_skipca_scores(const void* q, int ldq, const void* kv, int ldkv, const int* plan, float* scores,
                                int B, int H, int max_nv, void* stream) {
  return lr_skipca_scores_ex(q, ldq, kv, ldkv, plan, scores, B, H, max_nv, 0.f, stream);
}

extern "C" int lr_skipca_head(const float* scores, const void* kv, int ldkv, const int* plan, const void* x, int ldx,
                              const void* ca_ln_w, const void* value_head_w, void* reward, int B, int H, int max_nv,
                              int vhd, float eps, void* stream) {
  // one thread per 8 columns: H <= 8192 (the 4096 / 5120 wide Vicuna decoders of the LLaVA-v1.6 branch included)
  LR_CHECK_ARG(x && value_head_w && reward && B > 0 && H > 0 && H % 8 == 0 && vhd > 0 && H / 8 <= kHeadMaxThreads);
  if (scores) LR_CHECK_ARG(kv && plan && ca_ln_w && max_nv > 0 && max_nv <= kHeadMaxNv);
  if ((ldx % 8) || !aligned16(x) || !aligned16(value_head_w) || (scores && ((ldkv % 8) || !aligned16(kv))))
    return LR_ERR_ALIGN;
  cudaLaunchConfig_t cfg = {};
  const int ncta = scores ? kHeadCluster : 1;
  cfg.gridDim = dim3(B * ncta);
  cfg.blockDim = dim3((H / 8 + 31) / 32 * 32);
  cfg.dynamicSmemBytes = scores ? H * sizeof(float) : 0;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, skipca_head_kernel, scores, reinterpret_cast<const bf16*>(kv), ldkv, plan,
                                     reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(ca_ln_w),
                                     reinterpret_cast<const bf16*>(value_head_w), reinterpret_cast<bf16*>(reward), H,
                                     max_nv, vhd, eps);
  if (e != cudaSuccess) return static_cast<int>(e);
  return lr_launch_status();
}

extern "C" int lr_preference(const void* chosen, const void* reject, float* prob, int n, int vhd, int is_gpm,
                             float tau, void* stream) {
  LR_CHECK_ARG(chosen && reject && prob && n > 0 && vhd > 0 && tau != 0.f);
  preference_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(chosen), reinterpret_cast<const bf16*>(reject), prob, n, vhd, is_gpm, tau);
  return lr_launch_status();
}

extern "C" int lr_softmax_rows_bf16(void* scores, int lds, int rows, int n_valid, int n_total, float inv_sqrt_d,
                                    void* stream) {
  LR_CHECK_ARG(scores && rows > 0 && n_valid > 0 && n_total >= n_valid && n_total % 8 == 0 && lds >= n_total);
  if ((lds % 8) || !aligned16(scores)) return LR_ERR_ALIGN;
  softmax_rows_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<bf16*>(scores), lds, rows, n_valid, n_total, inv_sqrt_d);
  return lr_launch_status();
}

extern "C" int lr_masked_mean_rows_bf16(const void* x, int ldx, const int64_t* attention_mask, void* out, int ldo,
                                        int B, int S, int H, void* stream) {
  LR_CHECK_ARG(x && attention_mask && out && B > 0 && S > 0 && H > 0 && H % 256 == 0 && ldx >= H && ldo >= H);
  if ((ldx % 8) || !aligned16(x)) return LR_ERR_ALIGN;
  masked_mean_kernel<<<dim3(H / 256, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), ldx, attention_mask, reinterpret_cast<bf16*>(out), ldo, S);
  return lr_launch_status();
}
