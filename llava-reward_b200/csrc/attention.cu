// Flash-style prefill attention (online softmax, no S x S materialisation), bf16 in/out, fp32 statistics.
//   head_dim 64, non-causal : CLIP ViT-L/14-336 tower (577 tokens per crop)
//   head_dim 96, causal     : Phi-3 decoder, one contiguous valid run per (left-padded) sequence slot
// Round-1 kernel: 128 query rows per CTA (8 warps x 16 rows), 64-row K/V tiles double-buffered with cp.async,
// QK^T and PV on mma.sync.m16n8k16 (bf16, fp32 accumulate), ldmatrix from padded (conflict-free) smem rows.
#include <cstdlib>

#include "common.cuh"

namespace lr {

constexpr int kAttBM = 128;
constexpr int kAttBN = 64;
constexpr int kAttThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(kAttThreads)
attn_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                bf16* __restrict__ o, int ld_qkv, int ld_o, int rows_per_seq, const int* __restrict__ seq_start,
                const int* __restrict__ seq_len, float scale_log2) {
  constexpr int LDS = HD + 8;      // padded smem row (elements): odd number of 16 B chunks -> conflict-free ldmatrix
  constexpr int CH = HD / 8;       // 16 B chunks per row
  constexpr int KS = HD / 16;      // k-steps of QK^T
  constexpr int DT = HD / 8;       // n8 tiles of the output
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* sQ = reinterpret_cast<bf16*>(att_smem);
  bf16* sK = sQ + kAttBM * LDS;             // [2][64][LDS]
  bf16* sV = sK + 2 * kAttBN * LDS;         // [2][64][LDS]

  const int seq = blockIdx.z, head = blockIdx.y;
  const int m_blk = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int m0 = m_blk * kAttBM;
  const int start = seq_start ? seq_start[seq] : 0;
  const int len = seq_len ? seq_len[seq] : rows_per_seq;
  const int end = start + len;  // valid rows are [start, end) inside the slot
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t slot_row0 = size_t(seq) * rows_per_seq;
  const bf16* qh = q + slot_row0 * ld_qkv + head * HD;
  const bf16* kh = k + slot_row0 * ld_qkv + head * HD;
  const bf16* vh = v + slot_row0 * ld_qkv + head * HD;
  bf16* oh = o + slot_row0 * ld_o + head * HD;

  const int q_lo = max(m0, start), q_hi = min(min(m0 + kAttBM, end), rows_per_seq);
  const bool any_valid = q_lo < q_hi;
  // kv range [start, kv_end)
  const int kv_end = any_valid ? (CAUSAL ? min(end, q_hi) : end) : start;
  const int n_kv_blocks = (kv_end - start + kAttBN - 1) / kAttBN;

  auto load_kv = [&](int blk, int buf) {
    const int r0 = start + blk * kAttBN;
    bf16* dk = sK + buf * kAttBN * LDS;
    bf16* dv = sV + buf * kAttBN * LDS;
    for (int i = threadIdx.x; i < kAttBN * CH; i += kAttThreads) {
      const int r = i / CH, c = i % CH;
      const int row = r0 + r;
      const bool ok = row < kv_end;
      const size_t off = size_t(ok ? row : start) * ld_qkv + c * 8;
      cp_async16(dk + r * LDS + c * 8, kh + off, ok);
      cp_async16(dv + r * LDS + c * 8, vh + off, ok);
    }
  };

  if (any_valid) {
    for (int i = threadIdx.x; i < kAttBM * CH; i += kAttThreads) {
      const int r = i / CH, c = i % CH;
      const int row = m0 + r;
      const bool ok = row >= start && row < q_hi;
      cp_async16(sQ + r * LDS + c * 8, qh + size_t(ok ? row : start) * ld_qkv + c * 8, ok);
    }
    load_kv(0, 0);
  }
  cp_async_commit();

  float oacc[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  float row_m[2] = {-INFINITY, -INFINITY}, row_l[2] = {0.f, 0.f};
  uint32_t qf[KS][4];
  const int wrow0 = warp * 16;                  // this warp's first row inside the tile
  const int qrow_a = m0 + wrow0 + (lane >> 2);  // absolute slot row of accumulator rows c0/c1
  const int qrow_b = qrow_a + 8;                // ... of c2/c3

  for (int blk = 0; blk < n_kv_blocks; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_kv_blocks) load_kv(blk + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldsm_x4(qf[ks], sQ + (wrow0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8);
    }
    const bf16* bk = sK + buf * kAttBN * LDS;
    const bf16* bv = sV + buf * kAttBN * LDS;
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b[4];
        ldsm_x4(b, bk + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + ks * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[2 * np], qf[ks], b[0], b[1]);
        mma_bf16(s[2 * np + 1], qf[ks], b[2], b[3]);
      }
    }
    // scale + mask
    const int kv0 = start + blk * kAttBN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kv0 + t * 8 + (lane & 3) * 2 + (e & 1);
        const int qr = (e < 2) ? qrow_a : qrow_b;
        const bool ok = col < kv_end && (!CAUSAL || col <= qr);
        const float x = ok ? s[t][e] * scale_log2 : -INFINITY;
        s[t][e] = x;
        mx[e >> 1] = fmaxf(mx[e >> 1], x);
      }
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(row_m[h], mx[h]);
      msafe[h] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[h] = exp2f(row_m[h] - msafe[h]);  // row_m = -inf -> 0
      row_m[h] = m_new;
    }
    float ps[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[t][e] - msafe[e >> 1]);
        s[t][e] = p;
        ps[e >> 1] += p;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) row_l[h] = row_l[h] * corr[h] + ps[h];
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      oacc[i][0] *= corr[0];
      oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1];
      oacc[i][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kt][0], s[2 * kt][1]);
      a[1] = pack_bf16x2(s[2 * kt][2], s[2 * kt][3]);
      a[2] = pack_bf16x2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      a[3] = pack_bf16x2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < DT / 2; ++dp) {
        uint32_t b[4];
        ldsm_x4_t(b, bv + (kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8);
        mma_bf16(oacc[2 * dp], a, b[0], b[1]);
        mma_bf16(oacc[2 * dp + 1], a, b[2], b[3]);
      }
    }
    __syncthreads();  // all warps done with buf before it is refilled two iterations later
  }
  cp_async_wait<0>();

  // finalise: full row sums across the 4 lanes of a row, normalise, stage through this warp's sQ rows
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    row_l[h] += __shfl_xor_sync(0xffffffffu, row_l[h], 1);
    row_l[h] += __shfl_xor_sync(0xffffffffu, row_l[h], 2);
  }
  const float inv_a = row_l[0] > 0.f ? 1.f / row_l[0] : 0.f;
  const float inv_b = row_l[1] > 0.f ? 1.f / row_l[1] : 0.f;
  __syncwarp();
  bf16* stage = sQ + wrow0 * LDS;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const int col = i * 8 + (lane & 3) * 2;
    *reinterpret_cast<uint32_t*>(stage + (lane >> 2) * LDS + col) = pack_bf16x2(oacc[i][0] * inv_a, oacc[i][1] * inv_a);
    *reinterpret_cast<uint32_t*>(stage + ((lane >> 2) + 8) * LDS + col) =
        pack_bf16x2(oacc[i][2] * inv_b, oacc[i][3] * inv_b);
  }
  __syncwarp();
  for (int i = lane; i < 16 * CH; i += 32) {
    const int r = i / CH, c = i % CH;
    const int row = m0 + wrow0 + r;
    if (row < rows_per_seq) {
      const bool ok = any_valid && row >= start && row < end;
      uint4 val = ok ? *reinterpret_cast<const uint4*>(stage + r * LDS + c * 8) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(oh + size_t(row) * ld_o + c * 8) = val;
    }
  }
}

template <int HD, bool CAUSAL>
static int launch_attn(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int n_seq,
                       int rows_per_seq, const int* seq_start, const int* seq_len, int n_heads, float scale,
                       cudaStream_t stream) {
  constexpr int smem = (kAttBM + 4 * kAttBN) * (HD + 8) * 2;
  auto kern = attn_fwd_kernel<HD, CAUSAL>;
  {  // per-device attribute; setting it on every launch keeps multi-device processes correct
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  dim3 grid((rows_per_seq + kAttBM - 1) / kAttBM, n_heads, n_seq);
  kern<<<grid, kAttThreads, smem, stream>>>(reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k),
                                            reinterpret_cast<const bf16*>(v), reinterpret_cast<bf16*>(o), ld_qkv, ld_o,
                                            rows_per_seq, seq_start, seq_len, scale * 1.4426950408889634f);
  return lr_launch_status();
}

int attention_tc(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows, int n_seq,
                 int rows_per_seq, const int* seq_base, const int* seq_start, const int* seq_len, int n_heads,
                 int n_kv_heads, int head_dim, int causal, float scale, int split, cudaStream_t s);

int attention_tc_seg(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o, int total_rows,
                     const int* row_lo, const int* row_hi, int n_heads, int head_dim, float scale, cudaStream_t s);

}  // namespace lr

// product configuration at head_dim 128 (attention_tc variant code), chosen by tools/attn128_bench.py on the B200
static constexpr int kHd128Product = 1;  // two tiles per CTA: 5.81 ms vs 7.80 ms (one tile) at 64 x 3057 tokens, 32 heads

extern "C" int lr_attention_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o,
                                 int n_seq, int rows_per_seq, const int* seq_start, const int* seq_len, int n_heads,
                                 int head_dim, int causal, float scale, int impl, void* stream) {
  using namespace lr;
  LR_CHECK_ARG(q && k && v && o && n_seq > 0 && rows_per_seq > 0 && n_heads > 0);
  if ((ld_qkv % 8) || (ld_o % 8) || (reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15)
    return LR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == LR_ATTN_TCGEN05 || impl == LR_ATTN_TCGEN05_SPLIT || impl == LR_ATTN_TCGEN05_2TILE ||
      impl == LR_ATTN_TCGEN05_1TILE || impl == LR_ATTN_TCGEN05_MULTITILE) {
    // attention_tc's variant code: 1 = two tiles per CTA, 2 = split softmax, 3 = one tile per CTA,
    // 4 = the one-tile pipeline walking several query tiles per CTA
    int variant = impl == LR_ATTN_TCGEN05_SPLIT ? 2 : (impl == LR_ATTN_TCGEN05_2TILE ? 1 : 3);
    if (impl == LR_ATTN_TCGEN05_MULTITILE || impl == LR_ATTN_TCGEN05) variant = head_dim == 128 ? kHd128Product : 4;
    if (const char* e = getenv("LR_ATTN_VARIANT")) {  // kernel experiments only (tools/attn_variants.py)
      const int v = atoi(e);
      if (v > 0) variant = v;
    }
    return attention_tc(q, k, v, o, ld_qkv, ld_o, n_seq * rows_per_seq, n_seq, rows_per_seq, nullptr, seq_start, seq_len,
                        n_heads, n_heads, head_dim, causal, scale, variant, s);
  }
  if (impl != LR_ATTN_MMA_SYNC) return LR_ERR_BAD_ARG;
  if (head_dim == 64 && !causal)
    return launch_attn<64, false>(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, scale, s);
  if (head_dim == 96 && causal)
    return launch_attn<96, true>(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, scale, s);
  if (head_dim == 96 && !causal)
    return launch_attn<96, false>(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, scale, s);
  if (head_dim == 64 && causal)
    return launch_attn<64, true>(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, scale, s);
  if (head_dim == 128 && causal)  // Llama decoder of the LLaVA-v1.6 branch
    return launch_attn<128, true>(q, k, v, o, ld_qkv, ld_o, n_seq, rows_per_seq, seq_start, seq_len, n_heads, scale, s);
  return LR_ERR_BAD_ARG;
}


extern "C" int lr_attention_ex_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o,
                                    int total_rows, int n_seq, int max_len, const int* seq_base, const int* seq_start,
                                    const int* seq_len, int n_heads, int n_kv_heads, int head_dim, int causal,
                                    float scale, int impl, void* stream) {
  using namespace lr;
  LR_CHECK_ARG(q && k && v && o && n_seq > 0 && max_len > 0 && n_heads > 0 && n_kv_heads > 0 && total_rows > 0);
  if (seq_base) LR_CHECK_ARG(seq_len != nullptr);
  else LR_CHECK_ARG(total_rows == n_seq * max_len);
  if ((ld_qkv % 8) || (ld_o % 8) || (reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15)
    return LR_ERR_ALIGN;
  int variant = impl == LR_ATTN_TCGEN05_SPLIT ? 2 : (impl == LR_ATTN_TCGEN05_2TILE ? 1 : 3);
  if (impl == LR_ATTN_TCGEN05_MULTITILE || impl == LR_ATTN_TCGEN05) variant = head_dim == 128 ? kHd128Product : 4;
  if (impl != LR_ATTN_TCGEN05 && impl != LR_ATTN_TCGEN05_SPLIT && impl != LR_ATTN_TCGEN05_2TILE &&
      impl != LR_ATTN_TCGEN05_1TILE && impl != LR_ATTN_TCGEN05_MULTITILE)
    return LR_ERR_BAD_ARG;
  return attention_tc(q, k, v, o, ld_qkv, ld_o, total_rows, n_seq, max_len, seq_base, seq_start, seq_len, n_heads,
                      n_kv_heads, head_dim, causal, scale, variant, reinterpret_cast<cudaStream_t>(stream));
}


extern "C" int lr_attention_seg_bf16(const void* q, const void* k, const void* v, void* o, int ld_qkv, int ld_o,
                                     int total_rows, const int* row_lo, const int* row_hi, int n_heads, int head_dim,
                                     float scale, void* stream) {
  using namespace lr;
  LR_CHECK_ARG(q && k && v && o && row_lo && row_hi && total_rows > 0 && n_heads > 0);
  if ((ld_qkv % 8) || (ld_o % 8) || (reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                     reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15)
    return LR_ERR_ALIGN;
  return attention_tc_seg(q, k, v, o, ld_qkv, ld_o, total_rows, row_lo, row_hi, n_heads, head_dim, scale,
                          reinterpret_cast<cudaStream_t>(stream));
}
