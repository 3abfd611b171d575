// bf16 GEMM  C = epilogue(A . W^T)  on tcgen05 tensor cores, CTA-PAIR variant (cta_group::2).
//
// Two CTAs of a cluster (one TPC) cooperate on a 256(M) x 256(N) output tile: CTA r holds rows [r*128, +128) of A and
// rows [r*128, +128) of the W tile per 64-wide k-block (32 KB per stage instead of 48 KB -> 6 stages, and one third
// less L2->SM traffic per FLOP than the single-CTA 128x256 kernel). The leader CTA's MMA thread issues
// tcgen05.mma.cta_group::2 (M=256, N=256, K=16); each CTA's TMEM receives its own 128 rows x 256 fp32 columns and each
// CTA runs its own 8-warp epilogue + TMA store. Barriers: `full` lives in the leader (both CTAs' TMA loads complete_tx
// on it), `empty` / `tmem_full` are multicast-committed to both CTAs, `tmem_empty` lives in the leader and is arrived
// remotely by the peer's epilogue warps.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap_cache.cuh"

namespace lr {
namespace pair {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kDefaultL2Hints = 18;  // A evict_last | C evict_first << 4 (see l2_policy()): -5 ... -14 % DRAM reads, -1 % time
                                     // (profiles/r02_gemm_l2_hints.txt); LR_GEMM_L2_HINTS overrides
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
constexpr int kStagingBytes = 32 * 128;  // one [32 rows x 64 bf16] tile per epilogue warp

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = kBM * kBK * 2;
  static constexpr int kStageBytesB = (BN / 2) * kBK * 2;  // this CTA's half of the W tile
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = 6;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiWarps * kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) inside CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority hints (kind: 0 normal, 1 evict_first, 2 evict_last). The output tile is written once and never
// read by this kernel, the A slab of a raster group is re-read by every n-tile of the group, W streams once per group.
__device__ __forceinline__ uint64_t l2_policy(int kind) {
  uint64_t p;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const void* tmap, uint32_t leader_bar, int c0,
                                                      int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const void* tmap, const void* smem_src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far have retired) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(uint16_t(3))
      : "memory");
}

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
  const int per_group = group_m * num_n;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int gm = min(group_m, num_m - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, int M, int N, int K,
                    const bf16* __restrict__ bias, const bf16* __restrict__ R, int ldr, int group_m,
                    const int* __restrict__ pos, int rope_hd, const bf16* __restrict__ lin_bias) {
  // the RoPE epilogues reuse the generic slots: bias = cos table, R = sin table (bf16, or fp32 for
  // LR_EPI_BIAS_ROPE_F32), ldr = number of rotated columns; lin_bias = the nn.Linear bias (LR_EPI_BIAS_ROPE*), pos may
  // be NULL (table row = output row)
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  // bits 16.. of group_m carry the L2 hint kinds of the launch: A (2 bits) | W (2 bits) | C (2 bits)
  const int l2_hints = group_m >> 16;
  group_m &= 0xffff;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint8_t* smem_c = smem + kStages * Cfg::kStageBytes;  // 1024-aligned: every stage size is a multiple of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + kEpiWarps * kStagingBytes);
  uint64_t* full_bar = bars;                 // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;      // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;  // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs), 1 = peer
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_m = (M + 2 * kBM - 1) / (2 * kBM);  // 256-row pair tiles
  const int num_n = N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * kEpiWarps);  // one arrive per epilogue warp of BOTH CTAs (used in the leader)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised + TMEM allocated before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_a = l2_policy(l2_hints & 3), pol_b = l2_policy((l2_hints >> 2) & 3);
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);  // bytes of both CTAs
          const uint32_t lbar = map_to_cta(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair_hint(smem_a + stage * Cfg::kStageBytesA, &tma_a, lbar, kb * kBK,
                                m_blk * 2 * kBM + rank * kBM, pol_a);
          tma_load_2d_pair_hint(smem_b + stage * Cfg::kStageBytesB, &tma_b, lbar, kb * kBK,
                                n_blk * BN + rank * (BN / 2), pol_b);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kStageBytesA));
          const uint64_t db = umma_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kStageBytesB));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16_ss_pair(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit_pair(&empty_bar[stage]);  // frees this smem stage in both CTAs once the MMAs above retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tfull_bar[as]);  // accumulator complete (both CTAs)
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int q = warp & 3;      // TMEM lane quarter this warp may access (hardware rule: warp_id % 4)
    const int half = ew >> 2;    // which half of the tile's output columns
    constexpr int kOutN = epi_is_swiglu(EPI) ? BN / 2 : BN;
    constexpr int kChunks = kOutN / 2 / 64;  // 64-column chunks per warp
    static_assert(kChunks >= 1, "tile too narrow for 8 epilogue warps");
    uint8_t* stg = smem_c + ew * kStagingBytes;
    uint8_t* my_row = stg + lane * 128;
    const uint64_t pol_c = l2_policy((l2_hints >> 4) & 3);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row0 = m_blk * 2 * kBM + int(rank) * kBM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col_t = half * (kOutN / 2) + c * 64;  // first output column of this chunk inside the tile
        const int col_g = n_blk * kOutN + col_t;        // ... in C
        uint32_t packed[32];                            // 64 bf16 outputs of this thread's row
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                // two 32-column sub-chunks
          uint32_t acc[32];
          float v[32];
          tmem_ld_32x32(taddr + col_t + hh * 32, acc);
          if constexpr (epi_is_swiglu(EPI)) {
            uint32_t up[32];
            tmem_ld_32x32(taddr + BN / 2 + col_t + hh * 32, up);
            float bg[32], bu[32];
            if constexpr (EPI == LR_EPI_BIAS_SWIGLU) {  // bias packed like the W rows: [gate 128 | up 128] per n-tile
              const uint4* gp = reinterpret_cast<const uint4*>(bias + n_blk * BN + col_t + hh * 32);
              const uint4* up_ = reinterpret_cast<const uint4*>(bias + n_blk * BN + BN / 2 + col_t + hh * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                unpack8_bf16(__ldg(gp + j), bg + j * 8);
                unpack8_bf16(__ldg(up_ + j), bu + j * 8);
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if constexpr (EPI == LR_EPI_BIAS_SWIGLU)
                v[j] = epi_swiglu(__uint_as_float(acc[j]) + bg[j], __uint_as_float(up[j]) + bu[j]);
              else
                v[j] = epi_swiglu(__uint_as_float(acc[j]), __uint_as_float(up[j]));
            }
          } else if constexpr (epi_is_rope(EPI)) {
            const int cg = col_g + hh * 32;
            float lb[32];
            if constexpr (EPI != LR_EPI_ROPE) {
              const uint4* bp = reinterpret_cast<const uint4*>(lin_bias + cg);
#pragma unroll
              for (int j = 0; j < 4; ++j) unpack8_bf16(__ldg(bp + j), lb + j * 8);
            }
            tmem_ld_wait();
            if constexpr (EPI != LR_EPI_ROPE) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + lb[j]);
            }
            if (cg < ldr) {
              const int half = rope_hd >> 1;
              const int p = row_ok ? (pos ? pos[row] : row) : 0;
              const int i0 = (cg % rope_hd) >> 1;  // 16 consecutive rotation pairs start here (cg % 32 == 0)
              float cs[16], sn[16];
              if constexpr (EPI == LR_EPI_BIAS_ROPE_F32) {
                const float4* cp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(bias) + size_t(p) * half + i0);
                const float4* sp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(R) + size_t(p) * half + i0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 c4 = __ldg(cp + j), s4 = __ldg(sp + j);
                  cs[4 * j] = c4.x, cs[4 * j + 1] = c4.y, cs[4 * j + 2] = c4.z, cs[4 * j + 3] = c4.w;
                  sn[4 * j] = s4.x, sn[4 * j + 1] = s4.y, sn[4 * j + 2] = s4.z, sn[4 * j + 3] = s4.w;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {  // fp32 rotation, one rounding (apply_rotary_pos_emb_vision)
                  const float x1 = bf16_round(__uint_as_float(acc[2 * k])), x2 = bf16_round(__uint_as_float(acc[2 * k + 1]));
                  v[2 * k] = __fadd_rn(__fmul_rn(x1, cs[k]), __fmul_rn(-x2, sn[k]));
                  v[2 * k + 1] = __fadd_rn(__fmul_rn(x2, cs[k]), __fmul_rn(x1, sn[k]));
                }
              } else {
                const uint4* cp = reinterpret_cast<const uint4*>(bias + size_t(p) * half + i0);
                const uint4* sp = reinterpret_cast<const uint4*>(R + size_t(p) * half + i0);
                unpack8_bf16(__ldg(cp), cs);
                unpack8_bf16(__ldg(cp + 1), cs + 8);
                unpack8_bf16(__ldg(sp), sn);
                unpack8_bf16(__ldg(sp + 1), sn + 8);
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const float x1 = bf16_round(__uint_as_float(acc[2 * k])), x2 = bf16_round(__uint_as_float(acc[2 * k + 1]));
                  v[2 * k] = bf16_round(x1 * cs[k]) + bf16_round(-x2 * sn[k]);
                  v[2 * k + 1] = bf16_round(x2 * cs[k]) + bf16_round(x1 * sn[k]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
            }
          } else {
            float bv[32], rv[32];
            if constexpr (epi_has_bias(EPI)) {
              const uint4* bp = reinterpret_cast<const uint4*>(bias + col_g + hh * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) unpack8_bf16(__ldg(bp + j), bv + j * 8);
            }
            if constexpr (epi_has_res(EPI)) {
              if (row_ok) {
                const uint4* rp = reinterpret_cast<const uint4*>(R + size_t(row) * ldr + col_g + hh * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) unpack8_bf16(rp[j], rv + j * 8);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) rv[j] = 0.f;
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = epi_apply<EPI>(__uint_as_float(acc[j]), epi_has_bias(EPI) ? bv[j] : 0.f,
                                    epi_has_res(EPI) ? rv[j] : 0.f);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[hh * 16 + j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        // staging tile is free once the previous TMA store has finished reading it
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 8 x 16 B, chunk position XOR (row & 7) = the 128B TMA swizzle
          *reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 4)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_hint(&tma_c, stg, col_g, row0, pol_c);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tempty_bar[as]), 0));
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn2() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}

// 2D bf16 tensor [rows, cols] with row pitch ld elements; box = [64 cols, box_rows], 128B swizzle.
static int make_tmap(CUtensorMap* map, const void* ptr, int rows, int cols, int ld, int box_rows) {
  if (rows <= 0 || cols <= 0) return LR_ERR_BAD_ARG;
  EncodeTiledFn fn = get_encode_fn2();
  if (!fn) return LR_ERR_NO_DRIVER;
  return cached_tmap_bf16(fn, map, ptr, rows, cols, ld, kBK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B) ? LR_OK : LR_ERR_BAD_ARG;
}

static int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev);
  return cached[dev];
}

template <int BN, int EPI>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                       const void* bias, const void* R, int ldr, cudaStream_t stream, const int* pos = nullptr,
                       int rope_hd = 0, const void* lin_bias = nullptr) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb, tc;
  int st = make_tmap(&ta, A, M, K, lda, kBM);
  if (st != LR_OK) return st;
  st = make_tmap(&tb, W, N, K, ldw, BN / 2);
  if (st != LR_OK) return st;
  st = make_tmap(&tc, C, M, epi_is_swiglu(EPI) ? N / 2 : N, ldc, 32);
  if (st != LR_OK) return st;
  auto kern = gemm_pair_kernel<BN, EPI>;
  {  // per-device attribute, set once per (kernel, device)
    static bool attr_done[64] = {};   // one array per instantiation of this launch template = per kernel
    cudaError_t e = ensure_smem_attr(kern, Cfg::kSmemBytes, attr_done);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int num_tiles = ((M + 2 * kBM - 1) / (2 * kBM)) * (N / BN);
  const int pairs = sm_count() / 2;
  const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);
  // m-tiles per rasterisation group: keep the group's A slab (group_m x 128 x K bf16) around 32 MB so it stays
  // L2-resident while the group sweeps all n-tiles; W is then streamed from HBM once per group.
  int group_m = int((32ll << 20) / (int64_t(2 * kBM) * K * 2));
  group_m = group_m < 4 ? 4 : (group_m > 32 ? 32 : group_m);
  if (const char* e = getenv("LR_GEMM_GROUP_M")) {  // raster experiments (tools/gemm_raster_bench.py)
    const int g = atoi(e);
    if (g > 0) group_m = g;
  }
  // L2 eviction hints, 2 bits each for A | W << 2 | C << 4 (0 normal, 1 evict_first, 2 evict_last)
  static const int l2_hints = [] {
    const char* e = getenv("LR_GEMM_L2_HINTS");
    return e ? atoi(e) : kDefaultL2Hints;
  }();
  group_m |= (l2_hints & 63) << 16;
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, M, N, K, reinterpret_cast<const bf16*>(bias),
                                                        reinterpret_cast<const bf16*>(R), ldr, group_m, pos, rope_hd,
                                                        reinterpret_cast<const bf16*>(lin_bias));
  return lr_launch_status();
}

template <int BN>
static int dispatch_epi(int epi, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N,
                        int K, const void* bias, const void* R, int ldr, cudaStream_t s) {
  switch (epi) {
    case LR_EPI_NONE: return launch_gemm<BN, LR_EPI_NONE>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS: return launch_gemm<BN, LR_EPI_BIAS>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_QUICKGELU:
      return launch_gemm<BN, LR_EPI_BIAS_QUICKGELU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_GELU: return launch_gemm<BN, LR_EPI_BIAS_GELU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_RESIDUAL: return launch_gemm<BN, LR_EPI_RESIDUAL>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_RESIDUAL:
      return launch_gemm<BN, LR_EPI_BIAS_RESIDUAL>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    default: return LR_ERR_BAD_ARG;
  }
}

}  // namespace pair

int gemm_rope_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const int* pos,
                   const void* cos_tab, const void* sin_tab, int rope_cols, int head_dim, cudaStream_t s) {
  return pair::launch_gemm<256, LR_EPI_ROPE>(A, lda, W, ldw, C, ldc, M, N, K, cos_tab, sin_tab, rope_cols, s, pos, head_dim);
}

int gemm_rope_ex_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                      const void* lin_bias, const int* pos, const void* cos_tab, const void* sin_tab, int rope_cols,
                      int head_dim, int epi, cudaStream_t s) {
  if (epi == LR_EPI_BIAS_ROPE)
    return pair::launch_gemm<256, LR_EPI_BIAS_ROPE>(A, lda, W, ldw, C, ldc, M, N, K, cos_tab, sin_tab, rope_cols, s, pos,
                                                    head_dim, lin_bias);
  if (epi == LR_EPI_BIAS_ROPE_F32)
    return pair::launch_gemm<256, LR_EPI_BIAS_ROPE_F32>(A, lda, W, ldw, C, ldc, M, N, K, cos_tab, sin_tab, rope_cols, s,
                                                        pos, head_dim, lin_bias);
  return LR_ERR_BAD_ARG;
}

int gemm_tcgen05_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                      const void* bias, const void* R, int ldr, cudaStream_t s) {
  if (N % 256) return LR_ERR_BAD_ARG;
  if (epi == LR_EPI_SWIGLU) return pair::launch_gemm<256, LR_EPI_SWIGLU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
  if (epi == LR_EPI_BIAS_SWIGLU)
    return pair::launch_gemm<256, LR_EPI_BIAS_SWIGLU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
  return pair::dispatch_epi<256>(epi, A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
}

}  // namespace lr
