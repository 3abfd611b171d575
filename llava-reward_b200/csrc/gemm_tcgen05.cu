// bf16 GEMM  C = epilogue(A . W^T)  on tcgen05 tensor cores (sm_100a).
//
//   A [M,K] row-major (K-major), W [N,K] row-major (K-major, the nn.Linear layout), fp32 accumulate in TMEM.
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//   warp 0   : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, kStages-deep smem ring)
//   warp 1   : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warp 2   : TMEM allocator (2 accumulator buffers of BN fp32 columns -> the MMAs of tile i+1 overlap
//                              the epilogue of tile i)
//   warps 4-11: epilogue      (two warps per TMEM lane quarter, each owning half of the tile's columns:
//                              tcgen05.ld 32 lanes x 32 columns -> bias/activation/residual -> bf16 ->
//                              128B-swizzled smem staging tile [32 rows x 64 cols] -> TMA store)
//
// Tile order: m fastest inside groups of group_m m-tiles, so the CTAs running concurrently share a handful
// of weight tiles and a ~32 MB slab of A in L2 (A is read from HBM once, W once per group).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap_cache.cuh"

namespace lr {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
constexpr int kStagingBytes = 32 * 128;  // one [32 rows x 64 bf16] tile per epilogue warp

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = kBM * kBK * 2;
  static constexpr int kStageBytesB = BN * kBK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kEpiWarps * kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

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
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_c, int M, int N, int K,
                    const bf16* __restrict__ bias, const bf16* __restrict__ R, int ldr, int group_m) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
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
  const int num_m = (M + kBM - 1) / kBM;
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
      mbar_init(&tempty_bar[i], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kStageBytesA, &tma_a, &full_bar[stage], kb * kBK, m_blk * kBM);
          tma_load_2d(smem_b + stage * Cfg::kStageBytesB, &tma_b, &full_bar[stage], kb * kBK, n_blk * BN);
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
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
            umma_bf16_ss(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete
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
    constexpr int kOutN = (EPI == LR_EPI_SWIGLU) ? BN / 2 : BN;
    constexpr int kChunks = kOutN / 2 / 64;  // 64-column chunks per warp
    static_assert(kChunks >= 1, "tile too narrow for 8 epilogue warps");
    uint8_t* stg = smem_c + ew * kStagingBytes;
    uint8_t* my_row = stg + lane * 128;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row0 = m_blk * kBM + q * 32;
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
          if constexpr (EPI == LR_EPI_SWIGLU) {
            uint32_t up[32];
            tmem_ld_32x32(taddr + BN / 2 + col_t + hh * 32, up);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = epi_swiglu(__uint_as_float(acc[j]), __uint_as_float(up[j]));
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
          tma_store_2d(&tma_c, stg, col_g, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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
  EncodeTiledFn fn = get_encode_fn();
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
                       const void* bias, const void* R, int ldr, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb, tc;
  int st = make_tmap(&ta, A, M, K, lda, kBM);
  if (st != LR_OK) return st;
  st = make_tmap(&tb, W, N, K, ldw, BN);
  if (st != LR_OK) return st;
  st = make_tmap(&tc, C, M, EPI == LR_EPI_SWIGLU ? N / 2 : N, ldc, 32);
  if (st != LR_OK) return st;
  auto kern = gemm_tcgen05_kernel<BN, EPI>;
  {  // per-device attribute, set once per (kernel, device)
    static bool attr_done[64] = {};   // one array per instantiation of this launch template = per kernel
    cudaError_t e = ensure_smem_attr(kern, Cfg::kSmemBytes, attr_done);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int num_tiles = ((M + kBM - 1) / kBM) * (N / BN);
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  // m-tiles per rasterisation group: keep the group's A slab (group_m x 128 x K bf16) around 32 MB so it stays
  // L2-resident while the group sweeps all n-tiles; W is then streamed from HBM once per group.
  int group_m = int((32ll << 20) / (int64_t(kBM) * K * 2));
  group_m = group_m < 8 ? 8 : (group_m > 64 ? 64 : group_m);
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, M, N, K, reinterpret_cast<const bf16*>(bias),
                                                        reinterpret_cast<const bf16*>(R), ldr, group_m);
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

int gemm_tcgen05(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                 const void* bias, const void* R, int ldr, cudaStream_t s) {
  if (epi == LR_EPI_SWIGLU) {
    if (N % 256) return LR_ERR_BAD_ARG;
    return launch_gemm<256, LR_EPI_SWIGLU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
  }
  if (N % 256 == 0) return dispatch_epi<256>(epi, A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
  if (N % 128 == 0) return dispatch_epi<128>(epi, A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
  return LR_ERR_BAD_ARG;
}

}  // namespace lr
