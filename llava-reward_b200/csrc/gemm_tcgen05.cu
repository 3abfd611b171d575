// bf16 GEMM  C = epilogue(A . W^T)  on tcgen05 tensor cores (sm_100a).
//
//   A [M,K] row-major (K-major), W [N,K] row-major (K-major, the nn.Linear layout), fp32 accumulate in TMEM.
//
// One persistent CTA per SM, 256 threads, warp-specialised:
//   warp 0  : TMA producer   (cp.async.bulk.tensor 2D, 128B swizzle, kStages-deep smem ring)
//   warp 1  : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers)
//   warp 2  : TMEM allocator (2 accumulator buffers of BN fp32 columns -> MMA of tile i+1 overlaps
//                             the epilogue of tile i)
//   warps 4-7: epilogue      (tcgen05.ld 32 lanes x 32 columns -> bias/activation/residual -> bf16 -> global)
//
// Tile order: m fastest inside groups of kGroupM m-tiles, so the CTAs running concurrently share a handful
// of weight tiles and a 16 x 128-row slab of A in L2 (A is read from HBM once, W stays L2 resident).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace lr {

constexpr int kBM = 128;
constexpr int kBK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kGroupM = 16;
constexpr int kGemmThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = kBM * kBK * 2;
  static constexpr int kStageBytesB = BN * kBK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * num_n;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gm = min(kGroupM, num_m - first_m);
  const int r = tile - g * per_group;
  m_blk = first_m + r % gm;
  n_blk = r / gm;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    bf16* __restrict__ C, int ldc, int M, int N, int K, const bf16* __restrict__ bias,
                    const bf16* __restrict__ R, int ldr) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kStageBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
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
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
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
        tile_coords(tile, num_m, num_n, m_blk, n_blk);
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
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    constexpr int kOutN = (EPI == LR_EPI_SWIGLU) ? BN / 2 : BN;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, m_blk, n_blk);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int row = m_blk * kBM + q * 32 + lane;
      const bool row_ok = row < M;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
      bf16* crow = C + size_t(row) * ldc + size_t(n_blk) * kOutN;
      const bf16* rrow = epi_has_res(EPI) ? (R + size_t(row) * ldr + size_t(n_blk) * kOutN) : nullptr;
#pragma unroll 1
      for (int c = 0; c < kOutN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32(taddr + c * 32, acc);
        float v[32];
        if constexpr (EPI == LR_EPI_SWIGLU) {
          uint32_t up[32];
          tmem_ld_32x32(taddr + BN / 2 + c * 32, up);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = epi_swiglu(__uint_as_float(acc[j]), __uint_as_float(up[j]));
        } else {
          tmem_ld_wait();
          float bv[32], rv[32];
          if constexpr (epi_has_bias(EPI)) {
            const uint4* bp = reinterpret_cast<const uint4*>(bias + size_t(n_blk) * BN + c * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u = __ldg(bp + j);
              float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              bv[j * 8 + 0] = f0.x, bv[j * 8 + 1] = f0.y, bv[j * 8 + 2] = f1.x, bv[j * 8 + 3] = f1.y;
              bv[j * 8 + 4] = f2.x, bv[j * 8 + 5] = f2.y, bv[j * 8 + 6] = f3.x, bv[j * 8 + 7] = f3.y;
            }
          }
          if constexpr (epi_has_res(EPI)) {
            if (row_ok) {
              const uint4* rp = reinterpret_cast<const uint4*>(rrow + c * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 u = rp[j];
                float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                rv[j * 8 + 0] = f0.x, rv[j * 8 + 1] = f0.y, rv[j * 8 + 2] = f1.x, rv[j * 8 + 3] = f1.y;
                rv[j * 8 + 4] = f2.x, rv[j * 8 + 5] = f2.y, rv[j * 8 + 6] = f3.x, rv[j * 8 + 7] = f3.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) rv[j] = 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j)
            v[j] = epi_apply<EPI>(__uint_as_float(acc[j]), epi_has_bias(EPI) ? bv[j] : 0.f, epi_has_res(EPI) ? rv[j] : 0.f);
        }
        if (row_ok) {
          uint4* cp = reinterpret_cast<uint4*>(crow + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]);
            u.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
            u.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]);
            u.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
            cp[j] = u;
          }
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
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return LR_ERR_NO_DRIVER;
  cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? LR_OK : LR_ERR_BAD_ARG;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int BN, int EPI>
static int launch_gemm(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                       const void* bias, const void* R, int ldr, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  int st = make_tmap(&ta, A, M, K, lda, kBM);
  if (st != LR_OK) return st;
  st = make_tmap(&tb, W, N, K, ldw, BN);
  if (st != LR_OK) return st;
  auto kern = gemm_tcgen05_kernel<BN, EPI>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  const int num_tiles = ((M + kBM - 1) / kBM) * (N / BN);
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, reinterpret_cast<bf16*>(C), ldc, M, N, K,
                                                        reinterpret_cast<const bf16*>(bias),
                                                        reinterpret_cast<const bf16*>(R), ldr);
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
