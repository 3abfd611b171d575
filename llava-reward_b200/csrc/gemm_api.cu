// lr_gemm_bf16 entry point + the CUDA-core cross-check kernel (LR_GEMM_SIMT).
#include "common.cuh"

namespace lr {

int gemm_tcgen05(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                 const void* bias, const void* R, int ldr, cudaStream_t s);
int gemm_tcgen05_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                      const void* bias, const void* R, int ldr, cudaStream_t s);

int gemm_rope_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const int* pos,
                   const void* cos_tab, const void* sin_tab, int rope_cols, int head_dim, cudaStream_t s);

int gemm_rope_ex_pair(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                      const void* lin_bias, const int* pos, const void* cos_tab, const void* sin_tab, int rope_cols,
                      int head_dim, int epi, cudaStream_t s);

// Verification-only kernel: 32x32 output tile per CTA, fp32 accumulation on CUDA cores, same epilogue math.
// It exists so tests can tell a tcgen05/TMA descriptor bug from an epilogue/packing bug; the engine never
// selects it on its own.
template <int EPI>
__global__ void __launch_bounds__(1024)
gemm_simt_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, bf16* __restrict__ C,
                 int ldc, int M, int N, int K, const bf16* __restrict__ bias, const bf16* __restrict__ R, int ldr,
                 const int* __restrict__ pos = nullptr, int rope_hd = 0) {
  __shared__ float As[32][33], Ws[32][33], Us[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int m = blockIdx.y * 32 + ty;   // output row of this thread
  const int j = blockIdx.x * 32 + tx;   // output column of this thread
  auto wrow = [](int col) { return EPI == LR_EPI_SWIGLU ? (col / 128) * 256 + col % 128 : col; };
  const int lrow_m = blockIdx.y * 32 + ty;            // A row this thread loads
  const int lrow_w = wrow(blockIdx.x * 32 + ty);      // W row this thread loads
  float acc = 0.f, acc_u = 0.f;
  for (int k0 = 0; k0 < K; k0 += 32) {
    As[ty][tx] = (lrow_m < M) ? __bfloat162float(A[size_t(lrow_m) * lda + k0 + tx]) : 0.f;
    Ws[ty][tx] = __bfloat162float(W[size_t(lrow_w) * ldw + k0 + tx]);
    if (EPI == LR_EPI_SWIGLU) Us[ty][tx] = __bfloat162float(W[size_t(lrow_w + 128) * ldw + k0 + tx]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      acc += As[ty][kk] * Ws[tx][kk];
      if (EPI == LR_EPI_SWIGLU) acc_u += As[ty][kk] * Us[tx][kk];
    }
    __syncthreads();
  }
  if (EPI == LR_EPI_ROPE) {  // bias = cos table, R = sin table, ldr = rotated columns, pos/rope_hd extra
    const float xr = bf16_round(acc);
    const float other = __shfl_xor_sync(0xffffffffu, xr, 1);  // the rotation partner is the adjacent column
    if (m >= M) return;
    float out = xr;
    if (j < ldr) {
      const int half = rope_hd >> 1, i = (j % rope_hd) >> 1;
      const float cs = __bfloat162float(bias[size_t(pos[m]) * half + i]), sn = __bfloat162float(R[size_t(pos[m]) * half + i]);
      out = (j & 1) ? bf16_round(xr * cs) + bf16_round(other * sn) : bf16_round(xr * cs) + bf16_round(-other * sn);
    }
    C[size_t(m) * ldc + j] = __float2bfloat16_rn(out);
    return;
  }
  if (m >= M) return;
  float out;
  if (EPI == LR_EPI_SWIGLU) {
    out = epi_swiglu(acc, acc_u);
  } else {
    const float bv = epi_has_bias(EPI) ? __bfloat162float(bias[j]) : 0.f;
    const float rv = epi_has_res(EPI) ? __bfloat162float(R[size_t(m) * ldr + j]) : 0.f;
    out = epi_apply<EPI>(acc, bv, rv);
  }
  C[size_t(m) * ldc + j] = __float2bfloat16_rn(out);
}

template <int EPI>
static int launch_simt(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                       const void* bias, const void* R, int ldr, cudaStream_t s, const int* pos = nullptr,
                       int rope_hd = 0) {
  const int n_out = EPI == LR_EPI_SWIGLU ? N / 2 : N;
  dim3 grid(n_out / 32, (M + 31) / 32), block(32, 32);
  gemm_simt_kernel<EPI><<<grid, block, 0, s>>>(reinterpret_cast<const bf16*>(A), lda, reinterpret_cast<const bf16*>(W),
                                               ldw, reinterpret_cast<bf16*>(C), ldc, M, N, K,
                                               reinterpret_cast<const bf16*>(bias), reinterpret_cast<const bf16*>(R),
                                               ldr, pos, rope_hd);
  return lr_launch_status();
}

static int gemm_simt(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, int epi,
                     const void* bias, const void* R, int ldr, cudaStream_t s) {
  switch (epi) {
    case LR_EPI_NONE: return launch_simt<LR_EPI_NONE>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS: return launch_simt<LR_EPI_BIAS>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_QUICKGELU: return launch_simt<LR_EPI_BIAS_QUICKGELU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_GELU: return launch_simt<LR_EPI_BIAS_GELU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_RESIDUAL: return launch_simt<LR_EPI_RESIDUAL>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_BIAS_RESIDUAL: return launch_simt<LR_EPI_BIAS_RESIDUAL>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    case LR_EPI_SWIGLU: return launch_simt<LR_EPI_SWIGLU>(A, lda, W, ldw, C, ldc, M, N, K, bias, R, ldr, s);
    default: return LR_ERR_BAD_ARG;
  }
}

}  // namespace lr

using namespace lr;

extern "C" int lr_version(void) { return 100; }

extern "C" int lr_device_check(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return static_cast<int>(e);
  }
  return major == 10 ? LR_OK : LR_ERR_UNSUPPORTED;
}

extern "C" int lr_gemm_rope_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N,
                                 int K, const int* position_ids, const void* cos_tab, const void* sin_tab,
                                 int rope_cols, int head_dim, int impl, void* stream) {
  LR_CHECK_ARG(A && W && C && position_ids && cos_tab && sin_tab && M > 0 && N > 0 && K > 0 && K % 64 == 0);
  LR_CHECK_ARG(N % 256 == 0 && rope_cols % 256 == 0 && rope_cols >= 0 && rope_cols <= N && head_dim > 0 &&
               head_dim % 32 == 0 && lda >= K && ldw >= K && ldc >= N);
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  if (mis(A) || mis(W) || mis(C) || mis(cos_tab) || mis(sin_tab) || (lda % 8) || (ldw % 8) || (ldc % 8)) return LR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (impl == LR_GEMM_SIMT)
    return launch_simt<LR_EPI_ROPE>(A, lda, W, ldw, C, ldc, M, N, K, cos_tab, sin_tab, rope_cols, s, position_ids, head_dim);
  if (impl != LR_GEMM_TCGEN05 && impl != LR_GEMM_TCGEN05_PAIR) return LR_ERR_BAD_ARG;
  return gemm_rope_pair(A, lda, W, ldw, C, ldc, M, N, K, position_ids, cos_tab, sin_tab, rope_cols, head_dim, s);
}

extern "C" int lr_gemm_rope_ex_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N,
                                    int K, const void* bias, const int* position_ids, const void* cos_tab,
                                    const void* sin_tab, int rope_cols, int head_dim, int epilogue, void* stream) {
  LR_CHECK_ARG(A && W && C && bias && cos_tab && sin_tab && M > 0 && N > 0 && K > 0 && K % 64 == 0);
  LR_CHECK_ARG(epilogue == LR_EPI_BIAS_ROPE || epilogue == LR_EPI_BIAS_ROPE_F32);
  LR_CHECK_ARG(N % 256 == 0 && rope_cols % 256 == 0 && rope_cols >= 0 && rope_cols <= N && head_dim > 0 &&
               head_dim % 32 == 0 && lda >= K && ldw >= K && ldc >= N);
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  if (mis(A) || mis(W) || mis(C) || mis(bias) || mis(cos_tab) || mis(sin_tab) || (lda % 8) || (ldw % 8) || (ldc % 8))
    return LR_ERR_ALIGN;
  return gemm_rope_ex_pair(A, lda, W, ldw, C, ldc, M, N, K, bias, position_ids, cos_tab, sin_tab, rope_cols, head_dim,
                           epilogue, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int lr_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                            int epilogue, const void* bias, const void* R, int ldr, int impl, void* stream) {
  LR_CHECK_ARG(A && W && C && M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 128 == 0);
  // the RoPE epilogues: lr_gemm_rope_bf16 / lr_gemm_rope_ex_bf16 only
  LR_CHECK_ARG((epilogue >= LR_EPI_NONE && epilogue <= LR_EPI_SWIGLU) || epilogue == LR_EPI_BIAS_SWIGLU);
  if (epi_has_bias(epilogue) || epilogue == LR_EPI_BIAS_SWIGLU) LR_CHECK_ARG(bias != nullptr);
  if (epi_has_res(epilogue)) LR_CHECK_ARG(R != nullptr && ldr >= N);
  if (epi_is_swiglu(epilogue)) LR_CHECK_ARG(N % 256 == 0);
  LR_CHECK_ARG(lda >= K && ldw >= K && ldc >= (epi_is_swiglu(epilogue) ? N / 2 : N));
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) != 0; };
  if (mis(A) || mis(W) || mis(C) || (bias && mis(bias)) || (R && mis(R)) || (lda % 8) || (ldw % 8) || (ldc % 8) ||
      (R && (ldr % 8)))
    return LR_ERR_ALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (epilogue == LR_EPI_BIAS_SWIGLU) {  // CTA-pair kernel only (any M)
    if (impl == LR_GEMM_SIMT || impl == LR_GEMM_TCGEN05_SINGLE) return LR_ERR_BAD_ARG;
    return gemm_tcgen05_pair(A, lda, W, ldw, C, ldc, M, N, K, epilogue, bias, R, ldr, s);
  }
  if (impl == LR_GEMM_SIMT) return gemm_simt(A, lda, W, ldw, C, ldc, M, N, K, epilogue, bias, R, ldr, s);
  if (impl == LR_GEMM_TCGEN05_PAIR || (impl == LR_GEMM_TCGEN05 && N % 256 == 0 && M > 256)) {
    if (N % 256) return LR_ERR_BAD_ARG;
    return gemm_tcgen05_pair(A, lda, W, ldw, C, ldc, M, N, K, epilogue, bias, R, ldr, s);
  }
  if (impl != LR_GEMM_TCGEN05 && impl != LR_GEMM_TCGEN05_SINGLE) return LR_ERR_BAD_ARG;
  return gemm_tcgen05(A, lda, W, ldw, C, ldc, M, N, K, epilogue, bias, R, ldr, s);
}
