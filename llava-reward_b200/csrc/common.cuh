// Shared device helpers: bf16 packing, warp/block reductions, status codes.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/llava_reward_b200.h"

namespace lr {

typedef __nv_bfloat16 bf16;

#define LR_CHECK_ARG(cond) \
  do {                     \
    if (!(cond)) return LR_ERR_BAD_ARG; \
  } while (0)

inline int lr_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? LR_OK : static_cast<int>(e);
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// 128-bit global access helpers
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

// ---- GEMM epilogue math shared by the tcgen05 and SIMT kernels -------------------------------
// All variants first round the fp32 accumulator (+bias) to bf16, because the reference
// materialises the nn.Linear output in bf16 before the activation / residual add.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(z) = 0.5 + 0.5 tanh(z/2): one MUFU op instead of ex2 + rcp (the epilogues are MUFU/issue bound)
__device__ __forceinline__ float epi_quick_gelu(float x) {  // x * sigmoid(1.702 x)
  return x * fmaf(0.5f, tanh_approx(0.851f * x), 0.5f);
}
__device__ __forceinline__ float epi_gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float epi_silu(float x) { return x * fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

template <int EPI>
__device__ __forceinline__ float epi_apply(float acc, float bias, float res) {
  if constexpr (EPI == LR_EPI_NONE) return acc;
  if constexpr (EPI == LR_EPI_BIAS) return acc + bias;
  if constexpr (EPI == LR_EPI_BIAS_QUICKGELU) return epi_quick_gelu(bf16_round(acc + bias));
  if constexpr (EPI == LR_EPI_BIAS_GELU) return epi_gelu_erf(bf16_round(acc + bias));
  if constexpr (EPI == LR_EPI_RESIDUAL) return bf16_round(acc) + res;
  if constexpr (EPI == LR_EPI_BIAS_RESIDUAL) return bf16_round(acc + bias) + res;
  return acc;
}
__device__ __forceinline__ float epi_swiglu(float gate_acc, float up_acc) {
  return bf16_round(up_acc) * bf16_round(epi_silu(bf16_round(gate_acc)));
}

__host__ __device__ constexpr bool epi_is_swiglu(int e) { return e == LR_EPI_SWIGLU || e == LR_EPI_BIAS_SWIGLU; }
__host__ __device__ constexpr bool epi_is_rope(int e) {
  return e == LR_EPI_ROPE || e == LR_EPI_BIAS_ROPE || e == LR_EPI_BIAS_ROPE_F32;
}
__host__ __device__ constexpr bool epi_has_bias(int e) {
  return e == LR_EPI_BIAS || e == LR_EPI_BIAS_QUICKGELU || e == LR_EPI_BIAS_GELU || e == LR_EPI_BIAS_RESIDUAL;
}
__host__ __device__ constexpr bool epi_has_res(int e) { return e == LR_EPI_RESIDUAL || e == LR_EPI_BIAS_RESIDUAL; }

}  // namespace lr
