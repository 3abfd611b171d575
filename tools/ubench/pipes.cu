// Issue-rate micro-benchmark (per SM, per clock) of the instructions of the attention softmax on sm_100a:
// MUFU.EX2, F2FP.BF16.F32.PACK_AB (cvt.rn.bf16x2.f32), FFMA, integer round+PRMT packing, and mixtures.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cycles) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {  // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 1) {  // F2FP pack (two floats -> bf16x2)
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        acc ^= r;
      } else if (MODE == 2) {  // FFMA
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(0.5f));
      } else if (MODE == 3) {  // MUFU + F2FP interleaved 2:1 (the softmax mix)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) {
          uint32_t r;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i - 1]));
          acc ^= r;
        }
      } else if (MODE == 4) {  // MUFU + integer rounding pack 2:1 (add 0x8000, PRMT high halves)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if (i & 1) {
          const uint32_t u0 = __float_as_uint(a[i - 1]) + 0x8000u, u1 = __float_as_uint(a[i]) + 0x8000u;
          acc ^= __byte_perm(u0, u1, 0x7632);
        }
      } else if (MODE == 5) {  // integer rounding pack alone
        const uint32_t u0 = __float_as_uint(a[i]) + 0x8000u, u1 = __float_as_uint(a[(i + 1) & 15]) + 0x8000u;
        acc ^= __byte_perm(u0, u1, 0x7632);
        a[i] += 1.f;
      } else if (MODE == 6) {  // MUFU + FFMA 1:1
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(1.0001f), "f"(-0.5f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      } else if (MODE == 7) {  // cvt.rn.bf16.f32 single (F2F?)
        unsigned short r;
        asm volatile("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(a[i]));
        acc ^= r;
      } else if (MODE == 8) {  // MUFU.EX2 on packed f16x2 (two exponentials per instruction)
        uint32_t& u = reinterpret_cast<uint32_t&>(a[i]);
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
      } else if (MODE == 9) {  // MUFU.EX2 on packed bf16x2
        uint32_t& u = reinterpret_cast<uint32_t&>(a[i]);
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u));
      } else if (MODE == 10) {  // HFMA2 (f16x2)
        uint32_t& u = reinterpret_cast<uint32_t&>(a[i]);
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(0x3c003c00u), "r"(0x38003800u));
      } else if (MODE == 11) {  // HMNMX2 (f16x2 max)
        uint32_t& u = reinterpret_cast<uint32_t&>(a[i]);
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u) : "r"(reinterpret_cast<uint32_t&>(a[(i + 1) & 15])));
      } else if (MODE == 12) {  // cvt.rn.f16x2.f32 pack
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        acc ^= r;
      } else if (MODE == 13) {  // the f16x2 softmax mix per PAIR of elements: pack + max + fma + ex2
        uint32_t r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15]));
        asm volatile("max.f16x2 %0, %0, %1;" : "+r"(acc) : "r"(r));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r) : "r"(0x3c003c00u), "r"(0xb800b800u));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
        acc ^= r;
      } else if (MODE == 14) {  // the f32 softmax mix per PAIR of elements: 2 max + 2 ffma + 2 ex2 + 1 pack
        float x0 = a[i], x1 = a[(i + 1) & 15];
        float m = __uint_as_float(acc);
        asm volatile("max.f32 %0, %0, %1;" : "+f"(m) : "f"(x0));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(m) : "f"(x1));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x0) : "f"(1.0001f), "f"(-0.5f));
        asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x1) : "f"(1.0001f), "f"(-0.5f));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x0), "f"(x1));
        acc = __float_as_uint(m) ^ (r & 1u);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_inner) {
  float* out;
  long long* cyc;
  const int blocks = 148 * 2, threads = 512, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, threads>>>(out, 16, cyc);
  k<MODE><<<blocks, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[148 * 2];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < blocks; ++i) c += h[i];
  c /= blocks;
  // 2 CTAs x 512 threads per SM resident
  const double thread_ops = double(iters) * 16 * ops_per_inner * 2 * threads;
  printf("%-44s %8.1f ops/clk/SM  (%.0f cycles)\n", name, thread_ops / c, c);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 1);
  run<1>("cvt.rn.bf16x2.f32 (F2FP pack), per instr", 1);
  run<7>("cvt.rn.bf16.f32 (single), per instr", 1);
  run<2>("FFMA", 1);
  run<5>("int round + PRMT pack (2 IADD + PRMT + FADD)", 1);
  run<3>("MUFU + F2FP 2:1, per exp", 1);
  run<4>("MUFU + int pack 2:1, per exp", 1);
  run<6>("FFMA + MUFU 1:1, per exp", 1);
  run<8>("ex2.approx.f16x2, per INSTR (2 exps)", 1);
  run<9>("ex2.approx.ftz.bf16x2, per INSTR (2 exps)", 1);
  run<10>("fma.rn.f16x2 (HFMA2), per instr", 1);
  run<11>("max.f16x2 (HMNMX2), per instr", 1);
  run<12>("cvt.rn.f16x2.f32 pack, per instr", 1);
  run<13>("f16x2 softmax mix, per PAIR of elements", 1);
  run<14>("f32 softmax mix, per PAIR of elements", 1);
  printf("cuda status %d\n", (int)cudaGetLastError());
  return 0;
}
