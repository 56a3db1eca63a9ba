// Microbenchmark 2: mma.sync m16n8k8 TF32 with the register pattern of the conv kernels (distinct A / B fragments per MMA,
// 3-pass hi/lo split, 6 accumulators per warp), with and without the operand loads / splits around it.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int SPLIT>   // 0: operands straight from shared memory, 1: + trunc split (LOP3 + FADD), 2: + cvt.rna split
__global__ void __launch_bounds__(256, 2) k(float* out, int iters) {
  __shared__ float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = (float)(i % 17) * 0.125f;
  __syncthreads();
  float acc[2][3][4] = {};
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      uint32_t bh[3][2], bl[3][2];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float x = sm[(tap * 6 + nt * 2 + e) * 32 + lane + (it & 1)];
          if (SPLIT == 0) { bh[nt][e] = __float_as_uint(x); bl[nt][e] = __float_as_uint(x) ^ 0x1000; }
          else if (SPLIT == 1) { bh[nt][e] = __float_as_uint(x) & 0xffffe000u; bl[nt][e] = __float_as_uint(x - __uint_as_float(bh[nt][e])); }
          else { asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bh[nt][e]) : "f"(x)); bl[nt][e] = __float_as_uint(x - __uint_as_float(bh[nt][e])); }
        }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t ah[4], al[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = sm[2048 + (tap * 8 + j * 4 + e) * 8 + lane + (it & 1)];
          if (SPLIT == 0) { ah[e] = __float_as_uint(x); al[e] = __float_as_uint(x) ^ 0x1000; }
          else if (SPLIT == 1) { ah[e] = __float_as_uint(x) & 0xffffe000u; al[e] = __float_as_uint(x - __uint_as_float(ah[e])); }
          else { asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ah[e]) : "f"(x)); al[e] = __float_as_uint(x - __uint_as_float(ah[e])); }
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma(acc[j][nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma(acc[j][nt], ah, bl[nt][0], bl[nt][1]);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) mma(acc[j][nt], ah, bh[nt][0], bh[nt][1]);
      }
    }
  }
  float s = 0.f;
  for (int j = 0; j < 2; ++j) for (int nt = 0; nt < 3; ++nt) for (int e = 0; e < 4; ++e) s += acc[j][nt][e];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int SPLIT> void run(int bps, float* out) {
  const int iters = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<SPLIT><<<148 * bps, 256>>>(out, 4);
  cudaEventRecord(e0);
  k<SPLIT><<<148 * bps, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = 148.0 * bps * 8 * iters * 162;
  printf("split=%d blocks/SM=%d: %.3f ms, %.2f cycles/MMA/subcore, %.1f executed TFLOP/s\n", SPLIT, bps, ms,
         ms * 1e-3 * 1.965e9 / (mmas / (148 * 4.0)), mmas * 2048 / ms * 1e-9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2 * 256 * 4);
  run<0>(1, out); run<0>(2, out); run<1>(1, out); run<1>(2, out); run<2>(1, out); run<2>(2, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
