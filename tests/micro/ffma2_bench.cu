// Micro-benchmark (diagnostic): scalar FFMA vs packed FFMA2 issue rate on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, x.x, y.x);
        acc[i].y = fmaf(acc[i].y, x.y, y.y);
      } else {
        acc[i] = __ffma2_rn(acc[i], x, y);
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 256>>>(out, iters, 0.999f, 0.001f);
      else k<1><<<148 * 4, 256>>>(out, iters, 0.999f, 0.001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = 148.0 * 4 * 256 * (double)iters * 32;
      printf("mode %d: %.3f ms  %.1f TFLOP/s\n", mode, ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
