// Microbenchmark: issue rate of the warp-level mma.sync forms on sm_100a (the conv stack's tensor path).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_rate mma_sync_rate.cu && ./mma_sync_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int MODE, int NACC>
__global__ void __launch_bounds__(256) k(float* out, int iters, uint32_t seed) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else if (MODE == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC>
void run(const char* name, int blocks_per_sm, int sms, float* out) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NACC><<<sms * blocks_per_sm, 256>>>(out, 16, 1);
  cudaEventRecord(e0);
  k<MODE, NACC><<<sms * blocks_per_sm, 256>>>(out, iters, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = (double)sms * blocks_per_sm * 8 * iters * NACC;
  const double flop = mmas * 2.0 * 16 * 8 * (MODE == 0 ? 8 : 16);
  printf("%-28s acc=%2d blocks/SM=%d: %.3f ms, %.1f TFLOP/s, %.2f cycles/MMA/subcore @1.965GHz\n", name, NACC, blocks_per_sm, ms,
         flop / ms * 1e-9, ms * 1e-3 * 1.965e9 / (mmas / (sms * 4.0)));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 2);
  const int sms = p.multiProcessorCount;
  run<0, 4>("m16n8k8 tf32", 1, sms, out);
  run<0, 8>("m16n8k8 tf32", 1, sms, out);
  run<0, 8>("m16n8k8 tf32", 2, sms, out);
  run<0, 16>("m16n8k8 tf32", 2, sms, out);
  run<1, 8>("m16n8k16 f16", 1, sms, out);
  run<1, 8>("m16n8k16 f16", 2, sms, out);
  run<1, 16>("m16n8k16 f16", 2, sms, out);
  run<2, 16>("m16n8k16 bf16", 2, sms, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
