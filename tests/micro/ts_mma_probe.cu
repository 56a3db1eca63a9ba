// Micro-probe (diagnostic, not a test): facts the g-MLP chain kernel design depends on, measured on one B200.
//   1. numerics of tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form, A written with tcgen05.st as packed
//      fp16 pairs, lane = row, column = K/2) against the shared-memory form and a host reference;
//   2. issue-to-completion cycles per MMA for M128 x N{256,128} x K16, A from shared memory vs tensor memory;
//   3. per-SM ingest rate of cp.async.bulk weight-chunk streaming (32 KB chunks out of a 768 KB L2-resident image,
//      3-stage ring, all 148 SMs at once), alone -- the chain kernels re-stream 256 KB of weights per 128-row tile-layer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I relationnetworks_clevr_b200/csrc \
//        tests/micro/ts_mma_probe.cu -o tests/micro/ts_mma_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "tc_ptx.cuh"

using namespace rn::ptx;

__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kK = 64;       // one swizzle chunk
// smem: A chunk 128x64 fp16 (16 KB) | B chunk 256x64 fp16 (32 KB) | barrier
struct Out {
  float d_ss[128 * 256];
  float d_ts[128 * 256];
  long long cyc[16];
};

// mode bits: N in {128, 256}
__global__ void __launch_bounds__(128, 1) probe_kernel(const __half* A, const __half* B, Out* out, int reps) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  char* sA = smem;
  char* sB = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * kK; i += 128) {
    const int r = i / kK, c = i % kK;
    *reinterpret_cast<__half*>(sA + sw128_offset(r, c)) = A[r * kK + c];
  }
  for (int i = tid; i < 256 * kK; i += 128) {
    const int r = i / kK, c = i % kK;
    *reinterpret_cast<__half*>(sB + sw128_offset(r, c)) = B[r * kK + c];
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tbase), 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tbase;
  // A into tensor memory columns [384, 384 + 32): thread = row, column j holds K elements (2j, 2j+1)
  {
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) {
      __half2 h = __halves2half2(A[tid * kK + 2 * j], A[tid * kK + 2 * j + 1]);
      r[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 384, r);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t phase = 0;
  const uint32_t idesc256 = idesc_f16(128, 256, 0, 0), idesc128 = idesc_f16(128, 128, 0, 0);
  // ---- numerics: SS into columns [0,256), TS into the same columns afterwards ----
  for (int form = 0; form < 2; ++form) {
    if (tid == 0) {
      for (int k = 0; k < kK / 16; ++k) {
        const uint64_t bd = smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
        if (form == 0) mma_f16_ss(tmem, smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024), bd, idesc256, k > 0);
        else mma_f16_ts(tmem, tmem + 384 + k * 8, bd, idesc256, k > 0);
      }
      mma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), phase);
    phase ^= 1;
    tc_fence_after_sync();
    float* dst = form == 0 ? out->d_ss : out->d_ts;
    for (int cc = 0; cc < 8; ++cc) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, r);
      tmem_ld_wait();
      for (int e = 0; e < 32; ++e) dst[tid * 256 + cc * 32 + e] = __uint_as_float(r[e]);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }
  // ---- rates: `reps` back-to-back MMAs per variant, descriptors hoisted, unrolled by 4 ----
  uint64_t adv[4], bdv[4];
  for (int k = 0; k < 4; ++k) {
    adv[k] = smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024);
    bdv[k] = smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
  }
  const uint32_t idesc64 = idesc_f16(128, 64, 0, 0);
#define RATE(V, BODY)                                            \
  {                                                              \
    if (tid == 0) {                                              \
      const long long t0 = clock64();                            \
      for (int i = 0; i < reps; i += 4) {                        \
        _Pragma("unroll") for (int k = 0; k < 4; ++k) { BODY; }  \
      }                                                          \
      mma_commit(smem_u32(bar));                                 \
      mbar_wait(smem_u32(bar), phase);                           \
      out->cyc[V] = clock64() - t0;                              \
    } else {                                                     \
      mbar_wait(smem_u32(bar), phase);                           \
    }                                                            \
    phase ^= 1;                                                  \
    __syncthreads();                                             \
  }
  RATE(0, mma_f16_ss(tmem, adv[k], bdv[k], idesc256, 1))
  RATE(1, mma_f16_ss(tmem, adv[k], bdv[k], idesc128, 1))
  RATE(2, mma_f16_ts(tmem, tmem + 384 + k * 8, bdv[k], idesc256, 1))
  RATE(3, mma_f16_ts(tmem, tmem + 384 + k * 8, bdv[k], idesc128, 1))
  RATE(4, if (k & 1) mma_f16_ts(tmem, tmem + 384 + k * 8, bdv[k], idesc128, 1); else mma_f16_ss(tmem, adv[k], bdv[k], idesc128, 1))
  RATE(5, mma_f16_ss(tmem + (k & 1) * 128, adv[k], bdv[k], idesc128, 1))
  RATE(6, mma_f16_ss(tmem + (k & 1) * 256, adv[k], bdv[k], idesc256, 1))
  RATE(7, mma_f16_ss(tmem, adv[k], bdv[k], idesc64, 1))
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- weight-chunk streaming: every SM pulls `chunks` 32 KB chunks (cycling over a 768 KB image) through a 3-stage ring ----
__global__ void __launch_bounds__(64, 1) ingest_kernel(const char* img, int chunks, int chunk_bytes, int stages, long long* cyc) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 6 * 32768);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&full[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int per_img = 768 * 1024 / chunk_bytes;
    const long long t0 = clock64();
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < chunks + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) {                       // consume (wait) the chunk issued `stages` iterations ago
        mbar_wait(smem_u32(&full[s]), ph[s]);
        ph[s] ^= 1;
      }
      if (i < chunks) {
        mbar_expect_tx(smem_u32(&full[s]), chunk_bytes);
        bulk_g2s(smem_u32(smem + s * chunk_bytes), img + (size_t)(i % per_img) * chunk_bytes, chunk_bytes, smem_u32(&full[s]));
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

__global__ void __launch_bounds__(64, 1) ingest_split_kernel(const char* img, int chunks, int parts, long long* cyc) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 6 * 32768);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&full[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int pb = 32768 / parts;
    for (int i = 0; i < chunks + 3; ++i) {
      const int s = i % 3;
      if (i >= 3) {
        mbar_wait(smem_u32(&full[s]), ph[s]);
        ph[s] ^= 1;
      }
      if (i < chunks) {
        mbar_expect_tx(smem_u32(&full[s]), 32768);
        for (int p = 0; p < parts; ++p)
          bulk_g2s(smem_u32(smem + s * 32768 + p * pb), img + (size_t)(i % 24) * 32768 + p * pb, pb, smem_u32(&full[s]));
      }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  std::vector<__half> hA(128 * kK), hB(256 * kK);
  srand(1);
  for (auto& v : hA) v = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : hB) v = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dA, *dB;
  Out* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, sizeof(Out));
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 4096;
  probe_kernel<<<1, 128, smem>>>(dA, dB, dOut, reps);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe_kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<Out> ho(1);
  cudaMemcpy(ho.data(), dOut, sizeof(Out), cudaMemcpyDeviceToHost);
  double max_ss = 0, max_ts = 0, max_ref = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 256; ++n) {
      double ref = 0;
      for (int k = 0; k < kK; ++k) ref += (double)__half2float(hA[r * kK + k]) * (double)__half2float(hB[n * kK + k]);
      max_ref = fmax(max_ref, fabs(ref));
      max_ss = fmax(max_ss, fabs(ho[0].d_ss[r * 256 + n] - ref));
      max_ts = fmax(max_ts, fabs(ho[0].d_ts[r * 256 + n] - ref));
    }
  printf("numerics: max|ref| %.3f  max err SS %.3e  TS %.3e  (TS layout %s)\n", max_ref, max_ss, max_ts,
         max_ts < 1e-3 ? "CONFIRMED: lane=row, column=K/2 packed fp16 pairs" : "WRONG");
  const char* names[8] = {"SS N256", "SS N128", "TS N256", "TS N128", "SS/TS alternating N128", "SS N128 two accumulators", "SS N256 two accumulators", "SS N64"};
  for (int v = 0; v < 8; ++v) printf("rate %-28s %.1f cycles per MMA\n", names[v], (double)ho[0].cyc[v] / reps);

  // ingest
  char* img;
  long long* dcyc;
  cudaMalloc(&img, 768 * 1024);
  cudaMemset(img, 0, 768 * 1024);
  cudaMalloc(&dcyc, 160 * 8);
  const int ismem = 6 * 32768 + 64 + 1024;
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ismem);
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  for (int cb : {8192, 16384, 32768, 65536, 98304}) {
    for (int stages : {1, 2, 3, 6}) {
      if ((long long)cb * stages > 6 * 32768) continue;
      for (int grid : {1, nsm}) {
        const int chunks = 2048;
        ingest_kernel<<<grid, 64, ismem>>>(img, chunks, cb, stages, dcyc);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ingest: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> hc(grid);
        cudaMemcpy(hc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (auto c : hc) mx = c > mx ? c : mx;
        printf("ingest chunk %6d B  stages %d  grid %3d: %.1f B/cycle/SM (slowest SM), %.0f cycles per chunk\n", cb, stages, grid,
               (double)chunks * cb / mx, (double)mx / chunks);
      }
    }
  }
  // split issue: each 32 KB chunk as `parts` separate copies on one barrier (same thread)
  for (int parts : {2, 4, 8}) {
    ingest_split_kernel<<<nsm, 64, ismem>>>(img, 2048, parts, dcyc);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ingest_split: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> hc(nsm);
    cudaMemcpy(hc.data(), dcyc, nsm * 8, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : hc) mx = c > mx ? c : mx;
    printf("ingest 32 KB chunk as %d copies, 3 stages, grid %d: %.1f B/cycle/SM\n", parts, nsm, 2048.0 * 32768 / mx);
  }
  return 0;
}
