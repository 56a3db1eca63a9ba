// fmlp.cu -- f-MLP head: fc1 -> ReLU -> fc2 -> Dropout -> ReLU -> fc3 -> log_softmax
// (reference model.py:155-162; the dropout sits before the ReLU).  0.014 % of the model's FLOPs:
// fp32 SIMT GEMMs plus two small row kernels.
#include "common.cuh"
#include "sgemm.cuh"

namespace rn {

// h2 = relu(z2 * mask * keep_scale), in place
__global__ void dropout_relu_kernel(float* __restrict__ z, const uint8_t* __restrict__ mask, float keep_scale,
                                    long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = z[i];
  if (mask) v = mask[i] ? v * keep_scale : 0.f;
  z[i] = fmaxf(v, 0.f);
}

// one warp per row: logp = z - max - log(sum(exp(z - max)))
__global__ void log_softmax_rows_kernel(float* __restrict__ z, int B, int A) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= B) return;
  float* r = z + (long long)row * A;
  float m = -INFINITY;
  for (int j = lane; j < A; j += 32) m = fmaxf(m, r[j]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < A; j += 32) s += expf(r[j] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float lse = m + logf(s);
  for (int j = lane; j < A; j += 32) r[j] -= lse;
}

// dz = dlogp - exp(logp) * sum_j dlogp_j
__global__ void log_softmax_bwd_rows_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp,
                                            float* __restrict__ dz, int B, int A) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (row >= B) return;
  const float* d = dlogp + (long long)row * A;
  const float* lp = logp + (long long)row * A;
  float s = 0.f;
  for (int j = lane; j < A; j += 32) s += d[j];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int j = lane; j < A; j += 32) dz[(long long)row * A + j] = d[j] - expf(lp[j]) * s;
}


// ------------------------------------------------------------------------------------------
// Fused head for the from-pixels shapes (G = F1 = F2 = 256, A <= 32).  As five (forward) and four (backward data path)
// launches of 80 .. 640-row GEMMs the head cost 0.065 + 0.06 ms per step whatever the batch -- 8 % of the step at 80
// questions per GPU.  Here a block takes kFR samples through the whole chain: thread o owns output unit o, its weight row
// streams through registers as float4 (L1 keeps the 128-byte lines of the 256 rows), the kFR activation rows sit in
// shared memory and are read as broadcast float4.  Backward products use W as [in][out] slices (coalesced across o).
// ------------------------------------------------------------------------------------------
constexpr int kFD = 256;      // G = F1 = F2
constexpr int kFR = 8;        // samples per block

// acc[r] += sum_k act[r][k] * W[o][k] for the block's kFR rows and this thread's output unit o = tid.  W is [out][in]
// row-major: a thread walking its own row is 32 different cache lines per warp load (measured: the first version of this
// kernel was L1-wavefront bound at 50 us).  So W streams through shared memory in 32-column chunks, loaded coalesced (8
// threads per 128-byte row segment) and stored TRANSPOSED with a 257-float row stride (conflict-free both ways), the next
// chunk prefetched in registers under the FMAs.
constexpr int kFWs = 32 * 257 > kFD * 33 ? 32 * 257 : kFD * 33;      // weight stage (floats): [32][257] or fc3's [256][33]

__device__ __forceinline__ void f_layer(float (&acc)[kFR], const float* __restrict__ W, const float* __restrict__ act,
                                        float* __restrict__ Ws, int tid) {
  const int lrow = tid >> 3, lc4 = tid & 7;             // chunk load: rows lrow + 32 i, columns 4 lc4 .. 4 lc4 + 3
  float4 pre[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pre[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(lrow + 32 * i) * kFD + 4 * lc4));
#pragma unroll 1
  for (int c = 0; c < kFD / 32; ++c) {
    __syncthreads();                                    // the previous chunk has been consumed
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float* d = Ws + (4 * lc4) * 257 + lrow + 32 * i;
      d[0] = pre[i].x; d[257] = pre[i].y; d[2 * 257] = pre[i].z; d[3 * 257] = pre[i].w;
    }
    __syncthreads();
    if (c + 1 < kFD / 32) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        pre[i] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(lrow + 32 * i) * kFD + 32 * (c + 1) + 4 * lc4));
    }
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      const float w0 = Ws[k * 257 + tid], w1 = Ws[(k + 1) * 257 + tid], w2 = Ws[(k + 2) * 257 + tid], w3 = Ws[(k + 3) * 257 + tid];
#pragma unroll
      for (int r = 0; r < kFR; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(act + r * kFD + 32 * c + k);
        acc[r] = fmaf(a.x, w0, fmaf(a.y, w1, fmaf(a.z, w2, fmaf(a.w, w3, acc[r]))));
      }
    }
  }
}

__global__ void __launch_bounds__(kFD)
f_fused_fwd_kernel(const float* __restrict__ xg, const float* __restrict__ w1, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                   const float* __restrict__ b3, const uint8_t* __restrict__ drop_mask, float keep_scale, int B, int A,
                   float* __restrict__ logp, float* __restrict__ h1_out, float* __restrict__ h2_out) {
  extern __shared__ __align__(16) float f_smem[];
  float* act0 = f_smem;                    // [kFR][kFD]
  float* act1 = act0 + kFR * kFD;
  float* Ws = act1 + kFR * kFD;            // [kFWs]
  const int o = threadIdx.x, row0 = blockIdx.x * kFR;
  const int nr = min(kFR, B - row0);
  for (int i = threadIdx.x; i < kFR * kFD; i += kFD) {
    const int r = i / kFD;
    act0[i] = r < nr ? xg[(size_t)(row0 + r) * kFD + (i - r * kFD)] : 0.f;
  }
  float acc[kFR];
  // fc1 + ReLU  (f_layer starts with a barrier: act0 is complete before it is read)
#pragma unroll
  for (int r = 0; r < kFR; ++r) acc[r] = b1[o];
  f_layer(acc, w1, act0, Ws, o);
#pragma unroll
  for (int r = 0; r < kFR; ++r) {
    const float v = fmaxf(acc[r], 0.f);
    act1[r * kFD + o] = v;
    if (r < nr) h1_out[(size_t)(row0 + r) * kFD + o] = v;
  }
  // fc2 + dropout + ReLU
#pragma unroll
  for (int r = 0; r < kFR; ++r) acc[r] = b2[o];
  f_layer(acc, w2, act1, Ws, o);
#pragma unroll
  for (int r = 0; r < kFR; ++r) {
    float v = acc[r];
    if (drop_mask != nullptr && r < nr) v = drop_mask[(size_t)(row0 + r) * kFD + o] ? v * keep_scale : 0.f;
    v = fmaxf(v, 0.f);
    act0[r * kFD + o] = v;
    if (r < nr) h2_out[(size_t)(row0 + r) * kFD + o] = v;
  }
  __syncthreads();                         // h2 complete, the weight stage is free
  // fc3 + log_softmax: W3 transposed into the stage as [k][33], then warp r = sample, lane a = answer
  for (int i = threadIdx.x; i < A * (kFD / 4); i += kFD) {
    const int a = i / (kFD / 4), k4 = i - a * (kFD / 4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(w3 + (size_t)a * kFD + 4 * k4));
    float* d = Ws + (4 * k4) * 33 + a;
    d[0] = v.x; d[33] = v.y; d[66] = v.z; d[99] = v.w;
  }
  __syncthreads();
  const int r = threadIdx.x >> 5, a = threadIdx.x & 31;
  float z = -INFINITY;
  if (a < A) {
    float s0 = b3[a], s1 = 0.f;
#pragma unroll 4
    for (int k = 0; k < kFD; k += 4) {
      const float4 h = *reinterpret_cast<const float4*>(act0 + r * kFD + k);
      s0 = fmaf(h.x, Ws[k * 33 + a], fmaf(h.y, Ws[(k + 1) * 33 + a], s0));
      s1 = fmaf(h.z, Ws[(k + 2) * 33 + a], fmaf(h.w, Ws[(k + 3) * 33 + a], s1));
    }
    z = s0 + s1;
  }
  float m = z;
  for (int sft = 16; sft; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
  float e = a < A ? expf(z - m) : 0.f;
  for (int sft = 16; sft; sft >>= 1) e += __shfl_xor_sync(0xffffffffu, e, sft);
  if (a < A && r < nr) logp[(size_t)(row0 + r) * A + a] = z - (m + logf(e));
}

// data path of the backward: dz3 = log_softmax backward, dz2 = (dz3 W3) . mask2 . alpha, dz1 = (dz2 W2) . mask1,
// dxg = dz1 W1.  dz3 / dz2 / dz1 also go to global memory for the grouped weight-gradient launch.
__global__ void __launch_bounds__(kFD)
f_fused_bwd_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp, const float* __restrict__ w1,
                   const float* __restrict__ w2, const float* __restrict__ w3, const float* __restrict__ h1,
                   const float* __restrict__ h2, float alpha, int B, int A, float* __restrict__ dz3_out,
                   float* __restrict__ dz2_out, float* __restrict__ dz1_out, float* __restrict__ dxg) {
  __shared__ __align__(16) float act[2][kFR][kFD];
  __shared__ float dz3s[kFR][32];
  const int o = threadIdx.x, row0 = blockIdx.x * kFR;
  const int nr = min(kFR, B - row0);
  {
    const int r = threadIdx.x >> 5, a = threadIdx.x & 31;
    const bool ok = a < A && r < nr;
    const float d = ok ? dlogp[(size_t)(row0 + r) * A + a] : 0.f;
    float ssum = d;
    for (int sft = 16; sft; sft >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, sft);
    const float v = ok ? d - expf(logp[(size_t)(row0 + r) * A + a]) * ssum : 0.f;
    dz3s[r][a] = v;
    if (ok) dz3_out[(size_t)(row0 + r) * A + a] = v;
  }
  __syncthreads();
  float acc[kFR];
  // dz2[r][o] = alpha * [h2 > 0] * sum_a dz3[r][a] W3[a][o]
#pragma unroll
  for (int r = 0; r < kFR; ++r) acc[r] = 0.f;
  {
    float w[32];                     // all of this unit's W3 column in flight at once (A <= 32)
#pragma unroll
    for (int a = 0; a < 32; ++a) w[a] = a < A ? __ldg(w3 + (size_t)a * kFD + o) : 0.f;
#pragma unroll
    for (int a = 0; a < 32; ++a)
#pragma unroll
      for (int r = 0; r < kFR; ++r) acc[r] = fmaf(dz3s[r][a], w[a], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kFR; ++r) {
    const float v = r < nr && h2[(size_t)(row0 + r) * kFD + o] > 0.f ? alpha * acc[r] : 0.f;
    act[0][r][o] = v;
    if (r < nr) dz2_out[(size_t)(row0 + r) * kFD + o] = v;
  }
  __syncthreads();
  // dz1[r][j] = [h1 > 0] * sum_o dz2[r][o] W2[o][j];  dxg[r][g] = sum_j dz1[r][j] W1[j][g].  W rows are read [k][o]
  // (coalesced across o) 32 at a time, the next 32 prefetched under the FMAs: with 10 .. 80 blocks on the machine these
  // loops are pure L2 latency otherwise (measured 75 us with 8 loads in flight)
#pragma unroll 1
  for (int layer = 0; layer < 2; ++layer) {
    const float* w = layer == 0 ? w2 : w1;
    const float* src = &act[layer][0][0];
#pragma unroll
    for (int r = 0; r < kFR; ++r) acc[r] = 0.f;
    float wn[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) wn[k] = __ldg(w + (size_t)k * kFD + o);
#pragma unroll 1
    for (int k0 = 0; k0 < kFD; k0 += 32) {
      float wc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) wc[k] = wn[k];
      if (k0 + 32 < kFD) {
#pragma unroll
        for (int k = 0; k < 32; ++k) wn[k] = __ldg(w + (size_t)(k0 + 32 + k) * kFD + o);
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4)
#pragma unroll
        for (int r = 0; r < kFR; ++r) {
          const float4 d = *reinterpret_cast<const float4*>(src + r * kFD + k0 + k);
          acc[r] = fmaf(d.x, wc[k], fmaf(d.y, wc[k + 1], fmaf(d.z, wc[k + 2], fmaf(d.w, wc[k + 3], acc[r]))));
        }
    }
    if (layer == 0) {
#pragma unroll
      for (int r = 0; r < kFR; ++r) {
        const float v = r < nr && h1[(size_t)(row0 + r) * kFD + o] > 0.f ? acc[r] : 0.f;
        act[1][r][o] = v;
        if (r < nr) dz1_out[(size_t)(row0 + r) * kFD + o] = v;
      }
      __syncthreads();
    } else {
#pragma unroll
      for (int r = 0; r < kFR; ++r)
        if (r < nr) dxg[(size_t)(row0 + r) * kFD + o] = acc[r];
    }
  }
}

static bool f_fused_ok(const rn_f_cfg* c) { return c->G == kFD && c->F1 == kFD && c->F2 == kFD && c->A <= 32; }

}  // namespace rn

using namespace rn;

static int validate_f(const rn_f_cfg* c) {
  RN_CHECK_ARG(c != nullptr, "cfg is NULL");
  RN_CHECK_ARG(c->B > 0 && c->G > 0 && c->F1 > 0 && c->F2 > 0 && c->A > 0, "f cfg sizes must be positive");
  return RN_OK;
}

extern "C" int rn_f_fwd(const rn_f_cfg* cfg, const float* xg, const float* w1, const float* b1, const float* w2,
                        const float* b2, const float* w3, const float* b3, const uint8_t* drop_mask, float* logp,
                        float* saved, void* stream) {
  RN_TRY(validate_f(cfg));
  RN_CHECK_ARG(xg && w1 && b1 && w2 && b2 && w3 && b3 && logp && saved, "NULL pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = cfg->B;
  float* h1 = saved;
  float* h2 = saved + (size_t)B * cfg->F1;
  if (f_fused_ok(cfg) && aligned16(xg) && aligned16(w1) && aligned16(w2) && aligned16(w3)) {
    const size_t smem = (size_t)(2 * kFR * kFD + kFWs) * sizeof(float);
    RN_CUDA(cudaFuncSetAttribute(f_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    f_fused_fwd_kernel<<<cdiv(B, kFR), kFD, smem, st>>>(xg, w1, b1, w2, b2, w3, b3, drop_mask, cfg->keep_scale, B, cfg->A, logp, h1, h2);
    RN_LAUNCH_CHECK("f_fused_fwd_kernel");
    return RN_OK;
  }
  GemmEpilogue e1;
  e1.bias = b1;
  e1.relu = 1;
  RN_TRY(sgemm(false, true, B, cfg->F1, cfg->G, xg, cfg->G, w1, cfg->G, h1, cfg->F1, e1, st));
  GemmEpilogue e2;
  e2.bias = b2;
  RN_TRY(sgemm(false, true, B, cfg->F2, cfg->F1, h1, cfg->F1, w2, cfg->F1, h2, cfg->F2, e2, st));
  const long long total = (long long)B * cfg->F2;
  dropout_relu_kernel<<<cdiv(total, 256), 256, 0, st>>>(h2, drop_mask, cfg->keep_scale, total);
  RN_LAUNCH_CHECK("dropout_relu_kernel");
  GemmEpilogue e3;
  e3.bias = b3;
  RN_TRY(sgemm(false, true, B, cfg->A, cfg->F2, h2, cfg->F2, w3, cfg->F2, logp, cfg->A, e3, st));
  log_softmax_rows_kernel<<<cdiv(B, 4), 128, 0, st>>>(logp, B, cfg->A);
  RN_LAUNCH_CHECK("log_softmax_rows_kernel");
  return RN_OK;
}

extern "C" int rn_f_bwd(const rn_f_cfg* cfg, const float* dlogp, const float* logp, const float* xg, const float* w1,
                        const float* w2, const float* w3, const uint8_t* drop_mask, const float* saved, float* dxg,
                        float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3, float* scratch,
                        void* stream) {
  RN_TRY(validate_f(cfg));
  RN_CHECK_ARG(dlogp && logp && xg && w1 && w2 && w3 && saved && dxg && dw1 && db1 && dw2 && db2 && dw3 && db3 && scratch,
               "NULL pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = cfg->B;
  const float* h1 = saved;
  const float* h2 = saved + (size_t)B * cfg->F1;
  float* dz3 = scratch;
  float* dz2 = dz3 + (size_t)B * cfg->A;
  float* dz1 = dz2 + (size_t)B * cfg->F2;
  if (f_fused_ok(cfg)) {
    f_fused_bwd_kernel<<<cdiv(B, kFR), kFD, 0, st>>>(dlogp, logp, w1, w2, w3, h1, h2, drop_mask ? cfg->keep_scale : 1.f, B, cfg->A, dz3,
                                                    dz2, dz1, dxg);
    RN_LAUNCH_CHECK("f_fused_bwd_kernel");
    AtbBuilder wgf;
    wgf.add(dz3, cfg->A, h2, cfg->F2, dw3, cfg->F2, cfg->A, cfg->F2, db3);
    wgf.add(dz2, cfg->F2, h1, cfg->F1, dw2, cfg->F1, cfg->F2, cfg->F1, db2);
    wgf.add(dz1, cfg->F1, xg, cfg->G, dw1, cfg->G, cfg->F1, cfg->G, db1);
    return wgf.launch(B, st);
  }
  log_softmax_bwd_rows_kernel<<<cdiv(B, 4), 128, 0, st>>>(dlogp, logp, dz3, B, cfg->A);
  RN_LAUNCH_CHECK("log_softmax_bwd_rows_kernel");
  GemmEpilogue none;
  // data path: dz3 -> dz2 (through ReLU and dropout: h2 > 0 implies the unit was kept) -> dz1 -> dxg
  GemmEpilogue m2;
  m2.mask = h2;
  m2.ldmask = cfg->F2;
  m2.alpha = drop_mask ? cfg->keep_scale : 1.f;
  RN_TRY(sgemm(false, false, B, cfg->F2, cfg->A, dz3, cfg->A, w3, cfg->F2, dz2, cfg->F2, m2, st));
  GemmEpilogue m1;
  m1.mask = h1;
  m1.ldmask = cfg->F1;
  RN_TRY(sgemm(false, false, B, cfg->F1, cfg->F2, dz2, cfg->F2, w2, cfg->F1, dz1, cfg->F1, m1, st));
  RN_TRY(sgemm(false, false, B, cfg->G, cfg->F1, dz1, cfg->F1, w1, cfg->G, dxg, cfg->G, none, st));
  // the three weight gradients dW = dz^T h and the three bias gradients (column sums of dz) in ONE launch
  AtbBuilder wg;
  wg.add(dz3, cfg->A, h2, cfg->F2, dw3, cfg->F2, cfg->A, cfg->F2, db3);
  wg.add(dz2, cfg->F2, h1, cfg->F1, dw2, cfg->F1, cfg->F2, cfg->F1, db2);
  wg.add(dz1, cfg->F1, xg, cfg->G, dw1, cfg->G, cfg->F1, cfg->G, db1);
  RN_TRY(wg.launch(B, st));
  return RN_OK;
}
