// relation.cu -- relation layer: shared layer-0 pieces, the fp32 SIMT path and the C-ABI dispatch.
// Math: SURVEY.md 7.4 (restating reference model.py:104-152 and its autograd backward).
#include "relation.cuh"

#include <algorithm>

namespace rn {

// ------------------------------------------------------------------------------------------
// layer 0 ("pre"): U = X W0c^T, Vb = X W0a^T + beta0, Qb = q Wq^T + b_qinj
// ------------------------------------------------------------------------------------------
__global__ void add_rowgroup_bias_kernel(float* __restrict__ V, const float* __restrict__ bias, long long total, int G,
                                         int rows_per_group, long long group_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long r = i / G;
  const int c = i % G;
  V[i] += bias[(rows_per_group ? (r / rows_per_group) * group_stride : 0) + c];
}

// Fused layer-0 "pre" for the from-pixels shapes (k == 26, G <= 256): persistent blocks of G threads (thread = output
// feature g) keep W0c / W0a rows in registers and the question block transposed in shared memory, and walk the samples:
//   Qb[b, g] = b_qinj[g] + sum_j q[b, j] Wq[g, j];  U[b, o, g] = x[b, o, :] . W0c[g, :];  Vb = x . W0a[g, :] + beta0
// One launch instead of three GEMMs and a bias pass; U / Vb rows are written once, fully coalesced.
template <int K>
__global__ void __launch_bounds__(512)
rel_pre_fused_kernel(const float* __restrict__ x, const float* __restrict__ q, const float* __restrict__ w0, int fan0,
                     const float* __restrict__ wq, int ldq, const float* __restrict__ bq, const float* __restrict__ b0,
                     int q_at_layer0, int B, int n, int Q, int G, float* __restrict__ U, float* __restrict__ Vb,
                     float* __restrict__ Qb, float* __restrict__ U4) {
  extern __shared__ __align__(16) float pre_smem[];
  float* wqT = pre_smem;                          // [Q][G + 1]
  float* xs = wqT + (size_t)Q * (G + 1);          // [n][K]
  float* qs = xs + (((size_t)n * K + 3) & ~(size_t)3);
  const int g = threadIdx.x % G, ohalf = threadIdx.x / G;      // two threads per feature: each takes half of the objects
  for (int base = 0; base < G * Q; base += 8 * blockDim.x) {      // 8 loads in flight per thread
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * blockDim.x + threadIdx.x;
      t[u] = idx < G * Q ? wq[(size_t)(idx / Q) * ldq + idx % Q] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * blockDim.x + threadIdx.x;
      if (idx < G * Q) wqT[(size_t)(idx % Q) * (G + 1) + idx / Q] = t[u];
    }
  }
  float wc[K], wa[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    wc[j] = w0[(size_t)g * fan0 + j];
    wa[j] = w0[(size_t)g * fan0 + K + j];
  }
  const float bqv = bq[g], b0v = b0[g];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();            // previous sample's xs / qs are no longer read (also orders the wqT fill)
    for (int base = 0; base < n * K; base += 8 * blockDim.x) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * blockDim.x + threadIdx.x;
        t[u] = i < n * K ? x[(size_t)b * n * K + i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = base + u * blockDim.x + threadIdx.x;
        if (i < n * K) xs[i] = t[u];
      }
    }
    for (int i = threadIdx.x; i < Q; i += blockDim.x) qs[i] = q[(size_t)b * Q + i];
    __syncthreads();
    float acc = bqv;
#pragma unroll 8
    for (int j = 0; j < Q; ++j) acc = fmaf(qs[j], wqT[(size_t)j * (G + 1) + g], acc);
    if (ohalf == 0) Qb[(size_t)b * G + g] = acc;
    const float vb = q_at_layer0 ? acc : b0v;
    const int o_begin = ohalf * ((n + 1) / 2), o_end = ohalf == 0 ? (n + 1) / 2 : n;
#pragma unroll 2
    for (int o = o_begin; o < o_end; ++o) {
      float u = 0.f, v = vb;
#pragma unroll
      for (int j = 0; j < K; j += 2) {
        const float2 xv = *reinterpret_cast<const float2*>(&xs[o * K + j]);
        u = fmaf(xv.x, wc[j], u);
        u = fmaf(xv.y, wc[j + 1], u);
        v = fmaf(xv.x, wa[j], v);
        v = fmaf(xv.y, wa[j + 1], v);
      }
      U[((size_t)b * n + o) * G + g] = u;
      if (U4) U4[(((size_t)b * (G / 4) + (g >> 2)) * n + o) * 4 + (g & 3)] = u;
      Vb[((size_t)b * n + o) * G + g] = v;
    }
  }
}

// U [B, n, G] -> U4 [B][G/4][n][4]
__global__ void u4_transpose_kernel(const float* __restrict__ U, float* __restrict__ U4, int n, int G4, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // over B * n * G4 float4 groups of U
  if (i >= total4) return;
  const int g4 = i % G4;
  const long long r = i / G4;
  const int o = r % n;
  const long long b = r / n;
  reinterpret_cast<float4*>(U4)[(b * G4 + g4) * n + o] = reinterpret_cast<const float4*>(U)[i];
}

static size_t rel_pre_fused_smem(const RelShape& s) {
  return ((size_t)s.Q * (s.G + 1) + (((size_t)s.n * s.k + 3) & ~(size_t)3) + s.Q) * sizeof(float);
}

int relation_pre(const RelShape& s, const float* x, const float* q, const float* const* g_w, const float* const* g_b,
                 const RelPre& pre, cudaStream_t st) {
  const int fan0 = s.fan_in(0);
  if (s.k == 26 && s.G == 256 && rel_pre_fused_smem(s) <= 200 * 1024) {
    const float* wq = g_w[s.qinj] + (s.qinj == 0 ? 2 * s.k : s.G);
    const size_t smem = rel_pre_fused_smem(s);
    RN_CUDA(cudaFuncSetAttribute(rel_pre_fused_kernel<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rel_pre_fused_kernel<26><<<std::min(s.B, sm_count()), 2 * s.G, smem, st>>>(x, q, g_w[0], fan0, wq, s.fan_in(s.qinj),
                                                                          g_b[s.qinj], g_b[0], s.qinj == 0 ? 1 : 0, s.B,
                                                                          s.n, s.Q, s.G, pre.U, pre.Vb, pre.Qb, pre.U4);
    RN_LAUNCH_CHECK("rel_pre_fused_kernel");
    return RN_OK;
  }
  GemmEpilogue none;
  // U[b,c,:] = x[b,c,:] . W0[:, 0:k]^T ; V[b,a,:] = x[b,a,:] . W0[:, k:2k]^T
  RN_TRY(sgemm(false, true, s.B * s.n, s.G, s.k, x, s.k, g_w[0], fan0, pre.U, s.G, none, st));
  RN_TRY(sgemm(false, true, s.B * s.n, s.G, s.k, x, s.k, g_w[0] + s.k, fan0, pre.Vb, s.G, none, st));
  // Qb[b,:] = q[b] . Wq^T + b_qinj  (the question columns of layer qinj)
  const float* wq = g_w[s.qinj] + (s.qinj == 0 ? 2 * s.k : s.G);
  GemmEpilogue eb;
  eb.bias = g_b[s.qinj];
  RN_TRY(sgemm(false, true, s.B, s.G, s.Q, q, s.Q, wq, s.fan_in(s.qinj), pre.Qb, s.G, eb, st));
  // fold beta0 into V: per-sample Qb when the question enters at layer 0, else the plain bias b0
  const long long total = (long long)s.B * s.n * s.G;
  if (s.qinj == 0)
    add_rowgroup_bias_kernel<<<cdiv(total, 256), 256, 0, st>>>(pre.Vb, pre.Qb, total, s.G, s.n, s.G);
  else
    add_rowgroup_bias_kernel<<<cdiv(total, 256), 256, 0, st>>>(pre.Vb, g_b[0], total, s.G, 0, 0);
  RN_LAUNCH_CHECK("add_rowgroup_bias_kernel");
  if (pre.U4) {
    const long long total4 = (long long)s.B * s.n * (s.G / 4);
    u4_transpose_kernel<<<cdiv(total4, 256), 256, 0, st>>>(pre.U, pre.U4, s.n, s.G / 4, total4);
    RN_LAUNCH_CHECK("u4_transpose_kernel");
  }
  return RN_OK;
}

int relation_qinj_bwd(const RelShape& s, int l, const float* q, const float* const* g_w, const float* delta,
                      float* dq, float* const* dg_w, cudaStream_t st) {
  const int fan = s.fan_in(l);
  const int off = (l == 0 ? 2 * s.k : s.G);
  GemmEpilogue none;
  // dWq[o, j] = sum_b delta[b,o] q[b,j]
  AtbBuilder wg;
  wg.add(delta, s.G, q, s.Q, dg_w[l] + off, fan, s.G, s.Q, nullptr);
  RN_TRY(wg.launch(s.B, st));
  // dq[b, j] = sum_o delta[b,o] Wq[o,j]
  RN_TRY(sgemm(false, false, s.B, s.Q, s.G, delta, s.G, g_w[l] + off, fan, dq, s.Q, none, st));
  return RN_OK;
}

// Fused layer-0 backward for the from-pixels shapes (k == 26, n == 64, G == 256).  Persistent blocks of 256 threads walk
// the samples; dU[b] / dV[b] ([64, 256] each) are staged in shared memory with a 257-float row stride so that both
// access patterns below are conflict-free:
//   phase A (thread = feature g): delta0[b, g] = sum_a dV[a, g];  dW0c[g, :] += dU[c, g] x[c, :];  dW0a[g, :] += dV[a, g] x[a, :]
//                                 (52 register accumulators per thread, carried over all samples of the block)
//   phase B (thread = object o x K-quarter): dX[b, o, :] = dU[o, :] W0c + dV[o, :] W0a  (K = 256 features, 4-way split)
// Per-block partials of dW0 / db0 are summed in a fixed order by layer0_bwd_reduce_kernel.  The question-injection
// terms use delta0 (written to global) through the small GEMMs of relation_qinj_bwd.
constexpr int kL0N = 64, kL0K = 26, kL0G = 256, kL0Ld = kL0G + 1;
static size_t layer0_bwd_smem() {
  return ((size_t)2 * kL0N * kL0Ld + (size_t)kL0G * 2 * kL0K + (size_t)kL0N * kL0K + (size_t)4 * kL0N * 28) * sizeof(float);
}

__global__ void __launch_bounds__(512)
layer0_bwd_fused_kernel(const float* __restrict__ dU, const float* __restrict__ dV, const float* __restrict__ x,
                        const float* __restrict__ w0, int fan0, int B, float* __restrict__ delta, float* __restrict__ dx,
                        float* __restrict__ part) {
  extern __shared__ __align__(16) float l0_smem[];
  float* us = l0_smem;                              // [64][257]
  float* vs = us + kL0N * kL0Ld;                    // [64][257]
  float* ws = vs + kL0N * kL0Ld;                    // [256][52]: W0c | W0a rows
  float* xs = ws + kL0G * 2 * kL0K;                 // [64][26]
  float* red = xs + kL0N * kL0K;                    // [4][64][28]; its head doubles as the delta0 exchange [256]
  const int tid = threadIdx.x, g = tid & 255, ohalf = tid >> 8;    // phase A: two threads per feature (object halves)
  for (int base = 0; base < kL0G * 2 * kL0K; base += 8 * 512) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 512 + tid;
      t[u] = idx < kL0G * 2 * kL0K ? w0[(size_t)(idx / (2 * kL0K)) * fan0 + idx % (2 * kL0K)] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * 512 + tid;
      if (idx < kL0G * 2 * kL0K) ws[idx] = t[u];
    }
  }
  float wc[kL0K], wa[kL0K], db = 0.f;
#pragma unroll
  for (int j = 0; j < kL0K; ++j) wc[j] = wa[j] = 0.f;

  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();            // previous sample fully consumed (also orders the ws fill)
    const float* du = dU + (size_t)b * kL0N * kL0G;
    const float* dv = dV + (size_t)b * kL0N * kL0G;
#pragma unroll 1
    for (int o0 = ohalf * 32; o0 < ohalf * 32 + 32; o0 += 8) {          // 16 coalesced row loads in flight per thread
      float tu[8], tv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        tu[i] = du[(size_t)(o0 + i) * kL0G + g];
        tv[i] = dv[(size_t)(o0 + i) * kL0G + g];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        us[(o0 + i) * kL0Ld + g] = tu[i];
        vs[(o0 + i) * kL0Ld + g] = tv[i];
      }
    }
    for (int i = tid; i < kL0N * kL0K; i += 512) xs[i] = x[(size_t)b * kL0N * kL0K + i];
    __syncthreads();

    // ---- phase A ----
    float d0 = 0.f;
#pragma unroll 2
    for (int o = ohalf * 32; o < ohalf * 32 + 32; ++o) {
      const float u = us[o * kL0Ld + g], v = vs[o * kL0Ld + g];
      d0 += v;
#pragma unroll
      for (int j = 0; j < kL0K; j += 2) {
        const float2 xv = *reinterpret_cast<const float2*>(&xs[o * kL0K + j]);
        wc[j] = fmaf(u, xv.x, wc[j]);
        wc[j + 1] = fmaf(u, xv.y, wc[j + 1]);
        wa[j] = fmaf(v, xv.x, wa[j]);
        wa[j + 1] = fmaf(v, xv.y, wa[j + 1]);
      }
    }
    if (ohalf == 1) red[g] = d0;                    // delta0 = sum over both object halves
    __syncthreads();
    if (ohalf == 0) {
      d0 += red[g];
      delta[(size_t)b * kL0G + g] = d0;
      db += d0;
    }
    __syncthreads();                                // red is reused below

    // ---- phase B: thread = (object o, K-quarter kq, half of the 26 outputs) ----
    {
      constexpr int JH = kL0K / 2;                  // 13
      const int o = tid & 63, kq = (tid >> 6) & 3, j0 = (tid >> 8) * JH;
      float acc[JH];
#pragma unroll
      for (int j = 0; j < JH; ++j) acc[j] = 0.f;
#pragma unroll 2
      for (int gg = kq * 64; gg < kq * 64 + 64; ++gg) {
        const float u = us[o * kL0Ld + gg], v = vs[o * kL0Ld + gg];
        const float* wr = ws + gg * 2 * kL0K + j0;         // warp-uniform: broadcast reads
#pragma unroll
        for (int j = 0; j < JH; ++j) acc[j] = fmaf(u, wr[j], fmaf(v, wr[kL0K + j], acc[j]));
      }
#pragma unroll
      for (int j = 0; j < JH; ++j) red[(kq * kL0N + o) * 28 + j0 + j] = acc[j];
    }
    __syncthreads();
    for (int i = tid; i < kL0N * kL0K; i += 512) {
      const int o = i / kL0K, j = i % kL0K;
      dx[(size_t)b * kL0N * kL0K + i] = (red[(0 * kL0N + o) * 28 + j] + red[(1 * kL0N + o) * 28 + j]) +
                                        (red[(2 * kL0N + o) * 28 + j] + red[(3 * kL0N + o) * 28 + j]);
    }
  }
  float* pr = part + ((size_t)blockIdx.x * 2 + ohalf) * (kL0G * (2 * kL0K + 1));      // one partial per object half
#pragma unroll
  for (int j = 0; j < kL0K; ++j) {
    pr[g * (2 * kL0K + 1) + j] = wc[j];
    pr[g * (2 * kL0K + 1) + kL0K + j] = wa[j];
  }
  pr[g * (2 * kL0K + 1) + 2 * kL0K] = db;            // zero for the second half (delta0 is folded into the first)
}

// dW0[g, 0:52] and db0[g] = fixed-order sums of the per-block partials
__global__ void layer0_bwd_reduce_kernel(const float* __restrict__ part, int nblk, float* __restrict__ dw0, int fan0,
                                         float* __restrict__ db0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int W = 2 * kL0K + 1;
  if (idx >= kL0G * W) return;
  float v = 0.f;
  for (int i = 0; i < nblk; ++i) v += part[(size_t)i * kL0G * W + idx];
  const int g = idx / W, j = idx % W;
  if (j < 2 * kL0K) dw0[(size_t)g * fan0 + j] = v;
  else db0[g] = v;
}

int relation_layer0_bwd(const RelShape& s, const float* x, const float* q, const float* const* g_w, const float* dU,
                        const float* dV, float* delta, float* dx, float* dq, float* const* dg_w,
                        float* const* dg_b, float* ws, size_t ws_floats, cudaStream_t st) {
  const int fan0 = s.fan_in(0);
  const int nblk = std::min(s.B, sm_count());
  if (s.k == kL0K && s.n == kL0N && s.G == kL0G && ws && ws_floats >= (size_t)2 * nblk * kL0G * (2 * kL0K + 1)) {
    const size_t smem = layer0_bwd_smem();
    RN_CUDA(cudaFuncSetAttribute(layer0_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layer0_bwd_fused_kernel<<<nblk, 512, smem, st>>>(dU, dV, x, g_w[0], fan0, s.B, delta, dx, ws);
    RN_LAUNCH_CHECK("layer0_bwd_fused_kernel");
    layer0_bwd_reduce_kernel<<<cdiv(kL0G * (2 * kL0K + 1), 256), 256, 0, st>>>(ws, 2 * nblk, dg_w[0], fan0, dg_b[0]);
    RN_LAUNCH_CHECK("layer0_bwd_reduce_kernel");
    if (s.qinj == 0) RN_TRY(relation_qinj_bwd(s, 0, q, g_w, delta, dq, dg_w, st));
    return RN_OK;
  }
  GemmEpilogue none, acc;
  acc.beta = 1.f;
  // delta0[b] = sum_a dV[b,a]   (== sum over all pairs of dZ1) ; db0 = sum_b delta0
  RN_TRY(colsum(dV, delta, s.G, s.B, 1, s.n, 0, 1, s.n, st));
  RN_TRY(colsum(delta, dg_b[0], s.G, 1, 1, 0, 0, 1, s.B, st));
  // dW0c = dU^T X, dW0a = dV^T X  (sum over batch and objects)
  RN_TRY(sgemm(true, false, s.G, s.k, s.B * s.n, dU, s.G, x, s.k, dg_w[0], fan0, none, st, ws, ws_floats));
  RN_TRY(sgemm(true, false, s.G, s.k, s.B * s.n, dV, s.G, x, s.k, dg_w[0] + s.k, fan0, none, st, ws, ws_floats));
  if (s.qinj == 0) RN_TRY(relation_qinj_bwd(s, 0, q, g_w, delta, dq, dg_w, st));
  // dX = dU W0c + dV W0a
  RN_TRY(sgemm(false, false, s.B * s.n, s.k, s.G, dU, s.G, g_w[0], fan0, dx, s.k, none, st));
  RN_TRY(sgemm(false, false, s.B * s.n, s.k, s.G, dV, s.G, g_w[0] + s.k, fan0, dx, s.k, acc, st));
  return RN_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 SIMT path
// ------------------------------------------------------------------------------------------
// H1[b, a, c, :] = relu(U[b,c,:] + Vb[b,a,:])
__global__ void pairgen_kernel(const float* __restrict__ U, const float* __restrict__ Vb, float* __restrict__ H,
                               int n, int G4, long long total4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int g = i % G4;
  long long r = i / G4;
  const int c = r % n;
  r /= n;
  const int a = r % n;
  const long long b = r / n;
  const float4 u = reinterpret_cast<const float4*>(U)[(b * n + c) * G4 + g];
  const float4 v = reinterpret_cast<const float4*>(Vb)[(b * n + a) * G4 + g];
  float4 h;
  h.x = fmaxf(u.x + v.x, 0.f);
  h.y = fmaxf(u.y + v.y, 0.f);
  h.z = fmaxf(u.z + v.z, 0.f);
  h.w = fmaxf(u.w + v.w, 0.f);
  reinterpret_cast<float4*>(H)[i] = h;
}

// dZ[b, p, :] = dxg[b, :] * (H[b, p, :] > 0)
__global__ void relu_bwd_broadcast_kernel(const float* __restrict__ dxg, const float* __restrict__ H,
                                          float* __restrict__ dZ, long long pairs, int G, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = i % G;
  const long long b = i / G / pairs;
  dZ[i] = H[i] > 0.f ? dxg[b * G + g] : 0.f;
}

struct SimtSaved {
  RelPre pre;
  float* H[RN_MAX_G_LAYERS];
};

static SimtSaved simt_carve_saved(const RelShape& s, void* saved, int n_h) {
  Carver c(saved);
  SimtSaved out;
  out.pre.U = c.take<float>((size_t)s.B * s.n * s.G);
  out.pre.Vb = c.take<float>((size_t)s.B * s.n * s.G);
  out.pre.Qb = c.take<float>((size_t)s.B * s.G);
  for (int l = 0; l < n_h; ++l) out.H[l] = c.take<float>((size_t)s.rows * s.G);
  return out;
}

size_t simt_saved_bytes(const RelShape& s, bool training) {
  const size_t pre = 2 * round_up((size_t)s.B * s.n * s.G * 4, 256) + round_up((size_t)s.B * s.G * 4, 256);
  const size_t h = round_up((size_t)s.rows * s.G * 4, 256);
  return pre + (training ? s.L : 2) * h;     // eval ping-pongs two activation buffers
}

static size_t splitk_floats(const RelShape& s) { return (size_t)8 * 148 * 128 * 128; }

size_t simt_scratch_bytes(const RelShape& s, bool training) {
  size_t b = round_up(splitk_floats(s) * 4, 256);
  if (training) {
    b += 2 * round_up((size_t)s.rows * s.G * 4, 256);          // dZ ping-pong
    b += 2 * round_up((size_t)s.B * s.n * s.G * 4, 256);       // dU, dV
    b += round_up((size_t)s.B * s.G * 4, 256);                 // delta
  }
  return b;
}

int simt_relation_fwd(const RelShape& s, bool training, const float* x, const float* q, const float* const* g_w,
                      const float* const* g_b, float* xg, void* saved, void* scratch, cudaStream_t st) {
  const int n_h = training ? s.L : 2;
  SimtSaved sv = simt_carve_saved(s, saved, n_h);
  RN_TRY(relation_pre(s, x, q, g_w, g_b, sv.pre, st));
  const long long total4 = s.rows * (s.G / 4);
  pairgen_kernel<<<cdiv(total4, 256), 256, 0, st>>>(sv.pre.U, sv.pre.Vb, sv.H[0], s.n, s.G / 4, total4);
  RN_LAUNCH_CHECK("pairgen_kernel");
  const float* cur = sv.H[0];
  for (int l = 1; l < s.L; ++l) {
    float* nxt = sv.H[training ? l : (l & 1)];
    GemmEpilogue ep;
    ep.relu = 1;
    if (l == s.qinj) {
      ep.bias = sv.pre.Qb;
      ep.bias_group_rows = (int)s.pairs;
      ep.bias_group_stride = s.G;
    } else {
      ep.bias = g_b[l];
    }
    RN_TRY(sgemm(false, true, (int)s.rows, s.G, s.G, cur, s.G, g_w[l], s.fan_in(l), nxt, s.G, ep, st));
    cur = nxt;
  }
  RN_TRY(colsum(cur, xg, s.G, s.B, 1, s.pairs, 0, 1, (int)s.pairs, st));
  (void)scratch;
  return RN_OK;
}

int simt_relation_bwd(const RelShape& s, const float* dxg, const float* x, const float* q, const float* const* g_w,
                      const void* saved, float* dx, float* dq, float* const* dg_w, float* const* dg_b, void* scratch,
                      cudaStream_t st) {
  SimtSaved sv = simt_carve_saved(s, const_cast<void*>(saved), s.L);
  Carver c(scratch);
  float* ws = c.take<float>(splitk_floats(s));
  float* dZa = c.take<float>((size_t)s.rows * s.G);
  float* dZb = c.take<float>((size_t)s.rows * s.G);
  float* dU = c.take<float>((size_t)s.B * s.n * s.G);
  float* dV = c.take<float>((size_t)s.B * s.n * s.G);
  float* delta = c.take<float>((size_t)s.B * s.G);

  const long long total = s.rows * s.G;
  relu_bwd_broadcast_kernel<<<cdiv(total, 256), 256, 0, st>>>(dxg, sv.H[s.L - 1], dZa, s.pairs, s.G, total);
  RN_LAUNCH_CHECK("relu_bwd_broadcast_kernel");
  float* dZ = dZa;
  float* dZn = dZb;
  for (int l = s.L - 1; l >= 1; --l) {
    const int fan = s.fan_in(l);
    // delta[b] = sum_pairs dZ ; db_l = sum_b delta
    RN_TRY(colsum(dZ, delta, s.G, s.B, 1, s.pairs, 0, 1, (int)s.pairs, st));
    RN_TRY(colsum(delta, dg_b[l], s.G, 1, 1, 0, 0, 1, s.B, st));
    // dWh_l = dZ^T H_{l-1}
    GemmEpilogue none;
    RN_TRY(sgemm(true, false, s.G, s.G, (int)s.rows, dZ, s.G, sv.H[l - 1], s.G, dg_w[l], fan, none, st, ws,
                 splitk_floats(s)));
    if (l == s.qinj) RN_TRY(relation_qinj_bwd(s, l, q, g_w, delta, dq, dg_w, st));
    // dZ_{l} = (dZ Wh_l) .* (H_{l-1} > 0)
    GemmEpilogue ep;
    ep.mask = sv.H[l - 1];
    ep.ldmask = s.G;
    RN_TRY(sgemm(false, false, (int)s.rows, s.G, s.G, dZ, s.G, g_w[l], fan, dZn, s.G, ep, st));
    float* t = dZ; dZ = dZn; dZn = t;
  }
  // dZ is dZ1 [B, a, c, G]: dU[b,c] = sum_a, dV[b,a] = sum_c
  RN_TRY(colsum(dZ, dU, s.G, s.B, s.n, s.pairs, 1, s.n, s.n, st));
  RN_TRY(colsum(dZ, dV, s.G, s.B * s.n, 1, s.n, 0, 1, s.n, st));
  RN_TRY(relation_layer0_bwd(s, x, q, g_w, dU, dV, delta, dx, dq, dg_w, dg_b, ws, splitk_floats(s), st));
  return RN_OK;
}

}  // namespace rn

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace rn;

static int validate_cfg(const rn_relation_cfg* cfg) {
  RN_CHECK_ARG(cfg != nullptr, "cfg is NULL");
  RN_CHECK_ARG(cfg->B > 0 && cfg->n > 0 && cfg->k > 0 && cfg->Q > 0, "B, n, k, Q must be positive (B=%d n=%d k=%d Q=%d)",
               cfg->B, cfg->n, cfg->k, cfg->Q);
  RN_CHECK_ARG(cfg->G > 0 && cfg->G % 4 == 0, "G must be a positive multiple of 4 (G=%d)", cfg->G);
  RN_CHECK_ARG(cfg->L >= 2 && cfg->L <= RN_MAX_G_LAYERS, "L must be in [2,%d] (L=%d)", RN_MAX_G_LAYERS, cfg->L);
  RN_CHECK_ARG(cfg->qinj >= 0 && cfg->qinj < cfg->L, "qinj must be in [0,L) (qinj=%d)", cfg->qinj);
  RN_CHECK_ARG(cfg->precision >= RN_PRECISION_FP32 && cfg->precision <= RN_PRECISION_FAST, "unknown precision %d",
               cfg->precision);
  RN_CHECK_ARG((long long)cfg->B * cfg->n * cfg->n < (1LL << 31), "B*n*n must fit in int32");
  return RN_OK;
}

extern "C" int rn_relation_tc_supported(const rn_relation_cfg* cfg) {
  if (validate_cfg(cfg) != RN_OK) return 0;
  return tc_supported(RelShape(*cfg)) ? 1 : 0;
}

extern "C" int rn_relation_workspace(const rn_relation_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes) {
  RN_TRY(validate_cfg(cfg));
  RN_CHECK_ARG(saved_bytes && scratch_bytes, "output pointers are NULL");
  RelShape s(*cfg);
  if (cfg->precision == RN_PRECISION_FP32) {
    *saved_bytes = simt_saved_bytes(s, cfg->training != 0);
    *scratch_bytes = simt_scratch_bytes(s, cfg->training != 0);
  } else {
    if (!tc_supported(s))
      return fail(RN_ERR_UNSUPPORTED, "tcgen05 path needs G==256, L==4, n*n %% 128 == 0 (G=%d L=%d n=%d)", s.G, s.L, s.n);
    *saved_bytes = tc_saved_bytes(s, cfg->precision, cfg->training != 0);
    *scratch_bytes = tc_scratch_bytes(s, cfg->training != 0);
  }
  return RN_OK;
}

extern "C" int rn_relation_fwd(const rn_relation_cfg* cfg, const float* x, const float* q, const float* const* h_g_w,
                               const float* const* h_g_b, float* xg, void* saved, void* scratch, void* stream) {
  RN_TRY(validate_cfg(cfg));
  RN_CHECK_ARG(x && q && h_g_w && h_g_b && xg && saved && scratch, "NULL pointer argument");
  RN_CHECK_ARG(aligned16(x) && aligned16(q) && aligned16(xg) && aligned16(saved) && aligned16(scratch),
               "buffers must be 16-byte aligned");
  for (int l = 0; l < cfg->L; ++l) RN_CHECK_ARG(h_g_w[l] && h_g_b[l], "g layer %d pointer is NULL", l);
  RelShape s(*cfg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cfg->precision == RN_PRECISION_FP32)
    return simt_relation_fwd(s, cfg->training != 0, x, q, h_g_w, h_g_b, xg, saved, scratch, st);
  if (!tc_supported(s)) return fail(RN_ERR_UNSUPPORTED, "tcgen05 path does not support this shape");
  return tc_relation_fwd(s, cfg->precision, cfg->training != 0, x, q, h_g_w, h_g_b, xg, saved, scratch, st);
}

extern "C" int rn_relation_bwd(const rn_relation_cfg* cfg, const float* dxg, const float* x, const float* q,
                               const float* const* h_g_w, const void* saved, float* dx, float* dq,
                               float* const* h_dg_w, float* const* h_dg_b, void* scratch, void* stream) {
  RN_TRY(validate_cfg(cfg));
  RN_CHECK_ARG(cfg->training != 0, "rn_relation_bwd needs a cfg with training=1 (same as the forward call)");
  RN_CHECK_ARG(dxg && x && q && h_g_w && saved && dx && dq && h_dg_w && h_dg_b && scratch, "NULL pointer argument");
  for (int l = 0; l < cfg->L; ++l) RN_CHECK_ARG(h_g_w[l] && h_dg_w[l] && h_dg_b[l], "g layer %d pointer is NULL", l);
  RelShape s(*cfg);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cfg->precision == RN_PRECISION_FP32)
    return simt_relation_bwd(s, dxg, x, q, h_g_w, saved, dx, dq, h_dg_w, h_dg_b, scratch, st);
  if (!tc_supported(s)) return fail(RN_ERR_UNSUPPORTED, "tcgen05 path does not support this shape");
  return tc_relation_bwd(s, cfg->precision, dxg, x, q, h_g_w, saved, dx, dq, h_dg_w, h_dg_b, scratch, st);
}

// ------------------------------------------------------------------------------------------
// feature extraction support (reference extract.py:63-74): per sample, the max and the mean over all pair rows of the
// L2-normalised input of a g layer.  z [B * P, ld] fp32, the first `W` columns are used (the hook strips the question
// columns when the layer is the question-injection layer).  Two launches, fixed-order reductions.
// ------------------------------------------------------------------------------------------
namespace rn {
constexpr int kExChunks = 32, kExMaxPerLane = 16;      // W <= 512

__global__ void __launch_bounds__(256)
extract_partial_kernel(const float* __restrict__ z, long long P, int ld, int W, float* __restrict__ part) {
  __shared__ float smax[8][512], ssum[8][512];
  const int b = blockIdx.x, chunk = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per = (P + kExChunks - 1) / kExChunks;
  const long long r0 = chunk * per, r1 = min(P, r0 + per);
  float vmax[kExMaxPerLane], vsum[kExMaxPerLane];
#pragma unroll
  for (int k = 0; k < kExMaxPerLane; ++k) { vmax[k] = -INFINITY; vsum[k] = 0.f; }
  for (long long r = r0 + warp; r < r1; r += 8) {
    const float* row = z + ((long long)b * P + r) * ld;
    float v[kExMaxPerLane], ss = 0.f;
#pragma unroll
    for (int k = 0; k < kExMaxPerLane; ++k) {
      const int j = lane + 32 * k;
      v[k] = j < W ? row[j] : 0.f;
      ss = fmaf(v[k], v[k], ss);
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);        // F.normalize(p=2, eps=1e-12)
#pragma unroll
    for (int k = 0; k < kExMaxPerLane; ++k) {
      const float x = v[k] * inv;
      vmax[k] = fmaxf(vmax[k], x);
      vsum[k] += x;
    }
  }
#pragma unroll
  for (int k = 0; k < kExMaxPerLane; ++k) { smax[warp][lane + 32 * k] = vmax[k]; ssum[warp][lane + 32 * k] = vsum[k]; }
  __syncthreads();
  for (int j = threadIdx.x; j < W; j += 256) {
    float m = smax[0][j], s = ssum[0][j];
#pragma unroll
    for (int wv = 1; wv < 8; ++wv) { m = fmaxf(m, smax[wv][j]); s += ssum[wv][j]; }
    float* o = part + (((size_t)b * kExChunks + chunk) * 2) * W;
    o[j] = m;
    o[W + j] = s;
  }
}

__global__ void extract_final_kernel(const float* __restrict__ part, long long P, int W, float* __restrict__ maxf,
                                     float* __restrict__ avgf) {
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float m = -INFINITY, s = 0.f;
    for (int c = 0; c < kExChunks; ++c) {
      const float* o = part + (((size_t)b * kExChunks + c) * 2) * W;
      m = fmaxf(m, o[j]);
      s += o[W + j];
    }
    maxf[(size_t)b * W + j] = m;
    avgf[(size_t)b * W + j] = s / (float)P;
  }
}
}  // namespace rn

extern "C" int rn_extract_stats(const float* z, int B, long long P, int ld, int W, float* maxf, float* avgf, float* scratch,
                                void* stream) {
  RN_CHECK_ARG(z && maxf && avgf && scratch, "NULL pointer argument");
  RN_CHECK_ARG(B > 0 && B <= 65535 && P > 0 && W > 0 && W <= ld && W <= 32 * rn::kExMaxPerLane, "bad shape (B=%d P=%lld ld=%d W=%d)", B,
               P, ld, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rn::extract_partial_kernel<<<dim3(B, rn::kExChunks), 256, 0, st>>>(z, P, ld, W, scratch);
  RN_LAUNCH_CHECK("extract_partial_kernel");
  rn::extract_final_kernel<<<B, 256, 0, st>>>(scratch, P, W, maxf, avgf);
  RN_LAUNCH_CHECK("extract_final_kernel");
  return RN_OK;
}

// Device pointer of the materialised activation H_{l+1} = relu(layer l) [B * n * n, G] inside the `saved` buffer of an
// RN_PRECISION_FP32 training-mode forward (l in [0, L)): what a forward hook on g layer l + 1 receives as its input.
extern "C" int rn_relation_activation(const rn_relation_cfg* cfg, const void* saved, int l, const float** out) {
  RN_TRY(validate_cfg(cfg));
  RN_CHECK_ARG(saved && out, "NULL pointer argument");
  RN_CHECK_ARG(cfg->precision == RN_PRECISION_FP32 && cfg->training != 0, "activations are materialised by the fp32 training-mode forward only");
  RN_CHECK_ARG(l >= 0 && l < cfg->L, "l must be in [0, L) (l=%d)", l);
  RelShape s(*cfg);
  SimtSaved sv = simt_carve_saved(s, const_cast<void*>(saved), s.L);
  *out = sv.H[l];
  return RN_OK;
}
