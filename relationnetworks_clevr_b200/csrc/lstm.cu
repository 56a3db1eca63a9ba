// lstm.cu -- question encoder: Embedding -> 1-layer LSTM (zero initial state, batch_first) -> final hidden state
// (reference model.py:39-58), forward and backward, fp32 throughout, for hidden size 128.
//
// Why a kernel: the reference's T sequential time steps are T x (GEMM + pointwise) launches in any framework; at
// 80 questions per GPU (the 8-GPU split of the batch-640 step) that latency chain is as long as the whole g-MLP.  Here
// one persistent launch runs all T steps:
//   * the input projection is a TABLE: the vocabulary has V ~ 83 words, so P[v] = emb[v] W_ih^T + b_ih + b_hh is computed
//     once per call (V x 512 dot products of length E) and a time step only gathers a row of it;
//   * the recurrent matrix W_hh (512 x 128 fp32 = 256 KB = one SM's whole register file) is split over a CLUSTER OF TWO
//     CTAs (256 threads each, one gate row = 128 registers per thread): CTA r owns hidden units [64r, 64r + 64) and their
//     four gate rows, so the cell update is CTA-local and only the 64 new h values per sample cross to the peer, written
//     straight into its shared memory (DSMEM) before one cluster barrier per time step;
//   * samples are independent: cluster c takes S = ceil(B / clusters) of them and reuses every weight register S times.
// Backward is the same machine run from t = T-1 down: the pre-activation gradients go to HBM ([T*B, 512], 26 MB at
// B = 640, T = 20) for the two weight-gradient products; dh_{t-1} = dpre W_hh uses a transposed register layout (thread
// = output unit k x half of the CTA's gate rows) whose four partial sums per unit meet in the owner CTA's shared memory.
// dW_hh = dpre^T h_{t-1} is one split-K SIMT GEMM; the embedding / W_ih / bias gradients go through the table: dP[v] =
// sum of dpre rows whose token is v (fixed order -> deterministic, no atomics, no sort), then three tiny products.
#include "common.cuh"
#include "sgemm.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace rn {

using namespace ptx;

constexpr int kLH = 128;            // hidden size
constexpr int kLG = 4 * kLH;        // gate rows (i, f, g, o)
constexpr int kLSMax = 16;          // samples per cluster
constexpr int kLTMax = 64;          // time steps (tokens of a sample are staged in shared memory)
constexpr int kLThreads = 256;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

// P[v, r] = b_ih[r] + b_hh[r] + sum_e emb[v, e] w_ih[r, e]
__global__ void __launch_bounds__(kLG) lstm_table_kernel(const float* __restrict__ emb, const float* __restrict__ w_ih,
                                                          const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                                                          int E, float* __restrict__ P) {
  extern __shared__ float erow[];
  const int v = blockIdx.x, r = threadIdx.x;
  for (int e = threadIdx.x; e < E; e += blockDim.x) erow[e] = emb[(size_t)v * E + e];
  __syncthreads();
  float acc = b_ih[r] + b_hh[r];
  const float* w = w_ih + (size_t)r * E;
  for (int e = 0; e < E; ++e) acc = fmaf(erow[e], w[e], acc);
  P[(size_t)v * kLG + r] = acc;
}

struct LstmShared {
  float h[2][kLSMax][kLH];          // h_{t-1} / h_t of the cluster's samples (both CTAs hold all 128 units)
  float g[kLSMax][4][64];           // this CTA's activated gates (forward) / pre-activation gradients (backward)
  float red[2][kLSMax][4][64];      // backward: partial dh_{t-1} of this CTA's units (2 halves x 2 CTAs), double-buffered
  int tok[kLSMax][kLTMax];
};

// ---------------------------------------------------------------------------------------------------------------
// forward.  grid = 2 * clusters, cluster (2,1,1), 256 threads.  thread t of CTA r: gate (t >> 6), unit 64 r + (t & 63).
// saved (training): act [T, B, 512] activated gates, cs [T, B, 128] cell states, hs [T, B, 128] hidden states.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLThreads, 1)
lstm_fwd_kernel(const long long* __restrict__ tokens, const float* __restrict__ P, const float* __restrict__ w_hh, int B, int T,
                int V, int S, float* __restrict__ q_out, float* __restrict__ act, float* __restrict__ cs,
                float* __restrict__ hs) {
  extern __shared__ __align__(16) char lstm_smem[];
  LstmShared& sm = *reinterpret_cast<LstmShared*>(lstm_smem);
  const uint32_t rank = cluster_ctarank();
  const int t = threadIdx.x, gate = t >> 6, jj = t & 63;
  const int row = gate * kLH + (int)rank * 64 + jj;
  const int s0 = (int)(blockIdx.x >> 1) * S;
  const int ns = min(S, B - s0);
  if (ns <= 0) {                 // (whole cluster: both CTAs share s0) still take part in the cluster barriers below
    for (int step = 0; step < T; ++step) cluster_sync_all();
    return;
  }
  float w[kLH];
#pragma unroll
  for (int k = 0; k < kLH; k += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)row * kLH + k));
    w[k] = v.x; w[k + 1] = v.y; w[k + 2] = v.z; w[k + 3] = v.w;
  }
  for (int i = t; i < ns * T; i += kLThreads) {
    const long long tk = tokens[(size_t)(s0 + i / T) * T + i % T];
    sm.tok[i / T][i % T] = (int)min(max(tk, 0LL), (long long)(V - 1));
  }
  for (int i = t; i < kLSMax * kLH; i += kLThreads) (&sm.h[0][0][0])[i] = 0.f;
  float c_reg[kLSMax * 64 / kLThreads];
#pragma unroll
  for (int m = 0; m < kLSMax * 64 / kLThreads; ++m) c_reg[m] = 0.f;
  __syncthreads();
  const uint32_t peer = rank ^ 1u;

  for (int step = 0; step < T; ++step) {
    const int cur = step & 1;
    float pre[kLSMax];
#pragma unroll
    for (int s = 0; s < kLSMax; ++s) pre[s] = s < ns ? __ldg(P + (size_t)sm.tok[s][step] * kLG + row) : 0.f;
#pragma unroll 1
    for (int s = 0; s < ns; ++s) {
      const float4* hp = reinterpret_cast<const float4*>(sm.h[cur][s]);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < kLH / 4; ++k4) {
        const float4 hv = hp[k4];
        a0 = fmaf(w[4 * k4], hv.x, a0);
        a1 = fmaf(w[4 * k4 + 1], hv.y, a1);
        a2 = fmaf(w[4 * k4 + 2], hv.z, a2);
        a3 = fmaf(w[4 * k4 + 3], hv.w, a3);
      }
      float z = 0.f;
#pragma unroll
      for (int ss = 0; ss < kLSMax; ++ss) if (ss == s) z = pre[ss];
      z += (a0 + a1) + (a2 + a3);
      const float a = gate == 2 ? tanhf(z) : sigmoidf_acc(z);
      sm.g[s][gate][jj] = a;
      if (act) act[((size_t)step * B + s0 + s) * kLG + row] = a;
    }
    __syncthreads();
    // cell update of this CTA's 64 units: item = (sample, unit), fixed thread <-> item assignment (c stays in registers)
#pragma unroll
    for (int m = 0; m < kLSMax * 64 / kLThreads; ++m) {
      const int item = t + m * kLThreads, s = item >> 6, u = item & 63;
      if (s < ns) {
        const float ig = sm.g[s][0][u], fg = sm.g[s][1][u], gg = sm.g[s][2][u], og = sm.g[s][3][u];
        const float c = fmaf(fg, c_reg[m], ig * gg);
        c_reg[m] = c;
        const float h = og * tanhf(c);
        const int unit = (int)rank * 64 + u;
        float* dst = &sm.h[cur ^ 1][s][unit];
        *dst = h;
        st_cluster_f32(mapa_u32(smem_u32(dst), peer), h);
        const size_t o = ((size_t)step * B + s0 + s) * kLH + unit;
        if (cs) { cs[o] = c; hs[o] = h; }
        if (step == T - 1) q_out[(size_t)(s0 + s) * kLH + unit] = h;
      }
    }
    cluster_sync_all();            // h_t complete in both CTAs; sm.g may be overwritten
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward.  Same grid.  dG [T, B, 512] receives the pre-activation gradients.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLThreads, 1)
lstm_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ act, const float* __restrict__ cs,
                const float* __restrict__ w_hh, int B, int T, int S, float* __restrict__ dG) {
  extern __shared__ __align__(16) char lstm_smem[];
  LstmShared& sm = *reinterpret_cast<LstmShared*>(lstm_smem);
  const uint32_t rank = cluster_ctarank();
  const int t = threadIdx.x;
  const int s0 = (int)(blockIdx.x >> 1) * S;
  const int ns = min(S, B - s0);
  if (ns <= 0) {
    for (int step = 0; step < T; ++step) cluster_sync_all();
    return;
  }
  // transposed layout: thread = (output unit k, half hh of this CTA's 256 gate rows); local row lr = gate * 64 + unit
  const int k = t & 127, hh = t >> 7;
  float wt[kLH];
#pragma unroll
  for (int m = 0; m < kLH; ++m) {
    const int lr = hh * 128 + m;
    const int grow = (lr >> 6) * kLH + (int)rank * 64 + (lr & 63);
    wt[m] = __ldg(w_hh + (size_t)grow * kLH + k);
  }
  float dc_reg[kLSMax * 64 / kLThreads];
#pragma unroll
  for (int m = 0; m < kLSMax * 64 / kLThreads; ++m) dc_reg[m] = 0.f;
  for (int i = t; i < 2 * kLSMax * 4 * 64; i += kLThreads) (&sm.red[0][0][0][0])[i] = 0.f;
  __syncthreads();
  const uint32_t owner = (uint32_t)(k >> 6);          // CTA that owns output unit k

  for (int step = T - 1; step >= 0; --step) {
    const int rb = step & 1;                           // partial sums written at step + 1 sit in buffer (step + 1) & 1
    // ---- pointwise: item = (sample, unit of this CTA) ----
#pragma unroll
    for (int m = 0; m < kLSMax * 64 / kLThreads; ++m) {
      const int item = t + m * kLThreads, s = item >> 6, u = item & 63;
      if (s < ns) {
        const int unit = (int)rank * 64 + u;
        float dh;
        if (step == T - 1) dh = dq[(size_t)(s0 + s) * kLH + unit];
        else dh = (sm.red[rb ^ 1][s][0][u] + sm.red[rb ^ 1][s][1][u]) + (sm.red[rb ^ 1][s][2][u] + sm.red[rb ^ 1][s][3][u]);
        const size_t o = ((size_t)step * B + s0 + s);
        const float* a = act + o * kLG + unit;
        const float ig = a[0], fg = a[kLH], gg = a[2 * kLH], og = a[3 * kLH];
        const float c = cs[o * kLH + unit];
        const float cprev = step > 0 ? cs[(o - B) * kLH + unit] : 0.f;
        const float tc = tanhf(c);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dc_reg[m]);
        dc_reg[m] = dc * fg;
        const float di = dc * gg * ig * (1.f - ig);
        const float df = dc * cprev * fg * (1.f - fg);
        const float dg = dc * ig * (1.f - gg * gg);
        const float dO = dh * tc * og * (1.f - og);
        sm.g[s][0][u] = di; sm.g[s][1][u] = df; sm.g[s][2][u] = dg; sm.g[s][3][u] = dO;
        float* d = dG + o * kLG + unit;
        d[0] = di; d[kLH] = df; d[2 * kLH] = dg; d[3 * kLH] = dO;
      }
    }
    __syncthreads();
    if (step > 0) {
      // ---- dh_{t-1}[k] partial over this thread's 128 gate rows, sent to the CTA that owns unit k ----
#pragma unroll 1
      for (int s = 0; s < ns; ++s) {
        const float4* dp = reinterpret_cast<const float4*>(&sm.g[s][0][0]) + hh * 32;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int m4 = 0; m4 < kLH / 4; ++m4) {
          const float4 dv = dp[m4];
          a0 = fmaf(wt[4 * m4], dv.x, a0);
          a1 = fmaf(wt[4 * m4 + 1], dv.y, a1);
          a2 = fmaf(wt[4 * m4 + 2], dv.z, a2);
          a3 = fmaf(wt[4 * m4 + 3], dv.w, a3);
        }
        const float part = (a0 + a1) + (a2 + a3);
        float* dst = &sm.red[rb][s][(int)rank * 2 + hh][k & 63];
        if (owner == rank) *dst = part;
        else st_cluster_f32(mapa_u32(smem_u32(dst), owner), part);
      }
    }
    cluster_sync_all();
  }
}

// dP[v, r] = sum over positions p = t * B + s with tokens[s, t] == v of dG[p, r].  Two levels, both in a fixed order
// (deterministic, no atomics, no sort): block (v, sl) of lstm_dp_kernel takes slice sl of kDpSlices equal slices of the
// positions -- its 16 warps compact the hits of 16 sub-slices into shared-memory lists, then thread r sums column r of the
// listed rows -- and lstm_dp_reduce_kernel adds the kDpSlices partials.  One block per word (the first version) is a
// performance cliff on real questions: the padding word holds a third of all positions, and its block summed thousands of
// rows as one chain of dependent loads.
constexpr int kDpSlices = 16;

__global__ void __launch_bounds__(kLG)
lstm_dp_kernel(const long long* __restrict__ tokens, const float* __restrict__ dG, int B, int T, int V, int slice, int sub,
               float* __restrict__ part) {
  extern __shared__ int dp_list[];              // [16 warps][sub]
  __shared__ int counts[16];
  const int v = blockIdx.x, sl = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = B * T;
  int cnt = 0;
  const int s_end = min(n, (sl + 1) * slice);
  const int p_begin = sl * slice + warp * sub, p_end = min(s_end, p_begin + sub);
  for (int base = p_begin; base < p_end; base += 32) {
    const int p = base + lane;
    bool hit = false;
    if (p < p_end) {
      const int tt = p / B, s = p - tt * B;
      long long tk = tokens[(size_t)s * T + tt];
      tk = min(max(tk, 0LL), (long long)(V - 1));
      hit = (int)tk == v;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (hit) dp_list[warp * sub + cnt + __popc(m & ((1u << lane) - 1u))] = p;
    cnt += __popc(m);
  }
  if (lane == 0) counts[warp] = cnt;
  __syncthreads();
  const int r = threadIdx.x;
  float acc = 0.f;
  for (int wv = 0; wv < 16; ++wv) {
    const int c = counts[wv];
    const int* l = dp_list + wv * sub;
    int i = 0;
    for (; i + 8 <= c; i += 8) {
      float x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) x[u] = dG[(size_t)l[i + u] * kLG + r];
      acc += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    }
    for (; i < c; ++i) acc += dG[(size_t)l[i] * kLG + r];
  }
  part[((size_t)v * kDpSlices + sl) * kLG + r] = acc;
}

__global__ void __launch_bounds__(kLG) lstm_dp_reduce_kernel(const float* __restrict__ part, float* __restrict__ dP) {
  const int v = blockIdx.x, r = threadIdx.x;
  float x[kDpSlices];
#pragma unroll
  for (int sl = 0; sl < kDpSlices; ++sl) x[sl] = part[((size_t)v * kDpSlices + sl) * kLG + r];
  float acc = 0.f;
#pragma unroll
  for (int sl = 0; sl < kDpSlices; ++sl) acc += x[sl];
  dP[(size_t)v * kLG + r] = acc;
}

// blocks [0, 512): dW_ih[r, :] = sum_v dP[v, r] emb[v, :], db_ih[r] = db_hh[r] = sum_v dP[v, r];
// blocks [512, 512 + V): demb[v, :] = sum_r dP[v, r] W_ih[r, :].   256 threads: thread = (column e, part) with 256 / E parts
// that split the sum (a 512-term chain of dependent L2 loads per thread took 29 us with E threads per block); the parts
// are combined in shared memory in a fixed order.
__global__ void __launch_bounds__(256)
lstm_param_grad_kernel(const float* __restrict__ dP, const float* __restrict__ emb, const float* __restrict__ w_ih, int V, int E,
                       float* __restrict__ dw_ih, float* __restrict__ db_ih, float* __restrict__ db_hh,
                       float* __restrict__ demb) {
  __shared__ float red[256], redb[256];
  const int parts = 256 / E, e = threadIdx.x % E, part = threadIdx.x / E;
  const bool active = part < parts;
  float acc = 0.f, bsum = 0.f;
  const bool row_block = (int)blockIdx.x < kLG;
  if (active) {
    if (row_block) {
      const int r = blockIdx.x;
      for (int v = part; v < V; v += parts) {
        const float d = dP[(size_t)v * kLG + r];
        acc = fmaf(d, emb[(size_t)v * E + e], acc);
        bsum += d;
      }
    } else {
      const int v = blockIdx.x - kLG;
#pragma unroll 4
      for (int r = part; r < kLG; r += parts) acc = fmaf(dP[(size_t)v * kLG + r], w_ih[(size_t)r * E + e], acc);
    }
  }
  red[threadIdx.x] = acc;
  redb[threadIdx.x] = bsum;
  __syncthreads();
  if (part == 0) {
    float a = 0.f, bs = 0.f;
    for (int q = 0; q < parts; ++q) { a += red[q * E + e]; bs += redb[q * E + e]; }
    if (row_block) {
      dw_ih[(size_t)blockIdx.x * E + e] = a;
      if (e == 0) { db_ih[blockIdx.x] = bs; db_hh[blockIdx.x] = bs; }
    } else {
      demb[(size_t)(blockIdx.x - kLG) * E + e] = a;
    }
  }
}

static bool lstm_ok(const rn_lstm_cfg* c) {
  return c && c->H == kLH && c->E >= 1 && c->E <= 256 && c->T >= 1 && c->T <= kLTMax && c->V >= 1 && c->B >= 1;
}

static int lstm_clusters(int B, int* S) {
  const int max_clusters = std::max(1, sm_count() / 2);
  int s = cdiv(B, max_clusters);
  s = std::max(1, std::min(s, kLSMax));
  *S = s;
  return cdiv(B, s);
}

static size_t lstm_splitk_floats() { return (size_t)24 * kLG * kLH; }

template <typename... Args>
static int launch_cluster2(void (*kernel)(Args...), int clusters, size_t smem, cudaStream_t st, Args... args) {
  RN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kLThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RN_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  return RN_OK;
}

}  // namespace rn

using namespace rn;

extern "C" int rn_lstm_supported(const rn_lstm_cfg* cfg) { return lstm_ok(cfg) ? 1 : 0; }

extern "C" int rn_lstm_workspace(const rn_lstm_cfg* cfg, size_t* saved_floats, size_t* scratch_floats) {
  RN_CHECK_ARG(cfg && saved_floats && scratch_floats, "NULL pointer argument");
  if (!lstm_ok(cfg)) return fail(RN_ERR_UNSUPPORTED, "rn_lstm_*: needs H == 128, 1 <= T <= 64 (H=%d T=%d)", cfg->H, cfg->T);
  const size_t tb = (size_t)cfg->T * cfg->B;
  *saved_floats = round_up((size_t)cfg->V * kLG, 64) + (cfg->training ? tb * (kLG + 2 * kLH) : 0);
  *scratch_floats = cfg->training ? tb * kLG + round_up((size_t)cfg->V * kLG, 64) + lstm_splitk_floats() +
                                        (size_t)cfg->V * kDpSlices * kLG : 64;
  return RN_OK;
}

extern "C" int rn_lstm_fwd(const rn_lstm_cfg* cfg, const int64_t* tokens, const float* emb, const float* w_ih,
                           const float* w_hh, const float* b_ih, const float* b_hh, float* q, float* saved, void* stream) {
  RN_CHECK_ARG(cfg && tokens && emb && w_ih && w_hh && b_ih && b_hh && q && saved, "NULL pointer argument");
  if (!lstm_ok(cfg)) return fail(RN_ERR_UNSUPPORTED, "rn_lstm_fwd: needs H == 128, 1 <= T <= 64 (H=%d T=%d)", cfg->H, cfg->T);
  RN_CHECK_ARG(aligned16(w_hh) && aligned16(saved), "w_hh and saved must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* P = saved;
  lstm_table_kernel<<<cfg->V, kLG, cfg->E * sizeof(float), st>>>(emb, w_ih, b_ih, b_hh, cfg->E, P);
  RN_LAUNCH_CHECK("lstm_table_kernel");
  const size_t tb = (size_t)cfg->T * cfg->B;
  float* act = cfg->training ? saved + round_up((size_t)cfg->V * kLG, 64) : nullptr;
  float* cs = cfg->training ? act + tb * kLG : nullptr;
  float* hs = cfg->training ? cs + tb * kLH : nullptr;
  int S = 1;
  const int clusters = lstm_clusters(cfg->B, &S);
  RN_TRY(launch_cluster2(lstm_fwd_kernel, clusters, sizeof(LstmShared), st, reinterpret_cast<const long long*>(tokens),
                         (const float*)P, w_hh, (int)cfg->B, (int)cfg->T, (int)cfg->V, S, q, act, cs, hs));
  RN_LAUNCH_CHECK("lstm_fwd_kernel");
  return RN_OK;
}

extern "C" int rn_lstm_bwd(const rn_lstm_cfg* cfg, const int64_t* tokens, const float* emb, const float* w_ih,
                           const float* w_hh, const float* dq, const float* saved, float* demb, float* dw_ih, float* dw_hh,
                           float* db_ih, float* db_hh, float* scratch, void* stream) {
  RN_CHECK_ARG(cfg && tokens && emb && w_ih && w_hh && dq && saved && demb && dw_ih && dw_hh && db_ih && db_hh && scratch,
               "NULL pointer argument");
  if (!lstm_ok(cfg)) return fail(RN_ERR_UNSUPPORTED, "rn_lstm_bwd: needs H == 128, 1 <= T <= 64 (H=%d T=%d)", cfg->H, cfg->T);
  RN_CHECK_ARG(cfg->training != 0, "rn_lstm_bwd needs a cfg with training=1 (same as the forward call)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = cfg->B, T = cfg->T, V = cfg->V, E = cfg->E;
  const size_t tb = (size_t)T * B;
  const float* act = saved + round_up((size_t)V * kLG, 64);
  const float* cs = act + tb * kLG;
  const float* hs = cs + tb * kLH;
  float* dG = scratch;
  float* dP = dG + tb * kLG;
  float* ws = dP + round_up((size_t)V * kLG, 64);
  int S = 1;
  const int clusters = lstm_clusters(B, &S);
  RN_TRY(launch_cluster2(lstm_bwd_kernel, clusters, sizeof(LstmShared), st, dq, act, cs, w_hh, B, T, S, dG));
  RN_LAUNCH_CHECK("lstm_bwd_kernel");
  // dW_hh[r, k] = sum_{t >= 1, s} dG[t, s, r] h_{t-1}[s, k]
  if (T > 1) {
    GemmEpilogue none;
    RN_TRY(sgemm(true, false, kLG, kLH, (T - 1) * B, dG + (size_t)B * kLG, kLG, hs, kLH, dw_hh, kLH, none, st, ws,
                 lstm_splitk_floats()));
  } else {
    RN_CUDA(cudaMemsetAsync(dw_hh, 0, sizeof(float) * kLG * kLH, st));
  }
  // table gradient, then the embedding / W_ih / bias gradients
  float* dp_part = ws + lstm_splitk_floats();                 // [V][kDpSlices][512]
  const int slice = cdiv((long long)tb, kDpSlices);
  const int sub = (cdiv(slice, 16) + 31) / 32 * 32;           // positions per warp, a multiple of the warp size
  const size_t list_bytes = (size_t)16 * sub * sizeof(int);
  RN_CHECK_ARG(list_bytes <= 200 * 1024, "rn_lstm_bwd: T*B = %zu exceeds the position-list capacity", tb);
  RN_CUDA(cudaFuncSetAttribute(lstm_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)list_bytes));
  lstm_dp_kernel<<<dim3(V, kDpSlices), kLG, list_bytes, st>>>(reinterpret_cast<const long long*>(tokens), dG, B, T, V, slice, sub,
                                                               dp_part);
  RN_LAUNCH_CHECK("lstm_dp_kernel");
  lstm_dp_reduce_kernel<<<V, kLG, 0, st>>>(dp_part, dP);
  RN_LAUNCH_CHECK("lstm_dp_reduce_kernel");
  lstm_param_grad_kernel<<<kLG + V, 256, 0, st>>>(dP, emb, w_ih, V, E, dw_ih, db_ih, db_hh, demb);
  RN_LAUNCH_CHECK("lstm_param_grad_kernel");
  return RN_OK;
}
