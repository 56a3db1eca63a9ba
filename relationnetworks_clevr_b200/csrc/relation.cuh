// relation.cuh -- internal interfaces of the relation-layer implementation.
//
// The relation layer (reference model.py:104-152) is computed in the factorised form of
// SURVEY.md 7.4 / DESIGN.md:
//     U = X W0c^T, V' = X W0a^T + beta0[b]          (layer 0, fp32, "pre")
//     H1[a,c] = relu(U[c] + V'[a])                   (generated on the fly, never stored as pairs)
//     H_{l+1} = relu(H_l Wh_l^T + beta_l)            (layers 1..L-1: SIMT fp32 or tcgen05)
//     x_g = sum_{a,c} H_L[a,c]
#pragma once

#include "common.cuh"
#include "sgemm.cuh"

namespace rn {

struct RelShape {
  int B, n, k, Q, G, L, qinj;
  unsigned flags;
  long long pairs;       // n*n
  long long rows;        // B*n*n
  int fan_in(int l) const { return (l == 0 ? 2 * k : G) + (l == qinj ? Q : 0); }
  explicit RelShape(const rn_relation_cfg& c)
      : B(c.B), n(c.n), k(c.k), Q(c.Q), G(c.G), L(c.L), qinj(c.qinj), flags(c.flags), pairs((long long)c.n * c.n),
        rows((long long)c.B * c.n * c.n) {}
};

// layer-0 products shared by every precision mode ("pre"): U [B,n,G], Vb [B,n,G] (V + beta0 folded in),
// Qb [B,G] (per-sample bias of layer qinj when qinj > 0; unused otherwise).
struct RelPre {
  float* U;
  float* Vb;
  float* Qb;
  // optional second copy of U as [B][G/4][n][4] (column groups outermost): a warp whose lanes are consecutive objects
  // reads one column group of 32 objects as 512 contiguous bytes (the 3-pass forward generates its operand row-per-thread)
  float* U4 = nullptr;
};

int relation_pre(const RelShape& s, const float* x, const float* q, const float* const* g_w, const float* const* g_b,
                 const RelPre& pre, cudaStream_t st);

// layer-0 backward shared by every precision mode: from dU [B,n,G], dV [B,n,G] to dx, dq (when
// qinj == 0), dW0, db0.  `delta` [B,G] scratch.
// `ws` / `ws_floats`: split-K workspace for the long-K products (dU^T X, dV^T X: K = B*n).
int relation_layer0_bwd(const RelShape& s, const float* x, const float* q, const float* const* g_w, const float* dU,
                        const float* dV, float* delta, float* dx, float* dq, float* const* dg_w,
                        float* const* dg_b, float* ws, size_t ws_floats, cudaStream_t st);

// question-injection gradients at layer l == qinj > 0 from the per-sample column sums delta [B,G].
int relation_qinj_bwd(const RelShape& s, int l, const float* q, const float* const* g_w, const float* delta,
                      float* dq, float* const* dg_w, cudaStream_t st);

// fp32 SIMT path (any shape)
size_t simt_saved_bytes(const RelShape& s, bool training);
size_t simt_scratch_bytes(const RelShape& s, bool training);
int simt_relation_fwd(const RelShape& s, bool training, const float* x, const float* q, const float* const* g_w,
                      const float* const* g_b, float* xg, void* saved, void* scratch, cudaStream_t st);
int simt_relation_bwd(const RelShape& s, const float* dxg, const float* x, const float* q, const float* const* g_w,
                      const void* saved, float* dx, float* dq, float* const* dg_w, float* const* dg_b, void* scratch,
                      cudaStream_t st);

// tcgen05 path (G == 256, pairs % 128 == 0, L == 4)
bool tc_supported(const RelShape& s);
size_t tc_saved_bytes(const RelShape& s, int precision, bool training);
size_t tc_scratch_bytes(const RelShape& s, bool training);
int tc_relation_fwd(const RelShape& s, int precision, bool training, const float* x, const float* q,
                    const float* const* g_w, const float* const* g_b, float* xg, void* saved, void* scratch,
                    cudaStream_t st);
int tc_relation_bwd(const RelShape& s, int precision, const float* dxg, const float* x, const float* q,
                    const float* const* g_w, const void* saved, float* dx, float* dq, float* const* dg_w,
                    float* const* dg_b, void* scratch, cudaStream_t st);

}  // namespace rn
