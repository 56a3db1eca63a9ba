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
