// adam.cu -- optimiser tail on flat buffers: global-norm clip + Adam with L2 weight decay
// (reference train.py:45-48 clip_grad_norm(params, 50) and train.py:330 Adam(weight_decay=1e-4)).
// Two launches instead of ~100: a fixed-order two-level norm reduction, then one fused update that
// finishes the norm, derives the clip coefficient and applies Adam.
#include "common.cuh"

#include <algorithm>

namespace rn {

constexpr int kNormBlocks = 1024;

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float scale, float* __restrict__ part, int* __restrict__ d_step) {
  __shared__ double red[8];
  if (d_step && blockIdx.x == 0 && threadIdx.x == 0) *d_step += 1;      // device-side step counter (CUDA-graph replays)
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = (double)g[i] * scale;
    s += v * v;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    part[blockIdx.x] = (float)t;
  }
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(rn_adam_cfg c, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, const float* __restrict__ part, float* __restrict__ total_norm_out,
                 const int* __restrict__ d_step, const float* __restrict__ d_lr) {
  __shared__ float coef_s;
  if (d_step) c.step = *d_step;
  if (d_lr) c.lr = *d_lr;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int i = threadIdx.x; i < kNormBlocks; i += 32) s += part[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(s);
      float coef = 1.f;
      if (c.clip_norm > 0.f) coef = fminf(c.clip_norm / (norm + 1e-6f), 1.f);
      coef_s = coef * c.grad_scale;
      if (blockIdx.x == 0 && total_norm_out) *total_norm_out = norm;
    }
  }
  __syncthreads();
  const float coef = coef_s;
  const float bc1 = 1.f - powf(c.beta1, (float)c.step), bc2 = 1.f - powf(c.beta2, (float)c.step);
  const float step_size = c.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    const float gi = fmaf(c.weight_decay, w, g[i] * coef);
    const float mi = c.beta1 * m[i] + (1.f - c.beta1) * gi;
    const float vi = c.beta2 * v[i] + (1.f - c.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = w - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + c.eps);
  }
}

}  // namespace rn

using namespace rn;

extern "C" int rn_clip_adam(const rn_adam_cfg* cfg, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                            float* norm_scratch, float* total_norm_out, int32_t* d_step, const float* d_lr, void* stream) {
  RN_CHECK_ARG(cfg != nullptr && cfg->n > 0 && (cfg->step >= 1 || d_step != nullptr), "bad adam cfg");
  RN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && norm_scratch, "NULL pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sumsq_partial_kernel<<<kNormBlocks, 256, 0, st>>>(grads, cfg->n, cfg->grad_scale, norm_scratch, d_step);
  RN_LAUNCH_CHECK("sumsq_partial_kernel");
  const int blocks = (int)std::min<long long>((cfg->n + 255) / 256, 4LL * sm_count());
  clip_adam_kernel<<<blocks, 256, 0, st>>>(*cfg, params, grads, exp_avg, exp_avg_sq, norm_scratch, total_norm_out, d_step, d_lr);
  RN_LAUNCH_CHECK("clip_adam_kernel");
  return RN_OK;
}
