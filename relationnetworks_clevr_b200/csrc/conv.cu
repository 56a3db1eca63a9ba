// conv.cu -- the 4-layer feature extractor: 4 x [conv3x3 stride 2 pad 1 -> BatchNorm2d(24) -> ReLU]
// followed by the (x, y) coordinate channels and the [B, n, 26] object layout
// (reference model.py:9-36, 192-201, 208-213).
//
// Layout: NCHW fp32 like the reference.  The BatchNorm affine + ReLU of layer l is never written
// out: conv l stores its RAW output y_l plus per-block partial sums (sum y, sum y^2); a finalize
// kernel turns them into (mean, rstd, scale, shift), and conv l+1 applies relu(scale*y+shift) while
// staging its input patch in shared memory.  All reductions are two-pass (per-block partials, then
// a fixed-order sum) so results are deterministic.
#include "common.cuh"
#include "sgemm.cuh"

#include <algorithm>
#include <cstdlib>

namespace rn {

constexpr int kC = RN_CONV_CH;          // 24 channels everywhere except the RGB input
constexpr int kTile = 16;               // output tile edge per block
constexpr int kPatch = 2 * kTile + 1;   // input patch edge (stride 2, 3x3, pad 1)
constexpr int kChunk = 8;               // input channels staged per pass

struct Affine {                         // per-layer block in `saved`: mean, rstd, scale, shift (24 each)
  float v[4 * kC];
};

__device__ __forceinline__ float act_in(float y, const float* __restrict__ aff, int c) {
  // relu(scale*y + shift) of the previous layer's BatchNorm; aff == nullptr for the RGB image
  return aff ? fmaxf(fmaf(aff[2 * kC + c], y, aff[3 * kC + c]), 0.f) : y;
}

// BatchNorm backward applied on the fly: dy = k0 * (g - k1 - xhat * k2), g = dA where relu(bn(y)) > 0.
// `aff` = (mean, rstd, scale, shift) and `coef` = (k0, k1, k2) per channel (bn_bwd_finalize_kernel).
__device__ __forceinline__ float bn_bwd_dy(float yv, float da, const float* __restrict__ aff, const float* __restrict__ coef,
                                          int c) {
  const float g = fmaf(aff[2 * kC + c], yv, aff[3 * kC + c]) > 0.f ? da : 0.f;
  const float xhat = (yv - aff[c]) * aff[kC + c];
  return coef[c] * (g - coef[kC + c] - xhat * coef[2 * kC + c]);
}

// ---- asynchronous staging (LDGSTS): every thread keeps ALL its copies of a patch in flight at once, so the
// L2/HBM round trip is paid once per patch instead of once per element.  Patch rows are padded to kPW = 36 floats:
// column j holds input column iw0 - 3 + j, which makes every row nine ALIGNED 16-byte vectors (iw0 = 32*tile - 1)
// when the image edge is a multiple of 4 -- 3.7x fewer copies and index computations than element-wise staging
// (these kernels are instruction-issue bound).  Other edges fall back to 4-byte copies into the same layout. ------
constexpr int kPW = 36;                 // padded patch row
constexpr int kPO = 3;                  // patch column of input column iw0
constexpr int kPV = kPW / 4;            // vectors per row

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 4 : 0;                    // src-size 0: nothing is read, the bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stage a [CC][33][36] input patch (zero padding outside the image) and apply the previous layer's
// BatchNorm affine + ReLU in place.  Each thread transforms exactly the elements it copied, so no barrier is
// needed between the wait and the transform; the caller synchronises before the patch is consumed.
template <int CC>
__device__ __forceinline__ void stage_patch_issue(float (*patch)[kPatch][kPW], const float* __restrict__ inb, int ci0,
                                                  int ih0, int iw0, int hin, int tid) {
  if ((hin & 3) == 0) {
#pragma unroll 4
    for (int idx = tid; idx < CC * kPatch * kPV; idx += 256) {
      const int ci = idx / (kPatch * kPV), rem = idx - ci * (kPatch * kPV);
      const int r = rem / kPV, v = rem - r * kPV;
      const int ih = ih0 + r, iw = iw0 - kPO + 4 * v;                  // aligned: a vector is entirely in or out
      const bool valid = (unsigned)ih < (unsigned)hin && (unsigned)iw < (unsigned)hin;
      cp_async16(&patch[ci][r][4 * v], valid ? inb + ((size_t)(ci0 + ci) * hin + ih) * hin + iw : inb, valid);
    }
  } else {
#pragma unroll 4
    for (int idx = tid; idx < CC * kPatch * kPatch; idx += 256) {
      const int ci = idx / (kPatch * kPatch), rem = idx % (kPatch * kPatch);
      const int r = rem / kPatch, c = rem % kPatch;
      const int ih = ih0 + r, iw = iw0 + c;
      const bool valid = ih >= 0 && ih < hin && iw >= 0 && iw < hin;
      cp_async4(&patch[ci][r][c + kPO], valid ? inb + ((size_t)(ci0 + ci) * hin + ih) * hin + iw : inb, valid);
    }
  }
}

template <int CC>
__device__ __forceinline__ void stage_patch_finish(float (*patch)[kPatch][kPW], const float* __restrict__ in_aff,
                                                   int ci0, int ih0, int iw0, int hin, int tid) {
  cp_async_wait_all();
  if (!in_aff) return;
  if ((hin & 3) == 0) {
#pragma unroll 4
    for (int idx = tid; idx < CC * kPatch * kPV; idx += 256) {
      const int ci = idx / (kPatch * kPV), rem = idx - ci * (kPatch * kPV);
      const int r = rem / kPV, v = rem - r * kPV;
      const int ih = ih0 + r, iw = iw0 - kPO + 4 * v;
      if ((unsigned)ih < (unsigned)hin && (unsigned)iw < (unsigned)hin) {      // the zero padding applies to the post-activation tensor
        const float sc = in_aff[2 * kC + ci0 + ci], sh = in_aff[3 * kC + ci0 + ci];
        float4 t = *reinterpret_cast<float4*>(&patch[ci][r][4 * v]);
        t.x = fmaxf(fmaf(sc, t.x, sh), 0.f);
        t.y = fmaxf(fmaf(sc, t.y, sh), 0.f);
        t.z = fmaxf(fmaf(sc, t.z, sh), 0.f);
        t.w = fmaxf(fmaf(sc, t.w, sh), 0.f);
        *reinterpret_cast<float4*>(&patch[ci][r][4 * v]) = t;
      }
    }
  } else {
#pragma unroll 4
    for (int idx = tid; idx < CC * kPatch * kPatch; idx += 256) {
      const int ci = idx / (kPatch * kPatch), rem = idx % (kPatch * kPatch);
      const int r = rem / kPatch, c = rem % kPatch;
      const int ih = ih0 + r, iw = iw0 + c;
      if (ih >= 0 && ih < hin && iw >= 0 && iw < hin)
        patch[ci][r][c + kPO] = act_in(patch[ci][r][c + kPO], in_aff, ci0 + ci);
    }
  }
}

// uint8 images ([B,3,S,S] bytes, the pixels torchvision's ToTensor divides by 255): converted while staging, so the host
// copies a quarter of the bytes.  x = u / 255 with a true division -- bit-identical to ToTensor's float().div(255).
template <int CC>
__device__ __forceinline__ void stage_patch_u8(float (*patch)[kPatch][kPW], const unsigned char* __restrict__ inb, int ih0,
                                               int iw0, int hin, int tid) {
  if ((hin & 3) == 0) {
    for (int idx = tid; idx < CC * kPatch * kPV; idx += 256) {
      const int ci = idx / (kPatch * kPV), rem = idx - ci * (kPatch * kPV);
      const int r = rem / kPV, v = rem - r * kPV;
      const int ih = ih0 + r, iw = iw0 - kPO + 4 * v;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)ih < (unsigned)hin && (unsigned)iw < (unsigned)hin) {
        const uchar4 u = *reinterpret_cast<const uchar4*>(inb + ((size_t)ci * hin + ih) * hin + iw);
        t = make_float4(__fdiv_rn((float)u.x, 255.f), __fdiv_rn((float)u.y, 255.f), __fdiv_rn((float)u.z, 255.f),
                        __fdiv_rn((float)u.w, 255.f));
      }
      *reinterpret_cast<float4*>(&patch[ci][r][4 * v]) = t;
    }
  } else {
    for (int idx = tid; idx < CC * kPatch * kPatch; idx += 256) {
      const int ci = idx / (kPatch * kPatch), rem = idx % (kPatch * kPatch);
      const int r = rem / kPatch, c = rem % kPatch;
      const int ih = ih0 + r, iw = iw0 + c;
      const bool valid = ih >= 0 && ih < hin && iw >= 0 && iw < hin;
      patch[ci][r][c + kPO] = valid ? __fdiv_rn((float)inb[((size_t)ci * hin + ih) * hin + iw], 255.f) : 0.f;
    }
  }
}

template <int CC>
__device__ __forceinline__ void stage_patch(float (*patch)[kPatch][kPW], const float* __restrict__ inb,
                                            const float* __restrict__ in_aff, int ci0, int ih0, int iw0, int hin,
                                            int tid) {
  stage_patch_issue<CC>(patch, inb, ci0, ih0, iw0, hin, tid);
  stage_patch_finish<CC>(patch, in_aff, ci0, ih0, iw0, hin, tid);
}

// ------------------------------------------------------------------------------------------
// forward conv: y = conv(act(in)) + bias ; optional per-block (sum, sumsq) partials
// grid (tiles, B), block 256 = 64 quads (2x2 output pixels) x 4 channel groups (6 channels):
// 24 accumulators per thread; per input channel 25 input LDS + 18 broadcast weight LDS.128 feed 216 FFMA.
// Warp = 8 quads (x) x 4 channel groups: input reads hit 8 distinct banks and broadcast over the groups.
// ------------------------------------------------------------------------------------------
template <int CIN, bool U8 = false>
__global__ void __launch_bounds__(256)
conv_fwd_kernel(const float* __restrict__ in, const float* __restrict__ in_aff, const float* __restrict__ w,
                const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ stat_part, int hin,
                int hout, int tiles_x) {
  // 24-channel layers: 4-channel chunks, double buffered -- the copies of chunk k+1 are in flight while
  // chunk k is consumed.  The RGB layer is a single 3-channel chunk.
  constexpr int CC = CIN < 4 ? CIN : 4;
  constexpr int NBUF = CIN > CC ? 2 : 1;
  constexpr int NCHUNK = CIN / CC;
  __shared__ __align__(16) float patch[NBUF][CC][kPatch][kPW];
  __shared__ __align__(16) float wsm[NBUF][CC][9][kC];      // [ci][tap][co]: a thread's 6 channels are 3 adjacent pairs
  __shared__ float red[8][2 * kC];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qx = lane & 7, cg = lane >> 3, qy = warp;       // quad (qy, qx), channels [6*cg, 6*cg + 6)
  const int b = blockIdx.y;
  const int oh0 = (blockIdx.x / tiles_x) * kTile, ow0 = (blockIdx.x % tiles_x) * kTile;
  const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
  const float* inb = U8 ? in : in + (size_t)b * CIN * hin * hin;
  const unsigned char* inb8 = reinterpret_cast<const unsigned char*>(in) + (size_t)b * CIN * hin * hin;

  // packed fp32x2 accumulators (FFMA2: two FMAs per issue slot on sm_100): acc[pixel][channel pair]
  float2 acc[4][3];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[p][c] = make_float2(0.f, 0.f);

  auto issue = [&](int k) {
    const int buf = k % NBUF, ci0 = k * CC;
    if (U8) stage_patch_u8<CC>(patch[buf], inb8, ih0, iw0, hin, tid);
    else stage_patch_issue<CC>(patch[buf], inb, ci0, ih0, iw0, hin, tid);
    for (int idx = tid; idx < CC * kC * 9; idx += 256) {
      const int ci = idx / (kC * 9), rem = idx % (kC * 9);
      const int co = rem / 9, t = rem % 9;
      cp_async4(&wsm[buf][ci][t][co], w + ((size_t)co * CIN + ci0 + ci) * 9 + t, true);
    }
  };

  issue(0);
  for (int k = 0; k < NCHUNK; ++k) {
    const int buf = k % NBUF;
    if (U8) cp_async_wait_all();      // (the weights)
    else stage_patch_finish<CC>(patch[buf], in_aff, k * CC, ih0, iw0, hin, tid);
    __syncthreads();              // chunk k is complete; everyone is done reading the other buffer
    if (k + 1 < NCHUNK) issue(k + 1);
    if (oh0 + 2 * qy >= hout) continue;      // this warp's quad row is outside the image (small layers): staging only
#pragma unroll 1
    for (int ci = 0; ci < CC; ++ci) {
      float xv[5][5];
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) xv[r][c] = patch[buf][ci][4 * qy + r][4 * qx + c + kPO];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wr = &wsm[buf][ci][kh * 3 + kw][cg * 6];
          const float2 w01 = *reinterpret_cast<const float2*>(wr);
          const float2 w23 = *reinterpret_cast<const float2*>(wr + 2);
          const float2 w45 = *reinterpret_cast<const float2*>(wr + 4);
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
              const float x = xv[2 * py + kh][2 * px + kw];
              const float2 xx = make_float2(x, x);
              acc[py * 2 + px][0] = __ffma2_rn(xx, w01, acc[py * 2 + px][0]);
              acc[py * 2 + px][1] = __ffma2_rn(xx, w23, acc[py * 2 + px][1]);
              acc[py * 2 + px][2] = __ffma2_rn(xx, w45, acc[py * 2 + px][2]);
            }
        }
    }
  }

  const int oh = oh0 + 2 * qy, ow = ow0 + 2 * qx;
  const bool valid = oh < hout && ow < hout;        // hout is even: a quad is entirely in or out
  float* yb = y + (size_t)b * kC * hout * hout;
  float s1[6], s2[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int co = cg * 6 + c;
    const float bv = bias[co];
    const float v00 = (c & 1 ? acc[0][c >> 1].y : acc[0][c >> 1].x) + bv, v01 = (c & 1 ? acc[1][c >> 1].y : acc[1][c >> 1].x) + bv;
    const float v10 = (c & 1 ? acc[2][c >> 1].y : acc[2][c >> 1].x) + bv, v11 = (c & 1 ? acc[3][c >> 1].y : acc[3][c >> 1].x) + bv;
    if (valid) {
      *reinterpret_cast<float2*>(&yb[((size_t)co * hout + oh) * hout + ow]) = make_float2(v00, v01);
      *reinterpret_cast<float2*>(&yb[((size_t)co * hout + oh + 1) * hout + ow]) = make_float2(v10, v11);
    }
    s1[c] = valid ? (v00 + v01) + (v10 + v11) : 0.f;
    s2[c] = valid ? (v00 * v00 + v01 * v01) + (v10 * v10 + v11 * v11) : 0.f;
  }
  if (stat_part) {
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      float a = s1[c], a2 = s2[c];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {                 // over the 8 quads that share this channel group
        a += __shfl_xor_sync(0xffffffffu, a, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (qx == 0) { red[warp][cg * 6 + c] = a; red[warp][kC + cg * 6 + c] = a2; }
    }
    __syncthreads();
    if (tid < 2 * kC) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) s += red[wv][tid];
      stat_part[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * kC + tid] = s;
    }
  }
}

// one warp per channel: fixed-order double sum of the per-block partials -> mean/rstd/scale/shift,
// running-stat update (momentum, unbiased variance) as nn.BatchNorm2d does in training.
__global__ void bn_finalize_kernel(const float* __restrict__ part, int nblk, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ aff, float eps, float momentum,
                                   int training) {
  // one 256-thread block per channel; fixed-order (thread-strided, then tree) double-precision reduction
  __shared__ double sh[2][8];
  const int c = blockIdx.x, lane = threadIdx.x % 32, wrp = threadIdx.x / 32;
  float mean, var;
  if (training) {
    double s = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 256) {
      s += part[(size_t)i * 2 * kC + c];
      s2 += part[(size_t)i * 2 * kC + kC + c];
    }
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { sh[0][wrp] = s; sh[1][wrp] = s2; }
    __syncthreads();
    s = 0.0; s2 = 0.0;
    for (int i = 0; i < 8; ++i) { s += sh[0][i]; s2 += sh[1][i]; }
    const double m = s / count;
    double v = s2 / count - m * m;
    if (v < 0) v = 0;
    mean = (float)m;
    var = (float)v;
    if (threadIdx.x == 0) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(v * (count / (count > 1 ? count - 1 : 1)));
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  if (threadIdx.x == 0) {
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    aff[c] = mean;
    aff[kC + c] = rstd;
    aff[2 * kC + c] = sc;
    aff[3 * kC + c] = beta[c] - mean * sc;
  }
}

// objects[b, hw, 0..23] = relu(scale*y4 + shift), objects[b, hw, 24/25] = x/y coords (model.py:208-213)
__global__ void objects_fwd_kernel(const float* __restrict__ y, const float* __restrict__ aff, float* __restrict__ obj,
                                   int d, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = i % 26;
  const long long r = i / 26;
  const int hw = r % (d * d);
  const long long b = r / (d * d);
  float v;
  if (f < kC) {
    v = fmaxf(fmaf(aff[2 * kC + f], y[((size_t)b * kC + f) * d * d + hw], aff[3 * kC + f]), 0.f);
  } else {
    const int idx = (f == kC) ? hw % d : hw / d;      // channel 24 varies along W, 25 along H
    v = d > 1 ? (-d / 2.f + idx * ((float)d / (d - 1))) : -d / 2.f;   // linspace(-d/2, d/2, d)
  }
  obj[i] = v;
}

// dA[b, c, hw] = dobjects[b, hw, c]
__global__ void objects_bwd_kernel(const float* __restrict__ dobj, float* __restrict__ dA, int d, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int hw = i % (d * d);
  const long long r = i / (d * d);
  const int c = r % kC;
  const long long b = r / kC;
  dA[i] = dobj[((size_t)b * d * d + hw) * 26 + c];
}

// ------------------------------------------------------------------------------------------
// BatchNorm backward.  g = dA * (scale*y+shift > 0);  partial[b][c] = (sum g, sum g*xhat)
// grid (24, B), block 256
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ y, const float* __restrict__ dA, const float* __restrict__ aff,
                     float* __restrict__ part, int hw) {
  __shared__ float red[8][2];
  const int c = blockIdx.x, b = blockIdx.y;
  const float mean = aff[c], rstd = aff[kC + c], sc = aff[2 * kC + c], sh = aff[3 * kC + c];
  const size_t base = ((size_t)b * kC + c) * hw;
  float s = 0.f, sx = 0.f;
  for (int i = threadIdx.x; i < hw; i += 256) {
    const float yv = y[base + i];
    const float g = fmaf(sc, yv, sh) > 0.f ? dA[base + i] : 0.f;
    s += g;
    sx += g * (yv - mean) * rstd;
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
  }
  if (threadIdx.x % 32 == 0) { red[threadIdx.x / 32][0] = s; red[threadIdx.x / 32][1] = sx; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    part[((size_t)b * kC + c) * 2 + threadIdx.x] = v;
  }
}

// warp per channel: dbeta, dgamma and the coefficients of dy = k0 * (g - k1 - xhat * k2)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int B, double count, const float* __restrict__ gamma,
                                       const float* __restrict__ aff, float* __restrict__ coef, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ dbias, int training) {
  const int c = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (c >= kC) return;
  double s = 0.0, sx = 0.0;
  for (int b = lane; b < B; b += 32) {
    s += part[((size_t)b * kC + c) * 2];
    sx += part[((size_t)b * kC + c) * 2 + 1];
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
  }
  if (lane == 0) {
    dbeta[c] = (float)s;
    dgamma[c] = (float)sx;
    const float k0 = gamma[c] * aff[kC + c];
    coef[c] = k0;
    coef[kC + c] = training ? (float)(s / count) : 0.f;
    coef[2 * kC + c] = training ? (float)(sx / count) : 0.f;
    // the conv bias feeds straight into BatchNorm: with batch statistics its gradient is exactly 0
    dbias[c] = training ? 0.f : k0 * (float)s;
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ dA, const float* __restrict__ aff,
                                    const float* __restrict__ coef, float* __restrict__ dy, int hw, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (i / hw) % kC;
  const float yv = y[i];
  const float g = fmaf(aff[2 * kC + c], yv, aff[3 * kC + c]) > 0.f ? dA[i] : 0.f;
  const float xhat = (yv - aff[c]) * aff[kC + c];
  dy[i] = coef[c] * (g - coef[kC + c] - xhat * coef[2 * kC + c]);
}

// ------------------------------------------------------------------------------------------
// dy staging shared by the weight- and data-gradient kernels: dy = BatchNorm-backward(dA, y) of a
// [24][ROWS][COLS] window (zero outside the image), one aligned float4 of y and dA per item when the output edge is
// a multiple of 4 (shift-only index arithmetic), scalars otherwise.  dst row stride LDR, channel stride LDC floats.
// ------------------------------------------------------------------------------------------
template <int ROWS, int COLS, int LDR, int LDC>
__device__ __forceinline__ void stage_dy(float* __restrict__ dst, const float* __restrict__ yb, const float* __restrict__ dab,
                                         const float* __restrict__ aff, const float* __restrict__ coef, int oh0, int ow0,
                                         int hout, int tid) {
  constexpr int NV = COLS / 4, TAIL = COLS % 4;          // vectors per row, scalar tail columns
  if ((hout & 3) == 0) {
    constexpr int PER = NV + TAIL;
    for (int idx = tid; idx < kC * ROWS * PER; idx += 256) {
      const int co = idx / (ROWS * PER), rem = idx - co * (ROWS * PER);
      const int r = rem / PER, v = rem - r * PER;
      const int oh = oh0 + r;
      const float mean = aff[co], rstd = aff[kC + co], sc = aff[2 * kC + co], sh = aff[3 * kC + co];
      const float k0 = coef[co], k1 = coef[kC + co], k2 = coef[2 * kC + co];
      float* d = dst + co * LDC + r * LDR;
      if (v < NV) {
        const int ow = ow0 + 4 * v;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oh < hout && ow < hout) {
          const size_t o_ = ((size_t)co * hout + oh) * hout + ow;
          const float4 yv = *reinterpret_cast<const float4*>(yb + o_);
          const float4 da = *reinterpret_cast<const float4*>(dab + o_);
          o.x = k0 * ((fmaf(sc, yv.x, sh) > 0.f ? da.x : 0.f) - k1 - (yv.x - mean) * rstd * k2);
          o.y = k0 * ((fmaf(sc, yv.y, sh) > 0.f ? da.y : 0.f) - k1 - (yv.y - mean) * rstd * k2);
          o.z = k0 * ((fmaf(sc, yv.z, sh) > 0.f ? da.z : 0.f) - k1 - (yv.z - mean) * rstd * k2);
          o.w = k0 * ((fmaf(sc, yv.w, sh) > 0.f ? da.w : 0.f) - k1 - (yv.w - mean) * rstd * k2);
        }
        *reinterpret_cast<float4*>(d + 4 * v) = o;
      } else {
        const int c = 4 * NV + (v - NV), ow = ow0 + c;
        const size_t o_ = ((size_t)co * hout + oh) * hout + ow;
        d[c] = (oh < hout && ow < hout) ? bn_bwd_dy(yb[o_], dab[o_], aff, coef, co) : 0.f;
      }
    }
  } else {
    for (int idx = tid; idx < kC * ROWS * COLS; idx += 256) {
      const int co = idx / (ROWS * COLS), rem = idx % (ROWS * COLS);
      const int r = rem / COLS, c = rem % COLS;
      const int oh = oh0 + r, ow = ow0 + c;
      const size_t o_ = ((size_t)co * hout + oh) * hout + ow;
      dst[co * LDC + r * LDR + c] = (oh < hout && ow < hout) ? bn_bwd_dy(yb[o_], dab[o_], aff, coef, co) : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: per-block partial dW[24][CIN][9] over a 16x16 tile of dy.  grid (tiles, B), block 256.
// A thread owns 6 output channels x 9 taps of ONE input channel (54 accumulators): per pixel 9 input LDS + 6 dy
// LDS (broadcast) feed 54 FFMA.
//   CIN == 24: warp = pixel partition; lane = 4 output-channel groups x 8 input channels; three 8-channel chunks.
//   CIN == 3 : warp = TWO pixel partitions; lane = 2 x (4 groups x 3 channels) (24 of 32 lanes busy).
// Partitions are combined by a fixed-order tree (shuffle between the two sub-partitions, then shared memory).
// ------------------------------------------------------------------------------------------
constexpr int kDyLd = kTile * kTile + 4;          // channel stride of the staged dy: 6*260 mod 32 = 24 -> the 4 groups hit distinct banks
constexpr int kWgRedFloats = 4 * 32 * 54;
template <int CIN>
constexpr int wgrad_chunk() { return CIN < kChunk ? CIN : kChunk; }
template <int CIN>
static size_t wgrad_smem_bytes() {
  return ((size_t)wgrad_chunk<CIN>() * kPatch * kPW + (size_t)kC * kDyLd) * sizeof(float) < (size_t)kWgRedFloats * 4 + (size_t)kC * kDyLd * 4
             ? (size_t)kWgRedFloats * 4 + (size_t)kC * kDyLd * 4
             : ((size_t)wgrad_chunk<CIN>() * kPatch * kPW + (size_t)kC * kDyLd) * sizeof(float);
}

template <int CIN, bool U8 = false>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ in_aff, const float* __restrict__ yout,
                  const float* __restrict__ dA, const float* __restrict__ aff_out, const float* __restrict__ coef,
                  float* __restrict__ part, int hin, int hout, int tiles_x) {
  constexpr int CC = wgrad_chunk<CIN>();
  constexpr int PATCH_FLOATS = CC * kPatch * kPW > kWgRedFloats ? CC * kPatch * kPW : kWgRedFloats;
  constexpr int NSUB = CIN == 3 ? 2 : 1;             // pixel partitions per warp
  extern __shared__ __align__(16) float wg_smem[];
  float (*patch)[kPatch][kPW] = reinterpret_cast<float (*)[kPatch][kPW]>(wg_smem);
  float* red = wg_smem;                                          // aliases the patch between channel chunks
  float* dys = wg_smem + PATCH_FLOATS;                           // [24][kDyLd]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = CIN == 3 ? lane / 12 : 0;
  const int cg = CIN == 3 ? (lane % 12) / 3 : lane >> 3;
  const int ci = CIN == 3 ? lane % 3 : lane & 7;
  const bool lane_active = CIN != 3 || lane < 24;
  const int pp = warp * NSUB + (lane_active ? sub : 0);
  const int b = blockIdx.y;
  const int oh0 = (blockIdx.x / tiles_x) * kTile, ow0 = (blockIdx.x % tiles_x) * kTile;
  const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
  const float* inb = U8 ? in : in + (size_t)b * CIN * hin * hin;
  const unsigned char* inb8 = reinterpret_cast<const unsigned char*>(in) + (size_t)b * CIN * hin * hin;
  float* out = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (kC * CIN * 9);

  stage_dy<kTile, kTile, kTile, kDyLd>(dys, yout + (size_t)b * kC * hout * hout, dA + (size_t)b * kC * hout * hout, aff_out,
                                      coef, oh0, ow0, hout, tid);

  for (int ci0 = 0; ci0 < CIN; ci0 += CC) {
    __syncthreads();
    if (U8) stage_patch_u8<CC>(patch, inb8, ih0, iw0, hin, tid);
    else stage_patch<CC>(patch, inb, in_aff, ci0, ih0, iw0, hin, tid);
    __syncthreads();
    float2 acc2[3][9];           // packed (channel 2j, 2j+1) accumulators: FFMA2
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc2[j][t] = make_float2(0.f, 0.f);
    if (lane_active) {
#pragma unroll 2
      for (int p = pp; p < kTile * kTile; p += 8 * NSUB) {
        const int py = p / kTile, px = p % kTile;
        if (oh0 + py >= hout || ow0 + px >= hout) continue;      // dy is zero there
        float xv[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) xv[kh * 3 + kw] = patch[ci][2 * py + kh][2 * px + kw + kPO];
        float2 g2[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) g2[j] = make_float2(dys[(cg * 6 + 2 * j) * kDyLd + p], dys[(cg * 6 + 2 * j + 1) * kDyLd + p]);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float2 xx = make_float2(xv[t], xv[t]);
#pragma unroll
          for (int j = 0; j < 3; ++j) acc2[j][t] = __ffma2_rn(g2[j], xx, acc2[j][t]);
        }
      }
    }
    float acc[6][9];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        acc[2 * j][t] = acc2[j][t].x;
        acc[2 * j + 1][t] = acc2[j][t].y;
      }
    if (NSUB == 2) {           // fold sub-partition 1 (lanes 12..23) into sub-partition 0 (lanes 0..11)
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] += __shfl_down_sync(0xffffffffu, acc[j][t], 12);
    }
    // combine the 8 warps: 4..7 -> 0..3, 2..3 -> 0..1, 1 -> 0 (fixed order: deterministic)
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
      __syncthreads();               // previous readers of `red` (or of the patch it aliases) are done
      if (warp >= half && warp < 2 * half) {
        float* dst = red + ((size_t)(warp - half) * 32 + lane) * 54;
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
          for (int t = 0; t < 9; ++t) dst[j * 9 + t] = acc[j][t];
      }
      __syncthreads();
      if (warp < half) {
        const float* src = red + ((size_t)warp * 32 + lane) * 54;
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[j][t] += src[j * 9 + t];
      }
    }
    if (warp == 0 && (CIN == 3 ? lane < 12 : true)) {
#pragma unroll
      for (int j = 0; j < 6; ++j)
#pragma unroll
        for (int t = 0; t < 9; ++t) out[((size_t)(cg * 6 + j) * CIN + ci0 + ci) * 9 + t] = acc[j][t];
    }
  }
}

// ------------------------------------------------------------------------------------------
// data gradient (layers 2..4): dA_prev[b, ci, ih, iw] = sum_{co,kh,kw} dy[b, co, oh, ow] W[co, ci, kh, kw]
// with ih = 2*oh + kh - 1.  Block: 32x32 input pixels of one image, thread = 2x2 input quad x 8 ci.
// grid (tiles, B), block 256; the three 8-channel chunks of ci are looped inside the block
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_dgrad_kernel(const float* __restrict__ yout, const float* __restrict__ dAout, const float* __restrict__ aff_out,
                  const float* __restrict__ coef, const float* __restrict__ w, float* __restrict__ dA, int hin, int hout,
                  int tiles_x) {
  __shared__ __align__(16) float dys[kC][kTile + 1][kTile + 4];
  __shared__ __align__(16) float wsm[kC][9][kChunk];        // [co][tap][ci]: channel pairs adjacent (FFMA2)
  const int tid = threadIdx.x, tx = tid % kTile, ty = tid / kTile;
  const int b = blockIdx.y;
  const int q0y = (blockIdx.x / tiles_x) * kTile, q0x = (blockIdx.x % tiles_x) * kTile;   // quad == output coords
  const float* yb_o = yout + (size_t)b * kC * hout * hout;
  const float* dab = dAout + (size_t)b * kC * hout * hout;

  // dy of the window is staged ONCE and reused for the three 8-channel chunks of the input gradient
  stage_dy<kTile + 1, kTile + 1, kTile + 4, (kTile + 1) * (kTile + 4)>(&dys[0][0][0], yb_o, dab, aff_out, coef, q0y, q0x, hout, tid);
  const bool warp_active = 2 * (q0y + (tid / 32) * 2) < hin;      // warp = quad rows 2w, 2w+1: skip warps entirely below the image
  const int ih = 2 * (q0y + ty), iw = 2 * (q0x + tx);

  for (int ci0 = 0; ci0 < kC; ci0 += kChunk) {
    __syncthreads();                   // dys staged / previous chunk's weights no longer read
    for (int idx = tid; idx < kC * kChunk * 9; idx += 256) {
      const int co = idx / (kChunk * 9), rem = idx % (kChunk * 9);
      const int ci = rem / 9, t = rem % 9;
      wsm[co][t][ci] = w[((size_t)co * kC + ci0 + ci) * 9 + t];
    }
    __syncthreads();
    if (!warp_active) continue;

    // quad (i, j) covers input pixels (2i, 2j), (2i, 2j+1), (2i+1, 2j), (2i+1, 2j+1); accumulators packed over channel pairs
    float2 acc2[kChunk / 2][4];
#pragma unroll
    for (int c2 = 0; c2 < kChunk / 2; ++c2)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc2[c2][p] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int co = 0; co < kC; ++co) {
      const float d00 = dys[co][ty][tx], d01 = dys[co][ty][tx + 1], d10 = dys[co][ty + 1][tx], d11 = dys[co][ty + 1][tx + 1];
      const float2 e00 = make_float2(d00, d00), e01 = make_float2(d01, d01), e10 = make_float2(d10, d10), e11 = make_float2(d11, d11);
#pragma unroll
      for (int h = 0; h < 2; ++h) {          // channel pairs (0,1),(2,3) then (4,5),(6,7): one LDS.128 per tap
        float2 wt[9][2];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 v = *reinterpret_cast<const float4*>(&wsm[co][t][4 * h]);
          wt[t][0] = make_float2(v.x, v.y);
          wt[t][1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float2* a = acc2[2 * h + k];
          // taps: 0 (0,0) 1 (0,1) 2 (0,2) 3 (1,0) 4 (1,1) 5 (1,2) 6 (2,0) 7 (2,1) 8 (2,2)
          // (even, even): kh = 1, kw = 1
          a[0] = __ffma2_rn(e00, wt[4][k], a[0]);
          // (even, odd): kh = 1; kw = 0 -> ow = j + 1, kw = 2 -> ow = j
          a[1] = __ffma2_rn(e01, wt[3][k], __ffma2_rn(e00, wt[5][k], a[1]));
          // (odd, even): kw = 1; kh = 0 -> oh = i + 1, kh = 2 -> oh = i
          a[2] = __ffma2_rn(e10, wt[1][k], __ffma2_rn(e00, wt[7][k], a[2]));
          // (odd, odd): kh, kw in {0, 2}
          a[3] = __ffma2_rn(e11, wt[0][k], __ffma2_rn(e10, wt[2][k], __ffma2_rn(e01, wt[6][k], __ffma2_rn(e00, wt[8][k], a[3]))));
        }
      }
    }
    float acc[kChunk][4];
#pragma unroll
    for (int c2 = 0; c2 < kChunk / 2; ++c2)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        acc[2 * c2][p] = acc2[c2][p].x;
        acc[2 * c2 + 1][p] = acc2[c2][p].y;
      }
    if (ih < hin && iw < hin) {
      float* o = dA + ((size_t)b * kC + ci0) * hin * hin;
#pragma unroll
      for (int ci = 0; ci < kChunk; ++ci) {
        float* oc = o + (size_t)ci * hin * hin;
        *reinterpret_cast<float2*>(&oc[(size_t)ih * hin + iw]) = make_float2(acc[ci][0], acc[ci][1]);
        *reinterpret_cast<float2*>(&oc[(size_t)(ih + 1) * hin + iw]) = make_float2(acc[ci][2], acc[ci][3]);
      }
    }
  }
}

}  // namespace rn

#include "conv_tc.cuh"

namespace rn {

// ---- tensor-core path (conv_tc.cuh): side % 64 == 0 keeps every layer's rows float4-aligned -------------------------
static bool conv_tc_ok(const rn_conv_cfg& c) { return c.side % 64 == 0 && !(c.flags & RN_CONV_FLAG_SIMT); }

struct TcGrid {
  int tw, tiles_x, tiles, units, grid;
};
// fwd / dgrad: unit = TW x TW outputs (quads); wgrad: 8 x 16 outputs (TW = 16) or two 8x8 images
static TcGrid tc_grid(int B, int hout, int unit_rows16, int nimg8, int blocks_per_sm = 2) {
  TcGrid g;
  g.tw = hout <= 8 ? 8 : 16;
  if (g.tw == 16) {
    g.tiles_x = cdiv(hout, 16);
    g.tiles = g.tiles_x * cdiv(hout, unit_rows16);
    g.units = B * g.tiles;
  } else {
    g.tiles_x = 1;
    g.tiles = 1;
    g.units = cdiv(B, nimg8);
  }
  const int slots = blocks_per_sm * sm_count();
  const int per = cdiv(g.units, slots);
  g.grid = cdiv(g.units, per);
  return g;
}

template <int CIN, int TW, bool U8>
static int launch_fwd_tc(const TcGrid& g, const void* in, const float* in_aff, const rn_conv_layer& L, float* y, float* part,
                         int B, int hin, int hout, cudaStream_t st) {
  const size_t smem = ctc::FwdCfg<CIN, TW>::SMEM;
  RN_CUDA(cudaFuncSetAttribute(ctc::conv_fwd_tc_kernel<CIN, TW, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc::conv_fwd_tc_kernel<CIN, TW, U8><<<g.grid, 256, smem, st>>>(in, in_aff, L.w, L.bias, y, part, B, hin, hout, g.tiles_x,
                                                                 g.tiles, g.units);
  RN_LAUNCH_CHECK("conv_fwd_tc_kernel");
  return RN_OK;
}

constexpr int kRedChunks = 64;      // row chunks of the two-stage fixed-order reduction of the wgrad partials

struct ConvPlan {
  int h[RN_CONV_LAYERS + 1];        // h[0] = side, h[l+1] = output edge of layer l
  size_t y_off[RN_CONV_LAYERS];     // float offsets of y_l in `saved`
  size_t aff_off[RN_CONV_LAYERS];
  size_t saved_floats;
  int tiles[RN_CONV_LAYERS];        // 16x16 output tiles per image
  size_t scratch_floats;
  size_t spart_floats, wpart_floats;   // BatchNorm / weight-gradient partial regions of `scratch`
};

static ConvPlan make_plan(const rn_conv_cfg& c) {
  ConvPlan p;
  p.h[0] = c.side;
  size_t off = 0;
  for (int l = 0; l < RN_CONV_LAYERS; ++l) {
    p.h[l + 1] = p.h[l] / 2;
    p.y_off[l] = off;
    off += round_up((size_t)c.B * kC * p.h[l + 1] * p.h[l + 1], 64);
    const int t = cdiv(p.h[l + 1], kTile);
    p.tiles[l] = t * t;
  }
  for (int l = 0; l < RN_CONV_LAYERS; ++l) {
    p.aff_off[l] = off;
    off += 4 * kC;
  }
  p.saved_floats = off;
  // scratch: dA + dy (largest layer), BN partials, coef, wgrad partials
  const size_t big = round_up((size_t)c.B * kC * p.h[1] * p.h[1], 64);
  size_t wpart = 0, spart = 0;
  for (int l = 0; l < RN_CONV_LAYERS; ++l) {
    const size_t cin = l == 0 ? 3 : kC;
    // per-block weight-gradient / BatchNorm partials: one per 16x16 tile (SIMT kernels) or one per persistent block of the
    // tensor-core kernels (at most two blocks per SM, never more than their 8-row units)
    const size_t tc_blocks = std::min<size_t>((size_t)c.B * cdiv(p.h[l + 1], 16) * cdiv(p.h[l + 1], 8), (size_t)2 * sm_count());
    const size_t blocks = std::max<size_t>((size_t)c.B * p.tiles[l], tc_blocks);
    wpart = std::max(wpart, blocks * kC * cin * 9);
    spart = std::max(spart, blocks * 2 * kC);
  }
  p.spart_floats = round_up(std::max(spart, (size_t)c.B * kC * 2), 64);
  p.wpart_floats = round_up(wpart, 64);
  p.scratch_floats = 2 * big + p.spart_floats + 64 * 4 + p.wpart_floats + (size_t)kRedChunks * kC * kC * 9;
  return p;
}

}  // namespace rn

using namespace rn;

static int validate_conv(const rn_conv_cfg* c) {
  RN_CHECK_ARG(c != nullptr, "cfg is NULL");
  RN_CHECK_ARG(c->B > 0 && c->B <= 65535, "B must be in [1, 65535] (B=%d)", c->B);
  RN_CHECK_ARG(c->side >= 16 && c->side % 16 == 0, "side must be a positive multiple of 16 (side=%d)", c->side);
  return RN_OK;
}

extern "C" int rn_conv_workspace(const rn_conv_cfg* cfg, size_t* saved_floats, size_t* scratch_floats) {
  RN_TRY(validate_conv(cfg));
  RN_CHECK_ARG(saved_floats && scratch_floats, "output pointers are NULL");
  ConvPlan p = make_plan(*cfg);
  *saved_floats = p.saved_floats;
  *scratch_floats = p.scratch_floats;
  return RN_OK;
}

extern "C" int rn_conv_fwd(const rn_conv_cfg* cfg, const void* img_any, const rn_conv_layer* L, float* objects,
                           float* saved, float* scratch, void* stream) {
  const float* img = static_cast<const float*>(img_any);
  RN_TRY(validate_conv(cfg));
  RN_CHECK_ARG(img && L && objects && saved && scratch, "NULL pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvPlan p = make_plan(*cfg);
  const float* in = img;
  const float* in_aff = nullptr;
  for (int l = 0; l < RN_CONV_LAYERS; ++l) {
    RN_CHECK_ARG(L[l].w && L[l].bias && L[l].gamma && L[l].beta && L[l].running_mean && L[l].running_var,
                 "conv layer %d has a NULL parameter", l);
    float* y = saved + p.y_off[l];
    float* aff = saved + p.aff_off[l];
    const int hin = p.h[l], hout = p.h[l + 1], tx = cdiv(hout, kTile);
    dim3 grid(p.tiles[l], cfg->B);
    float* part = cfg->training ? scratch : nullptr;
    int nblk = p.tiles[l] * cfg->B;
    if (conv_tc_ok(*cfg) && (cfg->flags & RN_CONV_FLAG_TC_FWD)) {
      const TcGrid g = tc_grid(cfg->B, hout, 16, 4);
      nblk = g.grid;
      if (l == 0 && cfg->img_u8) RN_TRY((launch_fwd_tc<3, 16, true>(g, in, in_aff, L[l], y, part, cfg->B, hin, hout, st)));
      else if (l == 0) RN_TRY((launch_fwd_tc<3, 16, false>(g, in, in_aff, L[l], y, part, cfg->B, hin, hout, st)));
      else if (g.tw == 16) RN_TRY((launch_fwd_tc<kC, 16, false>(g, in, in_aff, L[l], y, part, cfg->B, hin, hout, st)));
      else RN_TRY((launch_fwd_tc<kC, 8, false>(g, in, in_aff, L[l], y, part, cfg->B, hin, hout, st)));
    } else if (l == 0 && cfg->img_u8)
      conv_fwd_kernel<3, true><<<grid, 256, 0, st>>>(in, in_aff, L[l].w, L[l].bias, y, part, hin, hout, tx);
    else if (l == 0)
      conv_fwd_kernel<3><<<grid, 256, 0, st>>>(in, in_aff, L[l].w, L[l].bias, y, part, hin, hout, tx);
    else
      conv_fwd_kernel<kC><<<grid, 256, 0, st>>>(in, in_aff, L[l].w, L[l].bias, y, part, hin, hout, tx);
    RN_LAUNCH_CHECK("conv_fwd_kernel");
    bn_finalize_kernel<<<kC, 256, 0, st>>>(scratch, nblk, (double)cfg->B * hout * hout, L[l].gamma,
                                             L[l].beta, L[l].running_mean, L[l].running_var, aff, cfg->eps,
                                             cfg->momentum, cfg->training);
    RN_LAUNCH_CHECK("bn_finalize_kernel");
    in = y;
    in_aff = aff;
  }
  const int d = p.h[RN_CONV_LAYERS];
  const long long total = (long long)cfg->B * d * d * 26;
  objects_fwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(in, in_aff, objects, d, total);
  RN_LAUNCH_CHECK("objects_fwd_kernel");
  return RN_OK;
}

extern "C" int rn_conv_bwd(const rn_conv_cfg* cfg, const void* img_any, const float* dobjects, const rn_conv_layer* L,
                           const float* saved, const rn_conv_grads* Gr, float* scratch, void* stream) {
  const float* img = static_cast<const float*>(img_any);
  RN_TRY(validate_conv(cfg));
  RN_CHECK_ARG(img && dobjects && L && saved && Gr && scratch, "NULL pointer argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvPlan p = make_plan(*cfg);
  const size_t big = round_up((size_t)cfg->B * kC * p.h[1] * p.h[1], 64);
  float* dA = scratch;
  float* dA_next = scratch + big;      // BatchNorm backward is applied on the fly by the consumers: no dy buffer
  float* bnpart = dA_next + big;
  float* coef = bnpart + p.spart_floats;
  float* wpart = coef + 64 * 4;
  const size_t wpart_floats = p.wpart_floats;

  const int d = p.h[RN_CONV_LAYERS];
  {
    const long long total = (long long)cfg->B * kC * d * d;
    objects_bwd_kernel<<<cdiv(total, 256), 256, 0, st>>>(dobjects, dA, d, total);
    RN_LAUNCH_CHECK("objects_bwd_kernel");
  }
  int bn_blocks = 0;      // > 0: `bnpart` already holds that many per-block partials of the layer about to be processed
  for (int l = RN_CONV_LAYERS - 1; l >= 0; --l) {
    RN_CHECK_ARG(Gr[l].dw && Gr[l].dbias && Gr[l].dgamma && Gr[l].dbeta, "conv grads of layer %d have a NULL pointer", l);
    const int hin = p.h[l], hout = p.h[l + 1], hw = hout * hout, tx = cdiv(hout, kTile);
    const float* y = saved + p.y_off[l];
    const float* aff = saved + p.aff_off[l];
    // BatchNorm-backward sums: a pass over (y, dA) for the last layer and the SIMT path; the tensor-core data gradient
    // of layer l + 1 has already left them as per-block partials (bn_blocks of them) in `bnpart`
    if (bn_blocks == 0) {
      bn_bwd_reduce_kernel<<<dim3(kC, cfg->B), 256, 0, st>>>(y, dA, aff, bnpart, hw);
      RN_LAUNCH_CHECK("bn_bwd_reduce_kernel");
      bn_blocks = cfg->B;
    }
    bn_bwd_finalize_kernel<<<1, 32 * kC, 0, st>>>(bnpart, bn_blocks, (double)cfg->B * hw, L[l].gamma, aff, coef, Gr[l].dgamma,
                                                 Gr[l].dbeta, Gr[l].dbias, cfg->training);
    RN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    bn_blocks = 0;
    // weight gradient: per-block partials, then a fixed-order sum over blocks
    const float* in = l == 0 ? img : saved + p.y_off[l - 1];
    const float* in_aff = l == 0 ? nullptr : saved + p.aff_off[l - 1];
    dim3 grid(p.tiles[l], cfg->B);
    const int cin = l == 0 ? 3 : kC;
    const bool tc = conv_tc_ok(*cfg);
    int nblk = p.tiles[l] * cfg->B;
    if (tc && l == 0) {
      // single-stage form, two blocks per SM: the per-unit MMA work of the RGB layer is too small for the producer /
      // consumer ring to pay (measured 0.20 ms against 0.28 ms warp-specialised at batch 640)
      const TcGrid g = tc_grid(cfg->B, hout, 16, 1, 2);
      nblk = g.grid;
      const size_t smem = ctc::Wg3Cfg::smem(false);
      if (cfg->img_u8) {
        RN_CUDA(cudaFuncSetAttribute(ctc::conv_wgrad3_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctc::conv_wgrad3_tc_kernel<true, false><<<g.grid, 256, smem, st>>>(in, y, dA, aff, coef, wpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
      } else {
        RN_CUDA(cudaFuncSetAttribute(ctc::conv_wgrad3_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctc::conv_wgrad3_tc_kernel<false, false><<<g.grid, 256, smem, st>>>(in, y, dA, aff, coef, wpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
      }
    } else if (tc) {
      // 32x32 / 16x16 outputs: warp-specialised ring, one block per SM (measured faster than two single-stage blocks per
      // SM at every batch from 80 to 640); 8x8 outputs: single-stage form
      if (hout > 8) {
        const TcGrid g = tc_grid(cfg->B, hout, 8, 2, 1);
        nblk = g.grid;
        const size_t smem = ctc::WgCfg<16>::smem(true);
        RN_CUDA(cudaFuncSetAttribute(ctc::conv_wgrad_tc_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctc::conv_wgrad_tc_kernel<16, true><<<g.grid, 448, smem, st>>>(in, in_aff, y, dA, aff, coef, wpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
      } else {
        const TcGrid g = tc_grid(cfg->B, hout, 8, 2, 2);
        nblk = g.grid;
        const size_t smem = ctc::WgCfg<8>::smem(false);
        RN_CUDA(cudaFuncSetAttribute(ctc::conv_wgrad_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctc::conv_wgrad_tc_kernel<8, false><<<g.grid, 224, smem, st>>>(in, in_aff, y, dA, aff, coef, wpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
      }
    } else if (l == 0 && cfg->img_u8) {
      const size_t smem = wgrad_smem_bytes<3>();
      RN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv_wgrad_kernel<3, true><<<grid, 256, smem, st>>>(in, in_aff, y, dA, aff, coef, wpart, hin, hout, tx);
    } else if (l == 0) {
      const size_t smem = wgrad_smem_bytes<3>();
      RN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv_wgrad_kernel<3><<<grid, 256, smem, st>>>(in, in_aff, y, dA, aff, coef, wpart, hin, hout, tx);
    } else {
      const size_t smem = wgrad_smem_bytes<kC>();
      RN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<kC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv_wgrad_kernel<kC><<<grid, 256, smem, st>>>(in, in_aff, y, dA, aff, coef, wpart, hin, hout, tx);
    }
    RN_LAUNCH_CHECK("conv_wgrad_kernel");
    {
      // dW = sum over blocks of the partials, two stages so the reduction itself fills the machine:
      // [nblk][N] -> [64][N] -> [N]  (both fixed order: deterministic)
      const int N = kC * cin * 9;
      float* red_tmp = wpart + round_up((size_t)wpart_floats, 64);
      if (nblk >= 4 * kRedChunks) {
        const int per = cdiv(nblk, kRedChunks);
        const int chunks = cdiv(nblk, per);
        // rows [c*per, min(nblk, (c+1)*per)): the last chunk may be short -> handle it with a second call
        const int full = nblk / per;
        RN_TRY(colsum(wpart, red_tmp, N, full, 1, per, 0, 1, per, st));
        if (chunks > full) RN_TRY(colsum(wpart + (size_t)full * per * N, red_tmp + (size_t)full * N, N, 1, 1, 0, 0, 1, nblk - full * per, st));
        RN_TRY(colsum(red_tmp, Gr[l].dw, N, 1, 1, 0, 0, 1, chunks, st));
      } else {
        RN_TRY(colsum(wpart, Gr[l].dw, N, 1, 1, 0, 0, 1, nblk, st));
      }
    }
    if (l > 0) {
      // data gradient into dA (now sized for layer l-1's output == this layer's input)
      const int qt = cdiv(hout, kTile);
      // reads (y_l, dA_l), writes dA_{l-1} into the other buffer
      if (tc) {
        const bool ws = hout > 8;
        const TcGrid g = tc_grid(cfg->B, hout, 16, 4, ws ? 1 : 2);
        const float* y_in = saved + p.y_off[l - 1];
        const float* aff_in = saved + p.aff_off[l - 1];
        if (ws) {
          const size_t smem = ctc::DgCfg<16>::smem(true);
          RN_CUDA(cudaFuncSetAttribute(ctc::conv_dgrad_tc_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          ctc::conv_dgrad_tc_kernel<16, true><<<g.grid, 512, smem, st>>>(y, dA, aff, coef, L[l].w, dA_next, y_in, aff_in, bnpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
        } else {
          const size_t smem = ctc::DgCfg<8>::smem(false);
          RN_CUDA(cudaFuncSetAttribute(ctc::conv_dgrad_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          ctc::conv_dgrad_tc_kernel<8, false><<<g.grid, 256, smem, st>>>(y, dA, aff, coef, L[l].w, dA_next, y_in, aff_in, bnpart, cfg->B, hin, hout, g.tiles_x, g.tiles, g.units);
        }
        bn_blocks = g.grid;
      } else {
        conv_dgrad_kernel<<<dim3(qt * qt, cfg->B), 256, 0, st>>>(y, dA, aff, coef, L[l].w, dA_next, hin, hout, qt);
      }
      RN_LAUNCH_CHECK("conv_dgrad_kernel");
      float* t = dA; dA = dA_next; dA_next = t;
    }
  }
  return RN_OK;
}
