// conv_tc.cuh -- tensor-core kernels of the conv stack (forward, weight gradient, data gradient of the 3x3 stride-2
// convolutions, reference model.py:22-36), included by conv.cu.
//
// The stack is 1 % of the step's FLOPs but ran at ~30 % of the fp32 SIMT peak (instruction-issue bound), 25 % of the
// step.  Here every convolution is an implicit GEMM on the warp-level tensor path (mma.sync m16n8k8, TF32 operands,
// fp32 accumulation) with the error-compensated 3-pass split  x*w = x_hi*w_hi + x_lo*w_hi + x_hi*w_lo  (x_hi = tf32(x),
// x_lo = x - x_hi), which carries fp32-level accuracy (2^-21 relative per product); tests hold the same 2e-4 / 5e-4
// bars as the fp32 SIMT kernels.  Why mma.sync and not tcgen05: the operands are GATHERED (stride-2 taps of an NCHW patch
// with the previous layer's BatchNorm affine + ReLU, or BatchNorm-backward, applied on the fly), N is 24, and the stack is
// HBM-bound once off the FFMA pipe -- register fragments need no im2col image in shared memory and no TMEM round trip.
// Measured rate (tests/micro/mma_sync_rate.cu): 8.6 cycles per m16n8k8 per SM sub-core = 276 TFLOP/s TF32.
//
// Shared-memory patch layout: per (channel, image) plane NR rows of RS floats; a row holds the EVEN patch columns
// E[e] (patch column 2e) at [1 + e] and the ODD ones O[o] (column 2o + 1) at [OB + o], so the stride-2 tap (kx) of 8
// consecutive output pixels is 8 consecutive floats (kx = 0: E[ox], 1: O[ox], 2: E[ox + 1]) and, with the plane stride
// congruent to 8 (pixels on the fragment's row axis) or 4 (pixels on its K axis) modulo 32, every fragment load is
// bank-conflict free.  Patch column j is input column 2*ow0 - 1 + j; rows are staged as aligned float4 (input columns
// 2*ow0 - 4 + 4v ..), so one vector = (O[2v-2], E[2v-1], O[2v-1], E[2v]) = two float2 stores.
#pragma once

namespace rn {
namespace ctc {

// x = hi + lo, hi = x rounded to TF32 (10 explicit mantissa bits).  cvt.rna.tf32.f32 has no SASS instruction on sm_100a: it
// expands to add / inf-test / select / mask (measured 12.0 vs 10.1 cycles per MMA in the 3-pass loop,
// tests/micro/mma_sync_rate2.cu), so the rounding is done by hand -- round half away from zero on the magnitude, two
// integer operations; activations and weights are finite and far from FLT_MAX, so no overflow case exists.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));      // exact; the MMA ignores the low 13 bits of lo (2^-21 relative)
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// c += A * B with both operands split: small terms first
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(c, al[0], al[1], al[2], al[3], bh0, bh1);
  mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

// the same for the three n-tiles of one A fragment, PASS-major: consecutive MMAs are independent (a dependent
// back-to-back chain on one accumulator leaves the tensor pipe idle for the MMA latency)
__device__ __forceinline__ void mma3x3(float (&c)[3][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       const uint32_t (&bh)[3][2], const uint32_t (&bl)[3][2]) {
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) mma_tf32(c[nt], al[0], al[1], al[2], al[3], bh[nt][0], bh[nt][1]);
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) mma_tf32(c[nt], ah[0], ah[1], ah[2], ah[3], bl[nt][0], bl[nt][1]);
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) mma_tf32(c[nt], ah[0], ah[1], ah[2], ah[3], bh[nt][0], bh[nt][1]);
}

constexpr int pad32(int x, int r) { return x + ((r - x % 32) + 32) % 32; }   // smallest y >= x with y % 32 == r

template <int TW>
struct Geo {                                  // TW = tile width in output pixels (16, or 8 for the 8x8 layer)
  static constexpr int NV = TW == 16 ? 9 : 5;     // aligned float4 per patch row
  static constexpr int RS = 4 * NV;               // 36 / 20 floats per patch row
  static constexpr int OB = TW == 16 ? 20 : 12;   // base of the odd columns (even, >= 2 + TW + 2: two dummy slots before it)
  __device__ static constexpr int kxoff(int kx) { return kx == 0 ? 1 : (kx == 1 ? OB : 2); }
};

// Register-staged patch loaders: NIMG images x NC channels x NR rows.  issue() puts every global load of the thread in
// flight, commit() applies the producer's BatchNorm affine + ReLU (the zero padding is applied AFTER the activation) and
// writes the de-interleaved rows.  NT = threads of the block.  Index arithmetic is what these kernels pay for (a flat-
// index decode per vector cost 3x the MMAs' instructions in the first version), hence two thread mappings:
//   RowStager  (TW = 16, one image): thread -> (channel plane, row slot, vector column), then a walk down the rows: one
//              address add and one row-bound compare per vector;
//   WalkStager (TW = 8, several images): thread -> vector column, strided walk over the (plane, row) pairs.
template <int NC, int NR, int PS, int NT, bool U8>
struct RowStager {
  static constexpr int NV = Geo<16>::NV, RS = Geo<16>::RS, OB = Geo<16>::OB;
  static constexpr int KS0 = NT / (NC * NV);                       // row slots per plane
  static constexpr int KS = KS0 > NR ? NR : KS0;
  static constexpr int PER = (NR + KS - 1) / KS;                   // rows per slot
  float4 v[PER];
  uint32_t ok;

  __device__ __forceinline__ void issue(const void* __restrict__ inb, int /*nimg_valid*/, int /*img_stride*/, int hin2, int ih0,
                                        int iw0, int hin, int tid) {
    const int q = tid / NV, vv = tid - q * NV;
    const int pl = q / KS, k = q - pl * KS;
    const int iw = iw0 - 3 + 4 * vv;                    // aligned: a vector is entirely in or out
    const bool colok = pl < NC && (unsigned)iw < (unsigned)hin;
    const int ih = ih0 + k * PER;
    const int off0 = pl * hin2 + ih * hin + iw;
    ok = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const bool valid = colok && (KS * PER == NR || k * PER + i < NR) && (unsigned)(ih + i) < (unsigned)hin;
      const int off = valid ? off0 + i * hin : 0;       // branch-free: an invalid vector reads element 0 and is masked
      float4 t;
      if (U8) t = make_float4(__uint_as_float(*reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(inb) + off)), 0.f, 0.f, 0.f);
      else t = *reinterpret_cast<const float4*>(static_cast<const float*>(inb) + off);
      v[i] = t;
      ok |= (valid ? 1u : 0u) << i;
    }
  }

  __device__ __forceinline__ void commit(float* __restrict__ dst, const float* __restrict__ aff_s, int tid) const {
    const int q = tid / NV, vv = tid - q * NV;
    const int pl = q / KS, k = q - pl * KS;
    if (pl >= NC) return;
    float sc = 1.f, sh = 0.f;
    if (aff_s != nullptr) { sc = aff_s[pl]; sh = aff_s[kC + pl]; }
    float* d = dst + pl * PS + (k * PER) * RS + 2 * vv;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (KS * PER != NR && k * PER + i >= NR) break;
      float4 t = v[i];
      const bool valid = (ok >> i) & 1u;
      if (U8) {                               // x = u / 255 with a true division: bit-identical to ToTensor
        const uint32_t u = __float_as_uint(t.x);
        t = make_float4(__fdiv_rn((float)(u & 0xff), 255.f), __fdiv_rn((float)((u >> 8) & 0xff), 255.f),
                        __fdiv_rn((float)((u >> 16) & 0xff), 255.f), __fdiv_rn((float)(u >> 24), 255.f));
      }
      if (aff_s != nullptr) {
        t.x = fmaxf(fmaf(sc, t.x, sh), 0.f);
        t.y = fmaxf(fmaf(sc, t.y, sh), 0.f);
        t.z = fmaxf(fmaf(sc, t.z, sh), 0.f);
        t.w = fmaxf(fmaf(sc, t.w, sh), 0.f);
      }
      if (!valid) t = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float2*>(d + i * RS) = make_float2(t.y, t.w);               // E[2v-1], E[2v]
      *reinterpret_cast<float2*>(d + i * RS + OB - 2) = make_float2(t.x, t.z);      // O[2v-2], O[2v-1]
    }
  }
};

template <int TW, int NIMG, int NC, int NR, int PS, int IS, int NT, bool U8>
struct WalkStager {
  static constexpr int NV = Geo<TW>::NV, RS = Geo<TW>::RS, OB = Geo<TW>::OB;
  static constexpr int TR = NT / NV;                    // thread rows
  static constexpr int ROWS = NIMG * NC * NR;           // (plane, row) pairs
  static constexpr int PER = (ROWS + TR - 1) / TR;
  static constexpr int DR = TR % NR, DP = TR / NR;      // (row, plane) advance per step
  float4 v[PER];
  uint32_t ok;

  // inb: channel c0 of the unit's first image (elements; bytes when U8); img_stride = cin * hin * hin
  __device__ __forceinline__ void issue(const void* __restrict__ inb, int nimg_valid, int img_stride, int hin2, int ih0, int iw0,
                                        int hin, int tid) {
    const int q0 = tid / NV, vv = tid - q0 * NV;
    const int iw = iw0 - 3 + 4 * vv;                    // aligned: a vector is entirely in or out
    const bool colok = q0 < TR && (unsigned)iw < (unsigned)hin;
    int pl = q0 / NR, r = q0 - pl * NR;
    ok = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int img = NIMG == 1 ? 0 : pl / NC, c = pl - img * NC;
      const int ih = ih0 + r;
      const bool valid = colok && pl < NIMG * NC && img < nimg_valid && (unsigned)ih < (unsigned)hin;
      const int off = valid ? img * img_stride + c * hin2 + ih * hin + iw : 0;
      if (U8) v[i] = make_float4(__uint_as_float(*reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(inb) + off)), 0.f, 0.f, 0.f);
      else v[i] = *reinterpret_cast<const float4*>(static_cast<const float*>(inb) + off);
      ok |= (valid ? 1u : 0u) << i;
      r += DR;
      pl += DP;
      if (r >= NR) { r -= NR; ++pl; }
    }
  }

  // dst: plane of channel c0; aff_s: shared-memory (scale[24], shift[24]) of the producing layer offset by c0, or nullptr
  __device__ __forceinline__ void commit(float* __restrict__ dst, const float* __restrict__ aff_s, int tid) const {
    const int q0 = tid / NV, vv = tid - q0 * NV;
    if (q0 >= TR) return;
    int pl = q0 / NR, r = q0 - pl * NR;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (pl >= NIMG * NC) break;
      const int img = NIMG == 1 ? 0 : pl / NC, c = pl - img * NC;
      float4 t = v[i];
      const bool valid = (ok >> i) & 1u;
      if (U8) {
        const uint32_t u = __float_as_uint(t.x);
        t = make_float4(__fdiv_rn((float)(u & 0xff), 255.f), __fdiv_rn((float)((u >> 8) & 0xff), 255.f),
                        __fdiv_rn((float)((u >> 16) & 0xff), 255.f), __fdiv_rn((float)(u >> 24), 255.f));
      }
      if (aff_s != nullptr) {
        const float sc = aff_s[c], sh = aff_s[kC + c];
        t.x = fmaxf(fmaf(sc, t.x, sh), 0.f);
        t.y = fmaxf(fmaf(sc, t.y, sh), 0.f);
        t.z = fmaxf(fmaf(sc, t.z, sh), 0.f);
        t.w = fmaxf(fmaf(sc, t.w, sh), 0.f);
      }
      if (!valid) t = make_float4(0.f, 0.f, 0.f, 0.f);
      float* d = dst + c * PS + img * IS + r * RS + 2 * vv;
      *reinterpret_cast<float2*>(d) = make_float2(t.y, t.w);               // E[2v-1], E[2v]
      *reinterpret_cast<float2*>(d + OB - 2) = make_float2(t.x, t.z);      // O[2v-2], O[2v-1]
      r += DR;
      pl += DP;
      if (r >= NR) { r -= NR; ++pl; }
    }
  }
};

template <int TW, int NIMG, int NC, int NR, int PS, int IS, int NT, bool U8>
struct PatchStagerSel { using type = WalkStager<TW, NIMG, NC, NR, PS, IS, NT, U8>; };
template <int NC, int NR, int PS, int IS, int NT, bool U8>
struct PatchStagerSel<16, 1, NC, NR, PS, IS, NT, U8> { using type = RowStager<NC, NR, PS, NT, U8>; };
template <int TW, int NIMG, int NC, int NR, int PS, int IS, int NT, bool U8>
using PatchStager = typename PatchStagerSel<TW, NIMG, NC, NR, PS, IS, NT, U8>::type;

// ------------------------------------------------------------------------------------------
// forward: y = conv(act(in)) + bias, per-block (sum, sum of squares) partials for BatchNorm.
// Implicit GEMM  D[pixel][co] = sum_{tap, ci} A[pixel][(tap, ci)] * W[(tap, ci)][co]:  M = 16 output pixels (two
// 8-pixel row segments), N = 24 = 3 n-tiles, K = 8 input channels of one tap per MMA.
// Persistent blocks of 256 threads loop over units (TW = 16: one 16x16 output tile of one image; TW = 8: the 8x8
// outputs of four images); warp w owns two m-tiles.  24-channel layers run three 8-channel chunks per unit through a
// register-prefetched double buffer: the loads of chunk k+1 are in flight while chunk k is multiplied, one barrier per
// chunk.  The RGB layer (K = 27 padded to 32) keeps its weight fragments in registers.
// ------------------------------------------------------------------------------------------
template <int CIN, int TW>
struct FwdCfg {
  static constexpr int NIMG = TW == 16 ? 1 : 4;
  static constexpr int NR = 2 * TW + 1;
  static constexpr int NC = CIN == kC ? 8 : CIN;
  static constexpr int NCHUNK = CIN / NC;
  static constexpr int IS = NR * Geo<TW>::RS;
  static constexpr int PS = pad32(NIMG * IS, 8);
  static constexpr int BUF = NC * PS;                               // floats per patch buffer
  static constexpr int WFLOATS = CIN == kC ? kC * kC * 9 : 0;       // weight fragments (24-channel layers)
  static constexpr size_t SMEM = (size_t)(2 * BUF + WFLOATS + 2 * kC + 8 * 2 * kC) * sizeof(float);
};

struct UnitPos {
  int b0, r0, c0;        // first image, first output row / column (quad row / column for the data gradient)
};
// Walk over the block's units u = blockIdx.x + k * gridDim.x, u = image group * tiles + tile, without divisions in the
// loop: (group, tile) advance by the precomputed (gridDim.x / tiles, gridDim.x % tiles); tile -> (row, column) through a
// 16-bit reciprocal (exact: tile * tiles_x < 2^16).
template <int NIMG, int TH, int TWID>
struct UnitWalk {
  int bg, tile, dbg, dtile, tiles, tiles_x, inv;
  __device__ __forceinline__ UnitWalk(int tiles_, int tiles_x_) : tiles(tiles_), tiles_x(tiles_x_) {
    bg = blockIdx.x / tiles;
    tile = blockIdx.x - bg * tiles;
    dbg = gridDim.x / tiles;
    dtile = gridDim.x - dbg * tiles;
    inv = (65536 + tiles_x - 1) / tiles_x;
  }
  __device__ __forceinline__ UnitPos pos() const {
    UnitPos p;
    const int ty = (tile * inv) >> 16;
    p.b0 = bg * NIMG;
    p.r0 = ty * TH;
    p.c0 = (tile - ty * tiles_x) * TWID;
    return p;
  }
  __device__ __forceinline__ void next() {
    bg += dbg;
    tile += dtile;
    if (tile >= tiles) { tile -= tiles; ++bg; }
  }
};

template <int CIN, int TW, bool U8>
__global__ void __launch_bounds__(256, 2)
conv_fwd_tc_kernel(const void* __restrict__ in, const float* __restrict__ in_aff, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ stat_part, int B, int hin,
                   int hout, int tiles_x, int tiles, int units) {
  using F = FwdCfg<CIN, TW>;
  using G = Geo<TW>;
  constexpr int PS = F::PS, IS = F::IS, RS = G::RS, NCHUNK = F::NCHUNK, NC = F::NC;
  extern __shared__ __align__(16) float fwd_smem[];
  float* patch = fwd_smem;                                  // [2][BUF]
  float2* wsm = reinterpret_cast<float2*>(fwd_smem + 2 * F::BUF);   // [chunk][tap][nt][lane] (b0, b1)
  float* affs = fwd_smem + 2 * F::BUF + F::WFLOATS;         // scale[24], shift[24] of the producing layer
  float* red = affs + 2 * kC;                               // [8][48]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hin2 = hin * hin, hw = hout * hout;

  if (in_aff != nullptr && tid < 2 * kC) affs[tid] = in_aff[2 * kC + tid];
  const float* aff_s = in_aff != nullptr ? affs : nullptr;

  // weight fragments: B[k = ci][n = co] of (chunk, tap, n-tile): b0 = W[8nt + g][8ch + t], b1 = W[8nt + g][8ch + t + 4]
  uint32_t wr_h[CIN == 3 ? 4 : 1][3][2], wr_l[CIN == 3 ? 4 : 1][3][2];   // RGB layer: fragments of k = tap*3 + ci in registers
  int koff[CIN == 3 ? 4 : 1][2];
  if (CIN == kC) {
    for (int idx = tid; idx < kC * kC * 9; idx += 256) {
      const int e = idx & 1, ln = (idx >> 1) & 31, q = idx >> 6;       // q = (chunk*9 + tap)*3 + nt
      const int nt = q % 3, tap = (q / 3) % 9, ch = q / 27;
      const int co = 8 * nt + (ln >> 2), ci = 8 * ch + (ln & 3) + 4 * e;
      reinterpret_cast<float*>(wsm)[idx] = w[((size_t)co * kC + ci) * 9 + tap];
    }
  } else {
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = 8 * s + t + 4 * e;
        const int tap = k / 3, ci = k - 3 * tap;
        koff[s][e] = k < 27 ? ci * PS + (tap / 3) * RS + G::kxoff(tap % 3) : 0;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const float wv = k < 27 ? w[((size_t)(8 * nt + g) * 3 + ci) * 9 + tap] : 0.f;
          split_tf32(wv, wr_h[s][nt][e], wr_l[s][nt][e]);
        }
      }
  }

  // this thread's pixels: m-tile j (0, 1), segment s (fragment rows g / g + 8)
  int so[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (TW == 16) so[j][s] = (2 * (2 * warp + j)) * RS + 8 * s + g;
      else so[j][s] = (warp >> 1) * IS + (2 * ((warp & 1) * 4 + 2 * j + s)) * RS + g;
    }

  float acc[2][3][4];
  float s1[3][2], s2[3][2];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) acc[j][nt][0] = acc[j][nt][1] = acc[j][nt][2] = acc[j][nt][3] = 0.f;

  PatchStager<TW, F::NIMG, NC, F::NR, PS, IS, 256, U8> stg;
  auto issue = [&](const UnitPos& p, int ch) {
    const size_t base = ((size_t)p.b0 * CIN + ch * NC) * hin2;
    const void* inb = U8 ? static_cast<const void*>(static_cast<const unsigned char*>(in) + base)
                         : static_cast<const void*>(static_cast<const float*>(in) + base);
    stg.issue(inb, B - p.b0, CIN * hin2, hin2, 2 * p.r0 - 1, 2 * p.c0 - 1, hin, tid);
  };

  __syncthreads();                       // affs (and the weight fragments) are visible
  int u = blockIdx.x, ch = 0, buf = 0;
  UnitWalk<F::NIMG, TW, TW> walk(tiles, tiles_x);
  UnitPos cur = walk.pos();
  if (u < units) {
    issue(cur, 0);
    stg.commit(patch, aff_s, tid);
  }
  __syncthreads();

  while (u < units) {
    const float* pb = patch + buf * F::BUF;
    // the next item: the next chunk of this unit, or the first chunk of the block's next unit
    int nu = u, nch = ch + 1;
    UnitPos nxt = cur;
    if (nch == NCHUNK) {
      nch = 0;
      nu = u + gridDim.x;
      walk.next();
      nxt = walk.pos();
    }
    const bool has_next = nu < units;
    if (has_next) issue(nxt, nch);

    if (CIN == kC) {
      const float* pa = pb + t * PS;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int toff = (tap / 3) * RS + G::kxoff(tap % 3);
        uint32_t bh[3][2], bl[3][2];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          const float2 wv = wsm[((ch * 9 + tap) * 3 + nt) * 32 + lane];
          split_tf32(wv.x, bh[nt][0], bl[nt][0]);
          split_tf32(wv.y, bh[nt][1], bl[nt][1]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t ah[4], al[4];
          split_tf32(pa[so[j][0] + toff], ah[0], al[0]);
          split_tf32(pa[so[j][1] + toff], ah[1], al[1]);
          split_tf32(pa[so[j][0] + toff + 4 * PS], ah[2], al[2]);
          split_tf32(pa[so[j][1] + toff + 4 * PS], ah[3], al[3]);
          mma3x3(acc[j], ah, al, bh, bl);
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint32_t ah[4], al[4];
          split_tf32(pb[koff[s][0] + so[j][0]], ah[0], al[0]);
          split_tf32(pb[koff[s][0] + so[j][1]], ah[1], al[1]);
          split_tf32(pb[koff[s][1] + so[j][0]], ah[2], al[2]);
          split_tf32(pb[koff[s][1] + so[j][1]], ah[3], al[3]);
          mma3x3(acc[j], ah, al, wr_h[s], wr_l[s]);
        }
    }

    if (ch == NCHUNK - 1) {               // the unit is complete: store y, accumulate the BatchNorm partials
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          int b, oy, ox;
          if (TW == 16) { b = cur.b0; oy = cur.r0 + 2 * warp + j; ox = cur.c0 + 8 * s + g; }
          else { b = cur.b0 + (warp >> 1); oy = (warp & 1) * 4 + 2 * j + s; ox = g; }
          const bool valid = b < B && oy < hout && ox < hout;
          float* yp = y + ((size_t)b * kC + 2 * t) * hw + oy * hout + ox;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float v = acc[j][nt][2 * s + e] + bias[8 * nt + 2 * t + e];
              if (valid) {
                yp[(8 * nt + e) * hw] = v;
                s1[nt][e] += v;
                s2[nt][e] += v * v;
              }
              acc[j][nt][2 * s + e] = 0.f;
            }
        }
    }

    if (has_next) stg.commit(patch + (buf ^ 1) * F::BUF, aff_s != nullptr ? aff_s + nch * NC : nullptr, tid);
    __syncthreads();
    buf ^= 1;
    u = nu;
    ch = nch;
    cur = nxt;
  }

  if (stat_part != nullptr) {
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = s1[nt][e], a2 = s2[nt][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {                  // over the 8 pixel rows (g) that share this channel
          a += __shfl_xor_sync(0xffffffffu, a, o);
          a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (g == 0) {
          red[warp * 2 * kC + 8 * nt + 2 * t + e] = a;
          red[warp * 2 * kC + kC + 8 * nt + 2 * t + e] = a2;
        }
      }
    __syncthreads();
    if (tid < 2 * kC) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) s += red[wv * 2 * kC + tid];
      stat_part[(size_t)blockIdx.x * 2 * kC + tid] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// dy staging for the gradient kernels: dy = BatchNorm-backward(dA, y) (bn_bwd_dy), zero outside the image / batch.
// The 7 per-channel coefficients live in shared memory as two float4 per channel.
// ------------------------------------------------------------------------------------------
struct BnBwdCoef {
  float mean, rstd, sc, sh, k0, k1, k2;
  __device__ __forceinline__ void load(const float* __restrict__ bnc, int c) {
    const float4 a = *reinterpret_cast<const float4*>(bnc + 8 * c), b = *reinterpret_cast<const float4*>(bnc + 8 * c + 4);
    mean = a.x; rstd = a.y; sc = a.z; sh = a.w;
    k0 = b.x; k1 = b.y; k2 = b.z;
  }
  __device__ __forceinline__ float apply(float yv, float da) const {
    return k0 * ((fmaf(sc, yv, sh) > 0.f ? da : 0.f) - k1 - (yv - mean) * rstd * k2);
  }
  __device__ __forceinline__ float4 apply4(const float4& yv, const float4& da) const {
    return make_float4(apply(yv.x, da.x), apply(yv.y, da.y), apply(yv.z, da.z), apply(yv.w, da.w));
  }
};
// bnc[24][8] = (mean, rstd, scale, shift, k0, k1, k2, -) from `aff` (4 x 24) and `coef` (3 x 24)
__device__ __forceinline__ void load_bnc(float* __restrict__ bnc, const float* __restrict__ aff, const float* __restrict__ coef,
                                         int tid) {
  if (tid < 8 * kC) {
    const int c = tid >> 3, f = tid & 7;
    bnc[tid] = f < 4 ? aff[f * kC + c] : (f < 7 ? coef[(f - 4) * kC + c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// Warp specialisation (WS = true) of the three backward kernels.  In the first version every warp alternated between
// staging (global loads, BatchNorm transform, shared-memory stores) and MMAs, and the two resident blocks of an SM drifted
// into phase: the tensor pipe idled ~60 % of the time (ncu: pipe_tensor 25 - 42 %).  Here a block has as many PRODUCER
// warps as consumer warps: producers stage unit i + 1 into the other half of a two-stage ring while the consumers
// multiply unit i; the hand-off is a pair of named barriers per stage (bar.arrive by one role, bar.sync by the other --
// full[s] = 1 + s, empty[s] = 3 + s), which also order the shared-memory traffic.  WS = false keeps the single-stage
// all-roles form (the 8x8 layer, whose units are too few to fill a ring).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Runs the unit loop of one block: stage(s, pos) by the producers, compute(s, pos) by the consumers.
template <bool WS, int NROLE, typename Walk, typename Stage, typename Compute>
__device__ __forceinline__ void unit_loop(int units, Walk walk, Stage&& stage, Compute&& compute) {
  const int n_my = (int)blockIdx.x < units ? (units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (WS) {
    if ((int)threadIdx.x >= NROLE) {
      for (int i = 0; i < n_my; ++i, walk.next()) {
        const int s = i & 1;
        if (i >= 2) bar_sync(3 + s, 2 * NROLE);          // the consumers have released this stage
        stage(s, walk.pos());
        bar_arrive(1 + s, 2 * NROLE);
      }
    } else {
      for (int i = 0; i < n_my; ++i, walk.next()) {
        const int s = i & 1;
        bar_sync(1 + s, 2 * NROLE);
        compute(s, walk.pos());
        if (i + 2 < n_my) bar_arrive(3 + s, 2 * NROLE);
      }
    }
  } else {
    for (int i = 0; i < n_my; ++i, walk.next()) {
      __syncthreads();                         // the previous unit's fragments have been read
      stage(0, walk.pos());
      __syncthreads();
      compute(0, walk.pos());
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient of the 24 -> 24 layers: dW[co][ci][tap] = sum_pixels dy[co][p] * act(in)[ci][tap(p)].
// GEMM  D[(tap, ci)][co] = sum_p A[(tap, ci)][p] * B[p][co]:  M = 216 = 27 row groups (8 channels of one tap) paired into
// 14 m-tiles (the 28th group is a dummy), N = 24 = 3 n-tiles, K = 8 consecutive output pixels of one row per MMA.
// 7 consumer warps, warp w owns m-tiles 2w, 2w+1 (24 accumulators) for ALL pixels; persistent over units of 128 output
// pixels (TW = 16: 8 rows x 16 columns of one image, TW = 8: the 8x8 outputs of two images); the accumulators live in
// registers across units, one partial [24][24][9] per block, summed by a fixed-order reduction afterwards.
// ------------------------------------------------------------------------------------------
template <int TW>
struct WgCfg {
  static constexpr int NT = 224;
  static constexpr int NIMG = TW == 16 ? 1 : 2;
  static constexpr int NR = 17;
  static constexpr int IS = NR * Geo<TW>::RS;
  static constexpr int PS = pad32(NIMG * IS, 4);
  static constexpr int LD = 132;                                     // dy: [24][128 pixels + 4]
  static constexpr int STAGE = kC * PS + kC * LD;                    // floats per ring stage
  static constexpr size_t smem(bool ws) { return (size_t)((ws ? 2 : 1) * STAGE + 2 * kC + 8 * kC) * sizeof(float); }
};

template <int TW, bool WS>
__global__ void __launch_bounds__(WS ? 448 : 224, WS ? 1 : 2)
conv_wgrad_tc_kernel(const float* __restrict__ in, const float* __restrict__ in_aff, const float* __restrict__ yout,
                     const float* __restrict__ dA, const float* __restrict__ aff_out, const float* __restrict__ coef,
                     float* __restrict__ part, int B, int hin, int hout, int tiles_x, int tiles, int units) {
  using W = WgCfg<TW>;
  using G = Geo<TW>;
  constexpr int PS = W::PS, IS = W::IS, RS = G::RS, LD = W::LD, NT = W::NT;
  extern __shared__ __align__(16) float wg_tc_smem[];
  float* ring = wg_tc_smem;                                // [stages][patch [24][PS] | dy [24][LD]]
  float* affs = ring + (WS ? 2 : 1) * W::STAGE;            // scale[24], shift[24] of the producing layer
  float* bnc = affs + 2 * kC;                              // [24][8]

  const int tid = WS && threadIdx.x >= NT ? threadIdx.x - NT : threadIdx.x;      // index within the role
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hin2 = hin * hin, hw = hout * hout;
  if (threadIdx.x < 2 * kC) affs[threadIdx.x] = in_aff[2 * kC + threadIdx.x];
  load_bnc(bnc, aff_out, coef, threadIdx.x);
  __syncthreads();

  // A rows: m-tile mt, half h (fragment rows g / g + 8) = row group q = 4*warp + 2*mt + h = chunk*9 + tap
  int arow[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int q = 4 * warp + 2 * mt + h;
      q = q < 27 ? q : 26;
      const int ch = q / 9, tap = q - 9 * ch;
      arow[mt][h] = (8 * ch + g) * PS + (tap / 3) * RS + G::kxoff(tap % 3) + t;
    }
  const int brow = g * LD + t;

  float acc[2][3][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  auto stage = [&](int s, const UnitPos& up) {
    float* patch = ring + s * W::STAGE;
    float* dys = patch + kC * PS;
    // dy: 24 channels x 128 pixels; lane -> 4 pixels of a row (fixed), warp -> channels w, w + 7, ..
    const int px = lane * 4;
    int b, oy, ox;
    if (TW == 16) { b = up.b0; oy = up.r0 + (px >> 4); ox = up.c0 + (px & 15); }
    else { b = up.b0 + (px >> 6); oy = (px >> 3) & 7; ox = px & 7; }
    const bool pv = b < B && oy < hout && ox < hout;
    const size_t off = pv ? (size_t)b * kC * hw + oy * hout + ox : 0;
    auto stage_dy = [&](const float4 (&yv)[4], const float4 (&dav)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = warp + 7 * i;
        if (co < kC) {
          BnBwdCoef k;
          k.load(bnc, co);
          *reinterpret_cast<float4*>(dys + co * LD + px) = pv ? k.apply4(yv[i], dav[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    float4 yv[4], dav[4];
    auto load_dy = [&]() {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = warp + 7 * i < kC ? warp + 7 * i : 0;
        yv[i] = *reinterpret_cast<const float4*>(yout + off + co * hw);
        dav[i] = *reinterpret_cast<const float4*>(dA + off + co * hw);
      }
    };
    if (TW == 16) {
      // ALL of the unit's loads in flight at once (one memory round trip per unit): rows 0..8 / 9..16 of the 24 input
      // planes (thread -> one plane and vector column, walking down the rows) and the (y, dA) vectors of dy
      const float* inb = in + (size_t)up.b0 * kC * hin2;
      PatchStager<TW, 1, kC, 9, PS, IS, NT, false> s1;
      PatchStager<TW, 1, kC, 8, PS, IS, NT, false> s2;
      s1.issue(inb, 1, kC * hin2, hin2, 2 * up.r0 - 1, 2 * up.c0 - 1, hin, tid);
      s2.issue(inb, 1, kC * hin2, hin2, 2 * up.r0 - 1 + 9, 2 * up.c0 - 1, hin, tid);
      load_dy();
      s1.commit(patch, affs, tid);
      s2.commit(patch + 9 * RS, affs, tid);
      stage_dy(yv, dav);
    } else {
      load_dy();
#pragma unroll 1
      for (int c0 = 0; c0 < kC; c0 += 12) {       // channels 0..11 / 12..23 (bounds the registers of the staging)
        PatchStager<TW, W::NIMG, 12, W::NR, PS, IS, NT, false> stg;
        stg.issue(in + ((size_t)up.b0 * kC + c0) * hin2, B - up.b0, kC * hin2, hin2, 2 * up.r0 - 1, 2 * up.c0 - 1, hin, tid);
        stg.commit(patch + c0 * PS, affs + c0, tid);
      }
      stage_dy(yv, dav);
    }
  };

  auto compute = [&](int s, const UnitPos&) {
    const float* patch = ring + s * W::STAGE;
    const float* dys = patch + kC * PS;
#pragma unroll 4
    for (int ks = 0; ks < 16; ++ks) {
      const int poff = TW == 16 ? (2 * (ks >> 1)) * RS + (ks & 1) * 8 : (ks >> 3) * IS + (2 * (ks & 7)) * RS;
      uint32_t bh[3][2], bl[3][2];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        split_tf32(dys[brow + 8 * nt * LD + ks * 8], bh[nt][0], bl[nt][0]);
        split_tf32(dys[brow + 8 * nt * LD + ks * 8 + 4], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t ah[4], al[4];
        split_tf32(patch[arow[mt][0] + poff], ah[0], al[0]);
        split_tf32(patch[arow[mt][1] + poff], ah[1], al[1]);
        split_tf32(patch[arow[mt][0] + poff + 4], ah[2], al[2]);
        split_tf32(patch[arow[mt][1] + poff + 4], ah[3], al[3]);
        mma3x3(acc[mt], ah, al, bh, bl);
      }
    }
  };

  unit_loop<WS, NT>(units, UnitWalk<W::NIMG, 8, 16>(tiles, tiles_x), stage, compute);

  if (WS && threadIdx.x >= NT) return;
  float* out = part + (size_t)blockIdx.x * (kC * kC * 9);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = 4 * warp + 2 * mt + h;
      if (q < 27) {
        const int ch = q / 9, tap = q - 9 * ch, ci = 8 * ch + g;
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) out[((size_t)(8 * nt + 2 * t + e) * kC + ci) * 9 + tap] = acc[mt][nt][2 * h + e];
      }
    }
}

// ------------------------------------------------------------------------------------------
// weight gradient of the RGB layer: dW[co][ci][tap], 24 x 27.  GEMM  D[co][(tap, ci)] = sum_p dy[co][p] * in[(tap, ci)][p]:
// M = 24 (two m-tiles, the upper half of the second one is zero), N = 27 padded to 32 = 4 n-tiles, K = 8 pixels.
// The output is tiny, so the PIXELS are split over the 8 consumer warps (warp w: rows 2w, 2w+1 of a 16x16 output tile) and
// the 8 per-warp results are combined by a fixed-order sum at the end of the persistent block.
// ------------------------------------------------------------------------------------------
struct Wg3Cfg {
  static constexpr int NR = 33;
  static constexpr int PS = pad32(NR * Geo<16>::RS, 4);              // 1188 (== 4 mod 32)
  static constexpr int LD = 260;                                     // dy: [24][256 pixels + 4]
  static constexpr int RED = 8 * kC * 27;                            // aliases stage 0 at the end
  static constexpr int STAGE = 3 * PS + kC * LD;
  static constexpr size_t smem(bool ws) { return (size_t)((ws ? 2 : 1) * STAGE + 8 * kC) * sizeof(float); }
  static_assert(STAGE >= RED, "the final reduction aliases one stage");
};

template <bool U8, bool WS>
__global__ void __launch_bounds__(WS ? 512 : 256, WS ? 1 : 2)
conv_wgrad3_tc_kernel(const void* __restrict__ in, const float* __restrict__ yout, const float* __restrict__ dA,
                      const float* __restrict__ aff_out, const float* __restrict__ coef, float* __restrict__ part, int B,
                      int hin, int hout, int tiles_x, int tiles, int units) {
  using G = Geo<16>;
  constexpr int PS = Wg3Cfg::PS, RS = G::RS, LD = Wg3Cfg::LD, NT = 256;
  extern __shared__ __align__(16) float wg3_smem[];
  float* ring = wg3_smem;                   // [stages][patch [3][PS] | dy [24][LD]]
  float* red = wg3_smem;                    // [8][24*27] at the end
  float* bnc = wg3_smem + (WS ? 2 : 1) * Wg3Cfg::STAGE;

  const int tid = WS && threadIdx.x >= NT ? threadIdx.x - NT : threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hin2 = hin * hin, hw = hout * hout;
  load_bnc(bnc, aff_out, coef, threadIdx.x);
  __syncthreads();

  int noff[4];                              // B rows: n = 8nt + g = tap*3 + ci
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int n = 8 * nt + g;
    const int tap = n / 3, ci = n - 3 * tap;
    noff[nt] = n < 27 ? ci * PS + (tap / 3) * RS + G::kxoff(tap % 3) + t : t;
  }

  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  auto stage = [&](int s, const UnitPos& up) {
    float* patch = ring + s * Wg3Cfg::STAGE;
    float* dys = patch + 3 * PS;
    PatchStager<16, 1, 3, Wg3Cfg::NR, PS, 0, NT, U8> stg;
    const size_t base = (size_t)up.b0 * 3 * hin2;
    const void* inb = U8 ? static_cast<const void*>(static_cast<const unsigned char*>(in) + base)
                         : static_cast<const void*>(static_cast<const float*>(in) + base);
    stg.issue(inb, 1, 3 * hin2, hin2, 2 * up.r0 - 1, 2 * up.c0 - 1, hin, tid);
    // dy of the 16x16 tile while the image loads are in flight: thread -> 4 pixels of a row (fixed), channels tid/64 + 4i
    const int px = (tid & 63) * 4;
    const int oy = up.r0 + (px >> 4), ox = up.c0 + (px & 15);
    const bool pv = oy < hout && ox < hout;
    const size_t off = (size_t)up.b0 * kC * hw + oy * hout + ox;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int co = (tid >> 6) + 4 * i;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pv) {
        BnBwdCoef k;
        k.load(bnc, co);
        o = k.apply4(*reinterpret_cast<const float4*>(yout + off + co * hw), *reinterpret_cast<const float4*>(dA + off + co * hw));
      }
      *reinterpret_cast<float4*>(dys + co * LD + px) = o;
    }
    stg.commit(patch, nullptr, tid);
  };

  auto compute = [&](int s, const UnitPos&) {
    const float* patch = ring + s * Wg3Cfg::STAGE;
    const float* dys = patch + 3 * PS;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int row = 2 * warp + (ks >> 1), ox0 = (ks & 1) * 8;
      const int px = row * 16 + ox0 + t, poff = (2 * row) * RS + ox0;
      uint32_t ah[2][4], al[2][4];
      split_tf32(dys[g * LD + px], ah[0][0], al[0][0]);
      split_tf32(dys[(g + 8) * LD + px], ah[0][1], al[0][1]);
      split_tf32(dys[g * LD + px + 4], ah[0][2], al[0][2]);
      split_tf32(dys[(g + 8) * LD + px + 4], ah[0][3], al[0][3]);
      split_tf32(dys[(g + 16) * LD + px], ah[1][0], al[1][0]);
      split_tf32(dys[(g + 16) * LD + px + 4], ah[1][2], al[1][2]);
      ah[1][1] = al[1][1] = ah[1][3] = al[1][3] = 0u;
      uint32_t bh[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        split_tf32(patch[noff[nt] + poff], bh[nt][0], bl[nt][0]);
        split_tf32(patch[noff[nt] + poff + 4], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int pass = 0; pass < 3; ++pass)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t (&a)[4] = pass == 0 ? al[mt] : ah[mt];
            const uint32_t (&b)[2] = pass == 1 ? bl[nt] : bh[nt];
            mma_tf32(acc[mt][nt], a[0], a[1], a[2], a[3], b[0], b[1]);
          }
    }
  };

  unit_loop<WS, NT>(units, UnitWalk<1, 16, 16>(tiles, tiles_x), stage, compute);

  __syncthreads();                          // every role is done with the ring: it becomes the reduction buffer
  if (!WS || threadIdx.x < NT) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int co = 16 * mt + 8 * h + g, n = 8 * nt + 2 * t + e;
            if (co < kC && n < 27) red[(warp * kC + co) * 27 + n] = acc[mt][nt][2 * h + e];
          }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < kC * 27; idx += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv * kC * 27 + idx];
    const int co = idx / 27, n = idx - 27 * co;
    const int tap = n / 3, ci = n - 3 * tap;
    part[(size_t)blockIdx.x * (kC * 27) + (co * 3 + ci) * 9 + tap] = s;
  }
}

// ------------------------------------------------------------------------------------------
// data gradient (layers 2..4): dA_prev[ci][ih][iw] = sum_{co, kh, kw} dy[co][oh][ow] W[co][ci][kh][kw], ih = 2 oh + kh - 1.
// Per parity class (ih & 1, iw & 1) a GEMM  D[quad][ci] = sum_{tap in class, co} dy[co][quad + shift(tap)] * W[co][ci][tap]
// (1, 2, 2, 4 taps): M = 16 quads (two 8-quad row segments), N = 24, K = 8 output channels of one tap.  The A fragment
// of a shift (0/1, 0/1) of the dy window is loaded once and feeds every tap that uses it.  8 consumer warps; unit =
// 16x16 quads (32x32 input pixels) of one image (TW = 16) or the 8x8 quads of four images (TW = 8); weight fragments
// pre-split (hi, lo) in shared memory.
// ------------------------------------------------------------------------------------------
template <int TW>
struct DgCfg {
  static constexpr int NIMG = TW == 16 ? 1 : 4;
  static constexpr int NR = TW + 1;
  static constexpr int RS = TW == 16 ? 20 : 12;
  static constexpr int IS = NR * RS;
  static constexpr int PS = pad32(NIMG * IS, 8);
  static constexpr int STAGE = kC * PS;
  static constexpr size_t smem(bool ws) {
    return (size_t)((ws ? 2 : 1) * STAGE + 2 * kC * kC * 9 + 8 * kC + 4 * kC + 8 * 2 * kC) * sizeof(float);
  }
};

// y_in / aff_in / bn_part: the raw output and BatchNorm block of the PRODUCING layer (the one whose dA this kernel writes).
// Its BatchNorm-backward sums  sum g, sum g * xhat  (g = dA where relu(bn(y)) > 0) are taken from the accumulators in the
// epilogue -- one partial [24][2] per block -- instead of a separate pass that re-reads dA and y_in (bn_bwd_reduce_kernel:
// 93 us for the first layer at batch 640).
template <int TW, bool WS>
__global__ void __launch_bounds__(WS ? 512 : 256, WS ? 1 : 2)
conv_dgrad_tc_kernel(const float* __restrict__ yout, const float* __restrict__ dAout, const float* __restrict__ aff_out,
                     const float* __restrict__ coef, const float* __restrict__ w, float* __restrict__ dA,
                     const float* __restrict__ y_in, const float* __restrict__ aff_in, float* __restrict__ bn_part, int B,
                     int hin, int hout, int tiles_x, int tiles, int units) {
  using D = DgCfg<TW>;
  constexpr int PS = D::PS, IS = D::IS, RS = D::RS, NR = D::NR, NT = 256;
  extern __shared__ __align__(16) float dg_smem[];
  float* ring = dg_smem;                                        // [stages][24][PS]
  float4* wsm = reinterpret_cast<float4*>(dg_smem + (WS ? 2 : 1) * D::STAGE);  // [tap][ks][nt][lane] (b0 hi, b1 hi, b0 lo, b1 lo)
  float* bnc = reinterpret_cast<float*>(wsm) + 2 * kC * kC * 9;
  float* affi = bnc + 8 * kC;                                  // (mean, rstd, scale, shift) x 24 of the producing layer
  float* bnred = affi + 4 * kC;                                // [8 warps][24][2]

  const int tid = WS && threadIdx.x >= NT ? threadIdx.x - NT : threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int hin2 = hin * hin, hw = hout * hout;
  load_bnc(bnc, aff_out, coef, threadIdx.x);
  if (threadIdx.x < 4 * kC) affi[threadIdx.x] = aff_in[threadIdx.x];
  for (int i = threadIdx.x; i < 8 * 2 * kC; i += blockDim.x) bnred[i] = 0.f;

  // B[k = co][n = ci] of (tap, ks, nt): b0 = W[8ks + t][8nt + g][tap], b1 = W[8ks + t + 4][8nt + g][tap]
  for (int idx = threadIdx.x; idx < 9 * 3 * 3 * 32; idx += blockDim.x) {
    const int ln = idx & 31, q = idx >> 5;
    const int nt = q % 3, ks = (q / 3) % 3, tap = q / 9;
    const int co = 8 * ks + (ln & 3), ci = 8 * nt + (ln >> 2);
    uint32_t h0, l0, h1, l1;
    split_tf32(w[((size_t)co * kC + ci) * 9 + tap], h0, l0);
    split_tf32(w[((size_t)(co + 4) * kC + ci) * 9 + tap], h1, l1);
    wsm[idx] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
  }
  __syncthreads();

  int dso[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (TW == 16) dso[j][s] = (2 * warp + j) * RS + 8 * s + g;
      else dso[j][s] = (warp >> 1) * IS + ((warp & 1) * 4 + 2 * j + s) * RS + g;
    }

  // dy window staging: thread -> one fixed (row, slot) of the (TW+1) x (TW+1) window (slot < TW/4: an aligned float4,
  // slot == TW/4: the last column), walking over the (image, channel) planes with stride LANES
  constexpr int PERROW = TW / 4 + 1, POS = NR * PERROW, LANES = 256 / POS, PLANES = D::NIMG * kC;
  constexpr int SITER = (PLANES + LANES - 1) / LANES;
  const int s_lane = tid / POS, s_pos = tid - s_lane * POS;
  const int s_r = s_pos / PERROW, s_v = s_pos - s_r * PERROW;

  auto stage = [&](int st, const UnitPos& up) {
    float* dys = ring + st * D::STAGE;
    if (s_lane < LANES) {
      const int oh = up.r0 + s_r, ow = up.c0 + (s_v < TW / 4 ? 4 * s_v : TW);
      const bool pv = oh < hout && ow < hout;
      const int poff = pv ? oh * hout + ow : 0;
      float* d = dys + s_r * RS + (s_v < TW / 4 ? 4 * s_v : TW);
      // every load of the thread in flight before the first use (branch-free: an invalid item reads a valid address and
      // is masked), then the BatchNorm-backward transform
      float4 yv[SITER], dv[SITER];
#pragma unroll
      for (int i = 0; i < SITER; ++i) {
        const int pl = min(s_lane + LANES * i, PLANES - 1);
        const int img = D::NIMG == 1 ? 0 : pl / kC, co = pl - img * kC;
        const size_t off = ((size_t)min(up.b0 + img, B - 1) * kC + co) * hw + poff;
        if (s_v < TW / 4) {
          yv[i] = *reinterpret_cast<const float4*>(yout + off);
          dv[i] = *reinterpret_cast<const float4*>(dAout + off);
        } else {
          yv[i] = make_float4(yout[off], 0.f, 0.f, 0.f);
          dv[i] = make_float4(dAout[off], 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < SITER; ++i) {
        const int pl = s_lane + LANES * i;
        if (pl < PLANES) {
          const int img = D::NIMG == 1 ? 0 : pl / kC, co = pl - img * kC;
          const bool ok = pv && up.b0 + img < B;
          BnBwdCoef k;
          k.load(bnc, co);
          float* dd = d + co * PS + img * IS;
          if (s_v < TW / 4) *reinterpret_cast<float4*>(dd) = ok ? k.apply4(yv[i], dv[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
          else *dd = ok ? k.apply(yv[i].x, dv[i].x) : 0.f;
        }
      }
    }
  };

  auto compute = [&](int st, const UnitPos& up) {
    const float* dys = ring + st * D::STAGE;
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      const float4* wq = wsm;
      asm volatile("" : "+l"(wq));          // keeps the 81 weight-fragment loads inside the loop (hoisted, they spill)
      float acc[4][3][4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) acc[c][nt][0] = acc[c][nt][1] = acc[c][nt][2] = acc[c][nt][3] = 0.f;

#pragma unroll
      for (int sh = 0; sh < 4; ++sh) {
        const int di = sh >> 1, dj = sh & 1;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
          const float* pa = dys + (8 * ks + t) * PS + di * RS + dj;
          uint32_t ah[4], al[4];
          split_tf32(pa[dso[j][0]], ah[0], al[0]);
          split_tf32(pa[dso[j][1]], ah[1], al[1]);
          split_tf32(pa[dso[j][0] + 4 * PS], ah[2], al[2]);
          split_tf32(pa[dso[j][1] + 4 * PS], ah[3], al[3]);
          // (tap, class) pairs fed by this shift; class = (ih & 1) * 2 + (iw & 1)
          constexpr int NP[4] = {4, 2, 2, 1};
          constexpr int TAPS[4][4] = {{4, 5, 7, 8}, {3, 6, 0, 0}, {1, 2, 0, 0}, {0, 0, 0, 0}};
          constexpr int CLS[4][4] = {{0, 1, 2, 3}, {1, 3, 0, 0}, {2, 3, 0, 0}, {3, 0, 0, 0}};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            if (p < NP[sh]) {
              uint32_t bh[3][2], bl[3][2];
#pragma unroll
              for (int nt = 0; nt < 3; ++nt) {
                const float4 wv = wq[((TAPS[sh][p] * 3 + ks) * 3 + nt) * 32 + lane];
                bh[nt][0] = __float_as_uint(wv.x); bh[nt][1] = __float_as_uint(wv.y);
                bl[nt][0] = __float_as_uint(wv.z); bl[nt][1] = __float_as_uint(wv.w);
              }
              mma3x3(acc[CLS[sh][p]], ah, al, bh, bl);
            }
          }
        }
      }

      float bs[3][2][2];                     // this m-tile's (sum g, sum g * xhat) of channels 8nt + 2t + e
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) bs[nt][0][0] = bs[nt][0][1] = bs[nt][1][0] = bs[nt][1][1] = 0.f;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        int b, qi, qj;
        if (TW == 16) { b = up.b0; qi = up.r0 + 2 * warp + j; qj = up.c0 + 8 * s + g; }
        else { b = up.b0 + (warp >> 1); qi = (warp & 1) * 4 + 2 * j + s; qj = g; }
        if (b < B && 2 * qi < hin && 2 * qj < hin) {
          const size_t off = ((size_t)b * kC + 2 * t) * hin2 + (2 * qi) * hin + 2 * qj;
          float* o = dA + off;
          const float* yi = y_in + off;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float2 d0 = make_float2(acc[0][nt][2 * s + e], acc[1][nt][2 * s + e]);
              const float2 d1 = make_float2(acc[2][nt][2 * s + e], acc[3][nt][2 * s + e]);
              *reinterpret_cast<float2*>(o + (8 * nt + e) * hin2) = d0;
              *reinterpret_cast<float2*>(o + (8 * nt + e) * hin2 + hin) = d1;
              const float2 y0 = *reinterpret_cast<const float2*>(yi + (8 * nt + e) * hin2);
              const float2 y1 = *reinterpret_cast<const float2*>(yi + (8 * nt + e) * hin2 + hin);
              const int c = 8 * nt + 2 * t + e;
              const float mean = affi[c], rstd = affi[kC + c], sc = affi[2 * kC + c], sh = affi[3 * kC + c];
              const float g0 = fmaf(sc, y0.x, sh) > 0.f ? d0.x : 0.f, g1 = fmaf(sc, y0.y, sh) > 0.f ? d0.y : 0.f;
              const float g2 = fmaf(sc, y1.x, sh) > 0.f ? d1.x : 0.f, g3 = fmaf(sc, y1.y, sh) > 0.f ? d1.y : 0.f;
              bs[nt][e][0] += (g0 + g1) + (g2 + g3);
              bs[nt][e][1] += ((g0 * (y0.x - mean) + g1 * (y0.y - mean)) + (g2 * (y1.x - mean) + g3 * (y1.y - mean))) * rstd;
            }
        }
      }
      // over the 8 quad lanes (g) that share a channel, then into this warp's row of the block partial
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            float v = bs[nt][e][k];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) bnred[(warp * kC + 8 * nt + 2 * t + e) * 2 + k] += v;
          }
    }
  };

  unit_loop<WS, NT>(units, UnitWalk<D::NIMG, TW, TW>(tiles, tiles_x), stage, compute);

  __syncthreads();
  if (threadIdx.x < 2 * kC) {                // fixed-order sum over the consumer warps: one partial per block
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += bnred[wv * 2 * kC + threadIdx.x];
    bn_part[(size_t)blockIdx.x * 2 * kC + threadIdx.x] = v;
  }
}

}  // namespace ctc
}  // namespace rn
