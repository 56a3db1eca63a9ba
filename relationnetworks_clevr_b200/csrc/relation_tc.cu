// relation_tc.cu -- tcgen05 / TMEM implementation of g-MLP layers 1..3 for G == 256 (original-fp, ir-fp and
// the grid sweep).  See DESIGN.md "Kernels" for the pipeline description.
//
// Chain kernel rn_g_chain_kernel<MODE> (training forward / eval forward / data gradient share one skeleton;
// persistent, one CTA per SM, 384 threads -- 512 with the optional generator warpgroup, see the GEN template flag):
//   warp 0      weight producer: streams pre-swizzled fp16 weight chunks (32 KB = 256 out x 64 in) from L2 into a
//               3-stage shared-memory ring with cp.async.bulk (TMA engine) + mbarrier complete_tx
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma M128 x N256 x K16 (kind::f16, fp32 accumulate in
//               TMEM), 4 per chunk; tcgen05.commit releases ring stages and publishes finished accumulators
//   warp 2      TMEM allocator (512 columns = two 128x256 fp32 accumulators)
//   warps 4-7   epilogue warpgroup of tile slot 0;  warps 8-11: slot 1.  Per tile: generate the layer-1 operand
//               H1 = relu(U[c] + Vb[a]) straight into the swizzled A buffer (the 4096x180 pair matrix never exists),
//               then per layer: tcgen05.ld accumulator -> +bias -> ReLU -> fp16 -> next layer's A operand in smem
//               (activations never leave the SM); last layer: ReLU + warp-shuffle column sums (the pair-sum).
//   Two tile slots ping-pong so one slot's epilogue overlaps the other slot's MMAs.
// Training streams the fp16 operand images H2, H3 / dZ1..dZ3 and the ReLU sign bits to HBM for the weight-gradient
// kernel rn_g_wgrad_kernel (256x256 fp32 accumulator resident in TMEM, images read back as MN-major operands), which
// REGENERATES H1 (from U / V') and dZ4 (from the sign bits) instead of reading images of them.
// Precision: A operands fp16; weights W = W_hi + W_lo (two fp16 MMAs per K-step, "parity" forward) or W_hi only
// ("fast", and the data gradient by default).
// Diagnostics: RN_B200_DBG bit 0/1/2 = timing ablations (half the weight traffic / no epilogue arithmetic / no weight
// streaming; results are garbage), bit 3 = in-kernel phase cycle counters read by tests/diag_tc_prof.py.
#include "relation.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace rn {

using namespace ptx;

constexpr int kG = 256;                 // g width
constexpr int kTileM = 128;             // pair rows per tile
constexpr int kKC = 64;                 // K elements per chunk (one 128-byte swizzle row)
constexpr int kNKC = kG / kKC;          // 4 chunks per layer
constexpr int kAChunk = kTileM * kKC * 2;     // 16 KB
constexpr int kATile = kAChunk * kNKC;        // 64 KB
constexpr int kWChunk = kG * kKC * 2;         // 32 KB
constexpr int kStages = 3;
constexpr int kTcLayers = 3;            // g layers 1..3 run on the tensor cores
constexpr int kFwdThreads = 384;
constexpr int kGenThreads = 512;          // + one generator warpgroup (warps 12..15)
constexpr int kSmemA = 0;
constexpr int kSmemW = 2 * kATile;
constexpr int kSmemBar = kSmemW + kStages * kWChunk;
constexpr int kSmemTotal = kSmemBar + 256;
constexpr int kSmemLaunch = kSmemTotal + 1024;   // slack to align the carve-up to 1024 B (SWIZZLE_128B atoms)
constexpr uint32_t kIdescFwd = idesc_f16(kTileM, kG, 0, 0);
constexpr uint32_t kIdescFwd2 = idesc_f16(2 * kTileM, kG, 0, 0);   // CTA pair: M = 256
constexpr int kStages2 = 6;                 // CTA-pair mode: each CTA stages only its N-half (16 KB) of a weight chunk
constexpr int kWHalf = kWChunk / 2;

// byte offset of element (row, col) inside a [rows x 64] fp16 K-major SWIZZLE_128B chunk
__host__ __device__ inline uint32_t sw128_offset(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ (row & 7)) & 7) << 4) + (col & 7) * 2);
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 [G, fan] -> fp16 hi/lo chunk images.  transpose == 0: B[N = out][K = in] (forward);
// transpose == 1: B[N = in][K = out] (data gradient).  One thread per 16-byte group (8 K elements).
// image index: ((layer * 2 + pass) * 4 + kchunk) * 32 KB
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* w[kTcLayers];
  int ld[kTcLayers];
};

__global__ void pack_weights_kernel(PackArgs args, __half* __restrict__ out, int transpose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over layers * 256 rows * 32 groups
  if (idx >= kTcLayers * kG * (kG / 8)) return;
  const int layer = idx / (kG * (kG / 8));
  const int rem = idx % (kG * (kG / 8));
  const int nrow = rem / (kG / 8);        // N index
  const int kg = rem % (kG / 8);          // group of 8 K elements
  const float* w = args.w[layer];
  const int ld = args.ld[layer];
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = kg * 8 + e * 2 + h;
      v[h] = transpose ? w[(size_t)kk * ld + nrow] : w[(size_t)nrow * ld + kk];
    }
    const __half h0 = __float2half_rn(v[0]), h1 = __float2half_rn(v[1]);
    const float r0 = v[0] - __half2float(h0), r1 = v[1] - __half2float(h1);
    __half2 hh = __halves2half2(h0, h1);
    __half2 ll = __floats2half2_rn(r0, r1);
    hi[e] = *reinterpret_cast<uint32_t*>(&hh);
    lo[e] = *reinterpret_cast<uint32_t*>(&ll);
  }
  const int kc = kg / 8;
  const uint32_t off = sw128_offset(nrow, (kg % 8) * 8);
  char* base = reinterpret_cast<char*>(out);
  *reinterpret_cast<uint4*>(base + ((size_t)(layer * 2 + 0) * kNKC + kc) * kWChunk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + ((size_t)(layer * 2 + 1) * kNKC + kc) * kWChunk + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ------------------------------------------------------------------------------------------------
// chain kernel: forward (eval / training) and data-gradient share one pipeline skeleton
// ------------------------------------------------------------------------------------------------
enum ChainMode { kFwdEval = 0, kFwdTrain = 1, kDgrad = 2 };

struct ChainParams {
  const __half* wpack;             // [3][2][4] x 32 KB chunk images (forward: W; dgrad: W^T)
  int n;
  int tiles_per_sample;
  int num_tiles;
  int passes;                      // 2 = parity (hi + lo), 1 = fast
  // forward
  const float* U;                  // [B, n, 256]
  const float* Vb;                 // [B, n, 256]  (V + beta0)
  const float* bias[kTcLayers];    // g layer l+1 bias; element [b * bias_stride + col]
  long long bias_stride[kTcLayers];
  float* xg_part;                  // [tiles][4][256]
  __half* saveH;                   // [tiles][3] x 64 KB operand images H1, H2, H3 (training)
  uint32_t* masks;                 // [4][tiles][128][8] sign bits of Z1..Z4 (training; read by dgrad)
  // dgrad
  const float* dxg;                // [B, 256]
  const float* scale;              // [0] = S (power of two applied to dxg), [1] = 1/S
  __half* dZ;                      // [tiles][4] x 64 KB images dZ1..dZ4 (scaled by S)
  int dbg;                         // timing ablations (RN_B200_DBG; results are garbage when nonzero)
  int skip_dz4_image;              // dgrad: the weight-gradient kernel regenerates dZ4 from the sign bits
  int skip_h1_image;               // training forward: the layer-1 weight-gradient kernel regenerates H1 from U / V'
  int sched;                       // MMA job order (for_each_chain_job)
  const float* U4;                 // 3-pass forward: U as [B][64 column groups][n][4]
  int m1_layout0;                  // dgrad: the Z1 sign bits use the same layout as Z2..Z4 (written by the 3-pass forward)
};

struct Bars {
  uint64_t w_full[kStages2];      // 1-CTA mode uses the first kStages entries
  uint64_t w_empty[kStages2];
  uint64_t peer_full[kStages2];   // CTA-pair mode, leader only: the peer CTA's half of the chunk has landed
  uint64_t a_full[2];
  uint64_t acc_full[2];
  uint64_t a_free[2];             // generator mode: the slot's A buffer may be overwritten with the next tile's operand
  uint64_t h1_done[2];            // generator mode: the image store of the generated operand has finished reading A
  uint32_t tmem_base;
};

__device__ __forceinline__ int tiles_of_cta(int num_tiles) {
  return (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
}

// Sum x[0..31] over the 32 lanes (rows) of the warp: afterwards lane L holds the total of element L.
__device__ __forceinline__ float warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int e = 0; e < off; ++e) {
      const float send = upper ? x[e] : x[e + off];
      const float keep = upper ? x[e + off] : x[e];
      x[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return x[0];
}

// fp32 pair -> packed fp16x2 with ReLU folded into the conversion (F2FP.RELU): lo -> low half, hi -> high half
__device__ __forceinline__ uint32_t pack_relu_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 0xFFFF in each half-word of the result where the corresponding fp16 of `w` is > 0 (one HSET2 per pair)
__device__ __forceinline__ uint32_t half2_pos_mask(uint32_t w) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  return __hgt2_mask(h, __half2half2(__ushort_as_half((unsigned short)0)));
}
// ReLU sign-bit words: a 32-bit word covers 32 consecutive columns; element e sits at bit (e/2) + 16*(e%2), i.e. the
// even columns fill the low half-word and the odd ones the high half-word, so one packed fp16 pair updates its two
// bits with a single LOP3: bits |= half2_pos_mask(pair) & mask_pair_const(i).
__host__ __device__ constexpr uint32_t mask_pair_const(int i) { return (1u << i) | (1u << (16 + i)); }
__host__ __device__ constexpr int mask_pos(int e) { return (e >> 1) + 16 * (e & 1); }

// Forward generation: H1 rows of one tile -> swizzled fp16 A operand (the pair matrix never exists anywhere).
// Warp `q` (0..3) of the warpgroup owns rows [32q, 32q+32).  Lane mapping: 4 rows x 8 sixteen-byte groups per
// step -> 256-byte coalesced reads of U, conflict-free STS.  M1 sign bits use a lane-local layout
// word j = (col % 64) / 8, bit = (col / 64) * 8 + col % 8.
template <bool SAVE, bool ONE_A>
__device__ __forceinline__ void generate_h1_rows(const ChainParams& p, int tile, char* a_tile, int q, int lane) {
  const int b = tile / p.tiles_per_sample;
  const int p0 = (tile % p.tiles_per_sample) * kTileM;
  const float* Ub = p.U + (size_t)b * p.n * kG;
  const float* Vb = p.Vb + (size_t)b * p.n * kG;
  const int sub = lane >> 3, j = lane & 7;
  float4 vr[ONE_A ? kNKC : 1][2];             // ONE_A: the tile lies inside one value of `a` (n % 128 == 0): its Vb row stays in registers
  if (ONE_A) {
    const float4* vp = reinterpret_cast<const float4*>(Vb + (size_t)(p0 / p.n) * kG + j * 8);
#pragma unroll
    for (int kc = 0; kc < kNKC; ++kc) {
      vr[kc][0] = __ldg(vp + kc * 16);
      vr[kc][1] = __ldg(vp + kc * 16 + 1);
    }
  }
#pragma unroll 2
  for (int g = 0; g < 8; ++g) {
    const int row = q * 32 + g * 4 + sub;
    const int pr = p0 + row;
    const int a = pr / p.n, c = pr - a * p.n;
    const float4* up = reinterpret_cast<const float4*>(Ub + (size_t)c * kG + j * 8);
    const float4* vp = reinterpret_cast<const float4*>(Vb + (size_t)a * kG + j * 8);
    uint32_t bits = 0;
#pragma unroll
    for (int kc = 0; kc < kNKC; ++kc) {
      const float4 u0 = __ldg(up + kc * 16), u1 = __ldg(up + kc * 16 + 1);
      const float4 v0 = ONE_A ? vr[ONE_A ? kc : 0][0] : __ldg(vp + kc * 16), v1 = ONE_A ? vr[ONE_A ? kc : 0][1] : __ldg(vp + kc * 16 + 1);
      const float h[8] = {u0.x + v0.x, u0.y + v0.y, u0.z + v0.z, u0.w + v0.w, u1.x + v1.x, u1.y + v1.y, u1.z + v1.z, u1.w + v1.w};
      uint4 o;
      o.x = pack_relu_half2(h[0], h[1]);
      o.y = pack_relu_half2(h[2], h[3]);
      o.z = pack_relu_half2(h[4], h[5]);
      o.w = pack_relu_half2(h[6], h[7]);
      if (SAVE) {      // M1, lane-local word j: element (kc, e) -> bit kc*4 + e/2 + 16*(e%2)
        bits |= half2_pos_mask(o.x) & mask_pair_const(kc * 4 + 0);
        bits |= half2_pos_mask(o.y) & mask_pair_const(kc * 4 + 1);
        bits |= half2_pos_mask(o.z) & mask_pair_const(kc * 4 + 2);
        bits |= half2_pos_mask(o.w) & mask_pair_const(kc * 4 + 3);
      }
      *reinterpret_cast<uint4*>(a_tile + kc * kAChunk + sw128_offset(row, j * 8)) = o;
    }
    if (SAVE) p.masks[((size_t)tile * kTileM + row) * 8 + j] = bits;      // masks[0] = M1
  }
}

template <bool SAVE>
__device__ __forceinline__ void generate_h1(const ChainParams& p, int tile, char* a_tile, int q, int lane) {
  if (p.n % kTileM == 0) generate_h1_rows<SAVE, true>(p, tile, a_tile, q, lane);
  else generate_h1_rows<SAVE, false>(p, tile, a_tile, q, lane);
}

// n == 64 fast path (the 8x8 grid): a tile is 2 values of `a` x all 64 values of `c`, so U[c] is read ONCE for both
// rows (c, a0) and (c, a1) and the two Vb rows live in registers: 66 KB of L2 reads per tile instead of 256 KB (the
// chain kernels are bound by L2 -> SM bandwidth).  Warp q takes c in [16q, 16q+16); lane owns columns
// [4*lane, 4*lane+4) and [128 + 4*lane, 128 + 4*lane + 4): fully coalesced 512-byte U reads, conflict-free STS.64.
// M1 sign bits by warp ballot: word (half*4 + e) of a row holds columns half*128 + 4*j + e at bit j (j = 0..31).
template <bool SAVE>
__device__ __forceinline__ void generate_h1_n64(const ChainParams& p, int tile, char* a_tile, int q, int lane) {
  const int b = tile / p.tiles_per_sample;
  const int a0 = (tile % p.tiles_per_sample) * 2;
  const float* Ub = p.U + (size_t)b * 64 * kG + 4 * lane;
  const float* Vb = p.Vb + ((size_t)b * 64 + a0) * kG + 4 * lane;
  float4 v[2][2];
#pragma unroll
  for (int ai = 0; ai < 2; ++ai)
#pragma unroll
    for (int h = 0; h < 2; ++h) v[ai][h] = __ldg(reinterpret_cast<const float4*>(Vb + ai * kG + h * 128));
  char* dst0 = a_tile + (lane >> 4) * kAChunk + ((4 * lane) & 7) * 2;       // + row part + swizzled group
  const int grp = ((4 * lane) & 63) >> 3;
  constexpr int RB = SAVE ? 4 : 8;              // rows per batch: 2*RB loads in flight (the ballot path is register-tight)
#pragma unroll 1
  for (int it = 0; it < 16; it += RB) {
    float4 u[RB][2];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int c = q * 16 + it + i;
#pragma unroll
      for (int h = 0; h < 2; ++h) u[i][h] = __ldg(reinterpret_cast<const float4*>(Ub + (size_t)c * kG + h * 128));
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int c = q * 16 + it + i;
      uint32_t mine = 0;
#pragma unroll
      for (int ai = 0; ai < 2; ++ai) {
        const int row = ai * 64 + c;
        const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((grp ^ (row & 7)) & 7) << 4));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float h0 = u[i][h].x + v[ai][h].x, h1 = u[i][h].y + v[ai][h].y;
          const float h2 = u[i][h].z + v[ai][h].z, h3 = u[i][h].w + v[ai][h].w;
          uint2 o;
          o.x = pack_relu_half2(h0, h1);
          o.y = pack_relu_half2(h2, h3);
          *reinterpret_cast<uint2*>(dst0 + h * 2 * kAChunk + off) = o;
          if (SAVE) {
            const uint32_t w0 = __ballot_sync(0xffffffffu, h0 > 0.f), w1 = __ballot_sync(0xffffffffu, h1 > 0.f);
            const uint32_t w2 = __ballot_sync(0xffffffffu, h2 > 0.f), w3 = __ballot_sync(0xffffffffu, h3 > 0.f);
            const int base = ai * 8 + h * 4;
            if (lane == base + 0) mine = w0;
            if (lane == base + 1) mine = w1;
            if (lane == base + 2) mine = w2;
            if (lane == base + 3) mine = w3;
          }
        }
      }
      if (SAVE && lane < 16) p.masks[((size_t)tile * kTileM + (lane >> 3) * 64 + c) * 8 + (lane & 7)] = mine;   // masks[0] = M1
    }
  }
}

// Backward generation: dZ4[r, :] = S * dxg[b, :] where Z4[r, :] > 0 -> swizzled fp16 A operand.
// Warp q owns rows [32q, 32q+32); lane owns columns [8*lane, 8*lane+8) of every row: the packed fp16 pairs of
// S * dxg for those columns are built once per tile, and a row costs one mask-word load (all 32 issued up front: one
// exposed L2 round trip per tile), four pair-mask ANDs and one conflict-free STS.128.
// (Column sums of the dZ images are taken by the weight-gradient kernel.)
__device__ __forceinline__ void generate_dz4(const ChainParams& p, int tile, char* a_tile, int q, int lane) {
  const int b = tile / p.tiles_per_sample;
  const float S = __ldg(p.scale);
  const float4 d0 = __ldg(reinterpret_cast<const float4*>(p.dxg + (size_t)b * kG + lane * 8));
  const float4 d1 = __ldg(reinterpret_cast<const float4*>(p.dxg + (size_t)b * kG + lane * 8 + 4));
  const uint32_t dp[4] = {pack_half2(d0.x * S, d0.y * S), pack_half2(d0.z * S, d0.w * S), pack_half2(d1.x * S, d1.y * S),
                          pack_half2(d1.z * S, d1.w * S)};
  // cols 8*lane + e of a row live in word lane / 4 at bit (lane % 4) * 4 + e / 2 + 16 * (e % 2)
  const uint32_t* m4 = p.masks + (((size_t)3 * p.num_tiles + tile) * kTileM + q * 32) * 8 + (lane >> 2);
  const int sh = (lane & 3) * 4;
  uint32_t w[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) w[r] = __ldg(m4 + r * 8);
  char* dst = a_tile + (lane >> 3) * kAChunk;
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    const int row = q * 32 + r;
    const uint32_t nib = w[r] >> sh;
    uint4 o;
    o.x = dp[0] & (((nib >> 0) & 0x00010001u) * 0xFFFFu);      // pair (e, e+1): bits e/2 and 16 + e/2 -> 0xFFFF per half
    o.y = dp[1] & (((nib >> 1) & 0x00010001u) * 0xFFFFu);
    o.z = dp[2] & (((nib >> 2) & 0x00010001u) * 0xFFFFu);
    o.w = dp[3] & (((nib >> 3) & 0x00010001u) * 0xFFFFu);
    *reinterpret_cast<uint4*>(dst + sw128_offset(row, (lane & 7) * 8)) = o;
  }
}

// Data-gradient epilogue of one tile row: dZ[row, :] = accumulator[row, :] where the ReLU sign bit is set, as fp16 into
// the swizzled A operand.  LAYOUT 0: word cc, bit mask_pos(e);  1: M1 of generate_h1;  2: M1 of generate_h1_n64
// (column c -> word (c / 128) * 4 + c % 4, bit (c % 128) / 4).
template <int LAYOUT>
__device__ __forceinline__ void dgrad_mask_tile(uint32_t taddr, const uint32_t (&mw)[8], char* a_tile, const uint32_t (&swz)[8]) {
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    uint32_t r[32];
    tmem_ld32(taddr + cc * 32, r);
    tmem_ld_wait();
    float x[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const uint32_t bit = LAYOUT == 0   ? (mw[cc] >> mask_pos(e)) & 1u
                           : LAYOUT == 1 ? (mw[(cc & 1) * 4 + (e >> 3)] >> ((cc >> 1) * 4 + mask_pos(e & 7))) & 1u
                                         : (mw[(cc >> 2) * 4 + (e & 3)] >> ((cc & 3) * 8 + (e >> 2))) & 1u;
      x[e] = bit ? __uint_as_float(r[e]) : 0.f;
    }
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
      uint4 o;
      o.x = pack_half2(x[g4 * 8 + 0], x[g4 * 8 + 1]);
      o.y = pack_half2(x[g4 * 8 + 2], x[g4 * 8 + 3]);
      o.z = pack_half2(x[g4 * 8 + 4], x[g4 * 8 + 5]);
      o.w = pack_half2(x[g4 * 8 + 6], x[g4 * 8 + 7]);
      const int col = cc * 32 + g4 * 8;
      *reinterpret_cast<uint4*>(a_tile + (col >> 6) * kAChunk + swz[(col >> 3) & 7]) = o;
    }
  }
}

// CTA2 = true: the kernel runs as clusters of two CTAs on one TPC (cta_group::2).  Each CTA keeps its own tile
// slots, A buffers, epilogue warps and TMEM accumulators, but ONE tcgen05.mma (M = 256) issued by the leader CTA
// drives both tiles, and each CTA stages only its N-half of every weight chunk: half the L2->smem weight traffic
// and half the shared-memory operand reads per SM -- the 1-CTA form is shared-memory-bandwidth bound.
__device__ long long g_chain_prof[160][16];
#define PROF_T0() long long _t0 = 0; if (prof) _t0 = clock64();
#define PROF_ACC(i) if (prof) { const long long _t1 = clock64(); pacc[(i) < (int)(sizeof(pacc) / sizeof(pacc[0])) ? (i) : 0] += _t1 - _t0; _t0 = _t1; }

// Static MMA job order shared by the weight producer and the MMA issuer (the epilogue warps are driven by barriers only).
// sched 0: round r = [s0 L1, s1 L1, s0 L2, s1 L2, s0 L3, s1 L3]  (strict alternation)
// sched 1: round r = [s0 L1, s1 L3 (previous tile), s0 L2, s0 L3, s1 L1, s1 L2]: every slot gets TWO MMA phases of the
//          other slot between its last layer and the first layer of its next tile, which covers the tile boundary
//          (pair-sum epilogue + operand generation, ~14 k cycles) instead of one (6.8 k), at the price of one
//          back-to-back pair per slot whose gap is a 4 k mid-layer epilogue.  Measured at B=640: 2.25 / 2.87 / 1.77 ms
//          (train forward / backward / eval forward) against 2.23 / 2.91 / 1.61 ms for sched 0, which stays the
//          default (RN_B200_SCHED).
template <typename F>
__device__ __forceinline__ void for_each_chain_job(int my_tiles, int sched, F&& f) {
  const int rounds = (my_tiles + 1) / 2;
  if (sched == 0) {
    for (int r = 0; r < rounds; ++r) {
      const int nslots = (2 * r + 1 < my_tiles) ? 2 : 1;
      for (int layer = 0; layer < kTcLayers; ++layer)
        for (int s = 0; s < nslots; ++s) f(s, layer);
    }
    return;
  }
  bool pending1 = false;                 // slot 1 still owes the last layer of its previous tile
  for (int r = 0; r < rounds; ++r) {
    const bool has1 = 2 * r + 1 < my_tiles;
    f(0, 0);
    if (pending1) f(1, 2);
    f(0, 1);
    f(0, 2);
    if (has1) {
      f(1, 0);
      f(1, 1);
    }
    pending1 = has1;
  }
  if (pending1) f(1, 2);
}

// GEN = true (single-CTA form only): a fourth warpgroup generates the layer-1 operand of each slot's NEXT tile as
// soon as the slot's last MMA has released the A buffer, concurrently with the slot's own last-layer epilogue (which
// only reads TMEM), instead of the epilogue warps doing both back to back on the slot's critical path.
template <int MODE, bool CTA2, bool GEN>
__global__ void __launch_bounds__(GEN ? kGenThreads : kFwdThreads, 1) rn_g_chain_kernel(const ChainParams p) {
  static_assert(!(GEN && CTA2), "generator warpgroup: single-CTA form only");
  constexpr bool SAVE = MODE != kFwdEval;        // training forward and dgrad stream operand images to HBM
  constexpr int NST = CTA2 ? kStages2 : kStages;
  constexpr int STAGE_BYTES = CTA2 ? kWHalf : kWChunk;
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + kSmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  // tiles of this CTA: 1-CTA: blockIdx, +grid, ...   CTA pair c of nc: tiles 2*(c + i*nc) + rank
  const int tile_first = CTA2 ? 2 * (int)(blockIdx.x >> 1) + (int)rank : (int)blockIdx.x;
  const int tile_step = CTA2 ? (int)gridDim.x : (int)gridDim.x;
  const int my_tiles = CTA2 ? (p.num_tiles / 2 - (int)(blockIdx.x >> 1) + (int)(gridDim.x >> 1) - 1) / (int)(gridDim.x >> 1)
                            : tiles_of_cta(p.num_tiles);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&bars->w_full[s]), 1);
      mbar_init(smem_u32(&bars->w_empty[s]), 1);
      mbar_init(smem_u32(&bars->peer_full[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      // pair mode: one arrival per warp, from both CTAs, at the leader.  generator mode: 128 generator + 128 epilogue arrivals
      mbar_init(smem_u32(&bars->a_full[s]), CTA2 ? 8 : GEN ? 256 : 128);
      mbar_init(smem_u32(&bars->acc_full[s]), 1);
      mbar_init(smem_u32(&bars->a_free[s]), 1);
      mbar_init(smem_u32(&bars->h1_done[s]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CTA2) tmem_alloc2(smem_u32(&bars->tmem_base), 512);
    else tmem_alloc(smem_u32(&bars->tmem_base), 512);
  }
  tc_fence_before_sync();
  if (CTA2) cluster_sync_all();        // barriers of both CTAs initialised before any remote arrive
  else __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;
  // job sequence shared by producer and issuer: for round r: for layer 0..2: for slot 0..1 (if its tile exists)
  // generator mode re-balances registers per warpgroup: 56*128 + 176*256 + 96*128 = 64512
  if (warp < 4) {
  if (GEN) reg_dec<56>();
  if (warp == 0) {
    if (lane == 0) {
      // ================= weight producer =================
      uint32_t stage = 0, phase = 0;
      for_each_chain_job(my_tiles, CTA2 ? 0 : p.sched, [&](int, int layer) {
        const int img_layer = MODE == kDgrad ? kTcLayers - 1 - layer : layer;    // dgrad walks W3^T, W2^T, W1^T
        for (int pass = 0; pass < ((p.dbg & 4) ? 0 : (p.dbg & 1) ? 1 : p.passes); ++pass)
          for (int kc = 0; kc < kNKC; ++kc) {
            mbar_wait(smem_u32(&bars->w_empty[stage]), phase ^ 1);
            const uint32_t full = smem_u32(&bars->w_full[stage]);
            mbar_expect_tx(full, STAGE_BYTES);
            // pair mode: this CTA's N-half = rows [128*rank, 128*rank + 128) = the rank-th 16 KB of the chunk image
            const char* src = reinterpret_cast<const char*>(p.wpack) + ((size_t)(img_layer * 2 + pass) * kNKC + kc) * kWChunk +
                              (CTA2 ? rank * kWHalf : 0);
            bulk_g2s(smem_u32(smem + kSmemW + stage * STAGE_BYTES), src, STAGE_BYTES, full);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
      });
    }
  } else if (warp == 1) {
    if (!CTA2 || rank == 0) {
      // ================= MMA issuer (pair mode: leader CTA only) =================
      // The whole warp runs the loop (warp-uniform control flow, descriptors stepped with one 32-bit add in uniform
      // registers); only the tcgen05 instructions are predicated on one elected lane -- see rn_g_fwd3_kernel.
      uint32_t stage = 0, phase = 0;
      uint32_t a_phase[2] = {0, 0};
      const bool prof = (p.dbg & 8) != 0;
      long long pacc[4] = {0, 0, 0, 0};
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const bool leader = elect_one();
      PROF_T0();
      for_each_chain_job(my_tiles, CTA2 ? 0 : p.sched, [&](int s, int) {
            if (CTA2) mbar_wait_cluster(smem_u32(&bars->a_full[s]), a_phase[s]);
            else mbar_wait(smem_u32(&bars->a_full[s]), a_phase[s]);
            a_phase[s] ^= 1;
            PROF_ACC(0);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + s * kG;
            const uint32_t a_lo = smem_desc_lo_sw128(smem_u32(smem + kSmemA + s * kATile), 16);
            uint32_t accumulate = 0;
            const int reps = (p.dbg & 1) ? p.passes : 1;
            for (int pass = 0; pass < ((p.dbg & 1) ? 1 : p.passes); ++pass)
              for (int kc = 0; kc < kNKC; ++kc) {
                if (!(p.dbg & 4)) mbar_wait(smem_u32(&bars->w_full[stage]), phase);
                if (CTA2) mbar_wait_cluster(smem_u32(&bars->peer_full[stage]), phase);
                PROF_ACC(1);
                tc_fence_after_sync();
                const uint32_t b_lo = smem_desc_lo_sw128(smem_u32(smem + kSmemW + stage * STAGE_BYTES), 16);
                const uint32_t a_kc = a_lo + kc * (kAChunk >> 4);
                if (leader) {
                  for (int rep = 0; rep < reps; ++rep)
#pragma unroll
                  for (int k = 0; k < kKC / 16; ++k) {
                    const uint64_t ad = desc_pack(a_kc + 2 * k, kDescHi);
                    const uint64_t bd = desc_pack(b_lo + 2 * k, kDescHi);
                    if (CTA2) mma_f16_ss_2cta(d_tmem, ad, bd, kIdescFwd2, accumulate | (uint32_t)(k > 0 || rep > 0));
                    else mma_f16_ss(d_tmem, ad, bd, kIdescFwd, accumulate | (uint32_t)(k > 0 || rep > 0));
                  }
                  // ring stage free (in both CTAs) once these MMAs retire
                  if (CTA2) mma_commit_2cta(smem_u32(&bars->w_empty[stage]), 3);
                  else if (!(p.dbg & 4)) mma_commit(smem_u32(&bars->w_empty[stage]));
                }
                accumulate = 1;
                if (!(p.dbg & 4) && ++stage == NST) { stage = 0; phase ^= 1; }
              }
            // accumulator of (slot, layer) complete (in both CTAs)
            if (leader) {
              if (CTA2) mma_commit_2cta(smem_u32(&bars->acc_full[s]), 3);
              else mma_commit(smem_u32(&bars->acc_full[s]));
            }
            __syncwarp();
            PROF_ACC(2);
      });
      if (prof && lane == 0) for (int i = 0; i < 3; ++i) g_chain_prof[blockIdx.x][8 + i] = pacc[i];
    } else if (CTA2 && lane == 0) {
      // ================= peer CTA relay: tell the leader when this CTA's half of each chunk has landed =================
      uint32_t stage = 0, phase = 0;
      for (int r = 0; 2 * r < my_tiles; ++r) {
        const int nslots = (2 * r + 1 < my_tiles) ? 2 : 1;
        for (int job = 0; job < kTcLayers * nslots * p.passes * kNKC; ++job) {
          mbar_wait(smem_u32(&bars->w_full[stage]), phase);
          mbar_arrive_cluster(mapa_u32(smem_u32(&bars->peer_full[stage]), 0));
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  }
  } else if (GEN && warp >= 12) {
    // ================= generator warpgroup =================
    reg_dec<96>();
    const int q = warp & 3, gtid = threadIdx.x - 384;
    uint32_t free_phase[2] = {0, 0};
    const bool prof = (p.dbg & 8) != 0 && gtid == 0;
    long long pacc[2] = {0, 0};
    PROF_T0();
    for (int i = 0; i < my_tiles; ++i) {
      const int s = i & 1;
      const int tile = tile_first + i * tile_step;
      char* a_tile = smem + kSmemA + s * kATile;
      mbar_wait(smem_u32(&bars->a_free[s]), free_phase[s] ^ 1);       // passes at once for the slot's first tile
      free_phase[s] ^= 1;
      PROF_ACC(0);
      if (MODE == kDgrad) generate_dz4(p, tile, a_tile, q, lane);
      else if (p.n == 64) generate_h1_n64<MODE == kFwdTrain>(p, tile, a_tile, q, lane);
      else generate_h1<MODE == kFwdTrain>(p, tile, a_tile, q, lane);
      fence_proxy_async_smem();
      if (SAVE) {
        named_bar_sync(3, 128);                   // whole image written and fenced by every generator thread
        if (gtid == 0) {
          char* dst = MODE == kFwdTrain ? reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 3 + 0) * kATile
                                        : reinterpret_cast<char*>(p.dZ) + ((size_t)tile * 4 + 3) * kATile;
          bulk_s2g(dst, smem_u32(a_tile), kATile);
          bulk_commit();
        }
      }
      mbar_arrive(smem_u32(&bars->a_full[s]));
      PROF_ACC(1);
      if (SAVE && gtid == 0) {
        bulk_wait_read0();
        mbar_arrive(smem_u32(&bars->h1_done[s]));
      }
    }
    if (SAVE && gtid == 0) bulk_wait0();
    if (prof) {
      g_chain_prof[blockIdx.x][6] = pacc[0];
      g_chain_prof[blockIdx.x][7] = pacc[1];
    }
  } else {
    // ================= generation / epilogue warpgroups =================
    if (GEN) reg_inc<176>();
    const int s = (warp - 4) >> 2;             // tile slot
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int wg_tid = threadIdx.x - 128 - s * 128;
    char* a_tile = smem + kSmemA + s * kATile;
    // pair mode: the A-operand-ready barrier lives in the leader CTA (it issues the MMAs for both tiles)
    const uint32_t a_full = CTA2 ? mapa_u32(smem_u32(&bars->a_full[s]), 0) : smem_u32(&bars->a_full[s]);
    const uint32_t acc_full = smem_u32(&bars->acc_full[s]);
    const uint32_t bar_id = 1 + s;
    uint32_t acc_phase = 0;
    uint32_t swz[8];                 // byte offset of this row's k-th 16-byte group inside a swizzled 128x64 chunk
#pragma unroll
    for (int k = 0; k < 8; ++k) swz[k] = sw128_offset(row, k * 8);

    // "this slot's A operand is ready": every thread has written + fenced its part.  Single-CTA: 128 local arrivals.
    // Pair mode: one (remote) arrival per warp at the leader's barrier keeps the cluster traffic small.
    auto arrive_a_full = [&]() {
      if (CTA2) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_full);
      } else if (GEN) {
        mbar_arrive_n(a_full, 2);              // the barrier expects 256 arrivals (generator mode)
      } else {
        mbar_arrive(a_full);
      }
    };
    uint32_t h1_phase = 0;
    // stream the freshly written operand image of this slot to HBM (one elected thread, bulk async store)
    auto store_image = [&](char* dst) {
      named_bar_sync(bar_id, 128);              // whole image written and fenced by every thread
      if (wg_tid == 0) {
        bulk_s2g(dst, smem_u32(a_tile), kATile);
        bulk_commit();
      }
    };
    // before overwriting the A buffer: the previous image store must have finished READING it
    auto wait_image_read = [&]() {
      if (wg_tid == 0) bulk_wait_read0();
      named_bar_sync(bar_id, 128);
    };

    const bool prof = (p.dbg & 8) != 0 && threadIdx.x == 128;
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    PROF_T0();
    for (int i = s; i < my_tiles; i += 2) {
      const int tile = tile_first + i * tile_step;
      const int b = tile / p.tiles_per_sample;
      if (prof) _t0 = clock64();
      if (GEN) {
        // this slot's accumulator has been drained (or never used): the epilogue half of "operand + accumulator ready"
        tc_fence_before_sync();
        mbar_arrive(a_full);
      } else {
        if (SAVE) wait_image_read();
        if (MODE == kDgrad) generate_dz4(p, tile, a_tile, q, lane);
        else if (p.n == 64) generate_h1_n64<MODE == kFwdTrain>(p, tile, a_tile, q, lane);
        else generate_h1<MODE == kFwdTrain>(p, tile, a_tile, q, lane);
        fence_proxy_async_smem();
        if (MODE == kFwdTrain && !p.skip_h1_image) store_image(reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 3 + 0) * kATile);
        if (MODE == kDgrad && !p.skip_dz4_image) store_image(reinterpret_cast<char*>(p.dZ) + ((size_t)tile * 4 + 3) * kATile);
        arrive_a_full();
      }
      PROF_ACC(0);

      for (int layer = 0; layer < kTcLayers; ++layer) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * kG;
        if (MODE != kDgrad) {
          // ---------------- forward epilogue ----------------
          const float* bias = p.bias[layer] + (size_t)b * p.bias_stride[layer];
          uint32_t* mrow = SAVE ? p.masks + (((size_t)(layer + 1) * p.num_tiles + tile) * kTileM + row) * 8 : nullptr;
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          PROF_ACC(1);
          tc_fence_after_sync();
          if (layer < kTcLayers - 1) {
            if (GEN && SAVE && layer == 0) {       // the generator's image store of H1 must have finished reading A
              mbar_wait(smem_u32(&bars->h1_done[s]), h1_phase);
              h1_phase ^= 1;
            } else if (SAVE) {
              wait_image_read();
            }
            uint32_t mw[8];
            if (p.dbg & 2) {
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) mw[cc] = 0;
            } else
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              uint32_t r[32];
              tmem_ld32(taddr + cc * 32, r);
              tmem_ld_wait();
              uint32_t bits = 0;
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                float v[8];
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + g4 * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + g4 * 8 + 4));
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g4 * 8 + e]) + bb[e];
                uint4 o;
                o.x = pack_relu_half2(v[0], v[1]);
                o.y = pack_relu_half2(v[2], v[3]);
                o.z = pack_relu_half2(v[4], v[5]);
                o.w = pack_relu_half2(v[6], v[7]);
                if (SAVE) {
                  bits |= half2_pos_mask(o.x) & mask_pair_const(g4 * 4 + 0);
                  bits |= half2_pos_mask(o.y) & mask_pair_const(g4 * 4 + 1);
                  bits |= half2_pos_mask(o.z) & mask_pair_const(g4 * 4 + 2);
                  bits |= half2_pos_mask(o.w) & mask_pair_const(g4 * 4 + 3);
                }
                const int col = cc * 32 + g4 * 8;
                *reinterpret_cast<uint4*>(a_tile + (col >> 6) * kAChunk + swz[(col >> 3) & 7]) = o;
              }
              mw[cc] = bits;
            }
            if (SAVE) {
              *reinterpret_cast<uint4*>(mrow) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
              *reinterpret_cast<uint4*>(mrow + 4) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
            }
            fence_proxy_async_smem();
            if (SAVE) store_image(reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 3 + layer + 1) * kATile);
            tc_fence_before_sync();
            arrive_a_full();
            PROF_ACC(2);
          } else {
            // last layer: ReLU + pair-sum.  Column sums over this warp's 32 rows by shuffle transpose-reduce.
            if (GEN && wg_tid == 0) {              // the last MMA has read A (and so has the H3 image store): hand it to the generator
              if (SAVE) bulk_wait_read0();
              mbar_arrive(smem_u32(&bars->a_free[s]));
            }
            float* part = p.xg_part + ((size_t)tile * 4 + q) * kG;
            uint32_t mw[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              uint32_t r[32];
              tmem_ld32(taddr + cc * 32, r);
              tmem_ld_wait();
              float x[32];
              uint32_t bits = 0;
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cc * 32 + e4 * 4));
                x[e4 * 4 + 0] = fmaxf(__uint_as_float(r[e4 * 4 + 0]) + bv.x, 0.f);
                x[e4 * 4 + 1] = fmaxf(__uint_as_float(r[e4 * 4 + 1]) + bv.y, 0.f);
                x[e4 * 4 + 2] = fmaxf(__uint_as_float(r[e4 * 4 + 2]) + bv.z, 0.f);
                x[e4 * 4 + 3] = fmaxf(__uint_as_float(r[e4 * 4 + 3]) + bv.w, 0.f);
              }
              if (SAVE) {
#pragma unroll
                for (int i = 0; i < 16; ++i) bits |= half2_pos_mask(pack_half2(x[2 * i], x[2 * i + 1])) & mask_pair_const(i);
              }
              mw[cc] = bits;
              part[cc * 32 + lane] = warp_transpose_sum(x, lane);     // lane L holds the sum of column cc*32 + L
            }
            if (SAVE) {
              *reinterpret_cast<uint4*>(mrow) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
              *reinterpret_cast<uint4*>(mrow + 4) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
            }
            tc_fence_before_sync();
            PROF_ACC(3);
          }
        } else {
          // ---------------- data-gradient epilogue: dZ_l = dH_l .* (Z_l > 0), l = 3 - layer ----------------
          const int l = kTcLayers - layer;                // 3, 2, 1
          const uint32_t* mrow = p.masks + (((size_t)(l - 1) * p.num_tiles + tile) * kTileM + row) * 8;
          const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(mrow));
          const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(mrow) + 1);
          const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
          mbar_wait(acc_full, acc_phase);
          acc_phase ^= 1;
          PROF_ACC(1);
          tc_fence_after_sync();
          char* img = reinterpret_cast<char*>(p.dZ) + ((size_t)tile * 4 + (l - 1)) * kATile;
          if (GEN && l == 1) {
            // generator mode: dZ1 goes straight from registers to its HBM image (same swizzled offsets), so the A buffer is
            // free for the next tile's dZ4 as soon as the dZ2 image store has finished reading it
            if (wg_tid == 0) {
              bulk_wait_read0();
              mbar_arrive(smem_u32(&bars->a_free[s]));
            }
            if (p.m1_layout0) dgrad_mask_tile<0>(taddr, mw, img, swz);
            else if (p.n == 64) dgrad_mask_tile<2>(taddr, mw, img, swz);
            else dgrad_mask_tile<1>(taddr, mw, img, swz);
            tc_fence_before_sync();
          } else {
            if (GEN && layer == 0) {                 // the generator's dZ4 image store must have finished reading A
              mbar_wait(smem_u32(&bars->h1_done[s]), h1_phase);
              h1_phase ^= 1;
            } else {
              wait_image_read();
            }
            if (!(p.dbg & 2)) {
              // Z2..Z4 masks: word cc, bit mask_pos(e).  Z1 mask: lane-local layout of generate_h1 (column c -> word
              // (c % 64) / 8, bit (c / 64) * 4 + (c % 8) / 2 + 16 * (c % 2)) or the ballot layout of generate_h1_n64.
              if (l != 1 || p.m1_layout0) dgrad_mask_tile<0>(taddr, mw, a_tile, swz);
              else if (p.n == 64) dgrad_mask_tile<2>(taddr, mw, a_tile, swz);
              else dgrad_mask_tile<1>(taddr, mw, a_tile, swz);
            }
            fence_proxy_async_smem();
            store_image(img);
            tc_fence_before_sync();
            if (layer < kTcLayers - 1) arrive_a_full();
          }
          PROF_ACC(2);
        }
      }
    }
    if (SAVE && wg_tid == 0) bulk_wait0();
    if (prof) {
      for (int i = 0; i < 4; ++i) g_chain_prof[blockIdx.x][i] = pacc[i];
      g_chain_prof[blockIdx.x][4] = clock64() - t_begin;
      g_chain_prof[blockIdx.x][5] = my_tiles;
    }
  }

  tc_fence_before_sync();
  if (CTA2) cluster_sync_all();        // the peer must not exit (or free TMEM) while the pair's MMAs may touch it
  else __syncthreads();
  if (warp == 2) {
    if (CTA2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------
// 3-pass training forward (RN_PRECISION_PARITY): fp32-level pre-activations on the tensor cores.
//
// Why: with fp16 activations (the 2-pass kernel above) a pre-activation carries ~2^-12 relative error, so ~1e-4 of the
// ReLU masks differ from an fp32 evaluation; that mask noise, not operand rounding in the backward, is what the
// gradient error of the round-1 kernels consisted of (2e-3 .. 1e-2 on dq / dW with trained weights, independent of
// the batch size -- profiles/r02_grad_error_vs_batch.json).  Here every layer computes
//     Z = A_hi W_hi + A_lo W_hi + A_hi W_lo,   A = A_hi + A_lo, W = W_hi + W_lo  (fp16 splits, fp32 accumulate)
// which leaves ~2^-21: the masks, and with them the gradients, agree with fp32 to the fp32 conditioning floor.
//
// Structure (persistent, one CTA per SM, ONE 128-row tile in flight, 512 threads):
//   A_hi of the current / next layer: two 64 KB swizzled shared-memory buffers;  A_lo: two 128-column tensor-memory
//   buffers (packed fp16 pairs, lane = row), read by tcgen05.mma as a TMEM A operand -- it costs no shared-memory
//   bandwidth and no shared-memory capacity;  accumulators: two N = 128 halves (tensor-memory columns 0..255).
//   warp 0      weight producer: 32 KB stages = (layer, N-half, K-chunk): [W_hi 128x64 | W_lo 128x64], 3-stage ring
//   warp 1      MMA issuer: per stage 4 x (SS A_hi W_hi, TS A_lo W_hi) + 4 x SS A_hi W_lo, M128 x N128 x K16
//   warp 2      TMEM allocator
//   warps 4-11  epilogue of one accumulator half (all 8 warps: 4 lane quarters x 2 column halves of 64): +bias, ReLU,
//               split into hi (-> shared memory, next layer's A_hi) and lo (-> tcgen05.st, next layer's A_lo), sign bits;
//               it runs under the MMAs of the other half, and the next layer starts on K-chunks 0,1 as soon as half 0
//               is written (epi_ready[0]) while half 1 is still in the epilogue.  Last layer: ReLU + pair-sum.
//   warps 12-15 operand generator: H1 = relu(U[c] + V'[a]) of the NEXT tile, row per thread (U read through its
//               column-group-major copy U4: 512 contiguous bytes per warp load), under the current tile's last layer.
// MMA work per tile-layer: 96 x 64 = 6144 cycles; weights: 256 KB per tile-layer as before, i.e. 42 B/cycle -- below
// the ~54 B/cycle one SM gets out of 32 KB bulk copies (tests/micro/ts_mma_probe.cu), so this kernel is tensor-bound
// where the 2-pass kernel is copy-bound.
// ------------------------------------------------------------------------------------------------
constexpr int k3Threads = 640;            // 4 control warps, 8 epilogue warps, 8 generator warps
constexpr int k3WSub = 16384;              // one weight sub-chunk: (layer, N-half, K-chunk, hi | lo) = 128 N rows x 64 K
constexpr int k3SubPerStage = 2;
constexpr int k3WStage = k3SubPerStage * k3WSub;      // 32 KB per bulk copy, 3 in the ring (2 x 48 KB measured slower: prefetch depth)
constexpr int k3Stages = 3;
constexpr int k3SmemW = 2 * kATile;
constexpr int k3SmemBar = k3SmemW + k3Stages * k3WStage;
constexpr int k3SmemLaunch = k3SmemBar + 256 + 1024;
constexpr uint32_t kIdescHalf = idesc_f16(kTileM, 128, 0, 0);
constexpr uint32_t k3AloCol = 256;            // tensor-memory columns [256, 512): two A_lo buffers of 128 columns

struct Bars3 {
  uint64_t w_full[k3Stages];
  uint64_t w_empty[k3Stages];
  uint64_t gen_ready;       // 256 generator arrivals: the tile's first operand (A_hi + A_lo) is written
  uint64_t gen_go;          // the buffer of the NEXT tile's first operand is free (MMA commit [+ its image store has been read])
  uint64_t epi_ready[2];    // 256 epilogue arrivals: K-half kh of the next layer's operand is written
  uint64_t acc_full[2];     // MMA commit: accumulator half h is complete
  uint64_t acc_free[2];     // 256 arrivals: the last-layer epilogue has drained accumulator half h
  uint64_t img_ready;       // 256 epilogue arrivals: a complete operand (H2 / H3) sits in shared memory, ready for its image store
  uint64_t h3_done;         // the H3 image store has finished reading its buffer (the next tile's H2 overwrites it)
  uint32_t tmem_base;
};

// image index (((layer * 2 + half) * 4 + kc) * 2 + pass) * 16 KB: 128 N rows x 64 K, SWIZZLE_128B
__global__ void pack_weights3_kernel(PackArgs args, __half* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kTcLayers * kG * (kG / 8)) return;
  const int layer = idx / (kG * (kG / 8));
  const int rem = idx % (kG * (kG / 8));
  const int nrow = rem / (kG / 8);
  const int kg = rem % (kG / 8);
  const float* w = args.w[layer] + (size_t)nrow * args.ld[layer] + kg * 8;
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float v0 = w[2 * e], v1 = w[2 * e + 1];
    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
    __half2 hh = __halves2half2(h0, h1);
    __half2 ll = __floats2half2_rn(v0 - __half2float(h0), v1 - __half2float(h1));
    hi[e] = *reinterpret_cast<uint32_t*>(&hh);
    lo[e] = *reinterpret_cast<uint32_t*>(&ll);
  }
  char* base = reinterpret_cast<char*>(out) + ((size_t)((layer * 2 + (nrow >> 7)) * kNKC + kg / 8) * 2) * 16384 +
               sw128_offset(nrow & 127, (kg % 8) * 8);
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + 16384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// relu(z) as an fp16 pair (hi) and the fp16 pair of what the rounding dropped (lo)
__device__ __forceinline__ void split_relu_half2(float z0, float z1, uint32_t& hi, uint32_t& lo) {
  const float r0 = fmaxf(z0, 0.f), r1 = fmaxf(z1, 0.f);
  const __half2 h = __floats2half2_rn(r0, r1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

#define PROF3_T0() uint32_t _t0 = 0; if (prof) _t0 = (uint32_t)clock();
#define PROF3_ACC(i) if (prof) { const uint32_t _t1 = (uint32_t)clock(); pacc[(i) < (int)(sizeof(pacc) / sizeof(pacc[0])) ? (i) : 0] += _t1 - _t0; _t0 = _t1; }
template <bool SAVE, bool PROF>
__global__ void __launch_bounds__(k3Threads, 1) rn_g_fwd3_kernel(const ChainParams p) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars3* bars = reinterpret_cast<Bars3*>(smem + k3SmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = tiles_of_cta(p.num_tiles);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < k3Stages; ++s) {
      mbar_init(smem_u32(&bars->w_full[s]), 1);
      mbar_init(smem_u32(&bars->w_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->gen_ready), 256);
    mbar_init(smem_u32(&bars->img_ready), 256);
    mbar_init(smem_u32(&bars->h3_done), 1);
    mbar_init(smem_u32(&bars->gen_go), SAVE ? 2 : 1);
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(&bars->epi_ready[h]), 256);
      mbar_init(smem_u32(&bars->acc_full[h]), 1);
      mbar_init(smem_u32(&bars->acc_free[h]), 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp < 4) {
  reg_dec<48>();          // (the register budgets of the three roles must not share code after a join: ptxas assumes the minimum)
  if (warp == 3) {
    if (SAVE && lane == 0) {
      // ================= image storer: streams H2 / H3 (the A_hi operands) to HBM for the weight-gradient kernel =================
      uint32_t ph = 0, j = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        for (int layer = 0; layer < kTcLayers; ++layer, ++j) {
          if (layer == kTcLayers - 1) continue;
          mbar_wait(smem_u32(&bars->img_ready), ph);
          ph ^= 1;
          if (!(p.dbg & 64))      // (timing ablation: no image stores)
            bulk_s2g(reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 3 + layer + 1) * kATile, smem_u32(smem + ((j + 1) & 1) * kATile), kATile);
          bulk_commit();
          bulk_wait_read0();
          mbar_arrive(smem_u32(layer == 0 ? &bars->gen_go : &bars->h3_done));
        }
      }
      bulk_wait0();
    }
  } else if (warp == 0) {
    if (lane == 0) {
      // ================= weight producer =================
      // the weight image is a stream of 48 sub-chunks per tile in consumption order (layer, half, kc, hi | lo); one copy = k3SubPerStage of them
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < my_tiles; ++i)
        for (int v = 0; v < kTcLayers * 2 * kNKC * 2 / k3SubPerStage; ++v) {
          mbar_wait(smem_u32(&bars->w_empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bars->w_full[stage]);
          mbar_expect_tx(full, k3WStage);
          bulk_g2s(smem_u32(smem + k3SmemW + stage * k3WStage), reinterpret_cast<const char*>(p.wpack) + (size_t)v * k3WStage,
                   k3WStage, full);
          if (++stage == k3Stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    {
      // ================= MMA issuer =================
      // The WHOLE warp runs this loop (warp-uniform control flow, descriptors in uniform registers); only the tcgen05
      // instructions are predicated on one elected lane.  A lane-0-only branch makes the compiler wrap every MMA in a
      // divergence loop and pass the descriptors through vector registers: measured, that instruction stream -- not the
      // tensor pipe, not the weight copies -- set the pace of the kernel.
      uint32_t stage = 0, phase = 0, gen_phase = 0, sub = 0;      // sub: sub-chunk within the current 48 KB stage
      uint32_t epi_phase = 0, free_phase = 0;      // bit h = phase of barrier [h] (no dynamically indexed arrays)
      uint32_t j = 0;                                  // operand counter: operand j lives in buffer j & 1
      constexpr bool prof = PROF;
      uint32_t pacc[PROF ? 8 : 1] = {0};
      const uint32_t t_begin = PROF ? (uint32_t)clock() : 0u;
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const bool leader = elect_one();
      PROF3_T0();
      for (int i = 0; i < my_tiles; ++i)
        for (int layer = 0; layer < kTcLayers; ++layer, ++j) {
          const uint32_t a_lo = smem_desc_lo_sw128(smem_u32(smem + (j & 1) * kATile), 16);
          const uint32_t alo = tmem_base + k3AloCol + (j & 1) * 128;
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_tmem = tmem_base + h * 128;
            if (layer == 0) {
              if (h == 0) {
                PROF3_ACC(5);
                mbar_wait(smem_u32(&bars->gen_ready), gen_phase);
                gen_phase ^= 1;
                PROF3_ACC(0);
              }
              if (i > 0) {                              // the previous tile's pair-sum epilogue has drained this half
                PROF3_ACC(5);
                mbar_wait(smem_u32(&bars->acc_free[h]), (free_phase >> h) & 1u);
                free_phase ^= 1u << h;
                PROF3_ACC(1);
              }
            }
            uint32_t accumulate = 0;
#pragma unroll 1
            for (int kc = 0; kc < kNKC; ++kc) {
              if (layer > 0 && h == 0 && (kc & 1) == 0) {      // K-half kc / 2 of this layer's operand is written
                PROF3_ACC(5);
                mbar_wait(smem_u32(&bars->epi_ready[kc >> 1]), (epi_phase >> (kc >> 1)) & 1u);
                epi_phase ^= 1u << (kc >> 1);
                PROF3_ACC(2 + (kc >> 1));
              }
              const uint32_t a_kc = a_lo + kc * (kAChunk >> 4);
              const uint32_t alo_kc = alo + kc * 32;
#pragma unroll
              for (int pass = 0; pass < 2; ++pass) {            // sub-chunk: W_hi then W_lo of (layer, h, kc)
                if (sub == 0) {
                  PROF3_ACC(5);
                  mbar_wait(smem_u32(&bars->w_full[stage]), phase);
                  PROF3_ACC(4);
                }
                tc_fence_after_sync();
                const uint32_t b_lo = smem_desc_lo_sw128(smem_u32(smem + k3SmemW + stage * k3WStage) + sub * k3WSub, 16);
                if (leader) {
                  if (pass == 0) {
#pragma unroll
                    for (int k = 0; k < kKC / 16; ++k) {
                      const uint64_t bd = desc_pack(b_lo + 2 * k, kDescHi);
                      mma_f16_ss(d_tmem, desc_pack(a_kc + 2 * k, kDescHi), bd, kIdescHalf, accumulate | (uint32_t)(k > 0));
                      if (!(p.dbg & 32)) mma_f16_ts(d_tmem, alo_kc + k * 8, bd, kIdescHalf, 1);
                    }
                  } else if (!(p.dbg & 16)) {
#pragma unroll
                    for (int k = 0; k < kKC / 16; ++k)
                      mma_f16_ss(d_tmem, desc_pack(a_kc + 2 * k, kDescHi), desc_pack(b_lo + 2 * k, kDescHi), kIdescHalf, 1);
                  }
                }
                accumulate = 1;
                if (++sub == k3SubPerStage) {
                  sub = 0;
                  if (leader) mma_commit(smem_u32(&bars->w_empty[stage]));
                  if (++stage == k3Stages) { stage = 0; phase ^= 1; }
                }
              }
            }
            if (leader) {
              mma_commit(smem_u32(&bars->acc_full[h]));
              // the second layer's operand buffer is where the generator writes the next tile's first operand
              if (layer == 1 && h == 1) mma_commit(smem_u32(&bars->gen_go));
            }
            __syncwarp();
          }
        }
      if (prof && lane == 0) {
        for (int c = 0; c < 5; ++c) g_chain_prof[blockIdx.x][c] = pacc[c];
        g_chain_prof[blockIdx.x][5] = (uint32_t)clock() - t_begin;
        g_chain_prof[blockIdx.x][13] = my_tiles;
      }
    }
  }
  } else if (warp >= 12) {
    // ================= operand generator: 8 warps, thread = (row, column half) =================
    // register pool of the CTA = 640 x 96 at launch: 128 x 48 (control) + 256 x 88 (generator) + 256 x 128 (epilogue) = 61440
    // (generator keeps the 96 registers of the launch)
    const int q = warp & 3, ch = (warp - 12) >> 2, gt = threadIdx.x - 384;
    const int row = q * 32 + lane;
    uint32_t swz[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) swz[k] = sw128_offset(row, k * 8);
    uint32_t go_phase = 0;
    const bool prof = PROF && gt == 0;
    uint32_t pacc[PROF ? 2 : 1] = {0};
    PROF3_T0();
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / p.tiles_per_sample;
      const int pr = (tile % p.tiles_per_sample) * kTileM + row;
      const int a = pr / p.n, c = pr - a * p.n;
      const uint32_t buf = (3u * (uint32_t)i) & 1u;
      const float4* up = reinterpret_cast<const float4*>(p.U4) + ((size_t)b * (kG / 4) + ch * 32) * p.n + c;      // + g4 * n
      const float4* vp = reinterpret_cast<const float4*>(p.Vb + ((size_t)b * p.n + a) * kG + ch * 128);
      // U is streamed through a 3-deep register ring of 16-column blocks (two blocks = 128 bytes per thread in flight:
      // the L2 round trip, not issue slots, paces this warp role); V' rows are warp-uniform L1 hits, loaded at use.
      float4 u[3][4];
#pragma unroll
      for (int pb = 0; pb < 2; ++pb)
#pragma unroll
        for (int g = 0; g < 4; ++g) u[pb][g] = __ldg(up + (size_t)(pb * 4 + g) * p.n);
      if (i > 0) {
        mbar_wait(smem_u32(&bars->gen_go), go_phase);
        go_phase ^= 1;
        tc_fence_after_sync();
      }
      PROF3_ACC(0);
      char* a_dst = smem + buf * kATile + (row >> 3) * 1024 + (row & 7) * 128;
      const int r7 = row & 7;
      const uint32_t alo = tmem_base + ((uint32_t)(q * 32) << 16) + k3AloCol + buf * 128 + ch * 64;
      uint32_t* mrow = p.masks + ((size_t)tile * kTileM + row) * 8 + ch * 4;       // masks[0] = M1, same layout as M2..M4
      uint32_t lo[16];
      uint32_t bits = 0;
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {                 // 16 columns per block
        if (blk + 2 < 8) {
#pragma unroll
          for (int g = 0; g < 4; ++g) u[(blk + 2) % 3][g] = __ldg(up + (size_t)((blk + 2) * 4 + g) * p.n);
        }
        float4 v[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) v[g] = __ldg(vp + blk * 4 + g);
        const int h16 = blk & 1;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const float4 u0 = u[blk % 3][2 * g8], u1 = u[blk % 3][2 * g8 + 1], v0 = v[2 * g8], v1 = v[2 * g8 + 1];
          uint4 o;
          uint32_t* l4 = lo + h16 * 8 + g8 * 4;
          split_relu_half2(u0.x + v0.x, u0.y + v0.y, o.x, l4[0]);
          split_relu_half2(u0.z + v0.z, u0.w + v0.w, o.y, l4[1]);
          split_relu_half2(u1.x + v1.x, u1.y + v1.y, o.z, l4[2]);
          split_relu_half2(u1.z + v1.z, u1.w + v1.w, o.w, l4[3]);
          if (SAVE) {
            bits |= half2_pos_mask(o.x) & mask_pair_const(h16 * 8 + g8 * 4 + 0);
            bits |= half2_pos_mask(o.y) & mask_pair_const(h16 * 8 + g8 * 4 + 1);
            bits |= half2_pos_mask(o.z) & mask_pair_const(h16 * 8 + g8 * 4 + 2);
            bits |= half2_pos_mask(o.w) & mask_pair_const(h16 * 8 + g8 * 4 + 3);
          }
          const int col = ch * 128 + blk * 16 + g8 * 8;
          *reinterpret_cast<uint4*>(a_dst + (col >> 6) * kAChunk + ((((col >> 3) & 7) ^ r7) << 4)) = o;
        }
        if (h16) {                                        // 32 columns = 16 tensor-memory columns per store
          if (SAVE) mrow[blk >> 1] = bits;
          bits = 0;
          tmem_st16(alo + (blk >> 1) * 16, lo);
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      if (SAVE && !p.skip_h1_image) {
        named_bar_sync(3, 256);
        if (gt == 0) {
          bulk_s2g(reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 3 + 0) * kATile, smem_u32(smem + buf * kATile), kATile);
          bulk_commit();
          bulk_wait_read0();
        }
      }
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&bars->gen_ready));
      PROF3_ACC(1);
    }
    if (prof) {
      g_chain_prof[blockIdx.x][11] = pacc[0];
      g_chain_prof[blockIdx.x][12] = pacc[1];
    }
    if (SAVE && gt == 0) bulk_wait0();
  } else {
    // ================= epilogue warps =================
    reg_inc<120>();
    const int q = warp & 3, ch = (warp - 4) >> 2, et = threadIdx.x - 128;
    const int row = q * 32 + lane;
    uint32_t swz[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) swz[k] = sw128_offset(row, k * 8);
    uint32_t acc_phase = 0, h3_phase = 0;
    uint32_t j = 0;
    const bool prof = PROF && et == 0;
    uint32_t pacc[PROF ? 8 : 1] = {0};
    PROF3_T0();
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / p.tiles_per_sample;
      for (int layer = 0; layer < kTcLayers; ++layer, ++j) {
        const uint32_t nbuf = (j + 1) & 1;
        const float* bias = p.bias[layer] + (size_t)b * p.bias_stride[layer];
        for (int h = 0; h < 2; ++h) {
          const int colbase = h * 128 + ch * 64;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + colbase;
          uint32_t* mrow = SAVE ? p.masks + (((size_t)(layer + 1) * p.num_tiles + tile) * kTileM + row) * 8 + (colbase >> 5) : nullptr;
          mbar_wait(smem_u32(&bars->acc_full[h]), (acc_phase >> h) & 1u);
          acc_phase ^= 1u << h;
          PROF3_ACC(h * 2);
          tc_fence_after_sync();
          if (layer < kTcLayers - 1) {
            if (SAVE && layer == 0 && h == 0 && i > 0) {      // the previous tile's H3 image store has finished reading this buffer
              mbar_wait(smem_u32(&bars->h3_done), h3_phase);
              h3_phase ^= 1;
            }
            char* a_dst = smem + nbuf * kATile;
            uint32_t lo[32], mw[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t r[32];
              tmem_ld32(taddr + cc * 32, r);
              tmem_ld_wait();
              uint32_t bits = 0;
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                const int col = colbase + cc * 32 + g4 * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
                uint4 o;
                uint32_t* l4 = lo + cc * 16 + g4 * 4;
                split_relu_half2(__uint_as_float(r[g4 * 8 + 0]) + b0.x, __uint_as_float(r[g4 * 8 + 1]) + b0.y, o.x, l4[0]);
                split_relu_half2(__uint_as_float(r[g4 * 8 + 2]) + b0.z, __uint_as_float(r[g4 * 8 + 3]) + b0.w, o.y, l4[1]);
                split_relu_half2(__uint_as_float(r[g4 * 8 + 4]) + b1.x, __uint_as_float(r[g4 * 8 + 5]) + b1.y, o.z, l4[2]);
                split_relu_half2(__uint_as_float(r[g4 * 8 + 6]) + b1.z, __uint_as_float(r[g4 * 8 + 7]) + b1.w, o.w, l4[3]);
                if (SAVE) {
                  bits |= half2_pos_mask(o.x) & mask_pair_const(g4 * 4 + 0);
                  bits |= half2_pos_mask(o.y) & mask_pair_const(g4 * 4 + 1);
                  bits |= half2_pos_mask(o.z) & mask_pair_const(g4 * 4 + 2);
                  bits |= half2_pos_mask(o.w) & mask_pair_const(g4 * 4 + 3);
                }
                *reinterpret_cast<uint4*>(a_dst + (col >> 6) * kAChunk + swz[(col >> 3) & 7]) = o;
              }
              mw[cc] = bits;
            }
            tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + k3AloCol + nbuf * 128 + (colbase >> 1), lo);
            tmem_st_wait();
            if (SAVE) *reinterpret_cast<uint2*>(mrow) = make_uint2(mw[0], mw[1]);
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(smem_u32(&bars->epi_ready[h]));
            if (SAVE && h == 1) mbar_arrive(smem_u32(&bars->img_ready));      // both halves written and fenced by this thread
            PROF3_ACC(h * 2 + 1);
          } else {
            // last layer: ReLU + pair-sum (column sums over this warp's 32 rows)
            float* part = p.xg_part + ((size_t)tile * 4 + q) * kG;
            uint32_t mw[2];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t r[32];
              tmem_ld32(taddr + cc * 32, r);
              tmem_ld_wait();
              float x[32];
              uint32_t bits = 0;
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + colbase + cc * 32 + e4 * 4));
                x[e4 * 4 + 0] = fmaxf(__uint_as_float(r[e4 * 4 + 0]) + bv.x, 0.f);
                x[e4 * 4 + 1] = fmaxf(__uint_as_float(r[e4 * 4 + 1]) + bv.y, 0.f);
                x[e4 * 4 + 2] = fmaxf(__uint_as_float(r[e4 * 4 + 2]) + bv.z, 0.f);
                x[e4 * 4 + 3] = fmaxf(__uint_as_float(r[e4 * 4 + 3]) + bv.w, 0.f);
              }
              if (SAVE) {
#pragma unroll
                for (int e = 0; e < 16; ++e) bits |= half2_pos_mask(pack_half2(x[2 * e], x[2 * e + 1])) & mask_pair_const(e);
              }
              mw[cc] = bits;
              part[colbase + cc * 32 + lane] = warp_transpose_sum(x, lane);
            }
            if (SAVE) *reinterpret_cast<uint2*>(mrow) = make_uint2(mw[0], mw[1]);
            tc_fence_before_sync();
            mbar_arrive(smem_u32(&bars->acc_free[h]));
            PROF3_ACC(4);
          }
        }
      }
    }
    if (prof) for (int c = 0; c < 5; ++c) g_chain_prof[blockIdx.x][6 + c] = pacc[c];
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// weight-gradient kernel: dW_l[o][i] = sum_rows dZ_{l+1}[r][o] * H_l[r][i]
// Both operands are the swizzled tile images the chain kernels streamed to HBM, read here as MN-major UMMA
// operands (A = dZ^T, B = H).  One 256x256 fp32 accumulator (two M=128 halves = all 512 TMEM columns) stays
// resident for the CTA's whole tile range; per-CTA partials are summed by a fixed-order reduce kernel.
//   warp 0: bulk-copy producer (half tiles: 64 rows x 256 cols of each operand = 64 KB per stage, 3 stages)
//   warp 1: MMA issuer (M128 x N256 x K16, 8 per stage)      warps 2-5: final TMEM -> global epilogue
// ------------------------------------------------------------------------------------------------
constexpr int kWgThreads = 192;
constexpr int kWgStageBytes = 65536;
constexpr int kWgStages = 3;
constexpr int kWgSmemBar = kWgStages * kWgStageBytes;
constexpr int kWgSmemLaunch = kWgSmemBar + 128 + 4096 + 1024;      // barriers, bit-count exchange (regen mode), alignment slack
constexpr uint32_t kIdescWgrad = idesc_f16(kTileM, kG, 1, 1);

struct WgradParams {
  const char* dZ;          // image of tile t at dZ + t * dz_stride
  size_t dz_stride;
  const char* H;
  size_t h_stride;
  float* partial;          // [grid][256][256]
  float* colpart;          // [tiles][256]: column sums of each dZ tile (bias / question-injection gradients)
  int num_tiles;
  // regen != 0 (layer 3): the dZ operand is dZ4 = S * dxg .* (Z4 > 0), regenerated here from the sign bits (32 B per
  // row instead of a 512 B image row that the dgrad kernel would have to write and this kernel read back)
  // regen == 2 (layer 1, n == 64): the H operand is H1 = relu(U[c] + V'[a]), regenerated here from U / V' (L2-resident,
  // read once per tile for both halves) instead of a tile image the forward kernel would have to write and this one read
  int regen;
  const float* U;          // [B, 64, 256]
  const float* Vb;         // [B, 64, 256]
  const uint32_t* masks4;  // [tiles][128][8]
  const float* dxg;        // [B, 256]
  const float* scale;
  int tiles_per_sample;
};

struct WgBars {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1) rn_g_wgrad_kernel(const WgradParams p) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgBars* bars = reinterpret_cast<WgBars*>(smem + kWgSmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = tiles_of_cta(p.num_tiles);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), p.regen ? 1 + 4 : 1);   // producer (+ the 4 generating warps)
      mbar_init(smem_u32(&bars->empty[s]), p.regen == 1 ? 1 : 1 + 4);      // MMA commit (+ the 4 column-sum warps)
    }
    mbar_init(smem_u32(&bars->done), 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const size_t tile = blockIdx.x + (size_t)i * gridDim.x;
        for (int half = 0; half < 2; ++half) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t full = smem_u32(&bars->full[stage]);
          mbar_expect_tx(full, p.regen ? kWgStageBytes / 2 : kWgStageBytes);
          const uint32_t dst = smem_u32(smem + stage * kWgStageBytes);
#pragma unroll
          for (int c = 0; c < kNKC; ++c) {      // rows [64*half, 64*half+64) of column chunk c: 8 KB contiguous
            if (p.regen != 1) bulk_g2s(dst + c * 8192, p.dZ + tile * p.dz_stride + c * kAChunk + half * 8192, 8192, full);
            if (p.regen != 2) bulk_g2s(dst + 32768 + c * 8192, p.H + tile * p.h_stride + c * kAChunk + half * 8192, 8192, full);
          }
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t accumulate = 0;
      for (int i = 0; i < 2 * my_tiles; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        tc_fence_after_sync();
        const uint32_t base = smem_u32(smem + stage * kWgStageBytes);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // MN-major SWIZZLE_128B: LBO = stride between 64-element MN atoms (8 KB), SBO = between 8-row K groups
            const uint64_t ad = smem_desc_sw128(base + (2 * h) * 8192 + k * 2048, 8192, 1024);
            const uint64_t bd = smem_desc_sw128(base + 32768 + k * 2048, 8192, 1024);
            mma_f16_ss(tmem_base + h * kG, ad, bd, kIdescWgrad, accumulate | (uint32_t)(k > 0));
          }
        accumulate = 1;
        mma_commit(smem_u32(&bars->empty[stage]));
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      mma_commit(smem_u32(&bars->done));
    }
  } else {
    // warps 2..5.  Main loop: column sums of every dZ half-tile straight from the staged operand (these warps
    // would otherwise idle; the tensor pipe is HBM-bound here) -> colpart[tile][256].  Thread t owns columns 2t, 2t+1.
    {
      const int t = threadIdx.x - 64;                  // 0..127
      const int col = 2 * t;
      const uint32_t coff = (uint32_t)(col >> 6) * 8192u + (uint32_t)(col & 7) * 2u;
      const uint32_t grp = (uint32_t)((col & 63) >> 3);
      uint32_t stage = 0, phase = 0;
      float2 acc = make_float2(0.f, 0.f);
      if (p.regen == 1) {
        // Regenerate every dZ4 half tile into the stage's A slot: warp gw takes rows [16*gw, 16*gw + 16), lane owns columns
        // [8*lane, 8*lane + 8) (chunk lane / 8, group lane % 8).  Mask words and dxg of the NEXT half tile are loaded while
        // this one is generated.  Column sums (bias gradient) come from per-column bit counts: count * S*dxg.
        const int gw = warp - 2;
        const float S = __ldg(p.scale);
        uint32_t* cnts = reinterpret_cast<uint32_t*>(smem + kWgSmemBar + 128);       // [2][4][128] packed 16-bit count pairs
        const int sh = (lane & 3) * 4;
        auto tile_of = [&](int i) { return (size_t)blockIdx.x + (size_t)(i >> 1) * gridDim.x; };
        auto load_masks = [&](int i, uint32_t (&w)[16]) {
          const uint32_t* m4 = p.masks4 + (tile_of(i) * kTileM + (i & 1) * 64 + gw * 16) * 8 + (lane >> 2);
#pragma unroll
          for (int r = 0; r < 16; ++r) w[r] = __ldg(m4 + r * 8);
        };
        uint32_t w[16], wn[16];
        float4 d0, d1;
        float2 dmine;
        auto load_dxg = [&](int i) {
          const float* row = p.dxg + (size_t)(tile_of(i) / p.tiles_per_sample) * kG;
          d0 = __ldg(reinterpret_cast<const float4*>(row + lane * 8));
          d1 = __ldg(reinterpret_cast<const float4*>(row + lane * 8 + 4));
          dmine = __ldg(reinterpret_cast<const float2*>(row + col));
        };
        if (my_tiles > 0) { load_masks(0, w); load_dxg(0); }
        for (int i = 0; i < 2 * my_tiles; ++i) {
          const uint32_t dp[4] = {pack_half2(d0.x * S, d0.y * S), pack_half2(d0.z * S, d0.w * S),
                                  pack_half2(d1.x * S, d1.y * S), pack_half2(d1.z * S, d1.w * S)};
          const float2 vmine = __half22float2(__floats2half2_rn(dmine.x * S, dmine.y * S));     // the fp16 values that are summed
          if (i + 1 < 2 * my_tiles) {
            load_masks(i + 1, wn);
            if (i & 1) load_dxg(i + 1);          // next tile
          }
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);          // the MMAs are done with this stage
          char* dst = smem + stage * kWgStageBytes + (lane >> 3) * 8192;
          uint32_t cnt[4] = {0, 0, 0, 0};
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const int rl = gw * 16 + r;
            const uint32_t nib = w[r] >> sh;
            uint32_t b4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              b4[k] = (nib >> k) & 0x00010001u;      // pair (2k, 2k+1) of this lane's 8 columns: bit 0 / bit 16
              cnt[k] += b4[k];
            }
            uint4 o;
            o.x = dp[0] & (b4[0] * 0xFFFFu);
            o.y = dp[1] & (b4[1] * 0xFFFFu);
            o.z = dp[2] & (b4[2] * 0xFFFFu);
            o.w = dp[3] & (b4[3] * 0xFFFFu);
            *reinterpret_cast<uint4*>(dst + sw128_offset(rl, (lane & 7) * 8)) = o;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->full[stage]));
          uint32_t* cb = cnts + (i & 1) * 512;
          *reinterpret_cast<uint4*>(cb + gw * 128 + lane * 4) = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
          named_bar_sync(1, 128);
          const uint32_t c4 = cb[t] + cb[128 + t] + cb[256 + t] + cb[384 + t];       // word t = columns 2t (low), 2t+1 (high)
          acc.x += (float)(c4 & 0xFFFFu) * vmine.x;
          acc.y += (float)(c4 >> 16) * vmine.y;
          if (i & 1) {                                   // second half of the tile
            *reinterpret_cast<float2*>(p.colpart + tile_of(i) * kG + col) = acc;
            acc = make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int r = 0; r < 16; ++r) w[r] = wn[r];
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      } else if (p.regen == 2) {
        // Regenerate the H1 halves of every tile into the B slots of TWO consecutive stages (rows (a0, c) and (a1, c)):
        // warp gw takes c in [16*gw, 16*gw + 16), lane owns columns [4*lane, +4) and [128 + 4*lane, +4).  The tile's U
        // rows live in registers; the NEXT tile's rows are loaded right after this tile is generated and land while the
        // column sums of the staged dZ halves are taken.
        const int gw = warp - 2;
        auto tile_of = [&](int it) { return (size_t)blockIdx.x + (size_t)it * gridDim.x; };
        float4 uu[16][2], vv[2][2];
        auto load_uv = [&](int it) {
          const size_t tile = tile_of(it);
          const size_t b = tile / p.tiles_per_sample;
          const int a0 = (int)(tile % p.tiles_per_sample) * 2;
          const float* Ub = p.U + (b * 64 + gw * 16) * kG + 4 * lane;
          const float* Vb = p.Vb + (b * 64 + a0) * kG + 4 * lane;
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            uu[r][0] = __ldg(reinterpret_cast<const float4*>(Ub + (size_t)r * kG));
            uu[r][1] = __ldg(reinterpret_cast<const float4*>(Ub + (size_t)r * kG + 128));
          }
#pragma unroll
          for (int ai = 0; ai < 2; ++ai) {
            vv[ai][0] = __ldg(reinterpret_cast<const float4*>(Vb + ai * kG));
            vv[ai][1] = __ldg(reinterpret_cast<const float4*>(Vb + ai * kG + 128));
          }
        };
        auto colsum_half = [&](uint32_t stg, uint32_t ph, int i) {
          mbar_wait(smem_u32(&bars->full[stg]), ph);
          const char* st = smem + stg * kWgStageBytes + coff;
#pragma unroll 8
          for (int r = 0; r < 64; ++r) {
            const __half2 h = *reinterpret_cast<const __half2*>(st + (r >> 3) * 1024 + (r & 7) * 128 + ((grp ^ (uint32_t)(r & 7)) << 4));
            const float2 f = __half22float2(h);
            acc.x += f.x;
            acc.y += f.y;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->empty[stg]));
          if (i & 1) {
            *reinterpret_cast<float2*>(p.colpart + tile_of(i >> 1) * kG + col) = acc;
            acc = make_float2(0.f, 0.f);
          }
        };
        if (my_tiles > 0) load_uv(0);
        const uint32_t goff = (uint32_t)(lane >> 4) * 8192u + (uint32_t)((4 * lane) & 7) * 2u;      // chunk + byte inside the 16-byte group
        const int ggrp = ((4 * lane) & 63) >> 3;
        for (int it = 0; it < my_tiles; ++it) {
          const uint32_t s_a = stage, ph_a = phase;
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          const uint32_t s_b = stage, ph_b = phase;
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          mbar_wait(smem_u32(&bars->empty[s_a]), ph_a ^ 1);
          mbar_wait(smem_u32(&bars->empty[s_b]), ph_b ^ 1);
#pragma unroll
          for (int ai = 0; ai < 2; ++ai) {
            char* dst = smem + (ai == 0 ? s_a : s_b) * kWgStageBytes + 32768 + goff;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
              const int c = gw * 16 + r;
              const uint32_t off = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128 + (((ggrp ^ (c & 7)) & 7) << 4));
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint2 o;
                o.x = pack_relu_half2(uu[r][h].x + vv[ai][h].x, uu[r][h].y + vv[ai][h].y);
                o.y = pack_relu_half2(uu[r][h].z + vv[ai][h].z, uu[r][h].w + vv[ai][h].w);
                *reinterpret_cast<uint2*>(dst + h * 2 * 8192 + off) = o;
              }
            }
            fence_proxy_async_smem();          // publish this half at once: its MMAs start while the other half is generated
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->full[ai == 0 ? s_a : s_b]));
          }
          if (it + 1 < my_tiles) load_uv(it + 1);
          colsum_half(s_a, ph_a, 2 * it);
          colsum_half(s_b, ph_b, 2 * it + 1);
        }
      } else {
      for (int i = 0; i < 2 * my_tiles; ++i) {
        mbar_wait(smem_u32(&bars->full[stage]), phase);
        const char* st = smem + stage * kWgStageBytes + coff;
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
          const __half2 h = *reinterpret_cast<const __half2*>(st + (r >> 3) * 1024 + (r & 7) * 128 + ((grp ^ (uint32_t)(r & 7)) << 4));
          const float2 f = __half22float2(h);
          acc.x += f.x;
          acc.y += f.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->empty[stage]));
        if (i & 1) {                                   // second half of the tile
          const size_t tile = blockIdx.x + (size_t)(i >> 1) * gridDim.x;
          *reinterpret_cast<float2*>(p.colpart + tile * kG + col) = acc;
          acc = make_float2(0.f, 0.f);
        }
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      }
    }
    // final epilogue: TMEM lane quarters 2, 3, 0, 1
    const int q = warp & 3;
    mbar_wait(smem_u32(&bars->done), 0);
    tc_fence_after_sync();
    float* out = p.partial + (size_t)blockIdx.x * kG * kG;
#pragma unroll 1
    for (int h = 0; h < 2; ++h)
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * kG + cc * 32, r);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(out + (size_t)(h * 128 + q * 32 + lane) * kG + cc * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]),
                               __uint_as_float(r[4 * e + 3]));
      }
    tc_fence_before_sync();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// dW[o * ld + i] = scale[1] * sum_parts partial[part][o][i]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dW, int ld,
                                    const float* __restrict__ scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kG * kG) return;
  float v = 0.f;
  for (int pidx = 0; pidx < nparts; ++pidx) v += partial[(size_t)pidx * kG * kG + idx];
  dW[(size_t)(idx / kG) * ld + idx % kG] = v * scale[1];
}

// S = 2^k with max|dxg| * S in (4, 8]: keeps the fp16 gradient operands in the normal range.
// Two stages (per-block maxima, then one block) so the scan is not a single block's chain of L2 round trips.
__global__ void __launch_bounds__(256) grad_absmax_kernel(const float* __restrict__ dxg, int n, float* __restrict__ part) {
  __shared__ float red[8];
  float m = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) m = fmaxf(m, fabsf(dxg[i]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    part[blockIdx.x] = m;
  }
}
__global__ void grad_scale_kernel(const float* __restrict__ part, int nparts, float* __restrict__ scale) {
  float m = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 32) m = fmaxf(m, part[i]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) {
    float S = 1.f;
    if (m > 0.f && isfinite(m)) S = exp2f(3.f - ceilf(log2f(m)));
    scale[0] = S;
    scale[1] = 1.f / S;
  }
}

__global__ void scale_inplace_kernel(float* __restrict__ x, long long n, const float* __restrict__ scale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= scale[1];
}

// dU[b, c, :] = (1/S) sum_a dZ1[(a, c), :],  dV[b, a, :] = (1/S) sum_c dZ1[(a, c), :]   from the dZ1 tile images.
// grid (B, 4 column chunks), block 256; each thread owns (object, 16-byte column group) accumulators and walks
// the other pair index in a fixed order -> deterministic.
__global__ void __launch_bounds__(256)
dz1_reduce_kernel(const char* __restrict__ dZ, size_t tile_stride, float* __restrict__ dU, float* __restrict__ dV,
                  const float* __restrict__ scale, int n, int tiles_per_sample) {
  const int b = blockIdx.x, kc = blockIdx.y;
  const float inv = scale[1];
  const char* base = dZ + (size_t)b * tiles_per_sample * tile_stride + (size_t)kc * kAChunk;
  for (int pass = 0; pass < 2; ++pass) {
    float* out = pass == 0 ? dU : dV;
    for (int item = threadIdx.x; item < n * 8; item += 256) {
      const int obj = item >> 3, j = item & 7;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int o = 0; o < n; ++o) {
        const int pr = pass == 0 ? o * n + obj : obj * n + o;      // pair row a*n + c
        const uint4 v = *reinterpret_cast<const uint4*>(base + (size_t)(pr / kTileM) * tile_stride +
                                                        sw128_offset(pr % kTileM, j * 8));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
      float* dst = out + ((size_t)b * n + obj) * kG + kc * 64 + j * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    }
  }
}

// n == 64: single pass over the dZ1 images (the generic kernel above reads them twice).  Block (b, kc) walks the 32
// tiles of sample b; chunk kc of each tile image (128 rows x 64 columns, 16 KB, contiguous) is staged by the TMA
// engine through a 4-deep ring.  A tile is (a0, a1) x all 64 c:
//   dU: thread (c = tid / 4, part = tid % 4) adds rows c and 64 + c of its 16 columns into registers over all tiles;
//   dV: warps 0 / 1 take the column sums of rows [0, 64) / [64, 128) (lane = column pair) and write dV[a] per tile.
constexpr int kDz1Stages = 4;
__global__ void __launch_bounds__(256)
dz1_reduce_n64_kernel(const char* __restrict__ dZ, size_t tile_stride, float* __restrict__ dU, float* __restrict__ dV,
                      const float* __restrict__ scale) {
  extern __shared__ __align__(128) char dz1_smem[];
  char (*stage)[kAChunk] = reinterpret_cast<char (*)[kAChunk]>(dz1_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(dz1_smem + kDz1Stages * kAChunk);
  const int b = blockIdx.x, kc = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kTiles = 32;
  const char* src = dZ + (size_t)b * kTiles * tile_stride + (size_t)kc * kAChunk;
  if (tid == 0) {
    for (int s = 0; s < kDz1Stages; ++s) mbar_init(smem_u32(&full[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int t = 0; t < kDz1Stages; ++t) {
      mbar_expect_tx(smem_u32(&full[t]), kAChunk);
      bulk_g2s(smem_u32(stage[t]), src + (size_t)t * tile_stride, kAChunk, smem_u32(&full[t]));
    }
  }
  const float inv = scale[1];
  const int c = tid >> 2, part = tid & 3;
  float au[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) au[e] = 0.f;
  for (int t = 0; t < kTiles; ++t) {
    const int s = t % kDz1Stages;
    mbar_wait(smem_u32(&full[s]), (t / kDz1Stages) & 1);
    const char* st = stage[s];
#pragma unroll
    for (int ai = 0; ai < 2; ++ai) {
      const int row = ai * 64 + c;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int g = part * 2 + i;
        const uint4 v = *reinterpret_cast<const uint4*>(st + (row >> 3) * 1024 + (row & 7) * 128 + ((g ^ (row & 7)) << 4));
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h[e]);
          au[i * 8 + 2 * e] += f.x;
          au[i * 8 + 2 * e + 1] += f.y;
        }
      }
    }
    if (warp < 2) {            // dV[a = 2t + warp][kc*64 + 2*lane, +1] = sum over the 64 rows of this half
      float2 acc = make_float2(0.f, 0.f);
      const uint32_t coff = (uint32_t)(lane & 3) * 4u, grp = (uint32_t)(lane >> 2);
#pragma unroll 8
      for (int r = 0; r < 64; ++r) {
        const int row = warp * 64 + r;
        const __half2 h = *reinterpret_cast<const __half2*>(st + (row >> 3) * 1024 + (row & 7) * 128 + ((grp ^ (uint32_t)(row & 7)) << 4) + coff);
        const float2 f = __half22float2(h);
        acc.x += f.x;
        acc.y += f.y;
      }
      *reinterpret_cast<float2*>(dV + ((size_t)b * 64 + 2 * t + warp) * kG + kc * 64 + 2 * lane) = make_float2(acc.x * inv, acc.y * inv);
    }
    __syncthreads();           // everyone is done with this stage
    if (tid == 0 && t + kDz1Stages < kTiles) {
      mbar_expect_tx(smem_u32(&full[s]), kAChunk);
      bulk_g2s(smem_u32(stage[s]), src + (size_t)(t + kDz1Stages) * tile_stride, kAChunk, smem_u32(&full[s]));
    }
  }
  float* dst = dU + ((size_t)b * 64 + c) * kG + kc * 64 + part * 16;
#pragma unroll
  for (int e = 0; e < 16; e += 4)
    *reinterpret_cast<float4*>(dst + e) = make_float4(au[e] * inv, au[e + 1] * inv, au[e + 2] * inv, au[e + 3] * inv);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool tc_supported(const RelShape& s) {
  return s.G == kG && s.L == 4 && s.pairs % kTileM == 0 && s.qinj >= 0 && s.qinj < 4;
}

static size_t wpack_bytes() { return (size_t)kTcLayers * 2 * kNKC * kWChunk; }

struct TcSaved {
  RelPre pre;
  __half* wpack;       // forward weight images
  __half* saveH;       // [tiles][3] x 64 KB
  uint32_t* masks;     // [4][tiles][128][8]
};

// the training forward of the parity mode is the 3-pass kernel unless the caller asks for the 2-pass one
static bool tc_fwd3(const RelShape& s, int precision, bool training) {
  return training && precision == RN_PRECISION_PARITY && !(s.flags & RN_REL_FLAG_FWD_2PASS);
}

static TcSaved tc_carve_saved(const RelShape& s, void* saved, bool training, bool fwd3) {
  Carver c(saved);
  TcSaved o;
  o.pre.U = c.take<float>((size_t)s.B * s.n * s.G);
  if (fwd3) o.pre.U4 = c.take<float>((size_t)s.B * s.n * s.G);
  o.pre.Vb = c.take<float>((size_t)s.B * s.n * s.G);
  o.pre.Qb = c.take<float>((size_t)s.B * s.G);
  o.wpack = reinterpret_cast<__half*>(c.take<char>(wpack_bytes()));
  const size_t tiles = s.rows / kTileM;
  o.saveH = training ? reinterpret_cast<__half*>(c.take<char>(tiles * 3 * kATile)) : nullptr;
  o.masks = training ? c.take<uint32_t>(tiles * 4 * kTileM * 8) : nullptr;
  return o;
}

size_t tc_saved_bytes(const RelShape& s, int precision, bool training) {
  const size_t tiles = s.rows / kTileM;
  size_t b = 2 * round_up((size_t)s.B * s.n * s.G * 4, 256) + round_up((size_t)s.B * s.G * 4, 256) + round_up(wpack_bytes(), 256);
  if (tc_fwd3(s, precision, training)) b += round_up((size_t)s.B * s.n * s.G * 4, 256);
  if (training) b += round_up(tiles * 3 * kATile, 256) + round_up(tiles * 4 * kTileM * 8 * 4, 256);
  return b;
}

constexpr int kScaleParts = 254;

struct TcBwdScratch {
  float* scale;
  __half* wpackT;
  __half* dZ;          // [tiles][4] x 64 KB
  float* colpart;      // [3][tiles][256] (written by the weight-gradient kernels)
  float* partial;      // [grid][256][256]
  float* delta;        // [3][B][256]
  float* dU;
  float* dV;
};

static TcBwdScratch tc_carve_bwd(const RelShape& s, void* scratch) {
  Carver c(scratch);
  TcBwdScratch o;
  const size_t tiles = s.rows / kTileM;
  o.scale = c.take<float>(2 + kScaleParts);       // S, 1/S, per-block |dxg| maxima
  o.wpackT = reinterpret_cast<__half*>(c.take<char>(wpack_bytes()));
  o.dZ = reinterpret_cast<__half*>(c.take<char>(tiles * 4 * kATile));
  o.colpart = c.take<float>(tiles * 3 * kG);
  o.partial = c.take<float>((size_t)256 * kG * kG);
  o.delta = c.take<float>((size_t)3 * s.B * kG);
  o.dU = c.take<float>((size_t)s.B * s.n * kG);
  o.dV = c.take<float>((size_t)s.B * s.n * kG);
  return o;
}

size_t tc_scratch_bytes(const RelShape& s, bool training) {
  const size_t tiles = s.rows / kTileM;
  const size_t fwd = round_up(tiles * 4 * kG * 4, 256);
  if (!training) return fwd;
  size_t bwd = round_up((2 + kScaleParts) * 4, 256) + round_up(wpack_bytes(), 256) + round_up(tiles * 4 * kATile, 256) + round_up(tiles * 3 * kG * 4, 256) +
               round_up((size_t)256 * kG * kG * 4, 256) + round_up((size_t)3 * s.B * kG * 4, 256) +
               2 * round_up((size_t)s.B * s.n * kG * 4, 256);
  return fwd > bwd ? fwd : bwd;
}

static void fill_pack_args(const RelShape& s, const float* const* g_w, PackArgs& pa) {
  for (int l = 0; l < kTcLayers; ++l) {
    pa.w[l] = g_w[l + 1];
    pa.ld[l] = s.fan_in(l + 1);
  }
}

}  // namespace rn
extern "C" int rn_debug_chain_profile(long long* out) {   // 160 x 16 counters of the last chain launch (RN_B200_DBG & 8)
  return cudaMemcpyFromSymbol(out, rn::g_chain_prof, sizeof(rn::g_chain_prof)) == cudaSuccess ? 0 : 1;
}
namespace rn {
// H1 is regenerated inside the layer-1 weight-gradient kernel (8x8 grid only; RN_B200_REGEN_H1=0 streams its image like
// H2 / H3; the generator-warpgroup form always stores it).  Forward and backward must agree.
static bool tc_regen_h1(const RelShape& s) {
  static const bool on = []() { const char* e = getenv("RN_B200_REGEN_H1"); return !(e && e[0] == '0'); }();
  static const int gen_env = []() { const char* e = getenv("RN_B200_GENWG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  return on && gen_env != 1 && s.n == 64;
}
static int chain_sched() {
  static const int v = []() { const char* e = getenv("RN_B200_SCHED"); return e ? atoi(e) : 0; }();
  return v;
}
static int chain_dbg() {
  static const int v = []() { const char* e = getenv("RN_B200_DBG"); return e ? atoi(e) : 0; }();
  return v;
}

template <int MODE>
static int launch_chain(const ChainParams& p, cudaStream_t st) {
  // CTA-pair (cta_group::2) kernel: opt-in with RN_B200_CTA2=1.  It is parity-tested, but measured 2.98 ms vs
  // 2.50 ms for the single-CTA form at B=640: the ALU-pipe-bound epilogue, not operand traffic, is the limiter.
  static const bool allow_pair = []() { const char* e = getenv("RN_B200_CTA2"); return e && e[0] == '1'; }();
  if (allow_pair && p.num_tiles % 2 == 0 && p.num_tiles >= 2) {
    const int pairs = std::min(p.num_tiles / 2, sm_count() / 2);
    RN_CUDA(cudaFuncSetAttribute(rn_g_chain_kernel<MODE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kFwdThreads);
    cfg.dynamicSmemBytes = kSmemLaunch;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RN_CUDA(cudaLaunchKernelEx(&cfg, rn_g_chain_kernel<MODE, true, false>, p));
    RN_LAUNCH_CHECK("rn_g_chain_kernel<pair>");
    return RN_OK;
  }
  const int grid = std::min(p.num_tiles, sm_count());
  // Generator-warpgroup form: measured at B=640 it hides the tile boundary in the eval forward (1.65 vs 1.68 ms) but not
  // in the training forward / data gradient, where the extra warpgroup competes with the epilogue warps for issue slots
  // and (dgrad) dZ1 has to be stored straight from registers (2.28 vs 2.22 ms, 3.38 vs 3.19 ms).  Default: eval only;
  // RN_B200_GENWG=0 / 1 forces the 384-thread / 512-thread form everywhere.
  static const int gen_env = []() { const char* e = getenv("RN_B200_GENWG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  const bool gen_wg = gen_env < 0 ? MODE == kFwdEval : gen_env == 1;
  if (gen_wg) {
    RN_CUDA(cudaFuncSetAttribute(rn_g_chain_kernel<MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    rn_g_chain_kernel<MODE, false, true><<<grid, kGenThreads, kSmemLaunch, st>>>(p);
  } else {
    RN_CUDA(cudaFuncSetAttribute(rn_g_chain_kernel<MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    rn_g_chain_kernel<MODE, false, false><<<grid, kFwdThreads, kSmemLaunch, st>>>(p);
  }
  RN_LAUNCH_CHECK("rn_g_chain_kernel");
  return RN_OK;
}

int tc_relation_fwd(const RelShape& s, int precision, bool training, const float* x, const float* q,
                    const float* const* g_w, const float* const* g_b, float* xg, void* saved, void* scratch,
                    cudaStream_t st) {
  const bool fwd3 = tc_fwd3(s, precision, training);
  TcSaved sv = tc_carve_saved(s, saved, training, fwd3);
  RN_TRY(relation_pre(s, x, q, g_w, g_b, sv.pre, st));
  PackArgs pa;
  fill_pack_args(s, g_w, pa);
  if (fwd3) pack_weights3_kernel<<<cdiv(kTcLayers * kG * (kG / 8), 256), 256, 0, st>>>(pa, sv.wpack);
  else pack_weights_kernel<<<cdiv(kTcLayers * kG * (kG / 8), 256), 256, 0, st>>>(pa, sv.wpack, 0);
  RN_LAUNCH_CHECK("pack_weights_kernel");

  ChainParams p = {};
  p.dbg = chain_dbg();
  p.sched = chain_sched();
  p.U = sv.pre.U;
  p.Vb = sv.pre.Vb;
  for (int l = 0; l < kTcLayers; ++l) {
    if (l + 1 == s.qinj) {
      p.bias[l] = sv.pre.Qb;
      p.bias_stride[l] = s.G;
    } else {
      p.bias[l] = g_b[l + 1];
      p.bias_stride[l] = 0;
    }
  }
  p.wpack = sv.wpack;
  p.xg_part = static_cast<float*>(scratch);
  p.saveH = sv.saveH;
  p.skip_h1_image = (training && tc_regen_h1(s)) ? 1 : 0;
  p.masks = sv.masks;
  p.n = s.n;
  p.tiles_per_sample = (int)(s.pairs / kTileM);
  p.num_tiles = (int)(s.rows / kTileM);
  p.passes = precision == RN_PRECISION_FAST ? 1 : 2;
  p.U4 = sv.pre.U4;
  if (fwd3) {
    if (p.dbg & 8) {
      RN_CUDA(cudaFuncSetAttribute(rn_g_fwd3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3SmemLaunch));
      rn_g_fwd3_kernel<true, true><<<std::min(p.num_tiles, sm_count()), k3Threads, k3SmemLaunch, st>>>(p);
    } else {
      RN_CUDA(cudaFuncSetAttribute(rn_g_fwd3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3SmemLaunch));
      rn_g_fwd3_kernel<true, false><<<std::min(p.num_tiles, sm_count()), k3Threads, k3SmemLaunch, st>>>(p);
    }
    RN_LAUNCH_CHECK("rn_g_fwd3_kernel");
  } else if (training) {
    RN_TRY(launch_chain<kFwdTrain>(p, st));
  } else {
    RN_TRY(launch_chain<kFwdEval>(p, st));
  }
  // x_g[b] = sum over the sample's tiles and the 4 row quarters (fixed order -> deterministic)
  RN_TRY(colsum(p.xg_part, xg, s.G, s.B, 1, (long long)p.tiles_per_sample * 4, 0, 1, p.tiles_per_sample * 4, st));
  return RN_OK;
}

int tc_relation_bwd(const RelShape& s, int precision, const float* dxg, const float* x, const float* q,
                    const float* const* g_w, const void* saved, float* dx, float* dq, float* const* dg_w,
                    float* const* dg_b, void* scratch, cudaStream_t st) {
  const bool fwd3 = tc_fwd3(s, precision, true);
  TcSaved sv = tc_carve_saved(s, const_cast<void*>(saved), true, fwd3);
  TcBwdScratch ws = tc_carve_bwd(s, scratch);
  const int tps = (int)(s.pairs / kTileM);
  const int tiles = (int)(s.rows / kTileM);

  {
    const int nparts = std::min(kScaleParts, cdiv((long long)s.B * s.G, 256));
    grad_absmax_kernel<<<nparts, 256, 0, st>>>(dxg, s.B * s.G, ws.scale + 2);
    RN_LAUNCH_CHECK("grad_absmax_kernel");
    grad_scale_kernel<<<1, 32, 0, st>>>(ws.scale + 2, nparts, ws.scale);
    RN_LAUNCH_CHECK("grad_scale_kernel");
  }
  PackArgs pa;
  fill_pack_args(s, g_w, pa);
  pack_weights_kernel<<<cdiv(kTcLayers * kG * (kG / 8), 256), 256, 0, st>>>(pa, ws.wpackT, 1);
  RN_LAUNCH_CHECK("pack_weights_kernel");

  // data gradient chain: dZ4 -> dZ3 -> dZ2 -> dZ1 (images to HBM), column sums of dZ2..dZ4
  ChainParams p = {};
  p.dbg = chain_dbg();
  p.sched = chain_sched();
  p.wpack = ws.wpackT;
  p.n = s.n;
  p.tiles_per_sample = tps;
  p.num_tiles = tiles;
  // one pass (W_hi only) for the data gradient unless RN_REL_FLAG_DGRAD_2PASS: fp16 weight rounding (2^-12 relative,
  // independent per weight) does not move the gradient error (measured: profiles/r02_grad_error_vs_batch.json)
  p.passes = (precision != RN_PRECISION_FAST && (s.flags & RN_REL_FLAG_DGRAD_2PASS)) ? 2 : 1;
  p.m1_layout0 = fwd3 ? 1 : 0;
  p.masks = sv.masks;
  p.dxg = dxg;
  p.scale = ws.scale;
  p.dZ = ws.dZ;
  // dZ4 is regenerated inside the layer-3 weight-gradient kernel (RN_B200_REGEN_DZ4=0: stream its image like dZ1..dZ3)
  static const bool regen_dz4 = []() { const char* e = getenv("RN_B200_REGEN_DZ4"); return !(e && e[0] == '0'); }();
  static const int gen_env = []() { const char* e = getenv("RN_B200_GENWG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  p.skip_dz4_image = (regen_dz4 && gen_env != 1) ? 1 : 0;      // (the generator-warpgroup form always stores it)
  RN_TRY(launch_chain<kDgrad>(p, st));

  // weight gradients of g layers 1..3: dW_l = dZ_{l+1}^T H_l
  const int wgrid = std::min(tiles, sm_count());
  RN_CUDA(cudaFuncSetAttribute(rn_g_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemLaunch));
  for (int l = 1; l <= kTcLayers; ++l) {
    WgradParams wp;
    wp.dZ = reinterpret_cast<const char*>(ws.dZ) + (size_t)l * kATile;            // dZ_{l+1}
    wp.dz_stride = (size_t)4 * kATile;
    wp.H = reinterpret_cast<const char*>(sv.saveH) + (size_t)(l - 1) * kATile;    // H_l
    wp.h_stride = (size_t)3 * kATile;
    wp.partial = ws.partial;
    wp.colpart = ws.colpart + (size_t)(l - 1) * tiles * kG;
    wp.num_tiles = tiles;
    wp.regen = (l == kTcLayers && p.skip_dz4_image) ? 1 : (l == 1 && tc_regen_h1(s)) ? 2 : 0;
    wp.U = sv.pre.U;
    wp.Vb = sv.pre.Vb;
    wp.masks4 = sv.masks + (size_t)3 * tiles * kTileM * 8;
    wp.dxg = dxg;
    wp.scale = ws.scale;
    wp.tiles_per_sample = tps;
    rn_g_wgrad_kernel<<<wgrid, kWgThreads, kWgSmemLaunch, st>>>(wp);
    RN_LAUNCH_CHECK("rn_g_wgrad_kernel");
    wgrad_reduce_kernel<<<cdiv(kG * kG, 256), 256, 0, st>>>(ws.partial, wgrid, dg_w[l], s.fan_in(l), ws.scale);
    RN_LAUNCH_CHECK("wgrad_reduce_kernel");
    // bias gradient and question-injection gradients from the column sums of dZ_{l+1}
    float* delta = ws.delta + (size_t)(l - 1) * s.B * kG;
    RN_TRY(colsum(ws.colpart + (size_t)(l - 1) * tiles * kG, delta, kG, s.B, 1, (long long)tps, 0, 1, tps, st));
    const long long nd = (long long)s.B * kG;
    scale_inplace_kernel<<<cdiv(nd, 256), 256, 0, st>>>(delta, nd, ws.scale);
    RN_LAUNCH_CHECK("scale_inplace_kernel");
    RN_TRY(colsum(delta, dg_b[l], kG, 1, 1, 0, 0, 1, s.B, st));
    if (l == s.qinj) RN_TRY(relation_qinj_bwd(s, l, q, g_w, delta, dq, dg_w, st));
  }

  // layer 0: dU / dV from the dZ1 images, then the small fp32 products
  if (s.n == 64) {
    constexpr int smem = kDz1Stages * kAChunk + 64;
    RN_CUDA(cudaFuncSetAttribute(dz1_reduce_n64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dz1_reduce_n64_kernel<<<dim3(s.B, kNKC), 256, smem, st>>>(reinterpret_cast<const char*>(ws.dZ), (size_t)4 * kATile, ws.dU,
                                                            ws.dV, ws.scale);
  } else {
    dz1_reduce_kernel<<<dim3(s.B, kNKC), 256, 0, st>>>(reinterpret_cast<const char*>(ws.dZ), (size_t)4 * kATile, ws.dU,
                                                        ws.dV, ws.scale, s.n, tps);
  }
  RN_LAUNCH_CHECK("dz1_reduce_kernel");
  return relation_layer0_bwd(s, x, q, g_w, ws.dU, ws.dV, ws.delta, dx, dq, dg_w, dg_b, ws.partial, (size_t)256 * kG * kG, st);
}

}  // namespace rn
