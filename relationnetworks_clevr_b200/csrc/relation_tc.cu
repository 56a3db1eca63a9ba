// relation_tc.cu -- tcgen05 / TMEM implementation of g-MLP layers 1..3 for G == 256 (original-fp, ir-fp and
// the grid sweep).  See DESIGN.md "Kernels" for the pipeline description.
//
// Forward kernel rn_g_fwd_kernel (persistent, one CTA per SM, 384 threads):
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
// Precision: A operands fp16; weights W = W_hi + W_lo (two fp16 MMAs per K-step, "parity") or W_hi only ("fast").
#include "relation.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

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
constexpr int kSmemA = 0;
constexpr int kSmemW = 2 * kATile;
constexpr int kSmemBar = kSmemW + kStages * kWChunk;
constexpr int kSmemTotal = kSmemBar + 128;
constexpr int kSmemLaunch = kSmemTotal + 1024;   // slack to align the carve-up to 1024 B (SWIZZLE_128B atoms)
constexpr uint32_t kIdescFwd = idesc_f16(kTileM, kG, 0, 0);

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
// forward kernel
// ------------------------------------------------------------------------------------------------
struct FwdParams {
  const float* U;                  // [B, n, 256]
  const float* Vb;                 // [B, n, 256]  (V + beta0)
  const float* bias[kTcLayers];    // layer l+1 bias; element [b * bias_stride + col]
  long long bias_stride[kTcLayers];
  const __half* wpack;             // [3][2][4] x 32 KB
  float* xg_part;                  // [tiles][4][256]
  __half* saveH;                   // [tiles][2] x 64 KB (H2, H3 operand images) or nullptr
  uint32_t* masks;                 // [3][tiles][128][8] bit c of word w <-> column 32w+c (Z2, Z3, Z4 > 0) or nullptr
  int n;
  int tiles_per_sample;
  int num_tiles;
  int passes;                      // 2 = parity (hi + lo), 1 = fast
};

struct Bars {
  uint64_t w_full[kStages];
  uint64_t w_empty[kStages];
  uint64_t a_full[2];
  uint64_t acc_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ int tiles_of_cta(int num_tiles) {
  return (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
}

// H1 rows of one tile -> swizzled fp16 A operand.  Warp `q` (0..3) of the warpgroup owns rows [32q, 32q+32).
// Lane mapping: 4 rows x 8 sixteen-byte groups per step -> 256-byte coalesced reads of U, conflict-free STS.
__device__ __forceinline__ void generate_h1(const FwdParams& p, int tile, char* a_tile, int q, int lane) {
  const int b = tile / p.tiles_per_sample;
  const int p0 = (tile % p.tiles_per_sample) * kTileM;
  const float* Ub = p.U + (size_t)b * p.n * kG;
  const float* Vb = p.Vb + (size_t)b * p.n * kG;
  const int sub = lane >> 3, j = lane & 7;
#pragma unroll 2
  for (int g = 0; g < 8; ++g) {
    const int row = q * 32 + g * 4 + sub;
    const int pr = p0 + row;
    const int a = pr / p.n, c = pr - a * p.n;
    const float4* up = reinterpret_cast<const float4*>(Ub + (size_t)c * kG + j * 8);
    const float4* vp = reinterpret_cast<const float4*>(Vb + (size_t)a * kG + j * 8);
#pragma unroll
    for (int kc = 0; kc < kNKC; ++kc) {
      const float4 u0 = __ldg(up + kc * 16), u1 = __ldg(up + kc * 16 + 1);
      const float4 v0 = __ldg(vp + kc * 16), v1 = __ldg(vp + kc * 16 + 1);
      uint4 o;
      o.x = pack_half2(fmaxf(u0.x + v0.x, 0.f), fmaxf(u0.y + v0.y, 0.f));
      o.y = pack_half2(fmaxf(u0.z + v0.z, 0.f), fmaxf(u0.w + v0.w, 0.f));
      o.z = pack_half2(fmaxf(u1.x + v1.x, 0.f), fmaxf(u1.y + v1.y, 0.f));
      o.w = pack_half2(fmaxf(u1.z + v1.z, 0.f), fmaxf(u1.w + v1.w, 0.f));
      *reinterpret_cast<uint4*>(a_tile + kc * kAChunk + sw128_offset(row, j * 8)) = o;
    }
  }
}

template <bool SAVE>
__global__ void __launch_bounds__(kFwdThreads, 1) rn_g_fwd_kernel(const FwdParams p) {
  extern __shared__ char smem_raw[];
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + kSmemBar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = tiles_of_cta(p.num_tiles);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars->w_full[s]), 1);
      mbar_init(smem_u32(&bars->w_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 128);
      mbar_init(smem_u32(&bars->acc_full[s]), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = bars->tmem_base;

  // job sequence shared by producer and issuer: for round r: for layer 0..2: for slot 0..1 (if its tile exists)
  if (warp == 0) {
    if (lane == 0) {
      // ================= weight producer =================
      uint32_t stage = 0, phase = 0;
      for (int r = 0; 2 * r < my_tiles; ++r) {
        const int nslots = (2 * r + 1 < my_tiles) ? 2 : 1;
        for (int layer = 0; layer < kTcLayers; ++layer)
          for (int s = 0; s < nslots; ++s)
            for (int pass = 0; pass < p.passes; ++pass)
              for (int kc = 0; kc < kNKC; ++kc) {
                mbar_wait(smem_u32(&bars->w_empty[stage]), phase ^ 1);
                const uint32_t full = smem_u32(&bars->w_full[stage]);
                mbar_expect_tx(full, kWChunk);
                const char* src = reinterpret_cast<const char*>(p.wpack) + ((size_t)(layer * 2 + pass) * kNKC + kc) * kWChunk;
                bulk_g2s(smem_u32(smem + kSmemW + stage * kWChunk), src, kWChunk, full);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      uint32_t stage = 0, phase = 0;
      uint32_t a_phase[2] = {0, 0};
      for (int r = 0; 2 * r < my_tiles; ++r) {
        const int nslots = (2 * r + 1 < my_tiles) ? 2 : 1;
        for (int layer = 0; layer < kTcLayers; ++layer)
          for (int s = 0; s < nslots; ++s) {
            mbar_wait(smem_u32(&bars->a_full[s]), a_phase[s]);
            a_phase[s] ^= 1;
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + s * kG;
            const uint32_t a_base = smem_u32(smem + kSmemA + s * kATile);
            uint32_t accumulate = 0;
            for (int pass = 0; pass < p.passes; ++pass)
              for (int kc = 0; kc < kNKC; ++kc) {
                mbar_wait(smem_u32(&bars->w_full[stage]), phase);
                tc_fence_after_sync();
                const uint32_t b_base = smem_u32(smem + kSmemW + stage * kWChunk);
#pragma unroll
                for (int k = 0; k < kKC / 16; ++k) {
                  const uint64_t ad = smem_desc_sw128(a_base + kc * kAChunk + k * 32, 16, 1024);
                  const uint64_t bd = smem_desc_sw128(b_base + k * 32, 16, 1024);
                  mma_f16_ss(d_tmem, ad, bd, kIdescFwd, accumulate);
                  accumulate = 1;
                }
                mma_commit(smem_u32(&bars->w_empty[stage]));      // ring stage free once these MMAs retire
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
            mma_commit(smem_u32(&bars->acc_full[s]));             // accumulator of (slot, layer) complete
          }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue warpgroups =================
    const int s = (warp - 4) >> 2;             // tile slot
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int wg_tid = threadIdx.x - 128 - s * 128;
    char* a_tile = smem + kSmemA + s * kATile;
    const uint32_t a_full = smem_u32(&bars->a_full[s]);
    const uint32_t acc_full = smem_u32(&bars->acc_full[s]);
    const uint32_t bar_id = 1 + s;
    uint32_t acc_phase = 0;
    for (int i = s; i < my_tiles; i += 2) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int b = tile / p.tiles_per_sample;
      if (SAVE) {      // the previous tile's H3 image may still be streaming out of this A buffer
        if (wg_tid == 0) bulk_wait_read0();
        named_bar_sync(bar_id, 128);
      }
      generate_h1(p, tile, a_tile, q, lane);
      fence_proxy_async_smem();
      mbar_arrive(a_full);
      for (int layer = 0; layer < kTcLayers; ++layer) {
        mbar_wait(acc_full, acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
        const float* bias = p.bias[layer] + (size_t)b * p.bias_stride[layer];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * kG;
        uint32_t* mrow = SAVE ? p.masks + (((size_t)layer * p.num_tiles + tile) * kTileM + row) * 8 : nullptr;
        if (layer < kTcLayers - 1) {
          if (SAVE) {
            if (wg_tid == 0) bulk_wait_read0();
            named_bar_sync(bar_id, 128);
          }
          uint32_t mw[8];
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
              for (int e = 0; e < 8; ++e) {
                v[e] = fmaxf(__uint_as_float(r[g4 * 8 + e]) + bb[e], 0.f);
                bits |= (v[e] > 0.f ? 1u : 0u) << (g4 * 8 + e);
              }
              uint4 o;
              o.x = pack_half2(v[0], v[1]);
              o.y = pack_half2(v[2], v[3]);
              o.z = pack_half2(v[4], v[5]);
              o.w = pack_half2(v[6], v[7]);
              const int col = cc * 32 + g4 * 8;
              *reinterpret_cast<uint4*>(a_tile + (col >> 6) * kAChunk + sw128_offset(row, col & 63)) = o;
            }
            mw[cc] = bits;
          }
          if (SAVE) {
            *reinterpret_cast<uint4*>(mrow) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            *reinterpret_cast<uint4*>(mrow + 4) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
          }
          fence_proxy_async_smem();
          if (SAVE) {
            named_bar_sync(bar_id, 128);          // whole operand image written
            if (wg_tid == 0) {
              bulk_s2g(reinterpret_cast<char*>(p.saveH) + ((size_t)tile * 2 + layer) * kATile, smem_u32(a_tile), kATile);
              bulk_commit();
            }
          }
          tc_fence_before_sync();
          mbar_arrive(a_full);
        } else {
          // last layer: ReLU + pair-sum.  Column sums over this warp's 32 rows by shuffle transpose-reduce.
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
            for (int e = 0; e < 32; ++e) {
              x[e] = fmaxf(__uint_as_float(r[e]) + __ldg(bias + cc * 32 + e), 0.f);
              bits |= (x[e] > 0.f ? 1u : 0u) << e;
            }
            mw[cc] = bits;
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
            part[cc * 32 + lane] = x[0];           // lane L holds the sum of column cc*32 + L
          }
          if (SAVE) {
            *reinterpret_cast<uint4*>(mrow) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            *reinterpret_cast<uint4*>(mrow + 4) = make_uint4(mw[4], mw[5], mw[6], mw[7]);
          }
          tc_fence_before_sync();
        }
      }
    }
    if (SAVE && wg_tid == 0) bulk_wait0();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool tc_supported(const RelShape& s) {
  return s.G == kG && s.L == 4 && s.pairs % kTileM == 0 && s.qinj >= 0 && s.qinj < 4;
}

struct TcSaved {
  RelPre pre;
  __half* wpack;       // forward weight images
  __half* saveH;
  uint32_t* masks;
  float* biases;       // [4][256] copy of the g biases (v1 backward recomputes the forward)
};

static size_t wpack_bytes() { return (size_t)kTcLayers * 2 * kNKC * kWChunk; }

static TcSaved tc_carve_saved(const RelShape& s, void* saved, bool training) {
  Carver c(saved);
  TcSaved o;
  o.pre.U = c.take<float>((size_t)s.B * s.n * s.G);
  o.pre.Vb = c.take<float>((size_t)s.B * s.n * s.G);
  o.pre.Qb = c.take<float>((size_t)s.B * s.G);
  o.wpack = reinterpret_cast<__half*>(c.take<char>(wpack_bytes()));
  const size_t tiles = s.rows / kTileM;
  o.saveH = training ? reinterpret_cast<__half*>(c.take<char>(tiles * 2 * kATile)) : nullptr;
  o.masks = training ? c.take<uint32_t>(tiles * kTcLayers * kTileM * 8) : nullptr;
  o.biases = training ? c.take<float>(4 * kG) : nullptr;
  return o;
}

size_t tc_saved_bytes(const RelShape& s, bool training) {
  const size_t tiles = s.rows / kTileM;
  size_t b = 2 * round_up((size_t)s.B * s.n * s.G * 4, 256) + round_up((size_t)s.B * s.G * 4, 256) + round_up(wpack_bytes(), 256);
  if (training) b += round_up(tiles * 2 * kATile, 256) + round_up(tiles * kTcLayers * kTileM * 8 * 4, 256) + 4 * kG * 4;
  return b;
}

size_t tc_scratch_bytes(const RelShape& s, bool training) {
  const size_t tiles = s.rows / kTileM;
  size_t fwd = round_up(tiles * 4 * kG * 4, 256);
  // v1 backward: fp32 SIMT recompute + backward (to be replaced by the tcgen05 dgrad/wgrad kernels)
  size_t bwd = training ? simt_saved_bytes(s, true) + simt_scratch_bytes(s, true) + round_up((size_t)s.B * s.G * 4, 256) : 0;
  return fwd > bwd ? fwd : bwd;
}

int tc_relation_fwd(const RelShape& s, int precision, bool training, const float* x, const float* q,
                    const float* const* g_w, const float* const* g_b, float* xg, void* saved, void* scratch,
                    cudaStream_t st) {
  TcSaved sv = tc_carve_saved(s, saved, training);
  RN_TRY(relation_pre(s, x, q, g_w, g_b, sv.pre, st));
  if (training)
    for (int l = 0; l < 4; ++l)
      RN_CUDA(cudaMemcpyAsync(sv.biases + l * kG, g_b[l], kG * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PackArgs pa;
  for (int l = 0; l < kTcLayers; ++l) {
    pa.w[l] = g_w[l + 1];
    pa.ld[l] = s.fan_in(l + 1);
  }
  pack_weights_kernel<<<cdiv(kTcLayers * kG * (kG / 8), 256), 256, 0, st>>>(pa, sv.wpack, 0);
  RN_LAUNCH_CHECK("pack_weights_kernel");

  FwdParams p;
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
  p.masks = sv.masks;
  p.n = s.n;
  p.tiles_per_sample = (int)(s.pairs / kTileM);
  p.num_tiles = (int)(s.rows / kTileM);
  p.passes = precision == RN_PRECISION_FAST ? 1 : 2;
  const int grid = std::min(p.num_tiles, sm_count());
  if (training) {
    RN_CUDA(cudaFuncSetAttribute(rn_g_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    rn_g_fwd_kernel<true><<<grid, kFwdThreads, kSmemLaunch, st>>>(p);
  } else {
    RN_CUDA(cudaFuncSetAttribute(rn_g_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    rn_g_fwd_kernel<false><<<grid, kFwdThreads, kSmemLaunch, st>>>(p);
  }
  RN_LAUNCH_CHECK("rn_g_fwd_kernel");
  // x_g[b] = sum over the sample's tiles and the 4 row quarters (fixed order -> deterministic)
  RN_TRY(colsum(p.xg_part, xg, s.G, s.B, 1, (long long)p.tiles_per_sample * 4, 0, 1, p.tiles_per_sample * 4, st));
  return RN_OK;
}

int tc_relation_bwd(const RelShape& s, int precision, const float* dxg, const float* x, const float* q,
                    const float* const* g_w, const void* saved, float* dx, float* dq, float* const* dg_w,
                    float* const* dg_b, void* scratch, cudaStream_t st) {
  // v1: recompute the forward in fp32 with the SIMT kernels, then the SIMT backward.
  (void)precision;
  TcSaved sv = tc_carve_saved(s, const_cast<void*>(saved), true);
  const float* g_b[4] = {sv.biases, sv.biases + kG, sv.biases + 2 * kG, sv.biases + 3 * kG};
  char* base = static_cast<char*>(scratch);
  void* simt_saved = base;
  void* simt_scratch = base + simt_saved_bytes(s, true);
  float* xg_tmp = reinterpret_cast<float*>(base + simt_saved_bytes(s, true) + simt_scratch_bytes(s, true));
  RN_TRY(simt_relation_fwd(s, true, x, q, g_w, g_b, xg_tmp, simt_saved, simt_scratch, st));
  return simt_relation_bwd(s, dxg, x, q, g_w, simt_saved, dx, dq, dg_w, dg_b, simt_scratch, st);
}

}  // namespace rn
