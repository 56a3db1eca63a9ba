// sgemm.cuh -- generic fp32 SIMT GEMM with fused epilogues.
//
// Used by the RN_PRECISION_FP32 relation path (any shape), the f-MLP head and the small layer-0
// products of the tcgen05 path.  C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] (+ epilogue).
// 128x128x16 / 64x64x16 / {128,64}x32x16 tiles, 256 threads, register-prefetched smem staging; arbitrary
// leading dimensions and sub-matrix offsets (scalar, bounds-checked global loads), optional
// split-K with a deterministic second-pass reduction.
#pragma once

#include "common.cuh"

namespace rn {

struct GemmEpilogue {
  const float* bias = nullptr;   // per-column bias [N]; row r uses bias + (r / bias_group_rows) * bias_group_stride
  int bias_group_rows = 0;       // 0: one bias vector for all rows
  long long bias_group_stride = 0;
  int relu = 0;                  // max(v, 0)
  const float* mask = nullptr;   // multiply by (mask[r*ldmask + c] > 0)
  long long ldmask = 0;
  float alpha = 1.f;
  float beta = 0.f;              // C = v + beta * C_old
};

constexpr int kGemmBK = 16, kGemmThreads = 256;

// Tile shapes: 128x128 (8x8 micro-tiles) for large products; 64x64 (4x4) when the 128x128 grid would leave most of
// the 148 SMs idle (the f-MLP and other [640, 256]-sized outputs); BN = 32 variants for skinny outputs (N <= 32:
// the layer-0 products with N = k = 26).  Threads form a 16x16 grid; thread (ty, tx) owns rows ty*4+i (and
// BM/2 + ty*4 + i when TM == 8) / ty*TM + i otherwise, and the same along N.
template <int T>
__device__ __forceinline__ int gemm_frag_index(int t, int i, int B) {
  return T == 8 ? (i < 4 ? t * 4 + i : B / 2 + t * 4 + (i - 4)) : t * T + i;
}

template <bool AT, bool BT, int BM, int BN>
__global__ void __launch_bounds__(kGemmThreads)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long lda, const float* __restrict__ B,
             long long ldb, float* __restrict__ C, long long ldc, GemmEpilogue ep, int k_chunk,
             long long split_stride) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int LA = BM * kGemmBK / kGemmThreads, LB = BN * kGemmBK / kGemmThreads;     // staged elements per thread
  __shared__ __align__(16) float As[kGemmBK][BM + 4];
  __shared__ __align__(16) float Bs[kGemmBK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int tx = tid % 16, ty = tid / 16;

  // packed fp32x2 accumulators over column pairs (FFMA2: two FMAs per issue slot on sm_100)
  float2 acc2[TM][TN / 2];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) acc2[i][j] = make_float2(0.f, 0.f);

  float ra[LA], rb[LB];

  // element i of this thread's share of the A tile: (m, k) inside the tile, contiguous in memory across threads
  auto a_coord = [&](int i, int& m, int& k) {
    if (AT) { m = tid % BM; k = tid / BM + (kGemmThreads / BM) * i; }       // memory contiguous along m
    else    { k = tid % 16; m = tid / 16 + 16 * i; }                        // memory contiguous along k
  };
  auto b_coord = [&](int i, int& n, int& k) {
    if (BT) { k = tid % 16; n = tid / 16 + 16 * i; }                        // B[n*ldb + k]
    else    { n = tid % BN; k = tid / BN + (kGemmThreads / BN) * i; }       // B[k*ldb + n]
  };
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int m, k;
      a_coord(i, m, k);
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = AT ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int n, k;
      b_coord(i, n, k);
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = BT ? B[(long long)gn * ldb + gk] : B[(long long)gk * ldb + gn];
      rb[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int m, k;
      a_coord(i, m, k);
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int n, k;
      b_coord(i, n, k);
      Bs[k][n] = rb[i];
    }
  };

  if (k_begin < k_end) {
    load_tiles(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += kGemmBK) {
      store_tiles();
      __syncthreads();
      if (k0 + kGemmBK < k_end) load_tiles(k0 + kGemmBK);
#pragma unroll
      for (int kk = 0; kk < kGemmBK; ++kk) {
        float a[TM], b[TN];
        if (TM >= 4) {
          *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
          if (TM == 8) *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[kk][BM / 2 + ty * 4]);
        } else {
          *reinterpret_cast<float2*>(&a[0]) = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
        }
        if (TN >= 4) {
          *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          if (TN == 8) *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[kk][BN / 2 + tx * 4]);
        } else {
          *reinterpret_cast<float2*>(&b[0]) = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
          for (int j = 0; j < TN / 2; ++j) acc2[i][j] = __ffma2_rn(aa, make_float2(b[2 * j], b[2 * j + 1]), acc2[i][j]);
        }
      }
      __syncthreads();
    }
  }
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN / 2; ++j) {
      acc[i][2 * j] = acc2[i][j].x;
      acc[i][2 * j + 1] = acc2[i][j].y;
    }

  float* Cz = C + (long long)blockIdx.z * split_stride;
  const bool raw = split_stride != 0;   // split-K partials: no epilogue here
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + gemm_frag_index<TM>(ty, i, BM);
    if (gm >= M) continue;
    const float* bias_row = nullptr;
    if (!raw && ep.bias)
      bias_row = ep.bias + (ep.bias_group_rows ? (long long)(gm / ep.bias_group_rows) * ep.bias_group_stride : 0);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + gemm_frag_index<TN>(tx, j, BN);
      if (gn >= N) continue;
      float v = acc[i][j];
      if (!raw) {
        v *= ep.alpha;
        if (bias_row) v += bias_row[gn];
        if (ep.relu) v = fmaxf(v, 0.f);
        if (ep.mask) v = ep.mask[(long long)gm * ep.ldmask + gn] > 0.f ? v : 0.f;
        if (ep.beta != 0.f) v += ep.beta * Cz[(long long)gm * ldc + gn];
      }
      Cz[(long long)gm * ldc + gn] = v;
    }
  }
}

// second pass of split-K: C = epilogue(sum_z partial[z])
static __global__ void splitk_reduce_kernel(int M, int N, int splits, const float* __restrict__ part, float* __restrict__ C,
                                     long long ldc, GemmEpilogue ep) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * N) return;
  const int gm = idx / N, gn = idx % N;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += part[(long long)z * M * N + idx];
  v *= ep.alpha;
  if (ep.bias) v += ep.bias[(ep.bias_group_rows ? (long long)(gm / ep.bias_group_rows) * ep.bias_group_stride : 0) + gn];
  if (ep.relu) v = fmaxf(v, 0.f);
  if (ep.mask) v = ep.mask[(long long)gm * ep.ldmask + gn] > 0.f ? v : 0.f;
  if (ep.beta != 0.f) v += ep.beta * C[(long long)gm * ldc + gn];
  C[(long long)gm * ldc + gn] = v;
}

template <int BM, int BN>
static void sgemm_launch(bool at, bool bt, dim3 grid, cudaStream_t st, int M, int N, int K, const float* A, long long lda,
                         const float* B, long long ldb, float* out, long long ld_out, const GemmEpilogue& ep, int k_chunk,
                         long long split_stride) {
  if (at && bt) sgemm_kernel<true, true, BM, BN><<<grid, kGemmThreads, 0, st>>>(M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else if (at) sgemm_kernel<true, false, BM, BN><<<grid, kGemmThreads, 0, st>>>(M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else if (bt) sgemm_kernel<false, true, BM, BN><<<grid, kGemmThreads, 0, st>>>(M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else sgemm_kernel<false, false, BM, BN><<<grid, kGemmThreads, 0, st>>>(M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
}

// Launch.  `splitk_ws` (floats, >= splits*M*N) enables split-K when K is long and the tile grid small.
inline int sgemm(bool at, bool bt, int M, int N, int K, const float* A, long long lda, const float* B, long long ldb,
                 float* C, long long ldc, const GemmEpilogue& ep, cudaStream_t st, float* splitk_ws = nullptr,
                 size_t splitk_ws_floats = 0) {
  if (M <= 0 || N <= 0) return RN_OK;
  // tile choice: skinny outputs get BN = 32; otherwise 128x128 unless that grid cannot fill the machine
  int bm = 128, bn = 128;
  if (N <= 32) {
    bn = 32;
    bm = (long long)cdiv(M, 128) >= 2LL * sm_count() ? 128 : 64;
  } else if ((long long)cdiv(M, 128) * cdiv(N, 128) < sm_count()) {
    bm = bn = 64;
  }
  dim3 grid(cdiv(M, bm), cdiv(N, bn), 1);
  int splits = 1;
  const long long tiles = (long long)grid.x * grid.y;
  if (splitk_ws && K >= 1024 && tiles < 2LL * sm_count()) {
    splits = (int)min((long long)cdiv(K, 256), max(1LL, (2LL * sm_count()) / tiles));
    while (splits > 1 && (size_t)splits * M * N > splitk_ws_floats) --splits;
  }
  int k_chunk = K;
  long long split_stride = 0;
  float* out = C;
  long long ld_out = ldc;
  if (splits > 1) {
    k_chunk = cdiv(cdiv(K, splits), kGemmBK) * kGemmBK;
    splits = cdiv(K, k_chunk);
    grid.z = splits;
    split_stride = (long long)M * N;
    out = splitk_ws;
    ld_out = N;
  }
  if (bm == 128 && bn == 128) sgemm_launch<128, 128>(at, bt, grid, st, M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else if (bm == 64 && bn == 64) sgemm_launch<64, 64>(at, bt, grid, st, M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else if (bm == 128) sgemm_launch<128, 32>(at, bt, grid, st, M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  else sgemm_launch<64, 32>(at, bt, grid, st, M, N, K, A, lda, B, ldb, out, ld_out, ep, k_chunk, split_stride);
  RN_LAUNCH_CHECK("sgemm_kernel");
  if (splits > 1) {
    const long long total = (long long)M * N;
    splitk_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(M, N, splits, splitk_ws, C, ldc, ep);
    RN_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return RN_OK;
}

// ---- grouped small weight-gradient products ------------------------------------------------------------------
// C_p[M_p x N_p] = A_p^T B_p  (A_p [K x M_p], B_p [K x N_p], K = batch rows) for up to four problems in ONE launch, plus
// the column sums of A_p (bias gradients).  These are the f-MLP / question-injection weight gradients: 256 x 256 outputs
// over K = batch, which as 64x64-tile GEMMs put 16 blocks on the machine for 54 us each.  Here a block owns one 32x32
// output tile for the whole K (no split-K workspace, fixed summation order: deterministic), so the three f-MLP
// gradients are 136 blocks = one wave.  Thread (ty, tx) of 16 x 16 owns rows 2ty, 2ty+1 and columns 2tx, 2tx+1.
struct AtbProblem {
  const float* A;
  const float* B;
  float* C;
  float* colsum;      // [M] column sums of A, or nullptr
  int lda, ldb, ldc, M, N;
  int tile_begin;     // first block of this problem
};
struct AtbGroup {
  AtbProblem p[4];
  int count;
};

static __global__ void __launch_bounds__(256) atb_group_kernel(AtbGroup g, int K) {
  __shared__ __align__(16) float As[2][32][36], Bs[2][32][36];
  int pi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < g.count && (int)blockIdx.x >= g.p[i].tile_begin) pi = i;
  const AtbProblem& P = g.p[pi];
  const int tiles_n = (P.N + 31) / 32;
  const int tile = blockIdx.x - P.tile_begin;
  const int m0 = (tile / tiles_n) * 32, n0 = (tile % tiles_n) * 32;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lr = tid >> 3, lc = (tid & 7) * 4;          // staging: row k = lr, columns lc .. lc + 3
  const bool a_vec = (P.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(P.A) & 15u) == 0 && m0 + 32 <= P.M;
  const bool b_vec = (P.ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(P.B) & 15u) == 0 && n0 + 32 <= P.N;

  auto load = [&](const float* X, int ld, int lim, int c0, bool vec, int k) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const float* r = X + (long long)k * ld + c0 + lc;
      if (vec) v = *reinterpret_cast<const float4*>(r);
      else {
        if (c0 + lc < lim) v.x = r[0];
        if (c0 + lc + 1 < lim) v.y = r[1];
        if (c0 + lc + 2 < lim) v.z = r[2];
        if (c0 + lc + 3 < lim) v.w = r[3];
      }
    }
    return v;
  };

  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float cs0 = 0.f, cs1 = 0.f;
  float4 ra = load(P.A, P.lda, P.M, m0, a_vec, lr), rb = load(P.B, P.ldb, P.N, n0, b_vec, lr);
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = rb;
    __syncthreads();
    if (k0 + 32 < K) {
      ra = load(P.A, P.lda, P.M, m0, a_vec, k0 + 32 + lr);
      rb = load(P.B, P.ldb, P.N, n0, b_vec, k0 + 32 + lr);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float2 a = *reinterpret_cast<const float2*>(&As[buf][k][2 * ty]);
      const float2 b = *reinterpret_cast<const float2*>(&Bs[buf][k][2 * tx]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]);
      acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]);
      acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
      cs0 += a.x;
      cs1 += a.y;
    }
    buf ^= 1;      // the next store goes to the other buffer: one barrier per chunk
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int gm = m0 + 2 * ty + i, gn = n0 + 2 * tx + j;
      if (gm < P.M && gn < P.N) P.C[(long long)gm * P.ldc + gn] = acc[i][j];
    }
  if (P.colsum != nullptr && n0 == 0 && tx == 0) {
    if (m0 + 2 * ty < P.M) P.colsum[m0 + 2 * ty] = cs0;
    if (m0 + 2 * ty + 1 < P.M) P.colsum[m0 + 2 * ty + 1] = cs1;
  }
}

struct AtbBuilder {
  AtbGroup g;
  int tiles = 0;
  AtbBuilder() { g.count = 0; }
  void add(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, float* colsum_out) {
    AtbProblem& p = g.p[g.count++];
    p.A = A; p.B = B; p.C = C; p.colsum = colsum_out;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.M = M; p.N = N;
    p.tile_begin = tiles;
    tiles += cdiv(M, 32) * cdiv(N, 32);
  }
  int launch(int K, cudaStream_t st) {
    atb_group_kernel<<<tiles, 256, 0, st>>>(g, K);
    RN_LAUNCH_CHECK("atb_group_kernel");
    return RN_OK;
  }
};

// ---- strided column sums -----------------------------------------------------------------
// out[(g1*n2 + g2), :] = sum_{i<count} A[(g1*s1 + g2*s2 + i*si), :]   (rows of width N, ld = N)
static __global__ void colsum_kernel(const float* __restrict__ A, float* __restrict__ out, int N, int n2, long long s1,
                              long long s2, long long si, int count) {
  __shared__ float red[8][33];
  const int col = blockIdx.y * 32 + threadIdx.x;
  const int g = blockIdx.x;
  const int g1 = g / n2, g2 = g % n2;
  const float* base = A + (g1 * s1 + g2 * s2) * N;
  float acc = 0.f;
  if (col < N)
    for (int i = threadIdx.y; i < count; i += 8) acc += base[i * si * N + col];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) v += red[y][threadIdx.x];
    out[(long long)g * N + col] = v;
  }
}

// float4 variant for wide rows (N % 4 == 0): block = 64 column groups x 4 row lanes, 1 KB coalesced row segments;
// eight independent loads in flight per thread so that tall single-output sums (bias gradients: one output row,
// hundreds of input rows) are not a chain of exposed L2 round trips.
static __global__ void __launch_bounds__(256)
colsum4_kernel(const float* __restrict__ A, float* __restrict__ out, int N, int n2, long long s1, long long s2,
               long long si, int count) {
  __shared__ float4 red[4][64];
  const int cg = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int col = (blockIdx.y * 64 + cg) * 4;
  const int g = blockIdx.x;
  const int g1 = g / n2, g2 = g % n2;
  const float* base = A + (g1 * s1 + g2 * s2) * N;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N) {
    int i = rl;
    for (; i + 28 < count; i += 32) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(base + (i + 4 * u) * si * N + col);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; i < count; i += 4) {
      const float4 v0 = *reinterpret_cast<const float4*>(base + i * si * N + col);
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
    }
  }
  red[rl][cg] = acc;
  __syncthreads();
  if (rl == 0 && col < N) {
    float4 v = red[0][cg];
#pragma unroll
    for (int y = 1; y < 4; ++y) { v.x += red[y][cg].x; v.y += red[y][cg].y; v.z += red[y][cg].z; v.w += red[y][cg].w; }
    *reinterpret_cast<float4*>(out + (long long)g * N + col) = v;
  }
}

// tall single-output sum (bias gradients): out[:] = sum_{i<count} A[i*si, :].  Block = 32 columns x 32 row lanes
// (1024 threads), so even a 256-column sum spreads over 8 SMs with ~8 K loads in flight each; fixed order.
static __global__ void __launch_bounds__(1024)
colsum_tall_kernel(const float* __restrict__ A, float* __restrict__ out, int N, long long si, int count) {
  __shared__ float red[32][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (col < N) {
    int i = threadIdx.y;
    for (; i + 96 < count; i += 128) {
      const float v0 = A[(long long)i * si * N + col], v1 = A[(long long)(i + 32) * si * N + col];
      const float v2 = A[(long long)(i + 64) * si * N + col], v3 = A[(long long)(i + 96) * si * N + col];
      acc += (v0 + v1) + (v2 + v3);
    }
    for (; i < count; i += 32) acc += A[(long long)i * si * N + col];
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float v = 0.f;
#pragma unroll
    for (int y = 0; y < 32; ++y) v += red[y][threadIdx.x];
    out[col] = v;
  }
}

inline int colsum(const float* A, float* out, int N, int n1, int n2, long long s1, long long s2, long long si,
                  int count, cudaStream_t st) {
  if (n1 * n2 <= 0) return RN_OK;
  if (n1 * n2 == 1 && count >= 64) {
    colsum_tall_kernel<<<cdiv(N, 32), dim3(32, 32), 0, st>>>(A, out, N, si, count);
    RN_LAUNCH_CHECK("colsum_tall_kernel");
    return RN_OK;
  }
  if (N % 4 == 0 && N >= 64 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    colsum4_kernel<<<dim3(n1 * n2, cdiv(N, 256)), 256, 0, st>>>(A, out, N, n2, s1, s2, si, count);
    RN_LAUNCH_CHECK("colsum4_kernel");
    return RN_OK;
  }
  dim3 grid(n1 * n2, cdiv(N, 32)), block(32, 8);
  colsum_kernel<<<grid, block, 0, st>>>(A, out, N, n2, s1, s2, si, count);
  RN_LAUNCH_CHECK("colsum_kernel");
  return RN_OK;
}

}  // namespace rn
