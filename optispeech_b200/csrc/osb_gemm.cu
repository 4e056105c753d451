// osb_gemm.cu — tcgen05 / TMA contraction kernels with fused epilogues.
//
//   gemm_nt_kernel   : out = epi( sum_tap A[b, t+tap-pad, :] . W[tap, n, :] )      (forward, dgrad)
//   gemm_wgrad_kernel: dw[tap, n, k] += sum_{b,t} dy[b,t,n] * a[b, t+tap-pad, k]   (weight gradient)
//
// Data layout in HBM: activations channels-last fp16 (B, T, ld), weights fp16 (taps, N, ld).
// One CTA owns a 128-row output tile; operands are staged by TMA into 128B-swizzled shared
// memory (a ring of full/empty mbarriers), one elected thread issues tcgen05.mma with the fp32
// accumulator in TMEM, and four epilogue warps read it back (one thread per output row, so
// LayerNorm / dot-product epilogues are thread-local).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {

unsigned long long g_launch_count = 0;

namespace {

constexpr int BM = 128;          // rows per CTA tile (UMMA M)
constexpr int ROW_BYTES = 128;   // bytes along the contraction dim per smem row (= swizzle span)
constexpr int BKE = 64;          // fp16 elements per ROW_BYTES
constexpr int UMMA_K_BYTES = 32; // bytes consumed along K by one tcgen05.mma (16 fp16)

struct GemmKParams {
  int T, B, N, K, taps, pad;
  int m_tiles;  // row tiles per batch
  int flags;
  void* out;
  void* aux;
  long long ldo;
  const float* bias;
  const float* resid;
  const float* gamma;
  const float* row_scale;
  const uint8_t* pad_mask;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  const float* dot_w;
  const float* dot_b;
  float* out_dot;
  const void* aux_in;
  const float* row_stat;
  float drop_p;
  float drop_inv_keep;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  int w_lo_slice;   // slice offset of the lo parts of W (split precision, shared weights)
  int w_lo_koff;    // K offset of the lo parts of W (split precision, batched weights stored [hi K | lo K])
  int w_batch;      // 1: slice index += batch (per-batch B operand)
  const long long* col_len;
  float* colsum;    // COLSUM: (N) column sums of the fp16 output
  int row_stride;   // > 1: strided convolution, output row t reads input row t*row_stride + tap - pad
  float lrelu;      // OSB_FLAG_LRELU slope
  int seq_pitch;    // > 0: rows t with t % seq_pitch >= seq_valid count as padded (KEEPMASK without a mask array)
  int seq_valid;
  int w_mn;         // 1: W is (taps, K, ldw >= N) with the OUTPUT index contiguous (MN-major B operand): dgrad on the forward pack
  int tap_rev;      // 1: tap t reads weight slice taps-1-t (transposed convolution)
  long long* trace; // optional (developer): clock64 timeline of CTA (0,0), see tools/probe_gemm_trace.py
};

#define GT_TRACE(idx) \
  do { if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0) p.trace[(idx)] = clock64(); } while (0)

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BN>
struct GemmCfg {
  static_assert(BN % 32 == 0 && BN <= 512, "BN");
  static constexpr int NINST = (BN <= 256) ? BN : BN / 2;  // N of one tcgen05.mma (<= 256, % 16 == 0)
  static constexpr int NCHUNK = BN / NINST;
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = cmin(6, (216 * 1024) / STAGE_BYTES);
  static constexpr int TMEM_COLS = pow2_cols(BN);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int VEC_BYTES = 4 * BN * 4;   // four per-column fp32 vectors (bias / gamma|ln_w / ln_b / dot_w) for the epilogue
  static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + VEC_BYTES;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(NINST % 16 == 0 && NINST <= 256, "NINST");
  static_assert(STAGES >= 2, "stages");
};

// ------------------------------------------------------------------------------------------
// epilogues: one thread == one output row; `taddr` addresses this thread's TMEM lane, col 0.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_chunk(uint32_t taddr, int c0, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld_h16x32(const __half* src, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      v[i + 2 * j] = f.x;
      v[i + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ void st_f32x32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
__device__ __forceinline__ void st_h16x32_lo(__half* dst, const float (&v)[32]) {
  // residual after fp16 rounding: lo = fp16(v - float(fp16(v)))
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = v[i] - __half2float(__float2half_rn(v[i]));
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    __half2 h0 = __floats2half2_rn(r[i], r[i + 1]);
    __half2 h1 = __floats2half2_rn(r[i + 2], r[i + 3]);
    __half2 h2 = __floats2half2_rn(r[i + 4], r[i + 5]);
    __half2 h3 = __floats2half2_rn(r[i + 6], r[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
}
__device__ __forceinline__ void st_h16x32(__half* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    __half2 h0 = __floats2half2_rn(v[i], v[i + 1]);
    __half2 h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
    __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]);
    __half2 h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative row I/O.  TMEM hands every thread one output ROW, but 32 threads storing 16 bytes each to 32 different
// rows is 32 memory transactions per instruction (measured: 7-12 us of a ~10 us GEMM launch were these stores).  Each
// epilogue warp therefore owns a small shared-memory slab (carved from the idle pipeline stages): a thread parks its
// 32-column chunk there, and the warp moves the 32 x 32 block between the slab and global memory row-wise, 4 rows x 128 B
// (fp32) or 8 rows x 64 B (fp16) per instruction.  All helpers are warp-collective; `nrows` = number of valid rows of
// this warp (rows are consecutive, so validity is a prefix).
// ------------------------------------------------------------------------------------------
constexpr int SLAB_F32_LD = 36;                       // floats per slab row (32 + 4: conflict-free float4 columns)
constexpr int SLAB_H16_LD = 40;                       // halves per slab row (fp16 view of the same slab)
constexpr int SLAB_BYTES = 32 * SLAB_F32_LD * 4;      // 4608 B per warp

__device__ __forceinline__ void warp_store_f32(float* sl, float* g, long long ld, int nrows, const float (&v)[32]) {
  const int lane = threadIdx.x & 31;
  float* mine = sl + lane * SLAB_F32_LD;
#pragma unroll
  for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(mine + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  __syncwarp();
  const int rr = lane >> 3, c = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rr;
    if (r < nrows) *reinterpret_cast<float4*>(g + r * ld + c) = *reinterpret_cast<const float4*>(sl + r * SLAB_F32_LD + c);
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_load_f32(float* sl, const float* g, long long ld, int nrows, float (&v)[32]) {
  const int lane = threadIdx.x & 31;
  const int rr = lane >> 3, c = (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 4 * i + rr;
    const float4 x = r < nrows ? *reinterpret_cast<const float4*>(g + r * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sl + r * SLAB_F32_LD + c) = x;
  }
  __syncwarp();
  const float* mine = sl + lane * SLAB_F32_LD;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(mine + i);
    v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_store_h16(float* sl, __half* g, long long ld, int nrows, const float (&v)[32]) {
  const int lane = threadIdx.x & 31;
  __half* sh = reinterpret_cast<__half*>(sl);
  st_h16x32(sh + lane * SLAB_H16_LD, v);
  __syncwarp();
  const int rr = lane >> 2, c = (lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + rr;
    if (r < nrows) *reinterpret_cast<uint4*>(g + r * ld + c) = *reinterpret_cast<const uint4*>(sh + r * SLAB_H16_LD + c);
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_load_h16(float* sl, const __half* g, long long ld, int nrows, float (&v)[32]) {
  const int lane = threadIdx.x & 31;
  __half* sh = reinterpret_cast<__half*>(sl);
  const int rr = lane >> 2, c = (lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + rr;
    const uint4 x = r < nrows ? *reinterpret_cast<const uint4*>(g + r * ld + c) : make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sh + r * SLAB_H16_LD + c) = x;
  }
  __syncwarp();
  ld_h16x32(sh + lane * SLAB_H16_LD, v);
  __syncwarp();
}

// warp_store_h16 that also adds the column sums of the (fp16-rounded, valid) rows into colsum[0..32): lane l owns column l.
__device__ __forceinline__ void warp_store_h16_colsum(float* sl, __half* g, long long ld, int nrows, const float (&v)[32], float* colsum) {
  const int lane = threadIdx.x & 31;
  __half* sh = reinterpret_cast<__half*>(sl);
  st_h16x32(sh + lane * SLAB_H16_LD, v);
  __syncwarp();
  const int rr = lane >> 2, c = (lane & 3) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + rr;
    if (r < nrows) *reinterpret_cast<uint4*>(g + r * ld + c) = *reinterpret_cast<const uint4*>(sh + r * SLAB_H16_LD + c);
  }
  float s = 0.f;
  for (int r = 0; r < nrows; ++r) s += __half2float(sh[r * SLAB_H16_LD + lane]);
  if (nrows > 0) atomicAdd(colsum + lane, s);
  __syncwarp();
}

// fp16 destination, this warp's rows, columns [n, n+32): plain (stride ldo) or split [hi ldo | lo ldo] (stride 2*ldo)
__device__ __forceinline__ void warp_store_h(const GemmKParams& p, float* sl, void* base, long long row0, int n, int nrows,
                                             const float (&v)[32]) {
  __half* hb = static_cast<__half*>(base);
  if (p.flags & OSB_FLAG_SPLIT_OUT) {
    warp_store_h16(sl, hb + row0 * (2 * p.ldo) + n, 2 * p.ldo, nrows, v);
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = v[i] - __half2float(__float2half_rn(v[i]));   // residual after fp16 rounding
    warp_store_h16(sl, hb + row0 * (2 * p.ldo) + p.ldo + n, 2 * p.ldo, nrows, r);
  } else {
    warp_store_h16(sl, hb + row0 * p.ldo + n, p.ldo, nrows, v);
  }
}

__device__ __forceinline__ void add_vec(const float* g, float (&v)[32]) {   // v[i] += g[i], 16-byte uniform loads
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(g + i));
    v[i] += x.x; v[i + 1] += x.y; v[i + 2] += x.z; v[i + 3] += x.w;
  }
}
__device__ __forceinline__ void ld_vec(const float* g, float (&w)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(g + i));
    w[i] = x.x; w[i + 1] = x.y; w[i + 2] = x.z; w[i + 3] = x.w;
  }
}

// Row-wise epilogue inputs (residual rows, saved activations): each warp bulk-loads ITS 32 rows x ncols into a warp-private
// shared-memory tile with every load in flight at once (one exposed memory latency per tile instead of one per 32-column
// chunk), then each thread reads its own row from shared memory.  Row pitch = ncols + 4 floats / ncols + 8 halves:
// 16-byte row reads by 8 consecutive lanes hit 32 distinct banks.
template <int NCOLS>
__device__ __forceinline__ void warp_tile_load_f32(float* tile, const float* g, long long ld, int nrows) {
  const int lane = threadIdx.x & 31;
  constexpr int PER_ROW = NCOLS / 4;                // float4 per row
  constexpr int ITERS = PER_ROW;                    // 32 rows * PER_ROW / 32 lanes
  constexpr int BATCH = 8;                          // independent 16-byte loads in flight per lane
  static_assert(ITERS % BATCH == 0, "tile width");
#pragma unroll 1
  for (int base = 0; base < ITERS; base += BATCH) {
    float4 x[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = lane + 32 * (base + j);
      const int r = i / PER_ROW, c = (i % PER_ROW) * 4;
      x[j] = r < nrows ? *reinterpret_cast<const float4*>(g + r * ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = lane + 32 * (base + j);
      const int r = i / PER_ROW, c = (i % PER_ROW) * 4;
      *reinterpret_cast<float4*>(tile + r * (NCOLS + 4) + c) = x[j];
    }
  }
  __syncwarp();
}
template <int NCOLS>
__device__ __forceinline__ void warp_tile_load_h16(__half* tile, const __half* g, long long ld, int nrows) {
  const int lane = threadIdx.x & 31;
  constexpr int PER_ROW = NCOLS / 8;                // uint4 (8 halves) per row
  constexpr int ITERS = PER_ROW;
  constexpr int BATCH = 8;
  static_assert(ITERS % BATCH == 0, "tile width");
#pragma unroll 1
  for (int base = 0; base < ITERS; base += BATCH) {
    uint4 x[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = lane + 32 * (base + j);
      const int r = i / PER_ROW, c = (i % PER_ROW) * 8;
      x[j] = r < nrows ? *reinterpret_cast<const uint4*>(g + r * ld + c) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      const int i = lane + 32 * (base + j);
      const int r = i / PER_ROW, c = (i % PER_ROW) * 8;
      *reinterpret_cast<uint4*>(tile + r * (NCOLS + 8) + c) = x[j];
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void ld_f32x32(const float* src, float (&v)[32]) {   // shared or global, 16-byte aligned
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(src + i);
    v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
  }
}
__device__ __forceinline__ void add_f32x32(const float* src, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 x = *reinterpret_cast<const float4*>(src + i);
    v[i] += x.x; v[i + 1] += x.y; v[i + 2] += x.z; v[i + 3] += x.w;
  }
}

// `taddr`: this thread's TMEM lane, column 0.  `t`: this thread's row inside batch b; `tw`: first row of this warp;
// `sl`: this warp's slab; `sv`: four per-column vectors of BN floats in shared memory (bias | gamma or ln_w | ln_b | dot_w,
// preloaded while the main loop ran); `tile`: this warp's input-tile area.  Every branch is warp-uniform (the helpers are
// warp-collective).  The accumulator is complete (and the pipeline memory free) once `acc_ready` has fired.
template <int EPI, int BN>
__device__ __forceinline__ void run_epilogue(const GemmKParams& p, uint32_t taddr, int b, int t, int tw, int n0, float* sl,
                                             const float* sv, uint8_t* tile, uint64_t* acc_ready) {
  using Cfg = GemmCfg<BN>;
  constexpr bool kTileF32 = 4 * SLAB_BYTES + 128 * (BN + 4) * 4 <= Cfg::PIPE_BYTES;
  constexpr int TLD32 = BN + 4, TLD16 = BN + 8;
  const int lane = threadIdx.x & 31;
  unsigned long long drop_seed = p.drop_seed;
  if (p.drop_p > 0.f && p.drop_seed_dev != nullptr) drop_seed += *p.drop_seed_dev;
  const bool valid = t < p.T;
  const int nrows = p.T - tw < 0 ? 0 : (p.T - tw > 32 ? 32 : p.T - tw);
  const long long row = static_cast<long long>(b) * p.T + t;
  const long long row0 = static_cast<long long>(b) * p.T + tw;
  const bool padded = valid && (p.pad_mask != nullptr ? (p.pad_mask[row] != 0)
                                                      : (p.seq_pitch > 0 && static_cast<int>(row % p.seq_pitch) >= p.seq_valid));
  const float* sv0 = sv;
  const float* sv1 = sv + BN;
  const float* sv2 = sv + 2 * BN;
  const float* sv3 = sv + 3 * BN;
  float* tile32 = reinterpret_cast<float*>(tile);
  __half* tile16 = reinterpret_cast<__half*>(tile);
  float v[32];

  mbar_wait(acc_ready, 0);
  tc_fence_after_sync();

  if constexpr (EPI == OSB_EPI_ATTN_LOGP) {
    // the CTA owns whole rows of the (Tm x Tx) attention: distance, masked log-softmax and prior are thread-local
    const int ncols = p.N;                                              // true number of columns (<= BN)
    const int nvalid = static_cast<int>(p.col_len[b]) < ncols ? static_cast<int>(p.col_len[b]) : ncols;
    const float nf = valid ? p.row_stat[row] : 0.f;
    const bool vec = (p.ldo & 3) == 0;
    const int ntile = vec ? (ncols & ~31) : 0;                           // whole 32-column chunks of 16-byte aligned rows
    const bool tiled = kTileF32 && ntile == BN;   // whole rows of BN columns (the usual case: Tx a multiple of 64)
    if (tiled) warp_tile_load_f32<BN>(tile32, p.resid + row0 * p.ldo, p.ldo, nrows);
    // pass 1: score = -distance, parked back in TMEM; running max / sum of exp (online softmax: one pass for both)
    float mx = -INFINITY, sum = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = c0 + i;
        v[i] = n < nvalid ? -sqrtf(fmaxf(nf + sv0[n] - 2.f * v[i], 0.f)) : -INFINITY;
        cm = fmaxf(cm, v[i]);
      }
      if (cm > mx) {
        sum *= __expf(mx - cm);  // exp(-inf) = 0 on the first chunk
        mx = cm;
      }
      float cs = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) cs += v[i] > -INFINITY ? __expf(v[i] - mx) : 0.f;  // also keeps an all-masked row NaN-free
      sum += cs;
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
      tmem_st_32x32(taddr + static_cast<uint32_t>(c0), r);
    }
    tmem_st_wait();
    const float lse = mx + logf(sum);
    if (valid && p.out_dot != nullptr) p.out_dot[row] = lse;
    // pass 2: log-probability + prior
    float pr[32];
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (c0 >= ncols) break;
      ld_chunk(taddr, c0, v);
      if (c0 + 32 <= ntile) {
        if (tiled) ld_f32x32(tile32 + lane * TLD32 + c0, pr);
        else warp_load_f32(sl, p.resid + row0 * p.ldo + c0, p.ldo, nrows, pr);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = c0 + i < nvalid ? (v[i] - lse) + pr[i] : -INFINITY;
        warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + c0, p.ldo, nrows, v);
      } else if (valid) {
        float* orow = static_cast<float*>(p.out) + row * p.ldo;
        const float* prow = p.resid + row * p.ldo;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int n = c0 + i;
          if (n < ncols) orow[n] = n < nvalid ? (v[i] - lse) + prow[n] : -INFINITY;
        }
      }
    }
  } else if constexpr (EPI == OSB_EPI_AXPY) {
    const float alpha = valid ? p.row_stat[row] : 0.f;
    float r[32];
    if constexpr (kTileF32) warp_tile_load_f32<BN>(tile32, p.resid + row0 * p.ldo + n0, p.ldo, nrows);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      const int n = n0 + c0;
      if constexpr (kTileF32) ld_f32x32(tile32 + lane * TLD32 + c0, r);
      else warp_load_f32(sl, p.resid + row0 * p.ldo + n, p.ldo, nrows, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(alpha, r[i], v[i]);
      warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + n, p.ldo, nrows, v);
    }
  } else if constexpr (EPI == OSB_EPI_GELU_BWD || EPI == OSB_EPI_RELU_BWD) {
    float pre[32];
    const float gate = p.drop_p > 0.f ? p.drop_inv_keep : 1.f;   // RELU_BWD: aux_in = relu output AFTER dropout (dropped = 0)
    warp_tile_load_h16<BN>(tile16, static_cast<const __half*>(p.aux_in) + row0 * p.ldo + n0, p.ldo, nrows);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      const int n = n0 + c0;
      ld_h16x32(tile16 + lane * TLD16 + c0, pre);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if constexpr (EPI == OSB_EPI_GELU_BWD) v[i] *= gelu_erf_grad(pre[i]);
        else v[i] = pre[i] > 0.f ? v[i] * gate : 0.f;
      }
      if (p.colsum != nullptr) warp_store_h16_colsum(sl, static_cast<__half*>(p.out) + row0 * p.ldo + n, p.ldo, nrows, v, p.colsum + n);
      else warp_store_h16(sl, static_cast<__half*>(p.out) + row0 * p.ldo + n, p.ldo, nrows, v);
    }
  } else if constexpr (EPI == OSB_EPI_LN_BWD) {
    // acc = d(xhat); out = (acc - mean(acc) - xhat * mean(acc * xhat)) * rstd       (rows are independent)
    float xh[32];
    const float inv_n = 1.f / static_cast<float>(BN);
    warp_tile_load_h16<BN>(tile16, static_cast<const __half*>(p.aux_in) + row0 * p.ldo, p.ldo, nrows);
    const __half* xrow = tile16 + lane * TLD16;
    float s1 = 0.f, s2 = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(xrow + c0, xh);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s1 += v[i];
        s2 = fmaf(v[i], xh[i], s2);
      }
    }
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    const float rstd = valid ? p.row_stat[row] : 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(xrow + c0, xh);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (v[i] - m1 - xh[i] * m2) * rstd;
      warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + c0, p.ldo, nrows, v);
    }
  } else if constexpr (EPI == OSB_EPI_RELU_LN_BWD) {
    // recompute the forward LN statistics of r = relu(conv) (fp16-saved), then LN backward and the ReLU gate
    float r[32];
    const float inv_n = 1.f / static_cast<float>(BN);
    warp_tile_load_h16<BN>(tile16, static_cast<const __half*>(p.aux_in) + row0 * p.ldo, p.ldo, nrows);
    const __half* rrow = tile16 + lane * TLD16;
    float s = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_h16x32(rrow + c0, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) s += r[i];
    }
    const float mean = s * inv_n;
    float q = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_h16x32(rrow + c0, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) q = fmaf(r[i] - mean, r[i] - mean, q);
    }
    const float rstd = rsqrtf(q * inv_n + p.ln_eps);
    float s1 = 0.f, s2 = 0.f;
    const unsigned long long dbase = static_cast<unsigned long long>(valid ? row : 0) * BN;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(rrow + c0, r);
      if (p.drop_p > 0.f) {
        dropout_apply(v, drop_seed, dbase + c0, p.drop_p, p.drop_inv_keep);
      }
      if (p.flags & OSB_FLAG_OUT_H16) warp_store_h16(sl, static_cast<__half*>(p.aux) + row0 * p.ldo + c0, p.ldo, nrows, v);
      float lw[32];
      ld_f32x32(sv1 + c0, lw);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float g = v[i] * lw[i];
        s1 += g;
        s2 = fmaf(g, (r[i] - mean) * rstd, s2);
      }
    }
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(rrow + c0, r);
      if (p.drop_p > 0.f) {
        dropout_apply(v, drop_seed, dbase + c0, p.drop_p, p.drop_inv_keep);
      }
      float lw[32];
      ld_f32x32(sv1 + c0, lw);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float g = v[i] * lw[i];
        const float d = (g - m1 - (r[i] - mean) * rstd * m2) * rstd;
        v[i] = r[i] > 0.f ? d : 0.f;
      }
      if (p.colsum != nullptr) warp_store_h16_colsum(sl, static_cast<__half*>(p.out) + row0 * p.ldo + c0, p.ldo, nrows, v, p.colsum + c0);
      else warp_store_h16(sl, static_cast<__half*>(p.out) + row0 * p.ldo + c0, p.ldo, nrows, v);
    }
  } else if constexpr (EPI == OSB_EPI_BIAS || EPI == OSB_EPI_GELU || EPI == OSB_EPI_RELU || EPI == OSB_EPI_RESID) {
    const float keep = ((p.flags & OSB_FLAG_KEEPMASK) && padded) ? 0.f : 1.f;
    const float rs = (EPI == OSB_EPI_RESID && p.row_scale != nullptr) ? p.row_scale[b] : 1.f;
    if constexpr (EPI == OSB_EPI_RESID && kTileF32) warp_tile_load_f32<BN>(tile32, p.resid + row0 * p.ldo + n0, p.ldo, nrows);
    // the TMEM load of chunk c+1 is in flight while chunk c is processed (two register buffers)
    auto process = [&](int c0) {
      const int n = n0 + c0;
      add_f32x32(sv0 + c0, v);                      // bias (zeros when absent)
      if constexpr (EPI == OSB_EPI_BIAS) {
        if (p.flags & OSB_FLAG_CLIP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], -1.f), 1.f);
        }
        if (p.flags & OSB_FLAG_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (p.flags & OSB_FLAG_LRELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.lrelu;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= keep;
        if (!(p.flags & OSB_FLAG_NO_F32)) warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + n, p.ldo, nrows, v);
        if (p.flags & OSB_FLAG_OUT_H16) warp_store_h(p, sl, p.aux, row0, n, nrows, v);
      } else if constexpr (EPI == OSB_EPI_GELU) {
        if (p.flags & OSB_FLAG_SAVE_PRE) warp_store_h16(sl, static_cast<__half*>(p.aux) + row0 * p.ldo + n, p.ldo, nrows, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        warp_store_h(p, sl, p.out, row0, n, nrows, v);
      } else if constexpr (EPI == OSB_EPI_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        if (p.drop_p > 0.f) {  // Dropout after the ReLU (MultiLayeredConv1d, multi_layer_conv.py:60-62)
          const unsigned long long dbase = static_cast<unsigned long long>(row) * p.N + n;
          dropout_apply(v, drop_seed, dbase, p.drop_p, p.drop_inv_keep);
        }
        warp_store_h(p, sl, p.out, row0, n, nrows, v);
      } else {  // RESID
        if (p.flags & OSB_FLAG_SAVE_PRE) warp_store_h16(sl, static_cast<__half*>(p.aux) + row0 * p.ldo + n, p.ldo, nrows, v);
        if (p.drop_p > 0.f) {  // element dropout on the branch before the residual add (EncoderLayer, encoder_layer.py:103,111)
          const unsigned long long dbase = static_cast<unsigned long long>(row) * p.N + n;
          dropout_apply(v, drop_seed, dbase, p.drop_p, p.drop_inv_keep);
        }
        float r[32];
        if constexpr (kTileF32) ld_f32x32(tile32 + lane * TLD32 + c0, r);
        else warp_load_f32(sl, p.resid + row0 * p.ldo + n, p.ldo, nrows, r);
        float gm[32];
        ld_f32x32(sv1 + c0, gm);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (r[i] + gm[i] * v[i] * rs) * keep;
        warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + n, p.ldo, nrows, v);
        if (p.flags & OSB_FLAG_OUT_H16) warp_store_h(p, sl, p.aux, row0, n, nrows, v);
      }
    };
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(taddr, ra);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
      tmem_ld_wait_regs(ra);
      if (c0 + 32 < BN) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 32), rb);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]);
      process(c0);
      if (c0 + 32 < BN) {
        tmem_ld_wait_regs(rb);
        if (c0 + 64 < BN) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 64), ra);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
        process(c0 + 32);
      }
    }
  } else {  // row-wise LayerNorm epilogues; the CTA owns the whole row (BN == N, n0 == 0)
    constexpr bool kRelu = (EPI == OSB_EPI_RELU_LN);
    const float inv_n = 1.f / static_cast<float>(BN);
    // TMEM reads are 64 B/clk per SM (a 128 x 256 fp32 accumulator takes >= 2048 cycles per sweep): when the row fits, the
    // pre-LN values are parked ONCE in this thread's row of the shared-memory tile and the later sweeps read them from there.
    float* myrow = tile32 + lane * TLD32;
    // pass 1: mean
    float s = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      add_f32x32(sv0 + c0, v);
      if (kRelu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) s += v[i];
      if constexpr (kTileF32) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(myrow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
    const float mean = s * inv_n;
    auto load_pre = [&](int c0) {   // pre-LN values of this row, columns [c0, c0 + 32)
      if constexpr (kTileF32) {
        ld_f32x32(myrow + c0, v);
      } else {
        ld_chunk(taddr, c0, v);
        add_f32x32(sv0 + c0, v);
        if (kRelu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
      }
    };
    // pass 2: biased variance around the mean (what at::layer_norm computes)
    float q = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      load_pre(c0);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float d = v[i] - mean;
        q += d * d;
      }
    }
    const float rstd = rsqrtf(q * inv_n + p.ln_eps);
    // pass 3: normalise, affine, store
    float dot = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      load_pre(c0);
      if (p.flags & OSB_FLAG_SAVE_PRE) warp_store_h16(sl, static_cast<__half*>(p.aux) + row0 * p.ldo + c0, p.ldo, nrows, v);
      {
        float lw[32], lb[32];
        ld_f32x32(sv1 + c0, lw);
        ld_f32x32(sv2 + c0, lb);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * lw[i] + lb[i];
      }
      if (kRelu && p.drop_p > 0.f) {
        const unsigned long long base = static_cast<unsigned long long>(valid ? row : 0) * BN + c0;
        dropout_apply(v, drop_seed, base, p.drop_p, p.drop_inv_keep);
      }
      if (p.flags & OSB_FLAG_DOT) {
        float dw[32];
        ld_f32x32(sv3 + c0, dw);
#pragma unroll
        for (int i = 0; i < 32; ++i) dot += v[i] * dw[i];
      }
      if constexpr (EPI == OSB_EPI_RELU_LN) {
        if (p.out != nullptr) warp_store_h(p, sl, p.out, row0, c0, nrows, v);
      } else {
        warp_store_f32(sl, static_cast<float*>(p.out) + row0 * p.ldo + c0, p.ldo, nrows, v);
        if (p.flags & OSB_FLAG_OUT_H16) warp_store_h(p, sl, p.aux, row0, c0, nrows, v);
      }
    }
    if ((p.flags & OSB_FLAG_DOT) && valid) p.out_dot[row] = padded ? 0.f : (dot + (p.dot_b != nullptr ? __ldg(p.dot_b) : 0.f));
  }
}

// ------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) GT_TRACE(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GT_TRACE(1);

  const int b = blockIdx.x / p.m_tiles;
  const int t0 = (blockIdx.x % p.m_tiles) * BM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = (p.K + BKE - 1) / BKE;
  // split precision: three products per k-block — (a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo).  The operands travel as TWO pipeline
  // stages per k-block, (a_hi, w_hi) and (a_lo, w_lo), and the MMA warp forms the three products from the pair: 2/3 of the
  // shared-memory fills of three separate passes (the main loop of these launches is bound by what one SM pulls from L2).
  // With fewer than four stages (BN = 384: 64 KB per stage) a pair would leave 1.5 k-blocks in flight; those tiles keep three
  // separate passes (pass-major order).
  const bool split_in = (p.flags & OSB_FLAG_SPLIT_IN) != 0;
  constexpr bool kPair = Cfg::STAGES >= 4;
  const int nsub = split_in ? (kPair ? 2 : 3) : 1;
  const int iters = p.taps * nsub * num_kb;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        const int tap = it / (nsub * num_kb);
        const int rem = it - tap * (nsub * num_kb);
        int kb, a_lo, w_lo;
        if (kPair || !split_in) {     // k-block major: (hi, hi) then (lo, lo)
          kb = rem / nsub;
          a_lo = w_lo = (rem - kb * nsub) == 1;
        } else {                      // pass major: (a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo)
          const int sub = rem / num_kb;
          kb = rem - sub * num_kb;
          a_lo = sub == 1;
          w_lo = sub == 2;
        }
        const int a_k = kb * BKE + (a_lo ? p.K : 0);
        const int w_slice = tap + (w_lo ? p.w_lo_slice : 0) + (p.w_batch ? b : 0);
        const int w_k = kb * BKE + (w_lo ? p.w_lo_koff : 0);
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        if (p.row_stride > 1) {
          // strided rows: a TMA box is at most 256 SOURCE rows, i.e. 64 loaded rows at stride <= 4: two boxes per 128-row tile
          tma_load_3d(sA, &tmA, &full_bar[s], a_k, t0 * p.row_stride + tap - p.pad, b);
          tma_load_3d(sA + 64 * ROW_BYTES, &tmA, &full_bar[s], a_k, (t0 + 64) * p.row_stride + tap - p.pad, b);
        } else {
          tma_load_3d(sA, &tmA, &full_bar[s], a_k, t0 + tap - p.pad, b);
        }
        if (p.w_mn) {
          // MN-major B: 64 contraction rows x BN output columns, as BN/64 chunks of [64 rows x 128 B]
          const int slice = p.tap_rev ? p.taps - 1 - tap : tap;
#pragma unroll
          for (int c = 0; c < BN / 64; ++c) tma_load_3d(sB + c * (BKE * ROW_BYTES), &tmW, &full_bar[s], n0 + c * 64, kb * BKE, slice);
        } else {
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c)
            tma_load_3d(sB + c * Cfg::NINST * ROW_BYTES, &tmW, &full_bar[s], w_k, n0 + c * Cfg::NINST, w_slice);
        }
        if (it == 0) GT_TRACE(2);
        if (it == iters - 1) GT_TRACE(3);
      }
    }
  } else if (warp == 1) {
    // whole warp, uniform control flow (descriptors live in uniform registers); one elected lane issues
    const uint32_t idesc = make_instr_desc(OSB_F16, BM, Cfg::NINST, 0, p.w_mn ? 1u : 0u);
    const uint64_t d0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t d0_mn = make_smem_desc_sw128(smem_u32(smem), BKE * ROW_BYTES, 1024);   // LBO = stride between 64-column chunks
    if (kPair && split_in) {
      // pairs of stages: s0 = (a_hi, w_hi), s1 = (a_lo, w_lo) of one k-block
      for (int it = 0; it < iters; it += 2) {
        const int s0 = it % Cfg::STAGES, s1 = (it + 1) % Cfg::STAGES;
        mbar_wait(&full_bar[s0], (it / Cfg::STAGES) & 1);
        mbar_wait(&full_bar[s1], ((it + 1) / Cfg::STAGES) & 1);
        tc_fence_after_sync();
        if (it == 0 && lane == 0) GT_TRACE(4);
        if (elect_one()) {
          const uint64_t a_hi = d0 + static_cast<uint64_t>((s0 * Cfg::STAGE_BYTES) >> 4);
          const uint64_t a_lo = d0 + static_cast<uint64_t>((s1 * Cfg::STAGE_BYTES) >> 4);
          const uint64_t w_hi = a_hi + static_cast<uint64_t>(Cfg::A_BYTES >> 4);
          const uint64_t w_lo = a_lo + static_cast<uint64_t>(Cfg::A_BYTES >> 4);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const uint64_t da0 = g == 1 ? a_lo : a_hi;
            const uint64_t db0 = g == 2 ? w_lo : w_hi;
#pragma unroll
            for (int k = 0; k < ROW_BYTES / UMMA_K_BYTES; ++k) {
#pragma unroll
              for (int c = 0; c < Cfg::NCHUNK; ++c) {
                umma_ss<false>(tmem_base + c * Cfg::NINST, da0 + static_cast<uint64_t>((k * UMMA_K_BYTES) >> 4),
                               db0 + static_cast<uint64_t>((c * Cfg::NINST * ROW_BYTES + k * UMMA_K_BYTES) >> 4), idesc,
                               (it | g | k) != 0 ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty_bar[s0]);
          umma_commit(&empty_bar[s1]);
        }
        __syncwarp();
      }
    } else
    for (int it = 0; it < iters; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after_sync();
      if (it == 0 && lane == 0) GT_TRACE(4);
      if (elect_one()) {
        const uint64_t da0 = d0 + static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
        const uint64_t db0 = da0 + static_cast<uint64_t>(Cfg::A_BYTES >> 4);
        if (p.w_mn) {
          const uint64_t db_mn = d0_mn + static_cast<uint64_t>((s * Cfg::STAGE_BYTES + Cfg::A_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < ROW_BYTES / UMMA_K_BYTES; ++k) {
#pragma unroll
            for (int c = 0; c < Cfg::NCHUNK; ++c) {
              // 16 contraction rows per MMA = 2048 B inside every chunk; instruction c starts NINST/64 chunks further
              umma_ss<false>(tmem_base + c * Cfg::NINST, da0 + static_cast<uint64_t>((k * UMMA_K_BYTES) >> 4),
                             db_mn + static_cast<uint64_t>((c * (Cfg::NINST / 64) * (BKE * ROW_BYTES) + k * 16 * ROW_BYTES) >> 4), idesc,
                             (it | k) != 0 ? 1u : 0u);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < ROW_BYTES / UMMA_K_BYTES; ++k) {
#pragma unroll
            for (int c = 0; c < Cfg::NCHUNK; ++c) {
              umma_ss<false>(tmem_base + c * Cfg::NINST, da0 + static_cast<uint64_t>((k * UMMA_K_BYTES) >> 4),
                             db0 + static_cast<uint64_t>((c * Cfg::NINST * ROW_BYTES + k * UMMA_K_BYTES) >> 4), idesc,
                             (it | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(tmem_full_bar);  // accumulator complete
    __syncwarp();
    if (lane == 0) GT_TRACE(5);
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    // While the main loop runs, the (otherwise idle) epilogue warps stage the per-column vectors in shared memory:
    // [0] bias (ATTN: |e_n|^2 of this batch), [1] gamma / ln_w, [2] ln_b, [3] dot_w — zeros where absent.
    float* sv = reinterpret_cast<float*>(smem + Cfg::PIPE_BYTES + 256);
    {
      const int et = threadIdx.x - 64;   // 0..127
      const float* v0 = p.bias;
      int lim0 = BN;
      if constexpr (EPI == OSB_EPI_ATTN_LOGP) { v0 = p.bias + static_cast<long long>(b) * p.N; lim0 = p.N; }
      const float* v1 = (EPI == OSB_EPI_RESID) ? p.gamma : p.ln_w;
      for (int i = et; i < BN; i += 128) {
        sv[i] = (v0 != nullptr && i < lim0) ? __ldg(v0 + (EPI == OSB_EPI_ATTN_LOGP ? 0 : n0) + i) : 0.f;
        sv[BN + i] = v1 != nullptr ? __ldg(v1 + n0 + i) : 0.f;
        sv[2 * BN + i] = p.ln_b != nullptr ? __ldg(p.ln_b + n0 + i) : 0.f;
        sv[3 * BN + i] = p.dot_w != nullptr ? __ldg(p.dot_w + n0 + i) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only
    }
    if (warp == 2 && lane == 0) GT_TRACE(6);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // once the accumulator is complete every TMA stage has been consumed: the pipeline memory then holds the per-warp slabs
    // and input tiles (run_epilogue waits on tmem_full_bar before touching either)
    float* slab = reinterpret_cast<float*>(smem + q * SLAB_BYTES);
    constexpr int TILE_WARP_BYTES = (Cfg::PIPE_BYTES - 4 * SLAB_BYTES) / 4 / 16 * 16;
    uint8_t* tile = smem + 4 * SLAB_BYTES + q * TILE_WARP_BYTES;
    run_epilogue<EPI, BN>(p, taddr, b, t0 + q * 32 + lane, t0 + q * 32, n0, slab, sv, tile, tmem_full_bar);
    if (warp == 2 && lane == 0) GT_TRACE(7);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) GT_TRACE(8);
}

// ------------------------------------------------------------------------------------------
// weight-gradient kernel: both operands MN-major (read straight from channels-last tensors)
// ------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 64;                      // contraction rows (b,t) per stage
constexpr int WG_TILE = 128;                     // output tile is 128 (n) x 128 (k)
constexpr int WG_OP_BYTES = WG_ROWS * 256;       // 64 rows x 128 elems x 2 B, stored as 2 chunks of 64 elems
constexpr int WG_STAGE_BYTES = 2 * WG_OP_BYTES;  // dy tile + a tile
constexpr int WG_STAGES = 6;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgradParams {
  int T, B, N, K, taps, pad;
  int row_blocks_per_batch;  // ceil(T / 64)
  int splits;
  int per_batch;             // 1: one (N, K) output per batch (batched matmul A^T B), taps == 1
  int a_stride;              // > 1: strided convolution — output row t pairs with input row t * a_stride + tap - pad
  float* dw;
};

__global__ void __launch_bounds__(192, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmA, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n0 = blockIdx.x * WG_TILE;
  const int k0 = blockIdx.y * WG_TILE;
  const int zsel = blockIdx.z / p.splits;       // tap, or batch in per-batch mode
  const int tap = p.per_batch ? 0 : zsel;
  const int split = blockIdx.z % p.splits;
  const int total_rb = p.per_batch ? p.row_blocks_per_batch : p.B * p.row_blocks_per_batch;
  const int rb_base = p.per_batch ? zsel * p.row_blocks_per_batch : 0;
  const int rb_begin = rb_base + static_cast<int>((static_cast<long long>(total_rb) * split) / p.splits);
  const int rb_end = rb_base + static_cast<int>((static_cast<long long>(total_rb) * (split + 1)) / p.splits);
  const int iters = rb_end - rb_begin;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], WG_STAGE_BYTES);
        const int rb = rb_begin + it;
        const int b = rb / p.row_blocks_per_batch;
        const int t0 = (rb - b * p.row_blocks_per_batch) * WG_ROWS;
        uint8_t* sDy = smem + s * WG_STAGE_BYTES;
        uint8_t* sA = sDy + WG_OP_BYTES;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(sDy + c * (WG_ROWS * ROW_BYTES), &tmDy, &full_bar[s], n0 + c * 64, t0, b);
          tma_load_3d(sA + c * (WG_ROWS * ROW_BYTES), &tmA, &full_bar[s], k0 + c * 64, t0 * p.a_stride + tap - p.pad, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_instr_desc(OSB_F16, WG_TILE, WG_TILE, 1, 1);
      for (int it = 0; it < iters; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t dy_addr = smem_u32(smem + s * WG_STAGE_BYTES);
        const uint32_t a_addr = dy_addr + WG_OP_BYTES;
#pragma unroll
        for (int k = 0; k < WG_ROWS / 16; ++k) {
          // 16 contraction rows = 2 swizzle groups of 8 rows (1024 B each); chunks of 64 elems along MN are
          // WG_ROWS*128 B apart (LBO).
          const uint64_t da = make_smem_desc_sw128(dy_addr + k * 16 * ROW_BYTES, WG_ROWS * ROW_BYTES, 1024);
          const uint64_t db = make_smem_desc_sw128(a_addr + k * 16 * ROW_BYTES, WG_ROWS * ROW_BYTES, 1024);
          umma_ss<false>(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else if (iters > 0) {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // TMEM hands every thread one output row; the reduction into dw wants a warp on one row (512 contiguous bytes per
    // vector red instead of 32 scattered 4-byte atomics).  Transpose the warp's 32 x 128 slab through the (now idle) pipeline
    // stages: row stride 132 floats keeps both the float4 row writes and the float4 row reads conflict-free.
    constexpr int LD = WG_TILE + 4;
    float* slab = reinterpret_cast<float*>(smem) + q * 32 * LD;
    float v[32];
    for (int c0 = 0; c0 < WG_TILE; c0 += 32) {
      ld_chunk(taddr, c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(slab + lane * LD + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    __syncwarp();
    const int kc = k0 + lane * 4;
    if (kc < p.K) {  // K is a multiple of 8: whole float4 groups are in or out
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int n = n0 + q * 32 + r;
        if (n >= p.N) break;
        const float4 x = *reinterpret_cast<const float4*>(slab + r * LD + lane * 4);
        float* dst = p.dw + (static_cast<long long>(zsel) * p.N + n) * p.K + kc;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

template <int BN, int EPI>
int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  dim3 grid(p.B * p.m_tiles, (p.N + BN - 1) / BN, 1);
  gemm_nt_kernel<BN, EPI><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmW, p);
  count_launch();
  return launch_status();
}

template <int EPI>
int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_nt<64, EPI>(tmA, tmW, p, stream);
    case 128: return launch_nt<128, EPI>(tmA, tmW, p, stream);
    case 192: return launch_nt<192, EPI>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, EPI>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}
template <int EPI>
int dispatch_bn_full(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 128: return launch_nt<128, EPI>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, EPI>(tmA, tmW, p, stream);
    case 384: return launch_nt<384, EPI>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}

int dispatch_attn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_nt<64, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 128: return launch_nt<128, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 192: return launch_nt<192, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 320: return launch_nt<320, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 384: return launch_nt<384, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 448: return launch_nt<448, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 512: return launch_nt<512, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}

// Largest supported tile width that divides N.
int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// TMA tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_tmap_3d(CUtensorMap* out, const void* base, TmaDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t elem_stride1) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return OSB_ERR_DRIVER;
  const uint64_t es = (dt == TMA_F32) ? 4 : 2;
  const CUtensorMapDataType cdt = dt == TMA_F16    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                  : dt == TMA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return OSB_ERR_ALIGN;
  if (((stride1_elems * es) & 15) != 0 || ((stride2_elems * es) & 15) != 0) return OSB_ERR_ALIGN;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * es, stride2_elems * es};
  cuuint32_t box[3] = {box0, box1, 1};
  if (elem_stride1 < 1 || elem_stride1 > 8 || box1 > 256) return OSB_ERR_SHAPE;
  cuuint32_t estr[3] = {1, elem_stride1, 1};
  CUresult r = fn(out, cdt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    // A thread that has not made a runtime call yet (e.g. the autograd engine's worker on its first backward node) has no
    // current context for this driver entry point: bind the primary context and retry once.
    cudaFree(nullptr);
    r = fn(out, cdt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? OSB_OK : OSB_ERR_DRIVER;
}

}  // namespace osb

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace osb;

static long long* g_gemm_trace = nullptr;
static int g_gemm_narrow_tiles = 0;
static int g_wgrad_min_iters = 48;
/* developer hook (not in the public header): minimum number of 64-row pipeline iterations per weight-gradient CTA */
extern "C" void osb_debug_set_wgrad_min_iters(int n) { g_wgrad_min_iters = n > 0 ? n : 48; }
/* developer hook (not in the public header): 0 = keep 256-wide tiles for small problems (A/B timing) */
extern "C" void osb_debug_set_gemm_narrow_tiles(int on) { g_gemm_narrow_tiles = on; }
/* developer hook (not in the public header): device buffer of 16 int64 receiving a clock64 timeline of CTA (0,0) */
extern "C" void osb_debug_set_gemm_trace(long long* buf) { g_gemm_trace = buf; }

extern "C" int osb_gemm(const osb_gemm_desc* d, void* stream_) {
  OSB_REQUIRE(d != nullptr && d->a != nullptr && d->w != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(d->B > 0 && d->T > 0 && d->N > 0 && d->K > 0 && d->taps > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(d->lda % 8 == 0 && d->ldw % 8 == 0 && (d->ldo % 8 == 0 || d->epi == OSB_EPI_ATTN_LOGP), OSB_ERR_ALIGN);
  const bool w_mn = (d->flags & OSB_FLAG_W_MN) != 0;
  OSB_REQUIRE(d->lda >= d->K && d->ldw >= (w_mn ? d->N : d->K) && d->ldo >= d->N, OSB_ERR_SHAPE);
  if (w_mn) OSB_REQUIRE(!(d->flags & OSB_FLAG_SPLIT_IN) && d->w_batched == 0 && d->K % 8 == 0, OSB_ERR_SHAPE);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  const bool full_row = (d->epi == OSB_EPI_RELU_LN || d->epi == OSB_EPI_BIAS_LN || d->epi == OSB_EPI_LN_BWD ||
                         d->epi == OSB_EPI_RELU_LN_BWD);
  const bool attn = d->epi == OSB_EPI_ATTN_LOGP;
  int bn = attn ? ((d->N + 63) / 64) * 64 : (full_row ? d->N : pick_bn(d->N));
  // Optional (developer hook, off): halve the tile width when there are few tiles.  The main loop of a small problem is bound
  // by what ONE SM can pull from L2 (A 16 KB + W 32 KB per 64-deep k-block at BN = 256; ~64 B/clk), not by the tensor pipe, so
  // twice the CTAs streaming 32 KB per k-block finish a launch sooner (measured: +1.5 % on the family's serial time) — but inside
  // the multi-stream step the extra SMs are taken from concurrent branches (2.345 -> 2.36 ms/step), so it stays off.
  if (!attn && !full_row && bn == 256 && g_gemm_narrow_tiles) {
    const long long tiles = static_cast<long long>(d->B) * ((d->T + BM - 1) / BM) * (d->N / 256);
    if (tiles * 2 <= 148) bn = 128;
  }
  // Split-precision launches (the synthesis path: one stream, nothing to take SMs from) with few tiles: 64-wide tiles.  A
  // B=1 utterance is 1-2 row tiles, so a 256-wide tile leaves ONE or two SMs pulling the whole hi + lo weight matrix through
  // their ~64 B/clk L2 port (1 MB for a 256 x 1024 pwconv2: 8 us before the first epilogue); 64-wide tiles spread the
  // weight rows over 4x the SMs (the A tile is re-read per column tile, from L2).
  if (!attn && !full_row && (d->flags & OSB_FLAG_SPLIT_IN) && !w_mn && d->N % 64 == 0 && bn > 64) {
    const long long tiles = static_cast<long long>(d->B) * ((d->T + BM - 1) / BM) * (d->N / bn);
    if (tiles * 4 <= 148) bn = 64;
    else if (tiles * 2 <= 148 && d->N % 128 == 0 && bn > 128) bn = 128;
  }
  OSB_REQUIRE(bn > 0 && bn <= 512, OSB_ERR_SHAPE);
  const int ninst = bn <= 256 ? bn : bn / 2;
  const bool w_batched = d->w_batched != 0;
  if (w_batched) OSB_REQUIRE(d->taps == 1, OSB_ERR_SHAPE);

  const bool split_in = (d->flags & OSB_FLAG_SPLIT_IN) != 0;
  if (split_in) OSB_REQUIRE(d->K % BKE == 0 && d->lda >= 2 * static_cast<int64_t>(d->K), OSB_ERR_SHAPE);
  CUtensorMap tmA, tmW;
  const int row_stride = d->row_stride > 1 ? d->row_stride : 1;
  int rc;
  if (row_stride > 1) {
    // strided convolution: the tensor map walks the INPUT rows with a traversal stride; 64 loaded rows per box
    OSB_REQUIRE(row_stride <= 4 && d->T_in > 0 && !split_in && d->epi == OSB_EPI_BIAS, OSB_ERR_SHAPE);
    OSB_REQUIRE(static_cast<int64_t>(d->T - 1) * row_stride + d->taps - 1 - d->pad < d->T_in + d->taps, OSB_ERR_SHAPE);
    rc = make_tmap_3d(&tmA, d->a, TMA_F16, d->K, d->T_in, d->B, d->lda, static_cast<uint64_t>(d->T_in) * d->lda, BKE, 64 * row_stride,
                      row_stride);
  } else {
    rc = make_tmap_3d(&tmA, d->a, TMA_F16, split_in ? 2 * d->K : d->K, d->T, d->B, d->lda, static_cast<uint64_t>(d->T) * d->lda, BKE, BM);
  }
  if (rc != OSB_OK) return rc;
  if (w_mn) {
    // (taps, K, ldw): contraction rows K, output columns N contiguous; box = 64 columns x 64 rows
    OSB_REQUIRE(ninst % 64 == 0, OSB_ERR_SHAPE);
    rc = make_tmap_3d(&tmW, d->w, TMA_F16, d->N, d->K, d->taps, d->ldw, static_cast<uint64_t>(d->K) * d->ldw, 64, BKE);
  } else if (w_batched) {
    if (split_in) OSB_REQUIRE(d->ldw >= 2 * static_cast<int64_t>(d->K), OSB_ERR_SHAPE);
    rc = make_tmap_3d(&tmW, d->w, TMA_F16, split_in ? 2 * d->K : d->K, d->N, d->B, d->ldw, static_cast<uint64_t>(d->N) * d->ldw, BKE,
                      ninst);
  } else {
    rc = make_tmap_3d(&tmW, d->w, TMA_F16, d->K, d->N, split_in ? 2 * d->taps : d->taps, d->ldw, static_cast<uint64_t>(d->N) * d->ldw,
                      BKE, ninst);
  }
  if (rc != OSB_OK) return rc;

  GemmKParams p;
  p.T = d->T; p.B = d->B; p.N = d->N; p.K = d->K; p.taps = d->taps; p.pad = d->pad;
  p.m_tiles = (d->T + BM - 1) / BM;
  p.flags = d->flags;
  p.out = d->out; p.aux = d->aux_h16; p.ldo = d->ldo;
  p.bias = d->bias; p.resid = d->resid; p.gamma = d->gamma; p.row_scale = d->row_scale;
  p.pad_mask = d->pad_mask; p.ln_w = d->ln_w; p.ln_b = d->ln_b; p.ln_eps = d->ln_eps;
  p.dot_w = d->dot_w; p.dot_b = d->dot_b; p.out_dot = d->out_dot;
  p.aux_in = d->aux_in_h16; p.row_stat = d->row_stat;
  p.w_batch = w_batched ? 1 : 0;
  p.w_lo_slice = w_batched ? 0 : d->taps;
  p.w_lo_koff = w_batched ? d->K : 0;
  p.col_len = reinterpret_cast<const long long*>(d->col_len);
  p.row_stride = row_stride;
  p.lrelu = d->lrelu_slope;
  p.seq_pitch = d->seq_pitch > 0 ? d->seq_pitch : 0;
  p.seq_valid = d->seq_valid;
  p.w_mn = w_mn ? 1 : 0;
  p.tap_rev = (d->flags & OSB_FLAG_TAP_REVERSE) ? 1 : 0;
  p.colsum = (d->flags & OSB_FLAG_COLSUM) ? d->out_colsum : nullptr;
  if ((d->flags & OSB_FLAG_COLSUM) && (d->out_colsum == nullptr || !(d->epi == OSB_EPI_GELU_BWD || d->epi == OSB_EPI_RELU_BWD ||
                                                                     d->epi == OSB_EPI_RELU_LN_BWD))) return OSB_ERR_ARG;
  p.trace = g_gemm_trace;
  p.drop_p = d->dropout_p; p.drop_seed = d->dropout_seed; p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(d->dropout_seed_dev);
  p.drop_inv_keep = d->dropout_p > 0.f && d->dropout_p < 1.f ? 1.f / (1.f - d->dropout_p) : 0.f;
  OSB_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f, OSB_ERR_ARG);

  // the epilogues read per-column vectors and move row chunks with 16-byte accesses
  for (const void* q : {static_cast<const void*>(d->gamma), static_cast<const void*>(d->ln_w), static_cast<const void*>(d->ln_b),
                        static_cast<const void*>(d->dot_w), static_cast<const void*>(d->out), static_cast<const void*>(d->aux_h16),
                        static_cast<const void*>(d->aux_in_h16), d->epi == OSB_EPI_ATTN_LOGP ? nullptr : static_cast<const void*>(d->bias),
                        d->epi == OSB_EPI_ATTN_LOGP ? nullptr : static_cast<const void*>(d->resid)})
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return OSB_ERR_ALIGN;
  if ((d->flags & (OSB_FLAG_OUT_H16 | OSB_FLAG_SAVE_PRE)) && d->aux_h16 == nullptr) return OSB_ERR_ARG;
  if ((d->flags & OSB_FLAG_KEEPMASK) && d->pad_mask == nullptr && d->seq_pitch <= 0) return OSB_ERR_ARG;

  switch (d->epi) {
    case OSB_EPI_BIAS:
      OSB_REQUIRE(d->out != nullptr || ((d->flags & OSB_FLAG_NO_F32) && (d->flags & OSB_FLAG_OUT_H16)), OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_BIAS>(bn, tmA, tmW, p, stream);
    case OSB_EPI_GELU:
      OSB_REQUIRE(d->out != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_GELU>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU:
      OSB_REQUIRE(d->out != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RELU>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RESID:
      OSB_REQUIRE(d->out != nullptr && d->resid != nullptr && d->gamma != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RESID>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_LN:
      OSB_REQUIRE(d->ln_w != nullptr && d->ln_b != nullptr, OSB_ERR_ARG);
      OSB_REQUIRE(!(d->flags & OSB_FLAG_DOT) || (d->dot_w != nullptr && d->out_dot != nullptr), OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_RELU_LN>(bn, tmA, tmW, p, stream);
    case OSB_EPI_BIAS_LN:
      OSB_REQUIRE(d->out != nullptr && d->ln_w != nullptr && d->ln_b != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_BIAS_LN>(bn, tmA, tmW, p, stream);
    case OSB_EPI_GELU_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_GELU_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RELU_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_ATTN_LOGP: {
      OSB_REQUIRE(d->out != nullptr && d->row_stat != nullptr && d->bias != nullptr && d->resid != nullptr && d->col_len != nullptr,
                  OSB_ERR_ARG);
      GemmKParams q = p;
      q.N = d->N;  // true column count; the tile is padded to a multiple of 64 (out-of-range rows of W read as zero)
      dim3 grid_check(1, 1, 1);
      (void)grid_check;
      return dispatch_attn(bn, tmA, tmW, q, stream);
    }
    case OSB_EPI_AXPY:
      OSB_REQUIRE(d->out != nullptr && d->row_stat != nullptr && d->resid != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_AXPY>(bn, tmA, tmW, p, stream);
    case OSB_EPI_LN_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr && d->row_stat != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_LN_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_LN_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr && d->ln_w != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_RELU_LN_BWD>(bn, tmA, tmW, p, stream);
    default:
      return OSB_ERR_ARG;
  }
}

static int wgrad_impl(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t N, int32_t K,
                      int32_t taps, int32_t pad, int per_batch, void* stream_, int32_t a_stride = 1, int32_t T_in = 0);

extern "C" int osb_gemm_wgrad(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                              int32_t N, int32_t K, int32_t taps, int32_t pad, void* stream_) {
  return wgrad_impl(dy, ldy, a, lda, dw, B, T, N, K, taps, pad, 0, stream_);
}

extern "C" int osb_gemm_wgrad_batched(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                                      int32_t N, int32_t K, void* stream_) {
  return wgrad_impl(dy, ldy, a, lda, dw, B, T, N, K, 1, 0, 1, stream_);
}

extern "C" int osb_gemm_wgrad_strided(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                                      int32_t T_in, int32_t N, int32_t K, int32_t taps, int32_t pad, int32_t stride, void* stream_) {
  OSB_REQUIRE(stride >= 1 && stride <= 4 && T_in > 0, OSB_ERR_SHAPE);
  return wgrad_impl(dy, ldy, a, lda, dw, B, T, N, K, taps, pad, 0, stream_, stride, T_in);
}

static int wgrad_impl(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t N, int32_t K,
                      int32_t taps, int32_t pad, int per_batch, void* stream_, int32_t a_stride, int32_t T_in) {
  OSB_REQUIRE(dy != nullptr && a != nullptr && dw != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && N > 0 && K > 0 && taps > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(ldy % 8 == 0 && lda % 8 == 0, OSB_ERR_ALIGN);
  // the tile is reduced into dw with 16-byte vector reds
  OSB_REQUIRE(K % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0, OSB_ERR_ALIGN);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tmDy, tmA;
  int rc = make_tmap_3d(&tmDy, dy, TMA_F16, N, T, B, ldy, static_cast<uint64_t>(T) * ldy, 64, WG_ROWS);
  if (rc != OSB_OK) return rc;
  if (a_stride > 1)   // the tensor map walks the INPUT rows (T_in per batch) with a traversal stride: 64 loaded rows per box
    rc = make_tmap_3d(&tmA, a, TMA_F16, K, T_in, B, lda, static_cast<uint64_t>(T_in) * lda, 64, WG_ROWS * a_stride, a_stride);
  else
    rc = make_tmap_3d(&tmA, a, TMA_F16, K, T, B, lda, static_cast<uint64_t>(T) * lda, 64, WG_ROWS);
  if (rc != OSB_OK) return rc;

  WgradParams p;
  p.a_stride = a_stride > 1 ? a_stride : 1;
  p.T = T; p.B = B; p.N = N; p.K = K; p.taps = taps; p.pad = pad;
  p.row_blocks_per_batch = (T + WG_ROWS - 1) / WG_ROWS;
  const int n_tiles = (N + WG_TILE - 1) / WG_TILE;
  const int k_tiles = (K + WG_TILE - 1) / WG_TILE;
  p.per_batch = per_batch;
  const int zcount = per_batch ? B : taps;
  const int total_rb = per_batch ? p.row_blocks_per_batch : B * p.row_blocks_per_batch;
  // Split the contraction rows over CTAs, but keep >= 48 pipeline iterations (64 rows each) per CTA.  Every CTA pays a fixed
  // prologue and flushes a 128x128 fp32 tile with atomics, and in the training step these launches run on side streams beside
  // the data-gradient chain: what counts there is the SM-time a launch takes from its neighbours, not its own latency.  Measured
  // on the B=32 step (6144-row problems): >= 6 iterations -> 0.46 ms of weight-gradient kernels alone, 2.27 ms/step;
  // >= 48 -> 0.69 ms alone, 2.18 ms/step; no split at all -> 1.3 ms alone, 2.26 ms/step.
  int splits = (148 + n_tiles * k_tiles * zcount - 1) / (n_tiles * k_tiles * zcount);
  if (splits > total_rb / g_wgrad_min_iters) splits = total_rb / g_wgrad_min_iters;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.dw = dw;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  dim3 grid(n_tiles, k_tiles, zcount * splits);
  gemm_wgrad_kernel<<<grid, 192, WG_SMEM_BYTES, stream>>>(tmDy, tmA, p);
  count_launch();
  return launch_status();
}

// v2: osb_gemm_desc grew (out_colsum, row_stride, T_in, lrelu_slope), new entry points (mha, pack_multi, losses)
// v3: osb_gemm_desc grew (seq_pitch, seq_valid), osb_pack_job grew (aux), osb_align_loss_fold / osb_ln_dwconv_bwd signatures,
//     new entry points (fused ConvNeXt training kernels, period discriminators, glue kernels)
extern "C" int osb_version(void) { return 3; }
extern "C" unsigned long long osb_launch_count(void) { return osb::g_launch_count; }
extern "C" const char* osb_strerror(int s) {
  switch (s) {
    case OSB_OK: return "ok";
    case OSB_ERR_SHAPE: return "unsupported or inconsistent shape";
    case OSB_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case OSB_ERR_ARCH: return "device is not sm_100 (B200)";
    case OSB_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case OSB_ERR_WORKSPACE: return "workspace too small";
    case OSB_ERR_ARG: return "null pointer or bad enum";
    default: return s > 0 ? "CUDA runtime error (see cudaGetErrorString)" : "unknown status";
  }
}
extern "C" int osb_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return static_cast<int>(e);
  return prop.major == 10 ? OSB_OK : OSB_ERR_ARCH;
}
