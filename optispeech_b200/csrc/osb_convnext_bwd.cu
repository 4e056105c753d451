// osb_convnext_bwd.cu — backward of one ConvNeXt block:
//
//   convnext_bwd_fused_kernel (tcgen05): the two data-gradient contractions of the block in ONE kernel, mirroring the fused
//   forward (osb_convnext.cu).  A CTA owns 128 consecutive positions of one sequence:
//     prologue  : dyg = dout * gamma * keep * rs[b]  (fp32 -> fp16) written straight into 128B-swizzled shared memory — the A
//                 operand of the pwconv2 data gradient — and to HBM (dy of the pwconv2 weight gradient);
//     main loop : the intermediate dimension I is walked in chunks of 64 columns.  dG_j = dyg . W2[:, j]  (tcgen05, accumulator
//                 64 TMEM columns, double buffered) -> worker warps: dH_j = dG_j * gelu'(pre_j) -> fp16 into swizzled shared
//                 memory (A operand of the pwconv1 data gradient) and to HBM (dy of the pwconv1 weight gradient);
//                 dXhat += dH_j . W1f[j, :]  accumulates over all chunks in a second TMEM region (C columns).
//                 Both B operands are the FORWARD weight packs read MN-major (no transposed copies exist).
//     epilogue  : dXhat (fp32) -> HBM; with the intermediate dimension split over blockIdx.y (few row tiles: the encoder's 48
//                 tiles on 148 SMs) the partial sums meet in L2 (red.global.add.v4.f32 into a zeroed buffer).
//   The (positions x I) gradient of the GELU output never exists in HBM; dH is written once (fp16) for the weight gradient.
//
//   ln_dwconv_bwd_kernel: LayerNorm backward (no affine: it is folded into W1f) + depthwise-conv7 backward + residual path +
//   the depthwise parameter gradients, one pass: a warp walks a run of consecutive positions with a 7-row register window
//   of the LayerNorm input gradient, so that gradient never exists in HBM either.
//
//   resid_param_grad_kernel: dgamma / db2 from dout, the block's input and its output (gamma * z * rs = out - x on unmasked
//   rows, so the pwconv2 output z is not saved by the forward pass).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int BB_M = 128;      // rows per CTA
constexpr int BB_NC = 64;      // intermediate columns per chunk
constexpr int BB_WORKERS = 8;
constexpr int BB_CTRL = 3;      // control warps: TMA producer, GEMM 1 issuer, GEMM 2 issuer
constexpr int BB_THREADS = BB_CTRL * 32 + BB_WORKERS * 32;

struct BwdParams {
  const float* dout;        // (B, T, C) fp32 gradient of the block output (carries the static loss scale)
  const float* gamma;       // (C)
  const float* row_scale;   // (B) DropPath scale or null
  const uint8_t* pad_mask;  // (B*T) or null
  const __half* pre;        // (B, T, I) saved GELU argument
  __half* dyg_out;          // (B, T, C)
  __half* dh_out;           // (B, T, I)
  float* dxh_out;           // (nsplit, B, T, C) gradient wrt the normalised dwconv output: one partial sum per split
  int B, T, m_tiles, nsplit;
};

template <int C>
struct BwdCfg {
  static constexpr int KB = C / 64;
  static constexpr int A_BYTES = KB * BB_M * 128;          // dyg tile
  static constexpr int WA_BYTES = C * 128;                 // W2[:, chunk]: C contraction rows x 64 columns (MN-major B of GEMM 1)
  static constexpr int WB_BYTES = KB * BB_NC * 128;        // W1f[chunk, :]: 64 contraction rows x C columns (MN-major B of GEMM 2)
  static constexpr int H_BYTES = BB_M * 128;               // dH chunk
  static constexpr int WS = (A_BYTES + 2 * (WA_BYTES + WB_BYTES) + 2 * H_BYTES + 1280 <= 227 * 1024) ? 2 : 1;
  static constexpr int SMEM = A_BYTES + WS * (WA_BYTES + WB_BYTES) + 2 * H_BYTES + 1024 + 256;
  static constexpr int OUT_LD = C + 4;
  static_assert(BB_M * OUT_LD * 4 <= A_BYTES + WS * (WA_BYTES + WB_BYTES) + 2 * H_BYTES, "staged output tile must fit");
  static constexpr int N2 = C <= 256 ? C : C / 2;
  static constexpr int N2_PARTS = C / N2;
  static constexpr uint32_t ACC1_COL = C;
};

__device__ __forceinline__ uint32_t sw128_off(int r, int c16) { return static_cast<uint32_t>(r * 128 + ((c16 ^ (r & 7)) << 4)); }

template <int C, int I>
__global__ void __launch_bounds__(BB_THREADS, 1)
convnext_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW1, const BwdParams p) {
  using Cfg = BwdCfg<C>;
  constexpr int NCH = I / BB_NC;
  constexpr int VPL = C / 128;
  constexpr int WS = Cfg::WS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sWA = sA + Cfg::A_BYTES;                 // [WS]
  uint8_t* sWB = sWA + WS * Cfg::WA_BYTES;          // [WS]
  uint8_t* sH = sWB + WS * Cfg::WB_BYTES;           // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * Cfg::H_BYTES);
  uint64_t* wa_full = bars + 0;     // [2]
  uint64_t* wa_empty = bars + 2;    // [2]
  uint64_t* wb_full = bars + 4;     // [2]
  uint64_t* wb_empty = bars + 6;    // [2]
  uint64_t* acc1_full = bars + 8;   // [2]
  uint64_t* acc1_empty = bars + 10; // [2]
  uint64_t* h_full = bars + 12;     // [2]
  uint64_t* h_empty = bars + 14;    // [2]
  uint64_t* a_ready = bars + 16;
  uint64_t* acc2_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.m_tiles;
  const int t0 = (blockIdx.x % p.m_tiles) * BB_M;
  const int split = blockIdx.y;
  const int ch_begin = (NCH * split) / p.nsplit;
  const int n_ch = (NCH * (split + 1)) / p.nsplit - ch_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&wa_full[i], 1); mbar_init(&wa_empty[i], 1); mbar_init(&wb_full[i], 1); mbar_init(&wb_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], BB_WORKERS);
      mbar_init(&h_full[i], BB_WORKERS);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(a_ready, BB_WORKERS);
    mbar_init(acc2_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: weight chunks (forward packs, boxes of 64 columns x 64 rows) =====================
    if (lane == 0) {
      for (int j = 0; j < n_ch; ++j) {
        const int st = j % WS;
        const uint32_t ph = (j / WS) & 1;
        const int ja = ch_begin + j;
        mbar_wait(&wa_empty[st], ph ^ 1);
        mbar_expect_tx(&wa_full[st], Cfg::WA_BYTES);
        // W2 pack is (C rows, I columns): contraction index c = row, output index i = column (contiguous) -> MN-major B
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb)
          tma_load_3d(sWA + st * Cfg::WA_BYTES + kb * (64 * 128), &tmW2, &wa_full[st], ja * BB_NC, kb * 64, 0);
        mbar_wait(&wb_empty[st], ph ^ 1);
        mbar_expect_tx(&wb_full[st], Cfg::WB_BYTES);
        // W1f pack is (I rows, C columns): contraction index i = row, output index c = column (contiguous) -> MN-major B
#pragma unroll
        for (int cc = 0; cc < Cfg::KB; ++cc)
          tma_load_3d(sWB + st * Cfg::WB_BYTES + cc * (64 * 128), &tmW1, &wb_full[st], cc * 64, ja * BB_NC, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer 1: dG chunk = dyg (128 x C) . W2[:, chunk]  (K = C) -> acc1[j & 1] =====================
    // (two issuing warps, as in the forward kernel: GEMM 1 of chunk j+1 does not queue behind the wait for dH of chunk j-1)
    constexpr uint32_t idesc1 = make_instr_desc(OSB_F16, BB_M, BB_NC, 0, 1);
    const uint64_t dA0 = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    // MN-major B: 64-column chunks are 64 * 128 B apart (LBO), 8-row groups 1024 B apart (SBO)
    const uint64_t dWA0 = make_smem_desc_sw128(smem_u32(sWA), 64 * 128, 1024);
    mbar_wait(a_ready, 0);
    tc_fence_after_sync();
    for (int j = 0; j < n_ch; ++j) {
      const int buf = j & 1;
      const int st = j % WS;
      mbar_wait(&wa_full[st], (j / WS) & 1);
      mbar_wait(&acc1_empty[buf], ((j >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      const uint32_t d = tmem_base + Cfg::ACC1_COL + buf * BB_NC;
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = dA0 + static_cast<uint64_t>((kb * (BB_M * 128) + k * 32) >> 4);
            // 16 contraction rows per MMA = 2048 B inside the k-block's [64 rows x 128 B] box
            const uint64_t db = dWA0 + static_cast<uint64_t>((st * Cfg::WA_BYTES + kb * (64 * 128) + k * (16 * 128)) >> 4);
            umma_ss<false>(d, da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
          }
        umma_commit(&wa_empty[st]);
        umma_commit(&acc1_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== MMA issuer 2: acc2 += dH chunk (128 x 64) . W1f[chunk, :]  (K = 64) =====================
    constexpr uint32_t idesc2 = make_instr_desc(OSB_F16, BB_M, Cfg::N2, 0, 1);
    const uint64_t dH0 = make_smem_desc_sw128(smem_u32(sH), 16, 1024);
    const uint64_t dWB0 = make_smem_desc_sw128(smem_u32(sWB), 64 * 128, 1024);
    for (int jj = 0; jj < n_ch; ++jj) {
      const int buf = jj & 1;
      const int st = jj % WS;
      mbar_wait(&wb_full[st], (jj / WS) & 1);
      mbar_wait(&h_full[buf], (jj >> 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = dH0 + static_cast<uint64_t>((buf * Cfg::H_BYTES + k * 32) >> 4);
#pragma unroll
          for (int part = 0; part < Cfg::N2_PARTS; ++part) {
            const uint64_t db = dWB0 + static_cast<uint64_t>((st * Cfg::WB_BYTES + part * (Cfg::N2 / 64) * (64 * 128) + k * (16 * 128)) >> 4);
            umma_ss<false>(tmem_base + part * Cfg::N2, da, db, idesc2, (jj | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&wb_empty[st]);
        umma_commit(&h_empty[buf]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc2_full);
    __syncwarp();
  } else {
    // ===================== worker warps =====================
    const int ww = warp - BB_CTRL;    // 0..7
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (ww >> 2);       // which 32-column half of a chunk / which half of C in the last epilogue
    // ---- prologue: dyg = dout * gamma * keep * rs -> fp16, swizzled smem (rows ww*16 .. +16) and HBM ----
    {
      float4 g4[VPL];
#pragma unroll
      for (int v = 0; v < VPL; ++v) g4[v] = __ldg(reinterpret_cast<const float4*>(p.gamma + v * 128 + lane * 4));
      const float rs = p.row_scale != nullptr ? p.row_scale[b] : 1.f;
      const int r_begin = ww * 16;
#pragma unroll 4
      for (int r = r_begin; r < r_begin + 16; ++r) {
        const int t = t0 + r;
        const long long grow = static_cast<long long>(b) * p.T + t;
        float s = 0.f;
        if (t < p.T) s = (p.pad_mask != nullptr && p.pad_mask[grow]) ? 0.f : rs;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * 128 + lane * 4;
          float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t < p.T) d = *reinterpret_cast<const float4*>(p.dout + grow * C + c);
          __half2 h0 = __floats2half2_rn(d.x * s * g4[v].x, d.y * s * g4[v].y);
          __half2 h1 = __floats2half2_rn(d.z * s * g4[v].z, d.w * s * g4[v].w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          const int kb = c >> 6, cc = c & 63;
          *reinterpret_cast<uint2*>(sA + kb * (BB_M * 128) + sw128_off(r, cc >> 3) + (cc & 7) * 2) = u;
          if (split == 0 && t < p.T) *reinterpret_cast<uint2*>(p.dyg_out + grow * C + c) = u;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
    // ---- per chunk: acc1 * gelu'(pre) -> fp16 -> swizzled smem (A operand of GEMM 2) and HBM ----
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool row_ok = t < p.T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const __half* pre_row = p.pre + (static_cast<long long>(b) * p.T + (row_ok ? t : 0)) * I + half * 32;
    __half* dh_row = p.dh_out + (static_cast<long long>(b) * p.T + (row_ok ? t : 0)) * I + half * 32;
    uint4 pre_cur[4], pre_nxt[4];   // this thread's 32 GELU arguments of the chunk, fetched one chunk ahead
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
      pre_cur[c4] = row_ok ? *reinterpret_cast<const uint4*>(pre_row + ch_begin * BB_NC + c4 * 8) : make_uint4(0u, 0u, 0u, 0u);
    for (int j = 0; j < n_ch; ++j) {
      const int buf = j & 1;
      if (j + 1 < n_ch) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          pre_nxt[c4] = row_ok ? *reinterpret_cast<const uint4*>(pre_row + (ch_begin + j + 1) * BB_NC + c4 * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&acc1_full[buf], (j >> 1) & 1);
      tc_fence_after_sync();
      uint32_t rr[32];
      tmem_ld_32x32(lane_addr + Cfg::ACC1_COL + buf * BB_NC + half * 32, rr);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[buf]);
      float v[32];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const __half2* hp = reinterpret_cast<const __half2*>(&pre_cur[c4]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(hp[e]);
          v[c4 * 8 + 2 * e] = __uint_as_float(rr[c4 * 8 + 2 * e]) * gelu_erf_grad(f.x);
          v[c4 * 8 + 2 * e + 1] = __uint_as_float(rr[c4 * 8 + 2 * e + 1]) * gelu_erf_grad(f.y);
        }
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) pre_cur[c4] = pre_nxt[c4];
      mbar_wait(&h_empty[buf], ((j >> 1) & 1) ^ 1);
      uint8_t* hrow = sH + buf * Cfg::H_BYTES;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        __half2 h0 = __floats2half2_rn(v[c4 * 8 + 0], v[c4 * 8 + 1]);
        __half2 h1 = __floats2half2_rn(v[c4 * 8 + 2], v[c4 * 8 + 3]);
        __half2 h2 = __floats2half2_rn(v[c4 * 8 + 4], v[c4 * 8 + 5]);
        __half2 h3 = __floats2half2_rn(v[c4 * 8 + 6], v[c4 * 8 + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        if (!row_ok) u = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(hrow + sw128_off(row, half * 4 + c4)) = u;
        if (row_ok) *reinterpret_cast<uint4*>(dh_row + (ch_begin + j) * BB_NC + c4 * 8) = u;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[buf]);
    }
    // ---- final epilogue: dXhat tile staged in shared memory (all MMA operands are dead), then coalesced row writes ----
    mbar_wait(acc2_full, 0);
    tc_fence_after_sync();
    float* stile = reinterpret_cast<float*>(smem);
    constexpr int OLD = Cfg::OUT_LD;
    constexpr int CH = C / 2;
    for (int c0 = half * CH; c0 < (half + 1) * CH; c0 += 32) {
      uint32_t rr[32];
      tmem_ld_32x32(lane_addr + c0, rr);
      tmem_ld_wait();
      float* dst = stile + row * OLD + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1]), __uint_as_float(rr[i + 2]),
                                                          __uint_as_float(rr[i + 3]));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(BB_WORKERS * 32) : "memory");   // worker warps only
    for (int r = ww; r < BB_M; r += BB_WORKERS) {          // one warp per row
      const int tt = t0 + r;
      if (tt >= p.T) break;
      const long long grow = static_cast<long long>(b) * p.T + tt;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = v * 128 + lane * 4;
        const float4 d4 = *reinterpret_cast<const float4*>(stile + r * OLD + c);
        // each split writes its own partial sum (plain stores: no zero fill, no atomics); osb_ln_dwconv_bwd adds them up
        *reinterpret_cast<float4*>(p.dxh_out + (static_cast<long long>(split) * p.B * p.T + grow) * C + c) = d4;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int C, int I>
int launch_bwd(const void* w2_h16, const void* w1f_h16, const BwdParams& p, cudaStream_t stream) {
  using Cfg = BwdCfg<C>;
  CUtensorMap tmW2, tmW1;
  // W2 pack (C, I): contiguous index i; boxes of 64 columns x 64 rows.  W1f pack (I, C): contiguous index c.
  int rc = make_tmap_3d(&tmW2, w2_h16, TMA_F16, I, C, 1, I, static_cast<uint64_t>(C) * I, 64, 64);
  if (rc != OSB_OK) return rc;
  rc = make_tmap_3d(&tmW1, w1f_h16, TMA_F16, C, I, 1, C, static_cast<uint64_t>(I) * C, 64, 64);
  if (rc != OSB_OK) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(convnext_bwd_fused_kernel<C, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  convnext_bwd_fused_kernel<C, I><<<dim3(p.B * p.m_tiles, p.nsplit), BB_THREADS, Cfg::SMEM, stream>>>(tmW2, tmW1, p);
  count_launch();
  return launch_status();
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward (no affine) + depthwise conv7 backward + residual path + depthwise parameter gradients.
//   dd[t,c]  = (g[t,c] - mean_c(g[t,:]) - xhat[t,c] * mean_c(g[t,:] * xhat[t,:])) * rstd[t]          g = dxh
//   dx[t,c]  = dout[t,c] * keep[t] + sum_j w[c,j] * dd[t - j + 3, c]
//   ddw[c,j] += sum_u dd[u - j + 3, c] * x[u, c] ;  ddb[c] += sum_t dd[t, c]
// A warp owns a run of RUN consecutive positions of one sequence (lanes own 4 channels per 128-channel group) and slides a
// 7-row register window of dd over it; the 6 halo rows of dd are recomputed by the neighbouring runs.
// ------------------------------------------------------------------------------------------
constexpr int LDB_TT = 32;              // positions per tile
constexpr int LDB_HR = LDB_TT + 6;      // staged rows of dd (3 halo rows each side)
constexpr int LDB_WARPS = 8;

// A block owns tiles of LDB_TT consecutive positions of one sequence (grid-stride over tiles).  Phase 1: the warps compute
// the LayerNorm backward of the tile's LDB_HR rows (independent rows: all loads of a warp's rows are in flight together)
// into a shared-memory tile; phase 2: every output row reads its 7 neighbours from that tile.  The 6 halo rows per tile
// are recomputed by the neighbouring tiles (19 % extra LayerNorm work, no dd round trip through HBM).  Parameter gradients
// are accumulated in registers over all tiles of the block and flushed once: (8, C) fp32, rows 0-6 = taps, row 7 = bias.
template <int VPL>
__global__ void __launch_bounds__(LDB_WARPS * 32)
ln_dwconv_bwd_kernel(const float* __restrict__ dxh, const __half* __restrict__ xhat, const float* __restrict__ rstd,
                     const float* __restrict__ dout, const float* __restrict__ x, const float* __restrict__ w /*(C,7)*/,
                     const uint8_t* __restrict__ pad_mask, float* __restrict__ dx, float* __restrict__ dparam /*(8,C)*/, int B, int T,
                     int tiles_per_seq, int ntiles, int nparts) {
  constexpr int C = 128 * VPL;
  extern __shared__ float ldb_smem[];
  float* sdd = ldb_smem;                 // [LDB_HR][C]
  float* stw = sdd + LDB_HR * C;         // [7][C] taps, tap-major
  const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 7 * C; i += LDB_WARPS * 32) stw[(i % 7) * C + i / 7] = w[i];
  float4 aw[7][VPL], ab[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 7; ++j) aw[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  constexpr int RPW = (LDB_HR + LDB_WARPS - 1) / LDB_WARPS;   // dd rows per warp (5)
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_seq;
    const int tbeg = (tile % tiles_per_seq) * LDB_TT;
    const long long base = static_cast<long long>(b) * T;
    // ---- phase 1: dd rows tbeg-3 .. tbeg+LDB_TT+2 -> shared memory ----
    {
      float4 g[RPW][VPL];
      uint2 xh[RPW][VPL];
      float rs[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int rr = wip + LDB_WARPS * i;
        const int t = tbeg + rr - 3;
        const bool ok = rr < LDB_HR && t >= 0 && t < T;
        rs[i] = ok ? rstd[base + t] : 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * 128 + lane * 4;
          g[i][v] = ok ? *reinterpret_cast<const float4*>(dxh + (base + t) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int part = 1; part < nparts && ok; ++part) {   // partial sums of the fused backward kernel's splits
            const float4 e4 = *reinterpret_cast<const float4*>(dxh + (static_cast<long long>(part) * B * T + base + t) * C + c);
            g[i][v].x += e4.x; g[i][v].y += e4.y; g[i][v].z += e4.z; g[i][v].w += e4.w;
          }
          xh[i][v] = ok ? *reinterpret_cast<const uint2*>(xhat + (base + t) * C + c) : make_uint2(0u, 0u);
        }
      }
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int rr = wip + LDB_WARPS * i;
        if (rr >= LDB_HR) continue;
        float4 xf[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&xh[i][v].x));
          const float2 a1 = __half22float2(*reinterpret_cast<const __half2*>(&xh[i][v].y));
          xf[v] = make_float4(a0.x, a0.y, a1.x, a1.y);
          s1 += (g[i][v].x + g[i][v].y) + (g[i][v].z + g[i][v].w);
          s2 += (g[i][v].x * xf[v].x + g[i][v].y * xf[v].y) + (g[i][v].z * xf[v].z + g[i][v].w * xf[v].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {     // the two row sums travel together
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 * (1.f / C), m2 = s2 * (1.f / C);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          float4 d;
          d.x = (g[i][v].x - m1 - xf[v].x * m2) * rs[i];
          d.y = (g[i][v].y - m1 - xf[v].y * m2) * rs[i];
          d.z = (g[i][v].z - m1 - xf[v].z * m2) * rs[i];
          d.w = (g[i][v].w - m1 - xf[v].w * m2) * rs[i];
          *reinterpret_cast<float4*>(sdd + rr * C + v * 128 + lane * 4) = d;
        }
      }
    }
    __syncthreads();
    // ---- phase 2: dx and the parameter-gradient partial sums; rows wip, wip + 8, ... of the tile ----
    {
      constexpr int OPW = LDB_TT / LDB_WARPS;   // output rows per warp (4)
      float4 xo[OPW][VPL], go[OPW][VPL];
      float keep[OPW];
#pragma unroll
      for (int i = 0; i < OPW; ++i) {
        const int t = tbeg + wip + LDB_WARPS * i;
        const bool ok = t < T;
        keep[i] = (ok && !(pad_mask != nullptr && pad_mask[base + t])) ? 1.f : 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * 128 + lane * 4;
          xo[i][v] = ok ? *reinterpret_cast<const float4*>(x + (base + t) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          go[i][v] = ok ? *reinterpret_cast<const float4*>(dout + (base + t) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < OPW; ++i) {
        const int r = wip + LDB_WARPS * i;
        const int t = tbeg + r;
        if (t >= T) continue;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * 128 + lane * 4;
          float4 acc = make_float4(go[i][v].x * keep[i], go[i][v].y * keep[i], go[i][v].z * keep[i], go[i][v].w * keep[i]);
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const float4 wj = *reinterpret_cast<const float4*>(stw + j * C + c);
            const float4 d = *reinterpret_cast<const float4*>(sdd + (r + 6 - j) * C + c);   // dd[t - j + 3]
            acc.x = fmaf(wj.x, d.x, acc.x); acc.y = fmaf(wj.y, d.y, acc.y);
            acc.z = fmaf(wj.z, d.z, acc.z); acc.w = fmaf(wj.w, d.w, acc.w);
            aw[j][v].x = fmaf(d.x, xo[i][v].x, aw[j][v].x); aw[j][v].y = fmaf(d.y, xo[i][v].y, aw[j][v].y);   // ddw[c,j] += dd[u-j+3] x[u]
            aw[j][v].z = fmaf(d.z, xo[i][v].z, aw[j][v].z); aw[j][v].w = fmaf(d.w, xo[i][v].w, aw[j][v].w);
          }
          *reinterpret_cast<float4*>(dx + (base + t) * C + c) = acc;
          const float4 dc = *reinterpret_cast<const float4*>(sdd + (r + 3) * C + c);
          ab[v].x += dc.x; ab[v].y += dc.y; ab[v].z += dc.z; ab[v].w += dc.w;
        }
      }
    }
    __syncthreads();   // the next tile overwrites sdd
  }
  // ---- flush the parameter gradients: the warps add into a shared accumulator in turn, then one vector reduction per 4 channels ----
  float* sacc = sdd;   // [8][C]
  for (int i = threadIdx.x; i < 8 * C; i += LDB_WARPS * 32) sacc[i] = 0.f;
  __syncthreads();
  for (int q = 0; q < LDB_WARPS; ++q) {
    if (wip == q) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = v * 128 + lane * 4;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          float4 t4 = *reinterpret_cast<float4*>(sacc + j * C + c);
          t4.x += aw[j][v].x; t4.y += aw[j][v].y; t4.z += aw[j][v].z; t4.w += aw[j][v].w;
          *reinterpret_cast<float4*>(sacc + j * C + c) = t4;
        }
        float4 t4 = *reinterpret_cast<float4*>(sacc + 7 * C + c);
        t4.x += ab[v].x; t4.y += ab[v].y; t4.z += ab[v].z; t4.w += ab[v].w;
        *reinterpret_cast<float4*>(sacc + 7 * C + c) = t4;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * C; i += LDB_WARPS * 32) {
    const float4 t4 = *reinterpret_cast<const float4*>(sacc + 4 * i);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dparam + 4 * i), "f"(t4.x), "f"(t4.y), "f"(t4.z), "f"(t4.w) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// dgamma[c] += (1 / gamma[c]) * sum_rows dout * keep * (out - x)      (gamma * z * rs = out - x where keep = 1)
// db2[c]    += sum_rows dout * keep * rs * gamma
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
resid_param_grad_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ x,
                        const float* __restrict__ gamma, const uint8_t* __restrict__ pad_mask, const float* __restrict__ row_scale,
                        float* __restrict__ dgamma, float* __restrict__ db2, long long rows, int T, int rows_per_warp) {
  constexpr int C = 128 * VPL;
  const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
  const long long w = static_cast<long long>(blockIdx.x) * 8 + wip;
  const long long r0 = w * rows_per_warp;
  const long long r1 = r0 + rows_per_warp < rows ? r0 + rows_per_warp : rows;
  float4 ag[VPL], ab[VPL], g[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    g[v] = *reinterpret_cast<const float4*>(gamma + v * 128 + lane * 4);
    ag[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = r0; r < r1; ++r) {
    if (pad_mask != nullptr && pad_mask[r]) continue;
    const float rs = row_scale != nullptr ? row_scale[r / T] : 1.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = v * 128 + lane * 4;
      const float4 d = *reinterpret_cast<const float4*>(dout + r * C + c);
      const float4 o = *reinterpret_cast<const float4*>(out + r * C + c);
      const float4 xi = *reinterpret_cast<const float4*>(x + r * C + c);
      ag[v].x = fmaf(d.x, o.x - xi.x, ag[v].x); ag[v].y = fmaf(d.y, o.y - xi.y, ag[v].y);
      ag[v].z = fmaf(d.z, o.z - xi.z, ag[v].z); ag[v].w = fmaf(d.w, o.w - xi.w, ag[v].w);
      ab[v].x = fmaf(d.x, rs, ab[v].x); ab[v].y = fmaf(d.y, rs, ab[v].y);
      ab[v].z = fmaf(d.z, rs, ab[v].z); ab[v].w = fmaf(d.w, rs, ab[v].w);
    }
  }
  __shared__ float4 s_red[8][2 * VPL * 32];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    s_red[wip][v * 32 + lane] = ag[v];
    s_red[wip][(VPL + v) * 32 + lane] = ab[v];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * VPL * 32; i += 256) {
    float4 t = s_red[0][i];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float4 u = s_red[q][i];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const bool is_b = i >= VPL * 32;
    const int ii = is_b ? i - VPL * 32 : i;
    const int c = (ii >> 5) * 128 + (ii & 31) * 4;
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    if (is_b) {
      t.x *= gm.x; t.y *= gm.y; t.z *= gm.z; t.w *= gm.w;
    } else {   // a layer scale of exactly 0 carries no information about z in (out - x): its gradient is reported as 0
      t.x = gm.x != 0.f ? t.x / gm.x : 0.f; t.y = gm.y != 0.f ? t.y / gm.y : 0.f;
      t.z = gm.z != 0.f ? t.z / gm.z : 0.f; t.w = gm.w != 0.f ? t.w / gm.w : 0.f;
    }
    float* dst = (is_b ? db2 : dgamma) + c;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
}

}  // namespace
}  // namespace osb

using namespace osb;

static int g_bwd_nsplit = 0;
/* developer hook (not in the public header): force the intermediate-dimension split of the fused backward (0 = automatic) */
extern "C" void osb_debug_set_bwd_nsplit(int n) { g_bwd_nsplit = n; }

/* number of partial buffers osb_convnext_block_bwd writes for this shape (the caller sizes dxhat with it) */
extern "C" int osb_convnext_block_bwd_parts(int32_t B, int32_t T, int32_t I) {
  const int tiles = B * ((T + BB_M - 1) / BB_M);
  const int nch = I / BB_NC;
  int ns = g_bwd_nsplit > 0 ? g_bwd_nsplit : (tiles * 2 > 148 ? 1 : (148 / tiles > 4 ? 4 : 148 / tiles));
  if (ns > nch / 2) ns = nch / 2;
  return ns < 1 ? 1 : ns;
}

extern "C" int osb_convnext_block_bwd(const float* dout, const float* gamma, const float* row_scale, const uint8_t* pad_mask,
                                      const void* pre_h16, const void* w2_h16, const void* w1f_h16, void* dyg_h16, void* dh_h16,
                                      float* dxhat, int32_t B, int32_t T, int32_t C, int32_t I, void* stream) {
  OSB_REQUIRE(dout && gamma && pre_h16 && w2_h16 && w1f_h16 && dyg_h16 && dh_h16 && dxhat, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0, OSB_ERR_SHAPE);
  for (const void* q : {static_cast<const void*>(dout), static_cast<const void*>(gamma), pre_h16, static_cast<const void*>(dyg_h16),
                        static_cast<const void*>(dh_h16), static_cast<const void*>(dxhat)})
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return OSB_ERR_ALIGN;
  BwdParams p;
  p.dout = dout; p.gamma = gamma; p.row_scale = row_scale; p.pad_mask = pad_mask;
  p.pre = static_cast<const __half*>(pre_h16);
  p.dyg_out = static_cast<__half*>(dyg_h16); p.dh_out = static_cast<__half*>(dh_h16); p.dxh_out = dxhat;
  p.B = B; p.T = T; p.m_tiles = (T + BB_M - 1) / BB_M;
  p.nsplit = osb_convnext_block_bwd_parts(B, T, I);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 256 && I == 1024) return launch_bwd<256, 1024>(w2_h16, w1f_h16, p, s);
  if (C == 384 && I == 1152) return launch_bwd<384, 1152>(w2_h16, w1f_h16, p, s);
  return OSB_ERR_SHAPE;
}

extern "C" int osb_ln_dwconv_bwd(const float* dxhat, int32_t nparts, const void* xhat_h16, const float* rstd, const float* dout,
                                 const float* x, const float* dw_w, const uint8_t* pad_mask, float* dx, float* dparam, int32_t B, int32_t T,
                                 int32_t C, void* stream) {
  OSB_REQUIRE(dxhat && xhat_h16 && rstd && dout && x && dw_w && dx && dparam, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && (C == 256 || C == 384) && nparts >= 1 && nparts <= 16, OSB_ERR_SHAPE);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(dparam) & 15) == 0, OSB_ERR_ALIGN);
  const int tiles_per_seq = (T + LDB_TT - 1) / LDB_TT;
  const long long ntiles = static_cast<long long>(B) * tiles_per_seq;
  const unsigned blocks = static_cast<unsigned>(ntiles < 148 * 2 ? ntiles : 148 * 2);
  const size_t smem = static_cast<size_t>(LDB_HR + 7) * C * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* xh = static_cast<const __half*>(xhat_h16);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(ln_dwconv_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LDB_HR + 7) * 256 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ln_dwconv_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (LDB_HR + 7) * 384 * 4);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  if (C == 256) ln_dwconv_bwd_kernel<2><<<blocks, LDB_WARPS * 32, smem, s>>>(dxhat, xh, rstd, dout, x, dw_w, pad_mask, dx, dparam, B, T, tiles_per_seq, static_cast<int>(ntiles), nparts);
  else ln_dwconv_bwd_kernel<3><<<blocks, LDB_WARPS * 32, smem, s>>>(dxhat, xh, rstd, dout, x, dw_w, pad_mask, dx, dparam, B, T, tiles_per_seq, static_cast<int>(ntiles), nparts);
  count_launch();
  return launch_status();
}

extern "C" int osb_resid_param_grad(const float* dout, const float* out, const float* x, const float* gamma, const uint8_t* pad_mask,
                                    const float* row_scale, float* dgamma, float* db2, int64_t rows, int32_t T, int32_t C, void* stream) {
  OSB_REQUIRE(dout && out && x && gamma && dgamma && db2, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && T > 0 && (C == 256 || C == 384), OSB_ERR_SHAPE);
  const int rows_per_warp = 4;
  const long long warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const unsigned blocks = static_cast<unsigned>((warps + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 256) resid_param_grad_kernel<2><<<blocks, 256, 0, s>>>(dout, out, x, gamma, pad_mask, row_scale, dgamma, db2, rows, T, rows_per_warp);
  else resid_param_grad_kernel<3><<<blocks, 256, 0, s>>>(dout, out, x, gamma, pad_mask, row_scale, dgamma, db2, rows, T, rows_per_warp);
  count_launch();
  return launch_status();
}
