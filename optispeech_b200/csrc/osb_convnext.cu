// osb_convnext.cu — ONE kernel per ConvNeXt block (forward):
//
//   out = (x + gamma * rs[b] * (W2 . gelu(W1f . LNhat(dwconv7(x)) + b1f) + b2)) * keep
//
// A CTA owns 128 consecutive positions of one sequence.
//   prologue   : 8 worker warps compute depthwise-conv7 + LayerNorm statistics on CUDA cores and write the fp16
//                normalised tile straight into 128B-swizzled shared memory — it IS the A operand of pwconv1;
//   main loop  : the intermediate dimension I is walked in chunks of 64 columns.  One thread issues
//                tcgen05.mma for pwconv1 of chunk j (accumulator: 64 TMEM columns, double buffered); the worker
//                warps read it back (tcgen05.ld), add the bias, apply exact-erf GELU and write the fp16 chunk into
//                swizzled shared memory, where it is the A operand of pwconv2 for that chunk; pwconv2 accumulates
//                over all chunks into a second TMEM region (C columns).  The (positions x I) intermediate never
//                leaves the SM.  Weight chunks are streamed by TMA (one producer thread) through full/empty
//                mbarriers; pwconv1 of chunk j+1 overlaps the GELU of chunk j and pwconv2 of chunk j-1.
//   epilogue   : bias, layer scale, DropPath scale, residual, pad mask -> fp32 HBM.
// HBM traffic per position: 4C bytes read (+ halo) and 4C written, i.e. the algorithmic minimum; the three-kernel
// path moves an extra 2C + 2*2I + 4C bytes per position.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int FB_M = 128;      // rows per CTA
constexpr int FB_NC = 64;      // intermediate columns per chunk (= one 128-byte swizzle row of fp16)
constexpr int FB_WORKERS = 8;  // worker warps (prologue + epilogues)
constexpr int FB_CTRL = 3;      // control warps: TMA producer, pwconv1 MMA issuer, pwconv2 MMA issuer
constexpr int FB_THREADS = FB_CTRL * 32 + FB_WORKERS * 32;

struct FusedParams {
  const float* x;        // (B, T, C) fp32 residual stream
  const float* dw_w;     // (C, 7)
  const float* dw_b;     // (C)
  const float* b1;       // (I) folded bias
  const float* b2;       // (C)
  const float* gamma;    // (C)
  const float* row_scale;   // (B) or null
  const uint8_t* pad_mask;  // (B*T) or null
  float* out;            // (B, T, C)
  // training mode (kTrain): activations the hand-written backward needs, written while they are on chip
  __half* xhat_out;      // (B, T, C) fp16 normalised dwconv output (A operand of pwconv1, LN affine folded into W1f)
  float* rstd_out;       // (B, T)
  __half* pre_out;       // (B, T, I) fp16 pwconv1 output + bias (GELU argument)
  __half* h_out;         // (B, T, I) fp16 GELU output (A operand of pwconv2)
  int I;
  int B, T, m_tiles;
  int pair;              // 1: (C = 384, T = 64) a 128-row tile holds TWO consecutive samples; B, T above are then B/2 and 128 (the
                         //    rows of two 64-frame samples are contiguous), only the input staging and row_scale see the real samples
  int nsplit;            // > 1: blockIdx.y owns a slice of the intermediate chunks and adds its partial result into `out` (pre-zeroed)
  float eps;
  long long* trace;      // optional (developer): clock64 timeline of CTA 0, [role][event] (see tools/probe_fused.py)
};

#define FB_TRACE(role, idx) \
  do { if (p.trace != nullptr && blockIdx.x == 0) p.trace[(role) * 256 + (idx)] = clock64(); } while (0)

template <int C>
struct FusedCfg {
  static constexpr int KB = C / 64;                       // 64-element k-blocks of the A operand
  static constexpr int A_BYTES = KB * FB_M * 128;          // xhat tile
  static constexpr int W1_BYTES = KB * FB_NC * 128;        // W1 chunk: 64 rows x C
  static constexpr int W2_BYTES = C * 128;                 // W2 chunk: C rows x 64
  static constexpr int H_BYTES = FB_M * 128;               // GELU chunk: 128 rows x 64
  // weight-chunk ring depth: two stages hide the TMA latency; C = 384 only has room for one (A alone is 96 KB)
  static constexpr int WS = (A_BYTES + 2 * (W1_BYTES + W2_BYTES) + 2 * H_BYTES + 1280 <= 227 * 1024) ? 2 : 1;
  static constexpr int SMEM = A_BYTES + WS * (W1_BYTES + W2_BYTES) + 2 * H_BYTES + 1024 + 256;
  static constexpr int OUT_LD = C + 4;                     // padded row stride (floats) of the staged output tile
  static_assert(FB_M * OUT_LD * 4 <= A_BYTES + WS * (W1_BYTES + W2_BYTES) + 2 * H_BYTES, "staged output tile must fit");
  static constexpr int N2 = C <= 256 ? C : C / 2;          // N of one pwconv2 MMA
  static constexpr int N2_PARTS = C / N2;
  static constexpr uint32_t ACC1_COL = C;                  // TMEM: [0, C) pwconv2 accumulator, then 2 x 64 for pwconv1
  // C = 256 leaves 128 TMEM columns: exactly the fp16 xhat tile (128 rows x 256 channels, two per column).  pwconv1 then takes
  // its A operand from tensor memory: the tile is read by all I/64 chunks, and as a shared-memory operand those re-reads
  // (64 KB per chunk) were the largest share of the shared-memory traffic that bounded the chunk loop (clock64 timeline, r2).
  static constexpr bool A_TMEM = (C == 256);
  static constexpr uint32_t A_COL = C + 2 * FB_NC;
  // prologue input staging: the (rows + 6 halo) x C fp32 input tile is brought in by TMA as C/32 boxes of [rows x 128 B]
  // (128B swizzle, zero fill outside [0, T) = the convolution's zero padding) into the weight / GELU buffers, which are idle
  // until the prologue is done; C = 384 takes two passes of 64 rows
  static constexpr int NXB = C / 32;
  static constexpr int NPASS = C <= 256 ? 1 : 2;
  static constexpr int RP = FB_M / NPASS;                  // output rows per pass
  static constexpr int XR = RP + 6;                        // staged input rows per pass
  static constexpr int XB_STRIDE = (XR * 128 + 1023) / 1024 * 1024;
  static constexpr int X_BYTES = NXB * XB_STRIDE;
  static constexpr int TAP_BYTES = 8 * C * 4;              // 7 taps + bias, tap-major
  static_assert(X_BYTES + TAP_BYTES <= WS * (W1_BYTES + W2_BYTES) + 2 * H_BYTES, "input tile must fit the idle buffers");
};

// 16-byte chunk `c16` (0..7) of row `r` inside a [rows x 128 B] tile with the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128_offset(int r, int c16) { return static_cast<uint32_t>(r * 128 + ((c16 ^ (r & 7)) << 4)); }

template <int C, int I, bool kTrain>
__global__ void __launch_bounds__(FB_THREADS, 1)
convnext_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                      const FusedParams p) {
  using Cfg = FusedCfg<C>;
  constexpr int NCH = I / FB_NC;
  constexpr int VPL = C / 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int WS = Cfg::WS;
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + Cfg::A_BYTES;                 // [WS]
  uint8_t* sW2 = sW1 + WS * Cfg::W1_BYTES;          // [WS]
  uint8_t* sH = sW2 + WS * Cfg::W2_BYTES;           // [2]
  uint8_t* sX = sW1;                                // prologue only: staged input tile, then the taps
  uint64_t* bars = reinterpret_cast<uint64_t*>(sH + 2 * Cfg::H_BYTES);
  uint64_t* w1_full = bars + 0;     // [2]
  uint64_t* w1_empty = bars + 2;    // [2]
  uint64_t* w2_full = bars + 4;     // [2]
  uint64_t* w2_empty = bars + 6;    // [2]
  uint64_t* acc1_full = bars + 8;   // [2]
  uint64_t* acc1_empty = bars + 10; // [2]
  uint64_t* h_full = bars + 12;     // [2]
  uint64_t* h_empty = bars + 14;    // [2]
  uint64_t* a_ready = bars + 16;
  uint64_t* acc2_full = bars + 17;
  uint64_t* x_full = bars + 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.m_tiles;
  const int t0 = (blockIdx.x % p.m_tiles) * FB_M;
  // Split of the intermediate dimension over blockIdx.y (small problems: fewer row tiles than SMs).  All pipeline indices
  // below are relative to this CTA's first chunk; only weight / bias coordinates use the absolute chunk index.
  const int split = blockIdx.y;
  const int ch_begin = (NCH * split) / p.nsplit;
  const int n_ch = (NCH * (split + 1)) / p.nsplit - ch_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], FB_WORKERS);
      mbar_init(&h_full[i], FB_WORKERS);
      mbar_init(&h_empty[i], 1);
    }
    mbar_init(a_ready, FB_WORKERS);
    mbar_init(acc2_full, 1);
    mbar_init(x_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: input tile (pass 0), then the weight chunks =====================
    if (lane == 0) {
      mbar_expect_tx(x_full, Cfg::NXB * Cfg::XR * 128);
#pragma unroll
      for (int i = 0; i < Cfg::NXB; ++i) tma_load_3d(sX + i * Cfg::XB_STRIDE, &tmX, x_full, i * 32, t0 - 3, p.pair ? 2 * b : b);
      mbar_wait(a_ready, 0);   // the weight ring overlays the staged input tile: the prologue has to be done with it
      for (int j = 0; j < n_ch; ++j) {
        const int st = j % WS;
        const uint32_t ph = (j / WS) & 1;
        const int ja = ch_begin + j;
        mbar_wait(&w1_empty[st], ph ^ 1);
        FB_TRACE(0, 2 * j);
        mbar_expect_tx(&w1_full[st], Cfg::W1_BYTES);
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb)
          tma_load_3d(sW1 + st * Cfg::W1_BYTES + kb * (FB_NC * 128), &tmW1, &w1_full[st], kb * 64, ja * FB_NC, 0);
        mbar_wait(&w2_empty[st], ph ^ 1);
        FB_TRACE(0, 2 * j + 1);
        mbar_expect_tx(&w2_full[st], Cfg::W2_BYTES);
#pragma unroll
        for (int part = 0; part < Cfg::N2_PARTS; ++part)
          tma_load_3d(sW2 + st * Cfg::W2_BYTES + part * (Cfg::N2 * 128), &tmW2, &w2_full[st], ja * FB_NC, part * Cfg::N2, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer 1: pwconv1 of every chunk -> acc1[j & 1] =====================
    // Two issuing warps: with a single one, pwconv1 of chunk j+1 queued behind the wait for the GELU of chunk j-1 (the
    // operand of pwconv2), and the ~60 cycles of scalar work per tcgen05.mma made the issuing thread itself the bound
    // (clock64 timeline, round 2).  The whole warp walks the loop with uniform control flow; one elected lane issues.
    constexpr uint32_t idesc1 = make_instr_desc(OSB_F16, FB_M, FB_NC, 0, 0);
    const uint64_t dA0 = make_smem_desc_sw128(smem_u32(sA), 16, 1024);
    const uint64_t dW10 = make_smem_desc_sw128(smem_u32(sW1), 16, 1024);
    if (lane == 0) FB_TRACE(1, 0);
    mbar_wait(a_ready, 0);
    if (lane == 0) FB_TRACE(1, 1);
    tc_fence_after_sync();
    for (int j = 0; j < n_ch; ++j) {
      const int buf = j & 1;
      const int st = j % WS;
      mbar_wait(&w1_full[st], (j / WS) & 1);
      if (lane == 0) FB_TRACE(1, 2 + 6 * j);
      mbar_wait(&acc1_empty[buf], ((j >> 1) & 1) ^ 1);
      if (lane == 0) FB_TRACE(1, 3 + 6 * j);
      tc_fence_after_sync();
      const uint32_t d = tmem_base + Cfg::ACC1_COL + buf * FB_NC;
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < Cfg::KB; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = dA0 + static_cast<uint64_t>((kb * (FB_M * 128) + k * 32) >> 4);
            const uint64_t db = dW10 + static_cast<uint64_t>((st * Cfg::W1_BYTES + kb * (FB_NC * 128) + k * 32) >> 4);
            if constexpr (Cfg::A_TMEM) umma_ts_f16(d, tmem_base + Cfg::A_COL + kb * 32 + k * 8, db, idesc1, (kb | k) != 0 ? 1u : 0u);
            else umma_ss<false>(d, da, db, idesc1, (kb | k) != 0 ? 1u : 0u);
          }
        umma_commit(&w1_empty[st]);
        umma_commit(&acc1_full[buf]);
      }
      __syncwarp();
      if (lane == 0) FB_TRACE(1, 4 + 6 * j);
    }
  } else if (warp == 2) {
    // ===================== MMA issuer 2: pwconv2, acc2 += gelu_chunk . W2[:, chunk]^T =====================
    constexpr uint32_t idesc2 = make_instr_desc(OSB_F16, FB_M, Cfg::N2, 0, 0);
    const uint64_t dW20 = make_smem_desc_sw128(smem_u32(sW2), 16, 1024);
    const uint64_t dH0 = make_smem_desc_sw128(smem_u32(sH), 16, 1024);
    for (int jj = 0; jj < n_ch; ++jj) {
      const int buf = jj & 1;
      const int st = jj % WS;
      mbar_wait(&w2_full[st], (jj / WS) & 1);
      if (lane == 0) FB_TRACE(1, 5 + 6 * jj);
      mbar_wait(&h_full[buf], (jj >> 1) & 1);
      if (lane == 0) FB_TRACE(1, 6 + 6 * jj);
      tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = dH0 + static_cast<uint64_t>((buf * Cfg::H_BYTES + k * 32) >> 4);
#pragma unroll
          for (int part = 0; part < Cfg::N2_PARTS; ++part) {
            const uint64_t db = dW20 + static_cast<uint64_t>((st * Cfg::W2_BYTES + part * (Cfg::N2 * 128) + k * 32) >> 4);
            umma_ss<false>(tmem_base + part * Cfg::N2, da, db, idesc2, (jj | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&w2_empty[st]);
        umma_commit(&h_empty[buf]);
      }
      __syncwarp();
      if (lane == 0) FB_TRACE(1, 7 + 6 * jj);
    }
    if (elect_one()) umma_commit(acc2_full);
    __syncwarp();
  } else {
    // ===================== worker warps =====================
    const int ww = warp - FB_CTRL;    // 0..7
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (ww >> 2);       // which 32-column half of a 64-column chunk / which half of C in the last epilogue
    // ---- prologue: dwconv7 + LayerNorm statistics -> fp16 xhat in swizzled smem ----
    {
      float* s_dw = reinterpret_cast<float*>(sX + Cfg::X_BYTES);   // [7][C] taps (tap-major), then [C] bias
      float* s_db = s_dw + 7 * C;
      const int wt = threadIdx.x - FB_CTRL * 32;                    // 0 .. 255
      for (int i = wt; i < 7 * C; i += FB_WORKERS * 32) s_dw[(i % 7) * C + i / 7] = p.dw_w[i];
      for (int i = wt; i < C; i += FB_WORKERS * 32) s_db[i] = p.dw_b[i];
      asm volatile("bar.sync 1, %0;" ::"n"(FB_WORKERS * 32) : "memory");   // worker warps only
      // this lane's 4 channels of every 128-channel group v live in box v*4 + lane/8, 16-byte chunk lane%8 of the box row
      const int xbox = lane >> 3, xchunk = lane & 7;
      constexpr int ROWS_PER_WARP = Cfg::RP / FB_WORKERS;
#pragma unroll 1
      for (int pass = 0; pass < Cfg::NPASS; ++pass) {
        if (pass > 0) {   // the next 64 rows: every worker is done with the previous pass's tile
          asm volatile("bar.sync 1, %0;" ::"n"(FB_WORKERS * 32) : "memory");
          if (ww == 0 && lane == 0) {
            mbar_expect_tx(x_full, Cfg::NXB * Cfg::XR * 128);
#pragma unroll
            for (int i = 0; i < Cfg::NXB; ++i)   // pair mode: the second half of the tile is the NEXT sample's rows 0..63 (+ its own zero halo)
              tma_load_3d(sX + i * Cfg::XB_STRIDE, &tmX, x_full, i * 32, p.pair ? -3 : t0 + pass * Cfg::RP - 3, p.pair ? 2 * b + pass : b);
          }
        }
        mbar_wait(x_full, pass & 1);
        // The staged rows are read ONCE each into a 7-row register window (fully unrolled: the window rotates at compile time);
        // for C = 256 the 7 x C taps of this lane's channels also live in registers — re-reading taps and rows from shared
        // memory for every output row made the prologue shared-memory-bandwidth bound (28 LDS.128 per row and warp, 8 warps:
        // ~14k cycles per tile, clock64 timeline round 2).
        constexpr bool kTapsInRegs = VPL <= 2;
        float4 tw[kTapsInRegs ? 7 : 1][VPL], tb[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          tb[v] = *reinterpret_cast<const float4*>(s_db + v * 128 + lane * 4);
          if constexpr (kTapsInRegs) {
#pragma unroll
            for (int jx = 0; jx < 7; ++jx) tw[jx][v] = *reinterpret_cast<const float4*>(s_dw + jx * C + v * 128 + lane * 4);
          }
        }
        auto ld_row = [&](int rr, float4 (&dst)[VPL]) {
#pragma unroll
          for (int v = 0; v < VPL; ++v)
            dst[v] = *reinterpret_cast<const float4*>(sX + (v * 4 + xbox) * Cfg::XB_STRIDE + rr * 128 + ((xchunk ^ (rr & 7)) << 4));
        };
        const int rl0 = ww * ROWS_PER_WARP;
        float4 win[7][VPL];           // slot (rl + jx) % 7 holds staged row rl + jx
#pragma unroll
        for (int jx = 0; jx < 6; ++jx) ld_row(rl0 + jx, win[jx]);
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; ++i) {
          const int rl = rl0 + i;      // row inside the pass; staged rows rl .. rl+6
          ld_row(rl + 6, win[(i + 6) % 7]);
          const int r = pass * Cfg::RP + rl;
          const int t = t0 + r;
          float4 d[VPL];
          float s = 0.f;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c = v * 128 + lane * 4;
            float4 acc = tb[v];
#pragma unroll
            for (int jx = 0; jx < 7; ++jx) {
              const float4 xv = win[(i + jx) % 7][v];
              float4 wj;
              if constexpr (kTapsInRegs) wj = tw[jx][v];
              else wj = *reinterpret_cast<const float4*>(s_dw + jx * C + c);
              acc.x = fmaf(wj.x, xv.x, acc.x);
              acc.y = fmaf(wj.y, xv.y, acc.y);
              acc.z = fmaf(wj.z, xv.z, acc.z);
              acc.w = fmaf(wj.w, xv.w, acc.w);
            }
            d[v] = acc;
            s += (acc.x + acc.y) + (acc.z + acc.w);
          }
          const float mean = warp_sum(s) * (1.f / C);
          float qv = 0.f;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            d[v].x -= mean; d[v].y -= mean; d[v].z -= mean; d[v].w -= mean;
            qv += (d[v].x * d[v].x + d[v].y * d[v].y) + (d[v].z * d[v].z + d[v].w * d[v].w);
          }
          const float rstd = (t < p.T) ? rsqrtf(warp_sum(qv) * (1.f / C) + p.eps) : 0.f;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const int c = v * 128 + lane * 4;
            const int kb = c >> 6, cc = c & 63;
            __half2 h0 = __floats2half2_rn(d[v].x * rstd, d[v].y * rstd);
            __half2 h1 = __floats2half2_rn(d[v].z * rstd, d[v].w * rstd);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(sA + kb * (FB_M * 128) + sw128_offset(r, cc >> 3) + (cc & 7) * 2) = u;
            if constexpr (kTrain) {
              if (split == 0 && t < p.T) *reinterpret_cast<uint2*>(p.xhat_out + (static_cast<long long>(b) * p.T + t) * C + c) = u;
            }
          }
          if constexpr (kTrain) {
            if (split == 0 && t < p.T && lane == 0) p.rstd_out[static_cast<long long>(b) * p.T + t] = rstd;
          }
        }
      }
      if constexpr (Cfg::A_TMEM) {
        // xhat tile: shared memory -> tensor memory.  A thread owns one row (TMEM lane) and two of the four 64-channel
        // k-blocks; the tile was written by other warps (rows are dealt differently in the prologue).
        asm volatile("bar.sync 1, %0;" ::"n"(FB_WORKERS * 32) : "memory");
        const int arow = q * 32 + lane;
#pragma unroll
        for (int kk = 0; kk < Cfg::KB / 2; ++kk) {
          const int kb = half * (Cfg::KB / 2) + kk;
          uint32_t ar[32];
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint4 u = *reinterpret_cast<const uint4*>(sA + kb * (FB_M * 128) + sw128_offset(arow, c16));
            ar[c16 * 4 + 0] = u.x; ar[c16 * 4 + 1] = u.y; ar[c16 * 4 + 2] = u.z; ar[c16 * 4 + 3] = u.w;
          }
          tmem_st_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + Cfg::A_COL + kb * 32, ar);
        }
        tmem_st_wait();
        tc_fence_before_sync();
      } else {
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor-core (async) proxy
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
      if (ww == 0 && lane == 0) FB_TRACE(2, 0);
    }
    // ---- per chunk: acc1 -> +bias -> GELU -> fp16 -> swizzled smem (A operand of pwconv2) ----
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float bias_cur[32], bias_nxt[32];   // the folded pwconv1 bias of this thread's 32 columns, fetched one chunk ahead (L1 is tiny here)
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_cur[i] = __ldg(p.b1 + ch_begin * FB_NC + half * 32 + i);
    for (int j = 0; j < n_ch; ++j) {
      const int buf = j & 1;
      if (j + 1 < n_ch) {
#pragma unroll
        for (int i = 0; i < 32; ++i) bias_nxt[i] = __ldg(p.b1 + (ch_begin + j + 1) * FB_NC + half * 32 + i);
      }
      mbar_wait(&acc1_full[buf], (j >> 1) & 1);
      if (ww == 0 && lane == 0) FB_TRACE(2, 1 + 5 * j);
      tc_fence_after_sync();
      uint32_t rr[32];
      tmem_ld_32x32(lane_addr + Cfg::ACC1_COL + buf * FB_NC + half * 32, rr);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[buf]);
      if (ww == 0 && lane == 0) FB_TRACE(2, 2 + 5 * j);
      float v[32];
      if constexpr (kTrain) {
        // pre-activation (fp16) -> HBM: this thread's 32 columns of its row are 64 contiguous bytes
        const int t = t0 + row;
        if (t < p.T) {
          __half* dst = p.pre_out + (static_cast<long long>(b) * p.T + t) * p.I + (ch_begin + j) * FB_NC + half * 32;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 u;
            __half2 h0 = __floats2half2_rn(__uint_as_float(rr[c4 * 8 + 0]) + bias_cur[c4 * 8 + 0], __uint_as_float(rr[c4 * 8 + 1]) + bias_cur[c4 * 8 + 1]);
            __half2 h1 = __floats2half2_rn(__uint_as_float(rr[c4 * 8 + 2]) + bias_cur[c4 * 8 + 2], __uint_as_float(rr[c4 * 8 + 3]) + bias_cur[c4 * 8 + 3]);
            __half2 h2 = __floats2half2_rn(__uint_as_float(rr[c4 * 8 + 4]) + bias_cur[c4 * 8 + 4], __uint_as_float(rr[c4 * 8 + 5]) + bias_cur[c4 * 8 + 5]);
            __half2 h3 = __floats2half2_rn(__uint_as_float(rr[c4 * 8 + 6]) + bias_cur[c4 * 8 + 6], __uint_as_float(rr[c4 * 8 + 7]) + bias_cur[c4 * 8 + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + c4 * 8) = u;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(__uint_as_float(rr[i]) + bias_cur[i]);
#pragma unroll
      for (int i = 0; i < 32; ++i) bias_cur[i] = bias_nxt[i];
      if (ww == 0 && lane == 0) FB_TRACE(2, 3 + 5 * j);
      mbar_wait(&h_empty[buf], ((j >> 1) & 1) ^ 1);
      if (ww == 0 && lane == 0) FB_TRACE(2, 4 + 5 * j);
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
        *reinterpret_cast<uint4*>(hrow + sw128_offset(row, half * 4 + c4)) = u;
        if constexpr (kTrain) {
          const int t = t0 + row;
          if (t < p.T)
            *reinterpret_cast<uint4*>(p.h_out + (static_cast<long long>(b) * p.T + t) * p.I + (ch_begin + j) * FB_NC + half * 32 + c4 * 8) = u;
        }
      }
      if (ww == 0 && lane == 0 && j < 8) FB_TRACE(2, 210 + j);
      fence_proxy_async_smem();
      if (ww == 0 && lane == 0 && j < 8) FB_TRACE(2, 220 + j);
      __syncwarp();
      if (lane == 0) mbar_arrive(&h_full[buf]);
      if (ww == 0 && lane == 0) FB_TRACE(2, 5 + 5 * j);
    }
    // ---- final epilogue: bias, layer scale, DropPath -> fp32 tile staged in shared memory (all MMA operands are dead
    //      now), then residual add + pad mask with fully coalesced row-wise global reads / writes ----
    // the residual rows of this warp are requested before the accumulator is awaited (their latency hides behind the tail)
    constexpr int EPR = FB_M / FB_WORKERS;    // rows per warp in the coalesced pass: r = ww + FB_WORKERS * i
    mbar_wait(acc2_full, 0);
    if (ww == 0 && lane == 0) FB_TRACE(2, 200);
    tc_fence_after_sync();
    float* stile = reinterpret_cast<float*>(smem);
    constexpr int OLD = Cfg::OUT_LD;
    const float rs = p.row_scale != nullptr ? p.row_scale[p.pair ? 2 * b + (row >> 6) : b] : 1.f;
    constexpr int CH = C / 2;  // columns per worker half
    for (int c0 = half * CH; c0 < (half + 1) * CH; c0 += 32) {
      uint32_t rr[32];
      tmem_ld_32x32(lane_addr + c0, rr);
      tmem_ld_wait();
      float* dst = stile + row * OLD + c0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + i));
        float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + c0 + i));
        if (split != 0) b4 = make_float4(0.f, 0.f, 0.f, 0.f);   // the pwconv2 bias enters once
        float4 o;
        o.x = g4.x * (__uint_as_float(rr[i + 0]) + b4.x) * rs;
        o.y = g4.y * (__uint_as_float(rr[i + 1]) + b4.y) * rs;
        o.z = g4.z * (__uint_as_float(rr[i + 2]) + b4.z) * rs;
        o.w = g4.w * (__uint_as_float(rr[i + 3]) + b4.w) * rs;
        *reinterpret_cast<float4*>(dst + i) = o;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(FB_WORKERS * 32) : "memory");   // worker warps only
    constexpr int EB = 4;                                    // rows in flight per warp (independent 512-byte loads)
#pragma unroll 1
    for (int i0 = 0; i0 < EPR; i0 += EB) {                   // one warp per row: 512-byte coalesced accesses
      float4 x4[EB][VPL];
      float keep[EB];
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        const int r = ww + FB_WORKERS * (i0 + e);
        const int t = t0 + r;
        const long long grow = static_cast<long long>(b) * p.T + t;
        keep[e] = (t < p.T && !(p.pad_mask != nullptr && p.pad_mask[grow])) ? 1.f : 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          x4[e][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (split == 0 && t < p.T) x4[e][v] = *reinterpret_cast<const float4*>(p.x + grow * C + v * 128 + lane * 4);   // so does the residual
        }
      }
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        const int r = ww + FB_WORKERS * (i0 + e);
        const int t = t0 + r;
        if (t >= p.T) continue;
        const long long grow = static_cast<long long>(b) * p.T + t;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int c = v * 128 + lane * 4;
          const float4 d4 = *reinterpret_cast<const float4*>(stile + r * OLD + c);
          float4 o;
          o.x = (x4[e][v].x + d4.x) * keep[e]; o.y = (x4[e][v].y + d4.y) * keep[e];
          o.z = (x4[e][v].z + d4.z) * keep[e]; o.w = (x4[e][v].w + d4.w) * keep[e];
          if (p.nsplit == 1) {
            *reinterpret_cast<float4*>(p.out + grow * C + c) = o;
          } else {  // partial sums of the splits meet in L2: one 16-byte vector reduction per thread, 512 B per warp
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out + grow * C + c), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w)
                         : "memory");
          }
        }
      }
    }
  }

  if (threadIdx.x == FB_CTRL * 32) FB_TRACE(2, 201);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int C, int I, bool kTrain>
int launch_fused(const void* w1_h16, const void* w2_h16, const FusedParams& p, cudaStream_t stream) {
  using Cfg = FusedCfg<C>;
  CUtensorMap tmW1, tmW2;
  // W1f: (I, C) K-major, box 64 x 64 ; W2: (C, I) K-major, box 64 x N2
  int rc = make_tmap_3d(&tmW1, w1_h16, TMA_F16, C, I, 1, C, static_cast<uint64_t>(I) * C, 64, FB_NC);
  if (rc != OSB_OK) return rc;
  rc = make_tmap_3d(&tmW2, w2_h16, TMA_F16, I, C, 1, I, static_cast<uint64_t>(C) * I, 64, Cfg::N2);
  if (rc != OSB_OK) return rc;
  // input (B, T, C) fp32: boxes of 32 channels (128 B) x XR rows; rows outside [0, T) read as zero (Conv1d zero padding)
  CUtensorMap tmX;
  const int T_true = p.pair ? p.T / 2 : p.T, B_true = p.pair ? p.B * 2 : p.B;
  rc = make_tmap_3d(&tmX, p.x, TMA_F32, C, T_true, B_true, C, static_cast<uint64_t>(T_true) * C, 32, Cfg::XR);
  if (rc != OSB_OK) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(convnext_fused_kernel<C, I, kTrain>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  if (p.nsplit > 1) {  // the splits accumulate into `out`
    cudaError_t e = cudaMemsetAsync(p.out, 0, static_cast<size_t>(p.B) * p.T * C * sizeof(float), stream);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  convnext_fused_kernel<C, I, kTrain><<<dim3(p.B * p.m_tiles, p.nsplit), FB_THREADS, Cfg::SMEM, stream>>>(tmW1, tmW2, tmX, p);
  count_launch();
  return launch_status();
}

}  // namespace
}  // namespace osb

using namespace osb;

static long long* g_fused_trace = nullptr;
static int g_fused_nsplit = 0;
static int g_fused_nsplit_cap = 4;
extern "C" void osb_debug_set_fused_nsplit_cap(int n) { g_fused_nsplit_cap = n > 0 ? n : 4; }
/* developer hook (not in the public header): force the intermediate-dimension split of the fused block (0 = automatic) */
extern "C" void osb_debug_set_fused_nsplit(int n) { g_fused_nsplit = n; }
/* developer hook (not in the public header): device buffer of 3*256 int64 receiving a clock64 timeline of CTA 0 */
extern "C" void osb_debug_set_fused_trace(long long* buf) { g_fused_trace = buf; }

static int fused_fwd_impl(const float* x, const float* dw_w, const float* dw_b, const void* w1f_h16, const float* b1f, const void* w2_h16,
                          const float* b2, const float* gamma, const float* row_scale, const uint8_t* pad_mask, float* out, void* xhat_out,
                          float* rstd_out, void* pre_out, void* h_out, int32_t B, int32_t T, int32_t C, int32_t I, float eps, void* stream) {
  OSB_REQUIRE(x && dw_w && dw_b && w1f_h16 && b1f && w2_h16 && b2 && gamma && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0, OSB_ERR_SHAPE);
  const bool train = xhat_out != nullptr;
  if (train) OSB_REQUIRE(rstd_out && pre_out && h_out, OSB_ERR_ARG);
  FusedParams p;
  p.x = x; p.dw_w = dw_w; p.dw_b = dw_b; p.b1 = b1f; p.b2 = b2; p.gamma = gamma; p.row_scale = row_scale; p.pad_mask = pad_mask;
  // Training vocoder: 64-frame segments.  One sample per 128-row tile would leave every tile half empty (twice the tensor and
  // epilogue work per useful row); two consecutive samples share a tile instead — their rows are contiguous in (B, 64, C).
  const bool pair = (C == 384 && T == 64 && B % 2 == 0);
  p.pair = pair ? 1 : 0;
  if (pair) { B /= 2; T = 128; }
  p.out = out; p.B = B; p.T = T; p.m_tiles = (T + FB_M - 1) / FB_M; p.eps = eps; p.I = I;
  p.xhat_out = static_cast<__half*>(xhat_out); p.rstd_out = rstd_out; p.pre_out = static_cast<__half*>(pre_out);
  p.h_out = static_cast<__half*>(h_out);
  // Few row tiles (less than half the SMs): split the intermediate dimension over blockIdx.y so that more SMs share the chunk
  // loop — 148 / tiles splits, at most 4: beyond that the prologue (replicated per split) and the L2 reductions cost what the
  // shorter loops save (16 tiles: 4 splits take the same 45 us as 9), and the SMs left free run the other branches of the
  // step (inside the captured step 48 tiles run 46 us with 3 splits, 63 us with 2).
  {
    const int tiles = B * p.m_tiles;
    const int nch = I / FB_NC;
    int ns = g_fused_nsplit > 0 ? g_fused_nsplit : (tiles * 2 > 148 ? 1 : (148 / tiles > g_fused_nsplit_cap ? g_fused_nsplit_cap : 148 / tiles));
    if (ns > nch / 2) ns = nch / 2;
    p.nsplit = ns < 1 ? 1 : ns;
  }
  p.trace = g_fused_trace;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 256 && I == 1024) return train ? launch_fused<256, 1024, true>(w1f_h16, w2_h16, p, s) : launch_fused<256, 1024, false>(w1f_h16, w2_h16, p, s);
  if (C == 384 && I == 1152) return train ? launch_fused<384, 1152, true>(w1f_h16, w2_h16, p, s) : launch_fused<384, 1152, false>(w1f_h16, w2_h16, p, s);
  return OSB_ERR_SHAPE;
}

extern "C" int osb_convnext_block_fwd(const float* x, const float* dw_w, const float* dw_b, const void* w1f_h16, const float* b1f,
                                      const void* w2_h16, const float* b2, const float* gamma, const float* row_scale,
                                      const uint8_t* pad_mask, float* out, int32_t B, int32_t T, int32_t C, int32_t I, float eps,
                                      void* stream) {
  return fused_fwd_impl(x, dw_w, dw_b, w1f_h16, b1f, w2_h16, b2, gamma, row_scale, pad_mask, out, nullptr, nullptr, nullptr, nullptr, B, T, C,
                        I, eps, stream);
}

extern "C" int osb_convnext_block_fwd_train(const float* x, const float* dw_w, const float* dw_b, const void* w1f_h16, const float* b1f,
                                            const void* w2_h16, const float* b2, const float* gamma, const float* row_scale,
                                            const uint8_t* pad_mask, float* out, void* xhat_h16, float* rstd, void* pre_h16, void* h_h16,
                                            int32_t B, int32_t T, int32_t C, int32_t I, float eps, void* stream) {
  OSB_REQUIRE(xhat_h16 && rstd && pre_h16 && h_h16, OSB_ERR_ARG);
  return fused_fwd_impl(x, dw_w, dw_b, w1f_h16, b1f, w2_h16, b2, gamma, row_scale, pad_mask, out, xhat_h16, rstd, pre_h16, h_h16, B, T, C, I,
                        eps, stream);
}
