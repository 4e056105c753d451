// osb_align.cu — alignment-learning kernels that the reference runs on the host CPU (numba):
// monotonic alignment search (Viterbi over the attention log-probabilities) and duration-span
// averaging, moved onto the device so a training step has no per-sample host round trips.
#include <cstdlib>

#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

// ------------------------------------------------------------------------------------------
// Monotonic alignment search, one CTA per sample, one thread per text token.
//   Q[0,j] = float32 running sum of lp[0..j, 0], widened to double (what the reference's numba code
//            computes for `log_prob[0, :j+1].sum()` on a float32 array)
//   Q[i,j] = max(Q[i-1,j-1], Q[i,j-1]) + lp[j,i]          (double), 1 <= i < min(j+1, N)
//   back-track from A[T-1] = N-1 with the '>=' tie rule (prefer the lower token).
// The decision bit (Q[i-1,j-1] >= Q[i,j-1]) of every cell is kept in shared memory (T x ceil(N/32) words).
// ------------------------------------------------------------------------------------------
__global__ void mas_kernel(const float* __restrict__ lp, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                           int* __restrict__ path, float* __restrict__ dur, int Tm, int Tx) {
  extern __shared__ unsigned char mas_smem[];
  const int b = blockIdx.x;
  const int N = static_cast<int>(x_len[b]);
  const int T = static_cast<int>(m_len[b]);
  const int i = threadIdx.x;
  const int nthr = blockDim.x;
  const int words = (nthr + 31) >> 5;
  double* qbuf = reinterpret_cast<double*>(mas_smem);                       // [2][nthr]
  unsigned* flags = reinterpret_cast<unsigned*>(qbuf + 2 * nthr);           // [T][words]
  int* cnt = reinterpret_cast<int*>(flags + static_cast<size_t>(T > 0 ? T : 1) * words);  // [nthr]
  const float* lpb = lp + static_cast<long long>(b) * Tm * Tx;
  int* pb = path + static_cast<long long>(b) * Tm;

  cnt[i] = 0;
  if (N <= 0 || T <= 0) {
    for (int t = i; t < Tm; t += nthr) pb[t] = -1;
    if (i < Tx) dur[static_cast<long long>(b) * Tx + i] = 0.f;
    return;
  }
  const double NEG = -INFINITY;
  float row0 = 0.f;
  double q = NEG;
  if (i == 0) {
    row0 = lpb[0];
    q = static_cast<double>(row0);
  }
  qbuf[i] = q;
  __syncthreads();
  // the log-probability of the next frame is fetched one iteration ahead: the L2 latency of that load would
  // otherwise sit on the serial critical path of the recursion
  constexpr int G = 8;
  float lpc[G], lpn[G];
#pragma unroll
  for (int g = 0; g < G; ++g) lpc[g] = (1 + g < T && i < N) ? lpb[static_cast<long long>(1 + g) * Tx + i] : 0.f;
  for (int base = 1; base < T; base += G) {
#pragma unroll
    for (int g = 0; g < G; ++g) lpn[g] = (base + G + g < T && i < N) ? lpb[static_cast<long long>(base + G + g) * Tx + i] : 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = base + g;
      if (j < T) {  // uniform across the block
        const int cur = j & 1, prev = cur ^ 1;
        const float lp_cur = lpc[g];
        const double left = (i > 0) ? qbuf[prev * nthr + i - 1] : NEG;
        const bool take_left = (i > 0) && (left >= q);
        const unsigned bits = __ballot_sync(0xffffffffu, take_left);
        if ((i & 31) == 0) flags[static_cast<size_t>(j) * words + (i >> 5)] = bits;
        if (i == 0) {
          row0 = __fadd_rn(row0, lp_cur);
          q = static_cast<double>(row0);
        } else if (i < N && i <= j) {
          q = fmax(left, q) + static_cast<double>(lp_cur);
        }
        qbuf[cur * nthr + i] = q;
        __syncthreads();
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) lpc[g] = lpn[g];
  }
  if (i == 0) {
    int a = N - 1;
    pb[T - 1] = a;
    for (int j = T - 2; j >= 0; --j) {
      if (a > 0) {
        const unsigned w = flags[static_cast<size_t>(j + 1) * words + (a >> 5)];
        a -= (w >> (a & 31)) & 1u;
      }
      pb[j] = a;
    }
  }
  __syncthreads();
  for (int t = i; t < Tm; t += nthr) {
    if (t < T) atomicAdd(&cnt[pb[t]], 1);
    else pb[t] = -1;
  }
  __syncthreads();
  if (i < Tx) dur[static_cast<long long>(b) * Tx + i] = static_cast<float>(cnt[i]);
}

// ------------------------------------------------------------------------------------------
// Same search with ONE WARP per sample: lane l owns the VPL consecutive tokens [l*VPL, (l+1)*VPL) and keeps their Q values in
// registers, so that a frame costs one 64-bit shuffle (the left neighbour of the lane's first token) instead of a shared-memory
// round trip plus a block barrier.  The serial chain per frame drops from ~440 to ~100 cycles.  Arithmetic, tie rule and the
// float32 running sum of row 0 are those of mas_kernel (bit-identical paths).  Decision bits: one word per (frame, lane).
// ------------------------------------------------------------------------------------------
template <int VPL>
struct MasFlag { using type = unsigned int; };
template <> struct MasFlag<2> { using type = unsigned char; };
template <> struct MasFlag<4> { using type = unsigned char; };
template <> struct MasFlag<6> { using type = unsigned char; };
template <> struct MasFlag<8> { using type = unsigned char; };
template <> struct MasFlag<12> { using type = unsigned short; };
template <> struct MasFlag<16> { using type = unsigned short; };

// W warps per sample (round 2): the search is issue bound on one warp (fp64 adds / compares / selects of VPL tokens per lane
// at ~0.45 instructions per cycle), so the tokens are spread over W warps with VPL = ceil(Tx / (32 W)) per lane; the one
// fp64 value that crosses a warp boundary per frame goes through a double-buffered shared-memory mailbox and one block
// barrier per frame.  W = 1 keeps the single-warp form (no barrier).
template <int VPL, int W>
__global__ void __launch_bounds__(32 * W)
mas_warp_kernel(const float* __restrict__ lp, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                int* __restrict__ path, float* __restrict__ dur, int Tm, int Tx) {
  using FlagT = typename MasFlag<VPL>::type;
  constexpr int NL = 32 * W;                              // lanes per sample
  constexpr int G = VPL <= 8 ? 8 : (VPL <= 16 ? 4 : 2);  // frames per prefetch batch
  extern __shared__ unsigned char mas_smem[];
  __shared__ double mbox[2][W];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int wip = threadIdx.x >> 5;
  const int gl = threadIdx.x;                              // lane index inside the sample
  const int N = static_cast<int>(x_len[b]);
  const int T = static_cast<int>(m_len[b]);
  int* pth = reinterpret_cast<int*>(mas_smem);                               // [Tm]
  int* cnt = pth + Tm;                                                       // [Tx]
  FlagT* flags = reinterpret_cast<FlagT*>(cnt + Tx);                         // [T][NL]
  const float* lpb = lp + static_cast<long long>(b) * Tm * Tx;
  int* pb = path + static_cast<long long>(b) * Tm;
  for (int n = gl; n < Tx; n += NL) cnt[n] = 0;
  if (N <= 0 || T <= 0) {
    for (int t = gl; t < Tm; t += NL) pb[t] = -1;
    for (int n = gl; n < Tx; n += NL) dur[static_cast<long long>(b) * Tx + n] = 0.f;
    return;
  }
  const double NEG = -INFINITY;
  const int c0 = gl * VPL;
  double q[VPL];
  int lim[VPL];    // token i = c0 + k takes part from frame j >= lim[k] on (i <= j), never when i >= N; token 0 is the float32 row sum
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    q[k] = NEG;
    lim[k] = (c0 + k < N && c0 + k > 0) ? c0 + k : 0x7fffffff;
  }
  float row0 = 0.f;
  if (gl == 0) {
    row0 = lpb[0];
    q[0] = static_cast<double>(row0);
  }
  const bool first = gl == 0;
  if (W > 1) {
    if (lane == 31) mbox[0][wip] = q[VPL - 1];
    __syncthreads();
  }
  // Frames are fetched G at a time, one whole batch ahead (two register batches): the wait at the top of a batch then only
  // covers loads issued a full batch earlier.  (Per-frame refills share hardware scoreboards with the newest loads and
  // serialise on the full memory latency every frame.)
  float cur[G][VPL], nxt[G][VPL];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int k = 0; k < VPL; ++k) cur[g][k] = (1 + g < T && c0 + k < N) ? lpb[static_cast<long long>(1 + g) * Tx + c0 + k] : 0.f;

  for (int base = 1; base < T; base += G) {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        nxt[g][k] = (base + G + g < T && c0 + k < N) ? lpb[static_cast<long long>(base + G + g) * Tx + c0 + k] : 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = base + g;
      if (j < T) {  // uniform across the block
        double up = __shfl_up_sync(0xffffffffu, q[VPL - 1], 1);
        if (lane == 0) up = (W > 1 && wip > 0) ? mbox[(j - 1) & 1][wip - 1] : NEG;
        unsigned bits = 0;
        row0 = __fadd_rn(row0, cur[g][0]);
#pragma unroll
        for (int k = VPL - 1; k >= 0; --k) {  // descending: q[k-1] still holds the previous frame
          const double left = k > 0 ? q[k - 1] : up;
          const bool take_left = left >= q[k];
          bits |= (take_left ? 1u : 0u) << k;
          // max(left, q) + lp == max(left + lp, q + lp) in IEEE arithmetic (rounding is monotonic): the two sums do not wait
          // for the comparison
          const double lpd = static_cast<double>(cur[g][k]);
          const double via_left = left + lpd, via_self = q[k] + lpd;
          if (j >= lim[k]) q[k] = take_left ? via_left : via_self;
        }
        if (first) {
          q[0] = static_cast<double>(row0);
          bits &= ~1u;  // token 0 has no left neighbour
        }
        flags[static_cast<size_t>(j) * NL + gl] = static_cast<FlagT>(bits);
        if (W > 1) {
          if (lane == 31) mbox[j & 1][wip] = q[VPL - 1];
          __syncthreads();
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int k = 0; k < VPL; ++k) cur[g][k] = nxt[g][k];
  }
  __syncthreads();
  if (gl == 0) {
    int a = N - 1;
    pth[T - 1] = a;
    for (int j = T - 2; j >= 0; --j) {
      if (a > 0) {
        const unsigned w = flags[static_cast<size_t>(j + 1) * NL + a / VPL];
        a -= (w >> (a % VPL)) & 1u;
      }
      pth[j] = a;
    }
  }
  __syncthreads();
  for (int t = gl; t < Tm; t += NL) {
    if (t < T) {
      const int a = pth[t];
      pb[t] = a;
      atomicAdd(&cnt[a], 1);
    } else {
      pb[t] = -1;
    }
  }
  __syncthreads();
  for (int n = gl; n < Tx; n += NL) dur[static_cast<long long>(b) * Tx + n] = static_cast<float>(cnt[n]);
}

template <int VPL, int W>
int launch_mas_warp(const float* lp, const long long* x_len, const long long* m_len, int* path, float* dur, int B, int Tm, int Tx,
                    cudaStream_t stream) {
  using FlagT = typename MasFlag<VPL>::type;
  const size_t smem = sizeof(int) * (static_cast<size_t>(Tm) + Tx) + sizeof(FlagT) * 32 * W * static_cast<size_t>(Tm);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mas_warp_kernel<VPL, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = 200 * 1024;
  }
  mas_warp_kernel<VPL, W><<<B, 32 * W, smem, stream>>>(lp, x_len, m_len, path, dur, Tm, Tx);
  count_launch();
  return launch_status();
}

// ------------------------------------------------------------------------------------------
// average_by_duration: token-level mean of a frame-level feature over each duration span.
// One CTA per sample; span sums are sequential float32 (numba's accumulation order), the division
// is done in double and rounded to float32 (numba: float32 / int64 -> float64).
// ------------------------------------------------------------------------------------------
__global__ void average_by_duration_kernel(const float* __restrict__ ds, const float* __restrict__ xs, const long long* __restrict__ x_len,
                                           const long long* __restrict__ m_len, float* __restrict__ out, int Tm, int Tx) {
  extern __shared__ int start_s[];  // Tx + 1
  const int b = blockIdx.x;
  const int N = static_cast<int>(x_len[b]);
  const int T = static_cast<int>(m_len[b]);
  if (threadIdx.x == 0) {
    int acc = 0;
    start_s[0] = 0;
    for (int n = 0; n < N; ++n) {
      acc += static_cast<int>(ds[static_cast<long long>(b) * Tx + n]);
      start_s[n + 1] = acc;
    }
  }
  __syncthreads();
  for (int n = threadIdx.x; n < Tx; n += blockDim.x) {
    float v = 0.f;
    if (n < N) {
      const int s = min(start_s[n], T), e = min(start_s[n + 1], T);
      if (e > s) {
        float c = 0.f;
        for (int t = s; t < e; ++t) c = __fadd_rn(c, xs[static_cast<long long>(b) * Tm + t]);
        v = static_cast<float>(static_cast<double>(c) / static_cast<double>(e - s));
      }
    }
    out[static_cast<long long>(b) * Tx + n] = v;
  }
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_mas(const float* log_p_attn, const int64_t* x_len, const int64_t* m_len, int32_t* path, float* durations,
                       int32_t B, int32_t Tm, int32_t Tx, void* stream) {
  OSB_REQUIRE(log_p_attn && x_len && m_len && path && durations, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0 && Tx <= 1024, OSB_ERR_SHAPE);
  {  // one warp per sample whenever its decision bits fit in shared memory
    const int vpl = (Tx + 31) / 32;
    const int flag_bytes = vpl <= 8 ? 1 : (vpl <= 16 ? 2 : 4);
    const size_t need = sizeof(int) * (static_cast<size_t>(Tm) + Tx) + static_cast<size_t>(flag_bytes) * 32 * Tm;
    static const bool force_block = getenv("OSB_MAS_BLOCK") != nullptr;  // developer switch: time the one-CTA-per-sample kernel
    if (need <= 200 * 1024 && !force_block) {
      const long long* xl = reinterpret_cast<const long long*>(x_len);
      const long long* ml = reinterpret_cast<const long long*>(m_len);
      cudaStream_t st = static_cast<cudaStream_t>(stream);
      // four warps per sample (2-4 tokens per lane) while their decision bytes fit shared memory; else one warp
      static const bool one_warp = getenv("OSB_MAS_ONE_WARP") != nullptr;   // developer switch
      const size_t need4 = sizeof(int) * (static_cast<size_t>(Tm) + Tx) + static_cast<size_t>(128) * Tm;
      if (!one_warp && Tx <= 512 && need4 <= 200 * 1024) {
        if (Tx <= 256) return launch_mas_warp<2, 4>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
        return launch_mas_warp<4, 4>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      }
      if (vpl <= 2) return launch_mas_warp<2, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 4) return launch_mas_warp<4, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 6) return launch_mas_warp<6, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 8) return launch_mas_warp<8, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 12) return launch_mas_warp<12, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 16) return launch_mas_warp<16, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      if (vpl <= 24) return launch_mas_warp<24, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
      return launch_mas_warp<32, 1>(log_p_attn, xl, ml, path, durations, B, Tm, Tx, st);
    }
  }
  const int nthr = ((Tx + 31) / 32) * 32;
  const size_t smem = sizeof(double) * 2 * nthr + sizeof(unsigned) * static_cast<size_t>(Tm) * (nthr / 32) + sizeof(int) * nthr;
  OSB_REQUIRE(smem <= 200 * 1024, OSB_ERR_SHAPE);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(mas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = 200 * 1024;
  }
  mas_kernel<<<B, nthr, smem, static_cast<cudaStream_t>(stream)>>>(log_p_attn, reinterpret_cast<const long long*>(x_len),
                                                                 reinterpret_cast<const long long*>(m_len), path, durations, Tm, Tx);
  count_launch();
  return launch_status();
}

extern "C" int osb_average_by_duration(const float* ds, const float* xs, const int64_t* x_len, const int64_t* m_len, float* out,
                                       int32_t B, int32_t Tm, int32_t Tx, void* stream) {
  OSB_REQUIRE(ds && xs && x_len && m_len && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0 && Tx <= 8192, OSB_ERR_SHAPE);
  average_by_duration_kernel<<<B, 256, sizeof(int) * (Tx + 1), static_cast<cudaStream_t>(stream)>>>(
      ds, xs, reinterpret_cast<const long long*>(x_len), reinterpret_cast<const long long*>(m_len), out, Tm, Tx);
  count_launch();
  return launch_status();
}

// ==========================================================================================
// Attention (pairwise distance + log-softmax) helpers: squared row norms, backward preparation,
// batched transpose-pack.  The contractions themselves run on the tcgen05 kernels of osb_gemm.cu.
// ==========================================================================================
namespace osb {
namespace {

// out[row] = sum_c x[row,c]^2, one warp per row
__global__ void rownorm_sq_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int C) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(x + row * C + c);
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// Backward of  lp = log_softmax_n(score) + prior,  score = -dist  with respect to dist, expressed as the matrix
//   Wn[t,n] = -(dscore[t,n] / score[t,n]),  dscore = G - softmax * rowsum(G)
// so that  dF = -rsn * F + Wn @ E  and  dE = -csn * E + Wn^T @ F  (rsn / csn = row / column sums of Wn).
// One warp per (b,t) row.  score is recovered as lp - prior + lse.
__global__ void attn_bwd_prep_kernel(const float* __restrict__ G, const float* __restrict__ lp, const float* __restrict__ prior,
                                     const float* __restrict__ lse, const long long* __restrict__ x_len,
                                     const long long* __restrict__ m_len, __half* __restrict__ Wn, float* __restrict__ neg_rsn,
                                     float* __restrict__ csn, int B, int Tm, int Tx, int ldw) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * Tm) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row % Tm);
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  __half* wrow = Wn + row * ldw;
  if (t >= T) {
    for (int n = lane; n < ldw; n += 32) wrow[n] = __float2half_rn(0.f);
    if (lane == 0) neg_rsn[row] = 0.f;
    return;
  }
  const float l = lse[row];
  float gsum = 0.f;
  for (int n = lane; n < N; n += 32) gsum += G[row * Tx + n];
  gsum = warp_sum(gsum);
  float rs = 0.f;
  for (int n = lane; n < ldw; n += 32) {
    float w = 0.f;
    if (n < N) {
      const float logp = lp[row * Tx + n] - prior[row * Tx + n];  // log softmax
      const float score = logp + l;                               // = -dist
      const float ds = G[row * Tx + n] - expf(logp) * gsum;
      w = score < 0.f ? -(ds / score) : 0.f;
      atomicAdd(csn + static_cast<long long>(b) * Tx + n, w);
    }
    rs += w;
    wrow[n] = __float2half_rn(w);
  }
  rs = warp_sum(rs);
  if (lane == 0) neg_rsn[row] = -rs;
}

// out[b, c, t] = fp16(x[b, t, c]) for t < T, zero for T <= t < Tp   (32x32 shared-memory tiles)
__global__ void transpose_pack_kernel(const float* __restrict__ x, __half* __restrict__ out, int T, int C, int Tp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + tx;
    tile[j][tx] = (t < T && c < C) ? x[(static_cast<long long>(b) * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + tx;
    if (c < C && t < Tp) out[(static_cast<long long>(b) * C + c) * Tp + t] = __float2half_rn(tile[tx][j]);
  }
}

// dE[b,n,:] = -csn[b,n] * E[b,n,:]   (the wgrad contraction then accumulates Wn^T @ F on top)
__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ row_scale, float* __restrict__ out, long long rows,
                                  int C, float sign) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  out[i] = sign * row_scale[i / C] * x[i];
}

}  // namespace
}  // namespace osb

namespace osb {
namespace {
// log BetaBinomial(k = n; N, a = t+1, b = T-t) for frame t < T and token n < N of each sample, -inf elsewhere.
// Every gamma-function argument of the scipy definition is an integer, so lf[m] = log(m!) (double) is exact.
__global__ void beta_binomial_prior_kernel(const double* __restrict__ lf, const long long* __restrict__ x_len,
                                           const long long* __restrict__ m_len, float* __restrict__ out, int Tm, int Tx) {
  const int b = blockIdx.z;
  const int t = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Tx) return;
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  float v = -INFINITY;
  if (t < T && n < N) {
    const int a = t + 1;  // alpha = t (1-based), beta = T - alpha + 1
    v = static_cast<float>(lf[N] - lf[n] - lf[N - n] + lf[n + a - 1] + lf[N - n + T - a] - lf[N + T] - lf[a - 1] - lf[T - a] + lf[T]);
  }
  out[(static_cast<long long>(b) * Tm + t) * Tx + n] = v;
}
}  // namespace
}  // namespace osb

extern "C" int osb_beta_binomial_prior(const double* log_factorial, int64_t table_len, const int64_t* x_len, const int64_t* m_len,
                                       float* out, int32_t B, int32_t Tm, int32_t Tx, void* stream) {
  OSB_REQUIRE(log_factorial && x_len && m_len && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0 && table_len >= static_cast<int64_t>(Tm) + Tx + 1, OSB_ERR_SHAPE);
  dim3 grid((Tx + 127) / 128, Tm, B);
  beta_binomial_prior_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      log_factorial, reinterpret_cast<const long long*>(x_len), reinterpret_cast<const long long*>(m_len), out, Tm, Tx);
  count_launch();
  return launch_status();
}

extern "C" int osb_rownorm_sq(const float* x, float* out, int64_t rows, int32_t C, void* stream) {
  OSB_REQUIRE(x && out, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, OSB_ERR_SHAPE);
  rownorm_sq_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, rows, C);
  count_launch();
  return launch_status();
}

extern "C" int osb_attn_bwd_prep(const float* G, const float* lp, const float* prior, const float* lse, const int64_t* x_len,
                                 const int64_t* m_len, void* wn_h16, float* neg_rsn, float* csn, int32_t B, int32_t Tm, int32_t Tx,
                                 int32_t ldw, void* stream) {
  OSB_REQUIRE(G && lp && prior && lse && x_len && m_len && wn_h16 && neg_rsn && csn, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0 && ldw >= Tx && ldw % 8 == 0, OSB_ERR_SHAPE);
  const long long rows = static_cast<long long>(B) * Tm;
  attn_bwd_prep_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      G, lp, prior, lse, reinterpret_cast<const long long*>(x_len), reinterpret_cast<const long long*>(m_len),
      static_cast<__half*>(wn_h16), neg_rsn, csn, B, Tm, Tx, ldw);
  count_launch();
  return launch_status();
}

extern "C" int osb_transpose_pack_h16(const float* x, void* out_h16, int32_t B, int32_t T, int32_t C, int32_t Tp, void* stream) {
  OSB_REQUIRE(x && out_h16, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0 && Tp >= T, OSB_ERR_SHAPE);
  dim3 grid((Tp + 31) / 32, (C + 31) / 32, B);
  transpose_pack_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__half*>(out_h16), T, C, Tp);
  count_launch();
  return launch_status();
}

extern "C" int osb_scale_rows(const float* x, const float* row_scale, float* out, int64_t rows, int32_t C, float sign, void* stream) {
  OSB_REQUIRE(x && row_scale && out, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C > 0, OSB_ERR_SHAPE);
  const long long n = rows * C;
  scale_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, row_scale, out, rows, C, sign);
  count_launch();
  return launch_status();
}

// ==========================================================================================
// Forward-sum alignment loss: CTC with the identity target (token n at position n), a constant
// blank logit, per-sample re-normalisation.  One CTA per sample, thread k owns the blank state 2k and
// the token state 2k+1 of the extended target; alpha is kept in a global workspace for the backward
// sweep, which runs in the same kernel and emits d(loss)/d(log_p_attn).
// ==========================================================================================
namespace osb {
namespace {

// log-sum-exp of the recursion: the arguments of exp are <= 0 and the sum lies in [1, 3], where the hardware
// ex2 / lg2 approximations are accurate to ~1e-7 absolute; this halves the length of the serial dependency chain.
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  return m == -INFINITY ? -INFINITY : m + __logf(__expf(a - m) + __expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  return m == -INFINITY ? -INFINITY : m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// per-frame normaliser over [blank | tokens < N]: one warp per (b, t) row, full grid
__global__ void fs_lse_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                              float blank_logit, float* __restrict__ lse_ws, int B, int Tm, int Tx) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * Tm) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row % Tm);
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  if (t >= T) {
    if (lane == 0) lse_ws[row] = 0.f;
    return;
  }
  const float* lp = lpa + row * Tx;
  float m = blank_logit;
  for (int i = lane; i < N; i += 32) m = fmaxf(m, lp[i]);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += expf(lp[i] - m);
  s = warp_sum(s) + expf(blank_logit - m);
  if (lane == 0) lse_ws[row] = m + logf(s);
}

constexpr int FS_G = 8;  // emissions are fetched FS_G frames ahead (registers) so that L2 latency stays off the serial chain

// The serial part: alpha (t = 1 .. T-1) and beta (t = T-2 .. 0) advance in the SAME iteration, on two halves of the CTA
// (threads [0, nthr) own the alpha chain, threads [nthr, 2 nthr) the beta chain: a thread carries ONE dependency chain of
// exp / log per step, the two recursions only share the barrier).  One CTA per sample; within a half, thread k owns blank
// state 2k and token state 2k+1.
__global__ void forward_sum_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                                   float blank_logit, const float* __restrict__ lse_ws, float* __restrict__ aw_all,
                                   float* __restrict__ bw_all, float* __restrict__ nll_ws, float* __restrict__ loss, int B, int Tm, int Tx,
                                   int nsplit_halves) {
  extern __shared__ float fs_smem[];
  const int b = blockIdx.x;
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  const int S = 2 * N + 1;
  const int Smax = 2 * Tx + 1;
  // split = 1 (2 (Tx + 1) threads fit a CTA): the two chains run side by side; otherwise every thread walks alpha, then beta
  const int split = nsplit_halves;
  const int half = split ? (blockDim.x >> 1) : blockDim.x;
  float* abuf0 = fs_smem;
  float* abuf1 = abuf0 + Smax;
  float* bbuf0 = abuf1 + Smax;
  float* bbuf1 = bbuf0 + Smax;
  float* lse = bbuf1 + Smax;  // [Tm]
  const float* lp = lpa + static_cast<long long>(b) * Tm * Tx;
  float* aw = aw_all + static_cast<long long>(b) * Tm * Tx;
  float* bw = bw_all + static_cast<long long>(b) * Tm * Tx;
  const float NEG = -INFINITY;
  if (N <= 0 || T <= 0) {
    if (threadIdx.x == 0) { loss[b] = 0.f; nll_ws[b] = INFINITY; }
    return;
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) lse[t] = lse_ws[static_cast<long long>(b) * Tm + t];
  __syncthreads();

  for (int pass = 0; pass < (split ? 1 : 2); ++pass) {
  const bool is_beta = split ? (static_cast<int>(threadIdx.x) >= half) : (pass == 1);
  const int k = static_cast<int>(threadIdx.x) - ((split && is_beta) ? half : 0);
  const bool has_tok = k < N;
  const bool active = k <= N;
  // this half's state buffers: cp = previous step, cc = current step
  float* cp = is_beta ? bbuf0 : abuf0;
  float* cc = is_beta ? bbuf1 : abuf1;
  float* ws = is_beta ? bw : aw;
  {
    const float l0 = lse[0], lT = lse[T - 1];
    if (active) {
      if (!is_beta) {
        cp[2 * k] = (k == 0) ? blank_logit - l0 : NEG;
        if (has_tok) {
          const float at = (k == 0) ? lp[0] - l0 : NEG;
          cp[2 * k + 1] = at;
          ws[k] = at;
        }
      } else {
        cp[2 * k] = (k == N) ? blank_logit - lT : NEG;
        if (has_tok) {
          const float bt = (k == N - 1) ? lp[static_cast<long long>(T - 1) * Tx + k] - lT : NEG;
          cp[2 * k + 1] = bt;
          ws[static_cast<long long>(T - 1) * Tx + k] = bt;
        }
      }
    }
  }
  __syncthreads();
  // frame visited at step i: alpha walks forward, beta backward
  auto frame = [&](int i) { return is_beta ? T - 1 - i : i; };
  float e[FS_G], e_n[FS_G];
#pragma unroll
  for (int g = 0; g < FS_G; ++g) {
    const int i = 1 + g;
    e[g] = (has_tok && i < T) ? lp[static_cast<long long>(frame(i)) * Tx + k] : 0.f;
  }
  for (int base = 1; base < T; base += FS_G) {
#pragma unroll
    for (int g = 0; g < FS_G; ++g) {
      const int i = base + FS_G + g;
      e_n[g] = (has_tok && i < T) ? lp[static_cast<long long>(frame(i)) * Tx + k] : 0.f;
    }
#pragma unroll
    for (int g = 0; g < FS_G; ++g) {
      const int i = base + g;
      if (i < T) {  // uniform across the block
        const int tf = frame(i);
        const float lf = lse[tf];
        if (active) {
          if (!is_beta) {
            const float pb = cp[2 * k];
            const float pm1 = k > 0 ? cp[2 * k - 1] : NEG;
            cc[2 * k] = lse2(pb, pm1) + (blank_logit - lf);
            if (has_tok) {
              const float nt = lse3(cp[2 * k + 1], pb, pm1) + (e[g] - lf);
              cc[2 * k + 1] = nt;
              ws[static_cast<long long>(tf) * Tx + k] = nt;
            }
          } else {
            const float qb = cp[2 * k];
            const float qt = has_tok ? cp[2 * k + 1] : NEG;
            cc[2 * k] = lse2(qb, qt) + (blank_logit - lf);
            if (has_tok) {
              const float mt = lse3(qt, cp[2 * k + 2], (k + 1 < N) ? cp[2 * k + 3] : NEG) + (e[g] - lf);
              cc[2 * k + 1] = mt;
              ws[static_cast<long long>(tf) * Tx + k] = mt;
            }
          }
        }
        __syncthreads();
        float* t0 = cp; cp = cc; cc = t0;
      }
    }
#pragma unroll
    for (int g = 0; g < FS_G; ++g) e[g] = e_n[g];
  }
  if (threadIdx.x == 0 && !is_beta) {   // alpha chain, k = 0: cp is the last alpha column
    const float ll = lse2(cp[S - 1], S >= 2 ? cp[S - 2] : NEG);
    const bool finite = ll > -INFINITY && ll < INFINITY;
    loss[b] = finite ? -ll / static_cast<float>(N) : 0.f;   // reduction='mean' (target length), zero_infinity
    nll_ws[b] = finite ? -ll : INFINITY;
  }
  __syncthreads();
  }  // pass
}

// ------------------------------------------------------------------------------------------
// Warp-synchronous forward-sum recursion (round 2).  The 2N+1 states of the extended target live in REGISTERS: lane L owns the
// SPL consecutive states s = SPL*L + j (SPL even, so register j is a blank state for even j and a token state for odd j, known
// at compile time); a step needs one neighbour value from lane L-1 (alpha) or two from lane L+1 (beta): shuffles, no shared
// memory, no __syncthreads.  The recursion stays in the LOG domain (base 2; a linear-domain recursion with a shared scale loses
// states that are small against the frame's maximum but are the only ones able to finish, e.g. T = N), with the sorted form
//     log2(2^a + 2^b + 2^c) = m + lg2(1 + 2^(mid - m) + 2^(lo - m)),
// i.e. three MUFU operations per token state and two per blank state, all of a lane's SPL states independent of each other
// within a step.  "Impossible" is a large finite negative number, so no branch guards a (-inf) - (-inf).
// Warp 0 walks alpha forward, warp 1 walks beta backward (same CTA, no communication).  The workspaces receive the natural-log
// alpha / beta of the token states, as the shared-memory kernel writes them (fs_grad_kernel reads either).
// ------------------------------------------------------------------------------------------
constexpr int FSW_G = 4;               // frames of emissions in flight per lane (register ring, two groups)
constexpr float FSW_NEG = -1.0e30f;    // log2 of an impossible state
constexpr float FSW_LOG2E = 1.4426950408889634f;
constexpr float FSW_LN2 = 0.69314718055994530942f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float l2add2(float a, float b) {          // log2(2^a + 2^b)
  const float m = fmaxf(a, b);
  return m + lg2_approx(1.f + ex2_approx(fminf(a, b) - m));
}
__device__ __forceinline__ float l2add3(float a, float b, float c) { // log2(2^a + 2^b + 2^c)
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  return m + lg2_approx(1.f + ex2_approx(mid - m) + ex2_approx(lo - m));
}

// W warps per recursion (a single warp issues at ~0.45 instructions / cycle: the recursion is issue bound, not latency bound —
// ncu, round 2): the chain of a sample is spread over W warps; the one value (alpha) / two values (beta) that cross a warp
// boundary go through a double-buffered shared-memory mailbox and ONE named barrier per step and chain.
// SPLIT: the alpha and the beta chain of a sample are two CTAs (grid 2B, 32 W threads each) instead of two halves of one: the
// recursion is bound by instruction issue, and two chains on one SM share its four schedulers (T = 864 steps: 252 us together).
template <int SPL, int W, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 32 * W : 64 * W)
forward_sum_warp_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                        float blank_logit, const float* __restrict__ lse_ws, float* __restrict__ aw_all, float* __restrict__ bw_all,
                        float* __restrict__ nll_ws, float* __restrict__ loss, int B, int Tm, int Tx) {
  constexpr int TPL = SPL / 2;                 // token states per lane
  __shared__ float mbox[2][2][W][2];           // [chain][slot][warp][value]
  __shared__ float fin[2];
  const int b = SPLIT ? static_cast<int>(blockIdx.x) % B : static_cast<int>(blockIdx.x);
  const int lane = threadIdx.x & 31;
  const bool is_beta = SPLIT ? static_cast<int>(blockIdx.x) >= B : threadIdx.x >= 32 * W;
  const int wc = (threadIdx.x >> 5) - ((!SPLIT && is_beta) ? W : 0);      // warp index inside its chain
  const int gl = wc * 32 + lane;                               // lane index inside the chain
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  if (N <= 0 || T <= 0) {
    if (threadIdx.x == 0 && !is_beta) { loss[b] = 0.f; nll_ws[b] = INFINITY; }
    return;
  }
  const int S = 2 * N + 1;
  const float* lp = lpa + static_cast<long long>(b) * Tm * Tx;
  const float* lse = lse_ws + static_cast<long long>(b) * Tm;
  float* ws = (is_beta ? bw_all : aw_all) + static_cast<long long>(b) * Tm * Tx;
  const int k0 = TPL * gl;                     // first token of this lane; the blank state of register 2i belongs to slot k0 + i
  bool tok_ok[TPL], blk_ok[TPL];               // token k < N, blank slot k <= N
#pragma unroll
  for (int i = 0; i < TPL; ++i) { tok_ok[i] = k0 + i < N; blk_ok[i] = k0 + i <= N; }
  auto frame = [&](int i) { return is_beta ? T - 1 - i : i; };   // frame visited at step i
  auto chain_barrier = [&]() {
    if (is_beta) asm volatile("bar.sync 2, %0;" ::"n"(32 * W) : "memory");
    else asm volatile("bar.sync 1, %0;" ::"n"(32 * W) : "memory");
  };
  const int ch = is_beta ? 1 : 0;

  float a[SPL];                                // log2 of the state values
#pragma unroll
  for (int j = 0; j < SPL; ++j) a[j] = FSW_NEG;
  {
    const int f0 = frame(0);
    const float l0 = lse[f0];
    const float eb = (blank_logit - l0) * FSW_LOG2E;
    if (!is_beta) {          // alpha_0: blank state 0 and token state 1
      if (gl == 0) { a[0] = eb; a[1] = (lp[static_cast<long long>(f0) * Tx] - l0) * FSW_LOG2E; }
    } else {                 // beta_{T-1}: last blank state 2N and last token state 2N-1
#pragma unroll
      for (int i = 0; i < TPL; ++i) {
        if (k0 + i == N) a[2 * i] = eb;
        if (k0 + i == N - 1) a[2 * i + 1] = (lp[static_cast<long long>(f0) * Tx + N - 1] - l0) * FSW_LOG2E;
      }
    }
#pragma unroll
    for (int i = 0; i < TPL; ++i)
      if (tok_ok[i]) ws[static_cast<long long>(f0) * Tx + k0 + i] = a[2 * i + 1];
    if (!is_beta && lane == 31) mbox[0][0][wc][0] = a[SPL - 1];
    if (is_beta && lane == 0) { mbox[1][0][wc][0] = a[0]; mbox[1][0][wc][1] = a[1]; }
  }
  chain_barrier();

  // emissions of the next frames, fetched FSW_G steps ahead (two register groups)
  float e[FSW_G][TPL], en[FSW_G][TPL], lf[FSW_G], lfn[FSW_G];
  auto fetch = [&](int step, float (&dst)[TPL], float& l) {
    if (step < T) {
      const int f = frame(step);
      l = lse[f];
#pragma unroll
      for (int i = 0; i < TPL; ++i) dst[i] = tok_ok[i] ? lp[static_cast<long long>(f) * Tx + k0 + i] : -INFINITY;
    } else {
      l = 0.f;
#pragma unroll
      for (int i = 0; i < TPL; ++i) dst[i] = -INFINITY;
    }
  };
#pragma unroll
  for (int g = 0; g < FSW_G; ++g) fetch(1 + g, e[g], lf[g]);

  for (int base = 1; base < T; base += FSW_G) {
#pragma unroll
    for (int g = 0; g < FSW_G; ++g) fetch(base + FSW_G + g, en[g], lfn[g]);
#pragma unroll
    for (int g = 0; g < FSW_G; ++g) {
      const int step = base + g;
      if (step < T) {   // uniform across the CTA
        const int f = frame(step);
        const int rs = (step - 1) & 1, wsl = step & 1;                   // mailbox slots: read the previous step's, write this step's
        const float nl2 = -lf[g] * FSW_LOG2E;
        const float eb = fmaf(blank_logit, FSW_LOG2E, nl2);
        float em[TPL];                                                   // log2 emission of this lane's tokens (impossible: FSW_NEG)
#pragma unroll
        for (int i = 0; i < TPL; ++i) em[i] = fmaxf(fmaf(e[g][i], FSW_LOG2E, nl2), FSW_NEG);
        float n[SPL];
        if (!is_beta) {
          float pm = __shfl_up_sync(0xffffffffu, a[SPL - 1], 1);        // state SPL*gl - 1
          if (lane == 0) pm = wc > 0 ? mbox[0][rs][wc - 1][0] : FSW_NEG;
          n[0] = blk_ok[0] ? l2add2(a[0], pm) + eb : FSW_NEG;
          n[1] = l2add3(a[1], a[0], pm) + em[0];
#pragma unroll
          for (int i = 1; i < TPL; ++i) {
            n[2 * i] = blk_ok[i] ? l2add2(a[2 * i], a[2 * i - 1]) + eb : FSW_NEG;
            n[2 * i + 1] = l2add3(a[2 * i + 1], a[2 * i], a[2 * i - 1]) + em[i];
          }
        } else {
          float nx0 = __shfl_down_sync(0xffffffffu, a[0], 1);            // states SPL*(gl+1), SPL*(gl+1) + 1
          float nx1 = __shfl_down_sync(0xffffffffu, a[1], 1);
          if (lane == 31) {
            nx0 = wc + 1 < W ? mbox[1][rs][wc + 1][0] : FSW_NEG;
            nx1 = wc + 1 < W ? mbox[1][rs][wc + 1][1] : FSW_NEG;
          }
#pragma unroll
          for (int i = 0; i < TPL - 1; ++i) {
            n[2 * i] = blk_ok[i] ? l2add2(a[2 * i], a[2 * i + 1]) + eb : FSW_NEG;
            n[2 * i + 1] = l2add3(a[2 * i + 1], a[2 * i + 2], a[2 * i + 3]) + em[i];
          }
          n[SPL - 2] = blk_ok[TPL - 1] ? l2add2(a[SPL - 2], a[SPL - 1]) + eb : FSW_NEG;
          n[SPL - 1] = l2add3(a[SPL - 1], nx0, nx1) + em[TPL - 1];
        }
        // impossible states drift below the sentinel by one sentinel per step at most (-1e30 * T: far from the fp32 range)
#pragma unroll
        for (int j = 0; j < SPL; ++j) a[j] = n[j];
        if (!is_beta && lane == 31) mbox[0][wsl][wc][0] = a[SPL - 1];
        if (is_beta && lane == 0) { mbox[1][wsl][wc][0] = a[0]; mbox[1][wsl][wc][1] = a[1]; }
#pragma unroll
        for (int i = 0; i < TPL; ++i)
          if (tok_ok[i]) ws[static_cast<long long>(f) * Tx + k0 + i] = a[2 * i + 1];    // log2 units; impossible <= FSW_NEG
        chain_barrier();
      }
    }
#pragma unroll
    for (int g = 0; g < FSW_G; ++g) {
      lf[g] = lfn[g];
#pragma unroll
      for (int i = 0; i < TPL; ++i) e[g][i] = en[g][i];
    }
  }
  if (!is_beta) {   // log-likelihood = log(alpha_{T-1}(S-1) + alpha_{T-1}(S-2))
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s_ = SPL * gl + j;
      if (s_ == S - 1) fin[0] = a[j];
      if (s_ == S - 2) fin[1] = a[j];
    }
    chain_barrier();
    if (gl == 0) {
      const float ll = l2add2(fin[0], S >= 2 ? fin[1] : FSW_NEG) * FSW_LN2;
      const bool finite = ll > 0.5f * FSW_NEG && ll < INFINITY;
      loss[b] = finite ? -ll / static_cast<float>(N) : 0.f;   // reduction='mean' (target length), zero_infinity
      nll_ws[b] = finite ? -ll : INFINITY;
    }
  }
}

// gradient for the log2-domain workspaces of the warp kernel: posterior = exp((aw + bw) ln 2 - e + nll), 0 for impossible states
__global__ void fs_grad_log2_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                                    const float* __restrict__ lse_ws, const float* __restrict__ aw, const float* __restrict__ bw,
                                    const float* __restrict__ nll_ws, float* __restrict__ grad, int B, int Tm, int Tx) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * Tm * Tx) return;
  const int n = static_cast<int>(idx % Tx);
  const long long row = idx / Tx;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row % Tm);
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  float g = 0.f;
  const float nll = nll_ws[b];
  if (t < T && n < N && nll < INFINITY) {
    const float e = lpa[idx] - lse_ws[row];
    const float av = aw[idx], bv = bw[idx];
    const float post = (av > 0.5f * FSW_NEG && bv > 0.5f * FSW_NEG) ? expf((av + bv) * FSW_LN2 - e + nll) : 0.f;
    g = (expf(e) - post) / (static_cast<float>(N) * static_cast<float>(B));
  }
  grad[idx] = g;
}

// the same, four consecutive tokens per thread (Tx % 4 == 0): 16-byte loads / stores — the kernel is on the step's critical path
// right behind the recursion (85 MB of traffic: 36 us with scalar accesses)
__global__ void __launch_bounds__(256)
fs_grad_log2_vec4_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                         const float* __restrict__ lse_ws, const float* __restrict__ aw, const float* __restrict__ bw,
                         const float* __restrict__ nll_ws, float* __restrict__ grad, int B, int Tm, int Tx) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int Tx4 = Tx >> 2;
  if (i4 >= static_cast<long long>(B) * Tm * Tx4) return;
  const int n0 = static_cast<int>(i4 % Tx4) * 4;
  const long long row = i4 / Tx4;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row % Tm);
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  const float nll = nll_ws[b];
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < T && n0 < N && nll < INFINITY) {
    const float l = lse_ws[row];
    const float4 lp = *reinterpret_cast<const float4*>(lpa + i4 * 4);
    const float4 av = *reinterpret_cast<const float4*>(aw + i4 * 4);
    const float4 bv = *reinterpret_cast<const float4*>(bw + i4 * 4);
    const float inv = 1.f / (static_cast<float>(N) * static_cast<float>(B));
    auto one = [&](float lpv, float a, float bb, int n) -> float {
      if (n >= N) return 0.f;
      const float e = lpv - l;
      const float post = (a > 0.5f * FSW_NEG && bb > 0.5f * FSW_NEG) ? expf((a + bb) * FSW_LN2 - e + nll) : 0.f;
      return (expf(e) - post) * inv;
    };
    g.x = one(lp.x, av.x, bv.x, n0);
    g.y = one(lp.y, av.y, bv.y, n0 + 1);
    g.z = one(lp.z, av.z, bv.z, n0 + 2);
    g.w = one(lp.w, av.w, bv.w, n0 + 3);
  }
  *reinterpret_cast<float4*>(grad + i4 * 4) = g;
}

// gradient, fully parallel over (b, t, n): softmax minus posterior occupancy, with the per-sample 1/N and the batch 1/B
__global__ void fs_grad_kernel(const float* __restrict__ lpa, const long long* __restrict__ x_len, const long long* __restrict__ m_len,
                               const float* __restrict__ lse_ws, const float* __restrict__ aw, const float* __restrict__ bw,
                               const float* __restrict__ nll_ws, float* __restrict__ grad, int B, int Tm, int Tx) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * Tm * Tx) return;
  const int n = static_cast<int>(idx % Tx);
  const long long row = idx / Tx;
  const int b = static_cast<int>(row / Tm), t = static_cast<int>(row % Tm);
  const int N = static_cast<int>(x_len[b]), T = static_cast<int>(m_len[b]);
  float g = 0.f;
  const float nll = nll_ws[b];
  if (t < T && n < N && nll < INFINITY) {
    const float e = lpa[idx] - lse_ws[row];
    g = (expf(e) - expf(aw[idx] + bw[idx] - e + nll)) / (static_cast<float>(N) * static_cast<float>(B));
  }
  grad[idx] = g;
}

}  // namespace
}  // namespace osb

static int g_fs_force_legacy = 0;
/* developer hook (not in the public header): launch shape of the warp recursion — 1 (default): one CTA per chain; 0: the alpha
 * and the beta chain of a sample in one CTA.  (8 warps x 2 states per chain was measured too: 275 us split, 382 us fused.) */
static int g_fs_variant = 1;
extern "C" void osb_debug_forward_sum_variant(int v) { g_fs_variant = v; }
/* developer hook (not in the public header): 1 = always use the shared-memory recursion (parity tests compare the two) */
extern "C" void osb_debug_forward_sum_legacy(int on) { g_fs_force_legacy = on; }

extern "C" int osb_forward_sum(const float* log_p_attn, const int64_t* x_len, const int64_t* m_len, float blank_logit, float* alpha_ws,
                               float* loss, float* grad, int32_t B, int32_t Tm, int32_t Tx, void* stream) {
  OSB_REQUIRE(log_p_attn && x_len && m_len && alpha_ws && loss && grad, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0 && Tx <= 1023, OSB_ERR_SHAPE);
  const int nthr = ((Tx + 1 + 31) / 32) * 32;
  const int halves = 2 * nthr <= 1024 ? 1 : 0;   // alpha and beta chains side by side when both fit one CTA
  const size_t smem = sizeof(float) * (4 * static_cast<size_t>(2 * Tx + 1) + Tm);
  OSB_REQUIRE(smem <= 200 * 1024, OSB_ERR_SHAPE);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(osb::forward_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = 200 * 1024;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long plane = static_cast<long long>(B) * Tm * Tx;
  float* aw = alpha_ws;
  float* bw = alpha_ws + plane;
  float* lse_ws = alpha_ws + 2 * plane;
  float* nll_ws = lse_ws + static_cast<long long>(B) * Tm;
  const long long* xl = reinterpret_cast<const long long*>(x_len);
  const long long* ml = reinterpret_cast<const long long*>(m_len);
  const long long rows = static_cast<long long>(B) * Tm;
  osb::fs_lse_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, B, Tm, Tx);
  // warp-synchronous register recursion when the 2 Tx + 1 states fit 32 lanes x SPL registers; else the shared-memory kernel
  const int S_max = 2 * Tx + 1;
  if (S_max <= 1024 && !g_fs_force_legacy) {
    // default: one CTA per CHAIN (alpha / beta of a sample on different SMs): 224 -> 165 us at B=32, Tm=864, Tx=192, bit-identical
    // results (tools/probe_fs.py); g_fs_variant 0 keeps both chains of a sample in one CTA (A/B)
    const bool split = g_fs_variant != 0;
    if (S_max <= 256) {
      if (split) osb::forward_sum_warp_kernel<4, 2, true><<<2 * B, 64, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
      else osb::forward_sum_warp_kernel<4, 2, false><<<B, 128, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
    } else if (S_max <= 512) {
      if (split) osb::forward_sum_warp_kernel<4, 4, true><<<2 * B, 128, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
      else osb::forward_sum_warp_kernel<4, 4, false><<<B, 256, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
    } else {
      if (split) osb::forward_sum_warp_kernel<8, 4, true><<<2 * B, 128, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
      else osb::forward_sum_warp_kernel<8, 4, false><<<B, 256, 0, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx);
    }
    const bool vec4 = Tx % 4 == 0 && (reinterpret_cast<uintptr_t>(log_p_attn) & 15) == 0 && (reinterpret_cast<uintptr_t>(aw) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(bw) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
    if (vec4) osb::fs_grad_log2_vec4_kernel<<<static_cast<unsigned>((plane / 4 + 255) / 256), 256, 0, s>>>(log_p_attn, xl, ml, lse_ws, aw, bw, nll_ws, grad, B, Tm, Tx);
    else osb::fs_grad_log2_kernel<<<static_cast<unsigned>((plane + 255) / 256), 256, 0, s>>>(log_p_attn, xl, ml, lse_ws, aw, bw, nll_ws, grad, B, Tm, Tx);
    osb::count_launch(3);
    return osb::launch_status();
  }
  osb::forward_sum_kernel<<<B, halves ? 2 * nthr : nthr, smem, s>>>(log_p_attn, xl, ml, blank_logit, lse_ws, aw, bw, nll_ws, loss, B, Tm, Tx,
                                                                   halves);
  osb::fs_grad_kernel<<<static_cast<unsigned>((plane + 255) / 256), 256, 0, s>>>(log_p_attn, xl, ml, lse_ws, aw, bw, nll_ws, grad, B, Tm, Tx);
  osb::count_launch(3);
  return osb::launch_status();
}
