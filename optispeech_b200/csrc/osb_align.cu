// osb_align.cu — alignment-learning kernels that the reference runs on the host CPU (numba):
// monotonic alignment search (Viterbi over the attention log-probabilities) and duration-span
// averaging, moved onto the device so a training step has no per-sample host round trips.
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
  for (int j = 1; j < T; ++j) {
    const int cur = j & 1, prev = cur ^ 1;
    const double left = (i > 0) ? qbuf[prev * nthr + i - 1] : NEG;
    const bool take_left = (i > 0) && (left >= q);
    const unsigned bits = __ballot_sync(0xffffffffu, take_left);
    if ((i & 31) == 0) flags[static_cast<size_t>(j) * words + (i >> 5)] = bits;
    if (i == 0) {
      row0 = __fadd_rn(row0, lpb[static_cast<long long>(j) * Tx]);
      q = static_cast<double>(row0);
    } else if (i < N && i <= j) {
      q = fmax(left, q) + static_cast<double>(lpb[static_cast<long long>(j) * Tx + i]);
    }
    qbuf[cur * nthr + i] = q;
    __syncthreads();
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
