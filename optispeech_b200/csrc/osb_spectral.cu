// osb_spectral.cu — fused STFT-domain reconstruction losses (multi-resolution STFT and log-mel L1) with
// their gradient.  One CTA owns one frame of one (prediction, target) signal pair:
//   reflect-padded framing + Hann window -> ONE complex FFT in shared memory for both real signals
//   (z = x_hat + i*y, spectra separated by Hermitian symmetry) -> magnitudes -> loss partial sums;
// the backward kernel recomputes the spectra, forms d(loss)/d(spectrum of x_hat), applies the adjoint
// (inverse, un-normalised) FFT, the window and the reflect-pad adjoint, and overlap-adds into d(x_hat).
// Nothing but the two waveforms is read from HBM and nothing but partial sums / d(x_hat) is written:
// the unfolded frames, complex spectra and magnitude tensors of the reference pipeline never exist.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int FFT_THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__device__ __forceinline__ int reflect_index(int p, int L) {
  if (p < 0) p = -p;
  if (p >= L) p = 2 * (L - 1) - p;
  return p;
}

// In-place decimation-in-time FFT over shared memory; input must already be in bit-reversed order.
// tw[k] = exp(-2*pi*i*k/N), k < N/2; INVERSE uses the conjugate twiddles (no 1/N normalisation).
// Two radix-2 stages (lengths len and 2*len) are fused per pass: the four elements i0, i0+len/2, i0+len, i0+len+len/2 form a
// closed set under both, so a thread carries them in registers — half the block-wide barriers and half the shared-memory
// round trips of the stage-per-pass loop (5 passes instead of 10 for N = 1024).  An odd log2(N) starts with one radix-2 stage.
template <int N, bool INVERSE>
__device__ __forceinline__ void fft_inplace(float2* s, const float2* tw) {
  constexpr int LOG2N = (N == 512) ? 9 : (N == 1024) ? 10 : 11;
  int len = 2;
  if constexpr (LOG2N & 1) {
    for (int j = threadIdx.x; j < N / 2; j += FFT_THREADS) {   // len = 2: twiddle 1
      const float2 a = s[2 * j], b = s[2 * j + 1];
      s[2 * j] = make_float2(a.x + b.x, a.y + b.y);
      s[2 * j + 1] = make_float2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
    len = 4;
  }
#pragma unroll 1
  for (; len <= N / 2; len <<= 2) {
    const int half = len >> 1;
    const int t1 = N / len, t2 = N / (2 * len);
    for (int j = threadIdx.x; j < N / 4; j += FFT_THREADS) {
      const int grp = j / half, pos = j - grp * half;
      const int i0 = grp * (2 * len) + pos;
      float2 w1 = tw[pos * t1], w2 = tw[pos * t2], w3 = tw[(pos + half) * t2];
      if (INVERSE) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
      const float2 a = s[i0], b = s[i0 + half], c = s[i0 + len], d = s[i0 + len + half];
      const float2 tb = cmul(w1, b), td = cmul(w1, d);
      const float2 a1 = make_float2(a.x + tb.x, a.y + tb.y), b1 = make_float2(a.x - tb.x, a.y - tb.y);
      const float2 c1 = make_float2(c.x + td.x, c.y + td.y), d1 = make_float2(c.x - td.x, c.y - td.y);
      const float2 u = cmul(w2, c1), v = cmul(w3, d1);
      s[i0] = make_float2(a1.x + u.x, a1.y + u.y);
      s[i0 + len] = make_float2(a1.x - u.x, a1.y - u.y);
      s[i0 + half] = make_float2(b1.x + v.x, b1.y + v.y);
      s[i0 + len + half] = make_float2(b1.x - v.x, b1.y - v.y);
    }
    __syncthreads();
  }
}

template <int N>
__device__ __forceinline__ int bitrev(int v) {
  constexpr int BITS = (N == 512) ? 9 : (N == 1024) ? 10 : 11;
  return static_cast<int>(__brev(static_cast<unsigned>(v)) >> (32 - BITS));
}

struct SpecParams {
  const float* x_hat;     // (B, L)
  const float* y;         // (B, L)
  const float* window;    // (win)
  int B, L, hop, win, frames;
  // mel mode (fb != nullptr): fb (N/2+1, n_mels) row-major, per-mel bin range [klo, khi], per-bin mel range [jlo, jhi]
  const float* fb;
  const int* klo; const int* khi; const int* jlo; const int* jhi;
  int n_mels;
  float clamp_min;        // 1e-7
  double* stats;          // fwd out / bwd in: [0]=sum (Ym-Xm)^2, [1]=sum Ym^2, [2]=sum |log Ym - log Xm| (or mel L1 in [2])
  // backward
  const float* coef;      // device (3): upstream dL/dSC, dL/dMAG (or dL/dMEL in [1]) ... see kernel
  float* dx_hat;          // (B, L) accumulated with atomics
};

// shared memory: float2 s[N]; float2 tw[N/2]; float mag[2][N/2+1] (mel mode) ; float red[...]
template <int N, bool MEL, bool BACKWARD>
__global__ void __launch_bounds__(FFT_THREADS) stft_loss_kernel(const SpecParams p) {
  extern __shared__ float2 sp_smem[];
  float2* s = sp_smem;
  float2* tw = s + N;
  float* magx = reinterpret_cast<float*>(tw + N / 2);   // [N/2+1]  (mel mode / backward scratch)
  float* magy = magx + (N / 2 + 1);                     // [N/2+1]
  float* dmel = magy + (N / 2 + 1);                     // [n_mels] (mel backward)
  __shared__ float red[3][FFT_THREADS / 32];

  const int f = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;
  const float* xh = p.x_hat + static_cast<long long>(b) * p.L;
  const float* yy = p.y + static_cast<long long>(b) * p.L;
  const int off = (N - p.win) / 2;  // torch.stft centres the window inside n_fft

  for (int k = tid; k < N / 2; k += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(N), &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  for (int n = tid; n < N; n += FFT_THREADS) {
    float2 v = make_float2(0.f, 0.f);
    const int m = n - off;
    if (m >= 0 && m < p.win) {
      const float w = p.window[m];
      const int j = reflect_index(f * p.hop + n - N / 2, p.L);
      v = make_float2(w * xh[j], w * yy[j]);
    }
    s[bitrev<N>(n)] = v;
  }
  __syncthreads();
  fft_inplace<N, false>(s, tw);

  // separate the two real spectra; bins 0..N/2
  constexpr int NB = N / 2 + 1;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if constexpr (!BACKWARD) {
    for (int k = tid; k < NB; k += FFT_THREADS) {
      const float2 zk = s[k], zn = s[(N - k) & (N - 1)];
      const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
      const float yr = 0.5f * (zk.y + zn.y), yi = -0.5f * (zk.x - zn.x);
      const float px = xr * xr + xi * xi, py = yr * yr + yi * yi;
      if constexpr (MEL) {
        magx[k] = sqrtf(px);
        magy[k] = sqrtf(py);
      } else {
        const float xm = sqrtf(fmaxf(px, p.clamp_min)), ym = sqrtf(fmaxf(py, p.clamp_min));
        const float d = ym - xm;
        a0 = fmaf(d, d, a0);
        a1 = fmaf(ym, ym, a1);
        a2 += fabsf(logf(ym) - logf(xm));
      }
    }
    if constexpr (MEL) {
      __syncthreads();
      for (int j = tid; j < p.n_mels; j += FFT_THREADS) {
        float mx = 0.f, my = 0.f;
        for (int k = p.klo[j]; k <= p.khi[j]; ++k) {
          const float w = p.fb[k * p.n_mels + j];
          mx = fmaf(w, magx[k], mx);
          my = fmaf(w, magy[k], my);
        }
        a2 += fabsf(logf(fmaxf(my, p.clamp_min)) - logf(fmaxf(mx, p.clamp_min)));
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if ((tid & 31) == 0) { red[0][tid >> 5] = a0; red[1][tid >> 5] = a1; red[2][tid >> 5] = a2; }
    __syncthreads();
    if (tid == 0) {
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      for (int i = 0; i < FFT_THREADS / 32; ++i) { r0 += red[0][i]; r1 += red[1][i]; r2 += red[2][i]; }
      if (!MEL) { atomicAdd(p.stats + 0, static_cast<double>(r0)); atomicAdd(p.stats + 1, static_cast<double>(r1)); }
      atomicAdd(p.stats + 2, static_cast<double>(r2));
    }
    return;
  } else {
    // ---------------- backward: G_k = d(loss)/d(re_k) + i d(loss)/d(im_k) of the x_hat spectrum ----------------
    // coef[0] = dL/dSC_r * 1/(sqrt(S1) sqrt(S2))   (spectral convergence: d sqrt(S1)/dXm = -(Ym-Xm)/sqrt(S1))
    // coef[1] = dL/dMAG_r * 1/count                (log-magnitude L1)       | mel: coef[1] = dL/dMEL * 1/count
    const float c_sc = p.coef[0], c_mag = p.coef[1];
    float2 g_local[(NB + FFT_THREADS - 1) / FFT_THREADS];
    if constexpr (MEL) {
      for (int k = tid; k < NB; k += FFT_THREADS) {
        const float2 zk = s[k], zn = s[(N - k) & (N - 1)];
        const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
        const float yr = 0.5f * (zk.y + zn.y), yi = -0.5f * (zk.x - zn.x);
        magx[k] = sqrtf(xr * xr + xi * xi);
        magy[k] = sqrtf(yr * yr + yi * yi);
      }
      __syncthreads();
      for (int j = tid; j < p.n_mels; j += FFT_THREADS) {
        float mx = 0.f, my = 0.f;
        for (int k = p.klo[j]; k <= p.khi[j]; ++k) {
          const float w = p.fb[k * p.n_mels + j];
          mx = fmaf(w, magx[k], mx);
          my = fmaf(w, magy[k], my);
        }
        const float lx = logf(fmaxf(mx, p.clamp_min)), ly = logf(fmaxf(my, p.clamp_min));
        const float sgn = (ly > lx) ? 1.f : ((ly < lx) ? -1.f : 0.f);
        dmel[j] = (mx > p.clamp_min) ? -sgn * c_mag / mx : 0.f;   // d|ly - lx|/d mx
      }
      __syncthreads();
    }
    int slot = 0;
    for (int k = tid; k < NB; k += FFT_THREADS, ++slot) {
      const float2 zk = s[k], zn = s[(N - k) & (N - 1)];
      const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
      float dmag;   // d(loss)/d(magnitude of x_hat at bin k), already divided by the magnitude below
      float2 g = make_float2(0.f, 0.f);
      if constexpr (MEL) {
        float acc = 0.f;
        for (int j = p.jlo[k]; j <= p.jhi[k]; ++j) acc = fmaf(p.fb[k * p.n_mels + j], dmel[j], acc);
        const float m = magx[k];
        dmag = m > 0.f ? acc / m : 0.f;
        g = make_float2(dmag * xr, dmag * xi);
      } else {
        const float yr = 0.5f * (zk.y + zn.y), yi = -0.5f * (zk.x - zn.x);
        const float px = xr * xr + xi * xi, py = yr * yr + yi * yi;
        if (px > p.clamp_min) {
          const float xm = sqrtf(px), ym = sqrtf(fmaxf(py, p.clamp_min));
          const float lx = logf(xm), ly = logf(ym);
          const float sgn = (ly > lx) ? 1.f : ((ly < lx) ? -1.f : 0.f);
          const float dxm = -c_sc * (ym - xm) - c_mag * sgn / xm;
          dmag = dxm / xm;
          g = make_float2(dmag * xr, dmag * xi);
        }
      }
      g_local[slot] = g;
    }
    __syncthreads();  // everyone has read the forward spectrum
    for (int n = tid; n < N; n += FFT_THREADS) s[n] = make_float2(0.f, 0.f);
    __syncthreads();
    slot = 0;
    for (int k = tid; k < NB; k += FFT_THREADS, ++slot) s[bitrev<N>(k)] = g_local[slot];
    __syncthreads();
    fft_inplace<N, true>(s, tw);   // du_n = Re sum_k G_k e^{+2 pi i k n / N}
    float* dx = p.dx_hat + static_cast<long long>(b) * p.L;
    for (int n = tid; n < N; n += FFT_THREADS) {
      const int m = n - off;
      if (m >= 0 && m < p.win) {
        const int j = reflect_index(f * p.hop + n - N / 2, p.L);
        atomicAdd(dx + j, s[n].x * p.window[m]);
      }
    }
  }
}

template <int N>
size_t spec_smem_bytes(int n_mels) {
  return sizeof(float2) * (N + N / 2) + sizeof(float) * (2 * (N / 2 + 1) + (n_mels > 0 ? n_mels : 1));
}

template <int N, bool MEL, bool BACKWARD>
int launch_spec(const SpecParams& p, cudaStream_t stream) {
  const size_t smem = spec_smem_bytes<N>(p.n_mels);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(stft_loss_kernel<N, MEL, BACKWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  stft_loss_kernel<N, MEL, BACKWARD><<<dim3(p.frames, p.B), FFT_THREADS, smem, stream>>>(p);
  count_launch();
  return launch_status();
}

template <bool MEL, bool BACKWARD>
int dispatch_spec(int n_fft, const SpecParams& p, cudaStream_t stream) {
  switch (n_fft) {
    case 512: return launch_spec<512, MEL, BACKWARD>(p, stream);
    case 1024: return launch_spec<1024, MEL, BACKWARD>(p, stream);
    case 2048: return launch_spec<2048, MEL, BACKWARD>(p, stream);
    default: return OSB_ERR_SHAPE;
  }
}

// ------------------------------------------------------------------------------------------
// Feature extraction (reference dataset/feature_extractors/__init__.py:110-200, `CommonFeatureExtractor.get_mel` and
// `FeatureExtractor.get_energy`): reflect padding by (n_fft - hop)/2 on both sides, framing WITHOUT centring, Hann window,
// magnitude sqrt(re^2 + im^2 + 1e-9), then   energy[f] = || mag[f, :] ||_2   and   mel[j, f] = log(max(sum_k fb[j,k] mag[f,k], 1e-5)).
// One CTA transforms two consecutive frames of one utterance with ONE complex FFT (z = frame_a + i*frame_b); every utterance
// is padded by its OWN length (a batch of ragged utterances gives what the per-utterance reference loop gives).
// ------------------------------------------------------------------------------------------
struct FeatParams {
  const float* wav;           // (B, Lmax)
  const long long* lengths;   // (B) samples
  const float* window;        // (win)
  const float* fb;            // (n_mels, N/2+1) row-major (librosa layout)
  const int* klo; const int* khi;   // (n_mels) non-zero bin range of every filter
  float* mel;                 // (B, n_mels, Fmax), zero for frames >= length/hop
  float* energy;              // (B, Fmax)
  int B, Lmax, hop, win, Fmax, n_mels;
  float mag_eps, clip_val;
};

template <int N>
__global__ void __launch_bounds__(FFT_THREADS) mel_energy_kernel(const FeatParams p) {
  extern __shared__ float2 sp_smem[];
  float2* s = sp_smem;
  float2* tw = s + N;
  constexpr int NB = N / 2 + 1;
  float* maga = reinterpret_cast<float*>(tw + N / 2);   // [NB]
  float* magb = maga + NB;                              // [NB]
  __shared__ float red[2][FFT_THREADS / 32];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int fa = 2 * blockIdx.x, fb_ = fa + 1;
  const int L = static_cast<int>(p.lengths[b]);
  const int frames = L / p.hop;
  const int pad = (N - p.hop) / 2;
  const int off = (N - p.win) / 2;
  const float* w = p.wav + static_cast<long long>(b) * p.Lmax;
  const bool va = fa < frames, vb = fb_ < frames;
  if (!va) {   // both frames past the utterance: zeros
    for (int j = tid; j < p.n_mels; j += FFT_THREADS) {
      if (fa < p.Fmax) p.mel[(static_cast<long long>(b) * p.n_mels + j) * p.Fmax + fa] = 0.f;
      if (fb_ < p.Fmax) p.mel[(static_cast<long long>(b) * p.n_mels + j) * p.Fmax + fb_] = 0.f;
    }
    if (tid == 0) {
      if (fa < p.Fmax) p.energy[static_cast<long long>(b) * p.Fmax + fa] = 0.f;
      if (fb_ < p.Fmax) p.energy[static_cast<long long>(b) * p.Fmax + fb_] = 0.f;
    }
    return;
  }
  for (int k = tid; k < N / 2; k += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(N), &sn, &cs);
    tw[k] = make_float2(cs, sn);
  }
  for (int n = tid; n < N; n += FFT_THREADS) {
    float2 v = make_float2(0.f, 0.f);
    const int m = n - off;
    if (m >= 0 && m < p.win) {
      const float wn = p.window[m];
      v.x = wn * w[reflect_index(fa * p.hop + n - pad, L)];
      if (vb) v.y = wn * w[reflect_index(fb_ * p.hop + n - pad, L)];
    }
    s[bitrev<N>(n)] = v;
  }
  __syncthreads();
  fft_inplace<N, false>(s, tw);
  float ea = 0.f, eb = 0.f;
  for (int k = tid; k < NB; k += FFT_THREADS) {
    const float2 zk = s[k], zn = s[(N - k) & (N - 1)];
    const float xr = 0.5f * (zk.x + zn.x), xi = 0.5f * (zk.y - zn.y);
    const float yr = 0.5f * (zk.y + zn.y), yi = -0.5f * (zk.x - zn.x);
    const float pa = xr * xr + xi * xi + p.mag_eps, pb = yr * yr + yi * yi + p.mag_eps;
    maga[k] = sqrtf(pa);
    magb[k] = sqrtf(pb);
    ea += pa;
    eb += pb;
  }
  ea = warp_sum(ea); eb = warp_sum(eb);
  if ((tid & 31) == 0) { red[0][tid >> 5] = ea; red[1][tid >> 5] = eb; }
  __syncthreads();
  if (tid == 0) {
    float ra = 0.f, rb = 0.f;
    for (int i = 0; i < FFT_THREADS / 32; ++i) { ra += red[0][i]; rb += red[1][i]; }
    p.energy[static_cast<long long>(b) * p.Fmax + fa] = sqrtf(ra);
    if (fb_ < p.Fmax) p.energy[static_cast<long long>(b) * p.Fmax + fb_] = vb ? sqrtf(rb) : 0.f;
  }
  for (int j = tid; j < p.n_mels; j += FFT_THREADS) {
    float ma = 0.f, mb = 0.f;
    const float* frow = p.fb + static_cast<long long>(j) * NB;
    for (int k = p.klo[j]; k <= p.khi[j]; ++k) {
      const float wk = frow[k];
      ma = fmaf(wk, maga[k], ma);
      mb = fmaf(wk, magb[k], mb);
    }
    float* mrow = p.mel + (static_cast<long long>(b) * p.n_mels + j) * p.Fmax;
    mrow[fa] = logf(fmaxf(ma, p.clip_val));
    if (fb_ < p.Fmax) mrow[fb_] = vb ? logf(fmaxf(mb, p.clip_val)) : 0.f;
  }
}

template <int N>
int launch_feat(const FeatParams& p, cudaStream_t stream) {
  const size_t smem = sizeof(float2) * (N + N / 2) + sizeof(float) * 2 * (N / 2 + 1);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mel_energy_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  mel_energy_kernel<N><<<dim3((p.Fmax + 1) / 2, p.B), FFT_THREADS, smem, stream>>>(p);
  count_launch();
  return launch_status();
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_stft_loss(const float* x_hat, const float* y, const float* window, int32_t B, int32_t L, int32_t n_fft, int32_t hop,
                             int32_t win, float clamp_min, double* stats, const float* coef, float* dx_hat, void* stream) {
  OSB_REQUIRE(x_hat && y && window && stats, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && L > n_fft / 2 && hop > 0 && win > 0 && win <= n_fft, OSB_ERR_SHAPE);
  SpecParams p{};
  p.x_hat = x_hat; p.y = y; p.window = window; p.B = B; p.L = L; p.hop = hop; p.win = win; p.frames = 1 + L / hop;
  p.fb = nullptr; p.n_mels = 0; p.clamp_min = clamp_min; p.stats = stats; p.coef = coef; p.dx_hat = dx_hat;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dx_hat != nullptr) {
    OSB_REQUIRE(coef != nullptr, OSB_ERR_ARG);
    return dispatch_spec<false, true>(n_fft, p, s);
  }
  return dispatch_spec<false, false>(n_fft, p, s);
}

extern "C" int osb_mel_loss(const float* x_hat, const float* y, const float* window, const float* fb, const int32_t* klo,
                            const int32_t* khi, const int32_t* jlo, const int32_t* jhi, int32_t n_mels, int32_t B, int32_t L,
                            int32_t n_fft, int32_t hop, int32_t win, float clamp_min, double* stats, const float* coef, float* dx_hat,
                            void* stream) {
  OSB_REQUIRE(x_hat && y && window && fb && klo && khi && jlo && jhi && stats, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && L > n_fft / 2 && hop > 0 && win > 0 && win <= n_fft && n_mels > 0 && n_mels <= 1024, OSB_ERR_SHAPE);
  SpecParams p{};
  p.x_hat = x_hat; p.y = y; p.window = window; p.B = B; p.L = L; p.hop = hop; p.win = win; p.frames = 1 + L / hop;
  p.fb = fb; p.klo = klo; p.khi = khi; p.jlo = jlo; p.jhi = jhi; p.n_mels = n_mels; p.clamp_min = clamp_min;
  p.stats = stats; p.coef = coef; p.dx_hat = dx_hat;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dx_hat != nullptr) {
    OSB_REQUIRE(coef != nullptr, OSB_ERR_ARG);
    return dispatch_spec<true, true>(n_fft, p, s);
  }
  return dispatch_spec<true, false>(n_fft, p, s);
}

extern "C" int osb_mel_energy(const float* wav, const int64_t* lengths, const float* window, const float* fb, const int32_t* klo,
                              const int32_t* khi, float* mel, float* energy, int32_t B, int32_t Lmax, int32_t Fmax, int32_t n_mels,
                              int32_t n_fft, int32_t hop, int32_t win, float mag_eps, float clip_val, void* stream) {
  OSB_REQUIRE(wav && lengths && window && fb && klo && khi && mel && energy, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Lmax > n_fft / 2 && hop > 0 && hop <= n_fft && win > 0 && win <= n_fft && Fmax > 0 && n_mels > 0, OSB_ERR_SHAPE);
  FeatParams p{};
  p.wav = wav; p.lengths = reinterpret_cast<const long long*>(lengths); p.window = window; p.fb = fb; p.klo = klo; p.khi = khi;
  p.mel = mel; p.energy = energy; p.B = B; p.Lmax = Lmax; p.hop = hop; p.win = win; p.Fmax = Fmax; p.n_mels = n_mels;
  p.mag_eps = mag_eps; p.clip_val = clip_val;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (n_fft) {
    case 512: return launch_feat<512>(p, s);
    case 1024: return launch_feat<1024>(p, s);
    case 2048: return launch_feat<2048>(p, s);
    default: return OSB_ERR_SHAPE;
  }
}
