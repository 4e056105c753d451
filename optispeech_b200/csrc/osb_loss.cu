// osb_loss.cu — the FastSpeech2 variance losses of the training step and their gradients in one launch.
//
// Replaces FastSpeech2Loss.forward (optispeech/model/generator/loss.py:83-140) AS THE REFERENCE EVALUATES IT: its masks carry a
// stray singleton axis, so `masked_select` broadcasts — the duration term takes element (b,i) len_b times for EVERY i < Tx
// (padded positions included, target log(0 + 1e-8)), the pitch / energy terms take element (b,i) once per sample whose length
// exceeds i — followed by 'mean' reductions:
//     d_loss = sum_b len_b * sum_i (d_hat[b,i] - log(ds[b,i] + 1e-8))^2 / (sum_b len_b * Tx)
//     w[i]   = #{b : len_b > i};   p_loss = sum_{b,i} smooth_l1(p_hat[b,i] - p_tgt[b,i]) * w[i] / (B * sum_i w[i])   (same for e)
// The ~25 small reductions / elementwise kernels of the PyTorch formulation (and as many in its backward) sit on the critical
// path between the joins of the prediction branches and the start of the backward pass; here they are one CTA.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int FL_THREADS = 1024;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < (FL_THREADS / 32) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  __syncthreads();
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

__global__ void __launch_bounds__(FL_THREADS)
fs2_losses_kernel(const float* __restrict__ d_hat, const float* __restrict__ p_hat, const float* __restrict__ e_hat,
                  const float* __restrict__ ds, const float* __restrict__ p_tgt, const float* __restrict__ e_tgt,
                  const long long* __restrict__ x_len, float* __restrict__ losses, float* __restrict__ g_d, float* __restrict__ g_p,
                  float* __restrict__ g_e, int B, int Tx) {
  extern __shared__ float fl_smem[];
  float* w = fl_smem;            // [Tx] samples longer than i
  float* lens = w + Tx;          // [B]
  float* red = lens + B;         // [32]
  for (int b = threadIdx.x; b < B; b += FL_THREADS) lens[b] = static_cast<float>(x_len[b]);
  __syncthreads();
  float wsum_part = 0.f;
  for (int i = threadIdx.x; i < Tx; i += FL_THREADS) {
    int c = 0;
    for (int b = 0; b < B; ++b) c += (lens[b] > static_cast<float>(i)) ? 1 : 0;
    w[i] = static_cast<float>(c);
    wsum_part += static_cast<float>(c);
  }
  float len_part = 0.f;
  for (int b = threadIdx.x; b < B; b += FL_THREADS) len_part += lens[b];
  const float wsum = block_sum(wsum_part, red);
  const float lsum = block_sum(len_part, red);
  const float d_den = lsum * static_cast<float>(Tx);
  const float pe_den = wsum * static_cast<float>(B);
  const float inv_d = d_den > 0.f ? 1.f / d_den : 0.f;
  const float inv_pe = pe_den > 0.f ? 1.f / pe_den : 0.f;
  float sd = 0.f, sp = 0.f, se = 0.f;
  const int n = B * Tx;
  for (int idx = threadIdx.x; idx < n; idx += FL_THREADS) {
    const int b = idx / Tx, i = idx - b * Tx;
    const float dd = d_hat[idx] - logf(ds[idx] + 1e-8f);
    sd = fmaf(dd * dd, lens[b], sd);
    g_d[idx] = 2.f * dd * lens[b] * inv_d;
    const float wi = w[i];
    const float dp = p_hat[idx] - p_tgt[idx];
    const float ap = fabsf(dp);
    sp = fmaf(ap < 1.f ? 0.5f * dp * dp : ap - 0.5f, wi, sp);          // SmoothL1, beta = 1 (loss.py:77-78)
    g_p[idx] = fminf(fmaxf(dp, -1.f), 1.f) * wi * inv_pe;
    const float de = e_hat[idx] - e_tgt[idx];
    const float ae = fabsf(de);
    se = fmaf(ae < 1.f ? 0.5f * de * de : ae - 0.5f, wi, se);
    g_e[idx] = fminf(fmaxf(de, -1.f), 1.f) * wi * inv_pe;
  }
  sd = block_sum(sd, red);
  sp = block_sum(sp, red);
  se = block_sum(se, red);
  if (threadIdx.x == 0) {
    losses[0] = sd * inv_d;
    losses[1] = sp * inv_pe;
    losses[2] = se * inv_pe;
  }
}

// align_loss = forward-sum loss + bin loss (reference generator/__init__.py:174-175, alignments.py:236-238):
//   bin_loss = -(1/B) sum_b mean_{t < T_b} log_p_attn[b, t, path[b,t]]
// One CTA per sample gathers along the alignment path (27 648 scattered reads at the benchmark shape: a single CTA took
// 50-160 us of pure memory latency on the step's critical path); the bin-loss gradient -1/(B T_b) is added IN PLACE to the
// forward-sum gradient (both terms enter align_loss with weight one).  The per-sample partial sums go to a small workspace
// and the LAST CTA to finish adds them up in sample order (deterministic, no float atomics) and re-arms the counter.
__global__ void __launch_bounds__(FL_THREADS)
align_loss_fold_kernel(const float* __restrict__ lp, const int* __restrict__ path, const long long* __restrict__ m_len,
                       const float* __restrict__ per_sample_fs, float* __restrict__ fs_grad, float* __restrict__ out,
                       float* __restrict__ ws /* [B] bin partials, then one counter word */, int B, int Tm, int Tx) {
  __shared__ float red[32];
  __shared__ int s_last;
  const int b = blockIdx.x;
  const float inv_b = 1.f / static_cast<float>(B);
  const int T = static_cast<int>(m_len[b]) < Tm ? static_cast<int>(m_len[b]) : Tm;
  const float g = T > 0 ? -inv_b / static_cast<float>(T) : 0.f;
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += FL_THREADS) {
    const int n = path[static_cast<long long>(b) * Tm + t];
    if (n >= 0 && n < Tx) {
      const long long idx = (static_cast<long long>(b) * Tm + t) * Tx + n;
      s += lp[idx];
      fs_grad[idx] += g;
    }
  }
  s = block_sum(s, red);
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws + B);
  if (threadIdx.x == 0) {
    ws[b] = T > 0 ? -(s / static_cast<float>(T)) * inv_b : 0.f;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == static_cast<unsigned int>(B - 1)) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    float fs_tot = 0.f, bin_tot = 0.f;
    for (int i = 0; i < B; ++i) {
      fs_tot += per_sample_fs[i] * inv_b;
      bin_tot += reinterpret_cast<volatile float*>(ws)[i];
    }
    out[0] = fs_tot + bin_tot;
    out[1] = fs_tot;
    out[2] = bin_tot;
    *counter = 0u;   // re-armed for the next launch (CUDA-graph replays reuse the workspace)
  }
}

}  // namespace
}  // namespace osb

extern "C" int osb_align_loss_fold(const float* log_p_attn, const int32_t* path, const int64_t* m_len, const float* per_sample_fs,
                                   float* fs_grad, float* out, float* workspace, int32_t B, int32_t Tm, int32_t Tx, void* stream) {
  OSB_REQUIRE(log_p_attn && path && m_len && per_sample_fs && fs_grad && out && workspace, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tm > 0 && Tx > 0, OSB_ERR_SHAPE);
  osb::align_loss_fold_kernel<<<B, osb::FL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      log_p_attn, path, reinterpret_cast<const long long*>(m_len), per_sample_fs, fs_grad, out, workspace, B, Tm, Tx);
  osb::count_launch();
  return osb::launch_status();
}

extern "C" int osb_fs2_losses(const float* d_hat, const float* p_hat, const float* e_hat, const float* ds, const float* p_tgt,
                              const float* e_tgt, const int64_t* x_len, float* losses, float* g_d, float* g_p, float* g_e, int32_t B,
                              int32_t Tx, void* stream) {
  OSB_REQUIRE(d_hat && p_hat && e_hat && ds && p_tgt && e_tgt && x_len && losses && g_d && g_p && g_e, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tx > 0 && (static_cast<size_t>(Tx) + B + 32) * sizeof(float) <= 48 * 1024, OSB_ERR_SHAPE);
  const size_t smem = (static_cast<size_t>(Tx) + B + 32) * sizeof(float);
  osb::fs2_losses_kernel<<<1, osb::FL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      d_hat, p_hat, e_hat, ds, p_tgt, e_tgt, reinterpret_cast<const long long*>(x_len), losses, g_d, g_p, g_e, B, Tx);
  osb::count_launch();
  return osb::launch_status();
}
