// osb_optim.cu — optimizer kernels over one flat fp32 parameter / gradient bucket:
// global gradient norm (with non-finite detection) and fused unscale + clip-by-global-norm + AdamW.
// HBM-bound: one pass over the gradient for the norm, one pass over (p, g, m, v) for the update.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

// stats[0] += sum g^2 ; stats[1] = 1 if any element is non-finite
// Sum of squares of the (all-reduced) flat gradient, DETERMINISTIC: block partials are combined by the last block in index
// order, so every rank of a data-parallel job derives bit-identical clip coefficients from its bit-identical copy of the
// reduced bucket (an atomic accumulation order would let the ranks' parameters drift apart by an ulp per step).
constexpr int SUMSQ_MAX_BLOCKS = 148 * 4;
__device__ float g_sumsq_partial[SUMSQ_MAX_BLOCKS];
__device__ int g_sumsq_bad[SUMSQ_MAX_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ stats) {
  float acc = 0.f;
  bool bad = false;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc += g[i] * g[i];
  bad = !isfinite(acc);
  acc = warp_sum(acc);
  __shared__ float red[8];
  __shared__ int sbad;
  __shared__ bool last;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) sbad = 1;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    g_sumsq_partial[blockIdx.x] = s;
    g_sumsq_bad[blockIdx.x] = (sbad || !isfinite(s)) ? 1 : 0;
    __threadfence();
    last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 32) {   // one warp, fixed order: lane l takes partials l, l+32, ...; then a fixed shuffle tree
    float s = 0.f;
    int b = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) {
      s += *(volatile float*)&g_sumsq_partial[i];
      b |= *(volatile int*)&g_sumsq_bad[i];
    }
    s = warp_sum(s);
    b = __any_sync(0xffffffffu, b != 0);
    if (threadIdx.x == 0) {
      stats[0] += s;
      if (b) stats[1] = 1.f;
      g_sumsq_ticket = 0;
    }
  }
}

// Multi-tensor gather: the step's gradient tensors (wherever autograd allocated them) -> the flat bucket, with the global
// sum of squares and the non-finite flag accumulated on the way (stats may be null under data parallelism, where the norm is
// taken after the all-reduce).  tab = [src pointer, destination offset, numel] per tensor; chunks = (tensor, chunk index),
// GATHER_CHUNK floats each; a null source (no gradient this step) writes zeros.
constexpr int GATHER_CHUNK = 2048;

__global__ void __launch_bounds__(256)
grad_gather_kernel(const long long* __restrict__ tab, const int2* __restrict__ chunks, long long n_chunks, float* __restrict__ flat,
                   float* __restrict__ stats) {
  float acc = 0.f;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int2 ch = chunks[c];
    const float* src = reinterpret_cast<const float*>(tab[3 * ch.x]);
    const long long start = static_cast<long long>(ch.y) * GATHER_CHUNK;
    const long long n = tab[3 * ch.x + 2];
    const int len = static_cast<int>(n - start < GATHER_CHUNK ? n - start : GATHER_CHUNK);
    float* dst = flat + tab[3 * ch.x + 1] + start;
    if (src == nullptr) {
      for (int i = threadIdx.x; i < len; i += 256) dst[i] = 0.f;
      continue;
    }
    src += start;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {  // dst is 16-byte aligned by construction (offsets and chunks are multiples of 4)
      const int len4 = len >> 2;
      for (int i = threadIdx.x; i < len4; i += 256) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        reinterpret_cast<float4*>(dst)[i] = v;
        acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      }
      for (int i = (len4 << 2) + threadIdx.x; i < len; i += 256) {
        const float v = src[i];
        dst[i] = v;
        acc += v * v;
      }
    } else {
      for (int i = threadIdx.x; i < len; i += 256) {
        const float v = src[i];
        dst[i] = v;
        acc += v * v;
      }
    }
  }
  if (stats == nullptr) return;
  const bool bad = !isfinite(acc);
  acc = warp_sum(acc);
  __shared__ float red[8];
  __shared__ int sbad;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) sbad = 1;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(stats, s);
    if (sbad || !isfinite(s)) stats[1] = 1.f;
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, max_norm, inv_scale;
  const float* hyper;  // optional device [lr, bc1, bc2_sqrt] overriding the by-value fields (CUDA-graph replay)
};

// torch.optim.AdamW (decoupled weight decay) preceded by clip_grad_norm_(max_norm) on the unscaled gradient:
//   total = sqrt(stats[0]) * inv_scale ; coef = min(1, max_norm / (total + 1e-6)) ; g = g * inv_scale * coef
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// The whole update is skipped when the gradient holds a non-finite value (stats[1] != 0).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             const float* __restrict__ stats, AdamArgs a) {
  if (stats[1] != 0.f) return;
  if (a.hyper != nullptr) {
    a.lr = a.hyper[0];
    a.bc1 = a.hyper[1];
    a.bc2_sqrt = a.hyper[2];
  }
  const float total = sqrtf(stats[0]) * a.inv_scale;
  float coef = a.max_norm > 0.f ? a.max_norm / (total + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f) * a.inv_scale;
  const float decay = 1.f - a.lr * a.weight_decay;
  const float step = a.lr / a.bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * decay;
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    pi -= step * mi / (sqrtf(vi) / a.bc2_sqrt + a.eps);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_grad_sumsq(const float* g, int64_t n, float* stats, void* stream) {
  OSB_REQUIRE(g && stats, OSB_ERR_ARG);
  OSB_REQUIRE(n > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, OSB_ERR_ALIGN);
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, stats);
  count_launch();
  return launch_status();
}

extern "C" int osb_grad_gather(const int64_t* table, const int32_t* chunks, int64_t n_chunks, float* flat_g, float* stats, void* stream) {
  OSB_REQUIRE(table && chunks && flat_g, OSB_ERR_ARG);
  OSB_REQUIRE(n_chunks > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE((reinterpret_cast<uintptr_t>(flat_g) & 15) == 0 && (reinterpret_cast<uintptr_t>(chunks) & 7) == 0, OSB_ERR_ALIGN);
  long long blocks = n_chunks < 148 * 8 ? n_chunks : 148 * 8;
  grad_gather_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(table), reinterpret_cast<const int2*>(chunks), n_chunks, flat_g, stats);
  count_launch();
  return launch_status();
}

extern "C" int osb_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* stats, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int64_t step, float max_norm, float inv_scale, void* stream) {
  OSB_REQUIRE(p && g && m && v && stats, OSB_ERR_ARG);
  OSB_REQUIRE(n > 0 && step > 0, OSB_ERR_SHAPE);
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, static_cast<float>(step));
  a.bc2_sqrt = sqrtf(1.f - powf(beta2, static_cast<float>(step)));
  a.max_norm = max_norm; a.inv_scale = inv_scale; a.hyper = nullptr;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, stats, a);
  count_launch();
  return launch_status();
}

extern "C" int osb_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* stats, const float* hyper,
                                  float beta1, float beta2, float eps, float weight_decay, float max_norm, float inv_scale,
                                  void* stream) {
  OSB_REQUIRE(p && g && m && v && stats && hyper, OSB_ERR_ARG);
  OSB_REQUIRE(n > 0, OSB_ERR_SHAPE);
  AdamArgs a;
  a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.bc1 = 1.f; a.bc2_sqrt = 1.f;
  a.max_norm = max_norm; a.inv_scale = inv_scale; a.hyper = hyper;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, stats, a);
  count_launch();
  return launch_status();
}
