// osb_pointwise.cu — HBM-bound kernels of the synthesis path (everything that is not a GEMM):
// text embedding, depthwise-conv7 + LayerNorm, standalone LayerNorm, variance embedding,
// duration rounding + scan, Gaussian upsampling, hard length-regulator gather, fp32->fp16 packing.
//
// Layout: channels-last rows of C contiguous fp32 values; one warp owns one row (position) and
// each lane owns C/32 channels as float4 groups, so every global access is a 512-byte coalesced
// warp transaction.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ void store_h4(__half* dst, float a, float b, float c, float d) {
  __half2 h0 = __floats2half2_rn(a, b);
  __half2 h1 = __floats2half2_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(dst) = u;
}

// hi/lo split store: dst_hi[0..3] = fp16(v), dst_lo[0..3] = fp16(v - hi)
__device__ __forceinline__ void store_h4_split(__half* dst_hi, __half* dst_lo, float a, float b, float c, float d) {
  store_h4(dst_hi, a, b, c, d);
  store_h4(dst_lo, a - __half2float(__float2half_rn(a)), b - __half2float(__float2half_rn(b)),
           c - __half2float(__float2half_rn(c)), d - __half2float(__float2half_rn(d)));
}
// row pointer + column for an fp16 destination that is either plain (rows of C) or split (rows of [hi C | lo C])
__device__ __forceinline__ void store_h4_row(__half* base, long long row, int C, int c, int split, float a, float b, float cc,
                                             float d) {
  if (split) {
    __half* r = base + row * (2LL * C);
    store_h4_split(r + c, r + C + c, a, b, cc, d);
  } else {
    store_h4(base + row * C + c, a, b, cc, d);
  }
}

// ------------------------------------------------------------------------------------------
// text embedding: out[b,t,:] = sqrt(dim) * E[id] + scale * [sin(t*f) | cos(t*f)]
// ------------------------------------------------------------------------------------------
__global__ void embed_text_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                  const float* __restrict__ inv_freq, const float* __restrict__ scale_p, float* __restrict__ out,
                                  int rows, int T, int dim, int n_vocab, float embed_scale) {
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int t = row % T;
  long long id = ids[row];
  if (id < 0 || id >= n_vocab) id = 0;
  const float scale = scale_p[0];
  const int half = dim >> 1;
  const float* e = table + id * dim;
  float* o = out + static_cast<long long>(row) * dim;
  for (int c = lane; c < dim; c += 32) {
    const int j = c < half ? c : c - half;
    const float ang = static_cast<float>(t) * inv_freq[j];
    const float pe = c < half ? sinf(ang) : cosf(ang);
    o[c] = embed_scale * e[c] + pe * scale;
  }
}

// ------------------------------------------------------------------------------------------
// depthwise conv (k=7, zero pad 3, bias) + LayerNorm statistics.
// Writes xhat = (d - mean) * rstd as fp16 (the LN affine is folded into the following pointwise
// GEMM at weight-packing time) and optionally rstd (needed by the backward pass).
// Each warp walks POS consecutive positions with a 7-row sliding window held in registers.
// ------------------------------------------------------------------------------------------
template <int VPL>  // float4 groups per lane: C = 128 * VPL
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
dwconv_ln_kernel(const float* __restrict__ x, const float* __restrict__ w /*(C,7)*/, const float* __restrict__ bias,
                 __half* __restrict__ xhat, float* __restrict__ rstd_out, int B, int T, int pos_per_warp, float eps, int split) {
  constexpr int C = 128 * VPL;
  __shared__ float sw[7][C];
  __shared__ float sb[C];
  for (int i = threadIdx.x; i < C * 7; i += blockDim.x) sw[i % 7][i / 7] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int groups_per_batch = (T + pos_per_warp - 1) / pos_per_warp;
  const int gw = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (gw >= B * groups_per_batch) return;
  const int b = gw / groups_per_batch;
  const int t_begin = (gw % groups_per_batch) * pos_per_warp;
  const int t_end = min(T, t_begin + pos_per_warp);
  const float* xb = x + static_cast<long long>(b) * T * C;

  float4 win[7][VPL];  // win[j] = row (t + j - 3)
  auto load_row = [&](int t, float4 (&dst)[VPL]) {
    if (t >= 0 && t < T) {
#pragma unroll
      for (int v = 0; v < VPL; ++v) dst[v] = *reinterpret_cast<const float4*>(xb + static_cast<long long>(t) * C + v * 128 + lane * 4);
    } else {
#pragma unroll
      for (int v = 0; v < VPL; ++v) dst[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
#pragma unroll
  for (int j = 0; j < 6; ++j) load_row(t_begin + j - 3, win[j + 1]);

  float4 nxt[VPL];   // row t + 3 of the coming iteration, requested one iteration ahead (its latency hides behind the FMAs)
  load_row(t_begin + 3, nxt);
  for (int t = t_begin; t < t_end; ++t) {
#pragma unroll
    for (int j = 0; j < 6; ++j)
#pragma unroll
      for (int v = 0; v < VPL; ++v) win[j][v] = win[j + 1][v];
#pragma unroll
    for (int v = 0; v < VPL; ++v) win[6][v] = nxt[v];
    if (t + 1 < t_end) load_row(t + 4, nxt);

    float4 d[VPL];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = v * 128 + lane * 4;
      float4 acc = *reinterpret_cast<const float4*>(&sb[c]);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const float4 wj = *reinterpret_cast<const float4*>(&sw[j][c]);
        acc.x = fmaf(wj.x, win[j][v].x, acc.x);
        acc.y = fmaf(wj.y, win[j][v].y, acc.y);
        acc.z = fmaf(wj.z, win[j][v].z, acc.z);
        acc.w = fmaf(wj.w, win[j][v].w, acc.w);
      }
      d[v] = acc;
      s += (acc.x + acc.y) + (acc.z + acc.w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      d[v].x -= mean; d[v].y -= mean; d[v].z -= mean; d[v].w -= mean;
      q += (d[v].x * d[v].x + d[v].y * d[v].y) + (d[v].z * d[v].z + d[v].w * d[v].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    const long long row = static_cast<long long>(b) * T + t;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
      store_h4_row(xhat, row, C, v * 128 + lane * 4, split, d[v].x * rstd, d[v].y * rstd, d[v].z * rstd, d[v].w * rstd);
    if (rstd_out != nullptr && lane == 0) rstd_out[row] = rstd;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim: one warp per row. Optional fp32 and fp16 outputs.
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out_f32,
                 __half* __restrict__ out_h16, long long rows, float eps, int split) {
  constexpr int C = 128 * VPL;
  const long long row = static_cast<long long>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 d[VPL];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    d[v] = *reinterpret_cast<const float4*>(x + row * C + v * 128 + lane * 4);
    s += (d[v].x + d[v].y) + (d[v].z + d[v].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    d[v].x -= mean; d[v].y -= mean; d[v].z -= mean; d[v].w -= mean;
    q += (d[v].x * d[v].x + d[v].y * d[v].y) + (d[v].z * d[v].z + d[v].w * d[v].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = v * 128 + lane * 4;
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    float4 y;
    y.x = d[v].x * rstd * ww.x + bb.x;
    y.y = d[v].y * rstd * ww.y + bb.y;
    y.z = d[v].z * rstd * ww.z + bb.z;
    y.w = d[v].w * rstd * ww.w + bb.w;
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + row * C + c) = y;
    if (out_h16 != nullptr) store_h4_row(out_h16, row, C, c, split, y.x, y.y, y.z, y.w);
  }
}

// ------------------------------------------------------------------------------------------
// y = LayerNorm(relu(x)) * w + b over the last dim (one warp per row) -> fp16 (plain or split) and / or the masked
// Linear(C -> 1) of a VariancePredictor's tail.  The stand-alone form of osb_gemm's RELU_LN epilogue, for problems of a few
// row tiles where that epilogue forces ONE CTA per 128 rows to stream the whole weight matrix (see osb_relu_layernorm).
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
relu_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, __half* __restrict__ out_h16,
                      long long rows, float eps, int split, const float* __restrict__ dot_w, const float* __restrict__ dot_b,
                      const uint8_t* __restrict__ pad_mask, float* __restrict__ out_dot) {
  constexpr int C = 128 * VPL;
  const long long row = static_cast<long long>(blockIdx.x) * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float4 d[VPL];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    d[v] = *reinterpret_cast<const float4*>(x + row * C + v * 128 + lane * 4);
    d[v].x = fmaxf(d[v].x, 0.f); d[v].y = fmaxf(d[v].y, 0.f); d[v].z = fmaxf(d[v].z, 0.f); d[v].w = fmaxf(d[v].w, 0.f);
    s += (d[v].x + d[v].y) + (d[v].z + d[v].w);
  }
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    d[v].x -= mean; d[v].y -= mean; d[v].z -= mean; d[v].w -= mean;
    q += (d[v].x * d[v].x + d[v].y * d[v].y) + (d[v].z * d[v].z + d[v].w * d[v].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  float dot = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = v * 128 + lane * 4;
    const float4 ww = *reinterpret_cast<const float4*>(w + c);
    const float4 bb = *reinterpret_cast<const float4*>(b + c);
    float4 y;
    y.x = d[v].x * rstd * ww.x + bb.x;
    y.y = d[v].y * rstd * ww.y + bb.y;
    y.z = d[v].z * rstd * ww.z + bb.z;
    y.w = d[v].w * rstd * ww.w + bb.w;
    if (out_h16 != nullptr) store_h4_row(out_h16, row, C, c, split, y.x, y.y, y.z, y.w);
    if (dot_w != nullptr) {
      const float4 dw = *reinterpret_cast<const float4*>(dot_w + c);
      dot += (y.x * dw.x + y.y * dw.y) + (y.z * dw.z + y.w * dw.w);
    }
  }
  if (out_dot != nullptr) {
    dot = warp_sum(dot);
    if (lane == 0) out_dot[row] = (pad_mask != nullptr && pad_mask[row]) ? 0.f : dot + (dot_b != nullptr ? dot_b[0] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------
// variance embedding: out = (x + bias + conv1d(val; 1->C, k taps, same)) * keep
// ------------------------------------------------------------------------------------------
__global__ void variance_embed_kernel(const float* __restrict__ x, const float* __restrict__ val, const float* __restrict__ w /*(C,k)*/,
                                      const float* __restrict__ bias, const uint8_t* __restrict__ pad_mask,
                                      const float* __restrict__ emb_scale, float* __restrict__ out_f32,
                                      __half* __restrict__ out_h16, int B, int T, int C, int ksize, int split) {
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= B * T) return;
  const int lane = threadIdx.x & 31;
  const int b = row / T, t = row % T;
  const int halfk = (ksize - 1) / 2;
  float v[16];
  for (int j = 0; j < ksize; ++j) {
    const int tt = t + j - halfk;
    v[j] = (tt >= 0 && tt < T) ? val[b * T + tt] : 0.f;
  }
  const float keep = (pad_mask != nullptr && pad_mask[row]) ? 0.f : 1.f;
  for (int c = lane; c < C; c += 32) {
    float acc = bias[c];
    for (int j = 0; j < ksize; ++j) acc = fmaf(w[c * ksize + j], v[j], acc);
    if (emb_scale != nullptr) acc *= emb_scale[static_cast<long long>(row) * C + c];  // dropout on the embedding branch
    const float y = (x[static_cast<long long>(row) * C + c] + acc) * keep;
    if (out_f32 != nullptr) out_f32[static_cast<long long>(row) * C + c] = y;
    if (out_h16 != nullptr) {
      const __half hi = __float2half_rn(y);
      if (split) {
        out_h16[static_cast<long long>(row) * 2 * C + c] = hi;
        out_h16[static_cast<long long>(row) * 2 * C + C + c] = __float2half_rn(y - __half2float(hi));
      } else {
        out_h16[static_cast<long long>(row) * C + c] = hi;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// durations: d = clamp(ceil((exp(logd) - clip) * factor), 0) as int64, 0 at pads; per-sample
// length, and Gaussian centres c = cumsum(d) - d/2.  One block per sample (Tx <= 4096).
// ------------------------------------------------------------------------------------------
__global__ void duration_kernel(const float* __restrict__ log_d, const uint8_t* __restrict__ pad_mask, long long* __restrict__ dur,
                                long long* __restrict__ lengths, int T, float factor, float clip_val) {
  const int b = blockIdx.x;
  long long local = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float d = ceilf((expf(log_d[b * T + t]) - clip_val) * factor);
    long long di = static_cast<long long>(d);  // matches .long() (truncation) for finite values
    if (!(d == d)) di = 0;
    if (di < 0) di = 0;
    if (pad_mask[b * T + t]) di = 0;
    dur[b * T + t] = di;
    local += di;
  }
  __shared__ long long red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
    lengths[b] = s;
  }
}

// cumsum of durations (int64 or fp32 input) -> centres (fp32) and inclusive cumsum (int64). One warp per sample.
template <typename DT>
__global__ void centres_kernel(const DT* __restrict__ dur, float* __restrict__ centres, long long* __restrict__ csum_out, int B, int T) {
  const int b = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  double carry = 0.0;  // exact for integer-valued durations
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const double d = t < T ? static_cast<double>(dur[b * T + t]) : 0.0;
    double s = d;
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    s += carry;
    if (t < T) {
      // reference: ds.cumsum(-1) - ds / 2, evaluated in fp32 (alignments.py:167)
      centres[b * T + t] = static_cast<float>(s) - static_cast<float>(d) * 0.5f;
      if (csum_out != nullptr) csum_out[b * T + t] = static_cast<long long>(s);
    }
    carry = __shfl_sync(0xffffffffu, s, 31);
  }
}

// ------------------------------------------------------------------------------------------
// Gaussian upsampling: y[b,t,:] = softmax_i(-delta (t*hmask - c_i)^2 | valid i) @ hs[b,:,:]
// Block = FR frames x C channels; the attention row is computed once per frame in shared memory.
// ------------------------------------------------------------------------------------------
template <int FR>
__global__ void __launch_bounds__(256)
gaussian_upsample_kernel(const float* __restrict__ hs, const float* __restrict__ centres, const long long* __restrict__ x_len,
                         const long long* __restrict__ y_len, float* __restrict__ out_f32, __half* __restrict__ out_h16, int Tx, int Tm,
                         int C, float delta, const long long* __restrict__ win_start, int halo, int Tfull) {
  // Window mode (win_start != NULL): output row j of sample b is frame win_start[b] - halo + j of the full (Tfull-frame)
  // sequence; rows that fall outside [0, Tfull) are zero (what a convolution's zero padding reads there).  Tm = rows per sample.
  extern __shared__ float sp[];  // FR * Tx attention weights
  __shared__ int s_rng[2];       // tokens [lo, hi] with a non-zero weight for any frame of the block
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int nx = static_cast<int>(x_len[b]);
  const int ny = static_cast<int>(y_len[b]);
  const int shift = win_start != nullptr ? static_cast<int>(win_start[b]) - halo : 0;   // full-sequence frame of output row 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { s_rng[0] = Tx; s_rng[1] = -1; }
  __syncthreads();
  // one warp per frame builds the softmax row
  for (int f = warp; f < FR; f += (blockDim.x >> 5)) {
    if (t0 + f >= Tm) continue;
    const int t = t0 + f + shift;
    if (t < 0 || t >= Tfull) {           // outside the sequence (window mode only): an all-zero row
      for (int i = lane; i < Tx; i += 32) sp[f * Tx + i] = 0.f;
      continue;
    }
    const float tf = (t < ny) ? static_cast<float>(t) : 0.f;
    float* p = sp + f * Tx;
    float mx = -INFINITY;
    for (int i = lane; i < Tx; i += 32) {
      float e = -INFINITY;
      if (i < nx) {
        const float diff = tf - centres[b * Tx + i];
        e = -delta * (diff * diff);
      }
      p[i] = e;
      mx = fmaxf(mx, e);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < Tx; i += 32) {
      const float e = (i < nx) ? expf(p[i] - mx) : 0.f;
      p[i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    // exp(-delta d^2) underflows to an exact zero a few dozen frames away from a token's centre: the product below skips the
    // tokens whose weight is zero for every frame of the block (adding 0 * h changes nothing: same sums, ~10x fewer terms)
    int lo = Tx, hi = -1;
    for (int i = lane; i < Tx; i += 32) {
      const float w = p[i] * inv;
      p[i] = w;
      if (w != 0.f) { lo = min(lo, i); hi = max(hi, i); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { atomicMin(&s_rng[0], lo); atomicMax(&s_rng[1], hi); }
  }
  __syncthreads();
  const int i_lo = s_rng[0], i_hi = min(s_rng[1], nx - 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[FR];
#pragma unroll
    for (int f = 0; f < FR; ++f) acc[f] = 0.f;
    const float* h = hs + static_cast<long long>(b) * Tx * C + c;
    for (int i = i_lo; i <= i_hi; ++i) {
      const float hv = h[static_cast<long long>(i) * C];
#pragma unroll
      for (int f = 0; f < FR; ++f) acc[f] = fmaf(sp[f * Tx + i], hv, acc[f]);
    }
#pragma unroll
    for (int f = 0; f < FR; ++f) {
      const int t = t0 + f;
      if (t < Tm) {
        const long long o = (static_cast<long long>(b) * Tm + t) * C + c;
        if (out_f32 != nullptr) out_f32[o] = acc[f];
        if (out_h16 != nullptr) out_h16[o] = __float2half_rn(acc[f]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// hard length regulator (expand_by_duration): out[b,t,:] = x[b, token(t), :], token(t) = first i
// with csum[i] > t; zero beyond the sample's length.  Also emits the index (bit-exact target).
// ------------------------------------------------------------------------------------------
__global__ void expand_gather_kernel(const float* __restrict__ x, const long long* __restrict__ csum, float* __restrict__ out,
                                     int* __restrict__ index_out, int B, int Tx, int Tm, int C) {
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= B * Tm) return;
  const int lane = threadIdx.x & 31;
  const int b = row / Tm, t = row % Tm;
  const long long* cs = csum + static_cast<long long>(b) * Tx;
  int idx = -1;
  if (t < cs[Tx - 1]) {
    int lo = 0, hi = Tx - 1;  // first i with cs[i] > t
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cs[mid] > t) hi = mid; else lo = mid + 1;
    }
    idx = lo;
  }
  if (index_out != nullptr && lane == 0) index_out[row] = idx;
  for (int c = lane; c < C; c += 32)
    out[static_cast<long long>(row) * C + c] = idx >= 0 ? x[(static_cast<long long>(b) * Tx + idx) * C + c] : 0.f;
}

// ------------------------------------------------------------------------------------------
// fp32 -> fp16 packing with optional per-column scale and row/column padding:
//   dst[r, c] = half(src[r * src_ld + c * src_cs] * (col_scale ? col_scale[c] : 1))   for r < rows, c < cols
//   zero elsewhere (c in [cols, dst_ld))
// ------------------------------------------------------------------------------------------
__global__ void pack_h16_kernel(const float* __restrict__ src, long long src_ld, long long src_cs, const float* __restrict__ col_scale,
                                __half* __restrict__ dst, __half* __restrict__ dst_lo, long long dst_rs, int dst_cols, long long rows,
                                int cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * dst_cols) return;
  const long long r = i / dst_cols;
  const int c = static_cast<int>(i % dst_cols);
  float v = 0.f;
  if (c < cols) {
    v = src[r * src_ld + c * src_cs];
    if (col_scale != nullptr) v *= col_scale[c];
  }
  const __half hi = __float2half_rn(v);
  dst[r * dst_rs + c] = hi;
  if (dst_lo != nullptr) dst_lo[r * dst_rs + c] = __float2half_rn(v - __half2float(hi));
}

// Contiguous-source fast path of pack_h16: a thread converts 8 consecutive elements of a row (two 16-byte loads, one 16-byte
// store per output).  Needs src_cs == 1, cols % 8 == 0 (so a group is either all data or all padding) and 16-byte aligned rows.
__global__ void __launch_bounds__(256)
pack_h16_vec8_kernel(const float* __restrict__ src, long long src_ld, const float* __restrict__ col_scale, __half* __restrict__ dst,
                     __half* __restrict__ dst_lo, long long dst_rs, int dst_cols, long long rows, int cols) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = dst_cols >> 3;
  if (i >= rows * groups) return;
  const long long r = i / groups;
  const int c = static_cast<int>(i - r * groups) << 3;
  float v[8];
  if (c < cols) {
    const float4 a = *reinterpret_cast<const float4*>(src + r * src_ld + c);
    const float4 b = *reinterpret_cast<const float4*>(src + r * src_ld + c + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    if (col_scale != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] *= col_scale[c + q];
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
  }
  __half h[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) h[q] = __float2half_rn(v[q]);
  *reinterpret_cast<uint4*>(dst + r * dst_rs + c) = *reinterpret_cast<const uint4*>(h);
  if (dst_lo != nullptr) {
    __half l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) l[q] = __float2half_rn(v[q] - __half2float(h[q]));
    *reinterpret_cast<uint4*>(dst_lo + r * dst_rs + c) = *reinterpret_cast<const uint4*>(l);
  }
}

// Conv1d weight (N, Cin, k) fp32 -> fp16 tensor-core operand, all taps in one launch.
//   forward form  (transpose_reverse = 0): dst[tap][n][c] = w[n][c][tap]            (k, N, Kp), zero for c >= Cin
//   dgrad form    (transpose_reverse = 1): dst[tap][c][n] = w[n][c][k-1-tap]        (k, Cin, N)
// dst_lo (optional) receives the fp16 rounding residual (split precision).
__global__ void pack_conv_h16_kernel(const float* __restrict__ w, __half* __restrict__ dst, __half* __restrict__ dst_lo, int N, int Cin,
                                     int k, int Kp, int transpose_reverse) {
  const long long total = transpose_reverse ? static_cast<long long>(k) * Cin * N : static_cast<long long>(k) * N * Kp;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = 0.f;
  if (transpose_reverse) {
    const int n = static_cast<int>(i % N);
    const int c = static_cast<int>((i / N) % Cin);
    const int tap = static_cast<int>(i / (static_cast<long long>(N) * Cin));
    v = w[(static_cast<long long>(n) * Cin + c) * k + (k - 1 - tap)];
  } else {
    const int c = static_cast<int>(i % Kp);
    const int n = static_cast<int>((i / Kp) % N);
    const int tap = static_cast<int>(i / (static_cast<long long>(Kp) * N));
    if (c < Cin) v = w[(static_cast<long long>(n) * Cin + c) * k + tap];
  }
  const __half hi = __float2half_rn(v);
  dst[i] = hi;
  if (dst_lo != nullptr) dst_lo[i] = __float2half_rn(v - __half2float(hi));
}

// Every weight pack of a training step in ONE launch: jobs[] (device) lists (source, destination, layout); a thread owns 8
// consecutive destination elements (one 16-byte store) and finds its job by binary search over the prefix sums of the
// destination sizes (multiples of 8).  kind 0: (rows, cols) fp32 row-major (optionally scaled per column) -> (rows, dst_cols)
// fp16; kind 1: Conv1d weight (rows = N, cols = Cin, k) -> (k, N, dst_cols) fp16; kind 2: fp32 matrix-vector product
// dst[r] = aux[r] + sum_c src[r, c] * col_scale[c] (the LayerNorm bias folded into the pwconv1 bias), 8 virtual elements
// per row.
__global__ void __launch_bounds__(256) pack_multi_kernel(const osb_pack_job* __restrict__ jobs, int n_jobs, long long total) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= total) return;
  int lo = 0, hi = n_jobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_elem <= i) lo = mid; else hi = mid - 1;
  }
  const osb_pack_job j = jobs[lo];
  const long long e = i - j.first_elem;
  const float* src = static_cast<const float*>(j.src);
  const float* cs = static_cast<const float*>(j.col_scale);
  if (j.kind == 2) {
    const long long r = e >> 3;
    const float* row = src + r * j.cols;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int c = 0;
    for (; c + 4 <= j.cols; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(row + c);
      const float4 v4 = *reinterpret_cast<const float4*>(cs + c);
      acc0 = fmaf(w4.x, v4.x, acc0); acc1 = fmaf(w4.y, v4.y, acc1); acc2 = fmaf(w4.z, v4.z, acc2); acc3 = fmaf(w4.w, v4.w, acc3);
    }
    for (; c < j.cols; ++c) acc0 = fmaf(row[c], cs[c], acc0);
    static_cast<float*>(j.dst)[r] = (j.aux != nullptr ? static_cast<const float*>(j.aux)[r] : 0.f) + ((acc0 + acc1) + (acc2 + acc3));
    return;
  }
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = 0.f;
  const int c = static_cast<int>(e % j.dst_cols);
  if (j.kind == 0) {
    const long long r = e / j.dst_cols;
    const float* row = src + r * j.cols;
    if (c + 8 <= j.cols && ((reinterpret_cast<uintptr_t>(row + c) & 15) == 0)) {
      const float4 a4 = *reinterpret_cast<const float4*>(row + c), b4 = *reinterpret_cast<const float4*>(row + c + 4);
      v[0] = a4.x; v[1] = a4.y; v[2] = a4.z; v[3] = a4.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) if (c + q < j.cols) v[q] = row[c + q];
    }
    if (cs != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) if (c + q < j.cols) v[q] *= cs[c + q];
    }
  } else {
    const int n = static_cast<int>((e / j.dst_cols) % j.rows);
    const int tap = static_cast<int>(e / (static_cast<long long>(j.dst_cols) * j.rows));
#pragma unroll
    for (int q = 0; q < 8; ++q) if (c + q < j.cols) v[q] = src[(static_cast<long long>(n) * j.cols + c + q) * j.k + tap];
  }
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(static_cast<__half*>(j.dst) + e) = u;
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_pack_multi(const osb_pack_job* jobs_dev, int32_t n_jobs, int64_t total_elems, void* stream) {
  OSB_REQUIRE(jobs_dev != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(n_jobs > 0 && total_elems > 0 && total_elems % 8 == 0, OSB_ERR_SHAPE);
  pack_multi_kernel<<<static_cast<unsigned>((total_elems / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, n_jobs,
                                                                                                                       total_elems);
  count_launch();
  return launch_status();
}

extern "C" int osb_pack_conv_h16(const float* w, void* dst, void* dst_lo, int32_t N, int32_t Cin, int32_t k, int32_t Kp,
                                 int32_t transpose_reverse, void* stream) {
  OSB_REQUIRE(w && dst, OSB_ERR_ARG);
  OSB_REQUIRE(N > 0 && Cin > 0 && k > 0 && (transpose_reverse || Kp >= Cin), OSB_ERR_SHAPE);
  const long long total = transpose_reverse ? static_cast<long long>(k) * Cin * N : static_cast<long long>(k) * N * Kp;
  pack_conv_h16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__half*>(dst), static_cast<__half*>(dst_lo), N, Cin, k, Kp, transpose_reverse);
  count_launch();
  return launch_status();
}

extern "C" int osb_embed_text(const int64_t* ids, const float* table, const float* inv_freq, const float* scale, float* out,
                              int32_t B, int32_t T, int32_t dim, int32_t n_vocab, void* stream) {
  OSB_REQUIRE(ids && table && inv_freq && scale && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && dim > 0 && dim % 2 == 0, OSB_ERR_SHAPE);
  const int rows = B * T;
  embed_text_kernel<<<(rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), table, inv_freq, scale, out, rows, T, dim, n_vocab, sqrtf(static_cast<float>(dim)));
  count_launch();
  return launch_status();
}

extern "C" int osb_dwconv_ln(const float* x, const float* w, const float* bias, void* xhat_h16, float* rstd, int32_t B, int32_t T,
                             int32_t C, float eps, int32_t split, void* stream) {
  OSB_REQUIRE(x && w && bias && xhat_h16, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && (C == 256 || C == 384 || C == 128 || C == 512), OSB_ERR_SHAPE);
  // positions per warp: long runs re-use the 7-row window (1.4x reads at 16), short runs fill the machine when there are few
  // positions (a B=1 utterance: 243 frames = 16 warps at 16 positions each, 13 us of serial row latency; 243 warps at 1)
  int ppw = 16;
  while (ppw > 1 && static_cast<long long>(B) * ((T + ppw - 1) / ppw) < 148LL * WARPS_PER_BLOCK * 2) ppw >>= 1;
  const int groups = B * ((T + ppw - 1) / ppw);
  const int blocks = (groups + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* o = static_cast<__half*>(xhat_h16);
  switch (C / 128) {
    case 1: dwconv_ln_kernel<1><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, bias, o, rstd, B, T, ppw, eps, split); break;
    case 2: dwconv_ln_kernel<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, bias, o, rstd, B, T, ppw, eps, split); break;
    case 3: dwconv_ln_kernel<3><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, bias, o, rstd, B, T, ppw, eps, split); break;
    case 4: dwconv_ln_kernel<4><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, bias, o, rstd, B, T, ppw, eps, split); break;
  }
  count_launch();
  return launch_status();
}

extern "C" int osb_layernorm(const float* x, const float* w, const float* b, float* out_f32, void* out_h16, int64_t rows, int32_t C,
                             float eps, int32_t split, void* stream) {
  OSB_REQUIRE(x && w && b && (out_f32 || out_h16), OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && (C == 128 || C == 256 || C == 384 || C == 512), OSB_ERR_SHAPE);
  const unsigned blocks = static_cast<unsigned>((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* oh = static_cast<__half*>(out_h16);
  switch (C / 128) {
    case 1: layernorm_kernel<1><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, out_f32, oh, rows, eps, split); break;
    case 2: layernorm_kernel<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, out_f32, oh, rows, eps, split); break;
    case 3: layernorm_kernel<3><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, out_f32, oh, rows, eps, split); break;
    case 4: layernorm_kernel<4><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, out_f32, oh, rows, eps, split); break;
  }
  count_launch();
  return launch_status();
}

extern "C" int osb_relu_layernorm(const float* x, const float* w, const float* b, void* out_h16, int64_t rows, int32_t C, float eps,
                                  int32_t split, const float* dot_w, const float* dot_b, const uint8_t* pad_mask, float* out_dot,
                                  void* stream) {
  OSB_REQUIRE(x && w && b && (out_h16 || out_dot), OSB_ERR_ARG);
  OSB_REQUIRE(out_dot == nullptr || dot_w != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && (C == 128 || C == 256 || C == 384 || C == 512), OSB_ERR_SHAPE);
  const unsigned blocks = static_cast<unsigned>((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __half* oh = static_cast<__half*>(out_h16);
  switch (C / 128) {
    case 1: relu_layernorm_kernel<1><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, oh, rows, eps, split, dot_w, dot_b, pad_mask, out_dot); break;
    case 2: relu_layernorm_kernel<2><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, oh, rows, eps, split, dot_w, dot_b, pad_mask, out_dot); break;
    case 3: relu_layernorm_kernel<3><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, oh, rows, eps, split, dot_w, dot_b, pad_mask, out_dot); break;
    case 4: relu_layernorm_kernel<4><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(x, w, b, oh, rows, eps, split, dot_w, dot_b, pad_mask, out_dot); break;
  }
  count_launch();
  return launch_status();
}

extern "C" int osb_variance_embed(const float* x, const float* val, const float* w, const float* bias, const uint8_t* pad_mask,
                                  const float* emb_scale, float* out_f32, void* out_h16, int32_t B, int32_t T, int32_t C, int32_t ksize, int32_t split,
                                  void* stream) {
  OSB_REQUIRE(x && val && w && bias && (out_f32 || out_h16), OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0 && ksize > 0 && ksize <= 16 && (ksize & 1), OSB_ERR_SHAPE);
  const int rows = B * T;
  variance_embed_kernel<<<(rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, val, w, bias, pad_mask, emb_scale, out_f32, static_cast<__half*>(out_h16), B, T, C, ksize, split);
  count_launch();
  return launch_status();
}

extern "C" int osb_durations(const float* log_d, const uint8_t* pad_mask, int64_t* dur, int64_t* lengths, int32_t B, int32_t T,
                             float factor, float clip_val, void* stream) {
  OSB_REQUIRE(log_d && pad_mask && dur && lengths, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0, OSB_ERR_SHAPE);
  duration_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(log_d, pad_mask, reinterpret_cast<long long*>(dur),
                                                                   reinterpret_cast<long long*>(lengths), T, factor, clip_val);
  count_launch();
  return launch_status();
}

extern "C" int osb_centres(const void* dur, int32_t dur_is_i64, float* centres, int64_t* csum, int32_t B, int32_t T, void* stream) {
  OSB_REQUIRE(dur && centres, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0, OSB_ERR_SHAPE);
  const int blocks = (B + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dur_is_i64)
    centres_kernel<long long><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(static_cast<const long long*>(dur), centres,
                                                                      reinterpret_cast<long long*>(csum), B, T);
  else
    centres_kernel<float><<<blocks, WARPS_PER_BLOCK * 32, 0, s>>>(static_cast<const float*>(dur), centres,
                                                                  reinterpret_cast<long long*>(csum), B, T);
  count_launch();
  return launch_status();
}

extern "C" int osb_gaussian_upsample(const float* hs, const float* centres, const int64_t* x_len, const int64_t* y_len, float* out_f32,
                                     void* out_h16, int32_t B, int32_t Tx, int32_t Tm, int32_t C, float delta, void* stream) {
  OSB_REQUIRE(hs && centres && x_len && y_len && (out_f32 || out_h16), OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tx > 0 && Tm > 0 && C > 0, OSB_ERR_SHAPE);
  constexpr int FR = 8;
  const size_t smem = static_cast<size_t>(FR) * Tx * sizeof(float);
  OSB_REQUIRE(smem <= 48 * 1024, OSB_ERR_SHAPE);
  dim3 grid((Tm + FR - 1) / FR, B);
  gaussian_upsample_kernel<FR><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      hs, centres, reinterpret_cast<const long long*>(x_len), reinterpret_cast<const long long*>(y_len), out_f32,
      static_cast<__half*>(out_h16), Tx, Tm, C, delta, nullptr, 0, Tm);
  count_launch();
  return launch_status();
}

extern "C" int osb_gaussian_upsample_window(const float* hs, const float* centres, const int64_t* x_len, const int64_t* y_len,
                                            const int64_t* win_start, float* out_f32, void* out_h16, int32_t B, int32_t Tx, int32_t Tm,
                                            int32_t W, int32_t halo, int32_t C, float delta, void* stream) {
  OSB_REQUIRE(hs && centres && x_len && y_len && win_start && (out_f32 || out_h16), OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tx > 0 && Tm > 0 && W > 0 && halo >= 0 && C > 0, OSB_ERR_SHAPE);
  constexpr int FR = 8;
  const size_t smem = static_cast<size_t>(FR) * Tx * sizeof(float);
  OSB_REQUIRE(smem <= 48 * 1024, OSB_ERR_SHAPE);
  dim3 grid((W + FR - 1) / FR, B);
  gaussian_upsample_kernel<FR><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      hs, centres, reinterpret_cast<const long long*>(x_len), reinterpret_cast<const long long*>(y_len), out_f32,
      static_cast<__half*>(out_h16), Tx, W, C, delta, reinterpret_cast<const long long*>(win_start), halo, Tm);
  count_launch();
  return launch_status();
}

extern "C" int osb_expand_gather(const float* x, const int64_t* csum, float* out, int32_t* index_out, int32_t B, int32_t Tx,
                                 int32_t Tm, int32_t C, void* stream) {
  OSB_REQUIRE(x && csum && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && Tx > 0 && Tm > 0 && C > 0, OSB_ERR_SHAPE);
  const int rows = B * Tm;
  expand_gather_kernel<<<(rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, WARPS_PER_BLOCK * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const long long*>(csum), out, index_out, B, Tx, Tm, C);
  count_launch();
  return launch_status();
}

extern "C" int osb_pack_h16(const float* src, int64_t src_ld, int64_t src_cs, const float* col_scale, void* dst, void* dst_lo,
                            int64_t dst_rs, int32_t dst_cols, int64_t rows, int32_t cols, void* stream) {
  OSB_REQUIRE(src && dst, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && cols > 0 && dst_cols >= cols && dst_rs >= dst_cols, OSB_ERR_SHAPE);
  const long long n = rows * dst_cols;
  const bool vec = src_cs == 1 && cols % 8 == 0 && dst_cols % 8 == 0 && src_ld % 4 == 0 && dst_rs % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 &&
                   (dst_lo == nullptr || (reinterpret_cast<uintptr_t>(dst_lo) & 15) == 0);
  if (vec)
    pack_h16_vec8_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, src_ld, col_scale, static_cast<__half*>(dst), static_cast<__half*>(dst_lo), dst_rs, dst_cols, rows, cols);
  else
    pack_h16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, src_ld, src_cs, col_scale, static_cast<__half*>(dst), static_cast<__half*>(dst_lo), dst_rs, dst_cols, rows, cols);
  count_launch();
  return launch_status();
}
