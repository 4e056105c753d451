// osb_disc.cu — the non-GEMM kernels of the multi-period discriminators (reference vocoder/wavenext/disc/_discriminators.py:41-97):
//
//   A period discriminator views the waveform (reflect-padded at the tail to a multiple of the period p) as p interleaved
//   sequences x_j[l] = wav[l*p + j] and runs (5,1)/(3,1) convolutions along l.  On this path every (signal, phase) pair is
//   one channels-last sequence and ALL sequences of a layer live in one flat fp16 matrix (NSEQ * P_i rows, C_i columns): a
//   sequence owns P_i consecutive rows, its L_i valid rows first, the rest zero.  With P_5 = P_4 = L_4 + 2 and P_i = 3 P_{i+1}
//   for the stride-3 layers, the zero tail of one sequence is the convolution padding of the next, and every layer is ONE
//   implicit GEMM over the flat matrix (osb_gemm with a row stride, LeakyReLU epilogue, gap rows re-zeroed by the keep mask)
//   with full 128-row tiles, whatever the period.  This file holds what is left:
//     mpd_first_*      layer 1 (C_in = 1): period split + reflect padding + 5-tap FIR + bias + LeakyReLU, and its backward
//     mpd_post_*       conv_post (C_out = 1): 3-tap dot product over 1024 channels, and its backward
//     lrelu_bwd        gradient gate of a LeakyReLU from the saved OUTPUT (the slope is positive: the sign survives)
//     col2im_k5        data gradient of a strided convolution: gathers the per-tap products of one GEMM (rows_out x 5*C_in)
//     l1_pair_*        feature-matching term mean|f_real - f_fake| and its gradient with respect to f_fake
//   fp16 gradient tensors carry the caller's loss scale times DISC_GRAD_SCALE (the 1/numel factors of the hinge / feature-
//   matching means would otherwise sit in the fp16 subnormal range); the kernels that produce fp32 results take the factor
//   back out (`inv_scale`).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ void st_h8(__half* dst, const float (&v)[8]) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void ld_h8(const __half* src, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(src);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// sample index of element l of phase j of the reflect-padded signal (F.pad(x, (0, n_pad), "reflect"): padded[T + i] = x[T - 2 - i])
__device__ __forceinline__ long long wav_index(int l, int j, int p, int T) {
  const long long w = static_cast<long long>(l) * p + j;
  return w < T ? w : 2LL * T - 2 - w;
}

// ------------------------------------------------------------------------------------------
// layer 1: out[(n*p + j), lo, c] = lrelu(b[c] + sum_k w[c,k] * x_j[3*lo + k - 2]),  c < 32 (columns 32..CP-1 are zero padding
// so that the next layer's contraction is a whole 64-element k-block).  One thread = 8 channels of one output row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mpd_first_fwd_kernel(const float* __restrict__ wav, const float* __restrict__ w /*(32,5)*/, const float* __restrict__ bias,
                     __half* __restrict__ out, int NS, int T, int p, int L0, int L1, int P1, int CP, int stride, float slope) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = CP / 8;
  const long long rows = static_cast<long long>(NS) * p * P1;
  if (i >= rows * groups) return;
  const int cg = static_cast<int>(i % groups);
  const long long row = i / groups;
  const int lo = static_cast<int>(row % P1);
  const int seq = static_cast<int>(row / P1);
  const int n = seq / p, j = seq % p;
  float v[8];
  if (cg * 8 >= 32 || lo >= L1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
  } else {
    float x[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int l = stride * lo + k - 2;
      x[k] = (l >= 0 && l < L0) ? wav[static_cast<long long>(n) * T + wav_index(l, j, p, T)] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = cg * 8 + q;
      float acc = bias[c];
#pragma unroll
      for (int k = 0; k < 5; ++k) acc = fmaf(w[c * 5 + k], x[k], acc);
      v[q] = lrelu(acc, slope);
    }
  }
  st_h8(out + row * CP + cg * 8, v);
}

// d wav[n, widx] += inv_scale * sum_{k, lo: 3 lo + k - 2 = l} sum_c w[c,k] * g[(n*p+j), lo, c]     (g already gated)
// one thread per (seq, l); at most two (k, lo) pairs hit an input position; reflected positions fold onto T-2-i: atomics.
__global__ void __launch_bounds__(256)
mpd_first_dx_kernel(const __half* __restrict__ g, const float* __restrict__ w, float* __restrict__ dwav, int NS, int T, int p,
                    int L0, int L1, int P1, int CP, int stride, float inv_scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(NS) * p * L0;
  if (i >= total) return;
  const int l = static_cast<int>(i % L0);
  const int seq = static_cast<int>(i / L0);
  const int n = seq / p, j = seq % p;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int num = l + 2 - k;
    if (num < 0 || num % stride != 0) continue;
    const int lo = num / stride;
    if (lo >= L1) continue;
    const __half* gr = g + (static_cast<long long>(seq) * P1 + lo) * CP;
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 8) {
      float gv[8];
      ld_h8(gr + c0, gv);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(w[(c0 + q) * 5 + k], gv[q], acc);
    }
  }
  if (acc != 0.f) atomicAdd(dwav + static_cast<long long>(n) * T + wav_index(l, j, p, T), acc * inv_scale);
}

// dW1[c,k] += inv_scale * sum g[row, c] * x_j[3 lo + k - 2] ; db1[c] += inv_scale * sum g[row, c].  A thread owns 8 channels of
// a row (one 16-byte load; 4 threads cover the 32 real channels, a warp 8 rows per trip); partial sums are folded across the
// warp by shuffles, across the block in shared memory, and leave with one atomic per element and block.
__global__ void __launch_bounds__(256)
mpd_first_dw_kernel(const __half* __restrict__ g, const float* __restrict__ wav, float* __restrict__ dw /*(32,5)*/, float* __restrict__ db,
                    int NS, int T, int p, int L0, int L1, int P1, int CP, int stride, float inv_scale, long long rows_per_block) {
  __shared__ float part[8][32 * 6];
  const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
  const int cg = lane & 3, rsub = lane >> 2;             // channel group (8 channels), row within the warp's 8-row trip
  const long long rows = static_cast<long long>(NS) * p * P1;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float aw[8][5], ab[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    ab[q] = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) aw[q][k] = 0.f;
  }
  for (long long row = r0 + wip * 8 + rsub; row < r1; row += 64) {
    const int lo = static_cast<int>(row % P1);
    if (lo >= L1) continue;
    const int seq = static_cast<int>(row / P1);
    const int n = seq / p, j = seq % p;
    float gv[8], x[5];
    ld_h8(g + row * CP + cg * 8, gv);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int l = stride * lo + k - 2;
      x[k] = (l >= 0 && l < L0) ? wav[static_cast<long long>(n) * T + wav_index(l, j, p, T)] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      ab[q] += gv[q];
#pragma unroll
      for (int k = 0; k < 5; ++k) aw[q][k] = fmaf(gv[q], x[k], aw[q][k]);
    }
  }
  // fold the 8 row slots of the warp (lanes with equal cg), then the 8 warps
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      float v = k < 5 ? aw[q][k] : ab[q];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (rsub == 0) part[wip][(cg * 8 + q) * 6 + k] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32 * 6) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += part[q][threadIdx.x];
    const int c = threadIdx.x / 6, k = threadIdx.x % 6;
    if (k < 5) atomicAdd(dw + c * 5 + k, s * inv_scale);
    else atomicAdd(db + c, s * inv_scale);
  }
}

// ------------------------------------------------------------------------------------------
// conv_post: score[n, l*p + j] = b + sum_k sum_c w[c,k] * x[seq*P + l + k - 1, c]   (seq = n*p + j, C = 1024, k = 3; the rows
// before / after a sequence are zero or out of range).  One warp per output.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mpd_post_fwd_kernel(const __half* __restrict__ x, const float* __restrict__ w /*(C,3)*/, const float* __restrict__ bias, float* __restrict__ out,
                    int NSEQ, int p, int L, int P, int C) {
  const long long o = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (o >= static_cast<long long>(NSEQ) * L) return;
  const int lane = threadIdx.x & 31;
  const int l = static_cast<int>(o % L);
  const long long seq = o / L;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int li = l + k - 1;
    if (li < 0 || li >= L) continue;
    const __half* xr = x + (seq * P + li) * C;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float xv[8];
      ld_h8(xr + c0, xv);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc = fmaf(w[(c0 + q) * 3 + k], xv[q], acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(seq / p) * (static_cast<long long>(L) * p) + static_cast<long long>(l) * p + (seq % p)] = acc + bias[0];
}

__device__ __forceinline__ float score_grad(const float* dout, long long seq, int l, int p, int L) {
  return dout[(seq / p) * (static_cast<long long>(L) * p) + static_cast<long long>(l) * p + (seq % p)];
}

// dx[seq*P + l, c] = scale * sum_k w[c,k] * dscore[seq, l - k + 1]   (fp16, zero on the gap rows);  one thread = 8 channels of one row
__global__ void __launch_bounds__(256)
mpd_post_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w, __half* __restrict__ dx, int NSEQ, int p, int L, int P, int C,
                   float scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = C / 8;
  if (i >= static_cast<long long>(NSEQ) * P * groups) return;
  const int cg = static_cast<int>(i % groups);
  const long long row = i / groups;
  const int l = static_cast<int>(row % P);
  const long long seq = row / P;
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = 0.f;
  if (l < L) {
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int lo = l - k + 1;
      d[k] = (lo >= 0 && lo < L) ? score_grad(dout, seq, lo, p, L) * scale : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = cg * 8 + q;
      v[q] = w[c * 3 + 0] * d[0] + w[c * 3 + 1] * d[1] + w[c * 3 + 2] * d[2];
    }
  }
  st_h8(dx + row * C + cg * 8, v);
}

// dw[c,k] += sum dscore[seq, l] * x[seq*P + l + k - 1, c] ; db += sum dscore.   Thread = 8 channels; a block owns a row range.
__global__ void __launch_bounds__(128)
mpd_post_dw_kernel(const float* __restrict__ dout, const __half* __restrict__ x, float* __restrict__ dw, float* __restrict__ db, int NSEQ, int p,
                   int L, int P, int C, long long rows_per_block) {
  const int cg = threadIdx.x;             // C / 8 == blockDim.x == 128
  const long long rows = static_cast<long long>(NSEQ) * P;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float aw[8][3];
#pragma unroll
  for (int q = 0; q < 8; ++q) aw[q][0] = aw[q][1] = aw[q][2] = 0.f;
  float ab = 0.f;
  for (long long row = r0; row < r1; ++row) {   // row = the INPUT row (seq, li); it meets dscore at l = li - k + 1
    const int li = static_cast<int>(row % P);
    if (li >= L) continue;
    const long long seq = row / P;
    float xv[8];
    ld_h8(x + row * C + cg * 8, xv);
    ab += score_grad(dout, seq, li, p, L);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = li - k + 1;
      const float d = (l >= 0 && l < L) ? score_grad(dout, seq, l, p, L) : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) aw[q][k] = fmaf(d, xv[q], aw[q][k]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(dw + (cg * 8 + q) * 3 + k, aw[q][k]);
  if (cg == 0) atomicAdd(db, ab);
}

// g = dy * (y > 0 ? 1 : slope) on the valid rows (row % P < L), 0 on the gap rows      (fp16, 8 elements per thread)
// A thread owns 8 channels and walks LB_ROWS rows; with `colsum` it also accumulates the column sums of g (the bias gradient of
// the layer, scaled) and leaves with 8 atomics — the separate column-sum pass re-read every gradient row from HBM.
constexpr int LB_ROWS = 8;
__global__ void __launch_bounds__(256)
lrelu_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ y, __half* __restrict__ g, long long rows, int C8, int P, int L,
                 float slope, float* __restrict__ colsum, float colsum_scale) {
  // column sums: threads of a block that own the same channel group meet in shared memory first (256 % C8 == 0: the group of
  // a thread is threadIdx.x % C8), so every address sees one global atomic per block instead of one per 8-row slab
  __shared__ float sacc[256 * 8];
  const bool block_reduce = colsum != nullptr && C8 <= 256 && 256 % C8 == 0;
  if (block_reduce) {
    for (int k = threadIdx.x; k < C8 * 8; k += blockDim.x) sacc[k] = 0.f;
    __syncthreads();
  }
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long slabs = (rows + LB_ROWS - 1) / LB_ROWS;
  const bool active = i < slabs * C8;
  const int cg = active ? static_cast<int>(i % C8) : 0;
  const long long r0 = active ? (i / C8) * LB_ROWS : rows;
  uint4 a4[LB_ROWS], b4[LB_ROWS];
  bool live[LB_ROWS];
#pragma unroll
  for (int j = 0; j < LB_ROWS; ++j) {     // all loads of the slab in flight before the first use
    const long long r = r0 + j;
    live[j] = r < rows && static_cast<int>(r % P) < L;
    if (live[j]) {
      const long long off = (r * C8 + cg) * 8;
      a4[j] = *reinterpret_cast<const uint4*>(dy + off);
      b4[j] = *reinterpret_cast<const uint4*>(y + off);
    }
  }
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
  for (int j = 0; j < LB_ROWS; ++j) {
    const long long r = r0 + j;
    if (r >= rows) break;
    float a[8];
    if (live[j]) {
      const __half2* ah = reinterpret_cast<const __half2*>(&a4[j]);
      const __half2* bh = reinterpret_cast<const __half2*>(&b4[j]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 av = __half22float2(ah[q]), bv = __half22float2(bh[q]);
        a[2 * q] = bv.x > 0.f ? av.x : av.x * slope;
        a[2 * q + 1] = bv.y > 0.f ? av.y : av.y * slope;
        acc[2 * q] += a[2 * q];
        acc[2 * q + 1] += a[2 * q + 1];
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = 0.f;
    }
    st_h8(g + (r * C8 + cg) * 8, a);
  }
  if (block_reduce) {
    if (active) {
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(&sacc[cg * 8 + q], acc[q]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C8 * 8; k += blockDim.x) atomicAdd(colsum + k, sacc[k] * colsum_scale);
  } else if (colsum != nullptr && active) {
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(colsum + cg * 8 + q, acc[q] * colsum_scale);
  }
}

// data gradient of a strided convolution from the per-tap products of ONE GEMM: col (rows_out, taps*C) holds, at column block j,
// sum_n g[r, n] W[n, :, tap(j)] with tap(j) = j or taps-1-j (`reversed`: the tap-reversed weight pack); the input row
// r_in = stride*r + tap - pad collects it:   dx[r_in, c] = sum_{tap: (r_in + pad - tap) % stride == 0} col[(r_in + pad - tap)/stride, j(tap)*C + c]
__global__ void __launch_bounds__(256)
col2im_kernel(const __half* __restrict__ col, __half* __restrict__ dx, long long rows_in, long long rows_out, int C, int taps, int pad, int stride,
              int reversed) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = C / 8;
  if (i >= rows_in * groups) return;
  const int cg = static_cast<int>(i % groups);
  const long long row = i / groups;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int k = 0; k < taps; ++k) {
    const long long num = row + pad - k;
    if (num < 0 || num % stride != 0) continue;
    const long long ro = num / stride;
    if (ro >= rows_out) continue;
    float v[8];
    ld_h8(col + (ro * taps + (reversed ? taps - 1 - k : k)) * C + cg * 8, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += v[q];
  }
  st_h8(dx + row * C + cg * 8, acc);
}

// sum |a - b| over n8*8 fp16 elements -> out[0] (atomic, one per block)
__global__ void __launch_bounds__(256)
l1_pair_fwd_kernel(const __half* __restrict__ a, const __half* __restrict__ b, float* __restrict__ out, long long n8) {
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    ld_h8(a + i * 8, x);
    ld_h8(b + i * 8, y);
#pragma unroll
    for (int q = 0; q < 8; ++q) s += fabsf(x[q] - y[q]);
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q];
    atomicAdd(out, t);
  }
}

// d b = coef[0] * scale * sign(b - a)   (gradient of coef * sum|a - b| with respect to b), fp16
__global__ void __launch_bounds__(256)
l1_pair_bwd_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const float* __restrict__ coef, float scale, __half* __restrict__ db,
                   long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float c = coef[0] * scale;
  float x[8], y[8];
  ld_h8(a + i * 8, x);
  ld_h8(b + i * 8, y);
#pragma unroll
  for (int q = 0; q < 8; ++q) x[q] = y[q] > x[q] ? c : (y[q] < x[q] ? -c : 0.f);
  st_h8(db + i * 8, x);
}

// ------------------------------------------------------------------------------------------
// Resolution discriminators (reference _discriminators.py:139-216): Conv2d stacks over magnitude spectrograms (NS, F, W).
// Same flat layout, with the (signal, frame) pairs as the sequences and the frequency axis along the rows:
//   row = (n * W_i + w) * P_i + h,  h < H_i valid, 64 channels per row;  P_5 = H_5 + 2, P_i = 2 P_(i+1) (every layer has
//   stride 2 along frequency).  The kw taps along the frame axis are folded into the contraction by a small gather
//   (wim2col: K = kw * 64), the kh taps along frequency are the implicit-GEMM taps with row stride 2.
// ------------------------------------------------------------------------------------------
constexpr int R1_KH = 7, R1_KW = 5, R1_TAPS = R1_KH * R1_KW;

// layer 1 (Conv2d(1, 64, (7,5), (2,2), (3,2)) + LeakyReLU): one thread = 8 channels of one output row
__global__ void __launch_bounds__(256)
mrd_first_fwd_kernel(const float* __restrict__ spec, const float* __restrict__ w /*(64,35)*/, const float* __restrict__ bias,
                     __half* __restrict__ out, int NS, int F, int W, int H1, int W1, int P1, float slope) {
  __shared__ float sw[R1_TAPS][64];   // tap-major
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * R1_TAPS; i += blockDim.x) sw[i % R1_TAPS][i / R1_TAPS] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(NS) * W1 * P1;
  if (i >= rows * 8) return;
  const int cg = static_cast<int>(i & 7);
  const long long row = i >> 3;
  const int h = static_cast<int>(row % P1);
  const int w1 = static_cast<int>((row / P1) % W1);
  const int n = static_cast<int>(row / (static_cast<long long>(P1) * W1));
  float v[8];
  if (h >= H1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = sb[cg * 8 + q];
    const float* sp = spec + static_cast<long long>(n) * F * W;
    for (int kh = 0; kh < R1_KH; ++kh) {
      const int f = 2 * h + kh - 3;
      if (f < 0 || f >= F) continue;
#pragma unroll
      for (int kw = 0; kw < R1_KW; ++kw) {
        const int t = 2 * w1 + kw - 2;
        if (t < 0 || t >= W) continue;
        const float x = sp[static_cast<long long>(f) * W + t];
        const float* wt = &sw[kh * R1_KW + kw][cg * 8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaf(wt[q], x, v[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = lrelu(v[q], slope);
  }
  st_h8(out + row * 64 + cg * 8, v);
}

// d spec[n, f, t] = inv_scale * sum_{kh, kw: parity ok} sum_c w[c, kh, kw] * g[(n, w1, h1), c]   (g gated); one thread per input
__global__ void __launch_bounds__(256)
mrd_first_dx_kernel(const __half* __restrict__ g, const float* __restrict__ w, float* __restrict__ dspec, int NS, int F, int W, int H1,
                    int W1, int P1, float inv_scale) {
  __shared__ float sw[R1_TAPS][64];
  for (int i = threadIdx.x; i < 64 * R1_TAPS; i += blockDim.x) sw[i % R1_TAPS][i / R1_TAPS] = w[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(NS) * F * W) return;
  const int t = static_cast<int>(i % W);
  const int f = static_cast<int>((i / W) % F);
  const int n = static_cast<int>(i / (static_cast<long long>(W) * F));
  float acc = 0.f;
  for (int kh = 0; kh < R1_KH; ++kh) {
    const int nh = f + 3 - kh;
    if (nh < 0 || (nh & 1)) continue;
    const int h1 = nh >> 1;
    if (h1 >= H1) continue;
    for (int kw = 0; kw < R1_KW; ++kw) {
      const int nw = t + 2 - kw;
      if (nw < 0 || (nw & 1)) continue;
      const int w1 = nw >> 1;
      if (w1 >= W1) continue;
      const __half* gr = g + ((static_cast<long long>(n) * W1 + w1) * P1 + h1) * 64;
      const float* wt = sw[kh * R1_KW + kw];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 8) {
        float gv[8];
        ld_h8(gr + c0, gv);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = fmaf(wt[c0 + q], gv[q], acc);
      }
    }
  }
  dspec[i] = acc * inv_scale;
}

// dW1[c, tap] += inv_scale * sum g[row, c] * spec[...];  db1[c] += inv_scale * sum g[row, c].  A lane owns 2 channels, a warp walks
// rows (the 35 spectrogram values of a row are warp-uniform loads); block partials in shared memory, one atomic per element.
__global__ void __launch_bounds__(256)
mrd_first_dw_kernel(const __half* __restrict__ g, const float* __restrict__ spec, float* __restrict__ dw /*(64,35)*/, float* __restrict__ db,
                    int NS, int F, int W, int H1, int W1, int P1, float inv_scale, long long rows_per_block) {
  __shared__ float part[64 * (R1_TAPS + 1)];
  for (int i = threadIdx.x; i < 64 * (R1_TAPS + 1); i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wip = threadIdx.x >> 5;
  const long long rows = static_cast<long long>(NS) * W1 * P1;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float a0[R1_TAPS + 1], a1[R1_TAPS + 1];
#pragma unroll
  for (int k = 0; k <= R1_TAPS; ++k) a0[k] = a1[k] = 0.f;
  for (long long row = r0 + wip; row < r1; row += 8) {
    const int h = static_cast<int>(row % P1);
    if (h >= H1) continue;
    const int w1 = static_cast<int>((row / P1) % W1);
    const int n = static_cast<int>(row / (static_cast<long long>(P1) * W1));
    const float2 gv = __half22float2(*reinterpret_cast<const __half2*>(g + row * 64 + lane * 2));
    const float* sp = spec + static_cast<long long>(n) * F * W;
    a0[R1_TAPS] += gv.x;
    a1[R1_TAPS] += gv.y;
#pragma unroll
    for (int kh = 0; kh < R1_KH; ++kh) {
      const int f = 2 * h + kh - 3;
#pragma unroll
      for (int kw = 0; kw < R1_KW; ++kw) {
        const int t = 2 * w1 + kw - 2;
        const float x = (f >= 0 && f < F && t >= 0 && t < W) ? sp[static_cast<long long>(f) * W + t] : 0.f;
        a0[kh * R1_KW + kw] = fmaf(gv.x, x, a0[kh * R1_KW + kw]);
        a1[kh * R1_KW + kw] = fmaf(gv.y, x, a1[kh * R1_KW + kw]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k <= R1_TAPS; ++k) {
    atomicAdd(&part[(lane * 2) * (R1_TAPS + 1) + k], a0[k]);
    atomicAdd(&part[(lane * 2 + 1) * (R1_TAPS + 1) + k], a1[k]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * (R1_TAPS + 1); i += blockDim.x) {
    const int c = i / (R1_TAPS + 1), k = i % (R1_TAPS + 1);
    if (k < R1_TAPS) atomicAdd(dw + c * R1_TAPS + k, part[i] * inv_scale);
    else atomicAdd(db + c, part[i] * inv_scale);
  }
}

// xcol[(n, wo, h), kw*C + c] = x[(n, wo*sw + kw - pw, h), c]  (zero outside [0, W_in)); 8 channels per thread
__global__ void __launch_bounds__(256)
wim2col_kernel(const __half* __restrict__ x, __half* __restrict__ xcol, int NS, int W_in, int W_out, int P, int C, int KW, int pw, int sw) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = KW * C / 8;
  const long long rows = static_cast<long long>(NS) * W_out * P;
  if (i >= rows * groups) return;
  const int gidx = static_cast<int>(i % groups);
  const long long row = i / groups;
  const int kw = gidx / (C / 8), cg = gidx % (C / 8);
  const int h = static_cast<int>(row % P);
  const int wo = static_cast<int>((row / P) % W_out);
  const long long n = row / (static_cast<long long>(P) * W_out);
  const int wi = wo * sw + kw - pw;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (wi >= 0 && wi < W_in) v = *reinterpret_cast<const uint4*>(x + ((n * W_in + wi) * P + h) * C + cg * 8);
  *reinterpret_cast<uint4*>(xcol + row * (KW * C) + kw * C + cg * 8) = v;
}

// dx[(n, w, h), c] = sum_{kw: (w + pw - kw) % sw == 0, wo in range} dxcol[(n, wo, h), kw*C + c]
__global__ void __launch_bounds__(256)
wcol2im_kernel(const __half* __restrict__ dxcol, __half* __restrict__ dx, int NS, int W_in, int W_out, int P, int C, int KW, int pw, int sw) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int groups = C / 8;
  const long long rows = static_cast<long long>(NS) * W_in * P;
  if (i >= rows * groups) return;
  const int cg = static_cast<int>(i % groups);
  const long long row = i / groups;
  const int h = static_cast<int>(row % P);
  const int wi = static_cast<int>((row / P) % W_in);
  const long long n = row / (static_cast<long long>(P) * W_in);
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int kw = 0; kw < KW; ++kw) {
    const int num = wi + pw - kw;
    if (num < 0 || num % sw != 0) continue;
    const int wo = num / sw;
    if (wo >= W_out) continue;
    float v[8];
    ld_h8(dxcol + ((n * W_out + wo) * P + h) * (KW * C) + kw * C + cg * 8, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += v[q];
  }
  st_h8(dx + row * C + cg * 8, acc);
}

// conv_post (Conv2d(64, 1, (3,3), padding 1)): score[n, h*W + w] = b + sum_{kh,kw,c} w[c,kh,kw] x[(n, w+kw-1, h+kh-1), c]; 8 lanes per output
__global__ void __launch_bounds__(256)
mrd_post_fwd_kernel(const __half* __restrict__ x, const float* __restrict__ w /*(64,9)*/, const float* __restrict__ bias, float* __restrict__ out,
                    int NS, int W, int H, int P) {
  __shared__ float sw9[9][64];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) sw9[i % 9][i / 9] = w[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long o = i >> 3;
  const int cg = static_cast<int>(i & 7);
  const bool live = o < static_cast<long long>(NS) * W * H;
  float acc = 0.f;
  int h = 0, wv = 0;
  long long n = 0;
  if (live) {
    h = static_cast<int>(o % H);
    wv = static_cast<int>((o / H) % W);
    n = o / (static_cast<long long>(H) * W);
    for (int kw = 0; kw < 3; ++kw) {
      const int wi = wv + kw - 1;
      if (wi < 0 || wi >= W) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = h + kh - 1;
        if (hi < 0 || hi >= H) continue;
        float xv[8];
        ld_h8(x + ((n * W + wi) * P + hi) * 64 + cg * 8, xv);
        const float* wt = &sw9[kh * 3 + kw][cg * 8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc = fmaf(wt[q], xv[q], acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && cg == 0) out[n * (static_cast<long long>(H) * W) + static_cast<long long>(h) * W + wv] = acc + bias[0];
}

// dx[(n, w, h), c] = scale * sum_{kh,kw} w[c,kh,kw] dscore[n, (h-kh+1)*W + (w-kw+1)]   (fp16; zero on the gap rows)
__global__ void __launch_bounds__(256)
mrd_post_dx_kernel(const float* __restrict__ dscore, const float* __restrict__ w, __half* __restrict__ dx, int NS, int W, int H, int P, float scale) {
  __shared__ float sw9[9][64];
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) sw9[i % 9][i / 9] = w[i];
  __syncthreads();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(NS) * W * P;
  if (i >= rows * 8) return;
  const int cg = static_cast<int>(i & 7);
  const long long row = i >> 3;
  const int h = static_cast<int>(row % P);
  const int wv = static_cast<int>((row / P) % W);
  const long long n = row / (static_cast<long long>(P) * W);
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = 0.f;
  if (h < H) {
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = h - kh + 1;
      if (ho < 0 || ho >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int wo = wv - kw + 1;
        if (wo < 0 || wo >= W) continue;
        const float d = dscore[n * (static_cast<long long>(H) * W) + static_cast<long long>(ho) * W + wo] * scale;
        const float* wt = &sw9[kh * 3 + kw][cg * 8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaf(wt[q], d, v[q]);
      }
    }
  }
  st_h8(dx + row * 64 + cg * 8, v);
}

// dw[c, kh, kw] += sum dscore[n, ho, wo] x[(n, wo+kw-1, ho+kh-1), c];  db += sum dscore.  Thread = 8 channels of an INPUT row.
__global__ void __launch_bounds__(256)
mrd_post_dw_kernel(const float* __restrict__ dscore, const __half* __restrict__ x, float* __restrict__ dw /*(64,9)*/, float* __restrict__ db,
                   int NS, int W, int H, int P, long long rows_per_block) {
  __shared__ float part[64 * 9 + 1];
  for (int i = threadIdx.x; i < 64 * 9 + 1; i += blockDim.x) part[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x & 7, rsub = threadIdx.x >> 3;   // 32 rows per trip
  const long long rows = static_cast<long long>(NS) * W * P;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float aw[8][9];
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int k = 0; k < 9; ++k) aw[q][k] = 0.f;
  float ab = 0.f;
  for (long long row = r0 + rsub; row < r1; row += 32) {
    const int hi = static_cast<int>(row % P);
    if (hi >= H) continue;
    const int wi = static_cast<int>((row / P) % W);
    const long long n = row / (static_cast<long long>(P) * W);
    float xv[8];
    ld_h8(x + row * 64 + cg * 8, xv);
    const float* ds = dscore + n * (static_cast<long long>(H) * W);
    if (cg == 0) ab += ds[static_cast<long long>(hi) * W + wi];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = hi - kh + 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wo = wi - kw + 1;
        const float d = (ho >= 0 && ho < H && wo >= 0 && wo < W) ? ds[static_cast<long long>(ho) * W + wo] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) aw[q][kh * 3 + kw] = fmaf(d, xv[q], aw[q][kh * 3 + kw]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int k = 0; k < 9; ++k) atomicAdd(&part[(cg * 8 + q) * 9 + k], aw[q][k]);
  if (cg == 0) atomicAdd(&part[64 * 9], ab);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) atomicAdd(dw + i, part[i]);
  if (threadIdx.x == 0) atomicAdd(db, part[64 * 9]);
}

// Layer 1 of a resolution discriminator as a GEMM: xcol[(n, w1, h1), tap] = spec[n, 2 h1 + kh - 3, 2 w1 + kw - 2], tap = kh*5 + kw < 35,
// zero for taps 35..63, outside the spectrogram and on the gap rows (h1 >= H1).  8 columns per thread.
__global__ void __launch_bounds__(256)
spec_im2col_kernel(const float* __restrict__ spec, __half* __restrict__ xcol, int NS, int F, int W, int H1, int W1, int P1) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = static_cast<long long>(NS) * W1 * P1;
  if (i >= rows * 8) return;
  const int cg = static_cast<int>(i & 7);
  const long long row = i >> 3;
  const int h = static_cast<int>(row % P1);
  const int w1 = static_cast<int>((row / P1) % W1);
  const int n = static_cast<int>(row / (static_cast<long long>(P1) * W1));
  float v[8];
  const float* sp = spec + static_cast<long long>(n) * F * W;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int tap = cg * 8 + q;
    const int kh = tap / R1_KW, kw = tap % R1_KW;
    const int f = 2 * h + kh - 3, t = 2 * w1 + kw - 2;
    v[q] = (tap < R1_TAPS && h < H1 && f >= 0 && f < F && t >= 0 && t < W) ? sp[static_cast<long long>(f) * W + t] : 0.f;
  }
  st_h8(xcol + row * 64 + cg * 8, v);
}

// adjoint: dspec[n, f, t] = inv_scale * sum_{kh, kw with matching parity} col[(n, w1, h1), kh*5 + kw]   (col = g . W1, fp16)
__global__ void __launch_bounds__(256)
spec_col2im_kernel(const __half* __restrict__ col, float* __restrict__ dspec, int NS, int F, int W, int H1, int W1, int P1, float inv_scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(NS) * F * W) return;
  const int t = static_cast<int>(i % W);
  const int f = static_cast<int>((i / W) % F);
  const int n = static_cast<int>(i / (static_cast<long long>(W) * F));
  float acc = 0.f;
  for (int kh = (f + 3) & 1; kh < R1_KH; kh += 2) {
    const int h1 = (f + 3 - kh) >> 1;
    if (f + 3 - kh < 0 || h1 >= H1) continue;
    for (int kw = (t + 2) & 1; kw < R1_KW; kw += 2) {
      const int w1 = (t + 2 - kw) >> 1;
      if (t + 2 - kw < 0 || w1 >= W1) continue;
      acc += __half2float(col[((static_cast<long long>(n) * W1 + w1) * P1 + h1) * 64 + kh * R1_KW + kw]);
    }
  }
  dspec[i] = acc * inv_scale;
}

inline unsigned grid_for(long long n, int block) { return static_cast<unsigned>((n + block - 1) / block); }

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_mpd_first_fwd(const float* wav, const float* w, const float* bias, void* out_h16, int32_t NS, int32_t T, int32_t period,
                                 int32_t L1, int32_t P1, int32_t CP, int32_t stride, float slope, void* stream) {
  OSB_REQUIRE(wav && w && bias && out_h16, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && T > 1 && period > 0 && L1 > 0 && P1 >= L1 && CP >= 32 && CP % 8 == 0 && stride >= 1, OSB_ERR_SHAPE);
  const int L0 = (T + period - 1) / period;
  OSB_REQUIRE(L0 * period - T < T - 1, OSB_ERR_SHAPE);   // reflect padding needs n_pad < T
  const long long n = static_cast<long long>(NS) * period * P1 * (CP / 8);
  mpd_first_fwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(wav, w, bias, static_cast<__half*>(out_h16), NS, T, period, L0,
                                                                                       L1, P1, CP, stride, slope);
  count_launch();
  return launch_status();
}

extern "C" int osb_mpd_first_bwd(const void* g_h16, const float* wav, const float* w, float* dwav, float* dw, float* db, int32_t NS, int32_t T,
                                 int32_t period, int32_t L1, int32_t P1, int32_t CP, int32_t stride, float inv_scale, void* stream) {
  OSB_REQUIRE(g_h16 && wav && w, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && T > 1 && period > 0 && L1 > 0 && P1 >= L1, OSB_ERR_SHAPE);
  const int L0 = (T + period - 1) / period;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* g = static_cast<const __half*>(g_h16);
  int launched = 0;
  if (dwav != nullptr) {   // accumulated (+=): the caller zeroes it
    const long long n = static_cast<long long>(NS) * period * L0;
    mpd_first_dx_kernel<<<grid_for(n, 256), 256, 0, s>>>(g, w, dwav, NS, T, period, L0, L1, P1, CP, stride, inv_scale);
    ++launched;
  }
  if (dw != nullptr && db != nullptr) {
    const long long rows = static_cast<long long>(NS) * period * P1;
    long long blocks = (rows + 2047) / 2048;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const long long rpb = (rows + blocks - 1) / blocks;
    mpd_first_dw_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(g, wav, dw, db, NS, T, period, L0, L1, P1, CP, stride, inv_scale, rpb);
    ++launched;
  }
  count_launch(launched);
  return launch_status();
}

extern "C" int osb_mpd_post_fwd(const void* x_h16, const float* w, const float* bias, float* out, int32_t NSEQ, int32_t period, int32_t L,
                                int32_t P, int32_t C, void* stream) {
  OSB_REQUIRE(x_h16 && w && bias && out, OSB_ERR_ARG);
  OSB_REQUIRE(NSEQ > 0 && period > 0 && NSEQ % period == 0 && L > 0 && P >= L && C % 256 == 0, OSB_ERR_SHAPE);
  const long long rows = static_cast<long long>(NSEQ) * L;
  mpd_post_fwd_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x_h16), w, bias, out, NSEQ, period,
                                                                                       L, P, C);
  count_launch();
  return launch_status();
}

extern "C" int osb_mpd_post_bwd(const float* dout, const void* x_h16, const float* w, void* dx_h16, float* dw, float* db, int32_t NSEQ,
                                int32_t period, int32_t L, int32_t P, int32_t C, float scale, void* stream) {
  OSB_REQUIRE(dout && w, OSB_ERR_ARG);
  OSB_REQUIRE(NSEQ > 0 && period > 0 && NSEQ % period == 0 && L > 0 && P >= L && C == 1024, OSB_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int launched = 0;
  if (dx_h16 != nullptr) {
    const long long n = static_cast<long long>(NSEQ) * P * (C / 8);
    mpd_post_dx_kernel<<<grid_for(n, 256), 256, 0, s>>>(dout, w, static_cast<__half*>(dx_h16), NSEQ, period, L, P, C, scale);
    ++launched;
  }
  if (dw != nullptr && db != nullptr && x_h16 != nullptr) {   // accumulated (+=): the caller zeroes them
    const long long rows = static_cast<long long>(NSEQ) * P;
    long long blocks = (rows + 63) / 64;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const long long rpb = (rows + blocks - 1) / blocks;
    mpd_post_dw_kernel<<<static_cast<unsigned>(blocks), 128, 0, s>>>(dout, static_cast<const __half*>(x_h16), dw, db, NSEQ, period, L, P, C, rpb);
    ++launched;
  }
  count_launch(launched);
  return launch_status();
}

extern "C" int osb_lrelu_bwd_h16(const void* dy, const void* y, void* g, int64_t rows, int32_t C, int32_t P, int32_t L, float slope,
                                 float* colsum, float colsum_scale, void* stream) {
  OSB_REQUIRE(dy && y && g, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && P > 0 && L > 0 && L <= P, OSB_ERR_SHAPE);
  const long long n = ((rows + LB_ROWS - 1) / LB_ROWS) * (C / 8);
  lrelu_bwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(dy), static_cast<const __half*>(y),
                                                                                  static_cast<__half*>(g), rows, C / 8, P, L, slope, colsum,
                                                                                  colsum_scale);
  count_launch();
  return launch_status();
}

extern "C" int osb_col2im_h16(const void* col, void* dx, int64_t rows_in, int64_t rows_out, int32_t C, int32_t taps, int32_t pad,
                              int32_t stride, int32_t reversed, void* stream) {
  OSB_REQUIRE(col && dx, OSB_ERR_ARG);
  OSB_REQUIRE(rows_in > 0 && rows_out > 0 && C % 8 == 0 && taps > 0 && stride >= 1, OSB_ERR_SHAPE);
  const long long n = rows_in * (C / 8);
  col2im_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(col), static_cast<__half*>(dx), rows_in,
                                                                              rows_out, C, taps, pad, stride, reversed);
  count_launch();
  return launch_status();
}

extern "C" int osb_l1_pair_fwd(const void* a_h16, const void* b_h16, float* out_sum, int64_t n, void* stream) {
  OSB_REQUIRE(a_h16 && b_h16 && out_sum, OSB_ERR_ARG);
  OSB_REQUIRE(n > 0 && n % 8 == 0, OSB_ERR_SHAPE);
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  l1_pair_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(a_h16),
                                                                                                static_cast<const __half*>(b_h16), out_sum, n / 8);
  count_launch();
  return launch_status();
}

extern "C" int osb_l1_pair_bwd(const void* a_h16, const void* b_h16, const float* coef, float scale, void* db_h16, int64_t n, void* stream) {
  OSB_REQUIRE(a_h16 && b_h16 && coef && db_h16, OSB_ERR_ARG);
  OSB_REQUIRE(n > 0 && n % 8 == 0, OSB_ERR_SHAPE);
  l1_pair_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(a_h16), static_cast<const __half*>(b_h16),
                                                                                      coef, scale, static_cast<__half*>(db_h16), n / 8);
  count_launch();
  return launch_status();
}

extern "C" int osb_mrd_first_fwd(const float* spec, const float* w, const float* bias, void* out_h16, int32_t NS, int32_t F, int32_t W,
                                 int32_t H1, int32_t W1, int32_t P1, float slope, void* stream) {
  OSB_REQUIRE(spec && w && bias && out_h16, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && F > 0 && W > 0 && H1 > 0 && W1 > 0 && P1 >= H1, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * W1 * P1 * 8;
  mrd_first_fwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(spec, w, bias, static_cast<__half*>(out_h16), NS, F, W, H1, W1,
                                                                                       P1, slope);
  count_launch();
  return launch_status();
}

extern "C" int osb_mrd_first_bwd(const void* g_h16, const float* spec, const float* w, float* dspec, float* dw, float* db, int32_t NS, int32_t F,
                                 int32_t W, int32_t H1, int32_t W1, int32_t P1, float inv_scale, void* stream) {
  OSB_REQUIRE(g_h16 && spec && w, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && F > 0 && W > 0 && H1 > 0 && W1 > 0 && P1 >= H1, OSB_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* g = static_cast<const __half*>(g_h16);
  int launched = 0;
  if (dspec != nullptr) {   // written (not accumulated)
    const long long n = static_cast<long long>(NS) * F * W;
    mrd_first_dx_kernel<<<grid_for(n, 256), 256, 0, s>>>(g, w, dspec, NS, F, W, H1, W1, P1, inv_scale);
    ++launched;
  }
  if (dw != nullptr && db != nullptr) {   // accumulated (+=): the caller zeroes them
    const long long rows = static_cast<long long>(NS) * W1 * P1;
    long long blocks = (rows + 511) / 512;
    if (blocks > 148 * 4) blocks = 148 * 4;
    const long long rpb = (rows + blocks - 1) / blocks;
    mrd_first_dw_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(g, spec, dw, db, NS, F, W, H1, W1, P1, inv_scale, rpb);
    ++launched;
  }
  count_launch(launched);
  return launch_status();
}

extern "C" int osb_wim2col_h16(const void* x, void* xcol, int32_t NS, int32_t W_in, int32_t W_out, int32_t P, int32_t C, int32_t KW, int32_t pw,
                               int32_t sw, void* stream) {
  OSB_REQUIRE(x && xcol, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && W_in > 0 && W_out > 0 && P > 0 && C % 8 == 0 && KW > 0 && sw >= 1, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * W_out * P * (KW * C / 8);
  wim2col_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x), static_cast<__half*>(xcol), NS, W_in,
                                                                               W_out, P, C, KW, pw, sw);
  count_launch();
  return launch_status();
}

extern "C" int osb_wcol2im_h16(const void* dxcol, void* dx, int32_t NS, int32_t W_in, int32_t W_out, int32_t P, int32_t C, int32_t KW, int32_t pw,
                               int32_t sw, void* stream) {
  OSB_REQUIRE(dxcol && dx, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && W_in > 0 && W_out > 0 && P > 0 && C % 8 == 0 && KW > 0 && sw >= 1, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * W_in * P * (C / 8);
  wcol2im_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(dxcol), static_cast<__half*>(dx), NS, W_in,
                                                                               W_out, P, C, KW, pw, sw);
  count_launch();
  return launch_status();
}

extern "C" int osb_mrd_post_fwd(const void* x_h16, const float* w, const float* bias, float* score, int32_t NS, int32_t W, int32_t H, int32_t P,
                                void* stream) {
  OSB_REQUIRE(x_h16 && w && bias && score, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && W > 0 && H > 0 && P >= H, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * W * H * 8;
  mrd_post_fwd_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x_h16), w, bias, score, NS, W, H, P);
  count_launch();
  return launch_status();
}

extern "C" int osb_mrd_post_bwd(const float* dscore, const void* x_h16, const float* w, void* dx_h16, float* dw, float* db, int32_t NS, int32_t W,
                                int32_t H, int32_t P, float scale, void* stream) {
  OSB_REQUIRE(dscore && w, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && W > 0 && H > 0 && P >= H, OSB_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int launched = 0;
  if (dx_h16 != nullptr) {
    const long long n = static_cast<long long>(NS) * W * P * 8;
    mrd_post_dx_kernel<<<grid_for(n, 256), 256, 0, s>>>(dscore, w, static_cast<__half*>(dx_h16), NS, W, H, P, scale);
    ++launched;
  }
  if (dw != nullptr && db != nullptr && x_h16 != nullptr) {   // accumulated (+=)
    const long long rows = static_cast<long long>(NS) * W * P;
    long long blocks = (rows + 255) / 256;
    if (blocks > 148 * 2) blocks = 148 * 2;
    const long long rpb = (rows + blocks - 1) / blocks;
    mrd_post_dw_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(dscore, static_cast<const __half*>(x_h16), dw, db, NS, W, H, P, rpb);
    ++launched;
  }
  count_launch(launched);
  return launch_status();
}

extern "C" int osb_spec_im2col_h16(const float* spec, void* xcol, int32_t NS, int32_t F, int32_t W, int32_t H1, int32_t W1, int32_t P1,
                                   void* stream) {
  OSB_REQUIRE(spec && xcol, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && F > 0 && W > 0 && H1 > 0 && W1 > 0 && P1 >= H1, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * W1 * P1 * 8;
  spec_im2col_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(spec, static_cast<__half*>(xcol), NS, F, W, H1, W1, P1);
  count_launch();
  return launch_status();
}

extern "C" int osb_spec_col2im(const void* col_h16, float* dspec, int32_t NS, int32_t F, int32_t W, int32_t H1, int32_t W1, int32_t P1,
                               float inv_scale, void* stream) {
  OSB_REQUIRE(col_h16 && dspec, OSB_ERR_ARG);
  OSB_REQUIRE(NS > 0 && F > 0 && W > 0 && H1 > 0 && W1 > 0 && P1 >= H1, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(NS) * F * W;
  spec_col2im_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(col_h16), dspec, NS, F, W, H1, W1,
                                                                                     P1, inv_scale);
  count_launch();
  return launch_status();
}
