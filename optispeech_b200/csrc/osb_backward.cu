// osb_backward.cu — HBM-bound backward kernels (everything in the backward pass that is not a GEMM).
//
// Gradients arrive in fp32 (possibly multiplied by the static loss scale the host applies); fp16
// outputs are operands of the following tcgen05 dgrad / wgrad contraction.  Parameter gradients are
// accumulated with fp32 atomics into caller-zeroed buffers: every warp first reduces over the rows
// it owns (grid-strided), so the number of atomics per element is (#warps in the grid), not (#rows).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int WPB = 8;  // warps per block

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldh4(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void sth4(__half* dst, float a, float b, float c, float d) {
  __half2 h0 = __floats2half2_rn(a, b);
  __half2 h1 = __floats2half2_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(dst) = u;
}
__device__ __forceinline__ void atomic_add4(float* dst, const float4& v) {
  atomicAdd(dst + 0, v.x);
  atomicAdd(dst + 1, v.y);
  atomicAdd(dst + 2, v.z);
  atomicAdd(dst + 3, v.w);
}
__device__ __forceinline__ float sum4(const float4& v) { return (v.x + v.y) + (v.z + v.w); }

// Per-column accumulators held by every warp of the block (column c = v*128 + lane*4) -> one vector reduction per column
// group and block: the warps' partials are summed through shared memory first, so a launch issues blocks x C/4 reds
// instead of warps x C scalar atomics on the same C addresses.  Block-collective (every warp calls it, also with zeros).
template <int VPL>
__device__ __forceinline__ void block_flush4(const float4 (&acc)[VPL], float* dst, float4* sred) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < VPL; ++v) sred[(w * VPL + v) * 32 + lane] = acc[v];
  __syncthreads();
  for (int i = threadIdx.x; i < VPL * 32; i += WPB * 32) {
    float4 t = sred[i];
#pragma unroll
    for (int ww = 1; ww < WPB; ++ww) {
      const float4 u = sred[ww * VPL * 32 + i];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* d = dst + (i >> 5) * 128 + (i & 31) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
  }
  __syncthreads();
}

// rows [r0, r1) of this warp: contiguous slab so that neighbouring warps touch neighbouring memory
__device__ __forceinline__ void warp_rows(long long rows, long long& r0, long long& r1) {
  const long long nw = static_cast<long long>(gridDim.x) * WPB;
  const long long w = static_cast<long long>(blockIdx.x) * WPB + (threadIdx.x >> 5);
  const long long per = (rows + nw - 1) / nw;
  r0 = w * per;
  r1 = r0 + per < rows ? r0 + per : rows;
}

inline int grid_for_rows(long long rows, int rows_per_warp) {
  long long warps = (rows + rows_per_warp - 1) / rows_per_warp;
  long long blocks = (warps + WPB - 1) / WPB;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------
// residual epilogue backward:  out = (x + gamma * z * rs[b]) * keep
//   dyg = fp16(dout * keep * rs * gamma)   -> A operand of the pwconv2 dgrad and dy of its wgrad
//   dgamma += sum_rows dout*keep*rs*z ;  db2 += sum_rows dout*keep*rs*gamma
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(WPB * 32)
resid_bwd_prep_kernel(const float* __restrict__ dout, const __half* __restrict__ z, const float* __restrict__ gamma,
                      const uint8_t* __restrict__ pad_mask, const float* __restrict__ row_scale, __half* __restrict__ dyg,
                      float* __restrict__ dgamma, float* __restrict__ db2, long long rows, int T) {
  constexpr int C = 128 * VPL;
  const int lane = threadIdx.x & 31;
  long long r0, r1;
  warp_rows(rows, r0, r1);
  float4 g[VPL], ag[VPL], ab[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    g[v] = ld4(gamma + v * 128 + lane * 4);
    ag[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = r0; r < r1; ++r) {
    float s = (pad_mask != nullptr && pad_mask[r]) ? 0.f : 1.f;
    if (row_scale != nullptr) s *= row_scale[r / T];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int c = v * 128 + lane * 4;
      float4 d = ld4(dout + r * C + c);
      d.x *= s; d.y *= s; d.z *= s; d.w *= s;
      const float4 zz = ldh4(z + r * C + c);
      ag[v].x = fmaf(d.x, zz.x, ag[v].x); ag[v].y = fmaf(d.y, zz.y, ag[v].y);
      ag[v].z = fmaf(d.z, zz.z, ag[v].z); ag[v].w = fmaf(d.w, zz.w, ag[v].w);
      d.x *= g[v].x; d.y *= g[v].y; d.z *= g[v].z; d.w *= g[v].w;
      ab[v].x += d.x; ab[v].y += d.y; ab[v].z += d.z; ab[v].w += d.w;
      sth4(dyg + r * C + c, d.x, d.y, d.z, d.w);
    }
  }
  __shared__ float4 sred[WPB * VPL * 32];
  block_flush4<VPL>(ag, dgamma, sred);
  block_flush4<VPL>(ab, db2, sred);
}

// ------------------------------------------------------------------------------------------
// column sums of an fp16 matrix (bias gradients): out[n] += sum_rows x[row, n].  N % 8 == 0.
// grid = (ceil(N / 2048), row-slabs), 256 threads each owning 8 columns.
// ------------------------------------------------------------------------------------------
// block = (128 column-threads of 8 columns each) x (8 row groups); the row groups are reduced through shared memory so
// that every block issues ONE atomic per column.
__global__ void __launch_bounds__(1024)
colsum_h16_kernel(const __half* __restrict__ x, float* __restrict__ out, long long rows, int N, int rows_per_block) {
  __shared__ float red[8][128 * 8 + 8];
  const int c = (blockIdx.x * 128 + threadIdx.x) * 8;
  const bool col_ok = c < N;
  const int rg = threadIdx.y;
  const long long rb0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long rb1 = rb0 + rows_per_block < rows ? rb0 + rows_per_block : rows;
  const long long per = (rb1 - rb0 + 7) / 8;
  const long long r0 = rb0 + rg * per;
  const long long r1 = col_ok ? (r0 + per < rb1 ? r0 + per : rb1) : r0;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {  // four independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = *reinterpret_cast<const uint4*>(x + (r + q) * N + c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[q]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
  for (; r < r1; ++r) {
    const uint4 u = *reinterpret_cast<const uint4*>(x + r * N + c);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rg][threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if (rg == 0 && col_ok) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x * 8 + j];
      atomicAdd(out + c + j, s);
    }
  }
}

// ------------------------------------------------------------------------------------------
// un-fold the LayerNorm affine that the forward pass folded into pwconv1:
//   W1f = W1 * diag(ln_w),  b1f = b1 + W1 @ ln_b
//   dW1[i,c] = dW1f[i,c] * ln_w[c] + db1[i] * ln_b[c]           (in place; b1f depends on W1 too)
//   dln_w[c] += sum_i dW1f[i,c] * W1[i,c] ;  dln_b[c] += sum_i db1[i] * W1[i,c]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_fold_bwd_kernel(float* __restrict__ dw1 /* in: dW1f, out: dW1 */, const float* __restrict__ w1, const float* __restrict__ ln_w,
                   const float* __restrict__ ln_b, const float* __restrict__ db1, float* __restrict__ dln_w, float* __restrict__ dln_b,
                   int I, int C, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int i0 = blockIdx.y * rows_per_block;
  const int i1 = min(I, i0 + rows_per_block);
  const float lw = ln_w[c], lb = ln_b[c];
  float aw = 0.f, ab = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float g = dw1[static_cast<long long>(i) * C + c];
    const float w = w1[static_cast<long long>(i) * C + c];
    const float d1 = db1[i];
    aw = fmaf(g, w, aw);
    ab = fmaf(d1, w, ab);
    dw1[static_cast<long long>(i) * C + c] = fmaf(g, lw, d1 * lb);
  }
  atomicAdd(dln_w + c, aw);
  atomicAdd(dln_b + c, ab);
}

// ------------------------------------------------------------------------------------------
// depthwise conv backward + residual path:
//   dx[b,t,c]  = dout[b,t,c] * keep[b,t] + sum_j w[c,j] * dd[b, t - j + 3, c]
//   ddw[c,j]  += sum_{b,t} dd[b,t,c] * x[b, t + j - 3, c] ;  ddb[c] += sum dd[b,t,c]
// one thread per channel, walking TT consecutive positions with sliding windows of dd and x.
// ------------------------------------------------------------------------------------------
__global__ void dwconv_bwd_kernel(const float* __restrict__ dd, const float* __restrict__ dout, const float* __restrict__ x,
                                  const float* __restrict__ w /*(C,7)*/, const uint8_t* __restrict__ pad_mask, float* __restrict__ dx,
                                  float* __restrict__ ddw, float* __restrict__ ddb, int T, int C, int TT) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.z;
  const int t_begin = blockIdx.y * TT;
  const int t_end = min(T, t_begin + TT);
  const long long base = static_cast<long long>(b) * T * C + c;
  float wj[7], aw[7], ab = 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    wj[j] = w[c * 7 + j];
    aw[j] = 0.f;
  }
  auto at = [&](const float* p, int t) -> float { return (t >= 0 && t < T) ? p[base + static_cast<long long>(t) * C] : 0.f; };
  float wd[7], wx[7];  // windows: index j <-> position t + j - 3
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    wd[j + 1] = at(dd, t_begin + j - 3);
    wx[j + 1] = at(x, t_begin + j - 3);
  }
  for (int t = t_begin; t < t_end; ++t) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      wd[j] = wd[j + 1];
      wx[j] = wx[j + 1];
    }
    wd[6] = at(dd, t + 3);
    wx[6] = at(x, t + 3);
    // dx: dd[t - j + 3] = wd[6 - j]
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) acc = fmaf(wj[j], wd[6 - j], acc);
    const float keep = (pad_mask != nullptr && pad_mask[static_cast<long long>(b) * T + t]) ? 0.f : 1.f;
    dx[base + static_cast<long long>(t) * C] = fmaf(dout[base + static_cast<long long>(t) * C], keep, acc);
    // parameter grads: dd[t] = wd[3], x[t + j - 3] = wx[j]
    const float d0 = wd[3];
#pragma unroll
    for (int j = 0; j < 7; ++j) aw[j] = fmaf(d0, wx[j], aw[j]);
    ab += d0;
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) atomicAdd(ddw + c * 7 + j, aw[j]);
  atomicAdd(ddb + c, ab);
}

// ------------------------------------------------------------------------------------------
// LayerNorm (with affine) backward; statistics recomputed from x.  One warp per row.
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(WPB * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ dx,
                     float* __restrict__ dw, float* __restrict__ db, long long rows, float eps) {
  constexpr int C = 128 * VPL;
  const int lane = threadIdx.x & 31;
  long long r0, r1;
  warp_rows(rows, r0, r1);
  float4 ww[VPL], aw[VPL], ab[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    ww[v] = ld4(w + v * 128 + lane * 4);
    aw[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = r0; r < r1; ++r) {
    float4 xv[VPL], g[VPL];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xv[v] = ld4(x + r * C + v * 128 + lane * 4);
      g[v] = ld4(dy + r * C + v * 128 + lane * 4);
      s += sum4(xv[v]);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xv[v].x -= mean; xv[v].y -= mean; xv[v].z -= mean; xv[v].w -= mean;
      q += (xv[v].x * xv[v].x + xv[v].y * xv[v].y) + (xv[v].z * xv[v].z + xv[v].w * xv[v].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xv[v].x *= rstd; xv[v].y *= rstd; xv[v].z *= rstd; xv[v].w *= rstd;  // xhat
      aw[v].x = fmaf(g[v].x, xv[v].x, aw[v].x); aw[v].y = fmaf(g[v].y, xv[v].y, aw[v].y);
      aw[v].z = fmaf(g[v].z, xv[v].z, aw[v].z); aw[v].w = fmaf(g[v].w, xv[v].w, aw[v].w);
      ab[v].x += g[v].x; ab[v].y += g[v].y; ab[v].z += g[v].z; ab[v].w += g[v].w;
      g[v].x *= ww[v].x; g[v].y *= ww[v].y; g[v].z *= ww[v].z; g[v].w *= ww[v].w;  // dxhat
      s1 += sum4(g[v]);
      s2 += (g[v].x * xv[v].x + g[v].y * xv[v].y) + (g[v].z * xv[v].z + g[v].w * xv[v].w);
    }
    const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float4 o;
      o.x = (g[v].x - m1 - xv[v].x * m2) * rstd;
      o.y = (g[v].y - m1 - xv[v].y * m2) * rstd;
      o.z = (g[v].z - m1 - xv[v].z * m2) * rstd;
      o.w = (g[v].w - m1 - xv[v].w * m2) * rstd;
      *reinterpret_cast<float4*>(dx + r * C + v * 128 + lane * 4) = o;
    }
  }
  __shared__ float4 sred[WPB * VPL * 32];
  block_flush4<VPL>(aw, dw, sred);
  block_flush4<VPL>(ab, db, sred);
}

// ------------------------------------------------------------------------------------------
// variance-predictor tail backward (last layer: y = LN(r), out = <y, lin_w> + lin_b, 0 at pads):
//   g_conv = fp16( LN_bwd(d_out * lin_w * ln_w ; r) * [r > 0] )
//   dlin_w += d_out * y ; dlin_b += d_out ; dln_w += d_out*lin_w * xhat ; dln_b += d_out*lin_w
// mode 1 (param grads only, for inner layers): gy (fp16) replaces d_out*lin_w and only dln_w/dln_b are produced.
// ------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(WPB * 32)
predictor_ln_bwd_kernel(const float* __restrict__ d_out, const __half* __restrict__ gy, const uint8_t* __restrict__ pad_mask,
                        const __half* __restrict__ r, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                        const float* __restrict__ lin_w, __half* __restrict__ g_conv, float* __restrict__ dlin_w,
                        float* __restrict__ dlin_b, float* __restrict__ dln_w, float* __restrict__ dln_b, long long rows, float eps,
                        float drop_p, float drop_inv_keep, unsigned long long drop_seed, const unsigned long long* drop_seed_dev) {
  constexpr int C = 128 * VPL;
  if (drop_p > 0.f && drop_seed_dev != nullptr) drop_seed += *drop_seed_dev;
  const int lane = threadIdx.x & 31;
  long long r0, r1;
  warp_rows(rows, r0, r1);
  const bool tail = d_out != nullptr;
  float4 lw[VPL], lb[VPL], li[VPL], a_lin[VPL], a_lnw[VPL], a_lnb[VPL];
  float a_linb = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int c = v * 128 + lane * 4;
    lw[v] = ld4(ln_w + c);
    lb[v] = tail ? ld4(ln_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    li[v] = tail ? ld4(lin_w + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    a_lin[v] = a_lnw[v] = a_lnb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = r0; row < r1; ++row) {
    float4 rv[VPL], xh[VPL], g[VPL];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      rv[v] = ldh4(r + row * C + v * 128 + lane * 4);
      s += sum4(rv[v]);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xh[v] = make_float4(rv[v].x - mean, rv[v].y - mean, rv[v].z - mean, rv[v].w - mean);
      q += (xh[v].x * xh[v].x + xh[v].y * xh[v].y) + (xh[v].z * xh[v].z + xh[v].w * xh[v].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    const float d = tail ? ((pad_mask != nullptr && pad_mask[row]) ? 0.f : d_out[row]) : 0.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      xh[v].x *= rstd; xh[v].y *= rstd; xh[v].z *= rstd; xh[v].w *= rstd;
      if (tail) {
        float4 mk = make_float4(1.f, 1.f, 1.f, 1.f);  // dropout mask (with 1/keep) of the forward pass
        if (drop_p > 0.f) {
          const unsigned long long base = static_cast<unsigned long long>(row) * C + v * 128 + lane * 4;
          float m4[4] = {1.f, 1.f, 1.f, 1.f};
          dropout_apply(m4, drop_seed, base, drop_p, drop_inv_keep);
          mk = make_float4(m4[0], m4[1], m4[2], m4[3]);
        }
        g[v] = make_float4(d * li[v].x * mk.x, d * li[v].y * mk.y, d * li[v].z * mk.z, d * li[v].w * mk.w);
        a_lin[v].x = fmaf(d * mk.x, fmaf(xh[v].x, lw[v].x, lb[v].x), a_lin[v].x);
        a_lin[v].y = fmaf(d * mk.y, fmaf(xh[v].y, lw[v].y, lb[v].y), a_lin[v].y);
        a_lin[v].z = fmaf(d * mk.z, fmaf(xh[v].z, lw[v].z, lb[v].z), a_lin[v].z);
        a_lin[v].w = fmaf(d * mk.w, fmaf(xh[v].w, lw[v].w, lb[v].w), a_lin[v].w);
      } else {
        g[v] = ldh4(gy + row * C + v * 128 + lane * 4);
      }
      a_lnw[v].x = fmaf(g[v].x, xh[v].x, a_lnw[v].x); a_lnw[v].y = fmaf(g[v].y, xh[v].y, a_lnw[v].y);
      a_lnw[v].z = fmaf(g[v].z, xh[v].z, a_lnw[v].z); a_lnw[v].w = fmaf(g[v].w, xh[v].w, a_lnw[v].w);
      a_lnb[v].x += g[v].x; a_lnb[v].y += g[v].y; a_lnb[v].z += g[v].z; a_lnb[v].w += g[v].w;
      if (tail) {
        g[v].x *= lw[v].x; g[v].y *= lw[v].y; g[v].z *= lw[v].z; g[v].w *= lw[v].w;
        s1 += sum4(g[v]);
        s2 += (g[v].x * xh[v].x + g[v].y * xh[v].y) + (g[v].z * xh[v].z + g[v].w * xh[v].w);
      }
    }
    if (tail) {
      a_linb += d;
      const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const float ox = rv[v].x > 0.f ? (g[v].x - m1 - xh[v].x * m2) * rstd : 0.f;
        const float oy = rv[v].y > 0.f ? (g[v].y - m1 - xh[v].y * m2) * rstd : 0.f;
        const float oz = rv[v].z > 0.f ? (g[v].z - m1 - xh[v].z * m2) * rstd : 0.f;
        const float ow = rv[v].w > 0.f ? (g[v].w - m1 - xh[v].w * m2) * rstd : 0.f;
        sth4(g_conv + row * C + v * 128 + lane * 4, ox, oy, oz, ow);
      }
    }
  }
  __shared__ float4 sred[WPB * VPL * 32];
  block_flush4<VPL>(a_lnw, dln_w, sred);
  block_flush4<VPL>(a_lnb, dln_b, sred);
  if (tail) {
    block_flush4<VPL>(a_lin, dlin_w, sred);
    if (lane == 0 && r0 < r1) atomicAdd(dlin_b, a_linb);
  }
}

// ------------------------------------------------------------------------------------------
// variance embedding backward: out = (x + bias + conv(val)) * keep
//   dx = dout * keep ; dw[c,j] += sum dout*keep*val[b,t+j-h] ; db[c] += sum dout*keep
// ------------------------------------------------------------------------------------------
__global__ void variance_embed_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ val,
                                          const uint8_t* __restrict__ pad_mask, const float* __restrict__ emb_scale,
                                          float* __restrict__ dx, float* __restrict__ dw,
                                          float* __restrict__ db, int T, int C, int ksize, int TT) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.z;
  const int t_begin = blockIdx.y * TT;
  const int t_end = min(T, t_begin + TT);
  const int halfk = (ksize - 1) / 2;
  float aw[16], ab = 0.f;
  for (int j = 0; j < ksize; ++j) aw[j] = 0.f;
  for (int t = t_begin; t < t_end; ++t) {
    const long long row = static_cast<long long>(b) * T + t;
    const float keep = (pad_mask != nullptr && pad_mask[row]) ? 0.f : 1.f;
    const float g = dout[row * C + c] * keep;
    if (dx != nullptr) dx[row * C + c] = g;
    const float ge = emb_scale != nullptr ? g * emb_scale[row * C + c] : g;  // gradient of the (dropped) embedding branch
    ab += ge;
    for (int j = 0; j < ksize; ++j) {
      const int tt = t + j - halfk;
      const float v = (tt >= 0 && tt < T) ? val[static_cast<long long>(b) * T + tt] : 0.f;
      aw[j] = fmaf(ge, v, aw[j]);
    }
  }
  for (int j = 0; j < ksize; ++j) atomicAdd(dw + c * ksize + j, aw[j]);
  atomicAdd(db + c, ab);
}

// ------------------------------------------------------------------------------------------
// text embedding backward: dtable[id] += sqrt(dim) * dout (padding row 0 receives nothing);
// dscale += sum dout * pe
// ------------------------------------------------------------------------------------------
__global__ void embed_text_bwd_kernel(const float* __restrict__ dout, const long long* __restrict__ ids,
                                      const float* __restrict__ inv_freq, float* __restrict__ dtable, float* __restrict__ dscale,
                                      int rows, int T, int dim, int n_vocab, int padding_idx, float embed_scale) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int t = row % T;
  const long long id = ids[row];
  const int half = dim >> 1;
  float acc = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float g = dout[static_cast<long long>(row) * dim + c];
    const int j = c < half ? c : c - half;
    const float ang = static_cast<float>(t) * inv_freq[j];
    acc = fmaf(g, c < half ? sinf(ang) : cosf(ang), acc);
    if (id != padding_idx && id >= 0 && id < n_vocab) atomicAdd(dtable + id * dim + c, embed_scale * g);
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(dscale, acc);
}

}  // namespace
}  // namespace osb

using namespace osb;

#define OSB_VPL_SWITCH(C, CALL)                      \
  switch ((C) / 128) {                               \
    case 1: { constexpr int VPL = 1; CALL; } break;  \
    case 2: { constexpr int VPL = 2; CALL; } break;  \
    case 3: { constexpr int VPL = 3; CALL; } break;  \
    case 4: { constexpr int VPL = 4; CALL; } break;  \
    default: return OSB_ERR_SHAPE;                   \
  }

extern "C" int osb_resid_bwd_prep(const float* dout, const void* z_h16, const float* gamma, const uint8_t* pad_mask,
                                  const float* row_scale, void* dyg_h16, float* dgamma, float* db2, int64_t rows, int32_t T,
                                  int32_t C, void* stream) {
  OSB_REQUIRE(dout && z_h16 && gamma && dyg_h16 && dgamma && db2, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && T > 0 && C % 128 == 0 && C <= 512, OSB_ERR_SHAPE);
  const int grid = grid_for_rows(rows, 6);   // short per-warp row runs: the row loop is a latency chain, parallelism hides it
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSB_VPL_SWITCH(C, (resid_bwd_prep_kernel<VPL><<<grid, WPB * 32, 0, s>>>(dout, static_cast<const __half*>(z_h16), gamma, pad_mask,
                                                                          row_scale, static_cast<__half*>(dyg_h16), dgamma, db2,
                                                                          rows, T)));
  count_launch();
  return launch_status();
}

extern "C" int osb_colsum_h16(const void* x_h16, float* out, int64_t rows, int32_t N, void* stream) {
  OSB_REQUIRE(x_h16 && out, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && N > 0 && N % 8 == 0, OSB_ERR_SHAPE);
  const int gx = (N / 8 + 127) / 128;
  int rpb = static_cast<int>((rows * gx + 147) / 148);  // ~one block per SM
  if (rpb < 128) rpb = 128;
  const int gy = static_cast<int>((rows + rpb - 1) / rpb);
  colsum_h16_kernel<<<dim3(gx, gy), dim3(128, 8), 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(x_h16), out, rows, N,
                                                                                         rpb);
  count_launch();
  return launch_status();
}

extern "C" int osb_ln_fold_bwd(float* dw1, const float* w1, const float* ln_w, const float* ln_b, const float* db1, float* dln_w,
                               float* dln_b, int32_t I, int32_t C, void* stream) {
  OSB_REQUIRE(dw1 && w1 && ln_w && ln_b && db1 && dln_w && dln_b, OSB_ERR_ARG);
  OSB_REQUIRE(I > 0 && C > 0, OSB_ERR_SHAPE);
  const int rpb = 32;
  ln_fold_bwd_kernel<<<dim3((C + 255) / 256, (I + rpb - 1) / rpb), 256, 0, static_cast<cudaStream_t>(stream)>>>(dw1, w1, ln_w, ln_b,
                                                                                                              db1, dln_w, dln_b, I, C, rpb);
  count_launch();
  return launch_status();
}

extern "C" int osb_dwconv_bwd(const float* dd, const float* dout, const float* x, const float* w, const uint8_t* pad_mask, float* dx,
                              float* ddw, float* ddb, int32_t B, int32_t T, int32_t C, void* stream) {
  OSB_REQUIRE(dd && dout && x && w && dx && ddw && ddb, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0, OSB_ERR_SHAPE);
  const int TT = 32;
  const int threads = C >= 256 ? 128 : 64;
  dim3 grid((C + threads - 1) / threads, (T + TT - 1) / TT, B);
  dwconv_bwd_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(dd, dout, x, w, pad_mask, dx, ddw, ddb, T, C, TT);
  count_launch();
  return launch_status();
}

extern "C" int osb_layernorm_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t rows,
                                 int32_t C, float eps, void* stream) {
  OSB_REQUIRE(dy && x && w && dx && dw && db, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C % 128 == 0 && C <= 512, OSB_ERR_SHAPE);
  const int grid = grid_for_rows(rows, 6);   // short per-warp row runs: the row loop is a latency chain, parallelism hides it
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSB_VPL_SWITCH(C, (layernorm_bwd_kernel<VPL><<<grid, WPB * 32, 0, s>>>(dy, x, w, dx, dw, db, rows, eps)));
  count_launch();
  return launch_status();
}

extern "C" int osb_predictor_tail_bwd(const float* d_out, const uint8_t* pad_mask, const void* r_h16, const float* ln_w,
                                      const float* ln_b, const float* lin_w, void* g_conv_h16, float* dlin_w, float* dlin_b,
                                      float* dln_w, float* dln_b, int64_t rows, int32_t C, float eps, float dropout_p,
                                      uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream) {
  OSB_REQUIRE(d_out && r_h16 && ln_w && ln_b && lin_w && g_conv_h16 && dlin_w && dlin_b && dln_w && dln_b, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C % 128 == 0 && C <= 512, OSB_ERR_SHAPE);
  OSB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, OSB_ERR_ARG);
  const float inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 0.f;
  const int grid = grid_for_rows(rows, 6);   // short per-warp row runs: the row loop is a latency chain, parallelism hides it
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSB_VPL_SWITCH(C, (predictor_ln_bwd_kernel<VPL><<<grid, WPB * 32, 0, s>>>(
                        d_out, nullptr, pad_mask, static_cast<const __half*>(r_h16), ln_w, ln_b, lin_w,
                        static_cast<__half*>(g_conv_h16), dlin_w, dlin_b, dln_w, dln_b, rows, eps, dropout_p, inv_keep,
                        static_cast<unsigned long long>(dropout_seed), reinterpret_cast<const unsigned long long*>(dropout_seed_dev))));
  count_launch();
  return launch_status();
}

extern "C" int osb_ln_param_grad(const void* gy_h16, const void* r_h16, const float* ln_w, float* dln_w, float* dln_b, int64_t rows,
                                 int32_t C, float eps, void* stream) {
  OSB_REQUIRE(gy_h16 && r_h16 && ln_w && dln_w && dln_b, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && C % 128 == 0 && C <= 512, OSB_ERR_SHAPE);
  const int grid = grid_for_rows(rows, 6);   // short per-warp row runs: the row loop is a latency chain, parallelism hides it
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSB_VPL_SWITCH(C, (predictor_ln_bwd_kernel<VPL><<<grid, WPB * 32, 0, s>>>(
                        nullptr, static_cast<const __half*>(gy_h16), nullptr, static_cast<const __half*>(r_h16), ln_w, nullptr,
                        nullptr, nullptr, nullptr, nullptr, dln_w, dln_b, rows, eps, 0.f, 0.f, 0ull, nullptr)));
  count_launch();
  return launch_status();
}

extern "C" int osb_variance_embed_bwd(const float* dout, const float* val, const uint8_t* pad_mask, const float* emb_scale, float* dx,
                                      float* dw, float* db,
                                      int32_t B, int32_t T, int32_t C, int32_t ksize, void* stream) {
  OSB_REQUIRE(dout && val && dw && db, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0 && ksize > 0 && ksize <= 16, OSB_ERR_SHAPE);
  const int TT = 32;
  const int threads = 128;
  dim3 grid((C + threads - 1) / threads, (T + TT - 1) / TT, B);
  variance_embed_bwd_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(dout, val, pad_mask, emb_scale, dx, dw, db, T, C, ksize, TT);
  count_launch();
  return launch_status();
}

extern "C" int osb_embed_text_bwd(const float* dout, const int64_t* ids, const float* inv_freq, float* dtable, float* dscale,
                                  int32_t B, int32_t T, int32_t dim, int32_t n_vocab, int32_t padding_idx, void* stream) {
  OSB_REQUIRE(dout && ids && inv_freq && dtable && dscale, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && dim > 0, OSB_ERR_SHAPE);
  const int rows = B * T;
  embed_text_bwd_kernel<<<(rows + WPB - 1) / WPB, WPB * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, reinterpret_cast<const long long*>(ids), inv_freq, dtable, dscale, rows, T, dim, n_vocab, padding_idx,
      sqrtf(static_cast<float>(dim)));
  count_launch();
  return launch_status();
}
