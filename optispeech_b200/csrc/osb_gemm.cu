// osb_gemm.cu — tcgen05 / TMA contraction kernels with fused epilogues.
//
//   gemm_nt_kernel   : out = epi( sum_tap A[b, t+tap-pad, :] . W[tap, n, :] )      (forward, dgrad)
//   gemm_wgrad_kernel: dw[tap, n, k] += sum_{b,t} dy[b,t,n] * a[b, t+tap-pad, k]   (weight gradient)
//
// Data layout in HBM: activations channels-last fp16 (B, T, ld), weights fp16 (taps, N, ld).
// One CTA owns a 128-row output tile; operands are staged by TMA into 128B-swizzled shared
// memory (a ring of full/empty mbarriers), one elected thread issues tcgen05.mma with the fp32
// accumulator in TMEM, and four epilogue warps read it back (one thread per output row, so
// LayerNorm / dot-product epilogues are thread-local).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {

unsigned long long g_launch_count = 0;

namespace {

constexpr int BM = 128;          // rows per CTA tile (UMMA M)
constexpr int ROW_BYTES = 128;   // bytes along the contraction dim per smem row (= swizzle span)
constexpr int BKE = 64;          // fp16 elements per ROW_BYTES
constexpr int UMMA_K_BYTES = 32; // bytes consumed along K by one tcgen05.mma (16 fp16)

struct GemmKParams {
  int T, B, N, K, taps, pad;
  int m_tiles;  // row tiles per batch
  int flags;
  void* out;
  void* aux;
  long long ldo;
  const float* bias;
  const float* resid;
  const float* gamma;
  const float* row_scale;
  const uint8_t* pad_mask;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  const float* dot_w;
  const float* dot_b;
  float* out_dot;
  const void* aux_in;
  const float* row_stat;
  float drop_p;
  float drop_inv_keep;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
  int w_lo_slice;   // slice offset of the lo parts of W (split precision, shared weights)
  int w_lo_koff;    // K offset of the lo parts of W (split precision, batched weights stored [hi K | lo K])
  int w_batch;      // 1: slice index += batch (per-batch B operand)
  const long long* col_len;
};

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int BN>
struct GemmCfg {
  static_assert(BN % 32 == 0 && BN <= 512, "BN");
  static constexpr int NINST = (BN <= 256) ? BN : BN / 2;  // N of one tcgen05.mma (<= 256, % 16 == 0)
  static constexpr int NCHUNK = BN / NINST;
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = cmin(6, (216 * 1024) / STAGE_BYTES);
  static constexpr int TMEM_COLS = pow2_cols(BN);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(NINST % 16 == 0 && NINST <= 256, "NINST");
  static_assert(STAGES >= 2, "stages");
};

// ------------------------------------------------------------------------------------------
// epilogues: one thread == one output row; `taddr` addresses this thread's TMEM lane, col 0.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld_chunk(uint32_t taddr, int c0, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld_h16x32(const __half* src, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + i);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      v[i + 2 * j] = f.x;
      v[i + 2 * j + 1] = f.y;
    }
  }
}

__device__ __forceinline__ void st_f32x32(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
__device__ __forceinline__ void st_h16x32_lo(__half* dst, const float (&v)[32]) {
  // residual after fp16 rounding: lo = fp16(v - float(fp16(v)))
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = v[i] - __half2float(__float2half_rn(v[i]));
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    __half2 h0 = __floats2half2_rn(r[i], r[i + 1]);
    __half2 h1 = __floats2half2_rn(r[i + 2], r[i + 3]);
    __half2 h2 = __floats2half2_rn(r[i + 4], r[i + 5]);
    __half2 h3 = __floats2half2_rn(r[i + 6], r[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
}
__device__ __forceinline__ void st_h16x32(__half* dst, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    __half2 h0 = __floats2half2_rn(v[i], v[i + 1]);
    __half2 h1 = __floats2half2_rn(v[i + 2], v[i + 3]);
    __half2 h2 = __floats2half2_rn(v[i + 4], v[i + 5]);
    __half2 h3 = __floats2half2_rn(v[i + 6], v[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2);
    u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
}

// fp16 destination row `row`, columns [n, n+32): plain (stride ldo) or split [hi ldo | lo ldo] (stride 2*ldo)
__device__ __forceinline__ void store_h(const GemmKParams& p, void* base, long long row, int n, const float (&v)[32]) {
  __half* hb = static_cast<__half*>(base);
  if (p.flags & OSB_FLAG_SPLIT_OUT) {
    __half* r = hb + row * (2 * p.ldo);
    st_h16x32(r + n, v);
    st_h16x32_lo(r + p.ldo + n, v);
  } else {
    st_h16x32(hb + row * p.ldo + n, v);
  }
}

template <int EPI, int BN>
__device__ __forceinline__ void run_epilogue(const GemmKParams& p, uint32_t taddr, int b, int t, int n0) {
  unsigned long long drop_seed = p.drop_seed;
  if (p.drop_p > 0.f && p.drop_seed_dev != nullptr) drop_seed += *p.drop_seed_dev;
  const bool valid = t < p.T;
  const long long row = static_cast<long long>(b) * p.T + t;
  const bool padded = (p.pad_mask != nullptr) && valid && (p.pad_mask[row] != 0);
  float v[32];

  if constexpr (EPI == OSB_EPI_ATTN_LOGP) {
    // the CTA owns whole rows of the (Tm x Tx) attention: distance, masked log-softmax and prior are thread-local
    const int ncols = p.N;                                              // true number of columns (<= BN)
    const int nvalid = static_cast<int>(p.col_len[b]) < ncols ? static_cast<int>(p.col_len[b]) : ncols;
    const float nf = valid ? p.row_stat[row] : 0.f;
    const float* ne = p.bias + static_cast<long long>(b) * ncols;
    // pass 1: score = -distance, parked back in TMEM; running max / sum of exp (online softmax: one pass for both)
    float mx = -INFINITY, sum = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int n = c0 + i;
        v[i] = n < nvalid ? -sqrtf(fmaxf(nf + __ldg(ne + n) - 2.f * v[i], 0.f)) : -INFINITY;
        cm = fmaxf(cm, v[i]);
      }
      if (cm > mx) {
        sum *= __expf(mx - cm);  // exp(-inf) = 0 on the first chunk
        mx = cm;
      }
      float cs = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) cs += v[i] > -INFINITY ? __expf(v[i] - mx) : 0.f;  // also keeps an all-masked row NaN-free
      sum += cs;
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
      tmem_st_32x32(taddr + static_cast<uint32_t>(c0), r);
    }
    tmem_st_wait();
    const float lse = mx + logf(sum);
    if (valid && p.out_dot != nullptr) p.out_dot[row] = lse;
    // pass 2: log-probability + prior; 16-byte accesses when the row pitch allows (one thread = one row of the output)
    const bool vec = (p.ldo & 3) == 0;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (c0 >= ncols) break;
      ld_chunk(taddr, c0, v);
      if (valid) {
        float* orow = static_cast<float*>(p.out) + row * p.ldo;
        const float* prow = p.resid + row * p.ldo;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int n = c0 + i;
          if (vec && n + 3 < ncols) {
            const float4 pr = *reinterpret_cast<const float4*>(prow + n);
            float4 o;
            o.x = (v[i] - lse) + pr.x;       // masked columns hold -inf already
            o.y = (v[i + 1] - lse) + pr.y;
            o.z = (v[i + 2] - lse) + pr.z;
            o.w = (v[i + 3] - lse) + pr.w;
            if (n >= nvalid) o.x = -INFINITY;
            if (n + 1 >= nvalid) o.y = -INFINITY;
            if (n + 2 >= nvalid) o.z = -INFINITY;
            if (n + 3 >= nvalid) o.w = -INFINITY;
            *reinterpret_cast<float4*>(orow + n) = o;
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < ncols) orow[n + u] = n + u < nvalid ? (v[i + u] - lse) + prow[n + u] : -INFINITY;
          }
        }
      }
    }
  } else if constexpr (EPI == OSB_EPI_AXPY) {
    const float alpha = valid ? p.row_stat[row] : 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      if (valid) {
        const int n = n0 + c0;
        const float* rp = p.resid + row * p.ldo + n;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
          v[i + 0] = fmaf(alpha, r4.x, v[i + 0]);
          v[i + 1] = fmaf(alpha, r4.y, v[i + 1]);
          v[i + 2] = fmaf(alpha, r4.z, v[i + 2]);
          v[i + 3] = fmaf(alpha, r4.w, v[i + 3]);
        }
        st_f32x32(static_cast<float*>(p.out) + row * p.ldo + n, v);
      }
    }
  } else if constexpr (EPI == OSB_EPI_GELU_BWD || EPI == OSB_EPI_RELU_BWD) {
    float pre[32];
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      if (valid) {
        const int n = n0 + c0;
        ld_h16x32(static_cast<const __half*>(p.aux_in) + row * p.ldo + n, pre);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if constexpr (EPI == OSB_EPI_GELU_BWD) v[i] *= gelu_erf_grad(pre[i]);
          else v[i] = pre[i] > 0.f ? v[i] * (p.drop_p > 0.f ? p.drop_inv_keep : 1.f) : 0.f;  // aux_in = relu output AFTER dropout:
        }                                                                                      // dropped elements are 0 there
        st_h16x32(static_cast<__half*>(p.out) + row * p.ldo + n, v);
      }
    }
  } else if constexpr (EPI == OSB_EPI_LN_BWD) {
    // acc = d(xhat); out = (acc - mean(acc) - xhat * mean(acc * xhat)) * rstd       (rows are independent)
    float xh[32];
    const float inv_n = 1.f / static_cast<float>(BN);
    const __half* xrow = static_cast<const __half*>(p.aux_in) + (valid ? row : 0) * p.ldo;
    float s1 = 0.f, s2 = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(xrow + c0, xh);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s1 += v[i];
        s2 = fmaf(v[i], xh[i], s2);
      }
    }
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    const float rstd = valid ? p.row_stat[row] : 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(xrow + c0, xh);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (v[i] - m1 - xh[i] * m2) * rstd;
      if (valid) st_f32x32(static_cast<float*>(p.out) + row * p.ldo + c0, v);
    }
  } else if constexpr (EPI == OSB_EPI_RELU_LN_BWD) {
    // recompute the forward LN statistics of r = relu(conv) (fp16-saved), then LN backward and the ReLU gate
    float r[32];
    const float inv_n = 1.f / static_cast<float>(BN);
    const __half* rrow = static_cast<const __half*>(p.aux_in) + (valid ? row : 0) * p.ldo;
    float s = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_h16x32(rrow + c0, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) s += r[i];
    }
    const float mean = s * inv_n;
    float q = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_h16x32(rrow + c0, r);
#pragma unroll
      for (int i = 0; i < 32; ++i) q = fmaf(r[i] - mean, r[i] - mean, q);
    }
    const float rstd = rsqrtf(q * inv_n + p.ln_eps);
    float s1 = 0.f, s2 = 0.f;
    const unsigned long long dbase = static_cast<unsigned long long>(valid ? row : 0) * BN;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(rrow + c0, r);
      if (p.drop_p > 0.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(drop_seed, dbase + c0 + i, p.drop_p, p.drop_inv_keep);
      }
      if (valid && (p.flags & OSB_FLAG_OUT_H16)) st_h16x32(static_cast<__half*>(p.aux) + row * p.ldo + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float g = v[i] * __ldg(p.ln_w + c0 + i);
        s1 += g;
        s2 = fmaf(g, (r[i] - mean) * rstd, s2);
      }
    }
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      ld_h16x32(rrow + c0, r);
      if (p.drop_p > 0.f) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(drop_seed, dbase + c0 + i, p.drop_p, p.drop_inv_keep);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float g = v[i] * __ldg(p.ln_w + c0 + i);
        const float d = (g - m1 - (r[i] - mean) * rstd * m2) * rstd;
        v[i] = r[i] > 0.f ? d : 0.f;
      }
      if (valid) st_h16x32(static_cast<__half*>(p.out) + row * p.ldo + c0, v);
    }
  } else if constexpr (EPI == OSB_EPI_BIAS || EPI == OSB_EPI_GELU || EPI == OSB_EPI_RELU || EPI == OSB_EPI_RESID) {
    const float keep = ((p.flags & OSB_FLAG_KEEPMASK) && padded) ? 0.f : 1.f;
    const float rs = (EPI == OSB_EPI_RESID && p.row_scale != nullptr) ? p.row_scale[b] : 1.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
      const int n = n0 + c0;
      if (p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + n + i);
      }
      if constexpr (EPI == OSB_EPI_BIAS) {
        if (p.flags & OSB_FLAG_CLIP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], -1.f), 1.f);
        }
        if (p.flags & OSB_FLAG_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= keep;
        if (valid) {
          if (!(p.flags & OSB_FLAG_NO_F32)) st_f32x32(static_cast<float*>(p.out) + row * p.ldo + n, v);
          if (p.flags & OSB_FLAG_OUT_H16) store_h(p, p.aux, row, n, v);
        }
      } else if constexpr (EPI == OSB_EPI_GELU) {
        if (valid && (p.flags & OSB_FLAG_SAVE_PRE)) st_h16x32(static_cast<__half*>(p.aux) + row * p.ldo + n, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
        if (valid) store_h(p, p.out, row, n, v);
      } else if constexpr (EPI == OSB_EPI_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        if (p.drop_p > 0.f) {  // Dropout after the ReLU (MultiLayeredConv1d, multi_layer_conv.py:60-62)
          const unsigned long long dbase = static_cast<unsigned long long>(row) * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(drop_seed, dbase + i, p.drop_p, p.drop_inv_keep);
        }
        if (valid) store_h(p, p.out, row, n, v);
      } else {  // RESID
        if (valid && (p.flags & OSB_FLAG_SAVE_PRE)) st_h16x32(static_cast<__half*>(p.aux) + row * p.ldo + n, v);
        if (p.drop_p > 0.f) {  // element dropout on the branch before the residual add (EncoderLayer, encoder_layer.py:103,111)
          const unsigned long long dbase = static_cast<unsigned long long>(row) * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(drop_seed, dbase + i, p.drop_p, p.drop_inv_keep);
        }
        if (valid) {
          const float* rp = p.resid + row * p.ldo + n;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
            v[i + 0] = (r4.x + __ldg(p.gamma + n + i + 0) * v[i + 0] * rs) * keep;
            v[i + 1] = (r4.y + __ldg(p.gamma + n + i + 1) * v[i + 1] * rs) * keep;
            v[i + 2] = (r4.z + __ldg(p.gamma + n + i + 2) * v[i + 2] * rs) * keep;
            v[i + 3] = (r4.w + __ldg(p.gamma + n + i + 3) * v[i + 3] * rs) * keep;
          }
          st_f32x32(static_cast<float*>(p.out) + row * p.ldo + n, v);
          if (p.flags & OSB_FLAG_OUT_H16) store_h(p, p.aux, row, n, v);
        }
      }
    }
  } else {  // row-wise LayerNorm epilogues; the CTA owns the whole row (BN == N, n0 == 0)
    constexpr bool kRelu = (EPI == OSB_EPI_RELU_LN);
    const float inv_n = 1.f / static_cast<float>(BN);
    // pass 1: mean
    float s = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = v[i] + (p.bias ? __ldg(p.bias + c0 + i) : 0.f);
        if (kRelu) x = fmaxf(x, 0.f);
        s += x;
      }
    }
    const float mean = s * inv_n;
    // pass 2: biased variance around the mean (what at::layer_norm computes)
    float q = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = v[i] + (p.bias ? __ldg(p.bias + c0 + i) : 0.f);
        if (kRelu) x = fmaxf(x, 0.f);
        const float d = x - mean;
        q += d * d;
      }
    }
    const float rstd = rsqrtf(q * inv_n + p.ln_eps);
    // pass 3: normalise, affine, store
    float dot = 0.f;
    for (int c0 = 0; c0 < BN; c0 += 32) {
      ld_chunk(taddr, c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = v[i] + (p.bias ? __ldg(p.bias + c0 + i) : 0.f);
        if (kRelu) x = fmaxf(x, 0.f);
        v[i] = x;
      }
      if (valid && (p.flags & OSB_FLAG_SAVE_PRE)) st_h16x32(static_cast<__half*>(p.aux) + row * p.ldo + c0, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = (v[i] - mean) * rstd * __ldg(p.ln_w + c0 + i) + __ldg(p.ln_b + c0 + i);
      }
      if (kRelu && p.drop_p > 0.f) {
        const unsigned long long base = static_cast<unsigned long long>(valid ? row : 0) * BN + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= dropout_scale(drop_seed, base + i, p.drop_p, p.drop_inv_keep);
      }
      if (p.flags & OSB_FLAG_DOT) {
#pragma unroll
        for (int i = 0; i < 32; ++i) dot += v[i] * __ldg(p.dot_w + c0 + i);
      }
      if (valid) {
        if constexpr (EPI == OSB_EPI_RELU_LN) {
          if (p.out != nullptr) store_h(p, p.out, row, c0, v);
        } else {
          st_f32x32(static_cast<float*>(p.out) + row * p.ldo + c0, v);
          if (p.flags & OSB_FLAG_OUT_H16) store_h(p, p.aux, row, c0, v);
        }
      }
    }
    if ((p.flags & OSB_FLAG_DOT) && valid) p.out_dot[row] = padded ? 0.f : (dot + (p.dot_b != nullptr ? __ldg(p.dot_b) : 0.f));
  }
}

// ------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------
template <int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmKParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int b = blockIdx.x / p.m_tiles;
  const int t0 = (blockIdx.x % p.m_tiles) * BM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = (p.K + BKE - 1) / BKE;
  // split precision: three passes over K per tap — (a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo)
  const bool split_in = (p.flags & OSB_FLAG_SPLIT_IN) != 0;
  const int nsub = split_in ? 3 : 1;
  const int iters = p.taps * nsub * num_kb;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        const int tap = it / (nsub * num_kb);
        const int rem = it - tap * (nsub * num_kb);
        const int sub = rem / num_kb;
        const int kb = rem - sub * num_kb;
        const int a_k = kb * BKE + (sub == 1 ? p.K : 0);
        const int w_slice = tap + (sub == 2 ? p.w_lo_slice : 0) + (p.w_batch ? b : 0);
        const int w_k = kb * BKE + (sub == 2 ? p.w_lo_koff : 0);
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        tma_load_3d(sA, &tmA, &full_bar[s], a_k, t0 + tap - p.pad, b);
#pragma unroll
        for (int c = 0; c < Cfg::NCHUNK; ++c)
          tma_load_3d(sB + c * Cfg::NINST * ROW_BYTES, &tmW, &full_bar[s], w_k, n0 + c * Cfg::NINST, w_slice);
      }
    }
  } else if (warp == 1) {
    // whole warp, uniform control flow (descriptors live in uniform registers); one elected lane issues
    constexpr uint32_t idesc = make_instr_desc(OSB_F16, BM, Cfg::NINST, 0, 0);
    const uint64_t d0 = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    for (int it = 0; it < iters; ++it) {
      const int s = it % Cfg::STAGES;
      const uint32_t ph = (it / Cfg::STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint64_t da0 = d0 + static_cast<uint64_t>((s * Cfg::STAGE_BYTES) >> 4);
        const uint64_t db0 = da0 + static_cast<uint64_t>(Cfg::A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < ROW_BYTES / UMMA_K_BYTES; ++k) {
#pragma unroll
          for (int c = 0; c < Cfg::NCHUNK; ++c) {
            umma_ss<false>(tmem_base + c * Cfg::NINST, da0 + static_cast<uint64_t>((k * UMMA_K_BYTES) >> 4),
                           db0 + static_cast<uint64_t>((c * Cfg::NINST * ROW_BYTES + k * UMMA_K_BYTES) >> 4), idesc,
                           (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(tmem_full_bar);  // accumulator complete
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    run_epilogue<EPI, BN>(p, taddr, b, t0 + q * 32 + lane, n0);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------
// weight-gradient kernel: both operands MN-major (read straight from channels-last tensors)
// ------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 64;                      // contraction rows (b,t) per stage
constexpr int WG_TILE = 128;                     // output tile is 128 (n) x 128 (k)
constexpr int WG_OP_BYTES = WG_ROWS * 256;       // 64 rows x 128 elems x 2 B, stored as 2 chunks of 64 elems
constexpr int WG_STAGE_BYTES = 2 * WG_OP_BYTES;  // dy tile + a tile
constexpr int WG_STAGES = 6;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgradParams {
  int T, B, N, K, taps, pad;
  int row_blocks_per_batch;  // ceil(T / 64)
  int splits;
  int per_batch;             // 1: one (N, K) output per batch (batched matmul A^T B), taps == 1
  float* dw;
};

__global__ void __launch_bounds__(192, 1)
gemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmA, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n0 = blockIdx.x * WG_TILE;
  const int k0 = blockIdx.y * WG_TILE;
  const int zsel = blockIdx.z / p.splits;       // tap, or batch in per-batch mode
  const int tap = p.per_batch ? 0 : zsel;
  const int split = blockIdx.z % p.splits;
  const int total_rb = p.per_batch ? p.row_blocks_per_batch : p.B * p.row_blocks_per_batch;
  const int rb_base = p.per_batch ? zsel * p.row_blocks_per_batch : 0;
  const int rb_begin = rb_base + static_cast<int>((static_cast<long long>(total_rb) * split) / p.splits);
  const int rb_end = rb_base + static_cast<int>((static_cast<long long>(total_rb) * (split + 1)) / p.splits);
  const int iters = rb_end - rb_begin;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], WG_STAGE_BYTES);
        const int rb = rb_begin + it;
        const int b = rb / p.row_blocks_per_batch;
        const int t0 = (rb - b * p.row_blocks_per_batch) * WG_ROWS;
        uint8_t* sDy = smem + s * WG_STAGE_BYTES;
        uint8_t* sA = sDy + WG_OP_BYTES;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(sDy + c * (WG_ROWS * ROW_BYTES), &tmDy, &full_bar[s], n0 + c * 64, t0, b);
          tma_load_3d(sA + c * (WG_ROWS * ROW_BYTES), &tmA, &full_bar[s], k0 + c * 64, t0 + tap - p.pad, b);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_instr_desc(OSB_F16, WG_TILE, WG_TILE, 1, 1);
      for (int it = 0; it < iters; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (it / WG_STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after_sync();
        const uint32_t dy_addr = smem_u32(smem + s * WG_STAGE_BYTES);
        const uint32_t a_addr = dy_addr + WG_OP_BYTES;
#pragma unroll
        for (int k = 0; k < WG_ROWS / 16; ++k) {
          // 16 contraction rows = 2 swizzle groups of 8 rows (1024 B each); chunks of 64 elems along MN are
          // WG_ROWS*128 B apart (LBO).
          const uint64_t da = make_smem_desc_sw128(dy_addr + k * 16 * ROW_BYTES, WG_ROWS * ROW_BYTES, 1024);
          const uint64_t db = make_smem_desc_sw128(a_addr + k * 16 * ROW_BYTES, WG_ROWS * ROW_BYTES, 1024);
          umma_ss<false>(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else if (iters > 0) {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // TMEM hands every thread one output row; the reduction into dw wants a warp on one row (512 contiguous bytes per
    // vector red instead of 32 scattered 4-byte atomics).  Transpose the warp's 32 x 128 slab through the (now idle) pipeline
    // stages: row stride 132 floats keeps both the float4 row writes and the float4 row reads conflict-free.
    constexpr int LD = WG_TILE + 4;
    float* slab = reinterpret_cast<float*>(smem) + q * 32 * LD;
    float v[32];
    for (int c0 = 0; c0 < WG_TILE; c0 += 32) {
      ld_chunk(taddr, c0, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(slab + lane * LD + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    __syncwarp();
    const int kc = k0 + lane * 4;
    if (kc < p.K) {  // K is a multiple of 8: whole float4 groups are in or out
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int n = n0 + q * 32 + r;
        if (n >= p.N) break;
        const float4 x = *reinterpret_cast<const float4*>(slab + r * LD + lane * 4);
        float* dst = p.dw + (static_cast<long long>(zsel) * p.N + n) * p.K + kc;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

template <int BN, int EPI>
int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  dim3 grid(p.B * p.m_tiles, (p.N + BN - 1) / BN, 1);
  gemm_nt_kernel<BN, EPI><<<grid, 192, Cfg::SMEM_BYTES, stream>>>(tmA, tmW, p);
  count_launch();
  return launch_status();
}

template <int EPI>
int dispatch_bn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_nt<64, EPI>(tmA, tmW, p, stream);
    case 128: return launch_nt<128, EPI>(tmA, tmW, p, stream);
    case 192: return launch_nt<192, EPI>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, EPI>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}
template <int EPI>
int dispatch_bn_full(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 128: return launch_nt<128, EPI>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, EPI>(tmA, tmW, p, stream);
    case 384: return launch_nt<384, EPI>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}

int dispatch_attn(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmKParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_nt<64, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 128: return launch_nt<128, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 192: return launch_nt<192, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 256: return launch_nt<256, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 320: return launch_nt<320, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 384: return launch_nt<384, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 448: return launch_nt<448, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    case 512: return launch_nt<512, OSB_EPI_ATTN_LOGP>(tmA, tmW, p, stream);
    default: return OSB_ERR_SHAPE;
  }
}

// Largest supported tile width that divides N.
int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// TMA tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_tmap_3d(CUtensorMap* out, const void* base, TmaDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return OSB_ERR_DRIVER;
  const uint64_t es = (dt == TMA_F32) ? 4 : 2;
  const CUtensorMapDataType cdt = dt == TMA_F16    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                  : dt == TMA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                   : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return OSB_ERR_ALIGN;
  if (((stride1_elems * es) & 15) != 0 || ((stride2_elems * es) & 15) != 0) return OSB_ERR_ALIGN;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * es, stride2_elems * es};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, cdt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? OSB_OK : OSB_ERR_DRIVER;
}

}  // namespace osb

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace osb;

extern "C" int osb_gemm(const osb_gemm_desc* d, void* stream_) {
  OSB_REQUIRE(d != nullptr && d->a != nullptr && d->w != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(d->B > 0 && d->T > 0 && d->N > 0 && d->K > 0 && d->taps > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(d->lda % 8 == 0 && d->ldw % 8 == 0 && (d->ldo % 8 == 0 || d->epi == OSB_EPI_ATTN_LOGP), OSB_ERR_ALIGN);
  OSB_REQUIRE(d->lda >= d->K && d->ldw >= d->K && d->ldo >= d->N, OSB_ERR_SHAPE);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  const bool full_row = (d->epi == OSB_EPI_RELU_LN || d->epi == OSB_EPI_BIAS_LN || d->epi == OSB_EPI_LN_BWD ||
                         d->epi == OSB_EPI_RELU_LN_BWD);
  const bool attn = d->epi == OSB_EPI_ATTN_LOGP;
  const int bn = attn ? ((d->N + 63) / 64) * 64 : (full_row ? d->N : pick_bn(d->N));
  OSB_REQUIRE(bn > 0 && bn <= 512, OSB_ERR_SHAPE);
  const int ninst = bn <= 256 ? bn : bn / 2;
  const bool w_batched = d->w_batched != 0;
  if (w_batched) OSB_REQUIRE(d->taps == 1, OSB_ERR_SHAPE);

  const bool split_in = (d->flags & OSB_FLAG_SPLIT_IN) != 0;
  if (split_in) OSB_REQUIRE(d->K % BKE == 0 && d->lda >= 2 * static_cast<int64_t>(d->K), OSB_ERR_SHAPE);
  CUtensorMap tmA, tmW;
  int rc = make_tmap_3d(&tmA, d->a, TMA_F16, split_in ? 2 * d->K : d->K, d->T, d->B, d->lda, static_cast<uint64_t>(d->T) * d->lda,
                        BKE, BM);
  if (rc != OSB_OK) return rc;
  if (w_batched) {
    if (split_in) OSB_REQUIRE(d->ldw >= 2 * static_cast<int64_t>(d->K), OSB_ERR_SHAPE);
    rc = make_tmap_3d(&tmW, d->w, TMA_F16, split_in ? 2 * d->K : d->K, d->N, d->B, d->ldw, static_cast<uint64_t>(d->N) * d->ldw, BKE,
                      ninst);
  } else {
    rc = make_tmap_3d(&tmW, d->w, TMA_F16, d->K, d->N, split_in ? 2 * d->taps : d->taps, d->ldw, static_cast<uint64_t>(d->N) * d->ldw,
                      BKE, ninst);
  }
  if (rc != OSB_OK) return rc;

  GemmKParams p;
  p.T = d->T; p.B = d->B; p.N = d->N; p.K = d->K; p.taps = d->taps; p.pad = d->pad;
  p.m_tiles = (d->T + BM - 1) / BM;
  p.flags = d->flags;
  p.out = d->out; p.aux = d->aux_h16; p.ldo = d->ldo;
  p.bias = d->bias; p.resid = d->resid; p.gamma = d->gamma; p.row_scale = d->row_scale;
  p.pad_mask = d->pad_mask; p.ln_w = d->ln_w; p.ln_b = d->ln_b; p.ln_eps = d->ln_eps;
  p.dot_w = d->dot_w; p.dot_b = d->dot_b; p.out_dot = d->out_dot;
  p.aux_in = d->aux_in_h16; p.row_stat = d->row_stat;
  p.w_batch = w_batched ? 1 : 0;
  p.w_lo_slice = w_batched ? 0 : d->taps;
  p.w_lo_koff = w_batched ? d->K : 0;
  p.col_len = reinterpret_cast<const long long*>(d->col_len);
  p.drop_p = d->dropout_p; p.drop_seed = d->dropout_seed; p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(d->dropout_seed_dev);
  p.drop_inv_keep = d->dropout_p > 0.f && d->dropout_p < 1.f ? 1.f / (1.f - d->dropout_p) : 0.f;
  OSB_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f, OSB_ERR_ARG);

  if ((d->flags & (OSB_FLAG_OUT_H16 | OSB_FLAG_SAVE_PRE)) && d->aux_h16 == nullptr) return OSB_ERR_ARG;
  if ((d->flags & OSB_FLAG_KEEPMASK) && d->pad_mask == nullptr) return OSB_ERR_ARG;

  switch (d->epi) {
    case OSB_EPI_BIAS:
      OSB_REQUIRE(d->out != nullptr || ((d->flags & OSB_FLAG_NO_F32) && (d->flags & OSB_FLAG_OUT_H16)), OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_BIAS>(bn, tmA, tmW, p, stream);
    case OSB_EPI_GELU:
      OSB_REQUIRE(d->out != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_GELU>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU:
      OSB_REQUIRE(d->out != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RELU>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RESID:
      OSB_REQUIRE(d->out != nullptr && d->resid != nullptr && d->gamma != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RESID>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_LN:
      OSB_REQUIRE(d->ln_w != nullptr && d->ln_b != nullptr, OSB_ERR_ARG);
      OSB_REQUIRE(!(d->flags & OSB_FLAG_DOT) || (d->dot_w != nullptr && d->out_dot != nullptr), OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_RELU_LN>(bn, tmA, tmW, p, stream);
    case OSB_EPI_BIAS_LN:
      OSB_REQUIRE(d->out != nullptr && d->ln_w != nullptr && d->ln_b != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_BIAS_LN>(bn, tmA, tmW, p, stream);
    case OSB_EPI_GELU_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_GELU_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_RELU_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_ATTN_LOGP: {
      OSB_REQUIRE(d->out != nullptr && d->row_stat != nullptr && d->bias != nullptr && d->resid != nullptr && d->col_len != nullptr,
                  OSB_ERR_ARG);
      GemmKParams q = p;
      q.N = d->N;  // true column count; the tile is padded to a multiple of 64 (out-of-range rows of W read as zero)
      dim3 grid_check(1, 1, 1);
      (void)grid_check;
      return dispatch_attn(bn, tmA, tmW, q, stream);
    }
    case OSB_EPI_AXPY:
      OSB_REQUIRE(d->out != nullptr && d->row_stat != nullptr && d->resid != nullptr, OSB_ERR_ARG);
      return dispatch_bn<OSB_EPI_AXPY>(bn, tmA, tmW, p, stream);
    case OSB_EPI_LN_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr && d->row_stat != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_LN_BWD>(bn, tmA, tmW, p, stream);
    case OSB_EPI_RELU_LN_BWD:
      OSB_REQUIRE(d->out != nullptr && d->aux_in_h16 != nullptr && d->ln_w != nullptr, OSB_ERR_ARG);
      return dispatch_bn_full<OSB_EPI_RELU_LN_BWD>(bn, tmA, tmW, p, stream);
    default:
      return OSB_ERR_ARG;
  }
}

static int wgrad_impl(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t N, int32_t K,
                      int32_t taps, int32_t pad, int per_batch, void* stream_);

extern "C" int osb_gemm_wgrad(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                              int32_t N, int32_t K, int32_t taps, int32_t pad, void* stream_) {
  return wgrad_impl(dy, ldy, a, lda, dw, B, T, N, K, taps, pad, 0, stream_);
}

extern "C" int osb_gemm_wgrad_batched(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                                      int32_t N, int32_t K, void* stream_) {
  return wgrad_impl(dy, ldy, a, lda, dw, B, T, N, K, 1, 0, 1, stream_);
}

static int wgrad_impl(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t N, int32_t K,
                      int32_t taps, int32_t pad, int per_batch, void* stream_) {
  OSB_REQUIRE(dy != nullptr && a != nullptr && dw != nullptr, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && N > 0 && K > 0 && taps > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(ldy % 8 == 0 && lda % 8 == 0, OSB_ERR_ALIGN);
  // the tile is reduced into dw with 16-byte vector reds
  OSB_REQUIRE(K % 4 == 0 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0, OSB_ERR_ALIGN);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tmDy, tmA;
  int rc = make_tmap_3d(&tmDy, dy, TMA_F16, N, T, B, ldy, static_cast<uint64_t>(T) * ldy, 64, WG_ROWS);
  if (rc != OSB_OK) return rc;
  rc = make_tmap_3d(&tmA, a, TMA_F16, K, T, B, lda, static_cast<uint64_t>(T) * lda, 64, WG_ROWS);
  if (rc != OSB_OK) return rc;

  WgradParams p;
  p.T = T; p.B = B; p.N = N; p.K = K; p.taps = taps; p.pad = pad;
  p.row_blocks_per_batch = (T + WG_ROWS - 1) / WG_ROWS;
  const int n_tiles = (N + WG_TILE - 1) / WG_TILE;
  const int k_tiles = (K + WG_TILE - 1) / WG_TILE;
  p.per_batch = per_batch;
  const int zcount = per_batch ? B : taps;
  const int total_rb = per_batch ? p.row_blocks_per_batch : B * p.row_blocks_per_batch;
  // Split the contraction rows over CTAs to fill the machine, but keep >= 6 pipeline iterations per CTA: every CTA pays a
  // fixed prologue and flushes a 128x128 fp32 tile with atomics, so over-splitting small problems costs more than it gains.
  int splits = (148 + n_tiles * k_tiles * zcount - 1) / (n_tiles * k_tiles * zcount);
  if (splits > total_rb / 6) splits = total_rb / 6;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.dw = dw;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  dim3 grid(n_tiles, k_tiles, zcount * splits);
  gemm_wgrad_kernel<<<grid, 192, WG_SMEM_BYTES, stream>>>(tmDy, tmA, p);
  count_launch();
  return launch_status();
}

extern "C" int osb_version(void) { return 1; }
extern "C" unsigned long long osb_launch_count(void) { return osb::g_launch_count; }
extern "C" const char* osb_strerror(int s) {
  switch (s) {
    case OSB_OK: return "ok";
    case OSB_ERR_SHAPE: return "unsupported or inconsistent shape";
    case OSB_ERR_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case OSB_ERR_ARCH: return "device is not sm_100 (B200)";
    case OSB_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case OSB_ERR_WORKSPACE: return "workspace too small";
    case OSB_ERR_ARG: return "null pointer or bad enum";
    default: return s > 0 ? "CUDA runtime error (see cudaGetErrorString)" : "unknown status";
  }
}
extern "C" int osb_check_device(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return static_cast<int>(e);
  return prop.major == 10 ? OSB_OK : OSB_ERR_ARCH;
}
