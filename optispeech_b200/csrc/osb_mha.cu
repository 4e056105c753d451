// osb_mha.cu — fused multi-head self-attention for the Transformer backbone (configs/model/transformer.yaml: 2 heads, d_k = 128).
//
// Replaces MultiHeadedAttention.forward_attention + the QK^T matmul of .forward
// (optispeech/model/generator/modules/_transformer/attention.py:84-125): softmax(masked(QK^T / sqrt(d_k))) V with keys
// beyond the sequence length masked, attention dropout on the probabilities.  The (B, h, T, T) score / probability tensors
// of the reference never reach HBM in the forward pass.
//
// mha_fwd_kernel — a CTA owns 128 query rows of one (batch, head).  Two sweeps over the key blocks (128 keys each):
//   sweep 1: S = Q K^T on the tensor cores (tcgen05, fp32 in TMEM, double buffered), worker threads (one per query row)
//            take the exact row maximum;
//   sweep 2: S again, p = exp2((s - max) * scale * log2 e) -> fp16 tile in 128B-swizzled shared memory = A operand of
//            O += P V (V consumed MN-major straight from its channels-last layout), row sums in fp32; O stays in TMEM over
//            all key blocks and is never rescaled (the maximum is already exact), so the result is the reference's softmax
//            up to fp16 operand rounding.
//   K / V tiles are TMA-streamed through full/empty mbarrier rings by one producer thread; one thread issues the MMAs.
// mha_bwd_kernel — same tiling, one sweep: S and dP = dO V^T on the tensor cores, dS = P o (dP o D - delta) * scale in the
//   epilogue (P recomputed from the saved row max / 1/sum, dropout mask D regenerated from the counter-based RNG), dS tile
//   -> shared memory -> dQ += dS K (the K tile is re-read MN-major from the same shared-memory bytes).  dS and P o D are
//   also written to HBM as fp16: dK = dS^T Q and dV = (P o D)^T dO are contractions over the query rows and run on
//   gemm_wgrad_kernel (both operands MN-major).
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

constexpr int MH_M = 128;        // query rows per CTA
constexpr int MH_N = 128;        // keys per block
constexpr int MH_D = 128;        // head dimension (d_k)
constexpr int MH_TILE = 128 * 256;  // bytes of one [128 x 128] fp16 operand tile (two 64-element k-blocks of [128 rows x 128 B])
constexpr int MH_KB = 128 * 128;    // bytes of one k-block
constexpr int MH_THREADS = 64 + 128;
constexpr float LOG2E = 1.4426950408889634f;

struct MhaParams {
  const long long* kv_len;   // (B) or null
  __half* ctx;               // fwd: (B, T, ld_ctx) output ; bwd: the forward output (for delta)
  long long ld_ctx;
  long long ctx_lo_off;      // fwd: > 0 -> also write the fp16 rounding residual at column offset ctx_lo_off (split precision)
  float* row_max;            // (B, H, T) raw (unscaled) row maximum
  float* row_inv_l;          // (B, H, T) 1 / sum exp
  // backward
  const __half* d_ctx;       // (B, T, ld_dctx)
  long long ld_dctx;
  __half* dq;                // (B, T, ld_dq) head h at columns [h*128, +128)
  long long ld_dq;
  __half* ds_out;            // (B, T, ld_p) head h at columns [h*Tp, +Tp): dS
  __half* pd_out;            // (B, T, ld_p): P o D
  long long ld_p;
  int Tp;
  int B, T, H;
  float scale;
  float drop_p, drop_inv_keep;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;
};

__device__ __forceinline__ uint32_t sw128(int r, int c16) { return static_cast<uint32_t>(r * 128 + ((c16 ^ (r & 7)) << 4)); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 h0 = __floats2half2_rn(v[0], v[1]);
  __half2 h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]);
  __half2 h3 = __floats2half2_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One [128 rows x 128 cols] fp16 tile = two TMA boxes of 64 columns.
__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int col0, int row0, int b) {
  tma_load_3d(dst, tm, bar, col0, row0, b);
  tma_load_3d(dst + MH_KB, tm, bar, col0 + 64, row0, b);
}

// D (+)= A[128 x 128, K-major] . B[128 x 128, K-major]^T : 8 MMAs of K = 16
__device__ __forceinline__ void mma_nt_128(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate_first) {
  constexpr uint32_t idesc = make_instr_desc(OSB_F16, MH_M, MH_N, 0, 0);
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_smem_desc_sw128(a_addr + kb * MH_KB + k * 32, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(b_addr + kb * MH_KB + k * 32, 16, 1024);
      umma_ss<false>(d_tmem, da, db, idesc, (accumulate_first || (kb | k) != 0) ? 1u : 0u);
    }
}
// D (+)= A[128 x 128 keys, K-major] . B[128 keys x 128, MN-major] : the B tile is (keys x d) with d contiguous
__device__ __forceinline__ void mma_nn_128(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, bool accumulate_first) {
  constexpr uint32_t idesc = make_instr_desc(OSB_F16, MH_M, MH_D, 0, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint64_t da = make_smem_desc_sw128(a_addr + (k >> 2) * MH_KB + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_smem_desc_sw128(b_addr + k * 16 * 128, MH_KB, 1024);
    umma_ss<false>(d_tmem, da, db, idesc, (accumulate_first || k != 0) ? 1u : 0u);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
constexpr int MHF_SMEM = MH_TILE * (1 + 2 + 2 + 2) + 1024 + 256;

__global__ void __launch_bounds__(MH_THREADS, 1)
mha_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const MhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + MH_TILE;        // [2]
  uint8_t* sV = sK + 2 * MH_TILE;    // [2]
  uint8_t* sP = sV + 2 * MH_TILE;    // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * MH_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* s_empty = bars + 11; // [2]
  uint64_t* p_full = bars + 13;  // [2]
  uint64_t* p_empty = bars + 15; // [2]
  uint64_t* o_full = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * MH_M;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  int kv = p.kv_len != nullptr ? static_cast<int>(p.kv_len[b]) : p.T;
  kv = kv < 0 ? 0 : (kv > p.T ? p.T : kv);
  const int nblk = kv > 0 ? (kv + MH_N - 1) / MH_N : 1;   // kv == 0: one fully masked block -> zero output

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t O_COL = 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, MH_TILE);
      load_tile(sQ, &tmQ, q_full, h * MH_D, t0, b);
      for (int s = 0; s < 2 * nblk; ++s) {
        const int j = s < nblk ? s : s - nblk;
        const int st = s & 1;
        mbar_wait(&k_empty[st], ((s >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], MH_TILE);
        load_tile(sK + st * MH_TILE, &tmK, &k_full[st], h * MH_D, j * MH_N, b);
        if (s >= nblk) {
          const int vs = j & 1;
          mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
          mbar_expect_tx(&v_full[vs], MH_TILE);
          load_tile(sV + vs * MH_TILE, &tmV, &v_full[vs], h * MH_D, j * MH_N, b);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
    mbar_wait(q_full, 0);
    tc_fence_after_sync();
    auto issue_pv = [&](int j) {
      const int buf = j & 1;
      mbar_wait(&v_full[buf], (j >> 1) & 1);
      mbar_wait(&p_full[buf], (j >> 1) & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_nn_128(tmem_base + O_COL, p_addr + buf * MH_TILE, v_addr + buf * MH_TILE, j != 0);
        umma_commit(&v_empty[buf]);
        umma_commit(&p_empty[buf]);
      }
      __syncwarp();
    };
    for (int s = 0; s < 2 * nblk; ++s) {
      const int st = s & 1;
      mbar_wait(&k_full[st], (s >> 1) & 1);
      mbar_wait(&s_empty[st], ((s >> 1) & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_nt_128(tmem_base + st * MH_N, q_addr, k_addr + st * MH_TILE, false);
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      }
      __syncwarp();
      if (s > nblk) issue_pv(s - 1 - nblk);
    }
    issue_pv(nblk - 1);
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale * LOG2E;
    float v[32];
    // ---- sweep 1: exact row maximum ----
    float m = -INFINITY;
    for (int s = 0; s < nblk; ++s) {
      const int st = s & 1;
      mbar_wait(&s_full[st], (s >> 1) & 1);
      tc_fence_after_sync();
      const int key0 = s * MH_N;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        ld32(lane_addr + st * MH_N + cc * 32, v);
        if (key0 + cc * 32 + 32 <= kv) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = (key0 + cc * 32 + i < kv) ? fmaxf(m, v[i]) : m;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
    }
    const float mc = (m == -INFINITY) ? 0.f : m * c;
    // ---- sweep 2: probabilities -> shared memory, row sums ----
    unsigned long long seed = p.drop_seed;
    if (p.drop_p > 0.f && p.drop_seed_dev != nullptr) seed += *p.drop_seed_dev;
    const unsigned long long drow = (static_cast<unsigned long long>(b * p.H + h) * p.T + (t < p.T ? t : 0)) * p.T;
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int s = nblk + j;
      const int st = s & 1;
      const int buf = j & 1;
      mbar_wait(&s_full[st], (s >> 1) & 1);
      tc_fence_after_sync();
      mbar_wait(&p_empty[buf], ((j >> 1) & 1) ^ 1);
      const int key0 = j * MH_N;
      uint8_t* prow = sP + buf * MH_TILE;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        ld32(lane_addr + st * MH_N + cc * 32, v);
        const int kc = key0 + cc * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = ex2(fmaf(v[i], c, -mc));
          v[i] = (kc + i < kv) ? e : 0.f;
          l += v[i];
        }
        if (p.drop_p > 0.f) dropout_apply(v, seed, drow + kc, p.drop_p, p.drop_inv_keep);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          *reinterpret_cast<uint4*>(prow + (cc >> 1) * MH_KB + sw128(row, (cc & 1) * 4 + c4)) = pack8(v + c4 * 8);
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[st]);
        mbar_arrive(&p_full[buf]);
      }
    }
    // ---- epilogue: O / l -> fp16 context ----
    const float inv_l = l > 0.f ? 1.f / l : 0.f;
    mbar_wait(o_full, 0);
    tc_fence_after_sync();
    const long long grow = static_cast<long long>(b) * p.T + t;
    if (t < p.T) {
      if (p.row_max != nullptr) {
        const long long sidx = (static_cast<long long>(b) * p.H + h) * p.T + t;
        p.row_max[sidx] = m;
        p.row_inv_l[sidx] = inv_l;
      }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      ld32(lane_addr + O_COL + cc * 32, v);
      if (t < p.T) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= inv_l;
        __half* dst = p.ctx + grow * p.ld_ctx + h * MH_D + cc * 32;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) *reinterpret_cast<uint4*>(dst + c4 * 8) = pack8(v + c4 * 8);
        if (p.ctx_lo_off > 0) {
          float r[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = v[i] - __half2float(__float2half_rn(v[i]));
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) *reinterpret_cast<uint4*>(dst + p.ctx_lo_off + c4 * 8) = pack8(r + c4 * 8);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
constexpr int MHB_SMEM = MH_TILE * (1 + 1 + 2 + 2 + 1) + 1024 + 256;

__global__ void __launch_bounds__(MH_THREADS, 1)
mha_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmDO, const MhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + MH_TILE;
  uint8_t* sK = sDO + MH_TILE;       // [2]
  uint8_t* sV = sK + 2 * MH_TILE;    // [2]
  uint8_t* sDS = sV + 2 * MH_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + MH_TILE);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* ds_empty = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * MH_M;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  int kv = p.kv_len != nullptr ? static_cast<int>(p.kv_len[b]) : p.T;
  kv = kv < 0 ? 0 : (kv > p.T ? p.T : kv);
  const int nblk = (p.T + MH_N - 1) / MH_N;   // every key block is visited: dS / P o D are written (zeros) for masked keys too

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(ds_full, 4); mbar_init(ds_empty, 1); mbar_init(dq_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t S_COL = 0, DP_COL = 128, DQ_COL = 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * MH_TILE);
      load_tile(sQ, &tmQ, q_full, h * MH_D, t0, b);
      load_tile(sDO, &tmDO, q_full, h * MH_D, t0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * MH_TILE);
        load_tile(sK + st * MH_TILE, &tmK, &kv_full[st], h * MH_D, j * MH_N, b);
        load_tile(sV + st * MH_TILE, &tmV, &kv_full[st], h * MH_D, j * MH_N, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK), v_addr = smem_u32(sV), ds_addr = smem_u32(sDS);
    mbar_wait(q_full, 0);
    tc_fence_after_sync();
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      mbar_wait(s_empty, (j & 1) ^ 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_nt_128(tmem_base + S_COL, q_addr, k_addr + st * MH_TILE, false);     // S  = Q K^T
        mma_nt_128(tmem_base + DP_COL, do_addr, v_addr + st * MH_TILE, false);   // dP = dO V^T
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(ds_full, j & 1);
      tc_fence_after_sync();
      if (elect_one()) {
        mma_nn_128(tmem_base + DQ_COL, ds_addr, k_addr + st * MH_TILE, j != 0);  // dQ += dS K
        umma_commit(&kv_empty[st]);
        umma_commit(ds_empty);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(dq_full);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool valid = t < p.T;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c = p.scale * LOG2E;
    const long long grow = static_cast<long long>(b) * p.T + (valid ? t : 0);
    const long long sidx = (static_cast<long long>(b) * p.H + h) * p.T + (valid ? t : 0);
    const float m = p.row_max[sidx];
    const float inv_l = valid ? p.row_inv_l[sidx] : 0.f;
    const float mc = (m == -INFINITY) ? 0.f : m * c;
    // delta = <dO_row, O_row>  (= sum_k dP_k (P o D)_k)
    float delta = 0.f;
    if (valid) {
      const __half* o = p.ctx + grow * p.ld_ctx + h * MH_D;
      const __half* g = p.d_ctx + grow * p.ld_dctx + h * MH_D;
#pragma unroll 4
      for (int i = 0; i < MH_D; i += 8) {
        const uint4 uo = *reinterpret_cast<const uint4*>(o + i);
        const uint4 ug = *reinterpret_cast<const uint4*>(g + i);
        const __half2* ho = reinterpret_cast<const __half2*>(&uo);
        const __half2* hg = reinterpret_cast<const __half2*>(&ug);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 a = __half22float2(ho[k]), d = __half22float2(hg[k]);
          delta = fmaf(a.x, d.x, fmaf(a.y, d.y, delta));
        }
      }
    }
    unsigned long long seed = p.drop_seed;
    if (p.drop_p > 0.f && p.drop_seed_dev != nullptr) seed += *p.drop_seed_dev;
    const unsigned long long drow = (static_cast<unsigned long long>(b * p.H + h) * p.T + (valid ? t : 0)) * p.T;
    float sv[32], dv[32];
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after_sync();
      mbar_wait(ds_empty, (j & 1) ^ 1);
      const int key0 = j * MH_N;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        ld32(lane_addr + S_COL + cc * 32, sv);
        ld32(lane_addr + DP_COL + cc * 32, dv);
        const int kc = key0 + cc * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float pr = (kc + i < kv) ? ex2(fmaf(sv[i], c, -mc)) * inv_l : 0.f;
          const float d = p.drop_p > 0.f ? dropout_scale(seed, drow + kc + i, p.drop_p, p.drop_inv_keep) : 1.f;
          sv[i] = pr * d;                                      // (P o D): the dV operand
          dv[i] = pr * (dv[i] * d - delta) * p.scale;          // dS
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          *reinterpret_cast<uint4*>(sDS + (cc >> 1) * MH_KB + sw128(row, (cc & 1) * 4 + c4)) = pack8(dv + c4 * 8);
        if (valid) {
          __half* dsr = p.ds_out + grow * p.ld_p + static_cast<long long>(h) * p.Tp + kc;
          __half* pdr = p.pd_out + grow * p.ld_p + static_cast<long long>(h) * p.Tp + kc;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            if (kc + c4 * 8 < p.Tp) {
              *reinterpret_cast<uint4*>(dsr + c4 * 8) = pack8(dv + c4 * 8);
              *reinterpret_cast<uint4*>(pdr + c4 * 8) = pack8(sv + c4 * 8);
            }
          }
        }
      }
      tc_fence_before_sync();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(ds_full);
      }
    }
    mbar_wait(dq_full, 0);
    tc_fence_after_sync();
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      ld32(lane_addr + DQ_COL + cc * 32, sv);
      if (valid) {
        __half* dst = p.dq + grow * p.ld_dq + h * MH_D + cc * 32;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) *reinterpret_cast<uint4*>(dst + c4 * 8) = pack8(sv + c4 * 8);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dqkv[b,t, off + h*128 + d] = fp16(src[h][b][t][d])   (src: (H, B, T, 128) fp32 from the per-head batched wgrad contractions)
__global__ void mha_pack_heads_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long ld_dst, int B, int T, int H) {
  const long long n4 = static_cast<long long>(H) * B * T * (MH_D / 4);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d4 = static_cast<int>(i % (MH_D / 4));
    const long long r = i / (MH_D / 4);
    const long long bt = r % (static_cast<long long>(B) * T);
    const int h = static_cast<int>(r / (static_cast<long long>(B) * T));
    const float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(dst + bt * ld_dst + h * MH_D + d4 * 4) = u;
  }
}

// dst[i] = fp16(src[i] * D(i)) over rows x N elements (4 per thread): the gradient entering a dropped-out branch
__global__ void dropout_pack_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n4, float drop_p, float inv_keep,
                                    unsigned long long seed, const unsigned long long* seed_dev) {
  if (drop_p > 0.f && seed_dev != nullptr) seed += *seed_dev;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = *reinterpret_cast<const float4*>(src + i * 4);
    if (drop_p > 0.f) {
      v.x *= dropout_scale(seed, i * 4 + 0, drop_p, inv_keep); v.y *= dropout_scale(seed, i * 4 + 1, drop_p, inv_keep);
      v.z *= dropout_scale(seed, i * 4 + 2, drop_p, inv_keep); v.w *= dropout_scale(seed, i * 4 + 3, drop_p, inv_keep);
    }
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(dst + i * 4) = u;
  }
}

// out[b,t,:] = (x[b,t,:] + alpha * pe[t,:]) * D   (ScaledPositionalEncoding.forward + its dropout, embedding.py:111-124)
__global__ void add_posenc_kernel(const float* __restrict__ x, const float* __restrict__ pe, const float* __restrict__ alpha,
                                  float* __restrict__ out, long long n4, int T, int C4, float drop_p, float inv_keep,
                                  unsigned long long seed, const unsigned long long* seed_dev) {
  if (drop_p > 0.f && seed_dev != nullptr) seed += *seed_dev;
  const float a = alpha != nullptr ? *alpha : 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % C4);
    const int t = static_cast<int>((i / C4) % T);
    float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    if (alpha != nullptr) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(pe + (static_cast<long long>(t) * C4 + c4) * 4));
      v.x = fmaf(a, e.x, v.x); v.y = fmaf(a, e.y, v.y); v.z = fmaf(a, e.z, v.z); v.w = fmaf(a, e.w, v.w);
    }
    if (drop_p > 0.f) {
      v.x *= dropout_scale(seed, i * 4 + 0, drop_p, inv_keep); v.y *= dropout_scale(seed, i * 4 + 1, drop_p, inv_keep);
      v.z *= dropout_scale(seed, i * 4 + 2, drop_p, inv_keep); v.w *= dropout_scale(seed, i * 4 + 3, drop_p, inv_keep);
    }
    *reinterpret_cast<float4*>(out + i * 4) = v;
  }
}

int make_head_map(CUtensorMap* tm, const void* base, int64_t ld, int B, int T, int H) {
  return make_tmap_3d(tm, base, TMA_F16, static_cast<uint64_t>(H) * MH_D, T, B, ld, static_cast<uint64_t>(T) * ld, 64, 128);
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_mha_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const int64_t* kv_len, void* ctx, int64_t ld_ctx,
                           int64_t ctx_lo_off, float* row_max, float* row_inv_l, int32_t B, int32_t T, int32_t H, int32_t d_k, float scale,
                           float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream) {
  OSB_REQUIRE(q && k && v && ctx, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && H > 0 && d_k == MH_D, OSB_ERR_SHAPE);
  OSB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && ctx_lo_off % 8 == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, OSB_ERR_ALIGN);
  OSB_REQUIRE((row_max == nullptr) == (row_inv_l == nullptr), OSB_ERR_ARG);
  OSB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, OSB_ERR_ARG);
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_head_map(&tmQ, q, ld_qkv, B, T, H);
  if (rc != OSB_OK) return rc;
  if ((rc = make_head_map(&tmK, k, ld_qkv, B, T, H)) != OSB_OK) return rc;
  if ((rc = make_head_map(&tmV, v, ld_qkv, B, T, H)) != OSB_OK) return rc;
  MhaParams p{};
  p.kv_len = reinterpret_cast<const long long*>(kv_len);
  p.ctx = static_cast<__half*>(ctx); p.ld_ctx = ld_ctx; p.ctx_lo_off = ctx_lo_off;
  p.row_max = row_max; p.row_inv_l = row_inv_l;
  p.B = B; p.T = T; p.H = H; p.scale = scale;
  p.drop_p = dropout_p; p.drop_inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_seed = dropout_seed; p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(dropout_seed_dev);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MHF_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  dim3 grid((T + MH_M - 1) / MH_M, H, B);
  mha_fwd_kernel<<<grid, MH_THREADS, MHF_SMEM, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  count_launch();
  return launch_status();
}

extern "C" int osb_mha_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const int64_t* kv_len, const void* ctx,
                           int64_t ld_ctx, const void* d_ctx, int64_t ld_dctx, const float* row_max, const float* row_inv_l, void* dq,
                           int64_t ld_dq, void* ds_out, void* pd_out, int64_t ld_p, int32_t Tp, int32_t B, int32_t T, int32_t H,
                           int32_t d_k, float scale, float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev,
                           void* stream) {
  OSB_REQUIRE(q && k && v && ctx && d_ctx && row_max && row_inv_l && dq && ds_out && pd_out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && H > 0 && d_k == MH_D && Tp >= T && ld_p >= static_cast<int64_t>(H) * Tp, OSB_ERR_SHAPE);
  OSB_REQUIRE(ld_qkv % 8 == 0 && ld_ctx % 8 == 0 && ld_dctx % 8 == 0 && ld_dq % 8 == 0 && ld_p % 8 == 0 && Tp % 8 == 0, OSB_ERR_ALIGN);
  OSB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, OSB_ERR_ARG);
  CUtensorMap tmQ, tmK, tmV, tmDO;
  int rc = make_head_map(&tmQ, q, ld_qkv, B, T, H);
  if (rc != OSB_OK) return rc;
  if ((rc = make_head_map(&tmK, k, ld_qkv, B, T, H)) != OSB_OK) return rc;
  if ((rc = make_head_map(&tmV, v, ld_qkv, B, T, H)) != OSB_OK) return rc;
  if ((rc = make_head_map(&tmDO, d_ctx, ld_dctx, B, T, H)) != OSB_OK) return rc;
  MhaParams p{};
  p.kv_len = reinterpret_cast<const long long*>(kv_len);
  p.ctx = const_cast<__half*>(static_cast<const __half*>(ctx)); p.ld_ctx = ld_ctx;
  p.row_max = const_cast<float*>(row_max); p.row_inv_l = const_cast<float*>(row_inv_l);
  p.d_ctx = static_cast<const __half*>(d_ctx); p.ld_dctx = ld_dctx;
  p.dq = static_cast<__half*>(dq); p.ld_dq = ld_dq;
  p.ds_out = static_cast<__half*>(ds_out); p.pd_out = static_cast<__half*>(pd_out); p.ld_p = ld_p; p.Tp = Tp;
  p.B = B; p.T = T; p.H = H; p.scale = scale;
  p.drop_p = dropout_p; p.drop_inv_keep = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f;
  p.drop_seed = dropout_seed; p.drop_seed_dev = reinterpret_cast<const unsigned long long*>(dropout_seed_dev);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MHB_SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr = true;
  }
  dim3 grid((T + MH_M - 1) / MH_M, H, B);
  mha_bwd_kernel<<<grid, MH_THREADS, MHB_SMEM, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, tmDO, p);
  count_launch();
  return launch_status();
}

extern "C" int osb_mha_pack_heads(const float* src, void* dst_h16, int64_t ld_dst, int32_t B, int32_t T, int32_t H, void* stream) {
  OSB_REQUIRE(src && dst_h16, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && H > 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(ld_dst % 4 == 0 && (reinterpret_cast<uintptr_t>(dst_h16) & 7) == 0, OSB_ERR_ALIGN);
  const long long n4 = static_cast<long long>(H) * B * T * (MH_D / 4);
  const int blocks = static_cast<int>((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  mha_pack_heads_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__half*>(dst_h16), ld_dst, B, T, H);
  count_launch();
  return launch_status();
}

extern "C" int osb_dropout_pack_h16(const float* src, void* dst_h16, int64_t rows, int32_t N, float dropout_p, uint64_t dropout_seed,
                                    const uint64_t* dropout_seed_dev, void* stream) {
  OSB_REQUIRE(src && dst_h16, OSB_ERR_ARG);
  OSB_REQUIRE(rows > 0 && N > 0 && N % 4 == 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, OSB_ERR_ARG);
  const long long n4 = rows * N / 4;
  const int blocks = static_cast<int>((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  dropout_pack_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__half*>(dst_h16), n4, dropout_p,
                                                                            dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f, dropout_seed,
                                                                            reinterpret_cast<const unsigned long long*>(dropout_seed_dev));
  count_launch();
  return launch_status();
}

extern "C" int osb_add_posenc(const float* x, const float* pe, const float* alpha, float* out, int32_t B, int32_t T, int32_t C,
                              float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream) {
  OSB_REQUIRE(x && out && ((pe != nullptr) == (alpha != nullptr)), OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0 && C % 4 == 0, OSB_ERR_SHAPE);
  OSB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, OSB_ERR_ARG);
  const long long n4 = static_cast<long long>(B) * T * C / 4;
  const int blocks = static_cast<int>((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
  add_posenc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, pe, alpha, out, n4, T, C / 4, dropout_p,
                                                                          dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 1.f, dropout_seed,
                                                                          reinterpret_cast<const unsigned long long*>(dropout_seed_dev));
  count_launch();
  return launch_status();
}
