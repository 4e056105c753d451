// osb_ptx.cuh — thin inline-PTX wrappers for sm_100a (Blackwell B200).
//
// Everything the kernels in this directory need from the Blackwell execution
// model lives here: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / fences) and the shared-memory + instruction descriptors.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" and
// "instruction descriptor" tables (same layout NVIDIA's CuTe exposes as
// UMMA::SmemDescriptor / UMMA::InstrDescriptor).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace osb {

// ----------------------------------------------------------------------------
// misc
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a mis-programmed pipeline traps instead of hanging the GPU box.
#ifndef OSB_MBAR_SPIN_LIMIT
#define OSB_MBAR_SPIN_LIMIT (1u << 28)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > OSB_MBAR_SPIN_LIMIT) {
      asm volatile("trap;");
    }
  }
}

// ----------------------------------------------------------------------------
// proxies / fences
// ----------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA store, UMMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ----------------------------------------------------------------------------
// TMA loads (tile mode). Coordinates are innermost-first, in elements.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "TMEM cols: pow2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(kCols) : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05: descriptors
// ----------------------------------------------------------------------------
enum : uint32_t { OSB_F16 = 0, OSB_BF16 = 1, OSB_TF32 = 2 };

// Shared-memory matrix descriptor, 128-byte swizzle (layout_type = 2), version 1.
//   bits [0,14)  start address >> 4
//   bits [16,30) leading-dimension byte offset >> 4
//   bits [32,46) stride-dimension byte offset >> 4
//   bits [46,48) version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand: rows of 128 B along K, 8-row groups 1024 B apart (SBO=1024); LBO unused (1).
// MN-major operand: rows of 128 B along MN, 8 k-rows per 1024 B group (SBO=1024), LBO = byte stride
// between consecutive 64-element (128 B) chunks along MN.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16 / kind::tf32, fp32 accumulate, dense.
//   [4,6) c_format (1=F32)  [7,10) a_format  [10,13) b_format  [15] a_major  [16] b_major
//   [17,23) N>>3  [24,29) M>>4       (major: 0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t make_instr_desc(uint32_t fmt, uint32_t m, uint32_t n, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
template <bool kTf32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// All previously issued MMAs of this thread arrive on `bar` when complete
// A operand in tensor memory (row i of A = TMEM lane i, 32-bit column c = elements k = 2c, 2c+1 of a 16-bit kind), B from
// shared memory: an A tile that is reused by many MMAs is then read from shared memory zero times instead of once per MMA.
__device__ __forceinline__ void umma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns; thread i of the warp
// receives lane (lane_base + i). A warp may only touch lanes [32*(warp_id%4), +32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but the destination registers of an earlier (still in flight) tcgen05.ld are tied to it as read-write operands:
// the compiler cannot schedule a use of them above the wait when other work sits between the load and the wait.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// registers -> TMEM, same shape (an epilogue can park a transformed accumulator row for a second pass).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// numerics shared by several kernels
// ----------------------------------------------------------------------------
// erf-GELU (torch.nn.GELU(approximate='none')) through the Abramowitz-Stegun 7.1.26 rational form of erfc:
//   erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + p z),  z >= 0,  |error| <= 1.5e-7.
// The negative branch uses erfc directly (1 + erf(-z) = erfc(z)), so the tail keeps its relative accuracy.  Absolute error
// of the GELU value <= 5e-7 for |x| <= 6 — far below the fp16 rounding of the tensor-core operand it feeds — at ~18
// instructions (one MUFU.RCP, one MUFU.EX2) instead of the ~100 of erff(): the GELU epilogue, not the MMA, bounds the
// fused ConvNeXt block.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float erfc_pos_as(float z, float& e_out) {
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));   // 1 ulp: far inside the 1.5e-7 error of the formula
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float e = __expf(-z * z);
  e_out = e;
  return t * poly * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float c = erfc_pos_as(fabsf(x) * 0.70710678118654752440f, e);  // erfc(|x|/sqrt2)
  return 0.5f * x * (x >= 0.f ? 2.0f - c : c);
}
// d/dx: Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;  // = exp(-x^2/2)
  const float c = erfc_pos_as(fabsf(x) * 0.70710678118654752440f, e);
  const float cdf = x >= 0.f ? 1.0f - 0.5f * c : 0.5f * c;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// Counter-based dropout: element `idx` of the tensor tagged `seed` is kept with probability 1-p and scaled by
// 1/(1-p).  Stateless, so the backward pass regenerates the forward mask.  Elements 2j and 2j+1 share one 32-bit hash of
// (j, seed) and take 16 bits each (p is resolved to 1/65536): the splitmix64 finaliser per ELEMENT this replaces cost ~40
// integer instructions against ~8 now, and the LayerNorm-backward epilogue of the predictor GEMMs evaluates the mask twice
// per element on 128 threads (39 -> 27 us for a 6144 x 256 x 1280 launch, most of it the mask).
//   h = fmix32(j ^ ka);  h = xs((h + kb) * c)        ka, kb: two independently mixed words of the 64-bit seed, so that the
// masks of two seeds are not index-permuted copies of each other (consecutive steps / layers use consecutive seeds).
__device__ __forceinline__ uint32_t drop_fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
struct DropKeys {
  uint32_t ka, kb;
};
__device__ __forceinline__ DropKeys dropout_keys(unsigned long long seed) {
  const uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
  DropKeys k;
  k.ka = drop_fmix32(lo ^ drop_fmix32(hi + 0x9E3779B9u));
  k.kb = drop_fmix32((lo + 0x85EBCA6Bu) * 0xC2B2AE35u ^ hi);
  return k;
}
// 32 mask bits of the element pair (2j, 2j+1): low half for 2j, high half for 2j+1
__device__ __forceinline__ uint32_t dropout_bits(DropKeys k, uint32_t j) {
  uint32_t h = drop_fmix32(j ^ k.ka);
  h = (h + k.kb) * 0x9E3779B1u;
  return h ^ (h >> 15);
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return static_cast<uint32_t>(p * 65536.0f); }
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx, float p, float inv_keep) {
  const uint32_t bits = dropout_bits(dropout_keys(seed), static_cast<uint32_t>(idx >> 1));
  const uint32_t u = (idx & 1ull) ? (bits >> 16) : (bits & 0xFFFFu);
  return u < dropout_threshold(p) ? 0.f : inv_keep;
}
// v[i] *= mask(base + i), i < N (N even): one hash per aligned pair when `base` is even (every call site: row * width + column
// block), the per-element definition above otherwise — same mask either way.
template <int N>
__device__ __forceinline__ void dropout_apply(float (&v)[N], unsigned long long seed, unsigned long long base, float p, float inv_keep) {
  const DropKeys k = dropout_keys(seed);
  const uint32_t thr = dropout_threshold(p);
  if ((base & 1ull) == 0ull) {
    const uint32_t j0 = static_cast<uint32_t>(base >> 1);
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const uint32_t bits = dropout_bits(k, j0 + static_cast<uint32_t>(i >> 1));
      v[i] *= (bits & 0xFFFFu) < thr ? 0.f : inv_keep;
      v[i + 1] *= (bits >> 16) < thr ? 0.f : inv_keep;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const unsigned long long idx = base + static_cast<unsigned long long>(i);
      const uint32_t bits = dropout_bits(k, static_cast<uint32_t>(idx >> 1));
      v[i] *= ((idx & 1ull) ? (bits >> 16) : (bits & 0xFFFFu)) < thr ? 0.f : inv_keep;
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace osb
