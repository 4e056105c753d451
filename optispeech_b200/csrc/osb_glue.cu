// osb_glue.cu — the small index / mask kernels of the path (HBM-bound, coalesced, one launch each):
//
//   sequence_mask    lengths -> (B, T) byte masks, valid and/or padding         (utils/model.py:12-21)
//   segment_starts   start = floor(rand * max(len - margin - S, 0))             (utils/segments.py:29-35)
//   gather_segments  out[b, s, :] = x[b, start[b]*scale + s, :], zero past T     (utils/segments.py:38-72)
//
// They replace chains of 3-6 eager elementwise / gather kernels each in the training step.
#include "osb_host.h"
#include "osb_ptx.cuh"

namespace osb {
namespace {

__global__ void __launch_bounds__(256)
sequence_mask_kernel(const long long* __restrict__ lengths, uint8_t* __restrict__ valid, uint8_t* __restrict__ pad, int B, int T) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * T) return;
  const int b = static_cast<int>(i / T), t = static_cast<int>(i % T);
  const uint8_t v = t < lengths[b] ? 1 : 0;
  if (valid != nullptr) valid[i] = v;
  if (pad != nullptr) pad[i] = v ^ 1;
}

__global__ void segment_starts_kernel(const float* __restrict__ rand, const long long* __restrict__ lengths, long long* __restrict__ start,
                                      int B, int margin, int S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // the reference computes in the float dtype of the decoder output: (len - 4).to(float) - S, clamp, rand * max, .long()
  float mx = static_cast<float>(lengths[b] - margin) - static_cast<float>(S);
  mx = mx < 0.f ? 0.f : mx;
  start[b] = static_cast<long long>(rand[b] * mx);   // truncation toward zero == .to(torch.long) for non-negative values
}

// one warp per output row (b, s); rows of C floats, float4 when C % 4 == 0
__global__ void __launch_bounds__(256)
gather_segments_kernel(const float* __restrict__ x, const long long* __restrict__ start, float* __restrict__ out, int B, long long T, int C,
                       int S, int scale) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * S) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / S), s = static_cast<int>(row % S);
  const long long t = start[b] * scale + s;
  const bool ok = t >= 0 && t < T;
  const float* src = x + (static_cast<long long>(b) * T + (ok ? t : 0)) * C;
  float* dst = out + row * C;
  if ((C & 3) == 0) {
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = ok ? *reinterpret_cast<const float4*>(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dst + c) = v;
    }
  } else {
    for (int c = lane; c < C; c += 32) dst[c] = ok ? src[c] : 0.f;
  }
}

// C == 1 (waveforms): one thread per output sample, contiguous along s
__global__ void __launch_bounds__(256)
gather_segments_c1_kernel(const float* __restrict__ x, const long long* __restrict__ start, float* __restrict__ out, int B, long long T, int S,
                          int scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * S) return;
  const int b = static_cast<int>(i / S), s = static_cast<int>(i % S);
  const long long t = start[b] * scale + s;
  out[i] = (t >= 0 && t < T) ? x[static_cast<long long>(b) * T + t] : 0.f;
}

}  // namespace
}  // namespace osb

using namespace osb;

extern "C" int osb_sequence_mask(const int64_t* lengths, uint8_t* valid, uint8_t* pad, int32_t B, int32_t T, void* stream) {
  OSB_REQUIRE(lengths && (valid || pad), OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0, OSB_ERR_SHAPE);
  const long long n = static_cast<long long>(B) * T;
  sequence_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(lengths), valid, pad, B, T);
  count_launch();
  return launch_status();
}

extern "C" int osb_segment_starts(const float* rand, const int64_t* lengths, int64_t* start, int32_t B, int32_t margin, int32_t S,
                                  void* stream) {
  OSB_REQUIRE(rand && lengths && start, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && S > 0 && margin >= 0, OSB_ERR_SHAPE);
  segment_starts_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      rand, reinterpret_cast<const long long*>(lengths), reinterpret_cast<long long*>(start), B, margin, S);
  count_launch();
  return launch_status();
}

extern "C" int osb_gather_segments(const float* x, const int64_t* start, float* out, int32_t B, int64_t T, int32_t C, int32_t S,
                                   int32_t scale, void* stream) {
  OSB_REQUIRE(x && start && out, OSB_ERR_ARG);
  OSB_REQUIRE(B > 0 && T > 0 && C > 0 && S > 0 && scale > 0, OSB_ERR_SHAPE);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(B) * S;
  if (C == 1) {
    gather_segments_c1_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, s>>>(x, reinterpret_cast<const long long*>(start), out, B,
                                                                                        T, S, scale);
  } else {
    if ((C & 3) == 0) OSB_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, OSB_ERR_ALIGN);
    gather_segments_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(x, reinterpret_cast<const long long*>(start), out, B, T, C,
                                                                                 S, scale);
  }
  count_launch();
  return launch_status();
}
