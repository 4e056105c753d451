// osb_host.h — host-side helpers shared by the .cu translation units:
// status codes, launch bookkeeping and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/osb200.h"

namespace osb {

// Number of kernels launched by this library since load (bench.py reports it as gpu_launches).
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += static_cast<unsigned long long>(n); }

// Positive return values are cudaError_t, negative are osb_status (see include/osb200.h).
inline int launch_status() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? OSB_OK : static_cast<int>(e);
}

#define OSB_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

enum TmaDtype { TMA_F16 = 0, TMA_BF16 = 1, TMA_F32 = 2 };

// 3-D row-major tensor (d2, d1, d0) with d0 contiguous; strides in elements for d1 and d2.
// Box = (box0, box1, 1), 128-byte swizzle, zero fill out of bounds.
// `elem_stride1` > 1: traversal stride along d1 (box1 then counts SOURCE rows: box1 = rows_to_load * elem_stride1 <= 256).
int make_tmap_3d(CUtensorMap* out, const void* base, TmaDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1, uint32_t elem_stride1 = 1);

}  // namespace osb
