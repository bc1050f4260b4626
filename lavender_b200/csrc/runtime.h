// Host-side runtime shared by all kernels: error reporting, launch accounting, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lavender_b200.h"

namespace lav {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();  // of the current device (cached per device)

// 2-D fp16 tensor map: `rows` x `cols` elements, row stride `ld` elements, box `box_rows` x `box_cols`,
// swizzle = CU_TENSOR_MAP_SWIZZLE_{32B,64B,128B}. Out-of-bounds elements read as zero.
int encode_tmap_2d_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

// same for 2-byte (fp16) or 4-byte (fp32) elements; also used for TMA stores (out-of-bounds elements are not written)
int encode_tmap_2d(CUtensorMap* map, const void* ptr, int elt_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

#define LAV_CHECK_CUDA(expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return lav::set_error(LAV_E_CUDA, "%s failed: %s (%s:%d)", #expr,          \
                                                 cudaGetErrorString(_e), __FILE__, __LINE__);         \
  } while (0)

#define LAV_REQUIRE(cond, ...)                                      \
  do {                                                              \
    if (!(cond)) return lav::set_error(LAV_E_INVALID, __VA_ARGS__); \
  } while (0)

}  // namespace lav
