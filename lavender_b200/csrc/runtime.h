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
// profiling aid (lav_debug_set_trace): device buffer that instrumented kernels fill with clock64() stamps, or null
unsigned long long* trace_buffer();
int64_t trace_capacity();

// 2-D fp16 tensor map: `rows` x `cols` elements, row stride `ld` elements, box `box_rows` x `box_cols`,
// swizzle = CU_TENSOR_MAP_SWIZZLE_{32B,64B,128B}. Out-of-bounds elements read as zero.
int encode_tmap_2d_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

// same for 2-byte (fp16) or 4-byte (fp32) elements; also used for TMA stores (out-of-bounds elements are not written)
int encode_tmap_2d(CUtensorMap* map, const void* ptr, int elt_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle);

// Launch with the programmatic-stream-serialization attribute: the kernel may become resident while its predecessor
// in the stream drains.  ONLY for kernels that execute griddepcontrol.wait (sm100.cuh griddep_wait) before their first
// access to global data produced by earlier kernels.  Measured on the training step: enabling it for the GEMMs only
// (gemm.cu, LAV_PDL=0 disables) gains 0.8 ms/step, enabling it for the row / attention kernels as well LOSES 2 ms
// (their many small CTAs become resident early and take SM slots from the kernel that is still running), so for
// these kernels it is opt-in: LAV_PDL_ALL=1.
bool pdl_enabled();
bool pdl_all_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_all_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

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
