#include "runtime.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace lav {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static unsigned long long* g_trace = nullptr;
static int64_t g_trace_cap = 0;
unsigned long long* trace_buffer() { return g_trace; }
int64_t trace_capacity() { return g_trace_cap; }
void set_trace(unsigned long long* p, int64_t n) { g_trace = p, g_trace_cap = n; }

bool pdl_enabled() {  // LAV_PDL=0 disables programmatic dependent launch
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAV_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool pdl_all_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LAV_PDL_ALL");
    v = (pdl_enabled() && e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

void set_trace(unsigned long long* p, int64_t n);

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // resolved through the runtime so the library does not link libcuda (absent on the build box)
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_2d_f16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  return encode_tmap_2d(map, ptr, 2, rows, cols, ld, box_rows, box_cols, swizzle);
}

int encode_tmap_2d(CUtensorMap* map, const void* ptr, int elt_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                   uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(LAV_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || ((ld * elt_bytes) % 16) != 0 || (elt_bytes != 2 && elt_bytes != 4))
    return set_error(LAV_E_INVALID, "TMA operand must be 16B aligned with a 16B-multiple row pitch (ptr=%p ld=%llu)", ptr,
                     (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elt_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(LAV_E_CUDA, "cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
  return LAV_OK;
}

}  // namespace lav

extern "C" {

int lav_abi_version(void) { return 3; }
int lav_debug_set_trace(void* buf, int64_t n_u64) {
  lav::set_trace(reinterpret_cast<unsigned long long*>(buf), buf ? n_u64 : 0);
  return LAV_OK;
}
const char* lav_last_error(void) { return lav::g_err; }
int64_t lav_launch_count(void) { return lav::g_launches.load(); }

int lav_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n)
    return lav::set_error(LAV_E_NO_DEVICE, "no CUDA device %d", device);
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return lav::set_error(LAV_E_CUDA, "cudaGetDeviceProperties");
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) return lav::set_error(LAV_E_NO_DEVICE, "device %d is sm_%d%d, need sm_100", device, p.major, p.minor);
  return LAV_OK;
}

}  // extern "C"
