// GPU input pipeline: decoded uint8 RGB frames -> resized, cropped, normalised fp32 clip tensor.
// Restates the per-frame transform of the reference's loader (dataset.py:118-175: torchvision / torch_videovision
// Resize(size_img) -> RandomCrop | CenterCrop -> ToTensor -> Normalize(mean, std)) on PIL images, whose Resize is PIL's
// two-pass antialiased bilinear resample (Pillow src/libImaging/Resample.c): horizontal pass, result rounded to uint8,
// then the vertical pass, rounded to uint8 again; filter support = max(scale, 1), triangle weights normalised to 1 and
// quantised to 22-bit fixed point.  The kernel reproduces exactly that arithmetic (integer accumulation, the
// intermediate uint8 rounding included), one thread per output pixel: HBM-bound byte work, ~16 taps x 3 channels each.
#include "runtime.h"
#include "sm100.cuh"

namespace lav {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

struct FrameParams {
  const uint8_t* src; int64_t frame_stride;  // [T][Hs][Ws][3] uint8 (row pitch Ws * 3)
  int T, Hs, Ws;                              // decoded frame size
  int Hr, Wr;                                 // size after Resize (shorter side = size_img)
  int S, top, left;                           // crop window [top, top + S) x [left, left + S) of the resized frame
  float mean[3], inv_std[3];
  float* out;                                 // [T][3][S][S]
};

// bounds and fixed-point coefficients of output coordinate `xx` (Pillow precompute_coeffs + normalize_coeffs_8bpc)
__device__ __forceinline__ int resample_coeffs(int xx, int in_size, int out_size, int& xmin, int (&kk)[24]) {
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;  // bilinear filter support 1.0
  const double center = (xx + 0.5) * scale;
  const double ss = 1.0 / filterscale;
  xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > 24) xmax = 24;
  double k[24];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double t = (x + xmin - center + 0.5) * ss;
    if (t < 0.0) t = -t;
    const double w = t < 1.0 ? 1.0 - t : 0.0;
    k[x] = w;
    ww += w;
  }
  for (int x = 0; x < xmax; ++x) {
    const double w = ww != 0.0 ? k[x] / ww : k[x];
    kk[x] = w < 0 ? (int)(-0.5 + w * (1 << kPrecisionBits)) : (int)(0.5 + w * (1 << kPrecisionBits));
  }
  return xmax;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__global__ void __launch_bounds__(256) frames_resize_crop_norm_kernel(const FrameParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;  // crop column
  const int y = blockIdx.y, t = blockIdx.z;
  if (x >= p.S) return;
  int xmin, ymin, kx[24], ky[24];
  const int nx = resample_coeffs(x + p.left, p.Ws, p.Wr, xmin, kx);
  const int ny = resample_coeffs(y + p.top, p.Hs, p.Hr, ymin, ky);
  const uint8_t* f = p.src + (int64_t)t * p.frame_stride;
  int acc[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  for (int j = 0; j < ny; ++j) {
    const uint8_t* row = f + ((int64_t)(ymin + j) * p.Ws + xmin) * 3;
    int h[3] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
    for (int i = 0; i < nx; ++i) {
      h[0] += row[3 * i] * kx[i], h[1] += row[3 * i + 1] * kx[i], h[2] += row[3 * i + 2] * kx[i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += clip8(h[c]) * ky[j];   // the horizontal pass is stored as uint8 by Pillow
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (float)clip8(acc[c]) / 255.0f;              // ToTensor
    p.out[(((int64_t)t * 3 + c) * p.S + y) * p.S + x] = (v - p.mean[c]) * p.inv_std[c];   // Normalize
  }
}

}  // namespace lav

using namespace lav;

extern "C" int lav_frames_resize_crop_norm_u8(const uint8_t* src, int T, int Hs, int Ws, int64_t frame_stride, int Hr, int Wr,
                                              int S, int top, int left, const float* mean, const float* stdv, float* out,
                                              void* stream) {
  LAV_REQUIRE(src && out && mean && stdv, "lav_frames_resize_crop_norm_u8: null pointer");
  LAV_REQUIRE(T > 0 && Hs > 0 && Ws > 0 && Hr > 0 && Wr > 0 && S > 0, "lav_frames_resize_crop_norm_u8: empty frame");
  LAV_REQUIRE(top >= 0 && left >= 0 && top + S <= Hr && left + S <= Wr, "lav_frames_resize_crop_norm_u8: crop outside the resized frame");
  LAV_REQUIRE((double)Hs / Hr <= 11.0 && (double)Ws / Wr <= 11.0, "lav_frames_resize_crop_norm_u8: down-scaling factor > 11 not supported");
  FrameParams p;
  p.src = src, p.frame_stride = frame_stride, p.T = T, p.Hs = Hs, p.Ws = Ws, p.Hr = Hr, p.Wr = Wr, p.S = S, p.top = top, p.left = left;
  for (int c = 0; c < 3; ++c) p.mean[c] = mean[c], p.inv_std[c] = 1.0f / stdv[c];
  p.out = out;
  dim3 grid((S + 255) / 256, S, T);
  frames_resize_crop_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  LAV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return LAV_OK;
}
