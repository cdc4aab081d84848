// kernels_refine.cuh — BoundaryRefiner.refine (Segmentation/full_evaluation_enhanced.py:332-393) on the device:
//   mask_u8 = (mask * 255).astype(uint8); boundary = (dilate > 0) xor (erode > 0); bilateral filter inside the boundary
//   band; MORPH_OPEN; MORPH_CLOSE; / 255.0.
// The reference calls OpenCV (opencv-python==4.8.0.76, requirements.txt:14); the algorithms restated here are the ones
// oracle/refine.py documents: elliptical structuring element (offsets computed on the host), erode / dilate that ignore
// pixels outside the image (BORDER_CONSTANT with the default border value), 8-bit bilateral filter with float32 running
// sums in tap order, colour-weight table, BORDER_REFLECT_101 and cvRound (round half to even).
// Byte planes, one thread per pixel, neighbours through L1 (32 consecutive bytes per warp and tap = one sector): at
// 1024^2 a pass is ~1 M x 21 byte loads, a few microseconds - the six passes are launch-bound, not HBM-bound.
#pragma once
#include "common.cuh"

namespace adp {

constexpr int kRefineMaxK = 15;                       // structuring element up to 15 x 15
constexpr int kRefineMaxD = 9;                        // bilateral diameter up to 9
struct RefineSE { int n; signed char dy[kRefineMaxK * kRefineMaxK], dx[kRefineMaxK * kRefineMaxK]; };
struct RefineTaps { int n; signed char dy[kRefineMaxD * kRefineMaxD], dx[kRefineMaxD * kRefineMaxD]; float sw[kRefineMaxD * kRefineMaxD]; };

// mask_u8 = (mask * 255).astype(np.uint8)  (:369): float32 product, truncation (values outside [0, 255] are clamped)
__global__ void __launch_bounds__(256) refine_quantize_kernel(const float *__restrict__ m, uint8_t *__restrict__ u, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __fmul_rn(m[i], 255.0f);
    u[i] = (uint8_t)min(max(__float2int_rz(v), 0), 255);
  }
}

ADP_DEVINL int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// eroded / dilated -> boundary band -> refined = boundary ? bilateral(mask_u8) : mask_u8   (:372-386)
__global__ void __launch_bounds__(256)
refine_band_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int H, int W, RefineSE se, RefineTaps bt,
                   const float *__restrict__ color_w /*[256]*/) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const uint8_t *img = src + (size_t)blockIdx.z * H * W;
  int mn = 255, mx = 0;
  for (int k = 0; k < se.n; ++k) {
    const int yy = y + se.dy[k], xx = x + se.dx[k];
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const int v = __ldg(img + (size_t)yy * W + xx);
    mn = min(mn, v); mx = max(mx, v);
  }
  const int v0 = __ldg(img + (size_t)y * W + x);
  int out = v0;
  if ((mx > 0) != (mn > 0)) {
    float sum = 0.f, wsum = 0.f;
    for (int k = 0; k < bt.n; ++k) {
      const int yy = reflect101(y + bt.dy[k], H), xx = reflect101(x + bt.dx[k], W);
      const int v = __ldg(img + (size_t)yy * W + xx);
      const float w = __fmul_rn(bt.sw[k], __ldg(color_w + abs(v - v0)));
      sum = __fadd_rn(sum, __fmul_rn((float)v, w));
      wsum = __fadd_rn(wsum, w);
    }
    out = min(max(__float2int_rn(__fdiv_rn(sum, wsum)), 0), 255);
  }
  dst[(size_t)blockIdx.z * H * W + (size_t)y * W + x] = (uint8_t)out;
}

// erode (ERODE) / dilate over the structuring element; fout != null: also write float32(v / 255.0) (:393)
template <bool ERODE>
__global__ void __launch_bounds__(256)
refine_morph_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, float *__restrict__ fout, int H, int W, RefineSE se) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const uint8_t *img = src + (size_t)blockIdx.z * H * W;
  int r = ERODE ? 255 : 0;
  for (int k = 0; k < se.n; ++k) {
    const int yy = y + se.dy[k], xx = x + se.dx[k];
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const int v = __ldg(img + (size_t)yy * W + xx);
    r = ERODE ? min(r, v) : max(r, v);
  }
  const size_t o = (size_t)blockIdx.z * H * W + (size_t)y * W + x;
  dst[o] = (uint8_t)r;
  if (fout) fout[o] = (float)((double)r / 255.0);
}

}  // namespace adp
