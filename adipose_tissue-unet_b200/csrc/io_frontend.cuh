// io_frontend.cuh — tile I/O front-end kernels, included by engine.cu:
//   * auxiliary whole-slide planes: the RGB mosaic (reconstruct_full_images.py:411-415: the three colour channels of every
//     tile, as float32 in [0,1], are blended with the SAME blender as the predictions) and the blended ground truth
//     (:404-409) accumulate on the device next to the probability accumulator and share its weight plane - no host-side
//     float32 tile lists (12 MB per RGB tile, 47 GB for a 32768^2 slide in the reference);
//   * uint8 export of a normalised plane: (value * 255).astype(uint8) (reconstruct_full_images.py:726, 733, 745);
//   * fat-% per tile: 100 * count(p > thr) / size (tile_classification_evaluation.py:211-225).
// fp32 arithmetic with explicit round-to-nearest ops (no FMA contraction): bit-identical to the NumPy statements.
#pragma once
#include "common.cuh"

namespace adp {

// One launch blends `nplanes` (1 or 3, interleaved in the tile) channels of one tile into consecutive aux planes.
// u8 != 0: tile bytes, value = float32(byte) / 255.0f  (cv2 tile .astype(np.float32) / 255.0, :368);  else float32 tile.
// mode 1: acc += v * w (Gaussian / Hann window);  mode 2: acc += v (linear).  The weight plane is NOT touched: it is the
// probability accumulator's (same tiles, same positions, same window => same sums).
__global__ void __launch_bounds__(256)
aux_blend_kernel(const void *__restrict__ tile, int u8, int nplanes, int S, int mode, float *__restrict__ planes, size_t plane_stride,
                 const float *__restrict__ window, int accW, int rlo, int rhi, int ty, int tx) {
  const size_t total = (size_t)S * S;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / S), j = (int)(p - (size_t)i * S);
    const int gy = ty + i, gx = tx + j;
    if (gy < rlo || gy >= rhi || gx < 0 || gx >= accW) continue;
    const size_t g = (size_t)gy * accW + gx;
    const float w = mode == 1 ? window[p] : 1.0f;
    for (int c = 0; c < nplanes; ++c) {
      float v;
      if (u8) v = __fdiv_rn((float)reinterpret_cast<const uint8_t *>(tile)[p * nplanes + c], 255.0f);
      else v = reinterpret_cast<const float *>(tile)[p * nplanes + c];
      float *a = planes + (size_t)c * plane_stride + g;
      *a = mode == 1 ? __fadd_rn(*a, __fmul_rn(v, w)) : __fadd_rn(*a, v);
    }
  }
}

ADP_DEVINL float aux_value(float a, float w, int linear) {
  if (linear) return (float)((double)a / (double)fmaxf(w, 1.0f));
  return __fdiv_rn(a, fmaxf(w, 1e-8f));
}

// out[px * nplanes + k] = uint8(normalised plane (order[k]) * 255)  — truncation like ndarray.astype(np.uint8) of a value in [0, 256)
__global__ void __launch_bounds__(256)
aux_export_u8_kernel(const float *__restrict__ planes, size_t plane_stride, const float *__restrict__ wsum, int linear, size_t n,
                     int nplanes, int reverse, uint8_t *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float w = wsum[i];
    for (int k = 0; k < nplanes; ++k) {
      const int c = reverse ? nplanes - 1 - k : k;
      const float v = __fmul_rn(aux_value(planes[(size_t)c * plane_stride + i], w, linear), 255.0f);
      out[i * nplanes + k] = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f);
    }
  }
}

// float32 export of one normalised plane (the blended ground truth, :404-409)
__global__ void __launch_bounds__(256)
aux_export_f32_kernel(const float *__restrict__ plane, const float *__restrict__ wsum, int linear, size_t n, float *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = aux_value(plane[i], wsum[i], linear);
}

// normalise + threshold + confusion counts with the ground truth taken from a blended aux plane:
// true = (gt_blend > 0.5) (calculate_pixel_metrics, full_evaluation_enhanced.py:737, on the blended full_gt of :404-409)
__global__ void __launch_bounds__(256)
finalize_auxgt_kernel(const float *__restrict__ acc, const float *__restrict__ wsum, const float *__restrict__ gt_plane, int linear, size_t n,
                      float thr, float *__restrict__ prob, uint8_t *__restrict__ mask, unsigned long long *__restrict__ counts) {
  unsigned tp = 0, fp = 0, fn = 0, tn = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float w = wsum[i];
    const float p = aux_value(acc[i], w, linear);
    const bool pb = p > thr, tb = aux_value(gt_plane[i], w, linear) > 0.5f;
    if (prob) prob[i] = p;
    if (mask) mask[i] = pb ? 1 : 0;
    tp += (pb && tb); fp += (pb && !tb); fn += (!pb && tb); tn += (!pb && !tb);
  }
  tp = warp_sum_u(tp); fp = warp_sum_u(fp); fn = warp_sum_u(fn); tn = warp_sum_u(tn);
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(&counts[0], (unsigned long long)tp);
    if (fp) atomicAdd(&counts[1], (unsigned long long)fp);
    if (fn) atomicAdd(&counts[2], (unsigned long long)fn);
    if (tn) atomicAdd(&counts[3], (unsigned long long)tn);
  }
}

// count[t] = number of pixels of tile t with value > thr; grid = (blocks per tile, tiles)
__global__ void __launch_bounds__(256)
tile_count_kernel(const float *__restrict__ v, size_t px, float thr, unsigned long long *__restrict__ count) {
  const float *t = v + (size_t)blockIdx.y * px;
  unsigned c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < px; i += (size_t)gridDim.x * blockDim.x) c += t[i] > thr;
  c = warp_sum_u(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&count[blockIdx.y], (unsigned long long)c);
}

}  // namespace adp
