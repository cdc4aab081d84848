// kernels_post.cuh — HBM-bound post-processing kernels: TTA de-augmentation + mean + window
// blend into the persistent slide accumulator, normalise/threshold/confusion counts, BCE+Dice
// reductions.  fp32 arithmetic is written with explicit __fmul_rn/__fadd_rn/__fdiv_rn so that the
// results are bit-identical to the NumPy statements of the reference (no FMA contraction).
#pragma once
#include "common.cuh"

namespace adp {

// ---------------------------------------------------------------------------------------------
// TTA combine (+ optional blend).  planes: n_ops probability planes of one tile in augmented
// space, S*S each, consecutive.  avg[i][j] = (sum_k deaug_k(P_k)[i][j]) / n_ops with sequential
// float32 adds in list order (np.mean(preds, axis=0), full_evaluation_enhanced.py:594).
// deaug_k = dihedral op inverse(ops[k]):  deaug(P)[i][j] = P[src(inv, i, j)].
// Each 32x32 output block reads, per plane, the one 32x32 source block it maps to with
// coalesced rows and transposes through shared memory.
//
// mode 0: write avg to out (S*S, row pitch S)
// mode 1: Gaussian blend: acc[y+i][x+j] += avg*w[i][j]; wsum[y+i][x+j] += w[i][j]
//         (GaussianBlender.reconstruct, full_evaluation_enhanced.py:165-173)
// mode 2: linear blend:   acc += avg; wsum(count as float) += 1   (LinearBlender, :196-199)
// Only accumulator rows [rlo, rhi) are touched (the whole accumulator normally; a strip that defers its boundary zone
// blends the rows below the zone first and replays the zone rows after the upper strip's partial sums have arrived).
struct TtaOps { int n; int inv[8]; };

// All source blocks (one 32x32 block per augmentation) are requested before the single barrier: 4 * n loads per thread in
// flight, issued from a register array so that no shared-memory store sits between two global loads.  ADP_TTA_SERIAL=1
// selects the one-plane-per-barrier walk (tta_blend_serial_kernel, eight dependent DRAM round trips per block) for A/B runs.
__global__ void __launch_bounds__(256)
tta_blend_kernel(const float *__restrict__ planes, TtaOps ops, int S, int mode, float *__restrict__ out,
                 float *__restrict__ acc, float *__restrict__ wsum, const float *__restrict__ window,
                 int accW, int rlo, int rhi, int ty, int tx) {
  __shared__ float tile[8][32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int lx = threadIdx.x, ly = threadIdx.y;     // 32 x 8
  const int ei = min(bi + 31, S - 1), ej = min(bj + 31, S - 1);
  float ld[8][4];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ops.n) {
      // source block origin: image of the block's corner region under the inverse op
      int s0i, s0j, s1i, s1j;
      d4_src(ops.inv[k], bi, bj, S, s0i, s0j);
      d4_src(ops.inv[k], ei, ej, S, s1i, s1j);
      const int sbi = min(s0i, s1i), sbj = min(s0j, s1j);
      const float *P = planes + (size_t)k * S * S;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = sbi + ly + 8 * r, j = sbj + lx;
        ld[k][r] = (i < S && j < S) ? __ldg(P + (size_t)i * S + j) : 0.f;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < ops.n) {
#pragma unroll
      for (int r = 0; r < 4; ++r) tile[k][ly + 8 * r][lx] = ld[k][r];
    }
  __syncthreads();
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ops.n) {
      const int op = ops.inv[k];
      int s0i, s0j, s1i, s1j;
      d4_src(op, bi, bj, S, s0i, s0j);
      d4_src(op, ei, ej, S, s1i, s1j);
      const int sbi = min(s0i, s1i), sbj = min(s0j, s1j);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = bi + ly + 8 * r, j = bj + lx;
        float v = 0.f;
        if (i < S && j < S) {
          int si, sj;
          d4_src(op, i, j, S, si, sj);
          v = tile[k][si - sbi][sj - sbj];
        }
        sum[r] = (k == 0) ? v : __fadd_rn(sum[r], v);
      }
    }
  }
  const float nf = (float)ops.n;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = bi + ly + 8 * r, j = bj + lx;
    if (i >= S || j >= S) continue;
    float avg = (ops.n == 1) ? sum[r] : __fdiv_rn(sum[r], nf);
    if (mode == 0) {
      out[(size_t)i * S + j] = avg;
    } else {
      int gy = ty + i, gx = tx + j;
      if (gy < rlo || gy >= rhi || gx < 0 || gx >= accW) continue;
      size_t g = (size_t)gy * accW + gx;
      if (mode == 1) {
        float w = window[(size_t)i * S + j];
        acc[g] = __fadd_rn(acc[g], __fmul_rn(avg, w));
        wsum[g] = __fadd_rn(wsum[g], w);
      } else {
        acc[g] = __fadd_rn(acc[g], avg);
        wsum[g] = __fadd_rn(wsum[g], 1.0f);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
tta_blend_serial_kernel(const float *__restrict__ planes, TtaOps ops, int S, int mode, float *__restrict__ out,
                        float *__restrict__ acc, float *__restrict__ wsum, const float *__restrict__ window,
                        int accW, int rlo, int rhi, int ty, int tx) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int lx = threadIdx.x, ly = threadIdx.y;     // 32 x 8
  float sum[4];
  for (int k = 0; k < ops.n; ++k) {
    const int op = ops.inv[k];
    // source block origin: image of the block's (bi,bj) corner region under op
    int s0i, s0j, s1i, s1j;
    d4_src(op, bi, bj, S, s0i, s0j);
    d4_src(op, min(bi + 31, S - 1), min(bj + 31, S - 1), S, s1i, s1j);
    const int sbi = min(s0i, s1i), sbj = min(s0j, s1j);
    const float *P = planes + (size_t)k * S * S;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int i = sbi + ly + 8 * r, j = sbj + lx;
      tile[ly + 8 * r][lx] = (i < S && j < S) ? P[(size_t)i * S + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int i = bi + ly + 8 * r, j = bj + lx;
      float v = 0.f;
      if (i < S && j < S) {
        int si, sj;
        d4_src(op, i, j, S, si, sj);
        v = tile[si - sbi][sj - sbj];
      }
      sum[r] = (k == 0) ? v : __fadd_rn(sum[r], v);
    }
  }
  const float nf = (float)ops.n;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = bi + ly + 8 * r, j = bj + lx;
    if (i >= S || j >= S) continue;
    float avg = (ops.n == 1) ? sum[r] : __fdiv_rn(sum[r], nf);
    if (mode == 0) {
      out[(size_t)i * S + j] = avg;
    } else {
      int gy = ty + i, gx = tx + j;
      if (gy < rlo || gy >= rhi || gx < 0 || gx >= accW) continue;
      size_t g = (size_t)gy * accW + gx;
      if (mode == 1) {
        float w = window[(size_t)i * S + j];
        acc[g] = __fadd_rn(acc[g], __fmul_rn(avg, w));
        wsum[g] = __fadd_rn(wsum[g], w);
      } else {
        acc[g] = __fadd_rn(acc[g], avg);
        wsum[g] = __fadd_rn(wsum[g], 1.0f);
      }
    }
  }
}

// Blend of an already de-augmented tile of arbitrary size th x tw (adp_blend_reconstruct,
// adp_wsi_push_probs).  One thread per tile pixel.
__global__ void __launch_bounds__(256)
blend_tile_kernel(const float *__restrict__ tile, int th, int tw, int mode, float *__restrict__ acc,
                  float *__restrict__ wsum, const float *__restrict__ window, int winW, int accW, int rlo, int rhi,
                  int ty, int tx) {
  const size_t total = (size_t)th * tw;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    int i = p / tw, j = p % tw;
    int gy = ty + i, gx = tx + j;
    if (gy < rlo || gy >= rhi || gx < 0 || gx >= accW) continue;
    size_t g = (size_t)gy * accW + gx;
    float v = tile[p];
    if (mode == 1) {
      float w = window[(size_t)i * winW + j];
      acc[g] = __fadd_rn(acc[g], __fmul_rn(v, w));
      wsum[g] = __fadd_rn(wsum[g], w);
    } else {
      acc[g] = __fadd_rn(acc[g], v);
      wsum[g] = __fadd_rn(wsum[g], 1.0f);
    }
  }
}

__global__ void __launch_bounds__(256)
add_partial_kernel(float *__restrict__ acc, float *__restrict__ wsum, const float *__restrict__ a2,
                   const float *__restrict__ w2, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    acc[i] = __fadd_rn(acc[i], a2[i]);
    wsum[i] = __fadd_rn(wsum[i], w2[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// Normalise + threshold + confusion counts in one pass.
//   gaussian: p = acc / max(wsum, 1e-8)            (full_evaluation_enhanced.py:176-181)
//   linear:   p = float32(float64(acc) / max(count,1))   (:201-204; NumPy promotes f32/i32 to f64)
//   raw (acc only, wsum == null): p = acc           (adp_threshold_metrics)
// mask = p > thr (strict, :716-718); truth = gt > 0.5 -> gt != 0 for uint8 masks (:737).
// counts[0..3] = tp, fp, fn, tn  (unsigned 64-bit atomics, one per warp after shuffle reduction)
ADP_DEVINL float finalize_value(float a, float w, bool has_w, int linear) {
  if (!has_w) return a;
  if (linear) return (float)((double)a / (double)fmaxf(w, 1.0f));
  return __fdiv_rn(a, fmaxf(w, 1e-8f));
}

// vec != 0: every pointer is 16-byte (float) / 4-byte (uint8) aligned - four pixels per thread through 128-bit / 32-bit
// accesses (the byte-wide scalar form reached 21 % of DRAM peak); the tail (n % 4) and unaligned calls take the scalar path.
__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ acc, const float *__restrict__ wsum, int linear, size_t n, float thr,
                float *__restrict__ prob, uint8_t *__restrict__ mask, const uint8_t *__restrict__ gt,
                unsigned long long *__restrict__ counts, int vec) {
  unsigned tp = 0, fp = 0, fn = 0, tn = 0;
  const size_t nq = vec ? n / 4 : 0;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
    const float4 a4 = reinterpret_cast<const float4 *>(acc)[q];
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wsum) w4 = reinterpret_cast<const float4 *>(wsum)[q];
    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
    uchar4 g4 = make_uchar4(0, 0, 0, 0);
    if (gt) g4 = reinterpret_cast<const uchar4 *>(gt)[q];
    const unsigned char g[4] = {g4.x, g4.y, g4.z, g4.w};
    float p[4]; unsigned char m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      p[k] = finalize_value(a[k], w[k], wsum != nullptr, linear);
      const bool pb = p[k] > thr, tb = gt ? (g[k] != 0) : false;
      m[k] = pb ? 1 : 0;
      tp += (pb && tb); fp += (pb && !tb); fn += (!pb && tb); tn += (!pb && !tb);
    }
    if (prob) reinterpret_cast<float4 *>(prob)[q] = make_float4(p[0], p[1], p[2], p[3]);
    if (mask) reinterpret_cast<uchar4 *>(mask)[q] = make_uchar4(m[0], m[1], m[2], m[3]);
  }
  for (size_t i = nq * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = finalize_value(acc[i], wsum ? wsum[i] : 0.f, wsum != nullptr, linear);
    if (prob) prob[i] = p;
    const bool pb = p > thr;
    if (mask) mask[i] = pb ? 1 : 0;
    const bool tb = gt ? (gt[i] != 0) : false;
    tp += (pb && tb); fp += (pb && !tb); fn += (!pb && tb); tn += (!pb && !tb);
  }
  tp = warp_sum_u(tp); fp = warp_sum_u(fp); fn = warp_sum_u(fn); tn = warp_sum_u(tn);
  if ((threadIdx.x & 31) == 0 && counts) {
    if (tp) atomicAdd(&counts[0], (unsigned long long)tp);
    if (fp) atomicAdd(&counts[1], (unsigned long long)fp);
    if (fn) atomicAdd(&counts[2], (unsigned long long)fn);
    if (tn) atomicAdd(&counts[3], (unsigned long long)tn);
  }
}

// ---------------------------------------------------------------------------------------------
// Threshold sweep (optimize_threshold_f1[_slide_level], full_evaluation_enhanced.py:891-980: the reference re-runs
// calculate_pixel_metrics for every candidate threshold).  One pass: every pixel is binned by k = number of candidate
// thresholds strictly below its probability (p > thr_j  <=>  j < k for ascending thr) separately for gt = 0 / 1; the host
// turns the two histograms into TP/FP/FN/TN of every candidate by suffix sums.  hist: [2][nthr + 1] unsigned 64-bit.
__global__ void __launch_bounds__(256)
threshold_sweep_kernel(const float *__restrict__ prob, const uint8_t *__restrict__ gt, size_t n, const float *__restrict__ thr,
                       int nthr, unsigned long long *__restrict__ hist) {
  __shared__ float st[64];
  __shared__ unsigned int h[2][65];
  for (int i = threadIdx.x; i < nthr; i += blockDim.x) st[i] = thr[i];
  for (int i = threadIdx.x; i < 2 * 65; i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float p = prob[i];
    int lo = 0, hi = nthr;                 // k = first index with !(p > thr[k])  (thr ascending)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (p > st[mid]) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&h[gt[i] != 0 ? 1 : 0][lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 65; i += blockDim.x) {
    const unsigned int v = (&h[0][0])[i];
    const int c = i / 65, k = i % 65;
    if (v && k <= nthr) atomicAdd(&hist[c * (nthr + 1) + k], (unsigned long long)v);
  }
}

// ---------------------------------------------------------------------------------------------
// Loss family of the reference (train_adipose_unet_v3.py:217-363) and dice_coef (src/utils/model.py:93-98):
//   combined_loss_standard                       bce_mean(y, p) + dice_loss(y, p)
//   combined_loss_with_label_smoothing           same on ys = y*(1 - eps_pos - eps_neg) + eps_neg          (:244-279)
//   online_hard_example_mining_loss[_with_smoothing]                                                        (:282-363)
//       tf.keras.losses.binary_crossentropy averages its LAST axis, and the model output / target are (B, H, W)
//       (:748-750, :611-613), so "per_pixel_bce" is the (B, H) tensor of per-ROW means; flat_loss is (B, H),
//       k = int(float32(H) * keep_ratio) (716 of 1024 rows at the default 0.7) and hard_bce is the mean of the k largest
//       row means of every image.  dice_loss runs over all pixels.
// Pass 1 (loss_reduce_kernel, one warp per image row): float64 sums  S0 = sum bce_i, S1 = sum ys*pc, S2 = sum ys,
//         S3 = sum pc, S4 = sum y*p, S5 = sum p, S6 = sum y  (pc = clip(p, 1e-7, 1-1e-7); the metric dice_coef uses the raw y);
//         with hard mining the per-row BCE means are kept (float32, like the tensor TensorFlow ranks).
// OHEM:   ohem_select_kernel, one block per image: rank of every row mean by counting (ties go to the lower row index, as in
//         tf.nn.top_k), rows with rank < k are selected: row_w = 1/W (the d(row mean)/d(pixel bce) factor) else 0; the
//         selected means are summed in a fixed order.
// Finish: loss_finish_kernel writes S0 (hard mining: sum of the selected row means) and S7 = number of BCE terms in the mean
//         (all pixels, or batch*k rows), so that the eight sums live on the device, additive over data-parallel ranks.
// Pass 2 (loss_grad_kernel, scalars read from the device sums): dL/dp_i = w_i * dbce_i / S7 + ddice_i with
//   dbce_i  = -( ys/(pc+eps) - (1-ys)/(1-pc+eps) ) * [eps <= p <= 1-eps]
//   ddice_i = -( 2*ys*D - (2I+1) ) / D^2 * [eps <= p <= 1-eps],  I = S1, D = S2 + S3 + 1
//   w_i = row_w[row of i] with hard mining, 1 otherwise
struct LossRecipe {
  float ohem_keep = 1.f;     // 1 = no hard-example mining
  float eps_pos = 0.f, eps_neg = 0.f;
  __host__ __device__ float ys_scale() const { return 1.0f - eps_pos - eps_neg; }
};

constexpr int kOhemMaxRows = 8192;      // rows per image the one-block select holds in shared memory

// rows of W pixels (the last one may be short when n % W != 0: flat calls without hard mining); one warp per row
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float *__restrict__ p, const float *__restrict__ y, size_t n, int W, float ys_a, float ys_b,
                   float *__restrict__ row_bce /* or null */, double *__restrict__ sums) {
  double sb = 0, syp = 0, sy = 0, spc = 0, sypr = 0, sp = 0, syr = 0;
  const float eps = 1e-7f;
  const int lane = threadIdx.x & 31;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t rows = (n + (size_t)W - 1) / (size_t)W;
  for (size_t r = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
    const size_t base = r * (size_t)W;
    const int len = (int)((n - base) < (size_t)W ? (n - base) : (size_t)W);
    double rb = 0;
    for (int j = lane; j < len; j += 32) {
      const float pv = p[base + j], yr = y[base + j];
      const float yv = yr * ys_a + ys_b;
      const float pc = fminf(fmaxf(pv, eps), 1.0f - eps);
      const float bce = -(yv * logf(pc + eps) + (1.0f - yv) * logf(1.0f - pc + eps));
      rb += bce; syp += (double)yv * pc; sy += yv; spc += pc; sypr += (double)yr * pv; sp += pv; syr += yr;
    }
    sb += rb;
    if (row_bce) {
      rb = warp_sum_d(rb);
      if (lane == 0) row_bce[r] = (float)(rb / (double)W);
    }
  }
  sb = warp_sum_d(sb); syp = warp_sum_d(syp); sy = warp_sum_d(sy); spc = warp_sum_d(spc);
  sypr = warp_sum_d(sypr); sp = warp_sum_d(sp); syr = warp_sum_d(syr);
  if (lane == 0) {
    atomicAdd(&sums[0], sb); atomicAdd(&sums[1], syp); atomicAdd(&sums[2], sy);
    atomicAdd(&sums[3], spc); atomicAdd(&sums[4], sypr); atomicAdd(&sums[5], sp); atomicAdd(&sums[6], syr);
  }
}

// One block (1024 threads) per image: top-k of the R row means by rank counting; dynamic shared memory = R floats.
__global__ void __launch_bounds__(1024)
ohem_select_kernel(const float *__restrict__ row_bce, int R, int k, float inv_w, float *__restrict__ row_w,
                   double *__restrict__ sel_sum) {
  extern __shared__ float sv[];
  __shared__ double part[32];
  const int img = blockIdx.x;
  const float *b = row_bce + (size_t)img * R;
  for (int i = threadIdx.x; i < R; i += blockDim.x) sv[i] = b[i];
  __syncthreads();
  double mine = 0;
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    const float v = sv[i];
    int rank = 0;
    for (int j = 0; j < R; ++j) {
      const float u = sv[j];
      rank += (u > v) || (u == v && j < i);
    }
    const bool sel = rank < k;
    row_w[(size_t)img * R + i] = sel ? inv_w : 0.f;
    if (sel) mine += (double)v;
  }
  mine = warp_sum_d(mine);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w];
    sel_sum[img] = s;
  }
}

// sums[7] = number of BCE terms in the mean; with hard mining sums[0] = sum over images of the selected row means
__global__ void loss_finish_kernel(double *__restrict__ sums, const double *__restrict__ sel_sum /* or null */, int batch,
                                   double n_terms) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (sel_sum) {
      double s = 0;
      for (int b = 0; b < batch; ++b) s += sel_sum[b];
      sums[0] = s;
    }
    sums[7] = n_terms;
  }
}

// Keras' 'binary_accuracy' metric of compile_model (train_adipose_unet_v3.py:877-878): mean over all pixels of
// equal(y_true, cast(y_pred > 0.5)); out[0] += number of matching pixels (float64), out[1] = n is written by the host side.
__global__ void __launch_bounds__(256)
binary_accuracy_kernel(const float *__restrict__ p, const float *__restrict__ y, size_t n, double *__restrict__ out) {
  unsigned c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += (y[i] == (p[i] > 0.5f ? 1.0f : 0.0f));
  c = warp_sum_u(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (double)c);
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const float *__restrict__ p, const float *__restrict__ y, size_t n, float ys_a, float ys_b,
                 const double *__restrict__ sums, const float *__restrict__ row_w /* or null */, int W,
                 float gain /* loss weight of this output */, float *__restrict__ dldp) {
  const float eps = 1e-7f;
  const double s1 = sums[1], s2 = sums[2], s3 = sums[3], s7 = sums[7];
  const float inv_n = (float)(1.0 / s7), inter2p1 = (float)(2.0 * s1 + 1.0), denom = (float)(s2 + s3 + 1.0);
  const float inv_d2 = 1.0f / (denom * denom);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float yv = y[i] * ys_a + ys_b;
    float g = 0.f;
    if (pv >= eps && pv <= 1.0f - eps) {
      const float w = row_w ? row_w[i / (size_t)W] : 1.f;
      const float dbce = -(yv / (pv + eps) - (1.0f - yv) / (1.0f - pv + eps));
      const float ddice = -(2.0f * yv * denom - inter2p1) * inv_d2;
      g = (w * dbce * inv_n + ddice) * gain;
    }
    dldp[i] = g;
  }
}

}  // namespace adp
