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
struct TtaOps { int n; int inv[8]; };

__global__ void __launch_bounds__(256)
tta_blend_kernel(const float *__restrict__ planes, TtaOps ops, int S, int mode, float *__restrict__ out,
                 float *__restrict__ acc, float *__restrict__ wsum, const float *__restrict__ window,
                 int accW, int accRows, int ty, int tx) {
  // all (up to 8) source blocks are requested before the single barrier: one HBM round trip per block instead of one
  // per augmentation
  __shared__ float tile[8][32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int lx = threadIdx.x, ly = threadIdx.y;     // 32 x 8
  int sbi[8], sbj[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ops.n) {
      // source block origin: image of the block's (bi,bj) corner region under op
      int s0i, s0j, s1i, s1j;
      d4_src(ops.inv[k], bi, bj, S, s0i, s0j);
      d4_src(ops.inv[k], min(bi + 31, S - 1), min(bj + 31, S - 1), S, s1i, s1j);
      sbi[k] = min(s0i, s1i); sbj[k] = min(s0j, s1j);
      const float *P = planes + (size_t)k * S * S;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = sbi[k] + ly + 8 * r, j = sbj[k] + lx;
        tile[k][ly + 8 * r][lx] = (i < S && j < S) ? P[(size_t)i * S + j] : 0.f;
      }
    }
  }
  __syncthreads();
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ops.n) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = bi + ly + 8 * r, j = bj + lx;
        float v = 0.f;
        if (i < S && j < S) {
          int si, sj;
          d4_src(ops.inv[k], i, j, S, si, sj);
          v = tile[k][si - sbi[k]][sj - sbj[k]];
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
      if (gy < 0 || gy >= accRows || gx < 0 || gx >= accW) continue;
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
                  float *__restrict__ wsum, const float *__restrict__ window, int winW, int accW, int accRows,
                  int ty, int tx) {
  const size_t total = (size_t)th * tw;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    int i = p / tw, j = p % tw;
    int gy = ty + i, gx = tx + j;
    if (gy < 0 || gy >= accRows || gx < 0 || gx >= accW) continue;
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
__global__ void __launch_bounds__(256)
finalize_kernel(const float *__restrict__ acc, const float *__restrict__ wsum, int linear, size_t n, float thr,
                float *__restrict__ prob, uint8_t *__restrict__ mask, const uint8_t *__restrict__ gt,
                unsigned long long *__restrict__ counts) {
  unsigned tp = 0, fp = 0, fn = 0, tn = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float p = acc[i];
    if (wsum) {
      if (linear) {
        float c = fmaxf(wsum[i], 1.0f);
        p = (float)((double)p / (double)c);
      } else {
        p = __fdiv_rn(p, fmaxf(wsum[i], 1e-8f));
      }
    }
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
// combined_loss_standard (train_adipose_unet_v3.py:217-241) and dice_coef (src/utils/model.py:93-98).
// Pass 1: five global sums in float64: S_bce = sum bce_i, S_yp = sum y*pc, S_y = sum y, S_pc = sum pc,
//         S_yp_raw = sum y*p, S_p = sum p  (pc = clip(p, 1e-7, 1-1e-7)).
// Pass 2 (host forms the scalars): dL/dp_i = dbce_i/N + ddice_i with
//   dbce_i  = -( y/(pc+eps) - (1-y)/(1-pc+eps) ) * [eps <= p <= 1-eps]
//   ddice_i = -( 2*y*D - (2I+1) ) / D^2 * [eps <= p <= 1-eps],  I = S_yp, D = S_y + S_pc + 1
__global__ void __launch_bounds__(256)
loss_reduce_kernel(const float *__restrict__ p, const float *__restrict__ y, size_t n, double *__restrict__ sums) {
  double sb = 0, syp = 0, sy = 0, spc = 0, sypr = 0, sp = 0;
  const float eps = 1e-7f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pv = p[i], yv = y[i];
    float pc = fminf(fmaxf(pv, eps), 1.0f - eps);
    float bce = -(yv * logf(pc + eps) + (1.0f - yv) * logf(1.0f - pc + eps));
    sb += bce; syp += (double)yv * pc; sy += yv; spc += pc; sypr += (double)yv * pv; sp += pv;
  }
  sb = warp_sum_d(sb); syp = warp_sum_d(syp); sy = warp_sum_d(sy); spc = warp_sum_d(spc);
  sypr = warp_sum_d(sypr); sp = warp_sum_d(sp);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], sb); atomicAdd(&sums[1], syp); atomicAdd(&sums[2], sy);
    atomicAdd(&sums[3], spc); atomicAdd(&sums[4], sypr); atomicAdd(&sums[5], sp);
  }
}

__global__ void __launch_bounds__(256)
loss_grad_kernel(const float *__restrict__ p, const float *__restrict__ y, size_t n, float inv_n, float inter2p1,
                 float denom, float *__restrict__ dldp) {
  const float eps = 1e-7f;
  const float inv_d2 = 1.0f / (denom * denom);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pv = p[i], yv = y[i];
    float g = 0.f;
    if (pv >= eps && pv <= 1.0f - eps) {
      float dbce = -(yv / (pv + eps) - (1.0f - yv) / (1.0f - pv + eps));
      float ddice = -(2.0f * yv * denom - inter2p1) * inv_d2;
      g = dbce * inv_n + ddice;
    }
    dldp[i] = g;
  }
}

}  // namespace adp
