// kernels_train.cuh — backward pass and optimizer kernels of the training step
// (Keras train_step via net.fit, train_adipose_unet_v3.py:1316-1324; loss :217-241; Adam :801-806).
// fp32, row-planar activations (kernels_simt.cuh).  The data-gradient of a conv is the forward
// SIMT conv kernel run with flipped/transposed weights; this file holds what has no forward twin.
#pragma once
#include "kernels_simt.cuh"

namespace adp {

// ---------------------------------------------------------------------------------------------
// Head backward.  p = softmax(z)[1] = sigmoid(z1 - z0);  dz1 = g*p*(1-p), dz0 = -dz1.
// G[c] = dz1 * (w1[c] - w0[c]) * [x[c] > 0] * scale  (x = post-dropout output of up1_conv3: the written tensor is already
// dL/d(pre-activation) of that layer);  dWh[1][c] = sum dz1 * x[c] = -dWh[0][c];  dbh[1] = sum dz1 = -dbh[0].
// A warp owns one channel group (gridDim.x * 8 warps is a multiple of the group count) and walks (image, row) lines
// with its lanes along x: 8 + 1 partial sums stay in registers over all its pixels and leave through shuffles, one
// shared-memory atomic per warp and one double atomic per block and channel.  (One thread per pixel with 49 warp
// reductions per pixel reached 3.2 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_kernel(View<T> x, int nb, const float *__restrict__ wh /*[2][C]*/, const float *__restrict__ prob,
                const float *__restrict__ dldp, View<T> g, double *__restrict__ dwh1 /*[C]*/, double *__restrict__ dbh1,
                float scale) {
  extern __shared__ float sm[];
  float *wd = sm;                       // C: w1 - w0
  float *red = sm + x.C;                // C + 1 partial sums of this block
  const int C = x.C, G = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { wd[i] = wh[C + i] - wh[i]; }
  for (int i = threadIdx.x; i <= C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nw = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  const int gi = wid % G;
  float wv[8], part[8], bsum = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { wv[k] = wd[gi * 8 + k]; part[k] = 0.f; }
  const int rows = nb * x.H;
  for (int r = wid / G; r < rows; r += nw / G) {
    const int n = r / x.H, yy = r - n * x.H;
    const size_t pbase = (size_t)r * x.W;
    // U columns per trip, all their loads issued before the first use (ncu on the one-column loop: 48 % warps active,
    // stalled on the loads, 59 % of DRAM peak)
    constexpr int U = 4;
    for (int xx0 = lane; xx0 < x.W; xx0 += 32 * U) {
      float pp[U], dd[U];
      Raw8<T> raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xx = xx0 + 32 * u;
        pp[u] = 0.f; dd[u] = 0.f; raw[u] = zero_raw8<T>();
        if (xx < x.W) {
          pp[u] = prob[pbase + xx]; dd[u] = dldp[pbase + xx];
          raw[u] = load_raw8<T>(x.p + x.at(n, yy, gi, xx));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int xx = xx0 + 32 * u;
        const float dz = dd[u] * pp[u] * (1.f - pp[u]);
        float a[8], o[8];
        unpack8<T>(raw[u], a);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o[k] = a[k] > 0.f ? dz * wv[k] * scale : 0.f;   // ReLU' (and dropout mask / keep) of up1_conv3 folded in
          part[k] = fmaf(dz, a[k], part[k]);
        }
        bsum += dz;
        if (xx < x.W) store8<T>(g.p + g.at(n, yy, gi, xx), o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum(part[k]);
    if (lane == 0) atomicAdd(&red[gi * 8 + k], s);
  }
  if (gi == 0) {
    const float s = warp_sum(bsum);
    if (lane == 0) atomicAdd(&red[C], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&dwh1[i], (double)red[i]);
  if (threadIdx.x == 0) atomicAdd(dbh1, (double)red[C]);
}

// G <- G * [X > 0] * scale   (ReLU backward; scale = 1/keep at a Dropout site, where X is the
// post-dropout tensor so that [X > 0] already carries the dropout mask)
template <typename T>
__global__ void __launch_bounds__(256) relu_mask_kernel(View<T> g, View<T> x, int nb, float scale) {
  const int G = g.C / 8;
  const size_t total = (size_t)nb * g.H * G * g.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, g.W, G, g.H);
    const int xx = ri.x, gi = ri.g, yy = ri.y, n = ri.n;
    float a[8], b[8];
    load8<T>(g.p + g.at(n, yy, gi, xx), a);
    load8<T>(x.p + x.at(n, yy, gi, xx), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = b[k] > 0.f ? a[k] * scale : 0.f;
    store8<T>(g.p + g.at(n, yy, gi, xx), a);
  }
}

// Dropout forward (train_adipose_unet_v3.py:682,696,703,710): X <- X * m / keep, m ~ Bernoulli(keep)
// from a counter-based hash (TF's RNG stream cannot be reproduced; parity tests supply masks instead).
// One 32-bit murmur3-finalizer hash decides TWO channels (16 bits each: P(keep) is keep rounded to 1/65536), so a
// group of 8 channels costs four short integer hashes and the kernel stays HBM-bound.
// (hash_u32 / dropout_salt live in common.cuh: the tcgen05 conv epilogue applies the same mask when the dropout is fused)
ADP_DEVINL uint32_t dropout_bits(uint64_t seed, size_t group_index, int pair, uint32_t salt0) {
  const uint64_t idx = (uint64_t)group_index * 4u + (uint64_t)pair;
  const uint32_t hi = (uint32_t)(idx >> 32);
  return hash_u32((uint32_t)idx * 0x9E3779B1u ^ (hi == 0 ? salt0 : dropout_salt(seed, hi)));
}
// Hash path on a dense tensor whose pair index fits 32 bits (every site of the training graph): the item index IS the
// element offset / 8, so there is no coordinate decomposition, and the hashes are pure 32-bit arithmetic.  Same mask as
// dropout_kernel for the same seed.  (ncu on the general kernel: ~310 instructions per 32-byte item, 69 % issue slots, 48 % DRAM.)
template <typename T>
__global__ void __launch_bounds__(256, 8)
dropout_dense_kernel(T *__restrict__ p, uint32_t total, float keep, uint64_t seed) {
  const float inv = 1.f / keep;
  const uint32_t thr = (uint32_t)fminf(keep * 65536.f, 65536.f);
  const uint32_t salt0 = dropout_salt(seed, 0u);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    T *ptr = p + (size_t)i * 8;
    float a[8];
    unpack8<T>(load_raw8<T>(ptr), a);
    const uint32_t idx0 = i * 4u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t h = hash_u32((idx0 + (uint32_t)q) * 0x9E3779B1u ^ salt0);
      a[2 * q] = (h & 0xFFFFu) < thr ? a[2 * q] * inv : 0.f;
      a[2 * q + 1] = (h >> 16) < thr ? a[2 * q + 1] * inv : 0.f;
    }
    store8<T>(ptr, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 8)
dropout_kernel(View<T> x, int nb, float keep, uint64_t seed, const uint8_t *__restrict__ mask_in /*NHWC real channels, or null*/,
               int creal) {
  const int G = x.C / 8;
  const size_t total = (size_t)nb * x.H * G * x.W;
  const float inv = 1.f / keep;
  const uint32_t thr = (uint32_t)fminf(keep * 65536.f, 65536.f);
  const uint32_t salt0 = dropout_salt(seed, 0u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, x.W, G, x.H);
    const int xx = ri.x, gi = ri.g, yy = ri.y, n = ri.n;
    float a[8];
    T *ptr = x.p + x.at(n, yy, gi, xx);
    load8<T>(ptr, a);
    if (mask_in) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = gi * 8 + k;
        const bool keep_it = c < creal ? mask_in[(((size_t)n * x.H + yy) * x.W + xx) * creal + c] != 0 : false;
        a[k] = keep_it ? a[k] * inv : 0.f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t h = dropout_bits(seed, i, q, salt0);
        a[2 * q] = (h & 0xFFFFu) < thr ? a[2 * q] * inv : 0.f;
        a[2 * q + 1] = (h >> 16) < thr ? a[2 * q + 1] * inv : 0.f;
      }
    }
    store8<T>(ptr, a);
  }
}

// MaxPooling2D backward: the gradient of a pooled pixel goes to the FIRST maximum of its 2x2 window
// in (dy,dx) scan order.  One thread per pooled pixel and channel group; gin is ADDED into (the skip
// tensor's gradient already holds the decoder branch, already multiplied by the ReLU mask of the layer that
// produced xin — so the pooled contribution is masked the same way here).
// bf16: the nine 16-byte vectors stay packed (36 registers instead of 72 floats -> 5 resident blocks per SM instead of 3; ncu on
// the float-array form: 33 % warps active, stalled on the loads, 50 % of DRAM peak) and channels are unpacked pair by pair.
template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 2) maxpool2_bwd_kernel(View<T> xin, View<T> gout, View<T> gin, int nb) {
  const int G = gout.C / 8;
  const size_t total = (size_t)nb * gout.H * G * gout.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, gout.W, G, gout.H);
    const int xx = ri.x, gi = ri.g, yy = ri.y, n = ri.n;
    if constexpr (sizeof(T) == 2) {
      uint32_t go[4], v[4][4], gw[4][4];
      *reinterpret_cast<uint4 *>(go) = *reinterpret_cast<const uint4 *>(gout.p + gout.at(n, yy, gi, xx));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<uint4 *>(v[q]) = *reinterpret_cast<const uint4 *>(xin.p + xin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1)));
        *reinterpret_cast<uint4 *>(gw[q]) = *reinterpret_cast<const uint4 *>(gin.p + gin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1)));
      }
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {               // element 2w + h = low / high half of word w
          const int sh = 16 * h;
          float vq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) vq[q] = __uint_as_float((v[q][w] >> sh) << 16);
          int best = 0; float vb = vq[0];
#pragma unroll
          for (int q = 1; q < 4; ++q) if (vq[q] > vb) { vb = vq[q]; best = q; }
          if (vb > 0.f) {                           // gin holds dL/d(pre-activation): ReLU' of the pooled layer
            const float g = __uint_as_float((go[w] >> sh) << 16);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (q == best) {
                const float nv = __uint_as_float((gw[q][w] >> sh) << 16) + g;
                const uint32_t nb16 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(nv));
                gw[q][w] = (gw[q][w] & (0xFFFF0000u >> sh)) | (nb16 << sh);
              }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4 *>(gin.p + gin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1))) = *reinterpret_cast<uint4 *>(gw[q]);
    } else {
    float go[8], v[4][8], gi4[4][8];
    load8<T>(gout.p + gout.at(n, yy, gi, xx), go);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      load8<T>(xin.p + xin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1)), v[q]);
      load8<T>(gin.p + gin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1)), gi4[q]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int best = 0;
#pragma unroll
      for (int q = 1; q < 4; ++q) if (v[q][k] > v[best][k]) best = q;
#pragma unroll
      for (int q = 0; q < 4; ++q) if (q == best && v[q][k] > 0.f) gi4[q][k] += go[k];   // gin holds dL/d(pre-activation): ReLU' of the pooled layer
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) store8<T>(gin.p + gin.at(n, 2 * yy + (q >> 1), gi, 2 * xx + (q & 1)), gi4[q]);
    }
  }
}

// UpSampling2D backward: glow[y][x] = sum of the 2x2 block of ghigh
// x.p != null: the result is also multiplied by [x > 0] * scale (ReLU' / dropout of the layer that produced the low-res tensor)
// resid.p != null: a second gradient into the same low-res tensor (deep-supervision head) is added before the mask
template <typename T>
__global__ void __launch_bounds__(256) upsample2_bwd_kernel(View<T> ghigh, View<T> glow, int nb, View<T> x, float scale, View<T> resid) {
  const int G = glow.C / 8;
  const size_t total = (size_t)nb * glow.H * G * glow.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, glow.W, G, glow.H);
    const int xx = ri.x, gi = ri.g, yy = ri.y, n = ri.n;
    // pairwise: (column 2x: row 2y + row 2y+1) + (column 2x+1: the same) - the order the fused epilogue of the data-gradient twin
    // (conv_tc.cuh, EPI_UPSUM) can form with one shuffle per value
    float s[8], a[8], b[8];
    load8<T>(ghigh.p + ghigh.at(n, 2 * yy, gi, 2 * xx), a);
    load8<T>(ghigh.p + ghigh.at(n, 2 * yy + 1, gi, 2 * xx), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = a[k] + b[k];
    load8<T>(ghigh.p + ghigh.at(n, 2 * yy, gi, 2 * xx + 1), a);
    load8<T>(ghigh.p + ghigh.at(n, 2 * yy + 1, gi, 2 * xx + 1), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += a[k] + b[k];
    if (resid.p) {
      load8<T>(resid.p + resid.at(n, yy, gi, xx), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += a[k];
    }
    if (x.p) {
      load8<T>(x.p + x.at(n, yy, gi, xx), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = a[k] > 0.f ? s[k] * scale : 0.f;
    }
    store8<T>(glow.p + glow.at(n, yy, gi, xx), s);
  }
}

// dst (view) <- a (view) + b (view)   — gradient fan-in of the Add / skip connections
template <typename T>
__global__ void __launch_bounds__(256) add_views_kernel(View<T> dst, View<T> a, View<T> b, int nb) {
  const int G = dst.C / 8;
  const size_t total = (size_t)nb * dst.H * G * dst.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, dst.W, G, dst.H);
    const int xx = ri.x, gi = ri.g, yy = ri.y, n = ri.n;
    float u[8], v[8];
    load8<T>(a.p + a.at(n, yy, gi, xx), u);
    load8<T>(b.p + b.at(n, yy, gi, xx), v);
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] += v[k];
    store8<T>(dst.p + dst.at(n, yy, gi, xx), u);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient of a 3x3 conv:  dW[t][ci][co] = sum_{n,y,x} Xin[n, y+dy_t, x+dx_t, ci] * dZ[n,y,x,co]
// (zero outside the image; UP: Xin is the nearest-upsampled view of xin), db[co] = sum dZ.
// grid = (9 taps, ci tiles x co tiles (64 x 64), pixel splits); block = 256 threads, each a 4x4
// register tile; partial sums leave through fp32 atomics into dW[9][cin_pad][cout_pad].
template <typename T, bool UP>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(View<T> xin, View<T> dz, float *__restrict__ dW, float *__restrict__ db, int dil, int nb, int cin_pad,
                  int cout_pad, int co_tiles) {
  __shared__ float xs[64][33];
  __shared__ float zs[64][33];
  const int t = blockIdx.x;
  const int ci0 = (blockIdx.y / co_tiles) * 64, co0 = (blockIdx.y % co_tiles) * 64;
  const int H = dz.H, W = dz.W;
  const int dy = (t / 3 - 1) * dil, dx = (t % 3 - 1) * dil;
  const int tid = threadIdx.x;
  const int tci = tid >> 4, tco = tid & 15;
  const int lp = tid & 31, lplane = tid >> 5;          // loader role: pixel in chunk, channel group in tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0, 0, 0, 0};
  const int chunks_per_row = (W + 31) / 32;
  const long long nchunks = (long long)nb * H * chunks_per_row;
  for (long long ch = blockIdx.z; ch < nchunks; ch += gridDim.z) {
    const int cx = (int)(ch % chunks_per_row) * 32;
    long long r = ch / chunks_per_row;
    const int y = (int)(r % H), n = (int)(r / H);
    __syncthreads();
    {
      const int x = cx + lp;
      float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const int iy = y + dy, ix = x + dx;
      const int gci = (ci0 >> 3) + lplane, gco = (co0 >> 3) + lplane;
      if (x < W && iy >= 0 && iy < H && ix >= 0 && ix < W && gci * 8 < cin_pad)
        load8<T>(xin.p + xin.at(n, UP ? (iy >> 1) : iy, gci, UP ? (ix >> 1) : ix), a);
      if (x < W && gco * 8 < cout_pad) load8<T>(dz.p + dz.at(n, y, gco, x), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) { xs[lplane * 8 + k][lp] = a[k]; zs[lplane * 8 + k][lp] = b[k]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int p = 0; p < 32; ++p) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = xs[tci * 4 + i][p]; b[i] = zs[tco * 4 + i][p]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (tci == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += b[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + tci * 4 + i;
    if (ci >= cin_pad) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tco * 4 + j;
      if (co < cout_pad) atomicAdd(&dW[((size_t)t * cin_pad + ci) * cout_pad + co], acc[i][j]);
    }
  }
  if (db && t == 4 && ci0 == 0 && tci == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tco * 4 + j;
      if (co < cout_pad) atomicAdd(&db[co], bsum[j]);
    }
  }
}

// First layer (Cin = 1) weight gradient: dW[t][co] = sum x_norm(p + tap) * dZ[p][co], db[co] = sum dZ.
// xnorm: [nb][S][S] float32 (already normalised).  A warp owns 32 consecutive pixels of a row and ONE channel
// group (gridDim.x * 8 warps is a multiple of the group count), keeps its 9x8 + 8 sums in registers over all its
// pixels and leaves through shuffles + one atomic per sum.
template <typename T>
__global__ void __launch_bounds__(256, 2)
first_wgrad_kernel(const float *__restrict__ xnorm, View<T> dz, int nb, float *__restrict__ dW /*[9][C]*/, float *__restrict__ db) {
  const int C = dz.C, S = dz.H, G = C / 8;
  const int lane = threadIdx.x & 31;
  const int wid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nw = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  const int g = wid % G;                        // this warp's channel group (gridDim.x * 8 is a multiple of G)
  const int slot = wid / G, nslots = nw / G;
  const int xchunks = (S + 31) / 32;
  const int nrc = nb * S * xchunks;             // (image, row, 32-pixel chunk) work items of one channel group
  float acc[9][8], bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    bs[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t][k] = 0.f;
  }
  // A warp walks a CONTIGUOUS range of (image, row, 32-pixel chunk) items, so coordinates advance by increments (no
  // divisions per item) and consecutive chunks re-use the input rows in L1; interior chunks (warp-uniform test) read their
  // nine input taps without bounds checks.  ncu on the strided, always-checked form: 227 instructions per item for 80 FMAs,
  // 45 % issue slots at 25 % occupancy.  The 16-byte gradient load of the item PD steps ahead is issued into the register
  // slot the current item has just been consumed from, so its HBM latency overlaps the FMAs of PD items.
  constexpr int PD = 2;
  const int per = (nrc + nslots - 1) / nslots;
  const int rc_begin = min(nrc, slot * per), rc_end = min(nrc, rc_begin + per);
  int cxc = rc_begin % xchunks, cy = (rc_begin / xchunks) % S, cn = (rc_begin / xchunks) / S;    // item being consumed
  int pxc = cxc, py = cy, pn = cn;                                                                // item being prefetched
  float d[PD][8];
  auto fetch = [&](int j, bool live) {
    const int x = pxc * 32 + lane;
    if (live && x < S) load8<T>(dz.p + dz.at(pn, py, g, x), d[j]);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) d[j][k] = 0.f;
    }
    if (++pxc == xchunks) { pxc = 0; if (++py == S) { py = 0; ++pn; } }
  };
#pragma unroll
  for (int j = 0; j < PD; ++j) fetch(j, rc_begin + j < rc_end);
  for (int rc = rc_begin; rc < rc_end; rc += PD) {
#pragma unroll
    for (int j = 0; j < PD; ++j) {
      if (rc + j < rc_end) {
        const int x = cxc * 32 + lane;
        float v[9];
        const float *xc0 = xnorm + ((size_t)cn * S + cy) * S + x;
        if (cy >= 1 && cy < S - 1 && cxc >= 1 && cxc * 32 + 32 < S) {
#pragma unroll
          for (int t = 0; t < 9; ++t) v[t] = __ldg(xc0 + (t / 3 - 1) * S + (t % 3 - 1));
        } else {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int iy = cy + t / 3 - 1, ix = x + t % 3 - 1;
            v[t] = (iy >= 0 && iy < S && ix >= 0 && ix < S) ? __ldg(xc0 + (t / 3 - 1) * S + (t % 3 - 1)) : 0.f;
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) bs[k] += d[j][k];
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[t][k] = fmaf(v[t], d[j][k], acc[t][k]);
        if (++cxc == xchunks) { cxc = 0; if (++cy == S) { cy = 0; ++cn; } }
        fetch(j, rc + j + PD < rc_end);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum(bs[k]);
    if (lane == 0) atomicAdd(&db[g * 8 + k], s);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float a = warp_sum(acc[t][k]);
      if (lane == 0) atomicAdd(&dW[t * C + g * 8 + k], a);
    }
  }
}

// First-layer weight gradient on the tensor cores (bf16 training path): per 16 pixels ONE mma.sync.m16n8k16
//   D[16 x 8] += A[16 x 16] * B[16 x 8],   rows of A = the nine input taps (row 9 = ones, so D row 9 is the bias
//   gradient; rows 10..15 zero), columns of A = 16 consecutive pixels, B = dZ of those pixels for the 8 channels of one
//   channel group, D rows 0..8 = dW[tap][8 channels] in fp32.
// Two MMAs per channel group replace the 80 FMAs per lane, 32-pixel chunk and channel group of first_wgrad_kernel (which
// stays the exact-fp32 / CUDA-core cross-check path), with 4 accumulator registers per group instead of 80.  The input
// pixels enter the MMA rounded to bf16 (dZ already is bf16); accumulation is fp32.
// Fragment layout of mma.m16n8k16 (PTX ISA): group = lane / 4, t = lane % 4;
//   A regs: {row group, cols 2t,2t+1}, {row group + 8, same cols}, {row group, cols 2t+8,2t+9}, {row group + 8, cols 2t+8,2t+9}
//   B regs: {k = 2t,2t+1; n = group}, {k = 2t+8,2t+9; n = group};  D: {row group, cols 2t,2t+1}, {row group + 8, same cols}.
ADP_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}
// A warp owns a contiguous range of (image, row, 32-pixel chunk) items for ALL channel groups: the A fragments (input taps)
// are built once per chunk and feed one MMA per channel group (a warp per channel group re-read the input rows G times
// through L2: 0.73 ms).  dZ streams through a per-warp ring of kFwStages slots of G x 512 bytes in shared memory, filled
// with cp.async (16 bytes per lane = one pixel's channel group) kFwStages - 1 chunks ahead: shared memory, not registers,
// holds the bytes in flight.
constexpr int kFwStages = 3;
constexpr int kFwMaxG = 8;                                   // init_nb <= 64 channels
inline size_t first_wgrad_mma_smem(int G) { return (size_t)8 * kFwStages * G * 512; }
ADP_DEVINL void cp_async16_zfill(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__global__ void __launch_bounds__(256)
first_wgrad_mma_kernel(const float *__restrict__ xnorm, View<__nv_bfloat16> dz, int nb, float *__restrict__ dW /*[9][C]*/,
                       float *__restrict__ db) {
  __shared__ float red[10][64];
  extern __shared__ __align__(16) uint8_t zring[];           // [8 warps][kFwStages][G][32 pixels][8 channels] bf16
  const int C = dz.C, S = dz.H, G = C / 8;
  for (int i = threadIdx.x; i < 10 * 64; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, grp = lane >> 2, t4 = lane & 3;
  const int wid = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nw = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  const int xchunks = (S + 31) / 32;
  const int nrc = nb * S * xchunks;
  const int per = (nrc + nw - 1) / nw;
  const int rc_begin = min(nrc, wid * per), rc_end = min(nrc, rc_begin + per);
  int cxc = rc_begin % xchunks, cy = (rc_begin / xchunks) % S, cn = (rc_begin / xchunks) / S;    // chunk being consumed
  int pxc = cxc, py = cy, pn = cn;                                                                // chunk being prefetched
  const uint32_t stage_bytes = (uint32_t)G * 512u;
  uint8_t *wz = zring + (size_t)(threadIdx.x >> 5) * kFwStages * stage_bytes;
  const uint32_t wz_u32 = (uint32_t)__cvta_generic_to_shared(wz);
  auto prefetch = [&](int stage, bool live) {
    const int p = pxc * 32 + lane;
    const bool ok = live && p < S;
    const __nv_bfloat16 *src = ok ? dz.p + dz.at(pn, py, 0, p) : dz.p;
    const size_t gstride = (size_t)dz.W * 8;                 // next channel group of the same pixel row
    for (int gg = 0; gg < G; ++gg)
      cp_async16_zfill(wz_u32 + (uint32_t)stage * stage_bytes + (uint32_t)(gg * 512 + lane * 16), ok ? (const void *)(src + gg * gstride) : (const void *)dz.p,
                       ok ? 16u : 0u);                       // 0 source bytes = zero fill
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    if (++pxc == xchunks) { pxc = 0; if (++py == S) { py = 0; ++pn; } }
  };
#pragma unroll
  for (int j = 0; j < kFwStages - 1; ++j) prefetch(j, rc_begin + j < rc_end);
  const int dy = grp / 3 - 1, dx = grp % 3 - 1;             // tap of A row `grp` (grp = 0..7); row 8 = tap (+1,+1)
  float c[kFwMaxG][4];
#pragma unroll
  for (int gg = 0; gg < kFwMaxG; ++gg) { c[gg][0] = 0.f; c[gg][1] = 0.f; c[gg][2] = 0.f; c[gg][3] = 0.f; }
  int st = 0;                                               // ring slot of the chunk being consumed
  for (int rc = rc_begin; rc < rc_end; ++rc) {
    prefetch(st == 0 ? kFwStages - 1 : st - 1, rc + kFwStages - 1 < rc_end);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(kFwStages - 1) : "memory");
    __syncwarp();
    const float *xrow = xnorm + ((size_t)cn * S + cy) * S;
    const bool interior = cy >= 1 && cy < S - 1 && cxc >= 1 && cxc * 32 + 32 < S;
    uint32_t a[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int px = cxc * 32 + 16 * s + 2 * t4;             // this lane's pixel pairs (px, px+1) and (px+8, px+9)
      float xa[4], xb[4];                                    // taps `grp` and 8 at those four pixels
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int p = px + (q & 1) + 8 * (q >> 1);
        if (interior) {
          xa[q] = __ldg(xrow + dy * S + p + dx);
          xb[q] = grp == 0 ? __ldg(xrow + S + p + 1) : 0.f;
        } else {
          const int ay = cy + dy, ax = p + dx;
          xa[q] = (p < S && ay >= 0 && ay < S && ax >= 0 && ax < S) ? __ldg(xrow + dy * S + ax) : 0.f;
          xb[q] = (grp == 0 && p < S && cy + 1 < S && p + 1 < S) ? __ldg(xrow + S + p + 1) : 0.f;
        }
      }
      a[s][0] = pack_bf16x2(xa[0], xa[1]); a[s][2] = pack_bf16x2(xa[2], xa[3]);
      a[s][1] = 0u; a[s][3] = 0u;
      if (grp == 0) { a[s][1] = pack_bf16x2(xb[0], xb[1]); a[s][3] = pack_bf16x2(xb[2], xb[3]); }
      else if (grp == 1) { a[s][1] = 0x3F803F80u; a[s][3] = 0x3F803F80u; }         // row 9 = ones -> bias gradient
    }
    const unsigned short *zs = reinterpret_cast<const unsigned short *>(wz + (size_t)st * stage_bytes) + grp;   // [G][32 pixels][8 channels]
#pragma unroll
    for (int gg = 0; gg < kFwMaxG; ++gg) {
      if (gg < G) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const unsigned short *zp = zs + gg * 256 + (16 * s + 2 * t4) * 8;       // zero-filled beyond the row end
          const uint32_t b0 = (uint32_t)zp[0] | ((uint32_t)zp[8] << 16), b1 = (uint32_t)zp[64] | ((uint32_t)zp[72] << 16);
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                       : "+f"(c[gg][0]), "+f"(c[gg][1]), "+f"(c[gg][2]), "+f"(c[gg][3])
                       : "r"(a[s][0]), "r"(a[s][1]), "r"(a[s][2]), "r"(a[s][3]), "r"(b0), "r"(b1));
        }
      }
    }
    __syncwarp();                                            // the slot is refilled by the next iteration's prefetch
    if (++st == kFwStages) st = 0;
    if (++cxc == xchunks) { cxc = 0; if (++cy == S) { cy = 0; ++cn; } }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  // D rows: grp -> tap grp (c0, c1), grp + 8 -> tap 8 (grp == 0) or bias (grp == 1) (c2, c3); columns = channels 2*t4, 2*t4 + 1
#pragma unroll
  for (int gg = 0; gg < kFwMaxG; ++gg) {
    if (gg < G) {
      const int ch = gg * 8 + 2 * t4;
      atomicAdd(&red[grp][ch], c[gg][0]); atomicAdd(&red[grp][ch + 1], c[gg][1]);
      if (grp < 2) { atomicAdd(&red[8 + grp][ch], c[gg][2]); atomicAdd(&red[8 + grp][ch + 1], c[gg][3]); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) {
    const int r = i / C, cc = i - r * C;
    const float v = red[r][cc];
    if (v != 0.f) atomicAdd(r < 9 ? &dW[r * C + cc] : &db[cc], v);
  }
}

// ---------------------------------------------------------------------------------------------
// Keras 2.13 Adam / AdamW update_step on the flat parameter buffer (epsilon OUTSIDE the bias correction):
//   [AdamW] theta -= theta * wd * lr
//   m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);  theta -= (m * alpha) / (sqrt(v) + eps),
//   alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)  (computed on the host in float32 like Keras does)
// trainable: per-parameter-segment flags are applied by launching only over trainable ranges.
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ theta, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, size_t n,
            float gscale, float alpha, float one_minus_b1, float one_minus_b2, float eps, float wd, float lr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float th = theta[i];
    const float gi = g[i] * gscale;
    if (wd != 0.f) th = __fsub_rn(th, __fmul_rn(__fmul_rn(th, wd), lr));
    float mi = m[i], vi = v[i];
    mi = __fadd_rn(mi, __fmul_rn(__fsub_rn(gi, mi), one_minus_b1));
    vi = __fadd_rn(vi, __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), vi), one_minus_b2));
    th = __fsub_rn(th, __fdiv_rn(__fmul_rn(mi, alpha), __fadd_rn(__fsqrt_rn(vi), eps)));
    theta[i] = th; m[i] = mi; v[i] = vi;
  }
}

// flat HWIO parameter segment <-> padded [9][cin_pad][cout_pad] kernel layout (concat layers: the second
// half of the real input channels starts at pad16(skip))
__global__ void __launch_bounds__(256)
pad_kernel_weights(const float *__restrict__ flat, float *__restrict__ padded, int taps, int cin, int cout, int cin_pad,
                   int cout_pad, int skip, int skip_pad, int to_flat, int transpose_flip) {
  const size_t total = (size_t)taps * cin * cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = i % cout; size_t r = i / cout;
    const int ci = r % cin; const int t = r / cin;
    const int cip = (skip && ci >= skip) ? skip_pad + (ci - skip) : ci;
    size_t pi;
    if (transpose_flip) pi = ((size_t)(taps - 1 - t) * cout_pad + co) * cin_pad + cip;   // dgrad weights [t'][co][ci]
    else pi = ((size_t)t * cin_pad + cip) * cout_pad + co;
    if (to_flat) const_cast<float *>(flat)[i] = padded[pi];
    else padded[pi] = flat[i];
  }
}

}  // namespace adp

namespace adp {

// head gradient sums (double) -> flat float gradients of output_softmax: kernel (1,1,C,2), bias (2)
__global__ void head_grad_finish_kernel(const double *__restrict__ acc /*[C+1]*/, int C, int acc_c, float *__restrict__ gk, float *__restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { const float v = (float)acc[i]; gk[i * 2 + 1] = v; gk[i * 2] = -v; }
  if (i == 0) { const float v = (float)acc[acc_c]; gb[1] = v; gb[0] = -v; }
}

// flat output_softmax kernel (C,2) -> [2][Cpad]
__global__ void head_pack_kernel(const float *__restrict__ flat, int C, int Cpad, float *__restrict__ wh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) { wh[i] = flat[i * 2]; wh[Cpad + i] = flat[i * 2 + 1]; }
}

__global__ void round_bf16_kernel(float *__restrict__ p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = __bfloat162float(__float2bfloat16_rn(p[i]));
}

// Device twin of the host packer of the tcgen05 B operand (engine.cu pack_layer): from padded fp32
// weights wp[9][cin_pad][cout_pad] to bf16 blocks [variant][chunk][tap][2 planes][N][8].
// up != 0: variant = output parity, tap = (ry,rx) of the 2x2 source window, value = sum of the 3x3 taps
// that read the same source pixel (summed in fp32, rounded once).
__global__ void __launch_bounds__(256)
pack_tc_kernel(const float *__restrict__ wp, __nv_bfloat16 *__restrict__ dst, int nvar, int nchunks, int ntaps, int N,
               int cin_pad, int cout_pad, int up, int kys) {
  const size_t blk = (size_t)16 * N;
  const size_t total = (size_t)nvar * nchunks * ntaps * blk;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = i % 8; size_t r = i / 8;
    const int n = r % N; r /= N;
    const int g = r % 2; r /= 2;
    const int t = r % ntaps; r /= ntaps;
    const int c = r % nchunks; const int v = r / nchunks;
    const int ci = c * 16 + g * 8 + j;
    float val = 0.f;
    if (ci < cin_pad) {
      if (up) {
        const int py = v >> 1, px = v & 1, ry = t >> 1, rx = t & 1;
        for (int ky = 0; ky < 3; ++ky) {
          const int sy = (py + ky - 1) >= 0 ? (py + ky - 1) / 2 : -1;
          if (sy - (py - 1) != ry) continue;
          for (int kx = 0; kx < 3; ++kx) {
            const int sx = (px + kx - 1) >= 0 ? (px + kx - 1) / 2 : -1;
            if (sx - (px - 1) != rx) continue;
            val += wp[((size_t)(ky * 3 + kx) * cin_pad + ci) * cout_pad + n];
          }
        }
      } else {
        val = wp[((size_t)t * cin_pad + ci) * cout_pad + v * N + n];
      }
    }
    // i enumerates the plain layout; the ky-stacked layout permutes positions inside the (variant, chunk) block
    const size_t o = kys ? ((size_t)v * nchunks + c) * ntaps * blk + tc_block_index(1, ntaps, N, t, g, n, j) : i;
    dst[o] = __float2bfloat16_rn(val);
  }
}

// ---------------------------------------------------------------------------------------------
// Deep supervision (train_adipose_unet_v3.py:712-745): aux_out = sigmoid(Conv2D(1, 1x1)(up)) at the decoder level's
// resolution, tf.image.resize(..., [S, S], 'bilinear') (half-pixel centres, no antialias) to full resolution.
// Forward: one thread per low-res pixel.  w: [C] (zero in pad channels), a_low: [nb][h][w] fp32.
template <typename T>
__global__ void __launch_bounds__(256)
aux_head_fwd_kernel(View<T> x, int nb, const float *__restrict__ w /*[Creal]*/, const float *__restrict__ b, int creal,
                    float *__restrict__ a_low) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < x.C; i += blockDim.x) sm[i] = i < creal ? w[i] : 0.f;
  __syncthreads();
  const size_t total = (size_t)nb * x.H * x.W;
  for (size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x; px < total; px += (size_t)gridDim.x * blockDim.x) {
    const int xx = px % x.W; size_t r = px / x.W; const int yy = r % x.H, n = r / x.H;
    float z = b[0];
    for (int gi = 0; gi < x.C / 8; ++gi) {
      float a[8];
      load8<T>(x.p + x.at(n, yy, gi, xx), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) z = fmaf(a[k], sm[gi * 8 + k], z);
    }
    a_low[px] = 1.f / (1.f + expf(-z));
  }
}

// source index pair and lerp weight of output index o for scale = in/out (TF2 / half_pixel_centers bilinear)
ADP_DEVINL void bilinear_src(int o, float scale, int n_in, int &lo, int &hi, float &t) {
  const float in = ((float)o + 0.5f) * scale - 0.5f;
  const float fl = floorf(in);
  lo = max((int)fl, 0);
  hi = min((int)ceilf(in), n_in - 1);
  t = in - fl;
}

__global__ void __launch_bounds__(256)
bilinear_up_kernel(const float *__restrict__ a_low, int nb, int h, int S, float *__restrict__ a_full) {
  const float scale = (float)h / (float)S;
  const size_t total = (size_t)nb * S * S;
  for (size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x; px < total; px += (size_t)gridDim.x * blockDim.x) {
    const int xx = px % S; size_t r = px / S; const int yy = r % S, n = r / S;
    int y0, y1, x0, x1; float ty, tx;
    bilinear_src(yy, scale, h, y0, y1, ty);
    bilinear_src(xx, scale, h, x0, x1, tx);
    const float *A = a_low + (size_t)n * h * h;
    const float top = A[(size_t)y0 * h + x0] + (A[(size_t)y0 * h + x1] - A[(size_t)y0 * h + x0]) * tx;
    const float bot = A[(size_t)y1 * h + x0] + (A[(size_t)y1 * h + x1] - A[(size_t)y1 * h + x0]) * tx;
    a_full[px] = top + (bot - top) * ty;
  }
}

// Adjoint of bilinear_up_kernel, gather form (deterministic): low-res pixel (jy, jx) collects w * g_full over the
// full-res pixels whose interpolation reads it; the weights come from the same bilinear_src, so border clamping is
// handled by construction.  Output is multiplied by `gain` (the loss weight of the auxiliary output).
__global__ void __launch_bounds__(256)
bilinear_up_bwd_kernel(const float *__restrict__ g_full, int nb, int h, int S, float gain, float *__restrict__ g_low) {
  const float scale = (float)h / (float)S;
  const int f = S / h;                               // integer upsampling factor (4 or 2)
  const size_t total = (size_t)nb * h * h;
  for (size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x; px < total; px += (size_t)gridDim.x * blockDim.x) {
    const int jx = px % h; size_t r = px / h; const int jy = r % h, n = r / h;
    const float *G = g_full + (size_t)n * S * S;
    float acc = 0.f;
    const int iy_lo = max(f * (jy - 1), 0), iy_hi = min(f * (jy + 2) - 1, S - 1);
    const int ix_lo = max(f * (jx - 1), 0), ix_hi = min(f * (jx + 2) - 1, S - 1);
    for (int iy = iy_lo; iy <= iy_hi; ++iy) {
      int y0, y1; float ty;
      bilinear_src(iy, scale, h, y0, y1, ty);
      const float wy = (y0 == jy ? 1.f - ty : 0.f) + (y1 == jy ? ty : 0.f);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int ix = ix_lo; ix <= ix_hi; ++ix) {
        int x0, x1; float tx;
        bilinear_src(ix, scale, h, x0, x1, tx);
        const float wx = (x0 == jx ? 1.f - tx : 0.f) + (x1 == jx ? tx : 0.f);
        if (wx != 0.f) row = fmaf(wx, G[(size_t)iy * S + ix], row);
      }
      acc = fmaf(wy, row, acc);
    }
    g_low[px] = acc * gain;
  }
}

// Backward of the 1x1 sigmoid head: dz = g_low * a * (1 - a); R[c] = dz * w[c] is written in the layout of the tensor's
// gradient (added by the data-gradient epilogue of the conv that consumes the tensor); dW[c] = sum dz * x[c], db = sum dz
// (block reduction, one double atomic per block and channel).
template <typename T>
__global__ void __launch_bounds__(256)
aux_head_bwd_kernel(View<T> x, int nb, const float *__restrict__ w, int creal, const float *__restrict__ a_low,
                    const float *__restrict__ g_low, View<T> resid, double *__restrict__ dwb /*[C + 1]*/) {
  extern __shared__ float sm[];
  float *wd = sm;                       // C
  float *red = sm + x.C;                // C + 1
  const int C = x.C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) wd[i] = i < creal ? w[i] : 0.f;
  for (int i = threadIdx.x; i <= C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const size_t total = (size_t)nb * x.H * x.W;
  const size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  float dz = 0.f;
  int xx = 0, yy = 0, n = 0;
  const bool live = px < total;
  if (live) {
    xx = px % x.W; size_t r = px / x.W; yy = r % x.H; n = r / x.H;
    const float a = a_low[px];
    dz = g_low[px] * a * (1.f - a);
  }
  for (int gi = 0; gi < C / 8; ++gi) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
    if (live) load8<T>(x.p + x.at(n, yy, gi, xx), a);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o[k] = dz * wd[gi * 8 + k];
      const float s = warp_sum(dz * a[k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&red[gi * 8 + k], s);
    }
    if (live) store8<T>(resid.p + resid.at(n, yy, gi, xx), o);
  }
  {
    const float s = warp_sum(dz);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[C], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= C; i += blockDim.x) atomicAdd(&dwb[i], (double)red[i]);
}

__global__ void aux_grad_finish_kernel(const double *__restrict__ dwb, int creal, int C, float *__restrict__ gk, float *__restrict__ gb) {
  for (int i = threadIdx.x; i < creal; i += blockDim.x) gk[i] = (float)dwb[i];
  if (threadIdx.x == 0) gb[0] = (float)dwb[C];
}

}  // namespace adp
