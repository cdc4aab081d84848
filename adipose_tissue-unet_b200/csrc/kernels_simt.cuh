// kernels_simt.cuh — CUDA-core kernels of the U-Net path: first conv (Cin=1, fused z-score +
// dihedral load), generic 3x3 conv (exact-fp32 parity path and bf16 cross-check), max-pool,
// six-way add, 1x1 softmax head.
//
// Activation layout in HBM ("row-planar"): [image][row y][channel group cg][column x][8 channels],
// channel count padded to a multiple of 16 (pad channels are exact zeros).  A warp that owns 32
// consecutive pixels of a row reads/writes 512 contiguous bytes per channel group, and a TMA box
// of 8 pixels x 8 channels is one 128-byte line — the layout is chosen for the tcgen05 conv's
// operand staging (conv_tc.cuh) and for coalesced epilogues, not for the reference's NHWC.
// `cgs` is the number of channel groups of the underlying buffer and `cg0` the first group of the
// view, so concat buffers are written in place (Concatenate([skip, up]) of
// train_adipose_unet_v3.py:693,700,707 costs nothing).
#pragma once
#include "common.cuh"

namespace adp {

template <typename T> struct View {
  T *p;
  int H, W;      // spatial size of one image
  int cgs;       // channel groups (of 8) per pixel row in the underlying buffer
  int cg0;       // first channel group of this view
  int C;         // channels of the view (multiple of 16)
  int lo;        // bf16x3 precision: channel-group distance from a value's hi half to its lo half (0 = plain tensor)
  __host__ __device__ size_t img_stride() const { return (size_t)H * cgs * W * 8; }
  // element offset of (image n, row y, channel group g of the view, column x)
  __host__ __device__ size_t at(int n, int y, int g, int x) const {
    return ((((size_t)n * H + y) * cgs + cg0 + g) * W + x) * 8;
  }
};

// ---------------------------------------------------------------------------------------------
// First layer: Reshape + Conv2D(1 -> C, 3x3, same) + ReLU  (train_adipose_unet_v3.py:665-668),
// fused with predict_single's normalisation x = (img - mean) / (std + 1e-10) in float32
// (full_evaluation_enhanced.py:1306) and the TTA input transform aug(img)
// (full_evaluation_enhanced.py:590).  One thread per output pixel.
// src is float32 or uint8 (1 or 3 interleaved channels -> OpenCV gray).
struct FirstConvSrc {
  const float *f32;     // n_tiles * S * S, or null
  const uint8_t *u8;    // n_tiles * S * S * ch, or null (ch = 1 or 3); or a slide region
  int ch;
  // slide mode (u8, ch = 1 or 3 interleaved): tile t starts at PIXEL slide_origin[t] = y*slideW + x of the region
  const int64_t *slide_origin;
  int slideW;           // row pitch in pixels of the slide region (0 = packed tiles)
};

ADP_DEVINL float first_conv_fetch(const FirstConvSrc &s, int tile, int S, int si, int sj) {
  if (s.f32) return s.f32[((size_t)tile * S + si) * S + sj];
  const size_t px = s.slideW ? (size_t)(s.slide_origin[tile] + (int64_t)si * s.slideW + sj) : ((size_t)tile * S + si) * S + sj;
  if (s.ch == 1) return (float)s.u8[px];
  const uint8_t *q = s.u8 + px * 3;
  int y = (9798 * (int)q[0] + 19235 * (int)q[1] + 3735 * (int)q[2] + 16384) >> 15;   // OpenCV 4.x RGB2GRAY, 8-bit
  return (float)y;
}

template <typename T> ADP_DEVINL void store8(T *o, const float *a) {
  if constexpr (sizeof(T) == 2) {
    __align__(16) __nv_bfloat16 h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = __float2bfloat16_rn(a[c]);
    *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<uint4 *>(h);
  } else {
    *reinterpret_cast<float4 *>(o) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4 *>(o + 4) = make_float4(a[4], a[5], a[6], a[7]);
  }
}
// hi/lo access of the bf16x3 precision: value = hi + lo; a store splits v into hi = bf16(v), lo = bf16(v - hi)
template <typename T> ADP_DEVINL void load8(const T *i, float *a);
template <typename T> ADP_DEVINL void load8v(const View<T> &v, int n, int y, int g, int x, float *a) {
  const T *p = v.p + v.at(n, y, g, x);
  load8<T>(p, a);
  if (v.lo) {
    float l[8];
    load8<T>(p + (size_t)v.lo * v.W * 8, l);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] += l[c];
  }
}
template <typename T> ADP_DEVINL void store8v(const View<T> &v, int n, int y, int g, int x, const float *a) {
  T *p = v.p + v.at(n, y, g, x);
  store8<T>(p, a);
  if constexpr (sizeof(T) == 2) {
    if (v.lo) {
      float l[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) l[c] = a[c] - __bfloat162float(__float2bfloat16_rn(a[c]));
      store8<T>(p + (size_t)v.lo * v.W * 8, l);
    }
  }
}
template <typename T> ADP_DEVINL void load8(const T *i, float *a) {
  if constexpr (sizeof(T) == 2) {
    __align__(16) __nv_bfloat16 h[8];
    *reinterpret_cast<uint4 *>(h) = *reinterpret_cast<const uint4 *>(i);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = __bfloat162float(h[c]);
  } else {
    float4 lo = *reinterpret_cast<const float4 *>(i), hi = *reinterpret_cast<const float4 *>(i + 4);
    a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
  }
}

// eight channels as loaded (16 bytes of bf16, or two float4): lets a kernel keep several loads in flight without holding
// their unpacked float copies in registers
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 v; };
template <> struct Raw8<float> { float4 a, b; };
template <typename T> ADP_DEVINL Raw8<T> load_raw8(const T *p) {
  Raw8<T> r;
  if constexpr (sizeof(T) == 2) r.v = *reinterpret_cast<const uint4 *>(p);
  else { r.a = *reinterpret_cast<const float4 *>(p); r.b = *reinterpret_cast<const float4 *>(p + 4); }
  return r;
}
template <typename T> ADP_DEVINL Raw8<T> zero_raw8() {
  Raw8<T> r;
  if constexpr (sizeof(T) == 2) r.v = make_uint4(0u, 0u, 0u, 0u);
  else { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
  return r;
}
template <typename T> ADP_DEVINL void unpack8(const Raw8<T> &r, float *a) {
  if constexpr (sizeof(T) == 2) {
    const uint32_t w[4] = {r.v.x, r.v.y, r.v.z, r.v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) { a[2 * c] = __uint_as_float(w[c] << 16); a[2 * c + 1] = __uint_as_float(w[c] & 0xFFFF0000u); }
  } else {
    a[0] = r.a.x; a[1] = r.a.y; a[2] = r.a.z; a[3] = r.a.w; a[4] = r.b.x; a[5] = r.b.y; a[6] = r.b.z; a[7] = r.b.w;
  }
}

// Packed dual fp32 FMA of sm_100 (SASS FFMA2): two independent IEEE round-to-nearest FMAs per instruction, i.e. the same bits
// as two fmaf() calls in half the issue slots - a three-register FFMA issues every other cycle per scheduler, which made the
// first conv issue-bound (ncu: 70 % issue slots at 39 % of DRAM peak).
ADP_DEVINL void ffma2(float2 &acc, const float2 a, const float2 b) {
  unsigned long long &d = reinterpret_cast<unsigned long long &>(acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

// Input of the fused first conv (conv_tc.cuh, FC variant): out[f][y][x] = (aug_f(img)[y][x] - mean) / (std + 1e-10) in float32,
// the statements of first_conv_kernel's window fill (full_evaluation_enhanced.py:1306, :590).  One thread = 4 consecutive x.
__global__ void __launch_bounds__(256)
tta_input_kernel(FirstConvSrc src, const int *__restrict__ fw_tile, const int *__restrict__ fw_op, int S, float mean_f, float sd_f,
                 float *__restrict__ out, int nfw) {
  const int S4 = S >> 2;
  const size_t total = (size_t)nfw * S * S4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % S4); size_t q = i / S4;
    const int y = (int)(q % S), f = (int)(q / S);
    const int tile = fw_tile[f], op = fw_op[f];
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int si, sj;
      d4_src(op, y, x4 * 4 + k, S, si, sj);
      v[k] = __fdiv_rn(__fsub_rn(first_conv_fetch(src, tile, S, si, sj), mean_f), sd_f);
    }
    *reinterpret_cast<float4 *>(out + ((size_t)f * S + y) * S + x4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// block = (32, 8), output tile = 32 columns x 32 rows: the normalised (and dihedrally transformed) 34 x 34 input window
// is staged once in shared memory (one z-score division per input pixel instead of nine), then a warp owns 32
// consecutive columns of FOUR rows so that every weight read from shared memory feeds four pixels (288 FMA per 18
// LDS.128) and every store is 512 contiguous bytes per channel group.
constexpr int kFcRows = 4;
template <typename T>
__global__ void __launch_bounds__(256, 3)
first_conv_kernel(FirstConvSrc src, const int *__restrict__ fw_tile, const int *__restrict__ fw_op,
                  int S, float mean_f, float sd_f, const float *__restrict__ w /*[9][C]*/,
                  const float *__restrict__ bias /*[C]*/, View<T> out) {
  extern __shared__ float sm[];
  float *ws = sm;                 // 9*C
  float *bs = sm + 9 * out.C;     // C
  float *win = bs + out.C;        // 34 x 34 normalised input window ('same' zero padding applies to the normalised image)
  const int tid = threadIdx.x + threadIdx.y * 32;
  for (int i = tid; i < 9 * out.C; i += 256) ws[i] = w[i];
  for (int i = tid; i < out.C; i += 256) bs[i] = bias[i];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int f = blockIdx.z;
  const int tile = fw_tile[f], op = fw_op[f];
  for (int i = tid; i < 34 * 34; i += 256) {
    const int wy = i / 34, wx = i - wy * 34;
    const int yy = y0 + wy - 1, xx = x0 + wx - 1;
    float val = 0.f;
    if (yy >= 0 && yy < S && xx >= 0 && xx < S) {
      int si, sj;
      d4_src(op, yy, xx, S, si, sj);
      const float raw = first_conv_fetch(src, tile, S, si, sj);
      val = __fdiv_rn(__fsub_rn(raw, mean_f), sd_f);
    }
    win[i] = val;
  }
  __syncthreads();
  const int x = x0 + threadIdx.x;
  const int yb = y0 + threadIdx.y * kFcRows;
  if (x >= S || yb >= S) return;
  float2 v[kFcRows + 2][3];                      // every window value in both halves of a register pair (FFMA2 operand)
#pragma unroll
  for (int r = 0; r < kFcRows + 2; ++r)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float q = win[(threadIdx.y * kFcRows + r) * 34 + threadIdx.x + kx];
      v[r][kx] = make_float2(q, q);
    }
  for (int g = 0; g < out.C / 8; ++g) {
    float2 a[kFcRows][4];
#pragma unroll
    for (int r = 0; r < kFcRows; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) a[r][c] = make_float2(bs[g * 8 + 2 * c], bs[g * 8 + 2 * c + 1]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4 *>(ws + t * out.C + g * 8);
      const float4 w1 = *reinterpret_cast<const float4 *>(ws + t * out.C + g * 8 + 4);
      const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int r = 0; r < kFcRows; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) ffma2(a[r][c], v[r + t / 3][t % 3], wv[c]);
    }
#pragma unroll
    for (int r = 0; r < kFcRows; ++r) {
      if (yb + r >= S) break;
      float o8[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) { o8[2 * c] = fmaxf(a[r][c].x, 0.f); o8[2 * c + 1] = fmaxf(a[r][c].y, 0.f); }
      store8v<T>(out, f, yb + r, g, x, o8);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Generic Conv2D(3x3, same, dilation d) + bias + ReLU on CUDA cores (fp32 accumulate).
// UP: the logical input is UpSampling2D((2,2)) nearest of `in` (out[y][x] = in[y//2][x//2],
// train_adipose_unet_v3.py:691,698,705), folded into the gather.
// One thread = one output pixel x 16 output channels.  block (32, 8), grid = (W/32, H/8, nb*Cout/16).
// w: [9][Cin][Cout] fp32 (padded, zero in pad rows/cols).
template <typename T, bool UP>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(View<T> in, View<T> out, const float *__restrict__ w, const float *__restrict__ bias,
                    int dil, int relu) {
  // Two output rows per thread (y and y + 8: the block covers 32 x 16 pixels): every weight vector read from shared memory
  // feeds two pixels, which moves the inner loop from LDS-bound (1 LDS.128 per 4 FMA) to FMA-bound (1 per 8).  Per pixel the
  // accumulation order (chunks, taps, channels) is unchanged; an out-of-range tap of one of the two rows contributes exact zeros.
  __shared__ float ws[9][16][16];
  const int H = out.H, W = out.W;
  const int cgroups = out.C >> 4;
  const int n = blockIdx.z / cgroups;
  const int co0 = (blockIdx.z % cgroups) << 4;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y0 = blockIdx.y * 16 + threadIdx.y;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const bool live[2] = {(x < W) && (y0 < H), (x < W) && (y0 + 8 < H)};
  float2 acc2[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[r][i] = make_float2(0.f, 0.f);
  for (int c0 = 0; c0 < in.C; c0 += 16) {
    __syncthreads();
    for (int i = tid; i < 9 * 256; i += 256) {
      int t = i >> 8, c = (i >> 4) & 15, co = i & 15;
      ws[t][c][co] = w[((size_t)t * in.C + c0 + c) * out.C + co0 + co];
    }
    __syncthreads();
    if (!live[0]) continue;
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {
      const int ix = x + (t % 3 - 1) * dil;
      if (ix < 0 || ix >= W) continue;
      const int sx = UP ? (ix >> 1) : ix;
      float a[2][16];
      bool ok[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int iy = y0 + 8 * r + (t / 3 - 1) * dil;
        ok[r] = live[r] && iy >= 0 && iy < H;
        if (ok[r]) {
          const int sy = UP ? (iy >> 1) : iy;
          load8<T>(in.p + in.at(n, sy, c0 >> 3, sx), a[r]);
          load8<T>(in.p + in.at(n, sy, (c0 >> 3) + 1, sx), a[r] + 8);
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) a[r][c] = 0.f;
        }
      }
      if (!ok[0] && !ok[1]) continue;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float4 *wr = reinterpret_cast<const float4 *>(&ws[t][c][0]);
        const float2 a2[2] = {make_float2(a[0][c], a[0][c]), make_float2(a[1][c], a[1][c])};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = wr[q];
#pragma unroll
          for (int r = 0; r < 2; ++r) {          // packed dual FMA (ffma2): the same two IEEE FMAs per accumulator pair
            ffma2(acc2[r][2 * q], a2[r], make_float2(wv.x, wv.y));
            ffma2(acc2[r][2 * q + 1], a2[r], make_float2(wv.z, wv.w));
          }
        }
      }
    }
  }
  float acc[2][16];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[r][2 * i] = acc2[r][i].x; acc[r][2 * i + 1] = acc2[r][i].y; }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (!live[r]) continue;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float v = acc[r][i] + bias[co0 + i];
      acc[r][i] = relu ? fmaxf(v, 0.f) : v;
    }
    store8<T>(out.p + out.at(n, y0 + 8 * r, co0 >> 3, x), acc[r]);
    store8<T>(out.p + out.at(n, y0 + 8 * r, (co0 >> 3) + 1, x), acc[r] + 8);
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPooling2D((2,2), strides=(2,2)) — train_adipose_unet_v3.py:670,674,678.
// One thread = one output pixel of one channel group (16-byte vectors for bf16, 2 x 16 for fp32).
template <typename T>
__global__ void __launch_bounds__(256) maxpool2_kernel(View<T> in, View<T> out, int nb) {
  const int G = out.C / 8;
  const size_t total = (size_t)nb * out.H * G * out.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, out.W, G, out.H);
    const int x = ri.x, g = ri.g, y = ri.y, n = ri.n;
    float a[8], b[8], c[8], d[8], m[8];
    load8v<T>(in, n, 2 * y, g, 2 * x, a);
    load8v<T>(in, n, 2 * y, g, 2 * x + 1, b);
    load8v<T>(in, n, 2 * y + 1, g, 2 * x, c);
    load8v<T>(in, n, 2 * y + 1, g, 2 * x + 1, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(c[k], d[k]));
    store8v<T>(out, n, y, g, x, m);
  }
}

// Add([dilate1..6]) — train_adipose_unet_v3.py:688.  Dense buffers of identical shape.
template <typename T>
__global__ void __launch_bounds__(256)
add6_kernel(const T *a0, const T *a1, const T *a2, const T *a3, const T *a4, const T *a5, T *out, size_t nvec) {
  constexpr int V = 16 / sizeof(T);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    __align__(16) T r[6][V], o[V];
    *reinterpret_cast<uint4 *>(r[0]) = reinterpret_cast<const uint4 *>(a0)[i];
    *reinterpret_cast<uint4 *>(r[1]) = reinterpret_cast<const uint4 *>(a1)[i];
    *reinterpret_cast<uint4 *>(r[2]) = reinterpret_cast<const uint4 *>(a2)[i];
    *reinterpret_cast<uint4 *>(r[3]) = reinterpret_cast<const uint4 *>(a3)[i];
    *reinterpret_cast<uint4 *>(r[4]) = reinterpret_cast<const uint4 *>(a4)[i];
    *reinterpret_cast<uint4 *>(r[5]) = reinterpret_cast<const uint4 *>(a5)[i];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float s = to_f(r[0][k]);
#pragma unroll
      for (int j = 1; j < 6; ++j) s += to_f(r[j][k]);
      o[k] = from_f<T>(s);
    }
    reinterpret_cast<uint4 *>(out)[i] = *reinterpret_cast<uint4 *>(o);
  }
}

// Add of six hi/lo tensors (bf16x3 precision): recombine, sum in fp32, split again.  One thread per pixel and channel group.
template <typename T>
__global__ void __launch_bounds__(256)
add6_split_kernel(View<T> a0, View<T> a1, View<T> a2, View<T> a3, View<T> a4, View<T> a5, View<T> out, int nb) {
  const int G = out.C / 8;
  const size_t total = (size_t)nb * out.H * G * out.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, out.W, G, out.H);
    const int x = ri.x, g = ri.g, y = ri.y, n = ri.n;
    float s[8], t[8];
    load8v<T>(a0, n, y, g, x, s);
    load8v<T>(a1, n, y, g, x, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += t[k];
    load8v<T>(a2, n, y, g, x, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += t[k];
    load8v<T>(a3, n, y, g, x, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += t[k];
    load8v<T>(a4, n, y, g, x, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += t[k];
    load8v<T>(a5, n, y, g, x, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += t[k];
    store8v<T>(out, n, y, g, x, s);
  }
}

// Conv2D(2, 1x1, softmax) -> channel 1 -> squeeze  (train_adipose_unet_v3.py:748-750).
// softmax(z)[1] == sigmoid(z1 - z0).  wh: [2][C] (padded with zeros), bh: [2].  One thread per pixel.
template <typename T>
__global__ void __launch_bounds__(256)
head_kernel(View<T> in, int nb, const float *__restrict__ wh, const float *__restrict__ bh, float *__restrict__ prob) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 2 * in.C; i += blockDim.x) sm[i] = wh[i];
  __syncthreads();
  const size_t total = (size_t)nb * in.H * in.W;
  const float b0 = bh[0], b1 = bh[1];
  for (size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x; px < total; px += (size_t)gridDim.x * blockDim.x) {
    const int x = px % in.W;
    const size_t r = px / in.W;
    const int y = r % in.H;
    const int n = r / in.H;
    float z0 = 0.f, z1 = 0.f;
    for (int g = 0; g < in.C / 8; ++g) {
      float a[8];
      load8<T>(in.p + in.at(n, y, g, x), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        z0 = fmaf(a[k], sm[g * 8 + k], z0);
        z1 = fmaf(a[k], sm[in.C + g * 8 + k], z1);
      }
    }
    z0 += b0; z1 += b1;
    prob[px] = 1.f / (1.f + expf(z0 - z1));
  }
}

}  // namespace adp
