// wgrad_tc.cuh — weight gradient of Conv2D(3x3, same, dilation d) as a tcgen05 GEMM whose reduction
// dimension is the pixel axis (the K11 "wgrad" row of SURVEY.md section 2a; Keras computes it by
// reverse-mode AD of train_adipose_unet_v3.py:668-709).
//
//   dW[ky][kx][ci][co] = sum_{n,y,x} X[n, y+(ky-1)d, x+(kx-1)d, ci] * dZ[n, y, x, co]
//
// Per X row y' the three vertical taps pair it with dZ rows y'+d, y', y'-d.  Both tensors are
// row-planar ([row][channel group][pixel][8 channels]), so in shared memory
//   * 8 channels x 8 pixels of X form one MN-major / SWIZZLE_NONE core matrix (8 pixel rows of 16 bytes):
//     A = X^T tile [M = up to 128 input channels] x [K = 16 pixels], SBO = plane stride, LBO = 128 B;
//     a horizontal tap is a start-address shift of 16 B per pixel;
//   * the three dZ rows y'-d, y', y'+d of one output-channel chunk lie contiguously as [row][group] planes,
//     i.e. ONE B operand with N = 3 x chunk columns: a single MMA produces the three vertical taps.
// So a K step of 16 pixels is 3 MMAs (one per horizontal tap) of shape 128 x (3*chunk) x 16 into three
// TMEM accumulators (3 x 144 = 432 of the 512 columns for chunk = 48).
//
// Bias gradient for free: an X block carries at most 15 channel groups (120 channels); the 16th plane of the A
// operand is filled once with bf16 ones, so accumulator row 120 of the centre tap is sum_pixels dZ = db.
//
// A CTA owns one (input-channel block, output-channel chunk) pair and a contiguous range of
// (image, row, 128-pixel strip) units; it accumulates its whole range in TMEM and then writes the
// 128 x 432 partial sums into ITS OWN slot of a scratch buffer [ctas_per_combo][9][cin_pad][cout_pad]; wgrad_reduce_kernel adds
// the slots in a fixed order.  (The first version added the partials into dW with red.global.add.f32: the order of those
// float additions changed from run to run, so the gradient was not reproducible.)
//
// Warp roles (192 threads): w0 TMA producer, w1 MMA issuer + TMEM allocator, w2-5 epilogue.
#pragma once
#include "ptx.cuh"

namespace adp {

struct WgradTcParams {
  int nb, H, W;                  // images, spatial size (X and dZ have the same size)
  int dil;
  int cin_pad, cout_pad;
  int n_ci_blk, n_co_chunk;      // 120-channel blocks of X, chunks of dZ channels
  int co_chunk;                  // channels per chunk (multiple of 16, 3*co_chunk <= 170)
  int ctas_per_combo;
  int cga_box;                   // channel groups per X box = min(15, cin_pad / 8); plane 15 holds ones
  int PW, margin8;               // X plane width in pixels (128 + 2*margin), halo in 8-pixel groups
  int S;                         // pipeline depth
  uint32_t a_bytes, b_row_bytes, stage_stride;
  float *dW;                     // scratch [ctas_per_combo][9][cin_pad][cout_pad] fp32: slot `part` of every combo is written, not accumulated
  float *db;                     // scratch [ctas_per_combo][cout_pad] fp32
};

constexpr int kWgThreads = 192;


__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmz, const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.S * p.stage_stride);
  uint64_t *full = bars, *empty = full + p.S, *acc_full = empty + p.S;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.S; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmx);
    ptx::prefetch_tmap(&tmz);
  }
  if (warp == 1) ptx::tmem_alloc_512(tmem_ptr);
  // plane 15 of every stage's A region := bf16 1.0 (never touched by the TMA boxes, which carry <= 15 groups)
  for (int st = 0; st < p.S; ++st) {
    uint32_t *ones = reinterpret_cast<uint32_t *>(smem + (size_t)st * p.stage_stride + (size_t)15 * p.PW * 16);
    for (int i = threadIdx.x; i < p.PW * 4; i += kWgThreads) ones[i] = 0x3F803F80u;
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // this CTA's combo and unit range
  const int combo = blockIdx.x / p.ctas_per_combo, part = blockIdx.x % p.ctas_per_combo;
  const int cib = combo / p.n_co_chunk, coc = combo % p.n_co_chunk;
  const int ntx = (p.W + 127) / 128;
  const long long nunits = (long long)p.nb * p.H * ntx;
  const long long u0 = nunits * part / p.ctas_per_combo, u1 = nunits * (part + 1) / p.ctas_per_combo;
  const int ci0 = cib * 120, co0 = coc * p.co_chunk;
  // TMA boxes have a fixed shape: channel groups past the end of the tensor arrive as zeros and only cost MMA columns
  const int ncga = p.cga_box;                                    // channel groups per X box (<= 16)
  const int nco = p.co_chunk;
  const int ncgb = nco / 8;
  const uint32_t plane_a = (uint32_t)p.PW * 16u, plane_b = 128u * 16u;
  const uint32_t N = 3u * (uint32_t)nco;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      const uint32_t a_tx = (uint32_t)ncga * plane_a, b_tx = 3u * (uint32_t)ncgb * plane_b;
      for (long long u = u0; u < u1; ++u) {
        const int tx = (int)(u % ntx); long long r = u / ntx;
        const int y = (int)(r % p.H), n = (int)(r / p.H);
        uint8_t *sa = smem + (size_t)st * p.stage_stride;
        ptx::mbar_wait(&empty[st], ph ^ 1, 11);
        ptx::mbar_expect_tx(&full[st], a_tx + b_tx);
        ptx::tma_load_5d(sa, &tmx, &full[st], 0, tx * 16 - p.margin8, ci0 / 8, y, n);
        // dZ rows y-d, y, y+d: three single-row boxes stored back to back (= rows of one stacked B operand)
        for (int s = 0; s < 3; ++s)
          ptx::tma_load_5d(sa + p.a_bytes + (size_t)s * ncgb * plane_b, &tmz, &full[st], 0, tx * 16, co0 / 8, y + (s - 1) * p.dil, n);
        if (++st == p.S) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      // MN-major, SWIZZLE_NONE: SBO = byte stride between 8-channel groups (planes), LBO = between 8-pixel blocks
      const uint32_t a_hi = (plane_a >> 4) | (1u << 14), b_hi = (plane_b >> 4) | (1u << 14);
      const uint32_t lo_lbo = ((128u >> 4) << 16);
      const uint32_t idesc = ptx::make_idesc(128, (int)N, 1) | (1u << 15) | (1u << 16);
      const uint32_t smem0 = ptx::smem_u32(smem);
      uint32_t a_shift[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) a_shift[j] = (uint32_t)(p.margin8 * 8 + (j - 1) * p.dil);   // pixels == 16-byte units
      bool first = true;
      for (long long u = u0; u < u1; ++u) {
        ptx::mbar_wait(&full[st], ph, 12);
        ptx::tc_fence_after();
        const uint32_t s_lo = ((smem0 + (uint32_t)st * p.stage_stride) & 0x3FFFFu) >> 4;
        const uint32_t b_lo = s_lo + (p.a_bytes >> 4);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bd = ((uint64_t)b_hi << 32) | (uint64_t)(lo_lbo | (b_lo + (uint32_t)ks * 16u));
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const uint64_t ad = ((uint64_t)a_hi << 32) | (uint64_t)(lo_lbo | (s_lo + a_shift[j] + (uint32_t)ks * 16u));
            ptx::mma_f16_ss(tmem_base + (uint32_t)j * N, ad, bd, idesc, (uint32_t)!(first && ks == 0));
          }
        }
        first = false;
        ptx::mma_commit(&empty[st]);
        if (++st == p.S) { st = 0; ph ^= 1; }
      }
      ptx::mma_commit(acc_full);
    }
  } else {
    // epilogue: TMEM lane = input channel, column = (kx, dZ row slot s, output channel); ky = 2 - s
    const int q4 = warp & 3;
    const bool have = u1 > u0;               // a CTA without units still writes its slot (zeros): the reduction reads every slot
    if (have) {
      ptx::mbar_wait(acc_full, 0, 13);
      ptx::tc_fence_after();
    }
    const int row = q4 * 32 + lane;
    const int ci = ci0 + row;
    const bool live = row < ncga * 8 && ci < p.cin_pad;
    const bool bias_row = row == 120 && cib == 0;
    const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int units = (int)(3u * N) >> 4;
    float *slotW = p.dW + (size_t)part * 9 * p.cin_pad * p.cout_pad;
    float *slotb = p.db + (size_t)part * p.cout_pad;
    for (int un = 0; un < units; ++un) {
      float v[16];
      if (have) ptx::tmem_ld16(t0 + (uint32_t)un * 16u, v);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      const int col = un * 16;
      const int j = col / (int)N, rem = col - j * (int)N;
      const int s = rem / nco, c = rem - s * nco;
      const int tap = (2 - s) * 3 + j;
      if (live && co0 + c < p.cout_pad) {
        float *dst = slotW + ((size_t)tap * p.cin_pad + ci) * p.cout_pad + co0 + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (bias_row && j == 1 && s == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (co0 + c + i < p.cout_pad) slotb[co0 + c + i] = v[i];
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc_512(tmem_base);
}

// Fixed-order sum over the slots of the per-CTA partial sums, written straight into the FLAT gradient (Keras order: kernel
// HWIO with the real channel counts, then bias): flat_k[(t*cin + ci)*cout + co] = sum_s slot_s[(t*cin_pad + cip)*cout_pad + co],
// cip = ci, or skip_pad + (ci - skip) for the upsampled-path half of a concat-fed layer; flat_b[co] = sum_s slot_b[co].
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ sW, const float *__restrict__ sb, int slots, int cin, int cout, int cin_pad, int cout_pad,
                    int skip, int skip_pad, float *__restrict__ flat_k, float *__restrict__ flat_b) {
  const size_t nk = (size_t)9 * cin * cout, np = (size_t)9 * cin_pad * cout_pad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nk + (size_t)cout; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nk) {
      const int co = (int)(i % cout); size_t r = i / cout;
      const int ci = (int)(r % cin); const int t = (int)(r / cin);
      const int cip = (skip && ci >= skip) ? skip_pad + (ci - skip) : ci;
      const size_t pi = ((size_t)t * cin_pad + cip) * cout_pad + co;
      float a = sW[pi];
      for (int s = 1; s < slots; ++s) a += sW[(size_t)s * np + pi];
      flat_k[i] = a;
    } else {
      const size_t c = i - nk;
      float a = sb[c];
      for (int s = 1; s < slots; ++s) a += sb[(size_t)s * cout_pad + c];
      flat_b[c] = a;
    }
  }
}

// db[co] = sum over pixels of dZ (bias gradient).  One warp owns 32 consecutive pixels of a row and walks
// the channel groups; per-block partial sums leave through shared memory and one atomic per channel.
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(View<T> dz, int nb, float *__restrict__ db) {
  extern __shared__ float sm[];     // [C]
  const int C = dz.C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int G = C / 8;
  const int xchunks = (dz.W + 31) / 32;
  const long long nwork = (long long)nb * dz.H * xchunks * G;
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  // a warp keeps one channel group while it strides over pixel chunks (work index = chunk * G + group, stride nw:
  // choose nw as a multiple of G so the group of a warp is fixed)
  const int g = (int)(wid % G);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long w = wid; w < nwork; w += nw) {
    long long r = w / G;
    const int xc = (int)(r % xchunks); r /= xchunks;
    const int y = (int)(r % dz.H), n = (int)(r / dz.H);
    const int x = xc * 32 + lane;
    if (x < dz.W) {
      float a[8];
      load8<T>(dz.p + dz.at(n, y, g, x), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += a[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) atomicAdd(&sm[g * 8 + k], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&db[i], sm[i]);
}

// nearest-neighbour 2x upsampling, materialised (input of the upsampled convs' weight gradient).  One thread per LOW-resolution
// pixel and channel group: one 16-byte load feeds four 16-byte stores (two adjacent columns of two rows), so the index
// arithmetic is paid once per 64 bytes written (one thread per output pixel was issue-bound: ncu 65 % issue slots, 45 % DRAM).
template <typename T>
__global__ void __launch_bounds__(256) upsample2_kernel(View<T> lo, View<T> hi, int nb) {
  const int G = hi.C / 8;
  const size_t total = (size_t)nb * lo.H * G * lo.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const RpIndex ri = rp_index(i, lo.W, G, lo.H);
    if constexpr (sizeof(T) == 2) {
      const uint4 v = *reinterpret_cast<const uint4 *>(lo.p + lo.at(ri.n, ri.y, ri.g, ri.x));
      const size_t rs = (size_t)hi.cgs * hi.W * 8;           // one output row further
      if ((lo.W & 31) == 0) {
        // a warp holds 32 consecutive source pixels of one row = 64 output pixels: lane l stores output pixels l and 32 + l
        // (values of source lanes l/2 and 16 + l/2), so every store instruction covers 512 contiguous bytes.  (Each thread
        // storing its own two adjacent 16-byte vectors made every instruction touch 32 half-written sectors: ncu lg_throttle 42.)
        const int lane = threadIdx.x & 31;
        uint4 a, b;
        a.x = __shfl_sync(0xffffffffu, v.x, lane >> 1); a.y = __shfl_sync(0xffffffffu, v.y, lane >> 1);
        a.z = __shfl_sync(0xffffffffu, v.z, lane >> 1); a.w = __shfl_sync(0xffffffffu, v.w, lane >> 1);
        b.x = __shfl_sync(0xffffffffu, v.x, 16 + (lane >> 1)); b.y = __shfl_sync(0xffffffffu, v.y, 16 + (lane >> 1));
        b.z = __shfl_sync(0xffffffffu, v.z, 16 + (lane >> 1)); b.w = __shfl_sync(0xffffffffu, v.w, 16 + (lane >> 1));
        T *o = hi.p + hi.at(ri.n, 2 * ri.y, ri.g, 2 * (ri.x - lane) + lane);
        *reinterpret_cast<uint4 *>(o) = a; *reinterpret_cast<uint4 *>(o + 32 * 8) = b;
        *reinterpret_cast<uint4 *>(o + rs) = a; *reinterpret_cast<uint4 *>(o + rs + 32 * 8) = b;
      } else {
        T *o = hi.p + hi.at(ri.n, 2 * ri.y, ri.g, 2 * ri.x);
        reinterpret_cast<uint4 *>(o)[0] = v; reinterpret_cast<uint4 *>(o)[1] = v;
        reinterpret_cast<uint4 *>(o + rs)[0] = v; reinterpret_cast<uint4 *>(o + rs)[1] = v;
      }
    } else {
      float a[8];
      load8<T>(lo.p + lo.at(ri.n, ri.y, ri.g, ri.x), a);
#pragma unroll
      for (int q = 0; q < 4; ++q) store8<T>(hi.p + hi.at(ri.n, 2 * ri.y + (q >> 1), ri.g, 2 * ri.x + (q & 1)), a);
    }
  }
}

}  // namespace adp
