// conv_tc.cuh — Conv2D(3x3, same, dilation d) + bias + ReLU as a tcgen05 implicit GEMM
// (train_adipose_unet_v3.py:668-709 — 20 of the 22 conv layers run through this kernel).
//
//   D[128 pixels x N couts] (TMEM, fp32)  +=  A[128 pixels x 16 cin] (smem) * B[N x 16 cin] (smem)
//
// GEMM-M = 128 consecutive pixels of one image row, GEMM-N = all (padded) output channels of the
// launch, GEMM-K = taps x input channels.  A persistent CTA walks "work items" = (image, block
// of T output rows, 128-pixel strip, variant) and keeps T accumulators (one per output row) in
// TMEM so that every weight block fetched from L2 feeds T MMAs; TMEM is split in two halves so
// the epilogue of item i overlaps the MMAs of item i+1.
//
// Operand staging (both operands K-major, SWIZZLE_NONE "interleaved" core matrices):
//   * activations: ONE 5-D TMA box per chunk of KC input channels brings the rows an item needs
//     (with halo) into shared memory as channel-group planes  [row][cin/8][pixel][8 x bf16].
//     Each filter tap is then just a different descriptor start address into the same planes
//     (+16 bytes per pixel of horizontal shift, another row slot per vertical shift), so an input
//     pixel is fetched from L2 once and reused by all taps.  'same' zero padding, image borders and
//     channel tails come from TMA out-of-bounds zero fill.
//   * weights: pre-packed on the host in exactly the shared-memory image of the B operand
//     ([cin/8][N][8 x bf16] per (variant, chunk, tap)), streamed with 1-D bulk copies.
//
// "Variants" make one launch cover (a) the two 176-wide halves of a 352-channel output and
// (b) the four output parities of UpSampling2D(2x2) + conv3x3 (train_adipose_unet_v3.py:691-692):
// on the low-resolution grid each output parity (py,px) is a 2x2-tap conv whose weights are sums
// of the 3x3 taps that land on the same source pixel — 4/9 of the MMA work and no upsampled tensor.
//
// Warp roles (256 threads): w0 activation producer (TMA), w1 weight producer (bulk copy),
// w2 MMA issuer (one elected lane), w3 TMEM allocator, w4-7 epilogue (tcgen05.ld -> bias -> ReLU
// -> bf16 -> global).
#pragma once
#include "ptx.cuh"

namespace adp {

struct ConvTcVariant {
  int y0, x0;        // origin of the activation box relative to the item's (row, col)
  int oy, ox;        // output pixel offset (output = work pixel * oscale + offset)
  int wbase;         // first weight block of this variant (block units)
  int bias_off;      // first bias element
  int out_coff;      // channel offset added to the output view
  int pad_;
};

struct ConvTcParams {
  int nb, Hin, Win;              // work grid = input-resolution grid
  int T, ntx, nty;               // output rows per item, strips per row, row blocks per image
  int nvar;
  ConvTcVariant var[4];
  int ntaps;
  int tap_box[9], tap_row[9], tap_xs[9];   // activation box index, first row slot, pixel shift of each tap
  int nbox, box_dy[3], BR;       // boxes per chunk, their row offsets, rows per box
  int PW, CG, nchunks;           // plane width (pixels), channel groups (of 8) per chunk, chunks
  int N;                         // GEMM-N (padded couts of one variant), multiple of 16, <= 256
  int SA, SB;                    // pipeline depths
  uint32_t a_box_stride, a_stride, b_stride, a_tx_bytes, b_bytes;
  const __nv_bfloat16 *wpk;      // packed weight blocks
  const float *bias;
  __nv_bfloat16 *out;
  int out_pitch, out_coff, Hout, Wout, oscale;
  int relu;
  uint32_t idesc;
  int dbg;                       // reserved for kernel experiments
};

__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;
  uint8_t *sB = sA + (size_t)p.SA * p.a_stride;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)p.SB * p.b_stride);
  uint64_t *a_full = bars, *a_empty = a_full + p.SA;
  uint64_t *b_full = a_empty + p.SA, *b_empty = b_full + p.SB;
  uint64_t *acc_full = b_empty + p.SB, *acc_empty = acc_full + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
  }
  if (warp == 3) ptx::tmem_alloc_512(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int per_img = p.nty * p.ntx * p.nvar;
  const int nitems = p.nb * per_img;

  if (warp == 0) {
    // ---------------- activation producer ----------------
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int v = item % p.nvar; int q = item / p.nvar;
        const int tx = q % p.ntx; q /= p.ntx;
        const int ty = q % p.nty; const int n = q / p.nty;
        const int xs = tx * 128 + p.var[v].x0, ys = ty * p.T + p.var[v].y0;
        for (int c = 0; c < p.nchunks; ++c) {
          ptx::mbar_wait(&a_empty[st], ph ^ 1, 1);
          ptx::mbar_expect_tx(&a_full[st], p.a_tx_bytes);
          for (int b = 0; b < p.nbox; ++b)
            ptx::tma_load_5d(sA + (size_t)st * p.a_stride + (size_t)b * p.a_box_stride, &tmap, &a_full[st], 0, xs,
                             c * p.CG, ys + p.box_dy[b], n);
          if (++st == p.SA) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- weight producer ----------------
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      const size_t blk_elems = p.b_bytes / 2;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int v = item % p.nvar;
        const __nv_bfloat16 *w0 = p.wpk + (size_t)p.var[v].wbase * blk_elems;
        const int nblk = p.nchunks * p.ntaps;
        for (int k = 0; k < nblk; ++k) {
          ptx::mbar_wait(&b_empty[st], ph ^ 1, 2);
          ptx::mbar_expect_tx(&b_full[st], p.b_bytes);
          ptx::bulk_load_1d(sB + (size_t)st * p.b_stride, w0 + (size_t)k * blk_elems, p.b_bytes, &b_full[st]);
          if (++st == p.SB) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
      const uint32_t plane_a = (uint32_t)p.PW * 16u;          // bytes between channel-group planes (A)
      const uint32_t plane_b = (uint32_t)p.N * 16u;           // bytes between channel-group planes (B)
      // K-major / SWIZZLE_NONE: LBO = distance between the two 16-byte K chunks of one MMA (the
      // next channel-group plane), SBO = distance between 8-row core matrices (8 pixels x 16 B).
      const uint32_t a_lbo = plane_a, a_sbo = 128u, b_lbo = plane_b, b_sbo = 128u;
      const int kk_n = p.CG >> 1;                             // MMAs (K=16) per chunk and tap
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&acc_empty[buf], acc_ph ^ 1, 3);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(buf * p.T * p.N);
        for (int c = 0; c < p.nchunks; ++c) {
          ptx::mbar_wait(&a_full[sa], pa, 4);
          ptx::tc_fence_after();
          const uint32_t a_st = sA_u + (uint32_t)sa * p.a_stride;
          for (int t = 0; t < p.ntaps; ++t) {
            ptx::mbar_wait(&b_full[sb], pb, 5);
            ptx::tc_fence_after();
            const uint32_t b_st = sB_u + (uint32_t)sb * p.b_stride;
            const uint32_t a_tap = a_st + (uint32_t)p.tap_box[t] * p.a_box_stride + (uint32_t)p.tap_xs[t] * 16u;
            for (int r = 0; r < p.T; ++r) {
              const uint32_t a_row = a_tap + (uint32_t)((p.tap_row[t] + r) * p.CG) * plane_a;
              for (int kk = 0; kk < kk_n; ++kk) {
                const uint64_t ad = ptx::smem_desc(a_row + (uint32_t)(2 * kk) * plane_a, a_lbo, a_sbo);
                const uint64_t bd = ptx::smem_desc(b_st + (uint32_t)(2 * kk) * plane_b, b_lbo, b_sbo);
                ptx::mma_f16_ss(d0 + (uint32_t)(r * p.N), ad, bd, p.idesc, (uint32_t)((c | t | kk) != 0));
              }
            }
            ptx::mma_commit(&b_empty[sb]);
            if (++sb == p.SB) { sb = 0; pb ^= 1; }
          }
          ptx::mma_commit(&a_empty[sa]);
          if (++sa == p.SA) { sa = 0; pa ^= 1; }
        }
        ptx::mma_commit(&acc_full[buf]);
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int ew = warp - 4;                                  // == warp % 4: TMEM lane quarter
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
      const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
      const int v = item % p.nvar; int q = item / p.nvar;
      const int tx = q % p.ntx; q /= p.ntx;
      const int ty = q % p.nty; const int n = q / p.nty;
      const int x = tx * 128 + ew * 32 + lane;
      const float *bias = p.bias + p.var[v].bias_off;
      ptx::mbar_wait(&acc_full[buf], acc_ph, 6);
      ptx::tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * p.T * p.N);
      for (int r = 0; r < p.T; ++r) {
        const int y = ty * p.T + r;
        const bool live = (y < p.Hin) && (x < p.Win);
        const size_t opix = ((size_t)n * p.Hout + (size_t)(y * p.oscale + p.var[v].oy)) * p.Wout +
                            (size_t)(x * p.oscale + p.var[v].ox);
        __nv_bfloat16 *o = p.out + opix * p.out_pitch + p.out_coff + p.var[v].out_coff;
        for (int ch = 0; ch < p.N; ch += 16) {
          float acc[16];
          ptx::tmem_ld16(t0 + (uint32_t)(r * p.N + ch), acc);
          if (live) {
            __align__(16) __nv_bfloat16 h[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float f = acc[i] + __ldg(bias + ch + i);
              h[i] = __float2bfloat16_rn(p.relu ? fmaxf(f, 0.f) : f);
            }
            *reinterpret_cast<uint4 *>(o + ch) = *reinterpret_cast<uint4 *>(h);
            *reinterpret_cast<uint4 *>(o + ch + 8) = *reinterpret_cast<uint4 *>(h + 8);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 3) ptx::tmem_dealloc_512(tmem_base);
}

}  // namespace adp
