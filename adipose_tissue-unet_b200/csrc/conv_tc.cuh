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
// Operand staging (both operands K-major, SWIZZLE_NONE "interleaved" core matrices).  One pipeline
// stage = one chunk of 16 input channels (= one MMA K step):
//   * activations live in HBM row-planar ([image][row][cin/8][column][8 x bf16], kernels_simt.cuh), so
//     a 5-D TMA box whose inner extent is 8 pixels x 8 channels = one 128-byte line brings the rows
//     an item needs (with an 8-pixel-aligned halo) into shared memory as channel-group planes
//     [row][cin/8][pixel][8 x bf16].  Each filter tap is then just a
//     different descriptor start address into the same planes (+16 bytes per pixel of horizontal
//     shift, another row slot per vertical shift), so an input pixel is fetched from L2 once and
//     reused by all taps.  'same' zero padding, image borders and channel tails come from TMA
//     out-of-bounds zero fill.
//   * weights: pre-packed on the host in exactly the shared-memory image of the B operand
//     ([tap][cin/8][N][8 x bf16] per (variant, chunk)), streamed with one 1-D bulk copy per stage.
// A stage therefore feeds NTAPS x T MMAs between two mbarrier round trips.
//
// "Variants" make one launch cover (a) the two 176-wide halves of a 352-channel output and
// (b) the four output parities of UpSampling2D(2x2) + conv3x3 (train_adipose_unet_v3.py:691-692):
// on the low-resolution grid each output parity (py,px) is a 2x2-tap conv whose weights are sums
// of the 3x3 taps that land on the same source pixel — 4/9 of the MMA work and no upsampled tensor.
//
// Warp roles (192 threads): w0 producer (TMA + bulk copy), w1 MMA issuer (one elected lane) and
// TMEM allocator, w2-5 epilogue (tcgen05.ld -> bias -> ReLU -> bf16 -> global).
#pragma once
#include "ptx.cuh"

namespace adp {

struct ConvTcVariant {
  int y0;            // first row of the activation box relative to the item's first output row
  int xs_add;        // extra pixel shift of every tap of this variant
  int oy, ox;        // output pixel offset (output = work pixel * oscale + offset)
  int wbase;         // first weight chunk-block of this variant (units of b_bytes)
  int bias_off;      // first bias element
  int out_cg;        // channel-group offset added to the output view
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
  int PW, margin8, nchunks;      // plane width (pixels), left halo in 8-pixel groups, chunks of 16 input channels
  int N;                         // GEMM-N (padded couts of one variant), multiple of 16, <= 256
  int S;                         // pipeline depth
  uint32_t a_box_stride, a_bytes, a_tx_bytes, b_bytes, stage_stride;
  const __nv_bfloat16 *wpk;      // packed weight blocks
  const float *bias;
  __nv_bfloat16 *out;
  int out_cgs, out_cg0, Hout, Wout, oscale;   // output view: channel groups per row of the buffer, first group
  int relu;
  uint32_t idesc;
  int dbg;                       // reserved for kernel experiments
};

constexpr int kTcThreads = 192;

template <int NTAPS, int T>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.S * p.stage_stride);
  uint64_t *full = bars, *empty = full + p.S;
  uint64_t *acc_full = empty + p.S, *acc_empty = acc_full + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.S; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 4); }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
  }
  if (warp == 1) ptx::tmem_alloc_512(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int nitems = p.nb * p.nty * p.ntx * p.nvar;

  if (warp == 0) {
    // ---------------- producer: activation boxes (TMA) + weight block (bulk copy) per stage ----------------
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      const size_t blk_elems = p.b_bytes / 2;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int v = item % p.nvar; int q = item / p.nvar;
        const int tx = q % p.ntx; q /= p.ntx;
        const int ty = q % p.nty; const int n = q / p.nty;
        const int xg = tx * 16 - p.margin8, ys = ty * T + p.var[v].y0;
        const __nv_bfloat16 *w0 = p.wpk + (size_t)p.var[v].wbase * blk_elems;
        for (int c = 0; c < p.nchunks; ++c) {
          uint8_t *sa = smem + (size_t)st * p.stage_stride;
          ptx::mbar_wait(&empty[st], ph ^ 1, 1);
          ptx::mbar_expect_tx(&full[st], p.a_tx_bytes + p.b_bytes);
          for (int b = 0; b < p.nbox; ++b)
            ptx::tma_load_5d(sa + (size_t)b * p.a_box_stride, &tmap, &full[st], 0, xg, c * 2, ys + p.box_dy[b], n);
          ptx::bulk_load_1d(sa + p.a_bytes, w0 + (size_t)c * blk_elems, p.b_bytes, &full[st]);
          if (++st == p.S) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      const uint32_t plane_a = (uint32_t)p.PW * 16u;          // bytes between the two channel-group planes (A)
      const uint32_t plane_b = (uint32_t)p.N * 16u;           // same for B
      // descriptor = hi (SBO = 128 B, version 1) : lo (start >> 4 | LBO >> 4 << 16); only the start moves
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((plane_a >> 4) << 16), b_lo0 = ((plane_b >> 4) << 16);
      uint32_t a_off[NTAPS], b_off[NTAPS];                    // byte offsets within a stage, >> 4
#pragma unroll
      for (int t = 0; t < NTAPS; ++t) {
        a_off[t] = ((uint32_t)p.tap_box[t] * p.a_box_stride + (uint32_t)(p.tap_row[t] * 2) * plane_a +
                    (uint32_t)p.tap_xs[t] * 16u) >> 4;
        b_off[t] = (p.a_bytes + (uint32_t)t * 2u * plane_b) >> 4;
      }
      const uint32_t row_step = (2u * plane_a) >> 4;          // next output row = next row slot
      const uint32_t smem0 = ptx::smem_u32(smem);
      const uint32_t idesc = p.idesc;
      const uint32_t n_cols = (uint32_t)p.N;
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
        const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
        ptx::mbar_wait(&acc_empty[buf], acc_ph ^ 1, 3);
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)buf * (uint32_t)T * n_cols;
        const uint32_t v_off = (uint32_t)p.var[item % p.nvar].xs_add;   // pixels == 16-byte units
        for (int c = 0; c < p.nchunks; ++c) {
          ptx::mbar_wait(&full[st], ph, 4);
          ptx::tc_fence_after();
          const uint32_t s_lo = ((smem0 + (uint32_t)st * p.stage_stride) & 0x3FFFFu) >> 4;
          const uint32_t sa_lo = s_lo + v_off;
#pragma unroll
          for (int t = 0; t < NTAPS; ++t) {
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo0 | (s_lo + b_off[t]));
#pragma unroll
            for (int r = 0; r < T; ++r) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo0 | (sa_lo + a_off[t] + (uint32_t)r * row_step));
              ptx::mma_f16_ss(d0 + (uint32_t)r * n_cols, ad, bd, idesc, (uint32_t)((c | t) != 0));
            }
          }
          ptx::mma_commit(&empty[st]);
          if (++st == p.S) { st = 0; ph ^= 1; }
        }
        ptx::mma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5; TMEM lane quarter = warp % 4) ----------------
    const int ew = warp & 3;
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
      const uint32_t t0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * T * p.N);
#pragma unroll 1
      for (int r = 0; r < T; ++r) {
        const int y = ty * T + r;
        const bool live = (y < p.Hin) && (x < p.Win);
        // row-planar output: ((((n*H + y)*cgs + cg)*W + x)*8; consecutive lanes = consecutive pixels
        const size_t orow = ((size_t)n * p.Hout + (size_t)(y * p.oscale + p.var[v].oy)) * p.out_cgs + p.out_cg0 + p.var[v].out_cg;
        const size_t plane = (size_t)p.Wout * 8;
        __nv_bfloat16 *o = p.out + orow * plane + (size_t)(x * p.oscale + p.var[v].ox) * 8;
#pragma unroll 1
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
            *reinterpret_cast<uint4 *>(o + (size_t)(ch >> 3) * plane) = *reinterpret_cast<uint4 *>(h);
            *reinterpret_cast<uint4 *>(o + (size_t)((ch >> 3) + 1) * plane) = *reinterpret_cast<uint4 *>(h + 8);
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
  if (warp == 1) ptx::tmem_dealloc_512(tmem_base);
}

}  // namespace adp
