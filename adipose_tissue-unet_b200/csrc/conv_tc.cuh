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
//   * resident weights (ConvTcParams::bres): where the blocks of all chunks (of the one variant a CTA meets) fit next to >= 4
//     activation-only stages, they are copied once per CTA and the stages carry activations only.
//
// "Variants" make one launch cover (a) the two 176-wide halves of a 352-channel output and
// (b) the four output parities of UpSampling2D(2x2) + conv3x3 (train_adipose_unet_v3.py:691-692):
// on the low-resolution grid each output parity (py,px) is a 2x2-tap conv whose weights are sums
// of the 3x3 taps that land on the same source pixel — 4/9 of the MMA work and no upsampled tensor.
//
// ky-stacked issue (KYS, dilation 1 and N <= 128): with N = 48 or 96 one MMA needs 44 / 56 shared-memory wavefronts
// (128 B each: 4 KB of A + N*32 B of B) but only N/2 = 24 / 48 tensor cycles, i.e. the operand read port, not the tensor
// pipe, bounds the layer (ncu: l1tex__data_pipe_tc_wavefronts_mem_shared 65-81 %, tensor pipe 35-70 %).  Input row i of
// an item contributes to the output rows o = i-2, i-1, i (vertical taps +1, 0, -1) and the accumulators of consecutive
// output rows are consecutive TMEM column ranges, so ONE MMA with the three vertical taps stacked along N
// (B = [W(+1) | W(0) | W(-1)], N' = 3N) adds row i into all three at once: T+2 reads of A per horizontal tap instead of
// 3T.  The first (chunk 0, kx 0) pass is issued unstacked because the accumulate flag is per instruction.
//
// bf16x3 precision (p.split): see ConvTcParams - three chunks per 16 input channels, hi/lo stores in the epilogue.
//
// Warp roles (352 threads): w0 producer (TMA + bulk copy), w1 MMA issuer A (one elected lane) and
// TMEM allocator, w2-9 epilogue: two warps per TMEM lane quarter, software-pipelined
// tcgen05.ld -> bias -> ReLU -> bf16 -> coalesced row-planar stores, w10 MMA issuer B (the two issuers
// alternate work items, see the issuer comment below).  Fused epilogues:
//   EPI_HEAD  Conv2D(2,1x1,softmax)[...,1] of train_adipose_unet_v3.py:748-750 on the fp32 accumulators
//             of up1_conv3 (the 44-channel activation never reaches HBM),
//   EPI_POOL  MaxPooling2D(2x2) (train_adipose_unet_v3.py:670,674) written next to the skip tensor,
//   EPI_BWD   data-gradient twin of the training step: (acc + residual) * [mask > 0] * scale, the mask tile staged in
//             shared memory by a TMA box per item for accumulators up to 96 wide (ConvTcParams::mask_bufs),
//   EPI_UPSUM data-gradient twin of an upsampled conv with two-row items: UpSampling2D's backward (2x2 sum, second gradient,
//             ReLU'/dropout mask) in the epilogue, the full-resolution gradient is never written,
//   dropout   (EPI_STORE, training forward) the Dropout that follows the layer, by the counter-based hash of kernels_train.cuh.
// FC variant (template flag, 512 threads, opt-in): five more warps compute the FIRST conv of the network into the pipeline stages
// of down1_conv2 (struct FirstConvFuse).  All fused forms are bit-identical to the separate kernels they replace.
#pragma once
#include "kernels_simt.cuh"
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

// FC variant (template flag): the A operand of down1_conv2 - the output of the FIRST conv, Reshape + Conv2D(1 -> C, 3x3) + ReLU
// (train_adipose_unet_v3.py:665-668) - is computed inside the kernel by five "stencil" warps straight into the pipeline stages,
// in exactly the layout the TMA boxes would have produced.  The 48-channel tensor (100 MB per forward) is then neither written
// nor read.  The stencil reads the normalised, dihedrally transformed image (float32 [forward][S][S], written by
// tta_input_kernel: full_evaluation_enhanced.py:1306, :590) through 8 x 136 TMA boxes (tensor map `tmask`) whose out-of-range
// fill is the conv's zero padding.  Inference only (training needs the tensor for the weight gradient).
struct FirstConvFuse {
  alignas(16) float w[9][64];    // [tap][output channel] fp32, zero padded (read through the constant bank, 128-bit uniform loads)
  alignas(16) float b[64];
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
  int epi_mode;                  // EPI_STORE / EPI_HEAD / EPI_POOL
  const float *head_w;           // [2][N] fp32 (zero padded), EPI_HEAD
  const float *head_b;           // [2]
  float *prob;                   // [nb][Hout][Wout] fp32, EPI_HEAD
  __nv_bfloat16 *pool_out;       // pooled row-planar tensor [nb][Hout/2][pool_cgs][Wout/2][8], EPI_POOL
  int pool_cgs, pool_cg0;
  int dbg;                       // timing experiments (ADP_TC_DEBUG): see the producer comment and tools/tc_experiments.sh
  long long *dbg_out;            // dbg & 16: per-CTA role timers (8 x int64 per CTA), tools/tc_timers.sh
  // ky-stacked issue (template KYS): the vertical taps of one horizontal tap are stacked along GEMM-N in the weight
  // block ([kx][cin/8][KY*N][8], vertical taps in DESCENDING order), tap_xs[kx] is the pixel shift of horizontal tap kx
  int kys;
  uint32_t idesc_stack[3];       // instruction descriptors for N, 2N, 3N
  // bf16x3 ("split") precision: every activation / weight value v is carried as hi = bf16(v), lo = bf16(v - hi) (hi groups
  // first, lo groups `*_lo` channel groups further in the same row-planar row) and a conv is the three bf16 GEMMs
  // A_hi*W_hi + A_hi*W_lo + A_lo*W_hi accumulated in fp32 (max-abs probability error ~2e-5 vs float64, CPU simulation
  // and tests/test_gpu_forward.py).  Implemented as THREE pipeline chunks per 16 input channels: nchunks and the packed
  // weight blocks are tripled by the host ([W_hi, W_lo, W_hi] per chunk); chunk j reads A from group (j/3)*2 (+ in_lo for
  // j%3 == 2).  The epilogue stores hi and lo.
  int split, in_lo, out_lo, pool_lo;
  // backward use (data-gradient twin, EPI_STORE only): out = (acc + resid) * [mask > 0] * mask_scale, where resid and
  // mask are tensors with exactly the layout of `out` (gradient fan-in of Add / ReLU' of the producing layer)
  const __nv_bfloat16 *resid;
  const __nv_bfloat16 *mask;
  float mask_scale;
  // EPI_BWD, N <= 96: the mask tile of an item (T rows x N/8 channel-group planes x 128 pixels x 16 B) is staged in shared
  // memory by the producer with one TMA box per item (tensor map `tmask`, 1 or 2 buffers after the pipeline stages) instead
  // of per-thread global loads: with register prefetch the loads' latency was exposed once per item (ablation: 11.5 k
  // cycles per item with the loads, 6.3 k without).  mask_bufs = 0 keeps the register-prefetch loads (wide accumulators)
  int mask_bufs;                 // 0, 1 or 2
  uint32_t mask_bytes;           // T * N * 256
  // Dropout fused into the EPI_STORE epilogue (training forward, hash masks): exactly dropout_dense_kernel on the stored
  // tensor - the value is rounded to bf16, multiplied by 1/keep or zeroed by the hash of its 8-channel group index
  // (= element offset / 8 of the dense output tensor), and rounded again.  drop_thr = 0: off.
  uint32_t drop_thr, drop_salt;
  float drop_inv;
  // Resident weights (bres = 1; launch_conv_tc turns it on when the packed weight blocks of ALL chunks and variants fit next to
  // >= 4 activation-only stages): the persistent CTA copies them into shared memory once (nvar * nchunks bulk copies on their own
  // mbarrier) and a pipeline stage then carries activations only.  Per work item of a 48-channel layer this takes the weight
  // refill (108 of ~1330 shared-memory wavefront-cycles per chunk, the port that binds these layers - DESIGN 4.1) and 13.8 KB of
  // L2 reads per chunk off the loop.  Same MMAs on the same operand bytes: results are bit-identical.
  // bres = 2: the grid is a multiple of nvar, so a CTA (items blockIdx.x + k * gridDim.x, variant = item % nvar) only ever sees
  // variant blockIdx.x % nvar and keeps that variant's blocks alone (the four parity variants of an upsampled conv).
  int bres;
  uint32_t bres_off, bres_bytes; // byte offset of the resident blocks in dynamic shared memory (128-byte aligned), their total size
  FirstConvFuse fc;              // FC variant only
};

// index of weight (tap t, row n of variant block) inside a (variant, chunk) weight block of `ntaps` taps, channel-group
// plane g, lane j: plain layout [t][g][N][8], ky-stacked layout [kx][g][KY*N][8] with vertical taps descending
__host__ __device__ inline size_t tc_block_index(int kys, int ntaps, int N, int t, int g, int n, int j) {
  if (!kys) return (((size_t)t * 2 + g) * N + n) * 8 + j;
  const int KY = ntaps == 9 ? 3 : 2;
  const int kyi = t / KY, kx = t % KY;
  return ((((size_t)kx * 2 + g) * KY + (KY - 1 - kyi)) * N + n) * 8 + j;
}

size_t tc_smem_bytes(const ConvTcParams &p) {   // FC variant: + kFcWinBytes
  // stages | mask tile buffers (EPI_BWD) | mbarriers (2S + 8) + tmem slot | bias (704 floats) | head weights (2*256 + 2 floats)
  const size_t base = (size_t)p.S * p.stage_stride + (size_t)p.mask_bufs * p.mask_bytes + (2 * p.S + 6 + 4) * 8 + (704 + 520) * 4;
  return p.bres ? (size_t)p.bres_off + p.bres_bytes : base;
}
// where the resident weight blocks start for the current S / mask_bufs (host: call after those are final)
inline uint32_t tc_bres_offset(const ConvTcParams &p) {
  return (uint32_t)(((size_t)p.S * p.stage_stride + (size_t)p.mask_bufs * p.mask_bytes + (2 * p.S + 6 + 4) * 8 + (704 + 520) * 4 + 127) & ~(size_t)127);
}
constexpr int kFcWinW = 136, kFcWinRows = 8;                     // input window of one item: 8 rows x 132 columns, loaded as a
constexpr int kFcWinX0 = 4;                                      // box of 136 that starts 4 columns left of the strip (16-byte aligned)
constexpr uint32_t kFcWinBoxBytes = kFcWinRows * kFcWinW * 4;
constexpr size_t kFcWinBytes = 2 * kFcWinBoxBytes + 128 + 10 * 64 * 4;   // double buffered, 128-byte aligned; + weights and bias
constexpr int kFcMaxChunks = 4;                                  // first conv of at most 64 channels
constexpr int kFcWarps = 5, kFcThreads = kFcWarps * 32;          // four warps own 32 columns each, the fifth the two halo columns

// Timing experiments (ADP_TC_DEBUG switches, role timers) are compiled in only by a debug build
// (`python -m adipose_unet_b200.build --force --debug`, -DADP_TC_DEBUG_BUILD=1): the production kernels carry no p.dbg
// branches in the MMA-issue and epilogue loops.
#ifndef ADP_TC_DEBUG_BUILD
#define ADP_TC_DEBUG_BUILD 0
#endif
ADP_DEVINL bool tc_dbg(const ConvTcParams &p, int bits) {
  if constexpr (ADP_TC_DEBUG_BUILD != 0) return (p.dbg & bits) != 0;
  else return false;
}

constexpr int kTcThreads = 352;   // producer, MMA issuer A, 8 epilogue warps, MMA issuer B
constexpr int kTcThreadsFc = kTcThreads + kFcThreads;   // FC variant: + five stencil warps (w11..15)

// f32 pair -> packed bf16x2 with ReLU (cvt.rn.relu: max(x, 0) then round to nearest even, what fmaxf + __float2bfloat16_rn give)
ADP_DEVINL uint32_t relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// ---- FC variant: the first conv as an in-kernel producer of the A operand (see FirstConvFuse) ----
// R consecutive rows x 8 channels (group cg) of the first conv at one column: acc = bias, then the nine taps in order with one
// IEEE FMA each (the accumulation order of first_conv_kernel: bit-identical values), ReLU, round to bf16: one packed 16-byte
// register quad per row, stored into the pipeline stage by the caller once the stage is free.  The weights sit in shared memory (one broadcast LDS.128 per four weights = one
// wavefront: ~450 per item next to the ~4000 of the MMA operand fetch).  Measured alternatives (1024^2, 16 forwards, whole
// kernel): weights as uniform-register FFMA2 operands from the parameter block, fully unrolled - 1.24 ms, 37 % of the stencil
// warps' issue slots lost to instruction fetch (24 KB of straight-line code per item); indexed LDC.64 in this loop - 1.46 ms.
template <int R>
ADP_DEVINL void fc_rows(const float *__restrict__ ws /*[9][64] then bias[64], shared*/, const float (*v)[3] /*R + 2 window rows*/, int cg,
                        uint32_t live_rows, uint4 (&o)[R]) {
  float2 a[R][4];
  {
    const float4 b0 = *reinterpret_cast<const float4 *>(ws + 9 * 64 + cg * 8), b1 = *reinterpret_cast<const float4 *>(ws + 9 * 64 + cg * 8 + 4);
    const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < R; ++r) a[r][c] = b[c];
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 w0 = *reinterpret_cast<const float4 *>(ws + t * 64 + cg * 8), w1 = *reinterpret_cast<const float4 *>(ws + t * 64 + cg * 8 + 4);
    const float2 w[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float q = v[r + t / 3][t % 3];
        ffma2(a[r][c], make_float2(q, q), w[c]);
      }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    o[r] = make_uint4(0u, 0u, 0u, 0u);                        // outside the image the operand is zero (what the TMA box fill gives)
    if ((live_rows >> r) & 1u)
      o[r] = make_uint4(relu_bf16x2(a[r][0].x, a[r][0].y), relu_bf16x2(a[r][1].x, a[r][1].y), relu_bf16x2(a[r][2].x, a[r][2].y),
                        relu_bf16x2(a[r][3].x, a[r][3].y));
  }
}

constexpr int EPI_STORE = 0, EPI_HEAD = 1, EPI_POOL = 2, EPI_BWD = 3;   // EPI_BWD = EPI_STORE with a mask (data-gradient twin)
constexpr int EPI_UPSUM = 4;   // data-gradient twin of an upsampled conv: the 2x2 sum of UpSampling2D's backward in the epilogue

// Software-pipelined walk over 16-column accumulator units u0, u0+step, ... < uend: the TMEM load
// of the next unit is in flight while `body` works on the current one.
template <typename F>
ADP_DEVINL void tmem_pipeline(uint32_t tbase, int u0, int step, int uend, F &&body) {
  uint32_t ra[16], rb[16];
  int u = u0;
  if (u < uend) ptx::tmem_ld16_issue(tbase + (uint32_t)u * 16u, ra);
  while (u < uend) {
    ptx::tmem_ld16_wait(ra);
    const int u2 = u + step;
    if (u2 < uend) ptx::tmem_ld16_issue(tbase + (uint32_t)u2 * 16u, rb);
    body(u, ra);
    if (u2 >= uend) break;
    ptx::tmem_ld16_wait(rb);
    const int u3 = u2 + step;
    if (u3 < uend) ptx::tmem_ld16_issue(tbase + (uint32_t)u3 * 16u, ra);
    body(u2, rb);
    u = u3;
  }
}

ADP_DEVINL void bias_relu16(const uint32_t (&r)[16], const float *sb, int relu, float (&f)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = *reinterpret_cast<const float4 *>(sb + 4 * q);
    f[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + b.x;
    f[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b.y;
    f[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b.z;
    f[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b.w;
  }
  if (relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
  }
}
ADP_DEVINL void load16_bf16(const __nv_bfloat16 *o, size_t plane, float (&f)[16]) {
  __align__(16) __nv_bfloat16 h[16];
  *reinterpret_cast<uint4 *>(h) = *reinterpret_cast<const uint4 *>(o);
  *reinterpret_cast<uint4 *>(h + 8) = *reinterpret_cast<const uint4 *>(o + plane);
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __bfloat162float(h[i]);
}
ADP_DEVINL void store16_bf16(__nv_bfloat16 *o, size_t plane, const float (&f)[16]) {
  __align__(16) __nv_bfloat16 h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = __float2bfloat16_rn(f[i]);
  *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<uint4 *>(h);
  *reinterpret_cast<uint4 *>(o + plane) = *reinterpret_cast<uint4 *>(h + 8);
}

// dropout_dense_kernel's arithmetic on 8 channels held in registers (group index g of the dense tensor)
ADP_DEVINL void dropout8(float *f, uint32_t g, uint32_t thr, float inv, uint32_t salt0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t h = hash_u32((g * 4u + (uint32_t)q) * 0x9E3779B1u ^ salt0);
    const float a = __bfloat162float(__float2bfloat16_rn(f[2 * q])), b = __bfloat162float(__float2bfloat16_rn(f[2 * q + 1]));
    f[2 * q] = (h & 0xFFFFu) < thr ? a * inv : 0.f;
    f[2 * q + 1] = (h >> 16) < thr ? b * inv : 0.f;
  }
}

// hi/lo store of the bf16x3 precision: lo_elems = element distance between a value's hi and lo halves (0 = plain bf16)
ADP_DEVINL void store16_out(__nv_bfloat16 *o, size_t lo_elems, size_t plane, const float (&f)[16]) {
  if (lo_elems == 0) { store16_bf16(o, plane, f); return; }
  __align__(16) __nv_bfloat16 h[16], l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    h[i] = __float2bfloat16_rn(f[i]);
    l[i] = __float2bfloat16_rn(f[i] - __bfloat162float(h[i]));
  }
  *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<uint4 *>(h);
  *reinterpret_cast<uint4 *>(o + plane) = *reinterpret_cast<uint4 *>(h + 8);
  *reinterpret_cast<uint4 *>(o + lo_elems) = *reinterpret_cast<uint4 *>(l);
  *reinterpret_cast<uint4 *>(o + lo_elems + plane) = *reinterpret_cast<uint4 *>(l + 8);
}

// item -> (variant, image, row block, strip)
ADP_DEVINL void decode_item(const ConvTcParams &p, int item, int &v, int &n, int &ty, int &tx) {
  v = item % p.nvar; int q = item / p.nvar;
  tx = q % p.ntx; q /= p.ntx;
  ty = q % p.nty; n = q / p.nty;
}

template <int NTAPS, int T, bool KYS, int EPI, bool FC = false>
__global__ void __launch_bounds__(FC ? kTcThreadsFc : kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmask, const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kThreads = FC ? kTcThreadsFc : kTcThreads;
  uint8_t *smask = smem + (size_t)p.S * p.stage_stride;           // EPI_BWD: mask_bufs x mask_bytes
  uint64_t *bars = reinterpret_cast<uint64_t *>(smask + (size_t)p.mask_bufs * p.mask_bytes);
  uint64_t *full = bars, *empty = full + p.S;
  uint64_t *acc_full = empty + p.S, *acc_empty = acc_full + 2;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 2);
  uint64_t *mask_full = acc_empty + 4, *mask_empty = mask_full + 2;
  uint64_t *b_full = acc_empty + 3;                               // resident weights have landed (bres)
  float *sbias = reinterpret_cast<float *>(acc_empty + 8);        // [nvar * N] (<= 704 floats)
  float *shead = sbias + 704;                                      // [2 * N] + 2, EPI_HEAD only
  // FC only: [2][kFcWinRows][kFcWinW] input windows (TMA destinations); their barriers are the mask tile's (unused by EPI_POOL)
  // (offset arithmetic on the shared array itself: a pointer that went through an integer loses its address space and
  // every access through it becomes a generic LD/ST)
  const size_t fc_off = ((size_t)p.S * p.stage_stride + (size_t)p.mask_bufs * p.mask_bytes + (size_t)(2 * p.S + 10) * 8 + (704 + 520) * 4 + 127) &
                        ~(size_t)127;
  float *swin = reinterpret_cast<float *>(smem + fc_off);
  float *sfw = swin + 2 * kFcWinRows * kFcWinW;                    // FC only: first-conv weights [9][64] + bias [64]
  if constexpr (FC) {
    const float *src = &p.fc.w[0][0];                              // w and b are contiguous in the parameter block
    for (int i = threadIdx.x; i < 10 * 64; i += kThreads) sfw[i] = src[i];
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.nvar * p.N; i += kThreads) {
    const int v = i / p.N;
    sbias[i] = p.bias[p.var[v].bias_off + (i - v * p.N)];
  }
  if constexpr (EPI == EPI_HEAD) {
    // softmax(z)[1] = 1 / (1 + exp(z0 - z1)): only the difference of the two 1x1 filters is needed, so the epilogue keeps
    // w0 - w1 (and b0 - b1) and reads it with 128-bit shared loads - the scalar form issued 2 x N 32-bit loads per pixel row,
    // 1.7 k shared-memory wavefronts per item that compete with the MMA operand fetch on the same port (ncu: 37 % of the
    // kernel's shared wavefronts were epilogue loads)
    for (int i = threadIdx.x; i < p.N; i += kThreads) shead[i] = p.head_w[i] - p.head_w[p.N + i];
    if (threadIdx.x == 0) shead[p.N] = p.head_b[0] - p.head_b[1];
  }

  if (warp == 0 && lane == 0) {
    // FC: a stage is full when the weight block has landed (producer thread: expect_tx) AND the five stencil warps have written A
    for (int i = 0; i < p.S; ++i) { ptx::mbar_init(&full[i], FC ? 1 + kFcWarps : 1); ptx::mbar_init(&empty[i], 2); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 8); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&mask_full[i], 1); ptx::mbar_init(&mask_empty[i], FC ? kFcWarps : 8); }
    ptx::mbar_init(b_full, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tmap);
    if constexpr (EPI == EPI_BWD || FC) ptx::prefetch_tmap(&tmask);
  }
  if (warp == 1) ptx::tmem_alloc_512(tmem_ptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int nitems = p.nb * p.nty * p.ntx * p.nvar;
  const int item0 = (int)blockIdx.x, item_end = nitems, item_step = (int)gridDim.x;   // strips of an image row adjacent in time
  if (warp == 0) {
    // ---------------- producer: activation boxes (TMA) + weight block (bulk copy) per stage ----------------
    if (ptx::elect_one()) {
      int st = 0; uint32_t ph = 0;
      long long t_w0 = 0; const long long t_start = clock64();
      int mit = 0;                                            // EPI_BWD: items whose mask tile has been requested
      const size_t blk_elems = p.b_bytes / 2;
      if (p.bres && item0 < item_end) {                       // weight blocks once: [variant][chunk] in wbase order, or this CTA's variant
        const int nblk = (p.bres == 2 ? 1 : p.nvar) * p.nchunks;
        const __nv_bfloat16 *wsrc = p.wpk + (p.bres == 2 ? (size_t)p.var[item0 % p.nvar].wbase * blk_elems : 0);
        ptx::mbar_expect_tx(b_full, p.bres_bytes);
        for (int b = 0; b < nblk; ++b)
          ptx::bulk_load_1d(smem + p.bres_off + (size_t)b * p.b_bytes, wsrc + (size_t)b * blk_elems, p.b_bytes, b_full);
      }
      for (int item = item0; item < item_end; item += item_step) {
        int v, n, ty, tx;
        decode_item(p, item, v, n, ty, tx);
        const int xg = tx * 16 - p.margin8, ys = ty * T + p.var[v].y0;
        const __nv_bfloat16 *w0 = p.wpk + (size_t)p.var[v].wbase * blk_elems;
        if (EPI == EPI_BWD && p.mask_bufs > 0) {
          const int mb = mit % p.mask_bufs; const uint32_t mph = (uint32_t)(mit / p.mask_bufs) & 1u;
          ptx::mbar_wait(&mask_empty[mb], mph ^ 1, 8);
          ptx::mbar_expect_tx(&mask_full[mb], p.mask_bytes);
          ptx::tma_load_5d(smask + (size_t)mb * p.mask_bytes, &tmask, &mask_full[mb], 0, tx * 16, p.var[v].out_cg, ty * T, n);
          ++mit;
        }
        if constexpr (FC) {                                   // input window of the item (rows ty*T - 2 .., columns tx*128 - 4 ..)
          const int wb = mit & 1; const uint32_t wph = (uint32_t)(mit >> 1) & 1u;
          ptx::mbar_wait(&mask_empty[wb], wph ^ 1, 13);
          ptx::mbar_expect_tx(&mask_full[wb], kFcWinBoxBytes);
          ptx::tma_load_5d(swin + (size_t)wb * (kFcWinRows * kFcWinW), &tmask, &mask_full[wb], tx * 128 - kFcWinX0, ty * T - 2, n, 0, 0);
          ++mit;
        }
        for (int c = 0; c < p.nchunks; ++c) {
          uint8_t *sa = smem + (size_t)st * p.stage_stride;
          { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&empty[st], ph ^ 1, 1); if (tc_dbg(p, 16)) t_w0 += clock64() - tw; }
          // p.dbg (ADP_TC_DEBUG, timing experiments only - results are wrong): 2 = weights only for the first item,
          // 8 = activations only for the first item, 1 = no epilogue stores, 4 = one MMA per stage, 16 = role timers
          const bool ld_b = (!tc_dbg(p, 2) || item == item0) && !p.bres, ld_a = !tc_dbg(p, 8) || item == item0;
          ptx::mbar_expect_tx(&full[st], ((ld_a && !FC) ? p.a_tx_bytes : 0u) + (ld_b ? p.b_bytes : 0u));
          int cgc = c * 2;                                     // first channel group of this chunk
          if (p.split) { const int cr = c / 3; cgc = cr * 2 + ((c - cr * 3) == 2 ? p.in_lo : 0); }
          if (ld_a && !FC)
            for (int b = 0; b < p.nbox; ++b)
              ptx::tma_load_5d(sa + (size_t)b * p.a_box_stride, &tmap, &full[st], 0, xg, cgc, ys + p.box_dy[b], n);
          if (ld_b) ptx::bulk_load_1d(sa + p.a_bytes, w0 + (size_t)c * blk_elems, p.b_bytes, &full[st]);
          if (++st == p.S) { st = 0; ph ^= 1; }
        }
      }
      if (tc_dbg(p, 16)) { p.dbg_out[blockIdx.x * 8 + 0] = t_w0; p.dbg_out[blockIdx.x * 8 + 1] = clock64() - t_start; }
    }
  } else if (warp == 1 || warp == 10) {
    // ---------------- MMA issuers ----------------
    // Two issuing threads alternate work items (A: even, B: odd).  An item owns its accumulator buffer (it & 1) and its
    // own run of pipeline stages, so the two instruction streams are independent; while one thread is blocked in a
    // tcgen05.mma issue slot the other prepares descriptors.  (Measured with tools/tc_timers.sh: one issuing thread was
    // > 85 % busy on the 1024^2 layers - ~50 cycles of descriptor arithmetic per MMA - while no pipe was saturated.)
    if (ptx::elect_one()) {
      const int issuer = (warp == 10) ? 1 : 0;
      const bool single_issuer = tc_dbg(p, 32) != 0;       // experiment switch: issuer B only observes
      int st = 0; uint32_t ph = 0;
      // The other issuer's stages are still OBSERVED: an mbarrier parity wait is only meaningful for a waiter that sees
      // every phase of the barrier in order, so the skipping thread waits for each stage's 'full' phase and then arrives on
      // its 'empty' barrier (count 2: the consuming issuer's tcgen05.commit + this arrive) - the producer cannot run a
      // stage two phases ahead of either issuer.
      auto skip_item = [&]() {
        for (int c = 0; c < p.nchunks; ++c) {
          ptx::mbar_wait(&full[st], ph, 5);
          ptx::mbar_arrive(&empty[st]);
          if (++st == p.S) { st = 0; ph ^= 1; }
        }
      };
      const uint32_t plane_a = (uint32_t)p.PW * 16u;          // bytes between the two channel-group planes (A)
      const uint32_t plane_b = (uint32_t)p.N * 16u;           // same for B
      // descriptor = hi (SBO = 128 B, version 1) : lo (start >> 4 | LBO >> 4 << 16); only the start moves
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((plane_a >> 4) << 16), b_lo0 = ((plane_b >> 4) << 16);
      const uint32_t row_step = (2u * plane_a) >> 4;          // next output row = next row slot
      const uint32_t smem0 = ptx::smem_u32(smem);
      const uint32_t n_cols = (uint32_t)p.N;
      const bool mma_all = !tc_dbg(p, 4);
      long long t_m0 = 0, t_m1 = 0, t_m2 = 0; const long long t_mstart = clock64();
      // first byte (>> 4) of the weight block of (variant, chunk): inside the stage, or in the resident region
      const uint32_t bres_lo = ((smem0 + p.bres_off) & 0x3FFFFu) >> 4, blk16 = p.b_bytes >> 4, a16 = p.a_bytes >> 4;
      if (p.bres && item0 < item_end) ptx::mbar_wait(b_full, 0, 15);
      if constexpr (KYS) {
        constexpr int KY = (NTAPS == 9) ? 3 : 2, KX = KY;
        const uint32_t plane_bs = (uint32_t)(KY * p.N) * 16u;  // stacked B: KY*N rows per channel-group plane
        const uint32_t bs_lo0 = ((plane_bs >> 4) << 16);
        const uint32_t n16 = (uint32_t)p.N;                    // N rows * 16 B, >> 4
        uint32_t a_xs[KX], b_blk[KX];
#pragma unroll
        for (int kx = 0; kx < KX; ++kx) {
          a_xs[kx] = (uint32_t)p.tap_xs[kx];                   // pixels == 16-byte units
          b_blk[kx] = ((uint32_t)kx * 2u * plane_bs) >> 4;
        }
        const uint32_t id1 = p.idesc_stack[0], id2 = p.idesc_stack[1], id3 = p.idesc_stack[2];
        int it = 0;
        for (int item = item0; item < item_end; item += item_step, ++it) {
          if (single_issuer ? issuer != 0 : (it & 1) != issuer) { skip_item(); continue; }
          const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
          { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&acc_empty[buf], acc_ph ^ 1, 3); if (tc_dbg(p, 16)) t_m0 += clock64() - tw; }
          ptx::tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)buf * (uint32_t)T * n_cols;
          const uint32_t v_off = (uint32_t)p.var[item % p.nvar].xs_add;
          const uint32_t wb0 = bres_lo + (p.bres == 2 ? 0u : (uint32_t)p.var[item % p.nvar].wbase * blk16);
          for (int c = 0; c < p.nchunks; ++c) {
            { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&full[st], ph, 4); if (tc_dbg(p, 16)) t_m1 += clock64() - tw; }
            ptx::tc_fence_after();
            const uint32_t s_lo = ((smem0 + (uint32_t)st * p.stage_stride) & 0x3FFFFu) >> 4;
            const uint32_t sa_lo = s_lo + v_off;
            const uint32_t sb_lo = p.bres ? wb0 + (uint32_t)c * blk16 : s_lo + a16;
#pragma unroll
            for (int kx = 0; kx < KX; ++kx) {
              if (kx == 0 && c == 0) {
                // first touch of every accumulator row: one MMA per (output row, vertical tap), overwrite on the first
#pragma unroll
                for (int o = 0; o < T; ++o)
#pragma unroll
                  for (int kyi = 0; kyi < KY; ++kyi) {
                    const uint64_t ad = ((uint64_t)desc_hi << 32) |
                                        (uint64_t)(a_lo0 | (sa_lo + a_xs[0] + (uint32_t)(o + kyi) * row_step));
                    const uint64_t bd = ((uint64_t)desc_hi << 32) |
                                        (uint64_t)(bs_lo0 | (sb_lo + b_blk[0] + (uint32_t)(KY - 1 - kyi) * n16));
                    if (mma_all || (o | kyi) == 0) ptx::mma_f16_ss(d0 + (uint32_t)o * n_cols, ad, bd, id1, (uint32_t)(kyi != 0));
                  }
              } else {
#pragma unroll
                for (int i = 0; i < T + KY - 1; ++i) {
                  const int o_lo = (i - (KY - 1)) > 0 ? (i - (KY - 1)) : 0;
                  const int o_hi = i < (T - 1) ? i : (T - 1);
                  const int cnt = o_hi - o_lo + 1;                       // 1..KY stacked vertical taps
                  const int sl = KY - 1 - i + o_lo;                      // first stacked tap slot
                  const uint64_t ad = ((uint64_t)desc_hi << 32) |
                                      (uint64_t)(a_lo0 | (sa_lo + a_xs[kx] + (uint32_t)i * row_step));
                  const uint64_t bd = ((uint64_t)desc_hi << 32) |
                                      (uint64_t)(bs_lo0 | (sb_lo + b_blk[kx] + (uint32_t)sl * n16));
                  if (mma_all) ptx::mma_f16_ss(d0 + (uint32_t)o_lo * n_cols, ad, bd, cnt == 1 ? id1 : (cnt == 2 ? id2 : id3), 1u);
                }
              }
            }
            { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mma_commit(&empty[st]); if (tc_dbg(p, 16)) t_m2 += clock64() - tw; }
            if (++st == p.S) { st = 0; ph ^= 1; }
          }
          { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mma_commit(&acc_full[buf]); if (tc_dbg(p, 16)) t_m2 += clock64() - tw; }
        }
      } else {
      uint32_t a_off[NTAPS], b_off[NTAPS];                    // byte offsets within a stage, >> 4
#pragma unroll
      for (int t = 0; t < NTAPS; ++t) {
        a_off[t] = ((uint32_t)p.tap_box[t] * p.a_box_stride + (uint32_t)(p.tap_row[t] * 2) * plane_a +
                    (uint32_t)p.tap_xs[t] * 16u) >> 4;
        b_off[t] = ((uint32_t)t * 2u * plane_b) >> 4;
      }
      const uint32_t idesc = p.idesc;
      int it = 0;
      for (int item = item0; item < item_end; item += item_step, ++it) {
        if (single_issuer ? issuer != 0 : (it & 1) != issuer) { skip_item(); continue; }
        const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
        { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&acc_empty[buf], acc_ph ^ 1, 3); if (tc_dbg(p, 16)) t_m0 += clock64() - tw; }
        ptx::tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)buf * (uint32_t)T * n_cols;
        const uint32_t v_off = (uint32_t)p.var[item % p.nvar].xs_add;   // pixels == 16-byte units
        const uint32_t wb0 = bres_lo + (p.bres == 2 ? 0u : (uint32_t)p.var[item % p.nvar].wbase * blk16);
        for (int c = 0; c < p.nchunks; ++c) {
          { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&full[st], ph, 4); if (tc_dbg(p, 16)) t_m1 += clock64() - tw; }
          ptx::tc_fence_after();
          const uint32_t s_lo = ((smem0 + (uint32_t)st * p.stage_stride) & 0x3FFFFu) >> 4;
          const uint32_t sa_lo = s_lo + v_off;
          const uint32_t sb_lo = p.bres ? wb0 + (uint32_t)c * blk16 : s_lo + a16;
#pragma unroll
          for (int t = 0; t < NTAPS; ++t) {
            const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo0 | (sb_lo + b_off[t]));
#pragma unroll
            for (int r = 0; r < T; ++r) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo0 | (sa_lo + a_off[t] + (uint32_t)r * row_step));
              if (mma_all || (t | r) == 0) ptx::mma_f16_ss(d0 + (uint32_t)r * n_cols, ad, bd, idesc, (uint32_t)((c | t) != 0));
            }
          }
          { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mma_commit(&empty[st]); if (tc_dbg(p, 16)) t_m2 += clock64() - tw; }
          if (++st == p.S) { st = 0; ph ^= 1; }
        }
        { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mma_commit(&acc_full[buf]); if (tc_dbg(p, 16)) t_m2 += clock64() - tw; }
      }
      }
      if (tc_dbg(p, 16) && issuer == 0) { p.dbg_out[blockIdx.x * 8 + 2] = t_m0; p.dbg_out[blockIdx.x * 8 + 3] = t_m1; p.dbg_out[blockIdx.x * 8 + 4] = clock64() - t_mstart; p.dbg_out[blockIdx.x * 8 + 7] = t_m2; }
    }
  } else if (!FC || warp < 10) {
    // ---------------- epilogue (warps 2..9; TMEM lane quarter = warp % 4, two warps per quarter) ----------------
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;
    const int NU = p.N >> 4;                                  // 16-column units per accumulator row
    long long t_e0 = 0; const long long t_estart = clock64();
    int it = 0;
    for (int item = item0; item < item_end; item += item_step, ++it) {
      const int buf = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
      int v, n, ty, tx;
      decode_item(p, item, v, n, ty, tx);
      const int x = tx * 128 + q4 * 32 + lane;
      const float *sb = sbias + v * p.N;
      const uint32_t t0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(buf * T * p.N);
      const size_t plane = (size_t)p.Wout * 8;
      const int ox = x * p.oscale + p.var[v].ox;
      // row-planar output row base of accumulator row r: ((((n*H + y)*cgs + cg)*W)*8
      // (no integer division and no 64-bit multiply chains per unit: row base + row stride, units walked incrementally)
      const size_t orstride = (size_t)p.oscale * p.out_cgs * plane;
      __nv_bfloat16 *const obase = p.out + (((size_t)n * p.Hout + (size_t)(ty * T * p.oscale + p.var[v].oy)) * p.out_cgs +
                                            p.out_cg0 + p.var[v].out_cg) * plane + (size_t)ox * 8;
      auto out_row = [&](int r) -> __nv_bfloat16 * { return obase + (size_t)r * orstride; };
      // unit u = row * NU + cu; this warp owns u = half, half + 2, ...
      auto unit_next = [&](int &row, int &cu) { cu += 2; while (cu >= NU) { cu -= NU; ++row; } };
      int row0 = 0, cu0 = half;
      while (cu0 >= NU) { cu0 -= NU; ++row0; }
      if constexpr (EPI == EPI_BWD) {
        // backward twin (data gradient): out = (acc + resid) * [mask > 0] * mask_scale
        const int UE = T * NU;
        if (p.mask_bufs > 0) {
        // the mask tile of the item sits in shared memory ([row][channel group][128 pixels][8]), landed there by the
        // producer's TMA box
        const int mb = it % p.mask_bufs; const uint32_t mph = (uint32_t)(it / p.mask_bufs) & 1u;
        { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&acc_full[buf], acc_ph, 6); if (tc_dbg(p, 16)) t_e0 += clock64() - tw; }
        ptx::mbar_wait(&mask_full[mb], mph, 9);
        ptx::tc_fence_after();
        const uint8_t *mtile = smask + (size_t)mb * p.mask_bytes + (size_t)(q4 * 32 + lane) * 16;
        int srow = row0, scu = cu0;
        tmem_pipeline(t0, half, 2, UE, [&](int, const uint32_t (&r)[16]) {
          const int row = srow, cu = scu;
          unit_next(srow, scu);
          float f[16];
          bias_relu16(r, sb + cu * 16, p.relu, f);
          if ((ty * T + row < p.Hin) && (x < p.Win)) {
            __nv_bfloat16 *o = out_row(row) + (size_t)(2 * cu) * plane;
            if (p.resid) {
              float g[16];
              load16_bf16(p.resid + (o - p.out), plane, g);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += g[i];
            }
            const uint8_t *mp = mtile + (size_t)(row * (p.N >> 3) + 2 * cu) * 2048;      // plane = 128 px * 16 B
            __align__(16) __nv_bfloat16 mh[16];
            *reinterpret_cast<uint4 *>(mh) = *reinterpret_cast<const uint4 *>(mp);
            *reinterpret_cast<uint4 *>(mh + 8) = *reinterpret_cast<const uint4 *>(mp + 2048);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __bfloat162float(mh[i]) > 0.f ? f[i] * p.mask_scale : 0.f;
            store16_bf16(o, plane, f);
          }
        });
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&mask_empty[mb]);
        } else {
        // wide accumulators (N > 96: the mask tile and a useful pipeline depth do not fit shared memory together): the
        // mask words of the first PF units this warp owns are requested from global memory before waiting for the
        // accumulator, so their latency overlaps the MMAs of the item; later units load unprefetched
        constexpr int PF = 6;
        uint4 mk[PF][2];
        int prow = row0, pcu = cu0;
#pragma unroll
        for (int k = 0; k < PF; ++k) {
          const int u = half + 2 * k;
          mk[k][0] = make_uint4(0, 0, 0, 0); mk[k][1] = mk[k][0];
          if (u < UE) {
            const int row = prow, cu = pcu;
            unit_next(prow, pcu);
            if ((ty * T + row < p.Hin) && (x < p.Win)) {
              const __nv_bfloat16 *m = p.mask + ((out_row(row) + (size_t)(2 * cu) * plane) - p.out);
              mk[k][0] = __ldg(reinterpret_cast<const uint4 *>(m));
              mk[k][1] = __ldg(reinterpret_cast<const uint4 *>(m + plane));
            }
          }
        }
        { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&acc_full[buf], acc_ph, 6); if (tc_dbg(p, 16)) t_e0 += clock64() - tw; }
        ptx::tc_fence_after();
        uint32_t acc[2][16];
        if (half < UE) ptx::tmem_ld16_issue(t0 + (uint32_t)half * 16u, acc[0]);
        prow = row0; pcu = cu0;
#pragma unroll
        for (int k = 0; k < PF; ++k) {
          const int u = half + 2 * k;
          if (u < UE) {
            ptx::tmem_ld16_wait(acc[k & 1]);
            if (u + 2 < UE) ptx::tmem_ld16_issue(t0 + (uint32_t)(u + 2) * 16u, acc[(k + 1) & 1]);
            const int row = prow, cu = pcu;
            unit_next(prow, pcu);
            float f[16];
            bias_relu16(acc[k & 1], sb + cu * 16, p.relu, f);
            if ((ty * T + row < p.Hin) && (x < p.Win)) {
              __nv_bfloat16 *o = out_row(row) + (size_t)(2 * cu) * plane;
              if (p.resid) {
                float g[16];
                load16_bf16(p.resid + (o - p.out), plane, g);
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] += g[i];
              }
              const __nv_bfloat16 *mh0 = reinterpret_cast<const __nv_bfloat16 *>(&mk[k][0]);
              const __nv_bfloat16 *mh1 = reinterpret_cast<const __nv_bfloat16 *>(&mk[k][1]);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                f[i] = __bfloat162float(mh0[i]) > 0.f ? f[i] * p.mask_scale : 0.f;
                f[8 + i] = __bfloat162float(mh1[i]) > 0.f ? f[8 + i] * p.mask_scale : 0.f;
              }
              store16_bf16(o, plane, f);
            }
          }
        }
        for (int u = half + 2 * PF; u < UE; u += 2) {
          const int row = u / NU, cu = u - row * NU;
          float f[16], m[16];
          ptx::tmem_ld16(t0 + (uint32_t)u * 16u, f);
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] += sb[cu * 16 + i];
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if ((ty * T + row < p.Hin) && (x < p.Win)) {
            __nv_bfloat16 *o = out_row(row) + (size_t)(2 * cu) * plane;
            if (p.resid) {
              load16_bf16(p.resid + (o - p.out), plane, m);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += m[i];
            }
            load16_bf16(p.mask + (o - p.out), plane, m);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = m[i] > 0.f ? f[i] * p.mask_scale : 0.f;
            store16_bf16(o, plane, f);
          }
        }
        }
      } else {
      // EPI_UPSUM: the low-resolution mask / second-gradient words of the (at most four) units this warp owns are requested
      // before the wait for the accumulator, so their latency overlaps the item's MMAs (even lanes own a 2x2 block)
      constexpr int UPF = 4;
      uint4 um[EPI == EPI_UPSUM ? UPF : 1][2], ur[EPI == EPI_UPSUM ? UPF : 1][2];
      size_t up_off = 0, up_plane = 0;
      bool up_live = false;
      if constexpr (EPI == EPI_UPSUM) {
        up_live = (ty * T + 1 < p.Hin) && (x < p.Win) && !(lane & 1) && !tc_dbg(p, 1);
        up_plane = (size_t)(p.Wout >> 1) * 8;
        up_off = (((size_t)n * (p.Hout >> 1) + ((ty * T) >> 1)) * p.pool_cgs + p.pool_cg0 + p.var[v].out_cg) * up_plane + (size_t)(x >> 1) * 8;
#pragma unroll
        for (int k = 0; k < UPF; ++k) {
          const int cu = half + 2 * k;
          um[k][0] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u); um[k][1] = um[k][0];      // "mask > 0" when there is no mask
          ur[k][0] = make_uint4(0u, 0u, 0u, 0u); ur[k][1] = ur[k][0];
          if (up_live && cu < NU) {
            const size_t o = up_off + (size_t)(2 * cu) * up_plane;
            if (p.mask) { um[k][0] = __ldg(reinterpret_cast<const uint4 *>(p.mask + o)); um[k][1] = __ldg(reinterpret_cast<const uint4 *>(p.mask + o + up_plane)); }
            if (p.resid) { ur[k][0] = __ldg(reinterpret_cast<const uint4 *>(p.resid + o)); ur[k][1] = __ldg(reinterpret_cast<const uint4 *>(p.resid + o + up_plane)); }
          }
        }
      }
      { const long long tw = tc_dbg(p, 16) ? clock64() : 0; ptx::mbar_wait(&acc_full[buf], acc_ph, 6); if (tc_dbg(p, 16)) t_e0 += clock64() - tw; }
      ptx::tc_fence_after();
      if constexpr (EPI == EPI_STORE) {
        int srow = row0, scu = cu0;
        tmem_pipeline(t0, half, 2, T * NU, [&](int, const uint32_t (&r)[16]) {
          const int row = srow, cu = scu;
          unit_next(srow, scu);
          float f[16];
          bias_relu16(r, sb + cu * 16, p.relu, f);
          if ((ty * T + row < p.Hin) && (x < p.Win) && !tc_dbg(p, 1)) {
            __nv_bfloat16 *o = out_row(row) + (size_t)(2 * cu) * plane;
            if (p.resid) {
              float g[16];
              load16_bf16(p.resid + (o - p.out), plane, g);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += g[i];
            }
            if (p.mask) {
              float m[16];
              load16_bf16(p.mask + (o - p.out), plane, m);
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = m[i] > 0.f ? f[i] * p.mask_scale : 0.f;
            }
            if (p.drop_thr) {
              const uint32_t g = (uint32_t)((size_t)(o - p.out) >> 3);
              dropout8(f, g, p.drop_thr, p.drop_inv, p.drop_salt);
              dropout8(f + 8, g + (uint32_t)(plane >> 3), p.drop_thr, p.drop_inv, p.drop_salt);
            }
            store16_out(o, (size_t)p.out_lo * plane, plane, f);
          }
        });
      } else if constexpr (EPI == EPI_HEAD) {
        // rows split between the two warps of a quarter so that one thread sees all channels of its pixel
        for (int row = half; row < T; row += 2) {
          float zd = 0.f;                                       // z0 - z1
          tmem_pipeline(t0, row * NU, 1, row * NU + NU, [&](int u, const uint32_t (&r)[16]) {
            const int cu = u - row * NU;
            float f[16];
            bias_relu16(r, sb + cu * 16, p.relu, f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 w = *reinterpret_cast<const float4 *>(shead + cu * 16 + 4 * q);
              zd = fmaf(f[4 * q + 0], w.x, zd); zd = fmaf(f[4 * q + 1], w.y, zd);
              zd = fmaf(f[4 * q + 2], w.z, zd); zd = fmaf(f[4 * q + 3], w.w, zd);
            }
          });
          const int y = ty * T + row;
          if (y < p.Hin && x < p.Win && !tc_dbg(p, 1))
            p.prob[((size_t)n * p.Hout + y) * p.Wout + x] = 1.f / (1.f + expf(zd + shead[p.N]));
        }
      } else if constexpr (EPI == EPI_UPSUM) {
        // Data gradient of UpSampling2D + conv (T = 2): the full-resolution gradient is never written.  Each value is rounded to
        // bf16 as the store would have done, the 2x2 block is summed in upsample2_bwd_kernel's pairwise order (column sums first),
        // then the second gradient into the low-resolution tensor (deep-supervision head) is added and the ReLU' / dropout mask of
        // the producing layer applied - the statements of upsample2_bwd_kernel, bit for bit.  pool_out = low-resolution gradient;
        // resid / mask have its layout.
        static_assert(EPI != EPI_UPSUM || T == 2, "the 2x2 sum needs one row pair per item");
#pragma unroll
        for (int k = 0; k < UPF; ++k) {
          const int cu = half + 2 * k;
          if (cu < NU) {
            uint32_t ra[16], rb[16];
            ptx::tmem_ld16_issue(t0 + (uint32_t)(cu * 16), ra);
            ptx::tmem_ld16_issue(t0 + (uint32_t)(p.N + cu * 16), rb);
            ptx::tmem_ld16_wait(ra);
            ptx::tmem_ld16_wait(rb);
            float fa[16], fb[16];
            bias_relu16(ra, sb + cu * 16, p.relu, fa);
            bias_relu16(rb, sb + cu * 16, p.relu, fb);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float t = __bfloat162float(__float2bfloat16_rn(fa[i])) + __bfloat162float(__float2bfloat16_rn(fb[i]));
              fa[i] = t + __shfl_xor_sync(0xffffffffu, t, 1);       // even lane: (this column's two rows) + (the next column's)
            }
            if (up_live) {
              const __nv_bfloat16 *m0 = reinterpret_cast<const __nv_bfloat16 *>(&um[k][0]), *m1 = reinterpret_cast<const __nv_bfloat16 *>(&um[k][1]);
              const __nv_bfloat16 *r0 = reinterpret_cast<const __nv_bfloat16 *>(&ur[k][0]), *r1 = reinterpret_cast<const __nv_bfloat16 *>(&ur[k][1]);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (p.resid) { fa[i] += __bfloat162float(r0[i]); fa[8 + i] += __bfloat162float(r1[i]); }
                if (p.mask) {
                  fa[i] = __bfloat162float(m0[i]) > 0.f ? fa[i] * p.mask_scale : 0.f;
                  fa[8 + i] = __bfloat162float(m1[i]) > 0.f ? fa[8 + i] * p.mask_scale : 0.f;
                }
              }
              store16_bf16(p.pool_out + up_off + (size_t)(2 * cu) * up_plane, up_plane, fa);
            }
          }
        }
      } else {
        // EPI_POOL: unit = (row pair k, 16 channels); both rows are stored, their max is reduced with
        // the neighbouring column (lane ^ 1) and even lanes write the pooled pixel
        const int HU = (T / 2) * NU;
        int pk = row0, pc = cu0;
        for (int pu = half; pu < HU; pu += 2) {
          const int k = pk, cu = pc;
          unit_next(pk, pc);
          uint32_t ra[16], rb[16];
          ptx::tmem_ld16_issue(t0 + (uint32_t)((2 * k) * p.N + cu * 16), ra);
          ptx::tmem_ld16_issue(t0 + (uint32_t)((2 * k + 1) * p.N + cu * 16), rb);
          ptx::tmem_ld16_wait(ra);
          ptx::tmem_ld16_wait(rb);
          float fa[16], fb[16];
          bias_relu16(ra, sb + cu * 16, p.relu, fa);
          bias_relu16(rb, sb + cu * 16, p.relu, fb);
          const bool live = (ty * T + 2 * k + 1 < p.Hin) && (x < p.Win) && !tc_dbg(p, 1);
          if (live) {
            store16_out(out_row(2 * k) + (size_t)(2 * cu) * plane, (size_t)p.out_lo * plane, plane, fa);
            store16_out(out_row(2 * k + 1) + (size_t)(2 * cu) * plane, (size_t)p.out_lo * plane, plane, fb);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float m = fmaxf(fa[i], fb[i]);
            fa[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
          }
          if (live && !(lane & 1)) {
            const int py = (ty * T + 2 * k) >> 1, px = x >> 1;
            const size_t pplane = (size_t)(p.Wout >> 1) * 8;
            const size_t prow = ((size_t)n * (p.Hout >> 1) + py) * p.pool_cgs + p.pool_cg0 + p.var[v].out_cg;
            store16_out(p.pool_out + prow * pplane + (size_t)px * 8 + (size_t)(2 * cu) * pplane, (size_t)p.pool_lo * pplane, pplane, fa);
          }
        }
      }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[buf]);
    }
    if (tc_dbg(p, 16) && warp == 2 && lane == 0) { p.dbg_out[blockIdx.x * 8 + 5] = t_e0; p.dbg_out[blockIdx.x * 8 + 6] = clock64() - t_estart; }
  } else {
    // ---------------- FC: stencil warps 11..15 compute the A operand (first conv) into the pipeline stages ----------------
    // Item = T (= 4) output rows x 128 columns: the A box is 6 rows x 130 columns of the first conv's output (stage row r <->
    // image row ty*4 - 1 + r, box pixel 7 + j <-> image column tx*128 - 1 + j), which needs an 8 x 132 window of the input:
    // the producer thread lands it in shared memory with one TMA box per item (two buffers).  Warps 0..3 of the group own 32
    // columns each - all rows of a column in one thread, so every window value feeds up to nine FMAs from a register - and
    // the fifth warp's first 12 lanes the two halo columns (one row each).
    if constexpr (FC) {
      const int sw = warp - 11;
      const uint32_t plane_a = (uint32_t)p.PW * 16u, row_stride = 2u * plane_a;
      const bool main_warp = sw < 4;
      const bool act = main_warp || lane < 12;
      const int j = main_warp ? sw * 32 + lane : 128 + (lane & 1);        // box column 0..129
      const int r0 = (main_warp || !act) ? 0 : (lane >> 1);               // first box row of this thread
      int st = 0; uint32_t ph = 0; int it = 0;
      for (int item = item0; item < item_end; item += item_step, ++it) {
        int vv, n, ty, tx;
        decode_item(p, item, vv, n, ty, tx);
        const int wbuf = it & 1; const uint32_t wph = (uint32_t)(it >> 1) & 1u;
        const float *wb = swin + wbuf * (kFcWinRows * kFcWinW) + r0 * kFcWinW + j + (kFcWinX0 - 2);
        float v[8][3];
        ptx::mbar_wait(&mask_full[wbuf], wph, 14);
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) v[r][kx] = (main_warp || r < 3) ? wb[r * kFcWinW + kx] : 0.f;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&mask_empty[wbuf]);    // the values are in registers: the producer may refill the buffer
        const int X = tx * 128 - 1 + j;
        uint32_t live = 0;
        if (X >= 0 && X < p.Win)
#pragma unroll
          for (int r = 0; r < 6; ++r) { const int Y = ty * T - 1 + r0 + r; live |= (Y >= 0 && Y < p.Hin) ? (1u << r) : 0u; }
        // One pass per group of 8 output channels, NOT unrolled: the loop body stays resident in the instruction cache - the
        // fully unrolled form (24 KB of straight-line code per item, four warps at four different places of it) spent 37 % of
        // its issue slots waiting for instruction fetch (ncu, stall_no_inst).  The FMAs of a pass run BEFORE the wait for the
        // stage, so the tensor pipe's drain of that stage overlaps them.
#pragma unroll 1
        for (int cg = 0; cg < 2 * p.nchunks; ++cg) {
          const uint32_t col = (uint32_t)(7 + j) * 16u + (uint32_t)(cg & 1) * plane_a;
          uint8_t *sa = smem + (size_t)st * p.stage_stride + col;
          if (main_warp) {
            uint4 o[6];
            fc_rows<6>(sfw, v, cg, live, o);
            if (!(cg & 1)) ptx::mbar_wait(&empty[st], ph ^ 1, 11);
#pragma unroll
            for (int r = 0; r < 6; ++r) *reinterpret_cast<uint4 *>(sa + (size_t)r * row_stride) = o[r];
          } else {
            uint4 o[1];
            if (act) fc_rows<1>(sfw, v, cg, live, o);
            if (!(cg & 1)) ptx::mbar_wait(&empty[st], ph ^ 1, 12);
            if (act) *reinterpret_cast<uint4 *>(sa + (size_t)r0 * row_stride) = o[0];
          }
          if (cg & 1) {
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&full[st]);
            if (++st == p.S) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc_512(tmem_base);
}

}  // namespace adp
