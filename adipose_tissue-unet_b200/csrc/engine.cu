// engine.cu — libadipose_b200.so: engine object, weight packing, forward orchestration and the
// C ABI of include/adipose_b200.h.  Host code is plain C++17 + CUDA runtime; the only driver API
// symbol (cuTensorMapEncodeTiled) is resolved at run time so the library loads on a CPU-only box.
#include "../../include/adipose_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "io_frontend.cuh"
#include "kernels_post.cuh"
#include "kernels_refine.cuh"
#include "kernels_simt.cuh"
#include "kernels_train.cuh"
#include "wgrad_tc.cuh"
#include "tiff_lzw.h"

using namespace adp;

namespace {

thread_local std::string g_last_error;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    ADP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) throw Error(ADP_ECUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// nvJPEG is resolved at run time (dlopen), like the one driver entry point above: the library loads on machines without
// it and contains sm_100a code of this repository only.
struct NvjpegApi {
  decltype(&nvjpegCreateEx) CreateEx = nullptr;
  decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
  decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
  decltype(&nvjpegDecode) Decode = nullptr;
  decltype(&nvjpegJpegStateDestroy) JpegStateDestroy = nullptr;
  decltype(&nvjpegDestroy) Destroy = nullptr;
};
const NvjpegApi &nvjpeg_api() {
  static NvjpegApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = nullptr;
    for (const char *name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"})
      if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
    if (h) {
      api.CreateEx = reinterpret_cast<decltype(api.CreateEx)>(dlsym(h, "nvjpegCreateEx"));
      api.JpegStateCreate = reinterpret_cast<decltype(api.JpegStateCreate)>(dlsym(h, "nvjpegJpegStateCreate"));
      api.GetImageInfo = reinterpret_cast<decltype(api.GetImageInfo)>(dlsym(h, "nvjpegGetImageInfo"));
      api.Decode = reinterpret_cast<decltype(api.Decode)>(dlsym(h, "nvjpegDecode"));
      api.JpegStateDestroy = reinterpret_cast<decltype(api.JpegStateDestroy)>(dlsym(h, "nvjpegJpegStateDestroy"));
      api.Destroy = reinterpret_cast<decltype(api.Destroy)>(dlsym(h, "nvjpegDestroy"));
    }
  }
  if (!api.CreateEx || !api.JpegStateCreate || !api.GetImageInfo || !api.Decode || !api.JpegStateDestroy || !api.Destroy)
    throw Error(ADP_ESTATE, "nvJPEG (libnvjpeg.so.12) is not available on this machine: decode the tiles on the host (--decode cv2)");
  return api;
}

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  void ensure(size_t n) {
    if (n <= bytes) return;
    release();
    ADP_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
  }
  ~DevBuf() { release(); }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

float bf16_round(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

struct FwTable {          // per-launch forward table, passed by value
  int tile[64];
  int op[64];
  long long origin[64];
};

__global__ void fw_table_store(FwTable t, int *tile, int *op, long long *origin, int n) {
  int i = threadIdx.x;
  if (i < n) { tile[i] = t.tile[i]; op[i] = t.op[i]; origin[i] = t.origin[i]; }
}

struct ProfEntry { int64_t launches = 0; double ms = 0, flops = 0, bytes = 0; };

struct HostWeight { std::vector<float> k; int64_t shape[4] = {0, 0, 0, 0}; std::vector<float> b; bool set = false; };

struct ConvLayer {
  std::string name;
  int cin, cout, dil;          // real channels
  bool up = false;             // input is UpSampling2D(2x2) of the source buffer
  int skip = 0;                // >0: concat-fed; first `skip` real input channels come from the skip path
  int cin_pad = 0, cout_pad = 0;
  DevBuf w_simt, bias, w_tc;   // [9][cin_pad][cout_pad] fp32 | [cout_pad] fp32 | packed bf16 blocks
  ConvTcParams tc{};           // geometry-independent part filled at pack time
  bool tc_ready = false;
};

// Activation buffers of one forward.  Inference aliases buffers whose lifetimes do not overlap
// (a1 holds down1_conv1 and later up1_conv2, ...); training keeps every tensor for the backward pass.
struct Acts {
  DevBuf *d1a, *cat1, *u1b, *u1c, *pl1;
  DevBuf *d2a, *cat2, *u2b, *u2c, *pl2;
  DevBuf *d3a, *cat3, *u3b, *u3c, *pl3;
  DevBuf *t[6], *ts, *prob;
  int cap;                       // images the buffers are sized for (outermost TMA dimension)
};

struct TrainState;

}  // namespace

struct adp_engine {
  int device = 0, prec = 0, init_nb = 44, max_fw = 16;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  int c[4], cp[4];                 // real / padded channel counts at the four levels
  std::map<std::string, HostWeight> hw;
  std::vector<ConvLayer> layers;   // the 20 generic 3x3 convs (all but down1_conv1 and the head)
  DevBuf w_first, b_first, w_head, b_head;
  bool packed = false;

  // activation arena for tile size S
  int S = 0;
  size_t esz = 4;
  DevBuf a1, cat1, b1, pl1, a2, cat2, b2, pl2, a3, cat3, b3, pl3, t[6], ts, prob;
  DevBuf in_stage, out_stage, fwt_tile, fwt_op, fwt_origin;
  DevBuf in_stage2[2], out_stage2[2];            // double-buffered host<->device staging of run_tiles
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in_ready[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_out_ready[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
  DevBuf fin_prob, fin_mask, fin_gt;             // persistent staging of run_finalize
  std::map<std::string, CUtensorMap> tmaps;
  int last_nfw = 0;
  Acts acts{};                                   // inference buffer assignment
  TrainState *tr = nullptr;                      // training arena (adp_train_begin .. adp_train_end)

  // whole-slide accumulator
  bool wsi_on = false;
  int wsi_rows = 0, wsi_W = 0, wsi_y0 = 0, wsi_tile = 0, wsi_mode = 0;
  DevBuf wsi_acc, wsi_wsum, wsi_window, counts, misc;
  // deferred boundary zone (adp_wsi_push_from_slide with defer_below_row): TTA-mean probabilities of the tiles that touch
  // the rows another strip's partial sums must reach FIRST, kept until adp_wsi_replay_deferred
  struct Deferred { std::unique_ptr<DevBuf> probs; std::vector<int32_t> ys, xs; int below = 0; };
  std::vector<Deferred> wsi_deferred;
  // auxiliary whole-slide planes (RGB mosaic, blended ground truth: io_frontend.cuh) sharing the probability accumulator's weights
  DevBuf wsi_aux; int wsi_naux = 0;
  // nvJPEG front-end (adp_jpeg_decode): handle, decoder state and the device-resident decoded tiles
  nvjpegHandle_t jpeg_h = nullptr; nvjpegJpegState_t jpeg_st = nullptr;
  DevBuf jpeg_gray, jpeg_rgb, aux_stage;

  // profiling
  bool prof = false;
  std::map<std::string, ProfEntry> prof_rows;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  int dbg = 0;
  bool fuse_head = true, fuse_pool = true;   // tcgen05 path only
  bool fuse_upsum = true;                    // tcgen05 backward: UpSampling2D's 2x2 gradient sum in the data-gradient twin's epilogue (T = 2 layers)
  bool fuse_dropout = true;                  // tcgen05 training forward: hash dropout applied in the producing conv's epilogue
  bool fuse_first = false;                   // tcgen05 inference: first conv computed inside down1_conv2 (conv_tc.cuh, FC variant): bit-identical,
                                             // opt-in.  Kernel time is break-even un-throttled; under the power cap the 3.2 GB less HBM traffic
                                             // per 16 forwards gave +0.1..1.4 % in bench.py A/B runs - inside the box-to-box spread, and it moves
                                             // the first conv's time into the dominant kernel (bench.py "first_conv_fusion")
  bool resident_weights = true;              // tcgen05 convs whose packed weights fit next to >= 4 activation stages keep them in shared
                                             // memory for the life of the persistent CTA (ConvTcParams::bres); bit-identical
  FirstConvFuse fc_host;                     // its fp32 weights / bias as kernel parameters (filled by pack_all)
  bool kys = true;                           // ky-stacked MMA issue for the N <= 128 layers
  bool split = false;                        // ADP_PREC_BF16X3: hi/lo bf16 activations and weights, three GEMM passes (conv_tc.cuh)
  int prec_public = 0;                       // what adp_precision() reports
  int mul() const { return split ? 2 : 1; }  // physical channel groups per logical group
  LossRecipe loss;                           // adp_train_set_loss: hard-example mining / label smoothing of the training loss
  bool deep_sup = false;                     // adp_train_set_deep_supervision: aux_out1 / aux_out2 heads while training
  float ds_w[3] = {1.0f, 0.4f, 0.3f};        // loss weights main / aux1 / aux2 (train_adipose_unet_v3.py:858-872)
  bool train_eval_mode = false;              // adp_set_option("train_eval_mode"): adp_train_forward without Dropout (validation pass)
  bool train_accuracy = false;               // adp_set_option("train_accuracy"): count Keras' binary_accuracy in every backward
  bool wgrad_simt = false, dgrad_simt = false;   // bf16 training: CUDA-core cross-check of the tcgen05 backward kernels

  // Grid of a grid-stride elementwise kernel: exactly one resident wave (blocks per SM from the occupancy calculator x SMs),
  // or fewer blocks when the work is smaller.  A fixed "SMs x 8 / x 16" grid ran 1.33 .. 5.33 waves for kernels that
  // hold 3, 5 or 6 blocks per SM, i.e. a last wave at 20-60 % occupancy.
  std::map<const void *, int> occ_cache;
  template <typename K> int wave_grid(K kern, size_t blocks_needed, int block = 256, size_t smem = 0) {
    const void *key = reinterpret_cast<const void *>(kern);
    auto it = occ_cache.find(key);
    if (it == occ_cache.end()) {
      int nb = 0;
      ADP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, block, smem));
      it = occ_cache.emplace(key, std::max(nb, 1)).first;
    }
    return (int)std::max<size_t>(1, std::min<size_t>(blocks_needed, (size_t)num_sms * it->second));
  }

  template <typename F> void launch(const char *kind, double flops, double bytes, F &&f) {
    if (prof) ADP_CUDA(cudaEventRecord(ev0, stream));
    f();
    ADP_CUDA(cudaGetLastError());
    ++launches;
    if (prof) {
      ADP_CUDA(cudaEventRecord(ev1, stream));
      ADP_CUDA(cudaEventSynchronize(ev1));
      float ms = 0;
      ADP_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
      auto &r = prof_rows[kind];
      r.launches++; r.ms += ms; r.flops += flops; r.bytes += bytes;
    }
  }
};

namespace {

// ------------------------------------------------------------------------------------------------
// graph description (train_adipose_unet_v3.py:668-709)
void build_layers(adp_engine *e) {
  const int c1 = e->c[0], c2 = e->c[1], c4 = e->c[2], c8 = e->c[3];
  auto add = [&](const char *n, int ci, int co, int d, bool up, int skip) {
    ConvLayer L;
    L.name = n; L.cin = ci; L.cout = co; L.dil = d; L.up = up; L.skip = skip;
    L.cout_pad = pad16(co);
    L.cin_pad = skip ? 2 * pad16(skip) : pad16(ci);
    e->layers.push_back(std::move(L));
  };
  add("down1_conv2", c1, c1, 1, false, 0);
  add("down2_conv1", c1, c2, 1, false, 0);
  add("down2_conv2", c2, c2, 1, false, 0);
  add("down3_conv1", c2, c4, 1, false, 0);
  add("down3_conv2", c4, c4, 1, false, 0);
  add("dilate1", c4, c8, 1, false, 0);
  add("dilate2", c8, c8, 2, false, 0);
  add("dilate3", c8, c8, 4, false, 0);
  add("dilate4", c8, c8, 8, false, 0);
  add("dilate5", c8, c8, 16, false, 0);
  add("dilate6", c8, c8, 32, false, 0);
  add("up3_conv1", c8, c4, 1, true, 0);
  add("up3_conv2", c8, c4, 1, false, c4);
  add("up3_conv3", c4, c4, 1, false, 0);
  add("up2_conv1", c4, c2, 1, true, 0);
  add("up2_conv2", c4, c2, 1, false, c2);
  add("up2_conv3", c2, c2, 1, false, 0);
  add("up1_conv1", c2, c1, 1, true, 0);
  add("up1_conv2", c2, c1, 1, false, c1);
  add("up1_conv3", c1, c1, 1, false, 0);
}

ConvLayer &layer(adp_engine *e, const std::string &n) {
  for (auto &L : e->layers)
    if (L.name == n) return L;
  throw Error(ADP_EINVAL, "unknown layer " + n);
}

const char *const kAuxNames[2] = {"aux_out1", "aux_out2"};   // deep-supervision heads: 1x1 convs on up3 / up2 (:712-725)
const char *const kAllNames[22] = {"down1_conv1", "down1_conv2", "down2_conv1", "down2_conv2", "down3_conv1",
                                   "down3_conv2", "dilate1", "dilate2", "dilate3", "dilate4", "dilate5", "dilate6",
                                   "up3_conv1", "up3_conv2", "up3_conv3", "up2_conv1", "up2_conv2", "up2_conv3",
                                   "up1_conv1", "up1_conv2", "up1_conv3", "output_softmax"};

// ------------------------------------------------------------------------------------------------
// tcgen05 plan: everything of ConvTcParams that does not depend on the tile size
void plan_tc(ConvLayer &L, bool allow_kys = true, bool split = false) {
  ConvTcParams &p = L.tc;
  memset(&p, 0, sizeof(p));
  const int N = (L.cout_pad <= 256) ? L.cout_pad : L.cout_pad / 2;
  const int nsplit = L.cout_pad / N;            // 1, or 2 for the 352-wide bottleneck
  ADP_REQUIRE(N % 16 == 0 && N <= 256 && nsplit <= 2, "unsupported output channel count for tcgen05 path");
  p.N = N;
  int T = (N <= 64) ? 4 : (N <= 128 ? 2 : 1);
  if (L.dil > 1) T = 1;
  p.T = T;
  if (L.up) {
    // four output parities, 2x2 taps each (see conv_tc.cuh); the activation box starts 8 pixels left
    // of the strip (TMA boxes are whole 8-pixel groups), tap (ry,rx) of parity (py,px) reads source
    // pixel (Y + py - 1 + ry, X + px - 1 + rx)
    p.nvar = 4; p.ntaps = 4; p.nbox = 1; p.BR = T + 1; p.margin8 = 1; p.PW = 144; p.oscale = 2;
    for (int v = 0; v < 4; ++v) {
      const int py = v >> 1, px = v & 1;
      p.var[v].y0 = py - 1; p.var[v].xs_add = px; p.var[v].oy = py; p.var[v].ox = px;
      p.var[v].bias_off = 0; p.var[v].out_cg = 0;
    }
    for (int t = 0; t < 4; ++t) { p.tap_box[t] = 0; p.tap_row[t] = t >> 1; p.tap_xs[t] = 7 + (t & 1); }
    p.box_dy[0] = 0;
  } else {
    p.nvar = nsplit; p.ntaps = 9; p.oscale = 1;
    const int margin = (L.dil + 7) / 8 * 8;
    p.margin8 = margin / 8; p.PW = 128 + 2 * margin;
    for (int v = 0; v < nsplit; ++v) {
      p.var[v].y0 = -L.dil; p.var[v].xs_add = 0; p.var[v].oy = 0; p.var[v].ox = 0;
      p.var[v].bias_off = v * N; p.var[v].out_cg = v * N / 8;
    }
    if (L.dil == 1) {
      p.nbox = 1; p.BR = T + 2; p.box_dy[0] = 0;
      for (int t = 0; t < 9; ++t) { p.tap_box[t] = 0; p.tap_row[t] = t / 3; p.tap_xs[t] = margin + (t % 3 - 1); }
    } else {
      p.nbox = 3; p.BR = T;
      for (int b = 0; b < 3; ++b) p.box_dy[b] = b * L.dil;
      for (int t = 0; t < 9; ++t) { p.tap_box[t] = t / 3; p.tap_row[t] = 0; p.tap_xs[t] = margin + (t % 3 - 1) * L.dil; }
    }
  }
  // one pipeline stage = 16 input channels: activation boxes + the weight blocks of all taps
  p.nchunks = (L.cin_pad / 16) * (split ? 3 : 1);      // bf16x3: chunks (A_hi,W_hi), (A_hi,W_lo), (A_lo,W_hi) per 16 channels
  p.split = split ? 1 : 0;
  p.a_box_stride = (uint32_t)(((size_t)p.BR * 2 * p.PW * 16 + 127) / 128 * 128);
  p.a_bytes = (uint32_t)(((size_t)p.nbox * p.a_box_stride + 127) / 128 * 128);
  p.a_tx_bytes = (uint32_t)((size_t)p.nbox * p.BR * 2 * p.PW * 16);
  p.b_bytes = (uint32_t)((size_t)p.ntaps * N * 32);
  p.stage_stride = (uint32_t)(((size_t)p.a_bytes + p.b_bytes + 1023) / 1024 * 1024);
  const size_t budget = 216 * 1024;
  p.S = (int)std::min<size_t>(8, budget / p.stage_stride);
  ADP_REQUIRE(p.S >= 2, "conv stage does not fit shared memory twice");
  p.idesc = ptx::make_idesc(128, N, 1);
  p.relu = 1;
  for (int v = 0; v < p.nvar; ++v) p.var[v].wbase = v * p.nchunks;
  // ky-stacked issue (conv_tc.cuh): dilation 1, at least two vertical taps fit GEMM-N <= 256
  const int KY = L.up ? 2 : 3;
  p.kys = (allow_kys && L.dil == 1 && T >= 2 && std::min(KY, T) * N <= 256) ? 1 : 0;
  if (p.kys) {
    for (int kx = 0; kx < KY; ++kx) p.tap_xs[kx] = L.up ? 7 + kx : p.margin8 * 8 + kx - 1;
    for (int k = 0; k < 3; ++k) p.idesc_stack[k] = ((k + 1) * N <= 256) ? ptx::make_idesc(128, (k + 1) * N, 1) : 0u;
  }
}

typedef void (*ConvTcKernel)(const CUtensorMap, const CUtensorMap, const ConvTcParams);
// every (taps, rows per item, ky-stacked, epilogue) combination the plans can ask for; nullptr = not instantiated
ConvTcKernel tc_kernel_lookup(int ntaps, int T, int kys, int epi) {
#define ADP_TC(NT, TT, KS, EP) if (ntaps == NT && T == TT && kys == KS && epi == EP) return conv_tc_kernel<NT, TT, (KS != 0), EP>;
#define ADP_TC_ROWS(NT, EP) ADP_TC(NT, 4, 0, EP) ADP_TC(NT, 4, 1, EP) ADP_TC(NT, 2, 0, EP) ADP_TC(NT, 2, 1, EP)
  ADP_TC_ROWS(9, EPI_STORE) ADP_TC(9, 1, 0, EPI_STORE)
  ADP_TC_ROWS(4, EPI_STORE) ADP_TC(4, 1, 0, EPI_STORE)
  ADP_TC_ROWS(9, EPI_HEAD)
  ADP_TC_ROWS(9, EPI_POOL)
  ADP_TC_ROWS(9, EPI_BWD) ADP_TC(9, 1, 0, EPI_BWD)
  ADP_TC(9, 2, 0, EPI_UPSUM) ADP_TC(9, 2, 1, EPI_UPSUM)
#undef ADP_TC_ROWS
#undef ADP_TC
  return nullptr;
}
ConvTcKernel tc_kernel_for(int ntaps, int T, int kys, int epi) {
  ConvTcKernel k = tc_kernel_lookup(ntaps, T, kys, epi);
  if (!k) throw Error(ADP_EINVAL, "no tcgen05 conv instantiation for this (taps, rows, stacking, epilogue) combination");
  return k;
}

// Effective padded fp32 weights wp[9][cin_pad][cout_pad] (concat layers: skip channels land in
// [0, pad16(skip)), upsampled-path channels in [pad16(skip), ...)) — Appendix A of SURVEY.md.
std::vector<float> padded_weights(const ConvLayer &L, const HostWeight &h, bool round_bf16) {
  std::vector<float> wp((size_t)9 * L.cin_pad * L.cout_pad, 0.f);
  const int sp = L.skip ? pad16(L.skip) : 0;
  for (int t = 0; t < 9; ++t)
    for (int ci = 0; ci < L.cin; ++ci) {
      const int cip = (L.skip && ci >= L.skip) ? sp + (ci - L.skip) : ci;
      for (int co = 0; co < L.cout; ++co) {
        float v = h.k[((size_t)t * L.cin + ci) * L.cout + co];
        wp[((size_t)t * L.cin_pad + cip) * L.cout_pad + co] = round_bf16 ? bf16_round(v) : v;
      }
    }
  return wp;
}

void pack_layer(adp_engine *e, ConvLayer &L) {
  const HostWeight &h = e->hw.at(L.name);
  const bool bf = e->prec != ADP_PREC_FP32;
  // SIMT weights + bias (bias stays fp32 on every path)
  std::vector<float> wp = padded_weights(L, h, bf);
  L.w_simt.ensure(wp.size() * 4);
  ADP_CUDA(cudaMemcpy(L.w_simt.p, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> bp(L.cout_pad, 0.f);
  std::copy(h.b.begin(), h.b.end(), bp.begin());
  L.bias.ensure(bp.size() * 4);
  ADP_CUDA(cudaMemcpy(L.bias.p, bp.data(), bp.size() * 4, cudaMemcpyHostToDevice));
  if (e->prec != ADP_PREC_BF16) return;

  plan_tc(L, e->kys, e->split);
  const ConvTcParams &p = L.tc;
  std::vector<float> w32 = padded_weights(L, h, false);   // sum taps in fp32, round once
  auto W = [&](int t, int ci, int co) -> float { return w32[((size_t)t * L.cin_pad + ci) * L.cout_pad + co]; };
  const int KC = 16, N = p.N;
  const size_t blk = (size_t)KC * N;      // one (chunk, tap) block: [2 planes][N][8]
  std::vector<__nv_bfloat16> pk((size_t)p.nvar * p.nchunks * p.ntaps * blk);
  for (int v = 0; v < p.nvar; ++v)
    for (int c = 0; c < p.nchunks; ++c)
      for (int t = 0; t < p.ntaps; ++t) {
        __nv_bfloat16 *base = pk.data() + ((size_t)p.var[v].wbase + c) * p.ntaps * blk;
        for (int g = 0; g < 2; ++g)
          for (int n = 0; n < N; ++n)
            for (int j = 0; j < 8; ++j) {
              const int creal = p.split ? c / 3 : c, pass = p.split ? c % 3 : 0;
              const int ci = creal * KC + g * 8 + j;
              float val = 0.f;
              if (ci < L.cin_pad) {
                if (L.up) {
                  // parity (py,px), tap (ry,rx) in {0,1}^2: sum of the 3x3 taps that read this source pixel
                  const int py = v >> 1, px = v & 1, ry = t >> 1, rx = t & 1;
                  for (int ky = 0; ky < 3; ++ky) {
                    const int sy = (py + ky - 1) >= 0 ? (py + ky - 1) / 2 : -1;   // floor((py+ky-1)/2)
                    if (sy - (py - 1) != ry) continue;
                    for (int kx = 0; kx < 3; ++kx) {
                      const int sx = (px + kx - 1) >= 0 ? (px + kx - 1) / 2 : -1;
                      if (sx - (px - 1) != rx) continue;
                      val += W(ky * 3 + kx, ci, n);
                    }
                  }
                } else {
                  val = W(t, ci, p.var[v].out_cg * 8 + n);
                }
              }
              __nv_bfloat16 hv = __float2bfloat16_rn(val);
              if (p.split && pass == 1) hv = __float2bfloat16_rn(val - __bfloat162float(hv));      // W_lo block
              base[tc_block_index(p.kys, p.ntaps, N, t, g, n, j)] = hv;
            }
      }
  L.w_tc.ensure(pk.size() * 2);
  ADP_CUDA(cudaMemcpy(L.w_tc.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
  L.tc_ready = true;
}

void pack_all(adp_engine *e) {
  for (const char *n : kAllNames)
    if (!e->hw.count(n) || !e->hw[n].set) throw Error(ADP_ESTATE, std::string("weights of layer ") + n + " not set");
  const bool bf = e->prec != ADP_PREC_FP32;
  // first conv: [9][cp0]; never rounded (computed in fp32 from the fp32 image on every path)
  {
    const HostWeight &h = e->hw["down1_conv1"];
    std::vector<float> w((size_t)9 * e->cp[0], 0.f), b(e->cp[0], 0.f);
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < e->c[0]; ++co) w[(size_t)t * e->cp[0] + co] = h.k[(size_t)t * e->c[0] + co];
    std::copy(h.b.begin(), h.b.end(), b.begin());
    e->w_first.ensure(w.size() * 4); e->b_first.ensure(b.size() * 4);
    ADP_CUDA(cudaMemcpy(e->w_first.p, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    ADP_CUDA(cudaMemcpy(e->b_first.p, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    memset(&e->fc_host, 0, sizeof(e->fc_host));
    if (e->cp[0] <= 64) {
      for (int t = 0; t < 9; ++t)
        for (int co = 0; co < e->cp[0]; ++co) e->fc_host.w[t][co] = w[(size_t)t * e->cp[0] + co];
      for (int co = 0; co < e->cp[0]; ++co) e->fc_host.b[co] = b[co];
    }
  }
  {
    const HostWeight &h = e->hw["output_softmax"];   // (1,1,c1,2)
    std::vector<float> w((size_t)2 * e->cp[0], 0.f);
    for (int ci = 0; ci < e->c[0]; ++ci)
      for (int k = 0; k < 2; ++k) w[(size_t)k * e->cp[0] + ci] = h.k[(size_t)ci * 2 + k];
    e->w_head.ensure(w.size() * 4); e->b_head.ensure(8);
    ADP_CUDA(cudaMemcpy(e->w_head.p, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    ADP_CUDA(cudaMemcpy(e->b_head.p, h.b.data(), 8, cudaMemcpyHostToDevice));
  }
  (void)bf;
  for (auto &L : e->layers) pack_layer(e, L);
  e->packed = true;
}

// ------------------------------------------------------------------------------------------------
// activation arena
void ensure_arena(adp_engine *e, int S) {
  if (e->S == S) return;
  ADP_REQUIRE(S >= 8 && S % 8 == 0 && S <= 8192, "tile size must be a multiple of 8");
  const size_t es = e->esz * e->mul(), nf = e->max_fw;      // bf16x3: hi + lo halves
  const size_t s1 = (size_t)S * S, s2 = s1 / 4, s3 = s1 / 16, s4 = s1 / 64;
  const int *cp = e->cp;
  e->a1.ensure(nf * s1 * cp[0] * es);   e->b1.ensure(nf * s1 * cp[0] * es);   e->cat1.ensure(nf * s1 * 2 * cp[0] * es);
  e->pl1.ensure(nf * s2 * cp[0] * es);
  e->a2.ensure(nf * s2 * cp[1] * es);   e->b2.ensure(nf * s2 * cp[1] * es);   e->cat2.ensure(nf * s2 * 2 * cp[1] * es);
  e->pl2.ensure(nf * s3 * cp[1] * es);
  e->a3.ensure(nf * s3 * cp[2] * es);   e->b3.ensure(nf * s3 * cp[2] * es);   e->cat3.ensure(nf * s3 * 2 * cp[2] * es);
  e->pl3.ensure(nf * s4 * cp[2] * es);
  for (int i = 0; i < 6; ++i) e->t[i].ensure(nf * s4 * cp[3] * es);
  e->ts.ensure(nf * s4 * cp[3] * es);
  e->prob.ensure(nf * s1 * 4);
  e->fwt_tile.ensure(64 * 4); e->fwt_op.ensure(64 * 4); e->fwt_origin.ensure(64 * 8);
  e->tmaps.clear();
  e->S = S;
  Acts &a = e->acts;
  a.d1a = &e->a1; a.cat1 = &e->cat1; a.u1b = &e->a1; a.u1c = &e->b1; a.pl1 = &e->pl1;
  a.d2a = &e->a2; a.cat2 = &e->cat2; a.u2b = &e->a2; a.u2c = &e->b2; a.pl2 = &e->pl2;
  a.d3a = &e->a3; a.cat3 = &e->cat3; a.u3b = &e->a3; a.u3c = &e->b3; a.pl3 = &e->pl3;
  for (int i = 0; i < 6; ++i) a.t[i] = &e->t[i];
  a.ts = &e->ts; a.prob = &e->prob; a.cap = e->max_fw;
}

// pitch / coff / C in channels (multiples of 8); the buffer is row-planar (kernels_simt.cuh)
template <typename T> View<T> view(const DevBuf &b, int H, int W, int pitch, int coff, int C) {
  View<T> v;
  v.p = b.as<T>(); v.H = H; v.W = W; v.cgs = pitch / 8; v.cg0 = coff / 8; v.C = C; v.lo = 0;
  return v;
}
// bf16x3: the buffer holds hi groups [0, pitch/8) and lo groups [pitch/8, 2*pitch/8) per pixel row
template <typename T> View<T> view_split(const DevBuf &b, int H, int W, int pitch, int coff, int C) {
  View<T> v;
  v.p = b.as<T>(); v.H = H; v.W = W; v.cgs = 2 * pitch / 8; v.cg0 = coff / 8; v.C = C; v.lo = pitch / 8;
  return v;
}

// Tensor map over a row-planar view [n][y][cg][x][8] (bf16): dim0 = 8 pixels x 8 channels (one 128-byte line),
// dim1 = 8-pixel groups along the row, dim2 = channel group, dim3 = row, dim4 = image.  Box = (64, box_px8,
// box_cg, box_rows, 1); everything outside the view (borders, channel tail) is zero-filled by the TMA unit.
const CUtensorMap &tmap_for(adp_engine *e, const void *buf, int H, int W, int cgs, int cg0, int C, int cap, int box_px8,
                            int box_cg, int box_rows) {
  char key[160];
  snprintf(key, sizeof(key), "%p/%d/%d/%d/%d/%d/%d/%d/%d/%d", buf, H, W, cgs, cg0, C, cap, box_px8, box_cg, box_rows);
  auto it = e->tmaps.find(key);
  if (it != e->tmaps.end()) return it->second;
  CUtensorMap m;
  ADP_REQUIRE(W % 8 == 0, "tcgen05 conv path needs every level's width to be a multiple of 8 (tile size % 64 == 0)");
  cuuint64_t gdim[5] = {64, (cuuint64_t)(W / 8), (cuuint64_t)(C / 8), (cuuint64_t)H, (cuuint64_t)cap};
  cuuint64_t gstr[4] = {128, (cuuint64_t)W * 16, (cuuint64_t)cgs * W * 16, (cuuint64_t)H * cgs * W * 16};
  cuuint32_t box[5] = {64, (cuuint32_t)box_px8, (cuuint32_t)box_cg, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  void *base = (void *)(reinterpret_cast<const __nv_bfloat16 *>(buf) + (size_t)cg0 * W * 8);
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(ADP_ECUDA, std::string("cuTensorMapEncodeTiled failed, code ") + std::to_string((int)r) + " key " + key);
  return e->tmaps.emplace(key, m).first->second;
}

// Tensor map over the normalised float32 input [forward][H][W] of the fused first conv: box = 8 rows x 136 columns, zero fill
// outside the image ('same' padding of the normalised image)
const CUtensorMap &tmap_input(adp_engine *e, const float *buf, int H, int W, int cap) {
  char key[160];
  snprintf(key, sizeof(key), "in/%p/%d/%d/%d", (const void *)buf, H, W, cap);
  auto it = e->tmaps.find(key);
  if (it != e->tmaps.end()) return it->second;
  CUtensorMap m;
  ADP_REQUIRE(W % 4 == 0, "fused first conv needs a row pitch of 16 bytes");
  cuuint64_t gdim[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cap, 1, 1};
  cuuint64_t gstr[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)cap * H * W * 4, (cuuint64_t)cap * H * W * 4};
  cuuint32_t box[5] = {(cuuint32_t)kFcWinW, (cuuint32_t)kFcWinRows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void *)buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(ADP_ECUDA, std::string("cuTensorMapEncodeTiled failed, code ") + std::to_string((int)r) + " key " + key);
  return e->tmaps.emplace(key, m).first->second;
}

// TTA combine / blend kernel (kernels_post.cuh); ADP_TTA_SERIAL=1 selects the one-plane-per-barrier variant for A/B timing
typedef void (*TtaKernel)(const float *, TtaOps, int, int, float *, float *, float *, const float *, int, int, int, int, int);
TtaKernel tta_kernel() {
  static const bool serial = getenv("ADP_TTA_SERIAL") && atoi(getenv("ADP_TTA_SERIAL")) != 0;
  return serial ? tta_blend_serial_kernel : tta_blend_kernel;
}

double conv_flops(const ConvLayer &L, int Hout, int Wout, int nb) {
  return 2.0 * nb * Hout * Wout * 9.0 * L.cin * L.cout;
}

// one generic 3x3 conv layer: src view (H,W are the SOURCE buffer dims) -> dst view
struct EpiSpec {
  int mode = EPI_STORE;
  const DevBuf *pool_dst = nullptr;   // EPI_POOL: dense pooled tensor (pitch = channels of the layer)
  float *prob = nullptr;              // EPI_HEAD: probability planes
  const void *resid = nullptr, *mask = nullptr;   // backward (EPI_STORE): see ConvTcParams
  float mask_scale = 1.f;
  void *up_dst = nullptr;             // EPI_UPSUM: low-resolution gradient (view: up_cgs groups per row, first group up_cg0);
  int up_cgs = 0, up_cg0 = 0;         // resid / mask above then have ITS layout
  float drop_keep = 1.f;              // EPI_STORE, training forward: hash dropout of the stored tensor fused into the epilogue
  uint64_t drop_seed = 0;             // (keep < 1: on; the output must be a dense tensor whose pair index fits 32 bits)
  const FirstConvFuse *fc = nullptr;  // EPI_POOL of down1_conv2: compute the source tensor (first conv) in the kernel from
  const float *fc_input = nullptr;    // the normalised float32 image [forward][Hs][Ws] (tta_input_kernel)
};

// tcgen05 conv launch shared by the forward layers and their data-gradient twins (train_host.cuh)
void launch_conv_tc(adp_engine *e, const ConvLayer &L, const std::string &label, double fl, double by, const void *src, int Hs,
                    int Ws, int s_cgs, int s_cg0, void *dst, int d_cgs, int d_cg0, int nb, int cap, const EpiSpec &epi,
                    const float *bias, int relu) {
  ConvTcParams p = L.tc;
  // bf16x3: s_cgs / d_cgs arrive as LOGICAL groups per row; the buffers hold 2x that, lo halves start one logical pitch later
  p.in_lo = 0; p.out_lo = 0; p.pool_lo = 0;
  if (p.split) {
    ADP_REQUIRE(!epi.mask && !epi.resid, "bf16x3 precision is inference-only");
    p.in_lo = s_cgs; p.out_lo = d_cgs;
    s_cgs *= 2; d_cgs *= 2;
  }
  const int Ho = L.up ? Hs * 2 : Hs, Wo = L.up ? Ws * 2 : Ws;
  p.nb = nb; p.Hin = Hs; p.Win = Ws;
  p.ntx = cdiv(Ws, 128); p.nty = cdiv(Hs, p.T);
  p.wpk = L.w_tc.as<__nv_bfloat16>(); p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16 *>(dst); p.out_cgs = d_cgs; p.out_cg0 = d_cg0; p.Hout = Ho; p.Wout = Wo;
  p.dbg = e->dbg;
  p.relu = relu;
  p.epi_mode = epi.mode;
  p.resid = reinterpret_cast<const __nv_bfloat16 *>(epi.resid);
  p.mask = reinterpret_cast<const __nv_bfloat16 *>(epi.mask);
  p.mask_scale = epi.mask_scale;
  p.drop_thr = 0; p.drop_salt = 0; p.drop_inv = 1.f;
  if (epi.drop_keep < 1.f) {
    ADP_REQUIRE(epi.mode == EPI_STORE && !epi.mask && !p.split && d_cg0 == 0 && d_cgs * 8 == L.cout_pad &&
                    (size_t)nb * Ho * Wo * d_cgs * 4 <= 0xFFFFFFFFull, "fused dropout needs a dense bf16 output tensor");
    p.drop_thr = (uint32_t)fminf(epi.drop_keep * 65536.f, 65536.f);
    p.drop_inv = 1.f / epi.drop_keep;
    p.drop_salt = dropout_salt(epi.drop_seed, 0u);
  }
  if (epi.mode == EPI_HEAD) {
    ADP_REQUIRE(p.T >= 2 && p.nvar == 1 && p.N == e->cp[0], "head fusion needs the 44-channel full-resolution layer");
    p.head_w = e->w_head.as<float>(); p.head_b = e->b_head.as<float>(); p.prob = epi.prob;
  } else if (epi.mode == EPI_UPSUM) {
    ADP_REQUIRE(p.T == 2 && p.oscale == 1 && Ho % 2 == 0 && Wo % 2 == 0 && !p.split, "2x2-sum epilogue needs a two-row item");
    p.pool_out = reinterpret_cast<__nv_bfloat16 *>(epi.up_dst); p.pool_cgs = epi.up_cgs; p.pool_cg0 = epi.up_cg0;
  } else if (epi.mode == EPI_POOL) {
    ADP_REQUIRE(p.T % 2 == 0 && p.oscale == 1 && Ho % 2 == 0 && Wo % 2 == 0, "pool fusion needs an even row block");
    p.pool_out = epi.pool_dst->as<__nv_bfloat16>(); p.pool_cgs = L.cout_pad / 8; p.pool_cg0 = 0;
    if (p.split) { p.pool_lo = p.pool_cgs; p.pool_cgs *= 2; }
  }
  // bf16x3: the map spans hi and lo groups of the view (groups in between belong to other tensors of a concat buffer)
  const int map_C = p.split ? (p.in_lo + L.cin_pad / 8) * 8 : L.cin_pad;
  const CUtensorMap &tm = tmap_for(e, src, Hs, Ws, s_cgs, s_cg0, map_C, cap, p.PW / 8, 2, p.BR);
  const int nitems = nb * p.nty * p.ntx * p.nvar;
  const int grid = std::min(nitems, e->num_sms);
  const int epi_kind = (epi.mode == EPI_STORE && epi.mask) ? EPI_BWD : epi.mode;
  ConvTcKernel kern = tc_kernel_for(p.ntaps, p.T, p.kys, epi_kind);
  int threads = kTcThreads;
  size_t smem_extra = 0;
  if (epi.fc) {
    ADP_REQUIRE(epi.mode == EPI_POOL && p.ntaps == 9 && p.T == 4 && p.kys == 1 && p.nvar == 1 && !p.split && p.PW == 144 && p.nbox == 1 &&
                    L.cin_pad <= 16 * kFcMaxChunks, "first-conv fusion needs the 3x3, 4-row, ky-stacked pool layer");
    p.fc = *epi.fc;
    kern = conv_tc_kernel<9, 4, true, EPI_POOL, true>;
    threads = kTcThreadsFc;
    smem_extra = kFcWinBytes;
    while (p.S > 2 && tc_smem_bytes(p) + smem_extra > 227 * 1024) --p.S;
  }
  // data-gradient twin, N <= 96 (T >= 2 rows per item): the mask tile of every item is staged in shared memory next to the
  // pipeline stages (two buffers when at least two stages still fit, else one).  Wider accumulators keep the
  // register-prefetch loads (mask_bufs = 0): a 44 KB tile per buffer would leave them two 63 KB stages, measured slower
  // (up2_conv2 twin 0.74 -> 0.98 ms) while the 48/96-wide twins gain (up1_conv2 1.35 -> 0.85 ms, down1_conv2 0.71 -> 0.43 ms).
  // ADP_BWD_MASK = 0 (never) / 1 / 2 (buffers where staged) / 3 (stage for every width) overrides for experiments.
  p.mask_bufs = 0; p.mask_bytes = 0;
  const CUtensorMap *tmk = &tm;
  if (epi.fc) tmk = &tmap_input(e, epi.fc_input, Hs, Ws, cap);
  if (epi_kind == EPI_BWD) {
    static const int mode = getenv("ADP_BWD_MASK") ? atoi(getenv("ADP_BWD_MASK")) : -1;
    const bool stage = p.oscale == 1 && mode != 0 && (p.N <= 96 || mode == 3);
    if (stage) {
      p.mask_bytes = (uint32_t)(p.T * p.N * 256);
      const size_t budget = 220 * 1024;
      p.mask_bufs = (2 * (size_t)p.mask_bytes + 2 * (size_t)p.stage_stride <= budget) ? 2 : 1;
      if (mode == 1 || mode == 2) p.mask_bufs = mode;
      const size_t left = budget - (size_t)p.mask_bufs * p.mask_bytes;
      p.S = (int)std::min<size_t>(p.S, left / p.stage_stride);
      ADP_REQUIRE(p.S >= 2, "data-gradient twin: pipeline stages and the mask tile do not fit shared memory");
      tmk = &tmap_for(e, epi.mask, Ho, Wo, d_cgs, d_cg0, L.cout_pad, cap, 16, p.N / 8, p.T);
    }
  }
  // resident weights: all (variant, chunk) blocks once per CTA instead of once per item and chunk, stages carry activations only
  p.bres = 0; p.bres_off = 0; p.bres_bytes = 0;
  if (e->resident_weights && !epi.fc) {
    const bool own_variant = p.nvar > 1 && grid % p.nvar == 0;      // a CTA then only meets variant blockIdx.x % nvar
    const size_t wb = (size_t)(own_variant ? 1 : p.nvar) * p.nchunks * p.b_bytes;
    const uint32_t a_stride = (uint32_t)(((size_t)p.a_bytes + 1023) / 1024 * 1024);
    const size_t fixed = (size_t)p.mask_bufs * p.mask_bytes + (2 * 8 + 10) * 8 + (704 + 520) * 4 + 256;
    const size_t total = 226 * 1024;
    static const int min_stages = getenv("ADP_BRES_MIN_STAGES") ? atoi(getenv("ADP_BRES_MIN_STAGES")) : 4;
    if (wb + fixed + (size_t)min_stages * a_stride <= total && wb < (1u << 20)) {
      p.bres = own_variant ? 2 : 1; p.bres_bytes = (uint32_t)wb;
      p.stage_stride = a_stride;
      p.S = (int)std::min<size_t>(8, (total - wb - fixed) / a_stride);
      p.bres_off = tc_bres_offset(p);
    }
  }
  const size_t smem = tc_smem_bytes(p) + smem_extra;
  ADP_REQUIRE(smem <= 227 * 1024, "tcgen05 conv: shared-memory plan exceeds 227 KB");
  if (e->dbg & 16) { e->misc.ensure(148 * 8 * 8 + 4096); p.dbg_out = reinterpret_cast<long long *>(e->misc.as<char>() + 4096); }
  e->launch(label.c_str(), fl, by, [&] { kern<<<grid, threads, smem, e->stream>>>(tm, *tmk, p); });
  if (e->dbg & 16) {
    // role timers (cycles, averaged over CTAs): where each warp role of the pipeline waits
    std::vector<long long> h((size_t)grid * 8);
    ADP_CUDA(cudaStreamSynchronize(e->stream));
    ADP_CUDA(cudaMemcpy(h.data(), p.dbg_out, h.size() * 8, cudaMemcpyDeviceToHost));
    double a[8] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)b * 8 + k] / grid;
    const double items = (double)nitems / grid;
    fprintf(stderr, "tc-timers %-34s items/CTA %6.1f chunks %2d | producer wait-empty %5.1f%% of %8.0f | mma wait-acc-empty %5.1f%% wait-full %5.1f%% commit %5.1f%% of %8.0f | epilogue wait-acc-full %5.1f%% of %8.0f | cycles/item %6.0f\n",
            label.c_str(), items, p.nchunks, 100 * a[0] / a[1], a[1], 100 * a[2] / a[4], 100 * a[3] / a[4], 100 * a[7] / a[4], a[4], 100 * a[5] / a[6], a[6], a[4] / items);
  }
}

void run_conv(adp_engine *e, const std::string &name, const DevBuf &src, int Hs, int Ws, int spitch, int scoff,
              const DevBuf &dst, int dpitch, int dcoff, int nb, int cap, EpiSpec epi = EpiSpec()) {
  ConvLayer &L = layer(e, name);
  const int Ho = L.up ? Hs * 2 : Hs, Wo = L.up ? Ws * 2 : Ws;
  const double fl = conv_flops(L, Ho, Wo, nb);
  const double by = (double)nb * ((double)Hs * Ws * L.cin_pad + (double)Ho * Wo * L.cout_pad) * e->esz;
  if (e->prec == ADP_PREC_BF16) {
    launch_conv_tc(e, L, "conv3x3_tcgen05/" + name, fl, by, src.p, Hs, Ws, spitch / 8, scoff / 8, dst.p, dpitch / 8, dcoff / 8, nb, cap,
                   epi, L.bias.as<float>(), 1);
    return;
  }
  dim3 grid(cdiv(Wo, 32), cdiv(Ho, 16), nb * (L.cout_pad / 16));   // block (32, 8) covers 32 x 16 pixels
  dim3 block(32, 8);
  if (e->prec == ADP_PREC_FP32) {
    auto in = view<float>(src, Hs, Ws, spitch, scoff, L.cin_pad);
    auto out = view<float>(dst, Ho, Wo, dpitch, dcoff, L.cout_pad);
    e->launch(("conv3x3_simt_fp32/" + name).c_str(), fl, by, [&] {
      if (L.up) conv3x3_simt_kernel<float, true><<<grid, block, 0, e->stream>>>(in, out, L.w_simt.as<float>(), L.bias.as<float>(), L.dil, 1);
      else conv3x3_simt_kernel<float, false><<<grid, block, 0, e->stream>>>(in, out, L.w_simt.as<float>(), L.bias.as<float>(), L.dil, 1);
    });
  } else {
    auto in = view<__nv_bfloat16>(src, Hs, Ws, spitch, scoff, L.cin_pad);
    auto out = view<__nv_bfloat16>(dst, Ho, Wo, dpitch, dcoff, L.cout_pad);
    e->launch(("conv3x3_simt_bf16/" + name).c_str(), fl, by, [&] {
      if (L.up) conv3x3_simt_kernel<__nv_bfloat16, true><<<grid, block, 0, e->stream>>>(in, out, L.w_simt.as<float>(), L.bias.as<float>(), L.dil, 1);
      else conv3x3_simt_kernel<__nv_bfloat16, false><<<grid, block, 0, e->stream>>>(in, out, L.w_simt.as<float>(), L.bias.as<float>(), L.dil, 1);
    });
  }
}

template <typename T>
void run_pool(adp_engine *e, const DevBuf &src, int Hs, int Ws, int spitch, int C, const DevBuf &dst, int nb) {
  auto in = e->split ? view_split<T>(src, Hs, Ws, spitch, 0, C) : view<T>(src, Hs, Ws, spitch, 0, C);
  auto out = e->split ? view_split<T>(dst, Hs / 2, Ws / 2, C, 0, C) : view<T>(dst, Hs / 2, Ws / 2, C, 0, C);
  const size_t total = (size_t)nb * (Hs / 2) * (Ws / 2) * (C / 8);
  const int grid = e->wave_grid(maxpool2_kernel<T>, cdiv64(total, 256));
  e->launch("maxpool2x2", 0, (double)nb * Hs * Ws * C * sizeof(T) * 1.25, [&] {
    maxpool2_kernel<T><<<grid, 256, 0, e->stream>>>(in, out, nb);
  });
}

// Dropout sites of the training graph (train_adipose_unet_v3.py:682,696,703,710); inference passes null.
struct DropSpec {
  float keep = 1.f;
  uint64_t seed = 0;
  const uint8_t *mask[4] = {nullptr, nullptr, nullptr, nullptr};   // supplied masks (NHWC, real channels) or null
};

template <typename T>
void run_dropout(adp_engine *e, const DevBuf &buf, int H, int C, int creal, int nb, const DropSpec &d, int site) {
  auto v = view<T>(buf, H, H, C, 0, C);
  const size_t total = (size_t)nb * H * H * (C / 8);
  const uint64_t seed = d.seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(site + 1);
  if (!d.mask[site] && total * 4 <= 0xFFFFFFFFull) {      // dense view (pitch == C, offset 0): item index == element offset / 8
    const int grid = e->wave_grid(dropout_dense_kernel<T>, cdiv64(total, 256));
    e->launch("dropout", 0, (double)total * 16 * sizeof(T), [&] {
      dropout_dense_kernel<T><<<grid, 256, 0, e->stream>>>(v.p, (uint32_t)total, d.keep, seed);
    });
    return;
  }
  const int grid = e->wave_grid(dropout_kernel<T>, cdiv64(total, 256));
  e->launch("dropout", 0, (double)total * 16 * sizeof(T), [&] {
    dropout_kernel<T><<<grid, 256, 0, e->stream>>>(v, nb, d.keep, seed, d.mask[site], creal);
  });
}

template <typename T> void forward_t(adp_engine *e, const Acts &A, int S, const FirstConvSrc &src, const FwTable &fw, int nfw,
                                     float mean, float std_, cudaEvent_t input_consumed, const DropSpec *drop) {
  const int S2 = S / 2, S3 = S / 4, S4 = S / 8;
  const int *cp = e->cp, *c = e->c;
  const int cap = A.cap;
  // forward table -> device (kernel parameter copy; no host buffer lifetime to worry about)
  e->launch("fw_table", 0, 0, [&] {
    fw_table_store<<<1, 64, 0, e->stream>>>(fw, e->fwt_tile.as<int>(), e->fwt_op.as<int>(), e->fwt_origin.as<long long>(), nfw);
  });
  FirstConvSrc s = src;
  s.slide_origin = reinterpret_cast<const int64_t *>(e->fwt_origin.p);
  const float mean_f = mean;
  const float sd_f = (float)((double)std_ + 1e-10);
  const bool tc = e->prec == ADP_PREC_BF16;
  const bool dropping = drop && (drop->keep < 1.f || drop->mask[0]);
  // training forward on the tcgen05 path: the four Dropout sites run inside the epilogue of the conv that produces the tensor
  // (same hash, same roundings as dropout_dense_kernel: bit-identical) unless parity tests supply explicit masks
  auto drop_fused = [&](int site, int H, int C) {
    return dropping && tc && e->fuse_dropout && !drop->mask[site] && drop->keep < 1.f && (size_t)nfw * H * H * (C / 8) * 4 <= 0xFFFFFFFFull;
  };
  auto drop_epi = [&](int site) {
    EpiSpec ep;
    ep.drop_keep = drop->keep;
    ep.drop_seed = drop->seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(site + 1);
    return ep;
  };
  EpiSpec pool1, pool2, head;
  if (tc && (e->fuse_pool || e->split)) { pool1.mode = EPI_POOL; pool1.pool_dst = A.pl1; pool2.mode = EPI_POOL; pool2.pool_dst = A.pl2; }
  if (tc && (e->fuse_head || e->split) && !drop) { head.mode = EPI_HEAD; head.prob = A.prob->as<float>(); }   // training keeps up1_conv3
  // Inference on the tcgen05 path: the first conv runs inside down1_conv2 (its 48-channel output is never materialised);
  // training needs that tensor for the weight gradient, bf16x3 carries hi/lo halves, both keep the separate kernel.
  const bool fuse_first = tc && e->fuse_first && !drop && !e->split && pool1.mode == EPI_POOL && cp[0] <= 16 * kFcMaxChunks && layer(e, "down1_conv2").tc.kys == 1 &&
                          layer(e, "down1_conv2").tc.T == 4;
  if (fuse_first) {
    // the d1a buffer is free in this mode (until up1_conv2 writes it): it holds the normalised float32 input instead
    float *xin = A.d1a->as<float>();
    const size_t total = (size_t)nfw * S * (S / 4);
    const int grid = e->wave_grid(tta_input_kernel, cdiv64(total, 256));
    e->launch("tta_input", 0, (double)nfw * S * S * 8, [&] {
      tta_input_kernel<<<grid, 256, 0, e->stream>>>(s, e->fwt_tile.as<int>(), e->fwt_op.as<int>(), S, mean_f, sd_f, xin, nfw);
    });
    if (input_consumed) ADP_CUDA(cudaEventRecord(input_consumed, e->stream));
    pool1.fc = &e->fc_host; pool1.fc_input = xin;
  } else {
    auto out = e->split ? view_split<T>(*A.d1a, S, S, cp[0], 0, cp[0]) : view<T>(*A.d1a, S, S, cp[0], 0, cp[0]);
    dim3 grid(cdiv(S, 32), cdiv(S, 32), nfw), block(32, 8);
    const size_t smem = ((size_t)10 * cp[0] + 34 * 34) * 4;
    e->launch("first_conv", 2.0 * nfw * S * S * 9.0 * e->c[0], (double)nfw * S * S * (4 + cp[0] * sizeof(T)), [&] {
      first_conv_kernel<T><<<grid, block, smem, e->stream>>>(s, e->fwt_tile.as<int>(), e->fwt_op.as<int>(), S, mean_f, sd_f,
                                                            e->w_first.as<float>(), e->b_first.as<float>(), out);
    });
    if (input_consumed) ADP_CUDA(cudaEventRecord(input_consumed, e->stream));
  }
  run_conv(e, "down1_conv2", *A.d1a, S, S, cp[0], 0, *A.cat1, 2 * cp[0], 0, nfw, cap, pool1);
  if (pool1.mode != EPI_POOL) run_pool<T>(e, *A.cat1, S, S, 2 * cp[0], cp[0], *A.pl1, nfw);
  run_conv(e, "down2_conv1", *A.pl1, S2, S2, cp[0], 0, *A.d2a, cp[1], 0, nfw, cap);
  run_conv(e, "down2_conv2", *A.d2a, S2, S2, cp[1], 0, *A.cat2, 2 * cp[1], 0, nfw, cap, pool2);
  if (pool2.mode != EPI_POOL) run_pool<T>(e, *A.cat2, S2, S2, 2 * cp[1], cp[1], *A.pl2, nfw);
  run_conv(e, "down3_conv1", *A.pl2, S3, S3, cp[1], 0, *A.d3a, cp[2], 0, nfw, cap);
  run_conv(e, "down3_conv2", *A.d3a, S3, S3, cp[2], 0, *A.cat3, 2 * cp[2], 0, nfw, cap);
  run_pool<T>(e, *A.cat3, S3, S3, 2 * cp[2], cp[2], *A.pl3, nfw);
  if (drop_fused(0, S4, cp[3])) run_conv(e, "dilate1", *A.pl3, S4, S4, cp[2], 0, *A.t[0], cp[3], 0, nfw, cap, drop_epi(0));
  else {
    run_conv(e, "dilate1", *A.pl3, S4, S4, cp[2], 0, *A.t[0], cp[3], 0, nfw, cap);
    if (dropping) run_dropout<T>(e, *A.t[0], S4, cp[3], c[3], nfw, *drop, 0);
  }
  const char *dn[5] = {"dilate2", "dilate3", "dilate4", "dilate5", "dilate6"};
  for (int i = 0; i < 5; ++i) run_conv(e, dn[i], *A.t[i], S4, S4, cp[3], 0, *A.t[i + 1], cp[3], 0, nfw, cap);
  if (e->split) {
    // hi/lo tensors: the six values are recombined in fp32, summed and split again
    View<T> v[6];
    for (int i = 0; i < 6; ++i) v[i] = view_split<T>(*A.t[i], S4, S4, cp[3], 0, cp[3]);
    auto o = view_split<T>(*A.ts, S4, S4, cp[3], 0, cp[3]);
    const size_t total = (size_t)nfw * S4 * S4 * (cp[3] / 8);
    e->launch("add6", 0, (double)total * 32 * 7, [&] {
      add6_split_kernel<T><<<e->wave_grid(add6_split_kernel<T>, cdiv64(total, 256)), 256, 0, e->stream>>>(
          v[0], v[1], v[2], v[3], v[4], v[5], o, nfw);
    });
  } else {
    const size_t nvec = (size_t)nfw * S4 * S4 * cp[3] * sizeof(T) / 16;
    const int grid = e->wave_grid(add6_kernel<T>, cdiv64(nvec, 256));
    e->launch("add6", 0, (double)nvec * 16 * 7, [&] {
      add6_kernel<T><<<grid, 256, 0, e->stream>>>(A.t[0]->as<T>(), A.t[1]->as<T>(), A.t[2]->as<T>(), A.t[3]->as<T>(),
                                                 A.t[4]->as<T>(), A.t[5]->as<T>(), A.ts->as<T>(), nvec);
    });
  }
  run_conv(e, "up3_conv1", *A.ts, S4, S4, cp[3], 0, *A.cat3, 2 * cp[2], cp[2], nfw, cap);
  run_conv(e, "up3_conv2", *A.cat3, S3, S3, 2 * cp[2], 0, *A.u3b, cp[2], 0, nfw, cap);
  if (drop_fused(1, S3, cp[2])) run_conv(e, "up3_conv3", *A.u3b, S3, S3, cp[2], 0, *A.u3c, cp[2], 0, nfw, cap, drop_epi(1));
  else {
    run_conv(e, "up3_conv3", *A.u3b, S3, S3, cp[2], 0, *A.u3c, cp[2], 0, nfw, cap);
    if (dropping) run_dropout<T>(e, *A.u3c, S3, cp[2], c[2], nfw, *drop, 1);
  }
  run_conv(e, "up2_conv1", *A.u3c, S3, S3, cp[2], 0, *A.cat2, 2 * cp[1], cp[1], nfw, cap);
  run_conv(e, "up2_conv2", *A.cat2, S2, S2, 2 * cp[1], 0, *A.u2b, cp[1], 0, nfw, cap);
  if (drop_fused(2, S2, cp[1])) run_conv(e, "up2_conv3", *A.u2b, S2, S2, cp[1], 0, *A.u2c, cp[1], 0, nfw, cap, drop_epi(2));
  else {
    run_conv(e, "up2_conv3", *A.u2b, S2, S2, cp[1], 0, *A.u2c, cp[1], 0, nfw, cap);
    if (dropping) run_dropout<T>(e, *A.u2c, S2, cp[1], c[1], nfw, *drop, 2);
  }
  run_conv(e, "up1_conv1", *A.u2c, S2, S2, cp[1], 0, *A.cat1, 2 * cp[0], cp[0], nfw, cap);
  run_conv(e, "up1_conv2", *A.cat1, S, S, 2 * cp[0], 0, *A.u1b, cp[0], 0, nfw, cap);
  if (drop_fused(3, S, cp[0]) && head.mode != EPI_HEAD) run_conv(e, "up1_conv3", *A.u1b, S, S, cp[0], 0, *A.u1c, cp[0], 0, nfw, cap, drop_epi(3));
  else {
    run_conv(e, "up1_conv3", *A.u1b, S, S, cp[0], 0, *A.u1c, cp[0], 0, nfw, cap, head);
    if (dropping) run_dropout<T>(e, *A.u1c, S, cp[0], c[0], nfw, *drop, 3);
  }
  if (head.mode != EPI_HEAD) {
    auto in = view<T>(*A.u1c, S, S, cp[0], 0, cp[0]);
    const size_t total = (size_t)nfw * S * S;
    const int grid = e->wave_grid(head_kernel<T>, cdiv64(total, 256), 256, (size_t)2 * cp[0] * 4);
    e->launch("head_softmax", 2.0 * total * e->c[0] * 2, (double)total * (cp[0] * sizeof(T) + 4), [&] {
      head_kernel<T><<<grid, 256, (size_t)2 * cp[0] * 4, e->stream>>>(in, nfw, e->w_head.as<float>(), e->b_head.as<float>(),
                                                                      A.prob->as<float>());
    });
  }
  e->last_nfw = nfw;
}

void forward(adp_engine *e, const FirstConvSrc &src, const FwTable &fw, int nfw, float mean, float std_,
             cudaEvent_t input_consumed = nullptr) {
  if (e->prec == ADP_PREC_FP32) forward_t<float>(e, e->acts, e->S, src, fw, nfw, mean, std_, input_consumed, nullptr);
  else forward_t<__nv_bfloat16>(e, e->acts, e->S, src, fw, nfw, mean, std_, input_consumed, nullptr);
}

// Runs n tiles through forward + TTA combine.
// Source kinds: 0 = packed float32 tiles, 1 = packed uint8 tiles (src.ch channels),
// 2 = device-resident uint8 slide region (tile origins given).
// Sink: out (mode 0) or the slide accumulator (ys/xs given).
void run_tiles(adp_engine *e, int kind, FirstConvSrc src, const void *src_base, int n, int S, float mean, float std_,
               const int *ops_in, int n_ops_in, float *out, const int32_t *ys, const int32_t *xs,
               const long long *origins) {
  if (!e->packed) pack_all(e);
  ensure_arena(e, S);
  int ops[8] = {0}; int n_ops = 1;
  if (ops_in && n_ops_in > 0) {
    ADP_REQUIRE(n_ops_in <= 8, "at most 8 TTA ops");
    n_ops = n_ops_in;
    for (int i = 0; i < n_ops; ++i) { ADP_REQUIRE(ops_in[i] >= 0 && ops_in[i] < 8, "bad dihedral op"); ops[i] = ops_in[i]; }
  }
  ADP_REQUIRE(n_ops <= e->max_fw, "max_forwards smaller than the number of TTA ops");
  const int tpc = std::max(1, e->max_fw / n_ops);      // tiles per chunk
  const bool out_host = out && !is_device_ptr(out);
  const size_t tile_px = (size_t)S * S;
  const size_t src_tile_bytes = kind == 0 ? tile_px * 4 : (kind == 1 ? tile_px * src.ch : 0);
  const bool src_host = kind != 2 && !is_device_ptr(src_base);
  TtaOps tops; tops.n = n_ops;
  for (int i = 0; i < 8; ++i) tops.inv[i] = d4_inverse(ops[i < n_ops ? i : 0]);
  // Host buffers are staged through two device slots on a copy stream, so the H2D of chunk i+1 and
  // the D2H of chunk i-1 overlap the kernels of chunk i (events order slot reuse).
  auto h2d = [&](int ci) {       // issue the input copy of chunk ci on the copy stream
    const int t0 = ci * tpc, nt = std::min(tpc, n - t0), slot = ci & 1;
    e->in_stage2[slot].ensure((size_t)tpc * src_tile_bytes);
    if (ci >= 2) ADP_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_in_free[slot], 0));
    const uint8_t *sp = reinterpret_cast<const uint8_t *>(src_base) + (size_t)t0 * src_tile_bytes;
    ADP_CUDA(cudaMemcpyAsync(e->in_stage2[slot].p, sp, (size_t)nt * src_tile_bytes, cudaMemcpyHostToDevice, e->copy_stream));
    ADP_CUDA(cudaEventRecord(e->ev_in_ready[slot], e->copy_stream));
  };
  const int nchunks = cdiv(n, tpc);
  if (src_host) h2d(0);
  for (int ci = 0; ci < nchunks; ++ci) {
    const int t0 = ci * tpc;
    const int nt = std::min(tpc, n - t0);
    const int slot = ci & 1;
    FirstConvSrc s = src;
    if (kind != 2) {
      const void *dp = reinterpret_cast<const uint8_t *>(src_base) + (size_t)t0 * src_tile_bytes;
      if (src_host) {
        if (ci + 1 < nchunks) h2d(ci + 1);
        ADP_CUDA(cudaStreamWaitEvent(e->stream, e->ev_in_ready[slot], 0));
        dp = e->in_stage2[slot].p;
      }
      s.f32 = kind == 0 ? reinterpret_cast<const float *>(dp) : nullptr;
      s.u8 = kind == 1 ? reinterpret_cast<const uint8_t *>(dp) : nullptr;
    }
    FwTable fw;
    memset(&fw, 0, sizeof(fw));
    for (int t = 0; t < nt; ++t) {
      fw.origin[t] = origins ? origins[t0 + t] : 0;     // indexed by local tile
      for (int k = 0; k < n_ops; ++k) { fw.tile[t * n_ops + k] = t; fw.op[t * n_ops + k] = ops[k]; }
    }
    forward(e, s, fw, nt * n_ops, mean, std_, src_host ? e->ev_in_free[slot] : nullptr);
    dim3 grid(cdiv(S, 32), cdiv(S, 32)), block(32, 8);
    float *dout = nullptr;
    if (out) {
      if (out_host) {
        e->out_stage2[slot].ensure((size_t)tpc * tile_px * 4);
        dout = e->out_stage2[slot].as<float>();
        if (ci >= 2) ADP_CUDA(cudaStreamWaitEvent(e->stream, e->ev_out_free[slot], 0));
      } else {
        dout = out + (size_t)t0 * tile_px;
      }
    }
    for (int t = 0; t < nt; ++t) {
      const float *planes = e->prob.as<float>() + (size_t)t * n_ops * tile_px;
      const double by = (double)tile_px * 4 * (n_ops + (out ? 1 : (e->wsi_mode == ADP_BLEND_GAUSSIAN ? 5 : 4)));
      if (out) {
        e->launch("tta_combine", 0, by, [&] {
          tta_kernel()<<<grid, block, 0, e->stream>>>(planes, tops, S, 0, dout + (size_t)t * tile_px, nullptr, nullptr,
                                                         nullptr, 0, 0, 0, 0, 0);
        });
      } else {
        const int mode = e->wsi_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
        e->launch("tta_blend", 0, by, [&] {
          tta_kernel()<<<grid, block, 0, e->stream>>>(planes, tops, S, mode, nullptr, e->wsi_acc.as<float>(),
                                                         e->wsi_wsum.as<float>(), e->wsi_window.as<float>(), e->wsi_W,
                                                         0, e->wsi_rows, ys[t0 + t] - e->wsi_y0, xs[t0 + t]);
        });
      }
    }
    if (out && out_host) {
      ADP_CUDA(cudaEventRecord(e->ev_out_ready[slot], e->stream));
      ADP_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_out_ready[slot], 0));
      ADP_CUDA(cudaMemcpyAsync(out + (size_t)t0 * tile_px, dout, (size_t)nt * tile_px * 4, cudaMemcpyDeviceToHost, e->copy_stream));
      ADP_CUDA(cudaEventRecord(e->ev_out_free[slot], e->copy_stream));
    }
  }
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->copy_stream));
}

// device-visible copy of a small/large host or device array
const void *to_device(adp_engine *e, DevBuf &buf, const void *p, size_t bytes) {
  if (!p) return nullptr;
  if (is_device_ptr(p)) return p;
  buf.ensure(bytes);
  ADP_CUDA(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, e->stream));
  return buf.p;
}

void run_finalize(adp_engine *e, const float *acc, const float *wsum, int linear, size_t n, float thr, float *prob,
                  uint8_t *mask, const uint8_t *gt, int64_t counts[4]) {
  const bool prob_host = prob && !is_device_ptr(prob), mask_host = mask && !is_device_ptr(mask);
  float *dp = prob; uint8_t *dm = mask;
  if (prob_host) { e->fin_prob.ensure(n * 4); dp = e->fin_prob.as<float>(); }
  if (mask_host) { e->fin_mask.ensure(n); dm = e->fin_mask.as<uint8_t>(); }
  const uint8_t *dg = reinterpret_cast<const uint8_t *>(to_device(e, e->fin_gt, gt, n));
  e->counts.ensure(32);
  ADP_CUDA(cudaMemsetAsync(e->counts.p, 0, 32, e->stream));
  const int grid = e->wave_grid(finalize_kernel, cdiv64(n, 256 * 4));
  const double by = (double)n * ((wsum ? 8 : 4) + (dp ? 4 : 0) + (dm ? 1 : 0) + (dg ? 1 : 0));
  e->launch("finalize_threshold_metrics", 0, by, [&] {
    auto al = [](const void *q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
    const int vec = al(acc, 16) && al(wsum, 16) && al(dp, 16) && al(dm, 4) && al(dg, 4);
    finalize_kernel<<<grid, 256, 0, e->stream>>>(acc, wsum, linear, n, thr, dp, dm, dg,
                                                             e->counts.as<unsigned long long>(), vec);
  });
  if (prob_host) ADP_CUDA(cudaMemcpyAsync(prob, dp, n * 4, cudaMemcpyDeviceToHost, e->stream));
  if (mask_host) ADP_CUDA(cudaMemcpyAsync(mask, dm, n, cudaMemcpyDeviceToHost, e->stream));
  unsigned long long hc[4];
  ADP_CUDA(cudaMemcpyAsync(hc, e->counts.p, 32, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  if (counts) for (int i = 0; i < 4; ++i) counts[i] = (int64_t)hc[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Loss recipes (kernels_post.cuh): sums[8] = {S0 sum of the BCE terms in the mean, S1 sum ys*pc, S2 sum ys, S3 sum pc, S4 sum y*p,
// S5 sum p, S6 sum y, S7 number of BCE terms in the mean}.  Every entry is additive over data-parallel ranks; the eight
// values are produced and consumed ON THE DEVICE (a host copy is optional), so a training step needs no host round trip.
struct LossState {
  DevBuf own_sums, row_bce, row_w, sel_sum;
  bool ohem = false;
};

void loss_from_sums8(const double s[8], double out[4]) {
  const double bce = s[0] / s[7];
  const double dice_loss = 1.0 - (2.0 * s[1] + 1.0) / (s[2] + s[3] + 1.0);
  out[0] = bce + dice_loss; out[1] = bce; out[2] = dice_loss;
  out[3] = (2.0 * s[4] + 1.0) / (s[6] + s[5] + 1.0);
}

// forward part: the eight sums of `batch` images of npi pixels each (p, y device pointers) into dsums (device, 8 doubles).
// row_len = length of the trailing axis of the reference's (B, H, W) tensors: hard-example mining ranks per-ROW BCE means
// (train_adipose_unet_v3.py:301-313, see kernels_post.cuh); without hard mining it only shapes the reduction.
void loss_forward(adp_engine *e, LossState &ls, const LossRecipe &r, const float *p, const float *y, int batch, size_t npi,
                  int row_len, double *dsums) {
  const size_t n = (size_t)batch * npi;
  ls.ohem = r.ohem_keep < 1.f;
  int W = row_len > 0 ? row_len : (int)std::min<size_t>(npi, 1024);
  int R = 0; long long k = 0;
  if (ls.ohem) {
    ADP_REQUIRE(row_len > 0 && npi % (size_t)row_len == 0, "hard-example mining needs the row length of the (B,H,W) tensors");
    R = (int)(npi / (size_t)row_len);
    ADP_REQUIRE(R <= kOhemMaxRows, "hard-example mining: more than 8192 rows per image");
    // k = int(float32(rows) * keep_ratio): tf.cast(tf.cast(num_pixels, tf.float32) * keep_ratio, tf.int32) on the (B, H) tensor
    k = (long long)((float)R * r.ohem_keep);
    ADP_REQUIRE(k >= 1 && k <= R, "hard-example ratio selects no row");
    ls.row_bce.ensure((size_t)batch * R * 4); ls.row_w.ensure((size_t)batch * R * 4); ls.sel_sum.ensure((size_t)batch * 8);
  }
  ADP_CUDA(cudaMemsetAsync(dsums, 0, 64, e->stream));
  const size_t rows = (size_t)cdiv64((long long)n, W);
  const int grid = e->wave_grid(loss_reduce_kernel, (size_t)cdiv64((long long)rows, 8));
  e->launch("loss_reduce", 0, (double)n * 8, [&] {
    loss_reduce_kernel<<<grid, 256, 0, e->stream>>>(p, y, n, W, r.ys_scale(), r.eps_neg, ls.ohem ? ls.row_bce.as<float>() : nullptr, dsums);
  });
  if (ls.ohem)
    e->launch("ohem_select", 0, (double)batch * R * 8, [&] {
      ohem_select_kernel<<<batch, 1024, (size_t)R * 4, e->stream>>>(ls.row_bce.as<float>(), R, (int)k, 1.0f / (float)W, ls.row_w.as<float>(),
                                                                    ls.sel_sum.as<double>());
    });
  e->launch("loss_finish", 0, 0, [&] {
    loss_finish_kernel<<<1, 32, 0, e->stream>>>(dsums, ls.ohem ? ls.sel_sum.as<double>() : nullptr, batch,
                                                ls.ohem ? (double)batch * (double)k : (double)n);
  });
}

// backward part: dL/dp for the loss defined by the (possibly rank-summed) device sums
void loss_backward(adp_engine *e, LossState &ls, const LossRecipe &r, const float *p, const float *y, int batch, size_t npi,
                   int row_len, const double *dsums, float *dldp, float gain = 1.f) {
  const size_t n = (size_t)batch * npi;
  const int grid = e->wave_grid(loss_grad_kernel, cdiv64(n, 256 * 4));
  e->launch("loss_grad", 0, (double)n * 12, [&] {
    loss_grad_kernel<<<grid, 256, 0, e->stream>>>(p, y, n, r.ys_scale(), r.eps_neg, dsums, ls.ohem ? ls.row_w.as<float>() : nullptr,
                                                 row_len > 0 ? row_len : 1, gain, dldp);
  });
}

#include "train_host.cuh"

// ================================================================================================
// C ABI
// ================================================================================================
#define ADP_TRY try {
#define ADP_CATCH                                              \
  }                                                            \
  catch (const adp::Error &ex) { g_last_error = ex.what(); return ex.code; } \
  catch (const std::exception &ex) { g_last_error = ex.what(); return ADP_ECUDA; } \
  return ADP_OK;

extern "C" {

int adp_abi_version(void) { return ADP_ABI_VERSION; }
const char *adp_last_error(void) { return g_last_error.c_str(); }

int adp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ++ok;
  }
  return ok;
}

int adp_create(int device, int precision, int init_nb, int max_forwards, adp_engine **out) {
  ADP_TRY
  ADP_REQUIRE(out, "out is null");
  ADP_REQUIRE(precision >= 0 && precision <= 3, "precision");
  if (init_nb <= 0) init_nb = 44;
  ADP_REQUIRE(init_nb % 4 == 0 && init_nb <= 64, "init_nb must be a multiple of 4, at most 64");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); throw Error(ADP_ENODEV, "no CUDA device (libadipose_b200 has no CPU fallback)"); }
  ADP_REQUIRE(device >= 0 && device < n, "device index");
  cudaDeviceProp pr;
  ADP_CUDA(cudaGetDeviceProperties(&pr, device));
  if (pr.major != 10) throw Error(ADP_ENODEV, std::string("device is sm_") + std::to_string(pr.major * 10 + pr.minor) + ", this library is sm_100a only");
  ADP_CUDA(cudaSetDevice(device));
  std::unique_ptr<adp_engine> e(new adp_engine());
  e->device = device; e->prec_public = precision; e->init_nb = init_nb;
  e->split = precision == ADP_PREC_BF16X3;
  e->prec = e->split ? ADP_PREC_BF16 : precision;       // bf16x3 runs the tcgen05 path on hi/lo bf16 tensors
  precision = e->prec;
  e->max_fw = max_forwards > 0 ? std::min(max_forwards, 64) : 16;
  e->num_sms = pr.multiProcessorCount;
  e->esz = precision == ADP_PREC_FP32 ? 4 : 2;
  for (int i = 0; i < 4; ++i) { e->c[i] = init_nb << i; e->cp[i] = pad16(e->c[i]); }
  ADP_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  ADP_CUDA(cudaEventCreate(&e->ev0));
  ADP_CUDA(cudaEventCreate(&e->ev1));
  ADP_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    ADP_CUDA(cudaEventCreateWithFlags(&e->ev_in_ready[i], cudaEventDisableTiming));
    ADP_CUDA(cudaEventCreateWithFlags(&e->ev_in_free[i], cudaEventDisableTiming));
    ADP_CUDA(cudaEventCreateWithFlags(&e->ev_out_ready[i], cudaEventDisableTiming));
    ADP_CUDA(cudaEventCreateWithFlags(&e->ev_out_free[i], cudaEventDisableTiming));
  }
  for (int nt : {9, 4})
    for (int T : {4, 2, 1})
      for (int kys : {0, 1})
        for (int epi : {EPI_STORE, EPI_HEAD, EPI_POOL, EPI_BWD, EPI_UPSUM})
          if (ConvTcKernel k = tc_kernel_lookup(nt, T, kys, epi))
            ADP_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  ADP_CUDA(cudaFuncSetAttribute(conv_tc_kernel<9, 4, true, EPI_POOL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  ADP_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (const char *f = getenv("ADP_FUSE_FIRST")) e->fuse_first = atoi(f) != 0;        // A/B runs of whole programs (bench.py)
  if (const char *f = getenv("ADP_FUSE_DROPOUT")) e->fuse_dropout = atoi(f) != 0;
  if (const char *f = getenv("ADP_FUSE_UPSUM")) e->fuse_upsum = atoi(f) != 0;
  if (const char *f = getenv("ADP_RESIDENT_WEIGHTS")) e->resident_weights = atoi(f) != 0;
  if (const char *d = getenv("ADP_TC_DEBUG")) {
    e->dbg = atoi(d);
    if (e->dbg && !ADP_TC_DEBUG_BUILD) {
      fprintf(stderr, "libadipose_b200: ADP_TC_DEBUG ignored - the timing switches exist only in a debug build (build.py --force --debug)\n");
      e->dbg = 0;
    }
  }
  build_layers(e.get());
  *out = e.release();
  ADP_CATCH
}

int adp_destroy(adp_engine *e) {
  ADP_TRY
  if (!e) return ADP_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  for (int i = 0; i < 2; ++i) {
    if (e->ev_in_ready[i]) cudaEventDestroy(e->ev_in_ready[i]);
    if (e->ev_in_free[i]) cudaEventDestroy(e->ev_in_free[i]);
    if (e->ev_out_ready[i]) cudaEventDestroy(e->ev_out_ready[i]);
    if (e->ev_out_free[i]) cudaEventDestroy(e->ev_out_free[i]);
  }
  if (e->jpeg_st) nvjpeg_api().JpegStateDestroy(e->jpeg_st);
  if (e->jpeg_h) nvjpeg_api().Destroy(e->jpeg_h);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e->tr;
  delete e;
  ADP_CATCH
}

int adp_precision(const adp_engine *e) { return e ? e->prec_public : ADP_EINVAL; }

int adp_synchronize(adp_engine *e) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

void *adp_stream(adp_engine *e) { return e ? (void *)e->stream : nullptr; }

int adp_set_option(adp_engine *e, const char *key, int value) {
  ADP_TRY
  ADP_REQUIRE(e && key, "null argument");
  std::string k = key;
  if (k == "fuse_first") e->fuse_first = value != 0;
  else if (k == "fuse_dropout") e->fuse_dropout = value != 0;
  else if (k == "fuse_upsum") e->fuse_upsum = value != 0;
  else if (k == "resident_weights") e->resident_weights = value != 0;
  else if (k == "fuse_head") e->fuse_head = value != 0;
  else if (k == "fuse_pool") e->fuse_pool = value != 0;
  else if (k == "wgrad_simt") e->wgrad_simt = value != 0;
  else if (k == "dgrad_simt") e->dgrad_simt = value != 0;
  else if (k == "train_accuracy") e->train_accuracy = value != 0;
  else if (k == "train_eval_mode") e->train_eval_mode = value != 0;
  else if (k == "debug") e->dbg = ADP_TC_DEBUG_BUILD ? value : 0;
  else if (k == "kys") { e->kys = value != 0; e->packed = false; }   // ky-stacked MMA issue (conv_tc.cuh); re-plans on next use
  else throw Error(ADP_EINVAL, "unknown option " + k);
  ADP_CATCH
}

int adp_set_weight(adp_engine *e, const char *layer_name, const float *kernel, const int64_t kshape[4], const float *bias,
                   int64_t nbias) {
  ADP_TRY
  ADP_REQUIRE(e && layer_name && kernel && kshape && bias, "null argument");
  ADP_CUDA(cudaSetDevice(e->device));
  std::string n = layer_name;
  int64_t want[4];
  if (n == "down1_conv1") { want[0] = 3; want[1] = 3; want[2] = 1; want[3] = e->c[0]; }
  else if (n == "output_softmax") { want[0] = 1; want[1] = 1; want[2] = e->c[0]; want[3] = 2; }
  else if (n == "aux_out1") { want[0] = 1; want[1] = 1; want[2] = e->c[2]; want[3] = 1; }
  else if (n == "aux_out2") { want[0] = 1; want[1] = 1; want[2] = e->c[1]; want[3] = 1; }
  else { ConvLayer &L = layer(e, n); want[0] = 3; want[1] = 3; want[2] = L.cin; want[3] = L.cout; }
  for (int i = 0; i < 4; ++i)
    if (kshape[i] != want[i])
      throw Error(ADP_EINVAL, "kernel shape mismatch for " + n + ": expected (" + std::to_string(want[0]) + "," + std::to_string(want[1]) +
                                  "," + std::to_string(want[2]) + "," + std::to_string(want[3]) + ")");
  ADP_REQUIRE(nbias == want[3], "bias length mismatch");
  if (e->tr) throw Error(ADP_ESTATE, "weights cannot be replaced while a training state is open (adp_train_end first)");
  HostWeight &h = e->hw[n];
  const size_t ne = (size_t)(want[0] * want[1] * want[2] * want[3]);
  h.k.assign(kernel, kernel + ne);
  h.b.assign(bias, bias + nbias);
  for (int i = 0; i < 4; ++i) h.shape[i] = want[i];
  h.set = true;
  e->packed = false;
  ADP_CATCH
}

int adp_get_weight(adp_engine *e, const char *layer_name, float *kernel, int64_t kernel_elems, float *bias, int64_t nbias) {
  ADP_TRY
  ADP_REQUIRE(e && layer_name, "null argument");
  if (e->tr) { ADP_CUDA(cudaSetDevice(e->device)); sync_host_weights(e); }
  auto it = e->hw.find(layer_name);
  if (it == e->hw.end() || !it->second.set) throw Error(ADP_ESTATE, std::string("layer not set: ") + layer_name);
  const HostWeight &h = it->second;
  if (kernel) { ADP_REQUIRE(kernel_elems == (int64_t)h.k.size(), "kernel_elems"); memcpy(kernel, h.k.data(), h.k.size() * 4); }
  if (bias) { ADP_REQUIRE(nbias == (int64_t)h.b.size(), "nbias"); memcpy(bias, h.b.data(), h.b.size() * 4); }
  ADP_CATCH
}

int adp_weights_ready(adp_engine *e) {
  if (!e) return 0;
  for (const char *n : kAllNames) {
    auto it = e->hw.find(n);
    if (it == e->hw.end() || !it->second.set) return 0;
  }
  return 1;
}

int adp_tta_ops(int mode, int ops[8]) {
  static const int full[8] = {0, 1, 2, 3, 4, 5, 6, 7}, basic[4] = {0, 4, 5, 1}, minimal[2] = {0, 4};
  if (!ops) return ADP_EINVAL;
  switch (mode) {
    case ADP_TTA_NONE: ops[0] = 0; return 1;
    case ADP_TTA_MINIMAL: memcpy(ops, minimal, sizeof(minimal)); return 2;
    case ADP_TTA_BASIC: memcpy(ops, basic, sizeof(basic)); return 4;
    case ADP_TTA_FULL: memcpy(ops, full, sizeof(full)); return 8;
  }
  return ADP_EINVAL;
}

int adp_predict(adp_engine *e, const float *tiles, int n, int size, float mean, float std_, const int *ops, int n_ops,
                float *out) {
  ADP_TRY
  ADP_REQUIRE(e && tiles && out && n > 0, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  FirstConvSrc s{};
  run_tiles(e, 0, s, tiles, n, size, mean, std_, ops, n_ops, out, nullptr, nullptr, nullptr);
  ADP_CATCH
}

int adp_predict_u8(adp_engine *e, const uint8_t *tiles, int n, int size, int channels, float mean, float std_,
                   const int *ops, int n_ops, float *out) {
  ADP_TRY
  ADP_REQUIRE(e && tiles && out && n > 0, "null/empty argument");
  ADP_REQUIRE(channels == 1 || channels == 3, "channels must be 1 or 3");
  ADP_CUDA(cudaSetDevice(e->device));
  FirstConvSrc s{};
  s.ch = channels;
  run_tiles(e, 1, s, tiles, n, size, mean, std_, ops, n_ops, out, nullptr, nullptr, nullptr);
  ADP_CATCH
}

int adp_tta_combine(adp_engine *e, const float *planes, int size, const int *ops, int n_ops, float *out) {
  ADP_TRY
  ADP_REQUIRE(e && planes && ops && out && size > 0 && n_ops > 0 && n_ops <= 8, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t px = (size_t)size * size;
  DevBuf dpl, dout;
  const float *pl = reinterpret_cast<const float *>(to_device(e, dpl, planes, px * 4 * n_ops));
  const bool host = !is_device_ptr(out);
  float *o = out;
  if (host) { dout.ensure(px * 4); o = dout.as<float>(); }
  TtaOps tops; tops.n = n_ops;
  for (int i = 0; i < 8; ++i) {
    const int op = ops[i < n_ops ? i : 0];
    ADP_REQUIRE(op >= 0 && op < 8, "bad dihedral op");
    tops.inv[i] = d4_inverse(op);
  }
  dim3 grid(cdiv(size, 32), cdiv(size, 32)), block(32, 8);
  e->launch("tta_combine", 0, (double)px * 4 * (n_ops + 1), [&] {
    tta_kernel()<<<grid, block, 0, e->stream>>>(pl, tops, size, 0, o, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0);
  });
  if (host) ADP_CUDA(cudaMemcpyAsync(out, o, px * 4, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_debug_layer(adp_engine *e, const char *name, int idx, float *out, int64_t out_elems, int64_t shape_hwc[3]) {
  ADP_TRY
  ADP_REQUIRE(e && name && shape_hwc, "null argument");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_REQUIRE(e->S > 0 && idx >= 0 && idx < e->last_nfw, "no forward to tap / idx out of range");
  const int S = e->S;
  const int *cp = e->cp, *c = e->c;
  struct Tap { const DevBuf *b; int H, pitch, coff, C; };
  std::map<std::string, Tap> m = {
      {"down1_conv2", {&e->cat1, S, 2 * cp[0], 0, c[0]}},      {"up1_conv1", {&e->cat1, S, 2 * cp[0], cp[0], c[0]}},
      {"down2_conv2", {&e->cat2, S / 2, 2 * cp[1], 0, c[1]}},  {"up2_conv1", {&e->cat2, S / 2, 2 * cp[1], cp[1], c[1]}},
      {"down3_conv2", {&e->cat3, S / 4, 2 * cp[2], 0, c[2]}},  {"up3_conv1", {&e->cat3, S / 4, 2 * cp[2], cp[2], c[2]}},
      {"dilate1", {&e->t[0], S / 8, cp[3], 0, c[3]}},          {"dilate2", {&e->t[1], S / 8, cp[3], 0, c[3]}},
      {"dilate3", {&e->t[2], S / 8, cp[3], 0, c[3]}},          {"dilate4", {&e->t[3], S / 8, cp[3], 0, c[3]}},
      {"dilate5", {&e->t[4], S / 8, cp[3], 0, c[3]}},          {"dilate6", {&e->t[5], S / 8, cp[3], 0, c[3]}},
      {"dilate_add", {&e->ts, S / 8, cp[3], 0, c[3]}},
      {"up3_conv2", {&e->a3, S / 4, cp[2], 0, c[2]}},          {"up3_conv3", {&e->b3, S / 4, cp[2], 0, c[2]}},
      {"up2_conv2", {&e->a2, S / 2, cp[1], 0, c[1]}},          {"up2_conv3", {&e->b2, S / 2, cp[1], 0, c[1]}},
      {"up1_conv2", {&e->a1, S, cp[0], 0, c[0]}},              {"up1_conv3", {&e->b1, S, cp[0], 0, c[0]}},
      {"pool1", {&e->pl1, S / 2, cp[0], 0, c[0]}},             {"pool2", {&e->pl2, S / 4, cp[1], 0, c[1]}},
      {"pool3", {&e->pl3, S / 8, cp[2], 0, c[2]}},
  };
  std::string n = name;
  if (n == "prob") {
    shape_hwc[0] = S; shape_hwc[1] = S; shape_hwc[2] = 1;
    if (out) {
      ADP_REQUIRE(out_elems == (int64_t)S * S, "out_elems");
      ADP_CUDA(cudaMemcpy(out, e->prob.as<float>() + (size_t)idx * S * S, (size_t)S * S * 4, cudaMemcpyDeviceToHost));
    }
    return ADP_OK;
  }
  auto it = m.find(n);
  if (it == m.end()) throw Error(ADP_EINVAL, "layer not tappable (overwritten later in the forward): " + n);
  const Tap &t = it->second;
  shape_hwc[0] = t.H; shape_hwc[1] = t.H; shape_hwc[2] = t.C;
  if (!out) return ADP_OK;
  ADP_REQUIRE(out_elems == (int64_t)t.H * t.H * t.C, "out_elems");
  const size_t npx = (size_t)t.H * t.H, img = npx * t.pitch * e->esz * e->mul();
  std::vector<uint8_t> host(img);
  ADP_CUDA(cudaMemcpy(host.data(), (const uint8_t *)t.b->p + (size_t)idx * img, img, cudaMemcpyDeviceToHost));
  const int W = t.H, cgs = t.pitch / 8 * e->mul(), lo = e->split ? t.pitch / 8 : 0;     // bf16x3: value = hi + lo
  for (int y = 0; y < t.H; ++y)
    for (int x = 0; x < W; ++x)
      for (int ch = 0; ch < t.C; ++ch) {
        const int cc = t.coff + ch;       // row-planar: (((y*cgs + cg)*W + x)*8 + c%8
        const size_t si = (((size_t)y * cgs + cc / 8) * W + x) * 8 + cc % 8;
        float v = e->esz == 4 ? reinterpret_cast<const float *>(host.data())[si]
                              : __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(host.data())[si]);
        if (lo) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(host.data())[si + (size_t)lo * W * 8]);
        out[((size_t)y * W + x) * t.C + ch] = v;
      }
  ADP_CATCH
}

int adp_threshold_metrics(adp_engine *e, const float *prob, const uint8_t *gt, int64_t n_px, float thr, uint8_t *mask,
                          int64_t counts[4]) {
  ADP_TRY
  ADP_REQUIRE(e && prob && n_px > 0, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  DevBuf dp;
  const float *p = reinterpret_cast<const float *>(to_device(e, dp, prob, (size_t)n_px * 4));
  run_finalize(e, p, nullptr, 0, (size_t)n_px, thr, nullptr, mask, gt, counts);
  ADP_CATCH
}

int adp_threshold_sweep(adp_engine *e, const float *prob, const uint8_t *gt, int64_t n_px, const float *thresholds, int n_thr,
                        int64_t *counts) {
  ADP_TRY
  ADP_REQUIRE(e && prob && gt && thresholds && counts && n_px > 0, "null/empty argument");
  ADP_REQUIRE(n_thr >= 1 && n_thr <= 64, "1..64 candidate thresholds");
  for (int i = 1; i < n_thr; ++i) ADP_REQUIRE(thresholds[i] > thresholds[i - 1], "thresholds must be strictly ascending");
  ADP_CUDA(cudaSetDevice(e->device));
  DevBuf dp, dg, dt, dh;
  const size_t n = (size_t)n_px;
  const float *p = reinterpret_cast<const float *>(to_device(e, dp, prob, n * 4));
  const uint8_t *g = reinterpret_cast<const uint8_t *>(to_device(e, dg, gt, n));
  dt.ensure((size_t)n_thr * 4); dh.ensure((size_t)2 * (n_thr + 1) * 8);
  ADP_CUDA(cudaMemcpyAsync(dt.p, thresholds, (size_t)n_thr * 4, cudaMemcpyHostToDevice, e->stream));
  ADP_CUDA(cudaMemsetAsync(dh.p, 0, (size_t)2 * (n_thr + 1) * 8, e->stream));
  const int grid = e->wave_grid(threshold_sweep_kernel, cdiv64(n, 256 * 8));
  e->launch("threshold_sweep", 0, (double)n * 5, [&] {
    threshold_sweep_kernel<<<grid, 256, 0, e->stream>>>(p, g, n, dt.as<float>(), n_thr, dh.as<unsigned long long>());
  });
  std::vector<unsigned long long> h((size_t)2 * (n_thr + 1));
  ADP_CUDA(cudaMemcpyAsync(h.data(), dh.p, h.size() * 8, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  // predicted positive at threshold j  <=>  k > j
  for (int j = 0; j < n_thr; ++j) {
    unsigned long long tp = 0, fp = 0, fn = 0, tn = 0;
    for (int k = 0; k <= n_thr; ++k) {
      const unsigned long long neg = h[k], pos = h[(size_t)(n_thr + 1) + k];
      if (k > j) { tp += pos; fp += neg; } else { fn += pos; tn += neg; }
    }
    counts[4 * j + 0] = (int64_t)tp; counts[4 * j + 1] = (int64_t)fp; counts[4 * j + 2] = (int64_t)fn; counts[4 * j + 3] = (int64_t)tn;
  }
  ADP_CATCH
}

int adp_boundary_refine(adp_engine *e, const float *mask, int n, int H, int W, int kernel_size, int bilateral_d, float sigma_color,
                        float sigma_space, float *out) {
  ADP_TRY
  ADP_REQUIRE(e && mask && out && n > 0 && H > 0 && W > 0, "null/empty argument");
  ADP_REQUIRE(kernel_size >= 1 && kernel_size <= kRefineMaxK, "kernel_size 1..15");
  ADP_REQUIRE(bilateral_d >= 1 && bilateral_d <= kRefineMaxD, "bilateral_d 1..9");
  ADP_REQUIRE(sigma_color > 0.f && sigma_space > 0.f, "sigmas must be positive");
  ADP_REQUIRE(H >= kRefineMaxD && W >= kRefineMaxD, "image smaller than the bilateral window");
  ADP_CUDA(cudaSetDevice(e->device));
  // cv2.getStructuringElement(MORPH_ELLIPSE, (k, k)): row i holds ones in [c - dx, c + dx], dx = round(c * sqrt((r^2 - dy^2) / r^2))
  RefineSE se; se.n = 0;
  {
    const int r = kernel_size / 2, c = kernel_size / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < kernel_size; ++i) {
      const int dy = i - r;
      if (std::abs(dy) > r) continue;
      const int dx = (int)std::lrint(c * std::sqrt(((double)r * r - (double)dy * dy) * inv_r2));
      for (int j = std::max(c - dx, 0); j < std::min(c + dx + 1, kernel_size); ++j) { se.dy[se.n] = (signed char)(i - r); se.dx[se.n] = (signed char)(j - c); ++se.n; }
    }
  }
  // cv2.bilateralFilter taps: (i, j) with sqrt(i^2 + j^2) <= radius, row-major; weights rounded to float32 from double exp
  RefineTaps bt; bt.n = 0;
  {
    const int radius = std::max(bilateral_d / 2, 1);
    const double gs = -0.5 / ((double)sigma_space * sigma_space);
    for (int i = -radius; i <= radius; ++i)
      for (int j = -radius; j <= radius; ++j) {
        const double rr = std::sqrt((double)i * i + (double)j * j);
        if (rr > radius) continue;
        bt.dy[bt.n] = (signed char)i; bt.dx[bt.n] = (signed char)j; bt.sw[bt.n] = (float)std::exp(rr * rr * gs); ++bt.n;
      }
  }
  float cw[256];
  {
    const double gc = -0.5 / ((double)sigma_color * sigma_color);
    for (int i = 0; i < 256; ++i) cw[i] = (float)std::exp((double)i * i * gc);
  }
  const size_t npx = (size_t)n * H * W;
  DevBuf din, dout, ua, ub, dcw;
  const float *m = reinterpret_cast<const float *>(to_device(e, din, mask, npx * 4));
  const bool out_host = !is_device_ptr(out);
  float *o = out;
  if (out_host) { dout.ensure(npx * 4); o = dout.as<float>(); }
  ua.ensure(npx); ub.ensure(npx); dcw.ensure(256 * 4);
  ADP_CUDA(cudaMemcpyAsync(dcw.p, cw, 256 * 4, cudaMemcpyHostToDevice, e->stream));
  uint8_t *A = ua.as<uint8_t>(), *B = ub.as<uint8_t>();
  const dim3 grid(cdiv(W, 32), cdiv(H, 8), n), block(32, 8);
  e->launch("refine_quantize", 0, (double)npx * 5, [&] {
    refine_quantize_kernel<<<e->wave_grid(refine_quantize_kernel, cdiv64(npx, 256)), 256, 0, e->stream>>>(m, A, npx);
  });
  e->launch("refine_band_bilateral", 0, (double)npx * 2, [&] { refine_band_kernel<<<grid, block, 0, e->stream>>>(A, B, H, W, se, bt, dcw.as<float>()); });
  // MORPH_OPEN = dilate(erode), MORPH_CLOSE = erode(dilate)  (:389-390); the last pass also writes refined / 255.0 (:393)
  e->launch("refine_morph", 0, (double)npx * 2, [&] { refine_morph_kernel<true><<<grid, block, 0, e->stream>>>(B, A, nullptr, H, W, se); });
  e->launch("refine_morph", 0, (double)npx * 2, [&] { refine_morph_kernel<false><<<grid, block, 0, e->stream>>>(A, B, nullptr, H, W, se); });
  e->launch("refine_morph", 0, (double)npx * 2, [&] { refine_morph_kernel<false><<<grid, block, 0, e->stream>>>(B, A, nullptr, H, W, se); });
  e->launch("refine_morph", 0, (double)npx * 6, [&] { refine_morph_kernel<true><<<grid, block, 0, e->stream>>>(A, B, o, H, W, se); });
  if (out_host) ADP_CUDA(cudaMemcpyAsync(out, o, npx * 4, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_blend_reconstruct(adp_engine *e, int blend_mode, const float *tiles, int n, int th, int tw, const int32_t *ys,
                          const int32_t *xs, const float *window, int H, int W, float *out) {
  ADP_TRY
  ADP_REQUIRE(e && out && H > 0 && W > 0 && n >= 0, "null/empty argument");
  ADP_REQUIRE(blend_mode == ADP_BLEND_GAUSSIAN || blend_mode == ADP_BLEND_LINEAR, "blend_mode");
  ADP_REQUIRE(blend_mode == ADP_BLEND_LINEAR || window, "Gaussian blending needs a window");
  ADP_REQUIRE(n == 0 || (tiles && ys && xs && th > 0 && tw > 0), "tiles/positions");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t npx = (size_t)H * W, tpx = (size_t)th * tw;
  DevBuf acc, ws, dwin, dtiles;
  acc.ensure(npx * 4); ws.ensure(npx * 4);
  ADP_CUDA(cudaMemsetAsync(acc.p, 0, npx * 4, e->stream));
  ADP_CUDA(cudaMemsetAsync(ws.p, 0, npx * 4, e->stream));
  const float *dw = reinterpret_cast<const float *>(to_device(e, dwin, window, tpx * 4));
  const bool th_host = n && !is_device_ptr(tiles);
  const int chunk = (int)std::max<size_t>(1, (256u << 20) / (tpx * 4));
  const int mode = blend_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
  const int grid = (int)std::min<size_t>(cdiv64(tpx, 256), (size_t)e->num_sms * 8);
  for (int t0 = 0; t0 < n; t0 += chunk) {
    const int nt = std::min(chunk, n - t0);
    const float *dt = tiles + (size_t)t0 * tpx;
    if (th_host) {
      dtiles.ensure((size_t)chunk * tpx * 4);
      ADP_CUDA(cudaMemcpyAsync(dtiles.p, dt, (size_t)nt * tpx * 4, cudaMemcpyHostToDevice, e->stream));
      dt = dtiles.as<float>();
    }
    for (int t = 0; t < nt; ++t)   // stream order == list order: same accumulation order as the NumPy loop
      e->launch("blend_tile", 0, (double)tpx * 4 * (mode == 1 ? 6 : 5), [&] {
        blend_tile_kernel<<<grid, 256, 0, e->stream>>>(dt + (size_t)t * tpx, th, tw, mode, acc.as<float>(), ws.as<float>(), dw,
                                                      tw, W, 0, H, ys[t0 + t], xs[t0 + t]);
      });
    if (th_host) ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  run_finalize(e, acc.as<float>(), ws.as<float>(), blend_mode == ADP_BLEND_LINEAR, npx, 0.5f, out, nullptr, nullptr, nullptr);
  ADP_CATCH
}

int adp_wsi_begin(adp_engine *e, int rows, int W, int y0, int tile, int blend_mode, const float *window) {
  ADP_TRY
  ADP_REQUIRE(e && rows > 0 && W > 0 && tile > 0, "null/empty argument");
  ADP_REQUIRE(blend_mode == ADP_BLEND_GAUSSIAN || blend_mode == ADP_BLEND_LINEAR, "blend_mode");
  ADP_REQUIRE(blend_mode == ADP_BLEND_LINEAR || window, "Gaussian blending needs a window");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t npx = (size_t)rows * W;
  e->wsi_acc.ensure(npx * 4); e->wsi_wsum.ensure(npx * 4);
  ADP_CUDA(cudaMemsetAsync(e->wsi_acc.p, 0, npx * 4, e->stream));
  ADP_CUDA(cudaMemsetAsync(e->wsi_wsum.p, 0, npx * 4, e->stream));
  if (window) {
    e->wsi_window.ensure((size_t)tile * tile * 4);
    ADP_CUDA(cudaMemcpyAsync(e->wsi_window.p, window, (size_t)tile * tile * 4, cudaMemcpyDefault, e->stream));
  }
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  e->wsi_rows = rows; e->wsi_W = W; e->wsi_y0 = y0; e->wsi_tile = tile; e->wsi_mode = blend_mode; e->wsi_on = true;
  ADP_CATCH
}

int adp_wsi_push_tiles(adp_engine *e, const float *tiles, int n, const int32_t *ys, const int32_t *xs, float mean,
                       float std_, const int *ops, int n_ops) {
  ADP_TRY
  ADP_REQUIRE(e && tiles && ys && xs && n > 0, "null/empty argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  FirstConvSrc s{};
  run_tiles(e, 0, s, tiles, n, e->wsi_tile, mean, std_, ops, n_ops, nullptr, ys, xs, nullptr);
  ADP_CATCH
}

int adp_wsi_push_from_slide(adp_engine *e, const uint8_t *slide, int channels, int region_y0, int region_rows, int n,
                            const int32_t *ys, const int32_t *xs, float mean, float std_, const int *ops, int n_ops,
                            int defer_below_row) {
  ADP_TRY
  ADP_REQUIRE(e && slide && ys && xs && n > 0, "null/empty argument");
  ADP_REQUIRE(channels == 1 || channels == 3, "channels must be 1 (gray) or 3 (interleaved RGB)");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(is_device_ptr(slide), "slide region must be device-resident (copy it once, then push tile positions)");
  ADP_CUDA(cudaSetDevice(e->device));
  const int S = e->wsi_tile;
  std::vector<long long> org(n);
  for (int i = 0; i < n; ++i) {
    ADP_REQUIRE(ys[i] >= region_y0 && ys[i] + S <= region_y0 + region_rows && xs[i] >= 0 && xs[i] + S <= e->wsi_W,
                "tile outside the resident slide region");
    org[i] = (long long)(ys[i] - region_y0) * e->wsi_W + xs[i];
  }
  FirstConvSrc s{};
  s.u8 = slide; s.ch = channels; s.slideW = e->wsi_W;
  const int zone = defer_below_row - e->wsi_y0;          // accumulator rows [0, zone) are deferred
  if (zone <= 0) {
    run_tiles(e, 2, s, nullptr, n, S, mean, std_, ops, n_ops, nullptr, ys, xs, org.data());
    return ADP_OK;
  }
  // deferred boundary zone: keep the tiles' probabilities, blend the rows below the zone now, the zone rows on replay
  ADP_REQUIRE(zone <= e->wsi_rows, "defer_below_row outside the accumulator");
  const size_t tpx = (size_t)S * S;
  adp_engine::Deferred d;
  d.probs.reset(new DevBuf());
  d.probs->ensure((size_t)n * tpx * 4);
  d.ys.assign(ys, ys + n); d.xs.assign(xs, xs + n); d.below = zone;
  run_tiles(e, 2, s, nullptr, n, S, mean, std_, ops, n_ops, d.probs->as<float>(), ys, xs, org.data());
  const int mode = e->wsi_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
  const int grid = (int)std::min<size_t>(cdiv64(tpx, 256), (size_t)e->num_sms * 8);
  for (int t = 0; t < n; ++t)
    e->launch("blend_tile", 0, (double)tpx * 4 * (mode == 1 ? 6 : 5), [&] {
      blend_tile_kernel<<<grid, 256, 0, e->stream>>>(d.probs->as<float>() + (size_t)t * tpx, S, S, mode, e->wsi_acc.as<float>(),
                                                    e->wsi_wsum.as<float>(), e->wsi_window.as<float>(), S, e->wsi_W, zone, e->wsi_rows,
                                                    ys[t] - e->wsi_y0, xs[t]);
    });
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  e->wsi_deferred.push_back(std::move(d));
  ADP_CATCH
}

int adp_wsi_replay_deferred(adp_engine *e) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  const int S = e->wsi_tile;
  const size_t tpx = (size_t)S * S;
  const int mode = e->wsi_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
  const int grid = (int)std::min<size_t>(cdiv64(tpx, 256), (size_t)e->num_sms * 8);
  for (auto &d : e->wsi_deferred)      // call order == list order: the zone rows see their tiles in row-major order
    for (size_t t = 0; t < d.ys.size(); ++t)
      e->launch("blend_tile", 0, (double)tpx * 4 * (mode == 1 ? 6 : 5), [&] {
        blend_tile_kernel<<<grid, 256, 0, e->stream>>>(d.probs->as<float>() + t * tpx, S, S, mode, e->wsi_acc.as<float>(),
                                                      e->wsi_wsum.as<float>(), e->wsi_window.as<float>(), S, e->wsi_W, 0, d.below,
                                                      d.ys[t] - e->wsi_y0, d.xs[t]);
      });
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  e->wsi_deferred.clear();
  ADP_CATCH
}

int adp_wsi_push_probs(adp_engine *e, const float *probs, int n, const int32_t *ys, const int32_t *xs) {
  ADP_TRY
  ADP_REQUIRE(e && probs && ys && xs && n > 0, "null/empty argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  const int S = e->wsi_tile;
  const size_t tpx = (size_t)S * S;
  const bool host = !is_device_ptr(probs);
  const int mode = e->wsi_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
  const int grid = (int)std::min<size_t>(cdiv64(tpx, 256), (size_t)e->num_sms * 8);
  const int chunk = 16;
  for (int t0 = 0; t0 < n; t0 += chunk) {
    const int nt = std::min(chunk, n - t0);
    const float *dt = probs + (size_t)t0 * tpx;
    if (host) {
      e->in_stage.ensure((size_t)chunk * tpx * 4);
      ADP_CUDA(cudaMemcpyAsync(e->in_stage.p, dt, (size_t)nt * tpx * 4, cudaMemcpyHostToDevice, e->stream));
      dt = e->in_stage.as<float>();
    }
    for (int t = 0; t < nt; ++t)
      e->launch("blend_tile", 0, (double)tpx * 4 * (mode == 1 ? 6 : 5), [&] {
        blend_tile_kernel<<<grid, 256, 0, e->stream>>>(dt + (size_t)t * tpx, S, S, mode, e->wsi_acc.as<float>(),
                                                      e->wsi_wsum.as<float>(), e->wsi_window.as<float>(), S, e->wsi_W,
                                                      0, e->wsi_rows, ys[t0 + t] - e->wsi_y0, xs[t0 + t]);
      });
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  ADP_CATCH
}

int adp_wsi_export(adp_engine *e, int y, int rows, float *acc, float *weight) {
  ADP_TRY
  ADP_REQUIRE(e && acc && weight, "null argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, nb = (size_t)rows * e->wsi_W * 4;
  ADP_CUDA(cudaMemcpyAsync(acc, e->wsi_acc.as<float>() + off, nb, cudaMemcpyDefault, e->stream));
  ADP_CUDA(cudaMemcpyAsync(weight, e->wsi_wsum.as<float>() + off, nb, cudaMemcpyDefault, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_wsi_import_add(adp_engine *e, int y, int rows, const float *acc, const float *weight) {
  ADP_TRY
  ADP_REQUIRE(e && acc && weight, "null argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, n = (size_t)rows * e->wsi_W;
  DevBuf da, dw;
  const float *a = reinterpret_cast<const float *>(to_device(e, da, acc, n * 4));
  const float *w = reinterpret_cast<const float *>(to_device(e, dw, weight, n * 4));
  const int grid = e->wave_grid(add_partial_kernel, cdiv64(n, 256));
  e->launch("wsi_add_partial", 0, (double)n * 24, [&] {
    add_partial_kernel<<<grid, 256, 0, e->stream>>>(e->wsi_acc.as<float>() + off, e->wsi_wsum.as<float>() + off, a, w, n);
  });
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_wsi_finalize(adp_engine *e, int y, int rows, float thr, float *prob, uint8_t *mask, const uint8_t *gt,
                     int64_t counts[4]) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, n = (size_t)rows * e->wsi_W;
  run_finalize(e, e->wsi_acc.as<float>() + off, e->wsi_wsum.as<float>() + off, e->wsi_mode == ADP_BLEND_LINEAR, n, thr,
               prob, mask, gt, counts);
  ADP_CATCH
}

int adp_wsi_end(adp_engine *e) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_CUDA(cudaSetDevice(e->device));
  e->wsi_acc.release(); e->wsi_wsum.release(); e->wsi_window.release();
  e->wsi_deferred.clear();
  e->wsi_aux.release(); e->wsi_naux = 0;
  e->wsi_on = false;
  ADP_CATCH
}

// ------------------------------------------------------------------------------------------------
// tile I/O front-end: nvJPEG decode, uint8 tile push, auxiliary planes, fat-%, TIFF-LZW
#define ADP_NVJPEG(expr)                                                                                     \
  do {                                                                                                       \
    nvjpegStatus_t _s = (expr);                                                                              \
    if (_s != NVJPEG_STATUS_SUCCESS) throw Error(ADP_ECUDA, std::string(#expr) + ": nvjpeg status " + std::to_string((int)_s)); \
  } while (0)

int adp_jpeg_decode(adp_engine *e, const uint8_t *const *data, const size_t *lengths, int n, int size, uint8_t *gray_host,
                    uint8_t *rgb_host, uint8_t **gray_dev, uint8_t **rgb_dev) {
  ADP_TRY
  ADP_REQUIRE(e && data && lengths && n > 0 && size > 0, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  if (!e->jpeg_h) {
    // interpolated chroma upsampling = libjpeg's "fancy upsampling", the default of the cv2.imread the reference decodes with
    ADP_NVJPEG(nvjpeg_api().CreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION, &e->jpeg_h));
    ADP_NVJPEG(nvjpeg_api().JpegStateCreate(e->jpeg_h, &e->jpeg_st));
  }
  const bool want_gray = gray_host || gray_dev, want_rgb = rgb_host || rgb_dev;
  const size_t px = (size_t)size * size;
  if (want_gray) e->jpeg_gray.ensure((size_t)n * px);
  if (want_rgb) e->jpeg_rgb.ensure((size_t)n * px * 3);
  for (int i = 0; i < n; ++i) {
    int ncomp = 0; nvjpegChromaSubsampling_t sub;
    int ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    ADP_NVJPEG(nvjpeg_api().GetImageInfo(e->jpeg_h, data[i], lengths[i], &ncomp, &sub, ws, hs));
    if (ws[0] != size || hs[0] != size)
      throw Error(ADP_EINVAL, "JPEG tile " + std::to_string(i) + " is " + std::to_string(ws[0]) + "x" + std::to_string(hs[0]) + ", expected " +
                                  std::to_string(size) + "x" + std::to_string(size));
    if (want_gray) {       // the luma plane: what libjpeg hands cv2.imread(..., IMREAD_GRAYSCALE) (out_color_space = JCS_GRAYSCALE)
      nvjpegImage_t img; memset(&img, 0, sizeof(img));
      img.channel[0] = e->jpeg_gray.as<unsigned char>() + (size_t)i * px; img.pitch[0] = (size_t)size;
      ADP_NVJPEG(nvjpeg_api().Decode(e->jpeg_h, e->jpeg_st, data[i], lengths[i], NVJPEG_OUTPUT_Y, &img, e->stream));
    }
    if (want_rgb) {
      nvjpegImage_t img; memset(&img, 0, sizeof(img));
      img.channel[0] = e->jpeg_rgb.as<unsigned char>() + (size_t)i * px * 3; img.pitch[0] = (size_t)size * 3;
      ADP_NVJPEG(nvjpeg_api().Decode(e->jpeg_h, e->jpeg_st, data[i], lengths[i], NVJPEG_OUTPUT_RGBI, &img, e->stream));
    }
    ++e->launches;
  }
  if (gray_host) ADP_CUDA(cudaMemcpyAsync(gray_host, e->jpeg_gray.p, (size_t)n * px, cudaMemcpyDeviceToHost, e->stream));
  if (rgb_host) ADP_CUDA(cudaMemcpyAsync(rgb_host, e->jpeg_rgb.p, (size_t)n * px * 3, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  if (gray_dev) *gray_dev = e->jpeg_gray.as<uint8_t>();
  if (rgb_dev) *rgb_dev = e->jpeg_rgb.as<uint8_t>();
  ADP_CATCH
}

int adp_wsi_push_tiles_u8(adp_engine *e, const uint8_t *tiles, int channels, int n, const int32_t *ys, const int32_t *xs, float mean,
                          float std_, const int *ops, int n_ops) {
  ADP_TRY
  ADP_REQUIRE(e && tiles && ys && xs && n > 0, "null/empty argument");
  ADP_REQUIRE(channels == 1 || channels == 3, "channels must be 1 or 3");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  FirstConvSrc s{};
  s.ch = channels;
  run_tiles(e, 1, s, tiles, n, e->wsi_tile, mean, std_, ops, n_ops, nullptr, ys, xs, nullptr);
  ADP_CATCH
}

int adp_wsi_aux_begin(adp_engine *e, int n_planes) {
  ADP_TRY
  ADP_REQUIRE(e && n_planes >= 1 && n_planes <= 4, "1..4 auxiliary planes");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t bytes = (size_t)n_planes * e->wsi_rows * e->wsi_W * 4;
  e->wsi_aux.ensure(bytes);
  ADP_CUDA(cudaMemsetAsync(e->wsi_aux.p, 0, bytes, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  e->wsi_naux = n_planes;
  ADP_CATCH
}

int adp_wsi_push_aux(adp_engine *e, int plane0, int n_planes, const void *tiles, int is_u8, int n, const int32_t *ys, const int32_t *xs) {
  ADP_TRY
  ADP_REQUIRE(e && tiles && ys && xs && n > 0, "null/empty argument");
  if (!e->wsi_on || e->wsi_naux == 0) throw Error(ADP_ESTATE, "adp_wsi_aux_begin not called");
  ADP_REQUIRE(plane0 >= 0 && (n_planes == 1 || n_planes == 3) && plane0 + n_planes <= e->wsi_naux, "plane range");
  ADP_CUDA(cudaSetDevice(e->device));
  const int S = e->wsi_tile;
  const size_t tpx = (size_t)S * S, tb = tpx * n_planes * (is_u8 ? 1 : 4);
  const size_t plane_stride = (size_t)e->wsi_rows * e->wsi_W;
  const int mode = e->wsi_mode == ADP_BLEND_GAUSSIAN ? 1 : 2;
  const int grid = (int)std::min<size_t>(cdiv64(tpx, 256), (size_t)e->num_sms * 8);
  const bool host = !is_device_ptr(tiles);
  const int chunk = 16;
  for (int t0 = 0; t0 < n; t0 += chunk) {
    const int nt = std::min(chunk, n - t0);
    const uint8_t *dt = reinterpret_cast<const uint8_t *>(tiles) + (size_t)t0 * tb;
    if (host) {
      e->aux_stage.ensure((size_t)chunk * tb);
      ADP_CUDA(cudaMemcpyAsync(e->aux_stage.p, dt, (size_t)nt * tb, cudaMemcpyHostToDevice, e->stream));
      dt = e->aux_stage.as<uint8_t>();
    }
    for (int t = 0; t < nt; ++t)     // stream order == list order: the accumulation order of the NumPy loop
      e->launch("aux_blend", 0, (double)tpx * n_planes * (is_u8 ? 9 : 12), [&] {
        aux_blend_kernel<<<grid, 256, 0, e->stream>>>(dt + (size_t)t * tb, is_u8, n_planes, S, mode, e->wsi_aux.as<float>() + (size_t)plane0 * plane_stride,
                                                     plane_stride, e->wsi_window.as<float>(), e->wsi_W, 0, e->wsi_rows, ys[t0 + t] - e->wsi_y0,
                                                     xs[t0 + t]);
      });
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  ADP_CATCH
}

int adp_wsi_export_u8(adp_engine *e, int plane0, int n_planes, int reverse, int y, int rows, uint8_t *out) {
  ADP_TRY
  ADP_REQUIRE(e && out, "null argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_REQUIRE((plane0 == -1 && n_planes == 1) || (plane0 >= 0 && n_planes >= 1 && plane0 + n_planes <= e->wsi_naux), "plane range");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, n = (size_t)rows * e->wsi_W;
  const size_t plane_stride = (size_t)e->wsi_rows * e->wsi_W;
  const float *planes = plane0 < 0 ? e->wsi_acc.as<float>() + off : e->wsi_aux.as<float>() + (size_t)plane0 * plane_stride + off;
  const bool host = !is_device_ptr(out);
  uint8_t *o = out;
  if (host) { e->fin_mask.ensure(n * n_planes); o = e->fin_mask.as<uint8_t>(); }
  const int grid = e->wave_grid(aux_export_u8_kernel, cdiv64(n, 256));
  e->launch("aux_export_u8", 0, (double)n * n_planes * 5 + (double)n * 4, [&] {
    aux_export_u8_kernel<<<grid, 256, 0, e->stream>>>(planes, plane_stride, e->wsi_wsum.as<float>() + off, e->wsi_mode == ADP_BLEND_LINEAR, n,
                                                     n_planes, reverse, o);
  });
  if (host) ADP_CUDA(cudaMemcpyAsync(out, o, n * n_planes, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_wsi_export_f32(adp_engine *e, int plane, int y, int rows, float *out) {
  ADP_TRY
  ADP_REQUIRE(e && out, "null argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_REQUIRE(plane >= 0 && plane < e->wsi_naux, "plane");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, n = (size_t)rows * e->wsi_W;
  const size_t plane_stride = (size_t)e->wsi_rows * e->wsi_W;
  const bool host = !is_device_ptr(out);
  float *o = out;
  if (host) { e->fin_prob.ensure(n * 4); o = e->fin_prob.as<float>(); }
  const int grid = e->wave_grid(aux_export_f32_kernel, cdiv64(n, 256));
  e->launch("aux_export_f32", 0, (double)n * 12, [&] {
    aux_export_f32_kernel<<<grid, 256, 0, e->stream>>>(e->wsi_aux.as<float>() + (size_t)plane * plane_stride + off, e->wsi_wsum.as<float>() + off,
                                                      e->wsi_mode == ADP_BLEND_LINEAR, n, o);
  });
  if (host) ADP_CUDA(cudaMemcpyAsync(out, o, n * 4, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_wsi_finalize_auxgt(adp_engine *e, int gt_plane, int y, int rows, float thr, float *prob, uint8_t *mask, int64_t counts[4]) {
  ADP_TRY
  ADP_REQUIRE(e && counts, "null argument");
  if (!e->wsi_on) throw Error(ADP_ESTATE, "adp_wsi_begin not called");
  ADP_REQUIRE(y >= e->wsi_y0 && rows > 0 && y + rows <= e->wsi_y0 + e->wsi_rows, "row range outside the accumulator");
  ADP_REQUIRE(gt_plane >= 0 && gt_plane < e->wsi_naux, "gt_plane");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t off = (size_t)(y - e->wsi_y0) * e->wsi_W, n = (size_t)rows * e->wsi_W;
  const size_t plane_stride = (size_t)e->wsi_rows * e->wsi_W;
  const bool prob_host = prob && !is_device_ptr(prob), mask_host = mask && !is_device_ptr(mask);
  float *dp = prob; uint8_t *dm = mask;
  if (prob_host) { e->fin_prob.ensure(n * 4); dp = e->fin_prob.as<float>(); }
  if (mask_host) { e->fin_mask.ensure(n); dm = e->fin_mask.as<uint8_t>(); }
  e->counts.ensure(32);
  ADP_CUDA(cudaMemsetAsync(e->counts.p, 0, 32, e->stream));
  const int grid = e->wave_grid(finalize_auxgt_kernel, cdiv64(n, 256));
  e->launch("finalize_threshold_metrics", 0, (double)n * 17, [&] {
    finalize_auxgt_kernel<<<grid, 256, 0, e->stream>>>(e->wsi_acc.as<float>() + off, e->wsi_wsum.as<float>() + off,
                                                      e->wsi_aux.as<float>() + (size_t)gt_plane * plane_stride + off, e->wsi_mode == ADP_BLEND_LINEAR,
                                                      n, thr, dp, dm, e->counts.as<unsigned long long>());
  });
  if (prob_host) ADP_CUDA(cudaMemcpyAsync(prob, dp, n * 4, cudaMemcpyDeviceToHost, e->stream));
  if (mask_host) ADP_CUDA(cudaMemcpyAsync(mask, dm, n, cudaMemcpyDeviceToHost, e->stream));
  unsigned long long hc[4];
  ADP_CUDA(cudaMemcpyAsync(hc, e->counts.p, 32, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 4; ++i) counts[i] = (int64_t)hc[i];
  ADP_CATCH
}

int adp_tile_fat_percent(adp_engine *e, const float *probs, int n, int64_t px, float thr, double *pct) {
  ADP_TRY
  ADP_REQUIRE(e && probs && pct && n > 0 && px > 0, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  DevBuf dp, dc;
  const float *p = reinterpret_cast<const float *>(to_device(e, dp, probs, (size_t)n * px * 4));
  dc.ensure((size_t)n * 8);
  ADP_CUDA(cudaMemsetAsync(dc.p, 0, (size_t)n * 8, e->stream));
  const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(cdiv64(px, 256 * 8), (int64_t)e->num_sms * 2)), (unsigned)n);
  e->launch("tile_fat_count", 0, (double)n * px * 4, [&] {
    tile_count_kernel<<<grid, 256, 0, e->stream>>>(p, (size_t)px, thr, dc.as<unsigned long long>());
  });
  std::vector<unsigned long long> h(n);
  ADP_CUDA(cudaMemcpyAsync(h.data(), dc.p, (size_t)n * 8, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  // (fat_pixels / total_pixels) * 100.0 in float64, tile_classification_evaluation.py:222-225
  for (int i = 0; i < n; ++i) pct[i] = ((double)h[i] / (double)px) * 100.0;
  ADP_CATCH
}

int adp_tiff_write_lzw(const char *path, const uint8_t *data, int64_t H, int64_t W, int channels, int rows_per_strip, int threads) {
  ADP_TRY
  std::string msg;
  const int rc = adp_tiff::write_lzw(path, data, H, W, channels, rows_per_strip, threads, msg);
  if (rc != 0) throw Error(ADP_EINVAL, msg);
  ADP_CATCH
}

int adp_loss_metrics_ex(adp_engine *e, const float *p, const float *y, int batch, int64_t px_per_image, int64_t row_len,
                        float ohem_keep_ratio, float eps_pos, float eps_neg, float *dldp, double out[4]) {
  ADP_TRY
  ADP_REQUIRE(e && p && y && out && batch > 0 && px_per_image > 0, "null/empty argument");
  ADP_REQUIRE(ohem_keep_ratio > 0.f && ohem_keep_ratio <= 1.f && eps_pos >= 0.f && eps_neg >= 0.f && eps_pos + eps_neg < 1.f,
              "loss recipe out of range");
  ADP_REQUIRE(row_len >= 0 && row_len <= px_per_image && row_len <= (1 << 30), "row_len");
  ADP_CUDA(cudaSetDevice(e->device));
  DevBuf dp, dy, dg;
  const size_t n = (size_t)batch * (size_t)px_per_image;
  const float *pp = reinterpret_cast<const float *>(to_device(e, dp, p, n * 4));
  const float *yy = reinterpret_cast<const float *>(to_device(e, dy, y, n * 4));
  LossRecipe r; r.ohem_keep = ohem_keep_ratio; r.eps_pos = eps_pos; r.eps_neg = eps_neg;
  LossState ls;
  ls.own_sums.ensure(64);
  double s[8];
  loss_forward(e, ls, r, pp, yy, batch, (size_t)px_per_image, (int)row_len, ls.own_sums.as<double>());
  ADP_CUDA(cudaMemcpyAsync(s, ls.own_sums.p, 64, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  loss_from_sums8(s, out);
  if (dldp) {
    const bool host = !is_device_ptr(dldp);
    float *g = dldp;
    if (host) { dg.ensure(n * 4); g = dg.as<float>(); }
    loss_backward(e, ls, r, pp, yy, batch, (size_t)px_per_image, (int)row_len, ls.own_sums.as<double>(), g);
    if (host) ADP_CUDA(cudaMemcpyAsync(dldp, g, n * 4, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  ADP_CATCH
}

int adp_loss_metrics(adp_engine *e, const float *p, const float *y, int64_t n_px, float *dldp, double out[4]) {
  return adp_loss_metrics_ex(e, p, y, 1, n_px, 0, 1.f, 0.f, 0.f, dldp, out);
}

int adp_train_begin(adp_engine *e, int batch, int size, float dropout_rate, uint64_t seed) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_CUDA(cudaSetDevice(e->device));
  if (e->split) throw Error(ADP_EINVAL, "the bf16x3 precision is inference-only; train with bf16 or fp32");
  if (e->tr) { sync_host_weights(e); delete e->tr; e->tr = nullptr; }
  e->tmaps.clear();
  train_alloc(e, batch, size, dropout_rate, seed);
  ADP_CATCH
}

int adp_train_forward(adp_engine *e, const float *x, const float *y, int batch, const uint8_t *const *dropout_masks,
                      double *sums) {
  ADP_TRY
  ADP_REQUIRE(e && x && y, "null argument");
  ADP_CUDA(cudaSetDevice(e->device));
  train_forward(e, x, y, batch, dropout_masks, sums);
  ADP_CATCH
}

int adp_train_loss(const double sums[8], double out[4]) {
  if (!sums || !out || !(sums[7] > 0)) return ADP_EINVAL;
  loss_from_sums8(sums, out);
  return ADP_OK;
}

int adp_train_set_loss(adp_engine *e, float ohem_keep_ratio, float eps_pos, float eps_neg) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_REQUIRE(ohem_keep_ratio > 0.f && ohem_keep_ratio <= 1.f && eps_pos >= 0.f && eps_neg >= 0.f && eps_pos + eps_neg < 1.f,
              "loss recipe out of range");
  e->loss.ohem_keep = ohem_keep_ratio; e->loss.eps_pos = eps_pos; e->loss.eps_neg = eps_neg;
  ADP_CATCH
}

int adp_train_set_deep_supervision(adp_engine *e, int on, float w_main, float w_aux1, float w_aux2) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (e->tr) throw Error(ADP_ESTATE, "deep supervision cannot change while a training state is open (adp_train_end first)");
  if (on)
    for (const char *n : kAuxNames)
      if (!e->hw.count(n) || !e->hw[n].set) throw Error(ADP_ESTATE, std::string("weights of layer ") + n + " not set");
  e->deep_sup = on != 0;
  e->ds_w[0] = w_main; e->ds_w[1] = w_aux1; e->ds_w[2] = w_aux2;
  ADP_CATCH
}

int adp_train_outputs(adp_engine *e) { return e && e->deep_sup ? 3 : 1; }

int adp_train_backward(adp_engine *e, const double *sums, int freeze_encoder) {
  ADP_TRY
  ADP_REQUIRE(e, "null argument");
  ADP_CUDA(cudaSetDevice(e->device));
  train_backward(e, sums, freeze_encoder != 0);
  ADP_CATCH
}

int adp_train_sums_buffer(adp_engine *e, double **dev_ptr, int *count) {
  ADP_TRY
  ADP_REQUIRE(e && dev_ptr && count, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  *dev_ptr = e->tr->dsums.as<double>();
  *count = e->deep_sup ? 24 : 8;
  ADP_CATCH
}

int adp_train_sums_read(adp_engine *e, double *host, int count) {
  ADP_TRY
  ADP_REQUIRE(e && host, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_REQUIRE(count == (e->deep_sup ? 24 : 8), "count != 8 x outputs");
  ADP_CUDA(cudaSetDevice(e->device));
  if (e->tr->sums_mirrored) {       // the values the last backward used: wait for their mirror only, not for the whole step
    ADP_CUDA(cudaEventSynchronize(e->tr->ev_sums));
    memcpy(host, e->tr->pinned_sums, (size_t)count * 8);
  } else {
    ADP_CUDA(cudaMemcpyAsync(host, e->tr->dsums.p, (size_t)count * 8, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  ADP_CATCH
}

int adp_train_grad_buffer(adp_engine *e, float **dev_ptr, int64_t *count) {
  ADP_TRY
  ADP_REQUIRE(e && dev_ptr && count, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  *dev_ptr = e->tr->grad.as<float>();
  *count = (int64_t)e->tr->P;
  ADP_CATCH
}

int adp_train_accuracy_read(adp_engine *e, double out[2]) {
  ADP_TRY
  ADP_REQUIRE(e && out, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  if (!e->train_accuracy || !(e->tr->have_forward || e->tr->sums_mirrored)) throw Error(ADP_ESTATE, "set option train_accuracy and run adp_train_forward first");
  ADP_CUDA(cudaSetDevice(e->device));
  if (e->tr->sums_mirrored) {        // after a backward: wait for the mirror only
    ADP_CUDA(cudaEventSynchronize(e->tr->ev_sums));
    out[0] = e->tr->pinned_sums[24]; out[1] = e->tr->pinned_sums[25];
  } else {                           // forward only (validation pass)
    ADP_CUDA(cudaMemcpyAsync(out, e->tr->dsums.as<double>() + 24, 16, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  ADP_CATCH
}

int adp_train_grad_buckets(adp_engine *e, int64_t *lo, int64_t *hi, int cap) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  const int n = (int)e->tr->buckets.size();
  if (lo && hi)
    for (int b = 0; b < n && b < cap; ++b) { lo[b] = (int64_t)e->tr->buckets[b].lo; hi[b] = (int64_t)e->tr->buckets[b].hi; }
  return n;
  ADP_CATCH
}

int adp_train_bucket_wait(adp_engine *e, int bucket, void *stream) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_REQUIRE(bucket >= 0 && bucket < (int)e->tr->buckets.size(), "bucket index");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_CUDA(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), e->tr->buckets[bucket].ev, 0));
  ADP_CATCH
}

int adp_train_join(adp_engine *e, void *stream) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_CUDA(cudaEventRecord(e->tr->ev_join, reinterpret_cast<cudaStream_t>(stream)));
  ADP_CUDA(cudaStreamWaitEvent(e->stream, e->tr->ev_join, 0));
  ADP_CATCH
}

int adp_train_grad_read(adp_engine *e, float *host, int64_t count) {
  ADP_TRY
  ADP_REQUIRE(e && host, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_REQUIRE(count == (int64_t)e->tr->P, "count != parameter count");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_CUDA(cudaMemcpyAsync(host, e->tr->grad.as<float>(), (size_t)count * 4, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_train_grad_write(adp_engine *e, const float *host, int64_t count) {
  ADP_TRY
  ADP_REQUIRE(e && host, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_REQUIRE(count == (int64_t)e->tr->P, "count != parameter count");
  ADP_CUDA(cudaSetDevice(e->device));
  ADP_CUDA(cudaMemcpyAsync(e->tr->grad.as<float>(), host, (size_t)count * 4, cudaMemcpyHostToDevice, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_train_get_grad(adp_engine *e, const char *layer_name, float *kernel, int64_t kernel_elems, float *bias, int64_t nbias) {
  ADP_TRY
  ADP_REQUIRE(e && layer_name, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  size_t off = 0;
  bool found = false;
  for (const std::string &n : param_names(e)) {
    const size_t ke = kernel_elems_of(e, n), be = bias_elems_of(e, n);
    if (n == layer_name) {
      ADP_CUDA(cudaStreamSynchronize(e->stream));
      if (kernel) { ADP_REQUIRE(kernel_elems == (int64_t)ke, "kernel_elems"); ADP_CUDA(cudaMemcpy(kernel, e->tr->grad.as<float>() + off, ke * 4, cudaMemcpyDeviceToHost)); }
      if (bias) { ADP_REQUIRE(nbias == (int64_t)be, "nbias"); ADP_CUDA(cudaMemcpy(bias, e->tr->grad.as<float>() + off + ke, be * 4, cudaMemcpyDeviceToHost)); }
      found = true;
      break;
    }
    off += ke + be;
  }
  if (!found) throw Error(ADP_EINVAL, std::string("unknown layer ") + layer_name);
  ADP_CATCH
}

int adp_train_probs(adp_engine *e, float *out, int64_t out_elems) {
  ADP_TRY
  ADP_REQUIRE(e && out, "null argument");
  if (!e->tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_CUDA(cudaSetDevice(e->device));
  const size_t n = (size_t)e->tr->nb * e->tr->S * e->tr->S;
  ADP_REQUIRE(out_elems == (int64_t)n, "out_elems");
  ADP_CUDA(cudaMemcpyAsync(out, e->tr->prob.p, n * 4, cudaMemcpyDefault, e->stream));
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_train_apply(adp_engine *e, int optimizer, float lr, float grad_scale, double beta1, double beta2, float eps,
                    float weight_decay, int freeze_encoder) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_REQUIRE(optimizer == ADP_OPT_ADAM || optimizer == ADP_OPT_ADAMW, "optimizer");
  ADP_CUDA(cudaSetDevice(e->device));
  train_apply(e, optimizer, lr, grad_scale, beta1 > 0 ? beta1 : 0.9, beta2 > 0 ? beta2 : 0.999, eps > 0 ? eps : 1e-7f, weight_decay,
              freeze_encoder != 0);
  ADP_CATCH
}

int adp_train_step(adp_engine *e, const float *x, const float *y, int batch, int optimizer, float lr, float weight_decay,
                   int freeze_encoder, double out[4]) {
  ADP_TRY
  ADP_REQUIRE(e && x && y, "null argument");
  ADP_REQUIRE(optimizer == ADP_OPT_ADAM || optimizer == ADP_OPT_ADAMW, "optimizer");
  ADP_CUDA(cudaSetDevice(e->device));
  double sums[24];
  train_forward(e, x, y, batch, nullptr, sums);
  if (out) {
    loss_from_sums8(sums, out);
    if (e->deep_sup) {       // Keras total = sum of weighted output losses; bce / dice_loss / dice_coef stay those of main_out
      double a1[4], a2[4];
      loss_from_sums8(sums + 8, a1); loss_from_sums8(sums + 16, a2);
      out[0] = e->ds_w[0] * out[0] + e->ds_w[1] * a1[0] + e->ds_w[2] * a2[0];
    }
  }
  train_backward(e, nullptr, freeze_encoder != 0);
  train_apply(e, optimizer, lr, 1.f, 0.9, 0.999, 1e-7f, weight_decay, freeze_encoder != 0);
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int64_t adp_train_iterations(adp_engine *e) { return (e && e->tr) ? e->tr->iter : -1; }

int adp_train_end(adp_engine *e) {
  ADP_TRY
  ADP_REQUIRE(e, "engine");
  ADP_CUDA(cudaSetDevice(e->device));
  if (e->tr) {
    sync_host_weights(e);
    delete e->tr;
    e->tr = nullptr;
    e->tmaps.clear();
    e->packed = false;        // rebuild the inference operand images from the (updated) host master copy
  }
  ADP_CATCH
}

int adp_adam_update(adp_engine *e, float *theta, const float *grad, float *m, float *v, int64_t n, int64_t t, int optimizer,
                    float lr, double beta1, double beta2, float eps, float weight_decay) {
  ADP_TRY
  ADP_REQUIRE(e && theta && grad && m && v && n > 0 && t >= 1, "null/empty argument");
  ADP_CUDA(cudaSetDevice(e->device));
  DevBuf dth, dg, dm, dv;
  const size_t nb = (size_t)n * 4;
  const bool host = !is_device_ptr(theta);
  float *pth = theta, *pm = m, *pv = v; const float *pg = grad;
  if (host) {
    dth.ensure(nb); dg.ensure(nb); dm.ensure(nb); dv.ensure(nb);
    ADP_CUDA(cudaMemcpyAsync(dth.p, theta, nb, cudaMemcpyHostToDevice, e->stream));
    ADP_CUDA(cudaMemcpyAsync(dg.p, grad, nb, cudaMemcpyHostToDevice, e->stream));
    ADP_CUDA(cudaMemcpyAsync(dm.p, m, nb, cudaMemcpyHostToDevice, e->stream));
    ADP_CUDA(cudaMemcpyAsync(dv.p, v, nb, cudaMemcpyHostToDevice, e->stream));
    pth = dth.as<float>(); pg = dg.as<float>(); pm = dm.as<float>(); pv = dv.as<float>();
  }
  if (beta1 <= 0) beta1 = 0.9;
  if (beta2 <= 0) beta2 = 0.999;
  if (eps <= 0) eps = 1e-7f;
  const float b1p = powf((float)beta1, (float)t), b2p = powf((float)beta2, (float)t);
  const float alpha = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  e->launch("adam_update", 0, (double)n * 28, [&] {
    adam_kernel<<<ew_grid(e, (size_t)n), 256, 0, e->stream>>>(pth, pg, pm, pv, (size_t)n, 1.f, alpha, (float)(1.0 - beta1), (float)(1.0 - beta2), eps,
                                                              optimizer == ADP_OPT_ADAMW ? weight_decay : 0.f, lr);
  });
  if (host) {
    ADP_CUDA(cudaMemcpyAsync(theta, pth, nb, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaMemcpyAsync(m, pm, nb, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaMemcpyAsync(v, pv, nb, cudaMemcpyDeviceToHost, e->stream));
  }
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CATCH
}

int adp_profile_enable(adp_engine *e, int on) { if (!e) return ADP_EINVAL; e->prof = on != 0; return ADP_OK; }
int adp_profile_reset(adp_engine *e) { if (!e) return ADP_EINVAL; e->prof_rows.clear(); return ADP_OK; }
int adp_profile_read(adp_engine *e, adp_prof_row *rows, int cap) {
  if (!e || (!rows && cap > 0)) return ADP_EINVAL;
  int i = 0;
  for (auto &kv : e->prof_rows) {
    if (i >= cap) break;
    memset(&rows[i], 0, sizeof(rows[i]));
    strncpy(rows[i].name, kv.first.c_str(), sizeof(rows[i].name) - 1);
    rows[i].launches = kv.second.launches; rows[i].ms = kv.second.ms; rows[i].flops = kv.second.flops; rows[i].bytes = kv.second.bytes;
    ++i;
  }
  return i;
}
int64_t adp_launch_count(adp_engine *e) { return e ? e->launches : 0; }

}  // extern "C"
