// train_host.cuh — host orchestration of the training step (included by engine.cu).
// Reference: Keras train_step via net.fit (train_adipose_unet_v3.py:1316-1324, 1413-1421),
// combined_loss_standard (:217-241), Adam/AdamW (:801-806), freeze/unfreeze (:760-778).
//
// State lives on the device for the whole training run: a flat fp32 parameter vector theta in
// Keras order (kernel HWIO then bias, layer order of kAllNames == creation order), a flat gradient
// of the same shape (the buffer a data-parallel wrapper all-reduces), Adam moments m and v, every
// forward activation (row-planar, the backward pass needs them) and their gradients.  After each
// optimizer step the operand images the kernels read (padded fp32 / packed bf16 / flipped dgrad
// weights) are rebuilt on the device from theta.
#pragma once

namespace {

struct TrainLayer {
  DevBuf wT;        // dgrad weights [9][cout_pad][cin_pad], taps flipped (fp32; bf16-rounded values off the fp32 path)
  DevBuf gw, gb;    // padded weight / bias gradients [9][cin_pad][cout_pad], [cout_pad]
  size_t koff = 0, boff = 0;   // offsets into the flat vectors
  ConvLayer twin;   // tcgen05 path: the data gradient is the same conv kernel run on dZ with wT (cin <-> cout)
  WgradTcParams wg{};
};

// plan of the tcgen05 weight-gradient kernel for one layer (wgrad_tc.cuh)
void plan_wgrad_tc(adp_engine *e, const ConvLayer &L, WgradTcParams &p) {
  memset(&p, 0, sizeof(p));
  p.dil = L.dil; p.cin_pad = L.cin_pad; p.cout_pad = L.cout_pad;
  p.co_chunk = std::min(48, L.cout_pad);
  p.n_co_chunk = cdiv(L.cout_pad, p.co_chunk);
  p.n_ci_blk = cdiv(L.cin_pad, 120);
  p.cga_box = std::min(15, L.cin_pad / 8);
  const int margin = (L.dil + 7) / 8 * 8;
  p.margin8 = margin / 8; p.PW = 128 + 2 * margin;
  p.a_bytes = (uint32_t)16 * p.PW * 16;                 // always 16 planes: an M = 128 operand spans 16 channel groups
  p.b_row_bytes = (uint32_t)(p.co_chunk / 8) * 2048;
  p.stage_stride = (p.a_bytes + 3 * p.b_row_bytes + 1023) / 1024 * 1024;
  p.S = (int)std::min<size_t>(4, (size_t)(220 * 1024) / p.stage_stride);
  ADP_REQUIRE(p.S >= 2, "wgrad stage does not fit shared memory twice");
  const int ncombo = p.n_ci_blk * p.n_co_chunk;
  p.ctas_per_combo = std::max(1, e->num_sms / ncombo);
}
size_t wgrad_smem_bytes(const WgradTcParams &p) { return (size_t)p.S * p.stage_stride + (2 * p.S + 1) * 8 + 16; }

struct TrainState {
  int nb = 0, S = 0;
  float keep = 1.f;
  uint64_t seed = 0;
  int64_t iter = 0;           // optimizer iterations done
  size_t P = 0;
  DevBuf theta, grad, m, v;
  size_t koff_first = 0, boff_first = 0, koff_head = 0, boff_head = 0, koff_aux[2] = {0, 0}, boff_aux[2] = {0, 0};
  // deep supervision: low-res sigmoid maps, their bilinear upsamplings, gradient scratch, residual tensors for the dgrad epilogues
  DevBuf a_low[2], aux_full[2], g_low, resid[2], g_aux;
  LossState ls_aux[2];
  std::vector<TrainLayer> tl;
  DevBuf gw_first, gb_first, g_head;     // [9][cp0], [cp0], double [cp0 + 1]
  DevBuf zeros;
  LossState ls;
  DevBuf dsums;               // 8 doubles per output (main, aux1, aux2): the loss sums, device-resident
  double *pinned_sums = nullptr;      // host mirror written by an async copy at the start of every backward
  cudaEvent_t ev_sums = nullptr;      // ... and the event that says it has landed (the loss of a step is read without draining the stream)
  bool sums_mirrored = false;
  // gradient buckets for a data-parallel caller: contiguous ranges of the flat gradient in the order the backward pass
  // completes them, each with an event recorded on the engine's stream when its last weight gradient has been written
  struct Bucket { size_t lo = 0, hi = 0; cudaEvent_t ev = nullptr; };
  std::vector<Bucket> buckets;
  cudaEvent_t ev_join = nullptr;
  // second stream of the backward pass: HBM-bound elementwise kernels that are off the critical chain (upsample
  // materialisation, upsample / max-pool backward) run under tensor-bound weight-gradient kernels of the main stream
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fwd = nullptr, ev_mat[3] = {nullptr, nullptr, nullptr}, ev_a = nullptr, ev_b = nullptr;
  DevBuf wg_scratch, wb_scratch;   // per-CTA partial sums of the weight-gradient GEMM (wgrad_tc.cuh), reduced in a fixed order
  DevBuf up_x[3];             // UpSampling2D(x) of the three upsampled convs' inputs (weight-gradient operands)
  bool overlap = true;
  ~TrainState() {
    if (side) cudaStreamDestroy(side);
    for (cudaEvent_t ev : {ev_fwd, ev_mat[0], ev_mat[1], ev_mat[2], ev_a, ev_b}) if (ev) cudaEventDestroy(ev);
    if (pinned_sums) cudaFreeHost(pinned_sums);
    if (ev_sums) cudaEventDestroy(ev_sums);
    if (ev_join) cudaEventDestroy(ev_join);
    for (auto &b : buckets) if (b.ev) cudaEventDestroy(b.ev);
  }
  DevBuf x, y, dldp;
  // activations
  DevBuf d1a, cat1, u1b, u1c, pl1, d2a, cat2, u2b, u2c, pl2, d3a, cat3, u3b, u3c, pl3, t[6], ts, prob;
  // gradients (same shapes); the dilate chain ping-pongs between gt[0] and gt[1]
  DevBuf g_d1a, g_cat1, g_u1b, g_u1c, g_pl1, g_d2a, g_cat2, g_u2b, g_u2c, g_pl2, g_d3a, g_cat3, g_u3b, g_u3c, g_pl3,
      gt[2], g_ts, g_hi;
  Acts acts{};
  bool have_forward = false, have_grads = false;
  bool host_stale = false;    // theta moved on since the host master copy was written
  DevBuf mask_stage[4];
};

// parameter tensors in flat-buffer order: the 22 layers of the graph, then the deep-supervision heads when enabled
std::vector<std::string> param_names(adp_engine *e) {
  std::vector<std::string> v(kAllNames, kAllNames + 22);
  if (e->deep_sup) { v.push_back(kAuxNames[0]); v.push_back(kAuxNames[1]); }
  return v;
}

size_t kernel_elems_of(adp_engine *e, const std::string &n) {
  if (n == "down1_conv1") return (size_t)9 * e->c[0];
  if (n == "output_softmax") return (size_t)2 * e->c[0];
  if (n == "aux_out1") return (size_t)e->c[2];
  if (n == "aux_out2") return (size_t)e->c[1];
  const ConvLayer &L = layer(e, n);
  return (size_t)9 * L.cin * L.cout;
}
size_t bias_elems_of(adp_engine *e, const std::string &n) {
  if (n == "down1_conv1") return e->c[0];
  if (n == "output_softmax") return 2;
  if (n == "aux_out1" || n == "aux_out2") return 1;
  return layer(e, n).cout;
}

// one resident wave of 256-thread blocks for a grid-stride kernel over n items (adp_engine::wave_grid)
template <typename K> int ew_grid(adp_engine *e, size_t n, K kern, size_t smem = 0) { return e->wave_grid(kern, (size_t)cdiv64((long long)n, 256), 256, smem); }
int ew_grid(adp_engine *e, size_t n) { return (int)std::max<size_t>(1, std::min<size_t>(cdiv64(n, 256), (size_t)e->num_sms * 8)); }

// theta -> every operand image the kernels read
void repack_from_theta(adp_engine *e) {
  TrainState *tr = e->tr;
  const float *th = tr->theta.as<float>();
  const bool bf = e->prec != ADP_PREC_FP32;
  const int cp0 = e->cp[0], c0 = e->c[0];
  e->launch("repack_weights", 0, (double)tr->P * 4 * 3, [&] {
    // first conv [9][cp0] and bias (never rounded), head [2][cp0] and bias
    pad_kernel_weights<<<ew_grid(e, 9 * c0), 256, 0, e->stream>>>(th + tr->koff_first, e->w_first.as<float>(), 9, 1, c0, 1, cp0, 0, 0, 0, 0);
    head_pack_kernel<<<1, 256, 0, e->stream>>>(th + tr->koff_head, c0, cp0, e->w_head.as<float>());
  });
  ADP_CUDA(cudaMemcpyAsync(e->b_first.p, th + tr->boff_first, (size_t)c0 * 4, cudaMemcpyDeviceToDevice, e->stream));
  ADP_CUDA(cudaMemcpyAsync(e->b_head.p, th + tr->boff_head, 8, cudaMemcpyDeviceToDevice, e->stream));
  for (size_t i = 0; i < e->layers.size(); ++i) {
    ConvLayer &L = e->layers[i];
    TrainLayer &T = tr->tl[i];
    const size_t ne = (size_t)9 * L.cin * L.cout, np = (size_t)9 * L.cin_pad * L.cout_pad;
    const int sp = L.skip ? pad16(L.skip) : 0;
    const int g = ew_grid(e, ne);
    pad_kernel_weights<<<g, 256, 0, e->stream>>>(th + T.koff, L.w_simt.as<float>(), 9, L.cin, L.cout, L.cin_pad, L.cout_pad, L.skip, sp, 0, 0);
    pad_kernel_weights<<<g, 256, 0, e->stream>>>(th + T.koff, T.wT.as<float>(), 9, L.cin, L.cout, L.cin_pad, L.cout_pad, L.skip, sp, 0, 1);
    if (e->prec == ADP_PREC_BF16) {  // pack from the unrounded padded copies (parity taps are summed in fp32)
      pack_tc_kernel<<<ew_grid(e, np * 2), 256, 0, e->stream>>>(L.w_simt.as<float>(), L.w_tc.as<__nv_bfloat16>(), L.tc.nvar, L.tc.nchunks,
                                                               L.tc.ntaps, L.tc.N, L.cin_pad, L.cout_pad, L.up ? 1 : 0, L.tc.kys);
      const ConvLayer &W = T.twin;
      pack_tc_kernel<<<ew_grid(e, np * 2), 256, 0, e->stream>>>(T.wT.as<float>(), W.w_tc.as<__nv_bfloat16>(), W.tc.nvar, W.tc.nchunks,
                                                               W.tc.ntaps, W.tc.N, W.cin_pad, W.cout_pad, 0, W.tc.kys);
    }
    if (bf) {
      round_bf16_kernel<<<ew_grid(e, np), 256, 0, e->stream>>>(L.w_simt.as<float>(), np);
      round_bf16_kernel<<<ew_grid(e, np), 256, 0, e->stream>>>(T.wT.as<float>(), np);
    }
    ADP_CUDA(cudaMemcpyAsync(L.bias.p, th + T.boff, (size_t)L.cout * 4, cudaMemcpyDeviceToDevice, e->stream));
    e->launches += bf ? 5 : 2;
  }
  ADP_CUDA(cudaGetLastError());
}

// device theta -> host master copies (adp_get_weight, adp_train_end)
void sync_host_weights(adp_engine *e) {
  TrainState *tr = e->tr;
  if (!tr || !tr->host_stale) return;
  std::vector<float> h(tr->P);
  ADP_CUDA(cudaStreamSynchronize(e->stream));
  ADP_CUDA(cudaMemcpy(h.data(), tr->theta.p, tr->P * 4, cudaMemcpyDeviceToHost));
  size_t off = 0;
  for (const std::string &n : param_names(e)) {
    HostWeight &w = e->hw[n];
    const size_t ke = kernel_elems_of(e, n), be = bias_elems_of(e, n);
    w.k.assign(h.begin() + off, h.begin() + off + ke); off += ke;
    w.b.assign(h.begin() + off, h.begin() + off + be); off += be;
  }
  tr->host_stale = false;
}

void train_alloc(adp_engine *e, int nb, int S, float dropout_rate, uint64_t seed) {
  ADP_REQUIRE(nb >= 1 && nb <= 64, "batch must be 1..64");
  ADP_REQUIRE(S >= 64 && S % 64 == 0 && S <= 4096, "training tile size must be a multiple of 64");
  ADP_REQUIRE(dropout_rate >= 0.f && dropout_rate < 1.f, "dropout rate");
  if (!e->packed) pack_all(e);
  std::unique_ptr<TrainState> tr(new TrainState());
  tr->nb = nb; tr->S = S; tr->keep = 1.f - dropout_rate; tr->seed = seed;
  const size_t es = e->esz, n = nb;
  const size_t s1 = (size_t)S * S, s2 = s1 / 4, s3 = s1 / 16, s4 = s1 / 64;
  const int *cp = e->cp;
  auto A = [&](DevBuf &act, DevBuf &g, size_t elems) { act.ensure(elems * es); g.ensure(elems * es); };
  A(tr->d1a, tr->g_d1a, n * s1 * cp[0]); A(tr->cat1, tr->g_cat1, n * s1 * 2 * cp[0]);
  A(tr->u1b, tr->g_u1b, n * s1 * cp[0]); A(tr->u1c, tr->g_u1c, n * s1 * cp[0]); A(tr->pl1, tr->g_pl1, n * s2 * cp[0]);
  A(tr->d2a, tr->g_d2a, n * s2 * cp[1]); A(tr->cat2, tr->g_cat2, n * s2 * 2 * cp[1]);
  A(tr->u2b, tr->g_u2b, n * s2 * cp[1]); A(tr->u2c, tr->g_u2c, n * s2 * cp[1]); A(tr->pl2, tr->g_pl2, n * s3 * cp[1]);
  A(tr->d3a, tr->g_d3a, n * s3 * cp[2]); A(tr->cat3, tr->g_cat3, n * s3 * 2 * cp[2]);
  A(tr->u3b, tr->g_u3b, n * s3 * cp[2]); A(tr->u3c, tr->g_u3c, n * s3 * cp[2]); A(tr->pl3, tr->g_pl3, n * s4 * cp[2]);
  for (int i = 0; i < 6; ++i) tr->t[i].ensure(n * s4 * cp[3] * es);
  tr->ts.ensure(n * s4 * cp[3] * es);
  tr->gt[0].ensure(n * s4 * cp[3] * es); tr->gt[1].ensure(n * s4 * cp[3] * es); tr->g_ts.ensure(n * s4 * cp[3] * es);
  // hi-res data gradient of an UpSampling2D-fed conv before its 2x2 reduction: largest is up1_conv1 (S^2 x cp1)
  tr->g_hi.ensure(n * std::max({s1 * cp[1], s2 * cp[2], s3 * cp[3]}) * es);
  tr->prob.ensure(n * s1 * 4); tr->x.ensure(n * s1 * 4); tr->y.ensure(n * s1 * 4); tr->dldp.ensure(n * s1 * 4);
  tr->zeros.ensure(4096); ADP_CUDA(cudaMemset(tr->zeros.p, 0, 4096));
  tr->dsums.ensure(32 * 8); ADP_CUDA(cudaMemset(tr->dsums.p, 0, 32 * 8));     // 24 loss sums + {matching pixels, pixels} of binary_accuracy
  ADP_CUDA(cudaMallocHost(&tr->pinned_sums, 32 * 8));
  ADP_CUDA(cudaEventCreateWithFlags(&tr->ev_sums, cudaEventDisableTiming));
  ADP_CUDA(cudaEventCreateWithFlags(&tr->ev_join, cudaEventDisableTiming));
  ADP_CUDA(cudaStreamCreateWithFlags(&tr->side, cudaStreamNonBlocking));
  for (cudaEvent_t *ev : {&tr->ev_fwd, &tr->ev_mat[0], &tr->ev_mat[1], &tr->ev_mat[2], &tr->ev_a, &tr->ev_b})
    ADP_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  if (const char *o = getenv("ADP_TRAIN_OVERLAP")) tr->overlap = atoi(o) != 0;
  if (e->prec == ADP_PREC_BF16) {      // operands of the upsampled convs' weight gradients, one buffer per level (written on the side stream)
    tr->up_x[0].ensure(n * s1 * cp[1] * es); tr->up_x[1].ensure(n * s2 * cp[2] * es); tr->up_x[2].ensure(n * s3 * cp[3] * es);
  }
  e->fwt_tile.ensure(64 * 4); e->fwt_op.ensure(64 * 4); e->fwt_origin.ensure(64 * 8);
  Acts &a = tr->acts;
  a.d1a = &tr->d1a; a.cat1 = &tr->cat1; a.u1b = &tr->u1b; a.u1c = &tr->u1c; a.pl1 = &tr->pl1;
  a.d2a = &tr->d2a; a.cat2 = &tr->cat2; a.u2b = &tr->u2b; a.u2c = &tr->u2c; a.pl2 = &tr->pl2;
  a.d3a = &tr->d3a; a.cat3 = &tr->cat3; a.u3b = &tr->u3b; a.u3c = &tr->u3c; a.pl3 = &tr->pl3;
  for (int i = 0; i < 6; ++i) a.t[i] = &tr->t[i];
  a.ts = &tr->ts; a.prob = &tr->prob; a.cap = nb;

  // flat parameter vector
  size_t off = 0;
  tr->tl.resize(e->layers.size());
  std::vector<float> h;
  for (const std::string &n2 : param_names(e)) {
    const size_t ke = kernel_elems_of(e, n2), be = bias_elems_of(e, n2);
    if (n2 == "down1_conv1") { tr->koff_first = off; tr->boff_first = off + ke; }
    else if (n2 == "output_softmax") { tr->koff_head = off; tr->boff_head = off + ke; }
    else if (n2 == "aux_out1") { tr->koff_aux[0] = off; tr->boff_aux[0] = off + ke; }
    else if (n2 == "aux_out2") { tr->koff_aux[1] = off; tr->boff_aux[1] = off + ke; }
    else {
      for (size_t i = 0; i < e->layers.size(); ++i)
        if (e->layers[i].name == n2) { tr->tl[i].koff = off; tr->tl[i].boff = off + ke; }
    }
    const HostWeight &w = e->hw.at(n2);
    h.insert(h.end(), w.k.begin(), w.k.end());
    h.insert(h.end(), w.b.begin(), w.b.end());
    off += ke + be;
  }
  tr->P = off;
  {
    // completion order of the backward pass: decoder + heads, dilate6..4, dilate3..1, encoder
    auto koff_of = [&](const char *nm) { for (size_t i = 0; i < e->layers.size(); ++i) if (e->layers[i].name == nm) return tr->tl[i].koff; return (size_t)0; };
    const size_t o_d1 = koff_of("dilate1"), o_d4 = koff_of("dilate4"), o_u3 = koff_of("up3_conv1");
    const size_t cuts[5] = {off, o_u3, o_d4, o_d1, 0};
    tr->buckets.resize(4);
    for (int b = 0; b < 4; ++b) {
      tr->buckets[b].lo = cuts[b + 1]; tr->buckets[b].hi = cuts[b];
      ADP_CUDA(cudaEventCreateWithFlags(&tr->buckets[b].ev, cudaEventDisableTiming));
    }
  }
  tr->theta.ensure(off * 4); tr->grad.ensure(off * 4); tr->m.ensure(off * 4); tr->v.ensure(off * 4);
  ADP_CUDA(cudaMemcpy(tr->theta.p, h.data(), off * 4, cudaMemcpyHostToDevice));
  ADP_CUDA(cudaMemset(tr->grad.p, 0, off * 4));
  ADP_CUDA(cudaMemset(tr->m.p, 0, off * 4));
  ADP_CUDA(cudaMemset(tr->v.p, 0, off * 4));
  for (size_t i = 0; i < e->layers.size(); ++i) {
    const ConvLayer &L = e->layers[i];
    const size_t np = (size_t)9 * L.cin_pad * L.cout_pad;
    tr->tl[i].wT.ensure(np * 4); tr->tl[i].gw.ensure(np * 4); tr->tl[i].gb.ensure((size_t)L.cout_pad * 4);
    ADP_CUDA(cudaMemset(tr->tl[i].wT.p, 0, np * 4));
    if (e->prec == ADP_PREC_BF16) {
      ConvLayer &W = tr->tl[i].twin;
      W.name = "dgrad/" + L.name; W.cin = L.cout; W.cout = L.cin; W.dil = L.dil; W.up = false; W.skip = 0;
      W.cin_pad = L.cout_pad; W.cout_pad = L.cin_pad;
      plan_tc(W, e->kys);
      W.w_tc.ensure((size_t)W.tc.nvar * W.tc.nchunks * W.tc.ntaps * 16 * W.tc.N * 2);
      W.tc_ready = true;
      plan_wgrad_tc(e, L, tr->tl[i].wg);
    }
  }
  tr->gw_first.ensure((size_t)9 * cp[0] * 4); tr->gb_first.ensure((size_t)cp[0] * 4); tr->g_head.ensure((size_t)(cp[0] + 1) * 8);
  if (e->deep_sup) {
    tr->a_low[0].ensure(n * s3 * 4); tr->a_low[1].ensure(n * s2 * 4);
    tr->aux_full[0].ensure(n * s1 * 4); tr->aux_full[1].ensure(n * s1 * 4);
    tr->g_low.ensure(n * s2 * 4);
    tr->resid[0].ensure(n * s3 * cp[2] * es); tr->resid[1].ensure(n * s2 * cp[1] * es);
    tr->g_aux.ensure((size_t)(cp[2] + 1) * 8);
  }
  delete e->tr;
  e->tr = tr.release();
  repack_from_theta(e);
  ADP_CUDA(cudaStreamSynchronize(e->stream));
}

// ---- forward -----------------------------------------------------------------------------------
void train_forward(adp_engine *e, const float *x, const float *y, int n, const uint8_t *const *masks, double *sums /* 8 per output, or null: stay on the device */) {
  TrainState *tr = e->tr;
  if (!tr) throw Error(ADP_ESTATE, "adp_train_begin not called");
  ADP_REQUIRE(n == tr->nb, "batch size differs from adp_train_begin");
  const int S = tr->S;
  const size_t npx = (size_t)n * S * S;
  ADP_CUDA(cudaMemcpyAsync(tr->x.p, x, npx * 4, cudaMemcpyDefault, e->stream));
  ADP_CUDA(cudaMemcpyAsync(tr->y.p, y, npx * 4, cudaMemcpyDefault, e->stream));
  DropSpec d;
  d.keep = tr->keep;
  d.seed = tr->seed + 0xD1B54A32D192ED03ULL * (uint64_t)(tr->iter + 1);
  if (masks) {
    const int Hs[4] = {S / 8, S / 4, S / 2, S};
    const int cr[4] = {e->c[3], e->c[2], e->c[1], e->c[0]};
    for (int i = 0; i < 4; ++i) {
      ADP_REQUIRE(masks[i], "all four dropout masks must be given");
      const size_t bytes = (size_t)n * Hs[i] * Hs[i] * cr[i];
      d.mask[i] = reinterpret_cast<const uint8_t *>(to_device(e, tr->mask_stage[i], masks[i], bytes));
    }
    if (d.keep >= 1.f) d.keep = 0.7f;      // supplied masks imply the reference's Dropout(0.3)
    tr->keep = d.keep;
  }
  FwTable fw;
  memset(&fw, 0, sizeof(fw));
  for (int i = 0; i < n; ++i) { fw.tile[i] = i; fw.op[i] = 0; }
  FirstConvSrc src{};
  src.f32 = tr->x.as<float>();
  // eval mode (adp_set_option "train_eval_mode"): the validation pass of net.fit - same graph and losses, Dropout inactive
  const bool dropping = (masks || tr->keep < 1.f) && !e->train_eval_mode;
  DropSpec none;   // keep == 1, no masks: the dropout kernels are skipped but the fused head is still off
  if (e->prec == ADP_PREC_FP32) forward_t<float>(e, tr->acts, S, src, fw, n, 0.f, 1.f, nullptr, dropping ? &d : &none);
  else forward_t<__nv_bfloat16>(e, tr->acts, S, src, fw, n, 0.f, 1.f, nullptr, dropping ? &d : &none);
  double *ds = tr->dsums.as<double>();
  loss_forward(e, tr->ls, e->loss, tr->prob.as<float>(), tr->y.as<float>(), n, (size_t)S * S, S, ds);
  if (e->deep_sup) {
    // aux_out1 <- post-dropout up3 (S/4), aux_out2 <- post-dropout up2 (S/2); their losses never use hard mining
    // (train_adipose_unet_v3.py:812-838: loss_fn_aux is the standard or the label-smoothing loss)
    LossRecipe ra = e->loss; ra.ohem_keep = 1.f;
    const int hs[2] = {S / 4, S / 2}, cps[2] = {e->cp[2], e->cp[1]}, crs[2] = {e->c[2], e->c[1]};
    DevBuf *src[2] = {&tr->u3c, &tr->u2c};
    for (int a = 0; a < 2; ++a) {
      const size_t lowpx = (size_t)n * hs[a] * hs[a];
      const float *th = tr->theta.as<float>();
      auto launch_fwd = [&](auto tag) {
        using T = decltype(tag);
        auto v = view<T>(*src[a], hs[a], hs[a], cps[a], 0, cps[a]);
        e->launch("aux_head", 2.0 * lowpx * crs[a], (double)lowpx * (cps[a] * sizeof(T) + 4), [&] {
          aux_head_fwd_kernel<T><<<ew_grid(e, lowpx), 256, (size_t)cps[a] * 4, e->stream>>>(v, n, th + tr->koff_aux[a], th + tr->boff_aux[a], crs[a],
                                                                                          tr->a_low[a].as<float>());
        });
      };
      if (e->prec == ADP_PREC_FP32) launch_fwd(float()); else launch_fwd(__nv_bfloat16());
      e->launch("aux_bilinear_up", 0, (double)npx * 4 + (double)lowpx * 4, [&] {
        bilinear_up_kernel<<<ew_grid(e, npx), 256, 0, e->stream>>>(tr->a_low[a].as<float>(), n, hs[a], S, tr->aux_full[a].as<float>());
      });
      loss_forward(e, tr->ls_aux[a], ra, tr->aux_full[a].as<float>(), tr->y.as<float>(), n, (size_t)S * S, S, ds + 8 * (a + 1));
    }
  }
  if (e->train_accuracy) {      // Keras' binary_accuracy of main_out on this rank's batch (mirrored to the host with the sums)
    const double init[2] = {0.0, (double)npx};
    ADP_CUDA(cudaMemcpyAsync(ds + 24, init, 16, cudaMemcpyHostToDevice, e->stream));
    e->launch("binary_accuracy", 0, (double)npx * 8, [&] {
      binary_accuracy_kernel<<<e->wave_grid(binary_accuracy_kernel, cdiv64(npx, 256 * 4)), 256, 0, e->stream>>>(tr->prob.as<float>(), tr->y.as<float>(), npx, ds + 24);
    });
  }
  if (sums) {
    ADP_CUDA(cudaMemcpyAsync(sums, ds, (size_t)(e->deep_sup ? 24 : 8) * 8, cudaMemcpyDeviceToHost, e->stream));
    ADP_CUDA(cudaStreamSynchronize(e->stream));
  }
  tr->have_forward = true; tr->have_grads = false; tr->sums_mirrored = false;
}

// ---- backward ----------------------------------------------------------------------------------
template <typename T> struct Bwd {
  adp_engine *e;
  TrainState *tr;
  int nb;

  View<T> V(const DevBuf &b, int H, int pitch, int coff, int C) { return view<T>(b, H, H, pitch, coff, C); }

  void relu_mask(View<T> g, View<T> x, float scale) {
    const size_t total = (size_t)nb * g.H * g.W * (g.C / 8);
    e->launch("relu_mask_bwd", 0, (double)total * 8 * sizeof(T) * 3, [&] {
      relu_mask_kernel<T><<<ew_grid(e, total, relu_mask_kernel<T>), 256, 0, e->stream>>>(g, x, nb, scale);
    });
  }
  void add(View<T> dst, View<T> a, View<T> b) {
    const size_t total = (size_t)nb * dst.H * dst.W * (dst.C / 8);
    e->launch("grad_add", 0, (double)total * 8 * sizeof(T) * 3, [&] {
      add_views_kernel<T><<<ew_grid(e, total, add_views_kernel<T>), 256, 0, e->stream>>>(dst, a, b, nb);
    });
  }
  // side != 0: enqueue on the side stream (TrainState::side) - the caller orders it with events
  cudaStream_t sel(bool side) const { return (side && tr->overlap && !e->prof) ? tr->side : e->stream; }
  void after(cudaEvent_t ev, bool on_side) {        // make the chosen stream wait for an event recorded on the other one
    if (tr->overlap && !e->prof) ADP_CUDA(cudaStreamWaitEvent(on_side ? tr->side : e->stream, ev, 0));
  }
  void mark(cudaEvent_t ev, bool on_side) {
    if (tr->overlap && !e->prof) ADP_CUDA(cudaEventRecord(ev, on_side ? tr->side : e->stream));
  }
  void pool_bwd(View<T> xin, View<T> gout, View<T> gin, bool side = false) {
    const size_t total = (size_t)nb * gout.H * gout.W * (gout.C / 8);
    cudaStream_t st = sel(side);
    e->launch("maxpool2x2_bwd", 0, (double)total * 8 * sizeof(T) * 13, [&] {
      maxpool2_bwd_kernel<T><<<ew_grid(e, total, maxpool2_bwd_kernel<T>), 256, 0, st>>>(xin, gout, gin, nb);
    });
  }
  // UpSampling2D(x) written out once per step (operand of the upsampled conv's weight gradient)
  void materialise(int level, View<T> xin, int cin_pad) {
    View<T> xs = V(tr->up_x[level], xin.H * 2, cin_pad, 0, cin_pad);
    const size_t total = (size_t)nb * xs.H * xs.W * (xs.C / 8);
    cudaStream_t st = sel(true);
    e->launch("upsample2x2_materialise", 0, (double)total * 16 * 1.25, [&] {
      upsample2_kernel<T><<<ew_grid(e, total / 4, upsample2_kernel<T>), 256, 0, st>>>(xin, xs, nb);
    });
  }

  // weight + bias gradient of layer li: xin = the conv's input (low-res source for an upsampled conv), dz = dL/d(pre-activation)
  void wgrad(size_t li, View<T> xin, View<T> dz) {
    ConvLayer &L = e->layers[li];
    TrainLayer &TL = tr->tl[li];
    const size_t np = (size_t)9 * L.cin_pad * L.cout_pad;
    const bool tc_path = e->prec == ADP_PREC_BF16 && !e->wgrad_simt && sizeof(T) == 2;
    if (!tc_path) {      // the CUDA-core kernel accumulates into the gradient; the tcgen05 path writes slots and reduces
      ADP_CUDA(cudaMemsetAsync(TL.gw.p, 0, np * 4, e->stream));
      ADP_CUDA(cudaMemsetAsync(TL.gb.p, 0, (size_t)L.cout_pad * 4, e->stream));
    }
    const double fl = conv_flops(L, dz.H, dz.W, nb);
    if (e->prec == ADP_PREC_BF16 && !e->wgrad_simt) {
      if constexpr (sizeof(T) == 2) {
        View<T> xs = xin;
        if (L.up) {     // UpSampling2D(x), materialised on the side stream at the start of the backward pass: the weight gradient
                        // is then a plain 9-tap reduction at high resolution
          const int level = dz.H == tr->S ? 0 : (dz.H == tr->S / 2 ? 1 : 2);
          xs = V(tr->up_x[level], dz.H, L.cin_pad, 0, L.cin_pad);
          after(tr->ev_mat[level], false);
        }
        WgradTcParams p = TL.wg;
        // deterministic accumulation: every CTA writes the partial sums of its unit range into its own slot, then the slots
        // are added in a fixed order (wgrad_reduce_kernel) - the gradient is bit-reproducible from run to run
        tr->wg_scratch.ensure((size_t)p.ctas_per_combo * np * 4); tr->wb_scratch.ensure((size_t)p.ctas_per_combo * L.cout_pad * 4);
        p.nb = nb; p.H = dz.H; p.W = dz.W; p.dW = tr->wg_scratch.as<float>(); p.db = tr->wb_scratch.as<float>();
        const CUtensorMap &tmx = tmap_for(e, xs.p, xs.H, xs.W, xs.cgs, xs.cg0, xs.C, nb, p.PW / 8, p.cga_box, 1);
        const CUtensorMap &tmz = tmap_for(e, dz.p, dz.H, dz.W, dz.cgs, dz.cg0, dz.C, nb, 16, p.co_chunk / 8, 1);
        const int grid = p.n_ci_blk * p.n_co_chunk * p.ctas_per_combo;
        const double by = (double)nb * dz.H * dz.W * (L.cin_pad + L.cout_pad) * 2.0;
        e->launch(("conv_wgrad_tcgen05/" + L.name).c_str(), fl, by, [&] {
          wgrad_tc_kernel<<<grid, kWgThreads, wgrad_smem_bytes(p), e->stream>>>(tmx, tmz, p);
        });
        const size_t nflat = (size_t)9 * L.cin * L.cout + L.cout;
        e->launch("wgrad_reduce", 0, (double)p.ctas_per_combo * np * 4 + (double)nflat * 4, [&] {
          wgrad_reduce_kernel<<<ew_grid(e, nflat, wgrad_reduce_kernel), 256, 0, e->stream>>>(
              tr->wg_scratch.as<float>(), tr->wb_scratch.as<float>(), p.ctas_per_combo, L.cin, L.cout, L.cin_pad, L.cout_pad, L.skip,
              L.skip ? pad16(L.skip) : 0, tr->grad.as<float>() + TL.koff, tr->grad.as<float>() + TL.boff);
        });
        return;          // the flat gradient is complete: no padded copy, no un-padding pass
      }
    } else {
    const int ci_tiles = cdiv(L.cin_pad, 64), co_tiles = cdiv(L.cout_pad, 64);
    const long long nchunks = (long long)nb * dz.H * cdiv(dz.W, 32);
    const int splits = (int)std::max<long long>(1, std::min<long long>(nchunks, std::max(8, e->num_sms * 8 / (9 * ci_tiles * co_tiles))));
    dim3 grid(9, ci_tiles * co_tiles, splits);
    e->launch(("conv_wgrad_simt/" + L.name).c_str(), fl, 0, [&] {
      if (L.up) conv_wgrad_kernel<T, true><<<grid, 256, 0, e->stream>>>(xin, dz, TL.gw.as<float>(), TL.gb.as<float>(), L.dil, nb, L.cin_pad, L.cout_pad, co_tiles);
      else conv_wgrad_kernel<T, false><<<grid, 256, 0, e->stream>>>(xin, dz, TL.gw.as<float>(), TL.gb.as<float>(), L.dil, nb, L.cin_pad, L.cout_pad, co_tiles);
    });
    }
    const size_t ne = (size_t)9 * L.cin * L.cout;
    const int sp = L.skip ? pad16(L.skip) : 0;
    e->launch("grad_unpad", 0, (double)ne * 8, [&] {
      pad_kernel_weights<<<ew_grid(e, ne), 256, 0, e->stream>>>(tr->grad.as<float>() + TL.koff, TL.gw.as<float>(), 9, L.cin, L.cout, L.cin_pad,
                                                               L.cout_pad, L.skip, sp, 1, 0);
    });
    ADP_CUDA(cudaMemcpyAsync(tr->grad.as<float>() + TL.boff, TL.gb.p, (size_t)L.cout * 4, cudaMemcpyDeviceToDevice, e->stream));
  }

  // data gradient of layer li: dz (output resolution) -> gin (input resolution):
  //   gin = (conv(dz, flipped W) [2x2-summed for an upsampled conv] + resid) * [mask > 0] * scale
  // mask = the forward tensor gin is the gradient of (ReLU', and the dropout mask when scale = 1/keep): the written
  // tensor is dL/d(pre-activation) of the producing layer.  On the tcgen05 path both live in the conv epilogue.
  // up_on_side: the 2x2 reduction of an upsampled conv's data gradient goes to the side stream (the caller overlaps it with
  // the same layer's weight gradient and waits for ev_b before the reduced gradient is read)
  void dgrad(size_t li, View<T> dz, View<T> gin, const View<T> *mask = nullptr, float scale = 1.f, const View<T> *resid = nullptr,
             bool up_on_side = false) {
    ConvLayer &L = e->layers[li];
    TrainLayer &TL = tr->tl[li];
    View<T> out = gin;
    if (L.up) out = V(tr->g_hi, dz.H, L.cin_pad, 0, L.cin_pad);
    dim3 grid(cdiv(dz.W, 32), cdiv(dz.H, 16), nb * (L.cin_pad / 16)), block(32, 8);
    const double fl = conv_flops(L, dz.H, dz.W, nb);
    bool fused = false;
    if (e->prec == ADP_PREC_BF16 && !e->dgrad_simt) {
      const double by = (double)nb * dz.H * dz.W * (L.cin_pad + L.cout_pad) * 2.0;
      EpiSpec epi;
      if (L.up && e->fuse_upsum && TL.twin.tc.T == 2 && (!mask || (mask->cgs == gin.cgs && mask->cg0 == gin.cg0 && mask->H == gin.H)) &&
          (!resid || (resid->cgs == gin.cgs && resid->cg0 == gin.cg0 && resid->H == gin.H))) {
        // UpSampling2D's backward (2x2 sum, second gradient, ReLU'/dropout mask) in the twin's epilogue: the full-resolution
        // gradient never reaches HBM and upsample2_bwd_kernel is not launched
        epi.mode = EPI_UPSUM;
        epi.up_dst = gin.p; epi.up_cgs = gin.cgs; epi.up_cg0 = gin.cg0;
        epi.mask = mask ? mask->p : nullptr; epi.mask_scale = scale; epi.resid = resid ? resid->p : nullptr;
        launch_conv_tc(e, TL.twin, "conv_dgrad_tcgen05/" + L.name, fl, (double)nb * dz.H * dz.W * (L.cin_pad / 4 + L.cout_pad) * 2.0, dz.p, dz.H, dz.W,
                       dz.cgs, dz.cg0, out.p, out.cgs, out.cg0, nb, nb, epi, tr->zeros.as<float>(), 0);
        return;
      }
      if (!L.up) {
        ADP_REQUIRE(!mask || (mask->cgs == gin.cgs && mask->cg0 == gin.cg0 && mask->H == gin.H), "mask layout must equal the gradient layout");
        ADP_REQUIRE(!resid || (resid->cgs == gin.cgs && resid->cg0 == gin.cg0 && resid->H == gin.H), "residual layout must equal the gradient layout");
        epi.mask = mask ? mask->p : nullptr; epi.mask_scale = scale; epi.resid = resid ? resid->p : nullptr;
        fused = true;
      }
      launch_conv_tc(e, TL.twin, "conv_dgrad_tcgen05/" + L.name, fl, by, dz.p, dz.H, dz.W, dz.cgs, dz.cg0, out.p, out.cgs, out.cg0, nb, nb,
                     epi, tr->zeros.as<float>(), 0);
    } else {
      e->launch(("conv_dgrad_simt/" + L.name).c_str(), fl, 0, [&] {
        conv3x3_simt_kernel<T, false><<<grid, block, 0, e->stream>>>(dz, out, TL.wT.as<float>(), tr->zeros.as<float>(), L.dil, 0);
      });
    }
    if (L.up) {
      const size_t total = (size_t)nb * gin.H * gin.W * (gin.C / 8);
      View<T> none{}; none.p = nullptr;
      if (up_on_side) { mark(tr->ev_a, false); after(tr->ev_a, true); }
      cudaStream_t st = sel(up_on_side);
      e->launch("upsample2x2_bwd", 0, (double)total * 8 * sizeof(T) * (mask ? 6 : 5), [&] {
        upsample2_bwd_kernel<T><<<ew_grid(e, total, upsample2_bwd_kernel<T>), 256, 0, st>>>(out, gin, nb, mask ? *mask : none, scale, resid ? *resid : none);
      });
      if (up_on_side) mark(tr->ev_b, true);
    } else if (!fused) {
      if (resid) add(gin, gin, *resid);
      if (mask) relu_mask(gin, *mask, scale);
    }
  }
};

size_t layer_index(adp_engine *e, const char *n) {
  for (size_t i = 0; i < e->layers.size(); ++i)
    if (e->layers[i].name == n) return i;
  throw Error(ADP_EINVAL, std::string("unknown layer ") + n);
}

template <typename T> void backward_t(adp_engine *e, bool freeze_encoder) {
  TrainState *tr = e->tr;
  Bwd<T> B{e, tr, tr->nb};
  const int nb = tr->nb, S = tr->S, S2 = S / 2, S3 = S / 4, S4 = S / 8;
  const int *cp = e->cp;
  const float inv_keep = 1.f / tr->keep;
  auto li = [&](const char *n) { return layer_index(e, n); };
  auto bucket_done = [&](int b) { ADP_CUDA(cudaEventRecord(tr->buckets[b].ev, e->stream)); };
  const bool tcw = e->prec == ADP_PREC_BF16 && !e->wgrad_simt && sizeof(T) == 2;
  if (tcw) {
    // operands of the three upsampled convs' weight gradients: forward activations only, so they are written on the side
    // stream while the head / up1 gradients run (ev_fwd orders them behind everything enqueued so far)
    B.mark(tr->ev_fwd, false); B.after(tr->ev_fwd, true);
    B.materialise(0, B.V(tr->u2c, S2, cp[1], 0, cp[1]), cp[1]); B.mark(tr->ev_mat[0], true);
    B.materialise(1, B.V(tr->u3c, S3, cp[2], 0, cp[2]), cp[2]); B.mark(tr->ev_mat[1], true);
    B.materialise(2, B.V(tr->ts, S4, cp[3], 0, cp[3]), cp[3]); B.mark(tr->ev_mat[2], true);
  }

  // head: dL/dp -> dL/d(pre-activation of up1_conv3) (ReLU' and dropout folded in), head weight gradients
  {
    auto x = B.V(tr->u1c, S, cp[0], 0, cp[0]);
    auto g = B.V(tr->g_u1c, S, cp[0], 0, cp[0]);
    ADP_CUDA(cudaMemsetAsync(tr->g_head.p, 0, (size_t)(cp[0] + 1) * 8, e->stream));
    const size_t total = (size_t)nb * S * S;
    // one resident wave, rounded to a multiple of the channel-group count (a warp keeps one group: gridDim.x * 8 % G == 0)
    const int G0 = cp[0] / 8;
    const size_t hsm = (size_t)(2 * cp[0] + 1) * 4;
    const int hgrid = std::max(1, e->wave_grid(head_bwd_kernel<T>, (size_t)cdiv64((long long)nb * S * G0, 8), 256, hsm) / G0) * G0;
    e->launch("head_bwd", 0, (double)total * (8 + 2.0 * cp[0] * sizeof(T)), [&] {
      head_bwd_kernel<T><<<hgrid, 256, hsm, e->stream>>>(
          x, nb, e->w_head.as<float>(), tr->prob.as<float>(), tr->dldp.as<float>(), g, tr->g_head.as<double>(), tr->g_head.as<double>() + cp[0],
          inv_keep);
    });
    e->launch("head_grad_finish", 0, 0, [&] {
      head_grad_finish_kernel<<<1, 256, 0, e->stream>>>(tr->g_head.as<double>(), e->c[0], cp[0], tr->grad.as<float>() + tr->koff_head,
                                                       tr->grad.as<float>() + tr->boff_head);
    });
  }
  // Every g_* tensor below holds dL/d(pre-activation) of the layer that produced the matching activation.
  struct Lvl { int H, c; DevBuf *cat, *g_cat, *ub, *g_ub, *uc, *g_uc, *da, *g_da, *pl, *g_pl; const char *c1, *c2, *c3, *d1, *d2; };
  const Lvl lv[3] = {
      {S, cp[0], &tr->cat1, &tr->g_cat1, &tr->u1b, &tr->g_u1b, &tr->u1c, &tr->g_u1c, &tr->d1a, &tr->g_d1a, &tr->pl1, &tr->g_pl1,
       "up1_conv1", "up1_conv2", "up1_conv3", "down1_conv1", "down1_conv2"},
      {S2, cp[1], &tr->cat2, &tr->g_cat2, &tr->u2b, &tr->g_u2b, &tr->u2c, &tr->g_u2c, &tr->d2a, &tr->g_d2a, &tr->pl2, &tr->g_pl2,
       "up2_conv1", "up2_conv2", "up2_conv3", "down2_conv1", "down2_conv2"},
      {S3, cp[2], &tr->cat3, &tr->g_cat3, &tr->u3b, &tr->g_u3b, &tr->u3c, &tr->g_u3c, &tr->d3a, &tr->g_d3a, &tr->pl3, &tr->g_pl3,
       "up3_conv1", "up3_conv2", "up3_conv3", "down3_conv1", "down3_conv2"}};
  // deep supervision: gradients of the two auxiliary outputs -> head weight gradients + residual tensors that the data
  // gradient into up3 / up2 adds before the ReLU'/dropout mask (both tensors are read post-dropout by their heads)
  if (e->deep_sup) {
    LossRecipe ra = e->loss; ra.ohem_keep = 1.f;
    const int hs[2] = {S3, S2}, cps[2] = {cp[2], cp[1]}, crs[2] = {e->c[2], e->c[1]};
    DevBuf *src[2] = {&tr->u3c, &tr->u2c};
    const size_t npx = (size_t)nb * S * S;
    for (int a = 0; a < 2; ++a) {
      const size_t lowpx = (size_t)nb * hs[a] * hs[a];
      loss_backward(e, tr->ls_aux[a], ra, tr->aux_full[a].as<float>(), tr->y.as<float>(), nb, (size_t)S * S, S,
                    tr->dsums.as<double>() + 8 * (a + 1), tr->dldp.as<float>(), e->ds_w[a + 1]);
      e->launch("aux_bilinear_up_bwd", 0, (double)npx * 4 + (double)lowpx * 4, [&] {
        bilinear_up_bwd_kernel<<<ew_grid(e, lowpx), 256, 0, e->stream>>>(tr->dldp.as<float>(), nb, hs[a], S, 1.0f, tr->g_low.as<float>());
      });
      ADP_CUDA(cudaMemsetAsync(tr->g_aux.p, 0, (size_t)(cps[a] + 1) * 8, e->stream));
      auto x = B.V(*src[a], hs[a], cps[a], 0, cps[a]);
      auto r = B.V(tr->resid[a], hs[a], cps[a], 0, cps[a]);
      const float *th = tr->theta.as<float>();
      e->launch("aux_head_bwd", 4.0 * lowpx * crs[a], (double)lowpx * (2.0 * cps[a] * sizeof(T) + 8), [&] {
        aux_head_bwd_kernel<T><<<(int)cdiv64(lowpx, 256), 256, (size_t)(2 * cps[a] + 1) * 4, e->stream>>>(
            x, nb, th + tr->koff_aux[a], crs[a], tr->a_low[a].as<float>(), tr->g_low.as<float>(), r, tr->g_aux.as<double>());
      });
      e->launch("aux_grad_finish", 0, 0, [&] {
        aux_grad_finish_kernel<<<1, 256, 0, e->stream>>>(tr->g_aux.as<double>(), crs[a], cps[a], tr->grad.as<float>() + tr->koff_aux[a],
                                                        tr->grad.as<float>() + tr->boff_aux[a]);
      });
    }
  }
  // decoder, levels 1..3
  for (int l = 0; l < 3; ++l) {
    const Lvl &q = lv[l];
    const int H = q.H, c = q.c;
    auto ub = B.V(*q.ub, H, c, 0, c), g_ub = B.V(*q.g_ub, H, c, 0, c), g_uc = B.V(*q.g_uc, H, c, 0, c);
    auto cat = B.V(*q.cat, H, 2 * c, 0, 2 * c), g_cat = B.V(*q.g_cat, H, 2 * c, 0, 2 * c);
    auto g_cat_up = B.V(*q.g_cat, H, 2 * c, c, c);
    B.wgrad(li(q.c3), ub, g_uc);
    B.dgrad(li(q.c3), g_uc, g_ub, &ub, 1.f);
    B.wgrad(li(q.c2), cat, g_ub);
    B.dgrad(li(q.c2), g_ub, g_cat, &cat, 1.f);          // both halves: [skip | up]
    if (l < 2) {       // input of up{l}_conv1 = post-dropout up{l+1}_conv3 at half resolution
      const Lvl &n = lv[l + 1];
      auto src = B.V(*n.uc, n.H, n.c, 0, n.c), g_src = B.V(*n.g_uc, n.H, n.c, 0, n.c);
      auto aux_r = B.V(tr->resid[1 - l], n.H, n.c, 0, n.c);      // l = 0 feeds up2 (aux_out2), l = 1 feeds up3 (aux_out1)
      // data gradient first: its 2x2 reduction (HBM-bound) runs on the side stream under the weight gradient (tensor-bound)
      B.dgrad(li(q.c1), g_cat_up, g_src, &src, inv_keep, e->deep_sup ? &aux_r : nullptr, true);
      B.wgrad(li(q.c1), src, g_cat_up);
      B.after(tr->ev_b, false);
    } else {           // up3_conv1 reads the Add of the six bottleneck tensors (no activation of its own)
      auto ts = B.V(tr->ts, S4, cp[3], 0, cp[3]), g_ts = B.V(tr->g_ts, S4, cp[3], 0, cp[3]);
      B.dgrad(li(q.c1), g_cat_up, g_ts, nullptr, 1.f, nullptr, true);
      B.wgrad(li(q.c1), ts, g_cat_up);
      B.after(tr->ev_b, false);
    }
  }
  bucket_done(0);
  // bottleneck: Add fans g_ts out to the six dilate outputs; the chain adds the downstream conv's data gradient
  const char *dn[6] = {"dilate1", "dilate2", "dilate3", "dilate4", "dilate5", "dilate6"};
  auto VT = [&](const DevBuf &b) { return B.V(b, S4, cp[3], 0, cp[3]); };
  ADP_CUDA(cudaMemcpyAsync(tr->gt[1].p, tr->g_ts.p, (size_t)nb * S4 * S4 * cp[3] * sizeof(T), cudaMemcpyDeviceToDevice, e->stream));
  B.relu_mask(VT(tr->gt[1]), VT(tr->t[5]), 1.f);
  int cur = 1;     // gt[cur] = dL/d(pre-activation of dilate{i+1})
  for (int i = 5; i >= 1; --i) {
    B.wgrad(li(dn[i]), VT(tr->t[i - 1]), VT(tr->gt[cur]));
    auto m = VT(tr->t[i - 1]), r = VT(tr->g_ts);
    B.dgrad(li(dn[i]), VT(tr->gt[cur]), VT(tr->gt[cur ^ 1]), &m, i == 1 ? inv_keep : 1.f, &r);
    cur ^= 1;
    if (i == 3) bucket_done(1);
  }
  // From here on every max-pool backward (HBM-bound) runs on the side stream under a weight gradient of the main stream:
  // the data gradient that feeds it is issued first, the weight gradient of the same layer (independent of it) second.
  auto pool_on_side = [&](int l) {
    const Lvl &q = lv[l];
    B.mark(tr->ev_a, false); B.after(tr->ev_a, true);
    B.pool_bwd(B.V(*q.cat, q.H, 2 * q.c, 0, q.c), B.V(*q.g_pl, q.H / 2, q.c, 0, q.c), B.V(*q.g_cat, q.H, 2 * q.c, 0, q.c), true);
    B.mark(tr->ev_b, true);
  };
  if (!freeze_encoder) {
    B.dgrad(li("dilate1"), VT(tr->gt[cur]), B.V(*lv[2].g_pl, lv[2].H / 2, lv[2].c, 0, lv[2].c));
    pool_on_side(2);
  }
  B.wgrad(li("dilate1"), B.V(tr->pl3, S4, cp[2], 0, cp[2]), VT(tr->gt[cur]));
  bucket_done(2);
  if (freeze_encoder) { bucket_done(3); return; }     // phase 1: nothing upstream is trainable (train_adipose_unet_v3.py:760-769)
  // encoder, levels 3..1
  for (int l = 2; l >= 0; --l) {
    const Lvl &q = lv[l];
    const int H = q.H, c = q.c;
    auto g_skip = B.V(*q.g_cat, H, 2 * c, 0, c);
    auto da = B.V(*q.da, H, c, 0, c), g_da = B.V(*q.g_da, H, c, 0, c);
    B.after(tr->ev_b, false);                          // max-pool backward of this level has written g_skip
    B.wgrad(li(q.d2), da, g_skip);
    B.dgrad(li(q.d2), g_skip, g_da, &da, 1.f);
    if (l > 0) {     // down{l}_conv1 reads the pooled output of the level above (down1_conv1: first-layer kernel below)
      const Lvl &u = lv[l - 1];
      B.dgrad(li(q.d1), g_da, B.V(*u.g_pl, u.H / 2, u.c, 0, u.c));
      pool_on_side(l - 1);
      B.wgrad(li(q.d1), B.V(*u.pl, H, u.c, 0, u.c), g_da);
    }
  }
  // first conv (Cin = 1): weight gradient only
  {
    ADP_CUDA(cudaMemsetAsync(tr->gw_first.p, 0, (size_t)9 * cp[0] * 4, e->stream));
    ADP_CUDA(cudaMemsetAsync(tr->gb_first.p, 0, (size_t)cp[0] * 4, e->stream));
    const size_t total = (size_t)nb * S * S;
    const int G0 = cp[0] / 8;
    const int grid = G0 * std::max(1, std::min(e->num_sms * 2 / G0, (int)cdiv64((long long)nb * S * cdiv(S, 32), 8)));
    bool on_tensor_cores = false;
    if constexpr (sizeof(T) == 2) on_tensor_cores = e->prec == ADP_PREC_BF16 && !e->wgrad_simt;
    e->launch("first_conv_wgrad", 2.0 * total * 9 * e->c[0], (double)total * (4 + cp[0] * sizeof(T)), [&] {
      if constexpr (sizeof(T) == 2) {
        if (on_tensor_cores) {     // nine taps + ones row as GEMM-M of mma.sync (kernels_train.cuh); one resident wave
          const size_t msm = first_wgrad_mma_smem(G0);
          ADP_CUDA(cudaFuncSetAttribute(first_wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)first_wgrad_mma_smem(kFwMaxG)));
          const int mgrid = e->wave_grid(first_wgrad_mma_kernel, (size_t)cdiv64((long long)nb * S * cdiv(S, 32), 8), 256, msm);
          first_wgrad_mma_kernel<<<mgrid, 256, msm, e->stream>>>(tr->x.as<float>(), B.V(tr->g_d1a, S, cp[0], 0, cp[0]), nb,
                                                                tr->gw_first.as<float>(), tr->gb_first.as<float>());
          return;
        }
      }
      first_wgrad_kernel<T><<<grid, 256, 0, e->stream>>>(tr->x.as<float>(), B.V(tr->g_d1a, S, cp[0], 0, cp[0]), nb,
                                                                              tr->gw_first.as<float>(), tr->gb_first.as<float>());
    });
    e->launch("grad_unpad", 0, 0, [&] {
      pad_kernel_weights<<<ew_grid(e, 9 * e->c[0]), 256, 0, e->stream>>>(tr->grad.as<float>() + tr->koff_first, tr->gw_first.as<float>(), 9, 1,
                                                                         e->c[0], 1, cp[0], 0, 0, 1, 0);
    });
    ADP_CUDA(cudaMemcpyAsync(tr->grad.as<float>() + tr->boff_first, tr->gb_first.p, (size_t)e->c[0] * 4, cudaMemcpyDeviceToDevice, e->stream));
  }
  bucket_done(3);
}

void train_backward(adp_engine *e, const double *sums /* 8 per output (host), or null: the device sums as they stand */, bool freeze_encoder) {
  TrainState *tr = e->tr;
  if (!tr || !tr->have_forward) throw Error(ADP_ESTATE, "adp_train_forward must run first");
  const size_t sbytes = (size_t)(e->deep_sup ? 24 : 8) * 8;
  if (sums) ADP_CUDA(cudaMemcpyAsync(tr->dsums.p, sums, sbytes, cudaMemcpyHostToDevice, e->stream));
  ADP_CUDA(cudaMemcpyAsync(tr->pinned_sums, tr->dsums.p, 32 * 8, cudaMemcpyDeviceToHost, e->stream));
  ADP_CUDA(cudaEventRecord(tr->ev_sums, e->stream));
  tr->sums_mirrored = true;
  loss_backward(e, tr->ls, e->loss, tr->prob.as<float>(), tr->y.as<float>(), tr->nb, (size_t)tr->S * tr->S, tr->S, tr->dsums.as<double>(),
                tr->dldp.as<float>(), e->deep_sup ? e->ds_w[0] : 1.f);
  if (freeze_encoder) {   // frozen tensors report zero gradient
    const size_t first_trainable = tr->tl[layer_index(e, "dilate1")].koff;
    ADP_CUDA(cudaMemsetAsync(tr->grad.p, 0, first_trainable * 4, e->stream));
  }
  if (e->prec == ADP_PREC_FP32) backward_t<float>(e, freeze_encoder);
  else backward_t<__nv_bfloat16>(e, freeze_encoder);
  tr->have_grads = true;       // stream-ordered: no host synchronisation inside a step (readers of the gradient synchronise)
}

// Keras Adam / AdamW (train_adipose_unet_v3.py:801-806; epsilon outside the bias correction)
void train_apply(adp_engine *e, int optimizer, float lr, float grad_scale, double beta1, double beta2, float eps, float weight_decay,
                 bool freeze_encoder) {
  TrainState *tr = e->tr;
  if (!tr || !tr->have_grads) throw Error(ADP_ESTATE, "adp_train_backward must run first");
  const int64_t t = tr->iter + 1;
  // alpha in float32 like Keras: lr * sqrt(1 - b2^t) / (1 - b1^t)
  const float b1p = powf((float)beta1, (float)t), b2p = powf((float)beta2, (float)t);
  const float alpha = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  const size_t first = freeze_encoder ? tr->tl[layer_index(e, "dilate1")].koff : 0;
  const size_t n = tr->P - first;
  const float wd = optimizer == ADP_OPT_ADAMW ? weight_decay : 0.f;
  e->launch("adam_update", 0, (double)n * 28, [&] {
    adam_kernel<<<ew_grid(e, n, adam_kernel), 256, 0, e->stream>>>(tr->theta.as<float>() + first, tr->grad.as<float>() + first, tr->m.as<float>() + first,
                                                      tr->v.as<float>() + first, n, grad_scale, alpha, (float)(1.0 - beta1), (float)(1.0 - beta2), eps, wd, lr);
  });
  tr->iter = t;
  tr->host_stale = true;
  repack_from_theta(e);
  tr->have_grads = false; tr->have_forward = false;
}

}  // namespace
