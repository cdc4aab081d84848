/*
 * adipose_b200.h — C ABI of libadipose_b200.so
 *
 * B200-native (sm_100a) replacement for the U-Net segmentation hot path of
 * MAGIC-SCAN/adipose_tissue-unet.  The reference has no FFI: its seam is a
 * duck-typed Python object.  Each entry point below names
 * the reference interface it replaces (paths relative to the reference root);
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: every function returns 0 on success and a negative ADP_E* code on
 * failure (adp_last_error() gives the message; no exception crosses the
 * boundary).  Buffers are caller-owned.  Every data pointer may be a host
 * pointer (pageable or pinned) or a CUDA device pointer on the engine's device;
 * the library detects which.  One CUDA stream per engine; an engine must not be
 * used from two threads at once.  All calls are synchronous on return unless
 * stated otherwise.  There is NO CPU fallback: adp_create fails without a
 * compute-capability-10.x device.
 */
#ifndef ADIPOSE_B200_H
#define ADIPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADP_ABI_VERSION 3

/* error codes */
#define ADP_OK 0
#define ADP_EINVAL -1   /* bad argument */
#define ADP_ECUDA -2    /* CUDA runtime / driver error */
#define ADP_ENODEV -3   /* no sm_100 device */
#define ADP_ESTATE -4   /* call out of order (e.g. predict before weights are loaded) */
#define ADP_ENOMEM -5

/* arithmetic of the convolution path */
#define ADP_PREC_FP32 0       /* fp32 activations, CUDA-core FFMA convs (exact-fp32 parity path) */
#define ADP_PREC_BF16 1       /* bf16 activations, tcgen05/TMEM implicit-GEMM convs, fp32 accumulate */
#define ADP_PREC_BF16_SIMT 2  /* bf16 activations, CUDA-core convs (cross-check of the tcgen05 path) */
#define ADP_PREC_BF16X3 3     /* fp32-class accuracy on the tensor cores: values carried as hi+lo bf16, a conv is the three
                                 bf16 GEMMs hi*hi + hi*lo + lo*hi with fp32 accumulation (inference only; max-abs
                                 probability error ~2e-5 vs float64, the <= 1e-4 "fp32/TF32 path" of BASELINE.json) */

/* blend modes — GaussianBlender / LinearBlender, Segmentation/full_evaluation_enhanced.py:115-204 */
#define ADP_BLEND_GAUSSIAN 0
#define ADP_BLEND_LINEAR 1

/* Dihedral op codes (aug image A[i][j] = img[src(i,j)], N = tile size):
 *  0 ident (i,j)       1 rot90 (j,N-1-i)    2 rot180 (N-1-i,N-1-j)   3 rot270 (N-1-j,i)
 *  4 flip_h (i,N-1-j)  5 flip_v (N-1-i,j)   6 flip_h∘rot90 (N-1-j,N-1-i)   7 flip_v∘rot90 (j,i)
 * TestTimeAugmentation modes (full_evaluation_enhanced.py:545-575, segmentation_inference.py:181-219):
 *  minimal = {0,4}   basic = {0,4,5,1}   full = {0,1,2,3,4,5,6,7}   (this order) */
#define ADP_TTA_NONE 0
#define ADP_TTA_MINIMAL 1
#define ADP_TTA_BASIC 2
#define ADP_TTA_FULL 3

typedef struct adp_engine adp_engine;

/* ---- library ------------------------------------------------------------------------------ */
int adp_abi_version(void);
const char *adp_last_error(void);           /* thread-local message of the last failing call */
int adp_device_count(void);                 /* number of sm_100 devices visible, 0 if none */

/* ---- engine lifetime ------------------------------------------------------------------------
 * Replaces AdiposeUNet.__init__/build_model (full_evaluation_enhanced.py:1156-1264,
 * segmentation_inference.py:82-146): the graph is fixed (init_nb, 22 convs), so "building"
 * is allocating the engine.  max_forwards = how many (tile, augmentation) forwards are kept in
 * flight per launch wave (activation arena is sized for it; 0 = default 16). */
int adp_create(int device, int precision, int init_nb, int max_forwards, adp_engine **out);
int adp_destroy(adp_engine *e);
int adp_precision(const adp_engine *e);
int adp_synchronize(adp_engine *e);
void *adp_stream(adp_engine *e);            /* the engine's cudaStream_t (for event timing by the caller) */
/* engine switches for tests and experiments (tcgen05 path only; every fused form is bit-identical to its separate kernels,
 * tests/test_gpu_forward.py, tests/test_gpu_train.py):
 *   "fuse_head"    softmax head in the last conv's epilogue (inference)                                    default 1
 *   "fuse_pool"    2x2 max-pool in the encoder convs' epilogue                                             default 1
 *   "fuse_dropout" the four Dropout sites in the producing conv's epilogue (training forward)              default 1
 *   "fuse_upsum"   UpSampling2D's backward in the epilogue of up1_conv1's data-gradient conv               default 1
 *   "fuse_first"   first conv computed inside down1_conv2 (inference; measured break-even)  default 0
 *   "resident_weights" packed weights of the narrow layers stay in shared memory for the life of a CTA     default 1
 * (also "train_accuracy", "train_eval_mode" below) */
int adp_set_option(adp_engine *e, const char *key, int value);

/* ---- weights ---------------------------------------------------------------------------------
 * Replaces net.load_weights / load_legacy_weights (full_evaluation_enhanced.py:1266-1301): the
 * host layer parses the .h5 and hands over one tensor pair per Keras layer name
 * ("down1_conv1" ... "up1_conv3", "dilate1".."dilate6", "output_softmax"): kernel float32 HWIO
 * (kh,kw,cin,cout) and bias (cout).  adp_get_weight copies the fp32 master copy back out. */
int adp_set_weight(adp_engine *e, const char *layer, const float *kernel_hwio, const int64_t kshape[4],
                   const float *bias, int64_t nbias);
int adp_get_weight(adp_engine *e, const char *layer, float *kernel_hwio, int64_t kernel_elems,
                   float *bias, int64_t nbias);
int adp_weights_ready(adp_engine *e);       /* 1 if all 22 layers are set, else 0 */

/* ---- tile inference ---------------------------------------------------------------------------
 * Replaces AdiposeUNet.predict_single (full_evaluation_enhanced.py:1303-1321,
 * segmentation_inference.py:153-158) and TestTimeAugmentation.predict_with_tta
 * (full_evaluation_enhanced.py:577-600) for a batch of n square tiles of side `size`
 * (multiple of 8; 1024 in the reference): x = (tile - mean) / (std + 1e-10) in float32,
 * forward, inverse dihedral op, mean over the n_ops augmentations in list order.
 * tiles: n*size*size float32 (gray 0..255).  out: n*size*size float32 probabilities.
 * ops == NULL or n_ops == 0 means a single identity forward. */
int adp_predict(adp_engine *e, const float *tiles, int n, int size, float mean, float std,
                const int *ops, int n_ops, float *out);
/* same, uint8 input; channels = 1 (gray) or 3 (interleaved RGB, converted with OpenCV 4.x's
 * 8-bit fixed-point formula Y=(9798R+19235G+3735B+16384)>>15 == cv2.cvtColor(RGB2GRAY); the
 * reference reads tiles with cv2.imread(..., IMREAD_GRAYSCALE), train_adipose_unet_v3.py:555) */
int adp_predict_u8(adp_engine *e, const uint8_t *tiles, int n, int size, int channels, float mean,
                   float std, const int *ops, int n_ops, float *out);
int adp_tta_ops(int mode, int ops[8]);      /* fills ops, returns their count (1,2,4,8) */
/* de-augment + average only: planes = n_ops probability maps (size*size each) predicted on the
 * augmented inputs aug_k(img) by ANY predict_single backend (the reference's ONNX seam,
 * segmentation_inference.py:161-178); out = mean_k deaug_k(planes[k]) — the tail of
 * TestTimeAugmentation.predict_with_tta (full_evaluation_enhanced.py:591-594). */
int adp_tta_combine(adp_engine *e, const float *planes, int size, const int *ops, int n_ops, float *out);

/* per-layer activation tap for parity tests: after adp_predict with n*n_ops <= max_forwards,
 * copies layer `name` of forward `idx` out as float32 NHWC (h*w*channels, unpadded channels).
 * Names: the 21 3x3 conv layers, "dilate_add", "prob". */
int adp_debug_layer(adp_engine *e, const char *name, int idx, float *out, int64_t out_elems,
                    int64_t shape_hwc[3]);

/* ---- threshold + pixel metrics -----------------------------------------------------------------
 * Replaces binarize_prediction / calculate_pixel_metrics (full_evaluation_enhanced.py:716-785):
 * mask = (p > thr) as uint8 (may be NULL), truth = (gt > 0.5) (gt uint8, may be NULL -> counts
 * only count predicted positives in counts[0] and negatives in counts[3]).
 * counts = {tp, fp, fn, tn}.  The ratios (+1e-10) are formed by the caller in float64. */
int adp_threshold_metrics(adp_engine *e, const float *prob, const uint8_t *gt, int64_t n_px, float thr,
                          uint8_t *mask, int64_t counts[4]);
/* TP/FP/FN/TN at every candidate threshold in ONE pass over (prob, gt): the reference's threshold search
 * (optimize_threshold_f1_slide_level / optimize_threshold_f1, full_evaluation_enhanced.py:891-980) evaluates
 * calculate_pixel_metrics once per candidate.  thresholds: n_thr (1..64) strictly ascending floats; gt: uint8, != 0 = positive;
 * counts: n_thr x {tp, fp, fn, tn}, identical to n_thr calls of adp_threshold_metrics. */
int adp_threshold_sweep(adp_engine *e, const float *prob, const uint8_t *gt, int64_t n_px, const float *thresholds, int n_thr,
                        int64_t *counts);

/* ---- boundary refinement -------------------------------------------------------
 * Replaces BoundaryRefiner.refine (full_evaluation_enhanced.py:332-393; callers :1575-1576 and
 * reconstruct_full_images.py:378-380): mask_u8 = (mask*255).astype(uint8); band = (dilate > 0) xor (erode > 0) with the
 * kernel_size x kernel_size MORPH_ELLIPSE element; bilateral filter (d, sigma_color, sigma_space) inside the band;
 * MORPH_OPEN; MORPH_CLOSE; / 255.0.  mask, out: n images of H*W float32 (host or device).  kernel_size 1..15,
 * bilateral_d 1..9.  The reference's unused `image` argument is not part of the call. */
int adp_boundary_refine(adp_engine *e, const float *mask, int n, int H, int W, int kernel_size, int bilateral_d, float sigma_color,
                        float sigma_space, float *out);

/* ---- blenders ----------------------------------------------------------------------------------
 * Replaces GaussianBlender.reconstruct / LinearBlender.reconstruct
 * (full_evaluation_enhanced.py:149-204).  tiles: n tiles of th*tw float32, positions (y,x) in list
 * order, window: th*tw float32 weight map slice (Gaussian; NULL for linear), out: H*W float32.
 * Bit-exact with the NumPy code for identical inputs (same fp32 mul, add, div, same tile order). */
int adp_blend_reconstruct(adp_engine *e, int blend_mode, const float *tiles, int n, int th, int tw,
                          const int32_t *ys, const int32_t *xs, const float *window, int H, int W,
                          float *out);

/* ---- whole-slide sliding window (persistent accumulator) ---------------------------------------
 * Replaces SlidingWindowInference.predict_with_sliding_window (full_evaluation_enhanced.py:286-329)
 * and reconstruct_slide (reconstruct_full_images.py:334-417) without ever holding the per-tile
 * predictions: tiles are predicted (TTA) and blend-accumulated on the device.
 * The accumulator covers slide rows [y0, y0+rows) x [0, W): a rank that owns a tile-row strip
 * passes its own row range (multi-GPU sharding).
 * window: tile*tile float32 (Gaussian) or NULL (linear). */
int adp_wsi_begin(adp_engine *e, int rows, int W, int y0, int tile, int blend_mode, const float *window);
/* predict n tiles (float32 gray 0..255) and accumulate them at slide positions (ys[i], xs[i]) */
int adp_wsi_push_tiles(adp_engine *e, const float *tiles, int n, const int32_t *ys, const int32_t *xs,
                       float mean, float std, const int *ops, int n_ops);
/* same, the tiles are cut on the device from a resident uint8 slide region (slide: region_rows x W pixels covering rows
 * [region_y0, ...); channels = 1: gray bytes; channels = 3: interleaved RGB, converted by the first conv kernel with
 * OpenCV's 8-bit RGB2GRAY formula, i.e. what cv2.imread(..., IMREAD_GRAYSCALE) of reconstruct_full_images.py:364 feeds
 * the model) - the sliding-window case.
 * defer_below_row > y0 (multi-GPU strips): slide rows [y0, defer_below_row) of this accumulator also
 * receive partial sums of the strip above, and float32 addition is order-dependent.  The reference adds tiles in
 * row-major order (full_evaluation_enhanced.py:165-173), i.e. the upper strip's tiles FIRST.  The call therefore keeps
 * the TTA-mean probabilities of these tiles on the device, blends only their rows >= defer_below_row now, and
 * adp_wsi_replay_deferred - called after adp_wsi_import_add has put the upper strip's partial sums into the still
 * untouched zone rows - blends their rows < defer_below_row in the original order: masks do not depend on the GPU count. */
int adp_wsi_push_from_slide(adp_engine *e, const uint8_t *slide, int channels, int region_y0, int region_rows,
                            int n, const int32_t *ys, const int32_t *xs, float mean, float std,
                            const int *ops, int n_ops, int defer_below_row);
int adp_wsi_replay_deferred(adp_engine *e);
/* add already-predicted probability tiles (blend only; used for ground-truth re-blending,
 * reconstruct_full_images.py:404-415) */
int adp_wsi_push_probs(adp_engine *e, const float *probs, int n, const int32_t *ys, const int32_t *xs);
/* raw partial sums of rows [y, y+rows) (slide coordinates): acc and weight (float32 each; for
 * linear blending weight holds the count as float32) — the boundary rows a strip owner ships */
int adp_wsi_export(adp_engine *e, int y, int rows, float *acc, float *weight);
/* add another strip's partial sums into this accumulator */
int adp_wsi_import_add(adp_engine *e, int y, int rows, const float *acc, const float *weight);
/* normalise rows [y, y+rows): prob = acc / max(weight, 1e-8) (linear: / max(count,1)), threshold,
 * count.  prob, mask, gt may each be NULL.  counts = {tp, fp, fn, tn}. */
int adp_wsi_finalize(adp_engine *e, int y, int rows, float thr, float *prob, uint8_t *mask,
                     const uint8_t *gt, int64_t counts[4]);
int adp_wsi_end(adp_engine *e);

/* ---- tile I/O front-end --------------------------------------------
 * adp_jpeg_decode: nvJPEG decode of n JPEG tiles (size x size) on the engine's stream.  gray = the luma plane, i.e. what
 * cv2.imread(path, IMREAD_GRAYSCALE) returns (libjpeg decodes with out_color_space = JCS_GRAYSCALE; reconstruct_full_images.py:364,
 * full_evaluation_enhanced.py:1383, segmentation_inference.py:435); rgb = interleaved RGB with interpolated chroma upsampling
 * (cv2.imread(IMREAD_COLOR) + COLOR_BGR2RGB, reconstruct_full_images.py:367-368).  The decoded tiles stay on the device
 * (*gray_dev / *rgb_dev, valid until the next decode) for adp_predict_u8 / adp_wsi_push_tiles_u8 / adp_wsi_push_aux; host
 * copies are optional.  nvJPEG's inverse DCT is not bit-identical to libjpeg-turbo's: tests/test_gpu_io.py measures the
 * difference (grey levels and resulting mask agreement). */
int adp_jpeg_decode(adp_engine *e, const uint8_t *const *data, const size_t *lengths, int n, int size, uint8_t *gray_host,
                    uint8_t *rgb_host, uint8_t **gray_dev, uint8_t **rgb_dev);
/* adp_wsi_push_tiles for packed uint8 tiles (host or device; channels = 1 gray, 3 interleaved RGB -> OpenCV gray formula) */
int adp_wsi_push_tiles_u8(adp_engine *e, const uint8_t *tiles, int channels, int n, const int32_t *ys, const int32_t *xs,
                          float mean, float std, const int *ops, int n_ops);
/* Auxiliary whole-slide planes next to the probability accumulator, sharing its weight plane (same tiles, positions and
 * window => same weight sums): the RGB mosaic (three blender.reconstruct calls on the tiles' colour channels as float32
 * in [0,1], reconstruct_full_images.py:411-415) and the blended ground truth (:404-409) without host-side float32 tile
 * lists.  adp_wsi_aux_begin after adp_wsi_begin; adp_wsi_push_aux blends n tiles (n_planes = 1 or 3 interleaved samples,
 * is_u8: bytes scaled by 1/255 in float32, else float32 values; host or device) into planes plane0..; push the same
 * tiles in the same order as the predictions.
 * adp_wsi_export_u8: (normalised plane * 255).astype(uint8) of n_planes consecutive planes, interleaved, optionally in
 * reverse plane order (the reference writes RGB2BGR-converted data, :726-728); plane0 = -1 exports the probability plane
 * (prediction_mask.tif, :733).  adp_wsi_export_f32: one normalised plane.  adp_wsi_finalize_auxgt: adp_wsi_finalize with
 * truth = (normalised gt_plane > 0.5), i.e. calculate_pixel_metrics(full_pred, full_gt, thr) (:745-749). */
int adp_wsi_aux_begin(adp_engine *e, int n_planes);
int adp_wsi_push_aux(adp_engine *e, int plane0, int n_planes, const void *tiles, int is_u8, int n, const int32_t *ys,
                     const int32_t *xs);
int adp_wsi_export_u8(adp_engine *e, int plane0, int n_planes, int reverse, int y, int rows, uint8_t *out);
int adp_wsi_export_f32(adp_engine *e, int plane, int y, int rows, float *out);
int adp_wsi_finalize_auxgt(adp_engine *e, int gt_plane, int y, int rows, float thr, float *prob, uint8_t *mask, int64_t counts[4]);
/* fat-% of n probability tiles of px pixels: 100 * count(p > thr) / px (calculate_fat_percentage,
 * tile_classification_evaluation.py:211-225), one launch for the batch */
int adp_tile_fat_percent(adp_engine *e, const float *probs, int n, int64_t px, float thr, double *pct);
/* Baseline TIFF with LZW compression, strips compressed in parallel on `threads` host threads (0 = all cores); replaces
 * tifffile.imwrite(path, array, compression='lzw') (reconstruct_full_images.py:724-734, segmentation_inference.py:455-464).
 * data: H x W x channels uint8 (channels 1 or 3, samples written in the caller's order).  Host only, no engine needed. */
int adp_tiff_write_lzw(const char *path, const uint8_t *data, int64_t H, int64_t W, int channels, int rows_per_strip, int threads);

/* ---- loss ---------------------------------------------------------------------------------------
 * adp_loss_metrics: combined_loss_standard = mean BCE + (1 - global Dice) and dice_coef
 * (train_adipose_unet_v3.py:217-241, src/utils/model.py:93-98) of probabilities p against y,
 * n_px pixels over the whole batch; dldp (may be NULL) receives dL/dp.
 * out = {loss, bce_mean, dice_loss, dice_coef}.
 * adp_loss_metrics_ex: the reference's other compile_model choices (train_adipose_unet_v3.py:808-855) on `batch`
 * images of px_per_image pixels whose trailing axis has row_len pixels (the W of the reference's (B,H,W) tensors; 0 = flat,
 * allowed without hard mining): ohem_keep_ratio < 1 = online_hard_example_mining_loss (:282-323).  As written in the
 * reference, tf.keras.losses.binary_crossentropy averages the LAST axis, so the "per-pixel" BCE it ranks is the (B,H)
 * tensor of per-row means: the loss is the mean of the top int(float32(H)*ratio) row means of every image (716 of 1024
 * rows at the default 0.7; ties to the lower row index like tf.nn.top_k) + Dice over all pixels, and a selected row's
 * pixels each receive 1/(W*B*k) of the BCE gradient.  eps_pos/eps_neg > 0 = asymmetric label smoothing
 * (:244-279, 326-363: ys = y*(1-eps_pos-eps_neg)+eps_neg in both terms; dice_coef keeps the raw y). */
int adp_loss_metrics(adp_engine *e, const float *p, const float *y, int64_t n_px, float *dldp, double out[4]);
int adp_loss_metrics_ex(adp_engine *e, const float *p, const float *y, int batch, int64_t px_per_image, int64_t row_len,
                        float ohem_keep_ratio, float eps_pos, float eps_neg, float *dldp, double out[4]);

/* ---- training step -----------------------------------------------------------------------------
 * Replaces Keras train_step as driven by net.fit (train_adipose_unet_v3.py:1316-1324, 1413-1421):
 * forward with Dropout(0.3) at the four sites (:682,696,703,710), combined_loss_standard (:228-241),
 * reverse-mode differentiation of the graph, Adam / AdamW update (:801-806).
 * The engine keeps theta, the flat gradient, the Adam moments and all activations on the device
 * between adp_train_begin and adp_train_end; adp_get_weight returns the current parameters.
 * x: batch*size*size float32 already normalised (the reference normalises on the host, :589-595),
 * y: batch*size*size float32 targets in [0,1].  Host or device pointers.
 *
 * A step is forward -> backward -> apply so that a data-parallel caller can (a) sum the six loss
 * sums over ranks before backward (exact whole-batch Dice) and (b) all-reduce the flat
 * gradient (adp_train_grad_buffer) before apply.  adp_train_step chains the three for one GPU. */
#define ADP_OPT_ADAM 0
#define ADP_OPT_ADAMW 1
int adp_train_begin(adp_engine *e, int batch, int size, float dropout_rate, uint64_t seed);
/* loss recipe of the following steps (default: standard loss = 1, 0, 0); see adp_loss_metrics_ex */
int adp_train_set_loss(adp_engine *e, float ohem_keep_ratio, float eps_pos, float eps_neg);
/* dropout_masks: NULL (masks drawn from the engine's counter-based generator when dropout_rate > 0) or four
 * uint8 0/1 arrays in NHWC with the real channel counts, sites in graph order
 * {dilate1 (size/8, 8*init_nb), up3 (size/4, 4*init_nb), up2 (size/2, 2*init_nb), up1 (size, init_nb)}.
 * sums = {sum of the BCE terms in the mean, sum ys*pc, sum ys, sum pc, sum y*p, sum p, sum y, number of BCE terms} over this
 * batch (pc = clip(p,1e-7,1-1e-7), ys = smoothed target; with hard mining the BCE terms are the selected per-row means);
 * every entry is additive over data-parallel ranks.  sums may be NULL: the values then stay in the device buffer
 * adp_train_sums_buffer() returns (no host synchronisation in the step). */
int adp_train_forward(adp_engine *e, const float *x, const float *y, int batch, const uint8_t *const *dropout_masks,
                      double *sums /* 8 per output: adp_train_outputs() x 8, or NULL */);
/* device pointer + count (8 x adp_train_outputs()) of the float64 loss sums of the last adp_train_forward: a data-parallel
 * caller all-reduces them in place on the engine's stream and calls adp_train_backward(e, NULL, ...) */
int adp_train_sums_buffer(adp_engine *e, double **dev_ptr, int *count);
/* host copy of that buffer (synchronises the engine's stream); count must equal 8 x adp_train_outputs() */
int adp_train_sums_read(adp_engine *e, double *host, int count);
/* loss = {loss, bce_mean, dice_loss, dice_coef} from (possibly rank-summed) sums */
int adp_train_loss(const double sums[8], double out[4]);
/* sums: the values the loss is defined over (own batch, or summed over data-parallel ranks) in host memory, or NULL = the
 * device sums buffer as it stands; stream-ordered (returns without synchronising the engine's stream);
 * freeze_encoder != 0 = phase 1 of the reference (down*_conv* frozen, :760-769): their gradients are zero */
int adp_train_backward(adp_engine *e, const double *sums /* 8 per output */, int freeze_encoder);
/* Deep supervision (train_adipose_unet_v3.py:712-745, 808-872): while training, aux_out1 = sigmoid(Conv2D(1,1x1)) on the
 * post-dropout up3 tensor (size/4) and aux_out2 on up2 (size/2), each bilinearly resized (half-pixel centres) to full
 * resolution and scored against y with the standard / label-smoothing loss (never hard mining); total loss =
 * w_main*L(main_out) + w_aux1*L(aux_out1) + w_aux2*L(aux_out2).  Needs the layers "aux_out1" (1,1,4*init_nb,1) and
 * "aux_out2" (1,1,2*init_nb,1) set with adp_set_weight; they join the flat parameter / gradient buffer after
 * output_softmax.  Inference ignores them (full_evaluation_enhanced.py:1313-1319 only reads main_out).
 * adp_train_outputs: 1, or 3 with deep supervision = number of 8-double groups in `sums` (main, aux1, aux2). */
int adp_train_set_deep_supervision(adp_engine *e, int on, float w_main, float w_aux1, float w_aux2);
int adp_train_outputs(adp_engine *e);
/* device pointer + element count of the flat fp32 gradient (Keras order: per layer kernel HWIO, bias) */
int adp_train_grad_buffer(adp_engine *e, float **dev_ptr, int64_t *count);
/* Keras' 'binary_accuracy' metric of compile_model (train_adipose_unet_v3.py:850-853, 877-878) for the batch of the last
 * adp_train_forward: out = {pixels where y == (p > 0.5), pixels}; counted on the device when
 * adp_set_option(e, "train_accuracy", 1) is on and mirrored to the host with the loss sums by adp_train_backward (no extra
 * synchronisation).  adp_set_option(e, "train_eval_mode", 1) makes adp_train_forward the validation pass of net.fit: same
 * graph, heads and losses with Dropout inactive. */
int adp_train_accuracy_read(adp_engine *e, double out[2]);
/* Overlapping the gradient exchange with the backward pass: the flat gradient is completed in adp_train_grad_buckets()
 * contiguous element ranges [lo, hi) (returned in completion order: decoder + heads, dilate6..4, dilate3..1, encoder); after
 * adp_train_backward has been enqueued, adp_train_bucket_wait(e, b, s) makes the caller's CUDA stream s wait for bucket b
 * (cudaStreamWaitEvent on an event the backward pass recorded on the engine's stream), the caller enqueues its collective
 * on s, and adp_train_join(e, s) makes the engine's stream wait for everything enqueued on s before adp_train_apply. */
int adp_train_grad_buckets(adp_engine *e, int64_t *lo, int64_t *hi, int cap);
int adp_train_bucket_wait(adp_engine *e, int bucket, void *stream);
int adp_train_join(adp_engine *e, void *stream);
/* whole flat gradient to / from host memory (count must equal the parameter count): the staging path of a
 * data-parallel caller whose collective runs on host buffers (gloo), and of tests that emulate ranks in-process */
int adp_train_grad_read(adp_engine *e, float *host, int64_t count);
int adp_train_grad_write(adp_engine *e, const float *host, int64_t count);
int adp_train_get_grad(adp_engine *e, const char *layer, float *kernel_hwio, int64_t kernel_elems, float *bias, int64_t nbias);
int adp_train_probs(adp_engine *e, float *out, int64_t out_elems);          /* probabilities of the last forward */
/* theta <- optimizer(theta, grad_scale * grad); Keras epsilon placement; beta1/beta2/eps <= 0 select 0.9/0.999/1e-7 */
int adp_train_apply(adp_engine *e, int optimizer, float lr, float grad_scale, double beta1, double beta2, float eps,
                    float weight_decay, int freeze_encoder);
int adp_train_step(adp_engine *e, const float *x, const float *y, int batch, int optimizer, float lr, float weight_decay,
                   int freeze_encoder, double out[4]);
int64_t adp_train_iterations(adp_engine *e);
int adp_train_end(adp_engine *e);
/* one optimizer update on caller-owned arrays (unit-test entry of the update rule): t is 1-based */
int adp_adam_update(adp_engine *e, float *theta, const float *grad, float *m, float *v, int64_t n, int64_t t, int optimizer,
                    float lr, double beta1, double beta2, float eps, float weight_decay);

/* ---- profiling ---------------------------------------------------------------------------------
 * When enabled, every kernel launch is bracketed by CUDA events on the engine stream and
 * accumulated per kernel kind.  adp_profile_read returns up to `cap` rows
 * (name, launches, total milliseconds, algorithmic flops, algorithmic bytes). */
typedef struct adp_prof_row {
  char name[48];
  int64_t launches;
  double ms;
  double flops;
  double bytes;
} adp_prof_row;
int adp_profile_enable(adp_engine *e, int on);
int adp_profile_reset(adp_engine *e);
int adp_profile_read(adp_engine *e, adp_prof_row *rows, int cap);
int64_t adp_launch_count(adp_engine *e);    /* kernels launched by this engine since creation */

#ifdef __cplusplus
}
#endif
#endif /* ADIPOSE_B200_H */
