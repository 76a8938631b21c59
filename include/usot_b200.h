/*
 * usot_b200 -- C ABI of the Blackwell-native (sm_100a) USOT forward path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  All `float*` arguments are DEVICE
 * pointers unless the name says `host_`; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Work is enqueued on `stream`; no call synchronises unless stated.  Every function returns 0 on success and a
 * non-zero code on failure, in which case usot_last_error() (thread-local) describes it -- the library never
 * calls exit() (the reference does, lib/models/prroi_pool/src/prroi_pooling_gpu_impl.cu:20-27).
 *
 * Layout conventions
 *   "nchw" : the reference's layout (contiguous N,C,H,W fp32)       -- images, rois, cls/bbox maps, loss scalars
 *   "nhwc" : contiguous N,H,W,C fp32 -- every 256-channel feature map that crosses the boundary (zf, xf, memory
 *            features).  The Python host exposes them as NCHW-shaped tensors with channels-last strides, which the
 *            reference's callers accept (lib/tracker/usot_tracker.py:105-106,196-199,258-261).
 *
 * Reference interfaces replaced (file:line under /root/reference):
 *   usot_prroi_pool_forward          <- prroi_pooling_forward_cuda       lib/models/prroi_pool/src/prroi_pooling_gpu.c:22-44
 *                                       (+ PrRoIPoolingForwardGpu        lib/models/prroi_pool/src/prroi_pooling_gpu_impl.cu:381-402)
 *   usot_prroi_pool_backward / _coor_backward <- prroi_pooling_backward_cuda / prroi_pooling_coor_backward_cuda   prroi_pooling_gpu.c:46-107
 *   usot_xcorr_depthwise             <- xcorr_depthwise                  lib/models/connect.py:147-157
 *   usot_groupdw_xcorr               <- GroupDW.forward                  lib/models/connect.py:86-102
 *   usot_conv2d_nhwc                 <- nn.Conv2d + BatchNorm2d (+ReLU)  lib/models/modules.py:37-58, connect.py:20-53
 *   usot_stem_conv / usot_maxpool3x3s2p1_nhwc <- conv1+bn1+relu / maxpool of ResNet_plus2   lib/models/modules.py:70-75,138-141
 *   usot_stem_maxpool                         <- the same four modules as one fused kernel   lib/models/modules.py:70-75,138-141
 *   usot_conf_fusion                 <- Conf_Fusion.forward (reduction)  lib/models/connect.py:123-144
 *   usot_cycle_glue                  <- forward-track argmax + box maps  lib/models/models.py:131-162,262-274
 *   usot_weighted_bce / usot_iou_loss <- _weighted_BCE / add_iouloss    lib/models/models.py:42-100
 *   usot_conv2d_wgrad_nhwc / usot_conv2d_dgrad_nhwc / usot_bn_* / usot_*_backward <- torch autograd of the same modules under
 *                                       loss.backward()                 scripts/train_usot.py:229-236
 *   usot_pred_conv                   <- bbox_pred / cls_pred / cls_memory_pred + their epilogues   lib/models/connect.py:235-241,274-275
 *   usot_engine_template             <- USOT_.template                   lib/models/models.py:173-177
 *   usot_engine_track                <- USOT_.track                      lib/models/models.py:179-198
 *   usot_engine_extract_memory_feature <- USOT_.extract_memory_feature   lib/models/models.py:200-206
 *   usot_engine_backbone_neck        <- feature_extractor + neck         lib/models/models.py:39-40,181-184
 *   usot_engine_forward_train        <- USOT_.forward                    lib/models/models.py:208-295
 *   usot_tracker_postprocess         <- USOTTracker.update tensor path   lib/tracker/usot_tracker.py:137-163
 *   usot_engine_track_frame          <- USOTTracker.track + update       lib/tracker/usot_tracker.py:133-276 (one call per frame)
 *   usot_crop_resize                 <- get_subwindow_tracking + cv2.resize  lib/utils/track_utils.py:30-119 (im_to_torch :24-27)
 *   usot_engine_load_tensor/finalize <- load_state_dict contract         lib/utils/train_utils.py:92-128
 */
#ifndef USOT_B200_H
#define USOT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USOT_API __attribute__((visibility("default")))

typedef struct usot_engine usot_engine;

/* precision modes of the dense convolutions */
enum {
    USOT_PREC_FP32_SIMT = 0, /* fp32 FMA on CUDA cores: exact reference arithmetic, cross-check path            */
    USOT_PREC_FP16X3_TC = 1, /* tcgen05 kind::f16, operands split hi+lo fp16, 3 MMAs: fp32-equivalent (default) */
    USOT_PREC_FP16_TC = 2    /* tcgen05 kind::f16 single pass, fp32 accumulate: fast mode (BASELINE config 3)    */
};

USOT_API const char* usot_last_error(void);
USOT_API int usot_abi_version(void);
/* Process-wide performance knobs (never change results): "groupdw_strips" = 2 | 3; "tc_bn_max" = 64 | 128 | 256;
 * "tc_cta_pair" = 0..7 (bit 0: single-fp16, bit 1: fp16x3 conv launches of the MMA-bound layers with large grids run as clusters of two CTAs executing tcgen05.mma.cta_group::2 with M = 256, bit 2: every eligible layer; same results bit for bit),
 * "tc_pdl" = 0 | 1 (programmatic dependent launch of the conv kernels when every tile has its own SM, default 1), "tc_latency_split" = 0 | 1 (small grids use narrower N tiles so more SMs work on a layer, default 1), "tc_l2_prefetch" = 0 | 1 (TMA L2-prefetch hints for the next tile's residual / 1x1 activations, default 0), "tc_tma_f32" = 0 | 1 (fp32-only conv outputs through smem staging + TMA store, default 1), "tc_fuse_cross" = 0 | 1 (split mode: a_hi x [w_hi|w_lo] as one N = 2*BN MMA, default 1), "tc_tma_store" = 0 | 1 (TMA-store epilogue), "tc_tma_res" = 0 | 1 (residual loaded by TMA), "stem_tc" = 0 | 1 (tensor-core stem),
 * "groupdw_tma" = 0 | 1 | 2 (register-staged / TMA ring + scalar FMA / TMA ring + packed FFMA2, default 2),
 * "groupdw_warps4" = 0 | 1 (FFMA2 GroupDW kernel with four consumer warps = one per SM sub-partition, default 1),
 * "stem_pool_fused" = 0 | 1 (stem + max-pool as one TMA-fed tensor-core kernel over the space-to-depth image, default 1; results differ from
 * the two-kernel path by summation order only), "conf_fusion_fused" = 0 | 1 (fp16x3: conf_gen || value_gen as one conv with the Conf_Fusion
 * reduction in its epilogue, default 1), "tc_multi_image_tiles" = 0 | 1 (conv M tiles may span several images, default 1),
 * "tc_skip_pad_rows" = 0 | 1 (one-row tiles skip filter rows inside the zero padding, default 1), "tc_res_ahead" = 1 | 2 (residual chunks
 * requested 1 / NB-1 chunks ahead, default 1),
 * "groupdw_row_split" = 0 | 1 (small batches split one map's output rows over several CTAs, default 1), "pred_tma_min_batch" = 0.. (batches >= this use the TMA-streamed per-image pred-conv kernel, default 48; 0 = never), "graph_max_batch" = 0..64 (track() with n <= this replays a CUDA graph).
 * One accuracy knob: "tc_split_bn_max" = 64 | 128 (default; separate cross-term accumulator) | 256 (single accumulator). */
USOT_API int usot_set_tunable(const char* name, int value);

/* Measurement hooks.  Launches of this library's kernels are always counted per kernel family; with on=1 every launch is
 * additionally bracketed by CUDA events on its own stream so that usot_profile_read can report the family's device time.
 * usot_profile_read synchronises the device; out[4] = {launches, ms, algorithmic FLOPs, algorithmic bytes} since reset. */
USOT_API int usot_profile_reset(int on);
USOT_API int usot_profile_family_count(void);
USOT_API const char* usot_profile_family_name(int family);
USOT_API int usot_profile_read(int family, double* out);
/* Adds `launches` to a family's counter: for replays of a CUDA graph the CALLER captured around this library's calls (the launchers only
 * see the capture, not the replays). */
USOT_API int usot_profile_count(int family, int64_t launches);

/* ------------------------------- stand-alone operators ------------------------------------------------ */

/* PrRoIPool forward, reference layout.  features (n_features,C,H,W) nchw; rois (n_rois,5) = [batch_idx,x1,y1,x2,y2];
 * output (n_rois,C,PH,PW) nchw, fully overwritten. */
USOT_API int usot_prroi_pool_forward(const float* features, const float* rois, float* output, int n_features, int n_rois,
                                     int channels, int height, int width, int pooled_height, int pooled_width,
                                     float spatial_scale, void* stream);

/* PrRoIPool backward w.r.t. the features (replaces prroi_pooling_backward_cuda, prroi_pooling_gpu.c:46-75; kernel
 * prroi_pooling_gpu_impl.cu:214-272).  features_diff (n_features,C,H,W) is zeroed and then accumulated with atomics. */
USOT_API int usot_prroi_pool_backward(const float* rois, const float* output_diff, float* features_diff, int n_features, int n_rois,
                                      int channels, int height, int width, int pooled_height, int pooled_width, float spatial_scale,
                                      void* stream);

/* PrRoIPool backward w.r.t. the roi coordinates (replaces prroi_pooling_coor_backward_cuda, prroi_pooling_gpu.c:77-107; kernel
 * prroi_pooling_gpu_impl.cu:274-379).  rois_diff (n_rois,5): column 0 (batch index) is 0. */
USOT_API int usot_prroi_pool_coor_backward(const float* features, const float* rois, const float* output, const float* output_diff,
                                           float* rois_diff, int n_rois, int channels, int height, int width, int pooled_height,
                                           int pooled_width, float spatial_scale, void* stream);

/* Depth-wise cross-correlation, reference layout.  x (bx,C,hx,wx), kernel (bk,C,hk,wk) with bk == bx or bk == 1
 * (broadcast, the view trick of connect.py:151-156); out (bx,C,hx-hk+1,wx-wk+1). */
USOT_API int usot_xcorr_depthwise(const float* x, const float* kernel, float* out, int bx, int bk, int channels, int hx, int wx,
                                  int hk, int wk, void* stream);

/* Backward of usot_xcorr_depthwise (training path; the reference gets it from autograd through F.conv2d, connect.py:147-157).
 * grad_out (bx,C,hx-hk+1,wx-wk+1); grad_x (bx,C,hx,wx) and grad_kernel (bk,C,hk,wk) are fully overwritten; either may be NULL.
 * With bk == 1 the kernel gradient is the sum over all bx samples (atomics: summation order is not fixed). */
USOT_API int usot_xcorr_depthwise_backward(const float* x, const float* kernel, const float* grad_out, float* grad_x, float* grad_kernel,
                                           int bx, int bk, int channels, int hx, int wx, int hk, int wk, void* stream);

/* Fused GroupDW, nhwc.  x11 (nx,F-2,F-2,C), x12 (nx,F-4,F-2,C), x21 (nx,F-2,F-4,C); z11 (nz,5,5,C), z12 (nz,3,5,C),
 * z21 (nz,5,3,C); weight = the raw 3-vector (softmax is applied inside; read back with one stream sync);
 * out (n_out,F-6,F-6,C).  Sample n uses x[n / (n_out/nx)] and z[n] (or z[0] when nz == 1). */
USOT_API int usot_groupdw_xcorr(const float* x11, const float* x12, const float* x21, const float* z11, const float* z12,
                                const float* z21, const float* weight, float* out, int nx, int nz, int n_out, int channels,
                                int feat_size, void* stream);

/* Dense conv + per-channel affine (+residual)(+ReLU), nhwc.  weight_kn is (kh*kw*cin, cout) with k = (r*kw+s)*cin + c.
 * residual may be NULL.  precision is one of USOT_PREC_*. */
USOT_API int usot_conv2d_nhwc(const float* in, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh, int kw,
                              int stride, int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift,
                              const float* residual, int relu, float* out, int precision, void* stream);

/* The same conv with its INPUT multiplied by a device scalar first (tensor-core precisions only): the dgrad route of the training path,
 * where `in` is a gradient map far below fp16's normal range, in_scale[0] = the power of two from usot_pow2_scale and `scale` carries its
 * inverse.  The multiplication rides on the fp32 -> split-fp16 conversion (no extra pass, exact).  No residual, no ReLU. */
USOT_API int usot_conv2d_nhwc_scaled(const float* in, const float* in_scale, int n, int h, int w, int cin, const float* weight_kn, int cout, int kh,
                                     int kw, int stride, int pad_h, int pad_w, int dil_h, int dil_w, const float* scale, const float* shift,
                                     float* out, int precision, void* stream);

/* Skinny prediction conv of the head: 3x3, pad 1, Cin = channels (256), Cout = 1 or 4, with bias, fused with the reference's
 * epilogue.  in (n,r,r,channels) nhwc; weight (9, cout, channels) with tap = kh*3+kw; out (n,cout,r,r) nchw.
 *   mode 0:  out = mul * (conv + bias)                          (`0.1 * cls_pred(x)`, connect.py:240-241,274-275)
 *   mode 1:  out = exp(adjust[0] * (conv + bias) + bias4[co])   (`exp(adjust * bbox_pred(x) + bias)`, connect.py:235-237)
 * adjust / bias4 are device pointers and may be NULL in mode 0. */
USOT_API int usot_pred_conv(const float* in, int n, int r, int channels, const float* weight, const float* bias, int cout, int mode,
                            float mul, const float* adjust, const float* bias4, float* out, void* stream);

/* SiamFC-style crops for the tracker (replaces the host path get_subwindow_tracking -> cv2.resize -> im_to_torch).
 * frames (n_frames,height,width,3) uint8 HWC (cv2 BGR frames) on the DEVICE; crops (n,4) int32 on the device =
 * [frame index, context_xmin, context_ymin, original_sz] with the context window in frame coordinates BEFORE padding
 * (track_utils.py:44-47; may lie partly or wholly outside the frame); fill (n,3) uint8 on the device = the channel means
 * truncated to uint8 (what the reference's uint8 canvas stores, track_utils.py:58-70).  out (n,3,model_sz,model_sz) float32 nchw,
 * every value an integer in [0,255], bit-identical to the reference's patch.  original_sz == model_sz copies, == 2*model_sz
 * averages 2x2 blocks (OpenCV's INTER_AREA route), anything else is OpenCV's 11-bit fixed-point bilinear. */
USOT_API int usot_crop_resize(const uint8_t* frames, int n_frames, int height, int width, const int32_t* crops, const uint8_t* fill,
                              int n, int model_sz, float* out, void* stream);

/* MaxPool 3x3 / stride 2 / pad 1 of the stem (lib/models/modules.py:75,141), nhwc.  in (n,h,w,C), C % 4 == 0; out (n,ho,wo,C) fp32
 * and/or out_split_sum (n,ho,wo,C): the engine's split-fp16 variant of the same kernel (hi + lo planes, summed back to fp32 by this
 * test/debug entry; allocates scratch and synchronises `stream`).  Either output may be NULL. */
USOT_API int usot_maxpool3x3s2p1_nhwc(const float* in, int n, int h, int w, int channels, float* out, float* out_split_sum, void* stream);

/* Stem conv1 7x7 / stride 2 / pad 0 + per-channel affine (folded BN) + ReLU (lib/models/modules.py:70-74,138-140).  x (n,3,size,size)
 * nchw on the device; host_weight_oihw (64,3,7,7), host_scale / host_shift (64) are HOST pointers (packed + uploaded by this
 * test/debug entry, which synchronises `stream`); out (n,HO,HO,64) nhwc fp32, HO = (size-7)/2+1.  precision = USOT_PREC_*. */
USOT_API int usot_stem_conv(const float* x, int n, int size, const float* host_weight_oihw, const float* host_scale, const float* host_shift,
                            float* out, int precision, void* stream);

/* conv1 + folded BN + ReLU + MaxPool 3x3/2 p1 as the engine runs them in the tensor-core modes (lib/models/modules.py:70-75,138-141): ONE
 * TMA-fed tcgen05 implicit GEMM over the space-to-depth image with the pooling in its epilogue.  Arguments as usot_stem_conv; size <= 261;
 * precision = USOT_PREC_FP16X3_TC or USOT_PREC_FP16_TC; out (n,PO,PO,64) nhwc fp32 = hi + lo of the split-fp16 planes the kernel writes,
 * PO = (HO-1)/2+1.  Test/debug entry: packs and uploads the weights itself and synchronises `stream`. */
USOT_API int usot_stem_maxpool(const float* x, int n, int size, const float* host_weight_oihw, const float* host_scale, const float* host_shift,
                               float* out, int precision, void* stream);

/* The bare stem convolution (7x7 / stride 2 / pad 0, 3 -> 64, no bias, no BN, no ReLU) in fp32 FMA arithmetic: the training path's conv1,
 * which is followed by train-mode BatchNorm (usot_bn_*).  x (n,3,size,size) nchw; weight_kn (147, 64) DEVICE pointer with
 * k = (c*7 + kh)*7 + kw (= weight.reshape(64, 147).T); out (n,HO,HO,64) nhwc. */
USOT_API int usot_stem_conv_raw(const float* x, int n, int size, const float* weight_kn, float* out, void* stream);

/* Conf_Fusion's reduction (lib/models/connect.py:123-144, after the two generator convs): conf / value (batch*nq, per_map) fp32
 * (any layout, elementwise); out[b] = sum_q exp(clamp(conf[b,q],-6,4)) * value[b,q] / sum_q exp(clamp(conf[b,q],-6,4)). */
USOT_API int usot_conf_fusion(const float* conf, const float* value, int batch, int nq, int64_t per_map, float* out, void* stream);

/* Forward-tracking glue of the cycle-memory forward (lib/models/models.py:262-274 with :131-162): per memory sample s of n,
 * res = cls_ratio*off_cls[s] + (1-cls_ratio)*mem_cls[s] over the (R,R) map; idx = first argmax; image box = grid(idx) -/+
 * off_bbox[s,:,idx] (nchw (n,4,R,R)); pool_box (n,4) = image_bbox_to_prpool_bbox(box).  best_score (n) / best_idx (n) may be NULL. */
USOT_API int usot_cycle_glue(const float* off_cls, const float* mem_cls, const float* off_bbox, int n, int score_size, int search_size,
                             int search_feature_size, float cls_ratio, float* pool_box, float* best_score, int32_t* best_idx, void* stream);

/* _weighted_BCE (lib/models/models.py:42-58): 0.5*mean BCEWithLogits over label==1 + 0.5*mean over label==0; loss[1] on the device. */
USOT_API int usot_weighted_bce(const float* pred, const float* label, int count, float* loss, void* stream);
/* add_iouloss / _IOULoss (lib/models/models.py:60-100): bbox (n,4,R,R) nchw, reg_target (n,R,R,4), reg_weight (n,R,R); cells = R*R. */
USOT_API int usot_iou_loss(const float* bbox, const float* reg_target, const float* reg_weight, int n, int cells, float* loss, void* stream);

/* ------------------------------- training path ---------------------------------------- */
/* The backward of USOT_.forward (scripts/train_usot.py:229-236 calls loss.backward(); the reference gets every gradient from torch
 * autograd over cuDNN).  All maps nhwc fp32; weights in the (kh*kw*cin, cout) "kn" layout of usot_conv2d_nhwc. */

/* autograd of nn.Conv2d w.r.t. its weight: grad_weight_kn (kh*kw*cin, cout) is overwritten.  Any stride / dilation / channel count.
 * precision = USOT_PREC_*: the tcgen05 modes run the tensor-core kernel (MN-major operands straight from the nhwc maps, fp16x3 split,
 * gradients pre-scaled by a power of two on the device) when Cin % 64 == 0 and Cout % 64 == 0; otherwise / fp32: fp32 FMA kernel. */
USOT_API int usot_conv2d_wgrad_nhwc(const float* in, const float* grad_out, int n, int h, int w, int cin, int cout, int kh, int kw, int stride,
                                    int pad_h, int pad_w, int dil_h, int dil_w, float* grad_weight_kn, int precision, void* stream);
/* Weight gradient of the stem conv (usot_stem_conv_raw): x (n,3,size,size) nchw, grad_out (n,HO,HO,64) nhwc -> grad_weight_kn (147, 64) with
 * k = (c*7 + kh)*7 + kw (the layout usot_stem_conv_raw takes), overwritten.  fp32 FMA, im2col on the fly. */
USOT_API int usot_stem_conv_wgrad(const float* x, const float* grad_out, int n, int size, float* grad_weight_kn, void* stream);
/* Power-of-two pre-scaling of a gradient map for the split-fp16 tensor-core kernels: s = 2^e with s * max|x| in [2^target_log2,
 * 2^(target_log2+1)) (s = 1 for an all-zero map); scale2[2] (device) = {s, 1/s}; y (optional, may alias nothing) = s * x.  No host sync. */
USOT_API int usot_pow2_scale(const float* x, int64_t numel, int target_log2, float* y, float* scale2, void* stream);
/* autograd of nn.Conv2d w.r.t. its input, generic gather form (the thin 256->1/4 prediction convs; wide layers run their dgrad on the
 * forward conv kernels: usot_conv2d_nhwc with transposed / flipped filters).  grad_in (n,h,w,cin) is overwritten. */
USOT_API int usot_conv2d_dgrad_nhwc(const float* grad_out, const float* weight_kn, int n, int h, int w, int cin, int cout, int kh, int kw,
                                    int stride, int pad_h, int pad_w, int dil_h, int dil_w, float* grad_in, void* stream);
/* nn.BatchNorm2d in train() mode over an (m, channels) map (m = N*H*W): batch mean and BIASED variance of x + bias (bias = the conv
 * bias, may be NULL); then y = (x + bias - mean) / sqrt(var + eps) * gamma + beta (+ residual)(ReLU).  With running statistics passed
 * as mean / var the same call is the eval() forward.  invstd_out (channels, optional) is what usot_bn_backward takes. */
USOT_API int usot_bn_stats(const float* x, const float* bias, int64_t m, int channels, float* mean, float* var, void* stream);
USOT_API int usot_bn_apply(const float* x, const float* bias, const float* mean, const float* var, float eps, const float* gamma, const float* beta,
                           const float* residual, int relu, int64_t m, int channels, float* y, float* invstd_out, void* stream);
/* Backward of usot_bn_apply.  grad_y is the gradient w.r.t. y; with relu != 0, y (the forward output) masks it first.  train != 0:
 * batch-statistics backward; train == 0: mean / invstd are constants (running statistics).  grad_x doubles as the gradient of the
 * conv bias' input; grad_gamma / grad_beta (channels); grad_residual (optional) receives the masked gradient for the shortcut. */
USOT_API int usot_bn_backward(const float* grad_y, const float* y, const float* x, const float* bias, const float* mean, const float* invstd,
                              const float* gamma, int train, int relu, int64_t m, int channels, float* grad_x, float* grad_gamma,
                              float* grad_beta, float* grad_residual, void* stream);
/* out[c] = sum over the m rows of x (m, channels): bias gradients. */
USOT_API int usot_channel_sum(const float* x, int64_t m, int channels, float* out, void* stream);
USOT_API int usot_maxpool3x3s2p1_backward_nhwc(const float* in, const float* grad_out, int n, int h, int w, int channels, float* grad_in,
                                               void* stream);
USOT_API int usot_conf_fusion_backward(const float* conf, const float* value, const float* grad_out, int batch, int nq, int64_t per_map,
                                       float* grad_conf, float* grad_value, void* stream);
/* GroupDW's weighted sum (lib/models/connect.py:96-102): out = w3[0]*x0 + w3[1]*x1 + w3[2]*x2 (w3 on the device) and its backward. */
USOT_API int usot_weighted_sum3(const float* x0, const float* x1, const float* x2, const float* w3, int64_t numel, float* out, void* stream);
USOT_API int usot_weighted_sum3_backward(const float* x0, const float* x1, const float* x2, const float* w3, const float* grad_out, int64_t numel,
                                         float* grad_x0, float* grad_x1, float* grad_x2, float* grad_w3, void* stream);
/* Backward of usot_weighted_bce / usot_iou_loss; grad_loss[1] on the device. */
USOT_API int usot_weighted_bce_backward(const float* pred, const float* label, int count, const float* grad_loss, float* grad_pred, void* stream);
USOT_API int usot_iou_loss_backward(const float* bbox, const float* reg_target, const float* reg_weight, int n, int cells, const float* grad_loss,
                                    float* grad_bbox, void* stream);

USOT_API int usot_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, void* stream);
USOT_API int usot_nhwc_to_nchw(const float* in, int n, int h, int w, int c, float* out, void* stream);

/* ------------------------------- engine ----------------------------------------------------------------- */

USOT_API int usot_engine_create(usot_engine** out, int device, int precision);
USOT_API int usot_engine_destroy(usot_engine* e);
/* Stage one state_dict tensor (HOST pointer, fp32, reference shape/order).  Name = reference state_dict key. */
USOT_API int usot_engine_load_tensor(usot_engine* e, const char* name, const float* host_data, int64_t numel);
/* Fold BN, repack and upload every staged tensor.  Synchronises the device.  May be called again after reloading. */
USOT_API int usot_engine_finalize(usot_engine* e);
/* Bytes of device memory currently owned by the engine (weights + workspace arena). */
/* Packed-weight image (offline BN folding + packed-weight cache).  After finalize() the engine can write
 * everything it uploaded -- folded scale/shift vectors, repacked fp32 weights, the fp16 hi/lo planes of the tensor-core path,
 * the prediction-head layouts -- into one host buffer; usot_engine_import_packed() restores an engine of the SAME precision
 * from such a buffer without the state_dict and without repeating the host-side packing (it replaces load_tensor + finalize).
 * The image is only meaningful for the library ABI version and precision recorded in its header; mismatches fail.
 * The Python host keys cache files by a content hash of the state_dict (usot_b200/engine.py, usot_b200/checkpoint.py). */
USOT_API int64_t usot_engine_packed_size(const usot_engine* e);
USOT_API int usot_engine_export_packed(usot_engine* e, void* host_buf, int64_t capacity);
USOT_API int usot_engine_import_packed(usot_engine* e, const void* host_buf, int64_t size);

USOT_API int64_t usot_engine_device_bytes(const usot_engine* e);

/* feature_extractor + neck:  x (n,3,size,size) nchw -> xf (n,F,F,256) nhwc, F = ((size-7)/2+1 -> pool -> s2) */
USOT_API int usot_engine_backbone_neck(usot_engine* e, const float* x, int n, int size, float* xf, void* stream);
USOT_API int usot_feature_size(int image_size);

/* USOT_.template.  z (n,3,size,size) nchw.  template_bbox (n,4) in feature coords or NULL (pr_pool=False: centre crop
 * [4:-4]).  zf (n,7,7,256) nhwc out.  x_ori (n,F,F,256) nhwc out, may be NULL. */
USOT_API int usot_engine_template(usot_engine* e, const float* z, int n, int size, const float* template_bbox, float* zf,
                                  float* x_ori, void* stream);

/* USOT_.track.  x (n,3,size,size) nchw; zf (nz,7,7,256) nhwc, nz == n or 1; template_mem (n*nq,7,7,256) nhwc or NULL
 * (nq = 0: offline only).  Outputs nchw: cls (n,1,R,R), bbox (n,4,R,R), cls_mem (n,1,R,R) [ignored if nq == 0];
 * xf (n,F,F,256) nhwc, may be NULL.  R = F - 6. */
USOT_API int usot_engine_track(usot_engine* e, const float* x, int n, int size, const float* zf, int nz, const float* template_mem,
                               int nq, float* cls, float* bbox, float* cls_mem, float* xf, void* stream);

/* USOT_.extract_memory_feature.  Exactly one of ori_x (n,3,size,size nchw) / xf (n,feat,feat,256 nhwc) is non-NULL.
 * search_bbox (n,4); out (n,7,7,256) nhwc. */
USOT_API int usot_engine_extract_memory_feature(usot_engine* e, const float* ori_x, int n, int size, const float* xf, int feat,
                                                const float* search_bbox, float* out, void* stream);

/* Head part of USOT_.forward (lib/models/models.py:223-295) with eval-mode BN: everything after the three backbone+neck
 * passes, which the caller runs with usot_engine_template / usot_engine_backbone_neck (so that a multi-GPU caller can overlap
 * its all-gather of zf with them).  zf (n,7,7,256), xf (n,feat,feat,256), xf_mem (n*m,feat,feat,256) nhwc; m = 0 selects the
 * naive-Siamese branch (models.py:288-295).  label (n,R,R), reg_target (n,R,R,4), reg_weight (n,R,R), search_bbox (n,4).
 * losses[3] (device) = {cls_loss, cls_memory_loss (0 if m == 0), reg_loss}.  Optional outputs (may be NULL): backward_map
 * (n,1,R,R) and pool_box (n*m,4) = the PrPool boxes of the forward-tracked targets. */
USOT_API int usot_engine_forward_train(usot_engine* e, const float* zf, const float* xf, const float* xf_mem, int n, int m, int feat,
                                       const float* label, const float* reg_target, const float* reg_weight, const float* search_bbox,
                                       float cls_ratio, float* losses, float* backward_map, float* pool_box, void* stream);

/* One whole tracker frame on the device, enqueued on `stream` without any host synchronisation (USOTTracker.track + update,
 * lib/tracker/usot_tracker.py:133-276, minus the scalar position / size smoothing that stays on the host):
 *   crop (context window [context_xmin, context_ymin, original_sz] of the uint8 HWC `frame`, fill = truncated channel means,
 *   resized to instance_size; bit-exact with get_subwindow_tracking)  ->  gather of the nq memory templates
 *   mem_buf[mem_rows[k]] (mem_rows is a HOST array of row indices into the device-resident queue buffer of (7,7,256) nhwc rows)
 *   ->  track() with the memory branch  ->  usot_tracker_postprocess  ->  pool_label_search of the winning box  ->  PrPool of
 *   the new memory feature from xf into feat_out ((7,7,256) nhwc, typically the next free row of the queue buffer).
 * zf (1,7,7,256) nhwc; window (R,R) float64; target_w/h already multiplied by scale_z; result[8] as usot_tracker_postprocess.
 * frame, fill (3 bytes), zf, mem_buf, window, result, feat_out are DEVICE pointers.  Calls on one engine are serialised on the
 * host and share one device workspace: issue the frames of one engine on one stream. */
USOT_API int usot_engine_track_frame(usot_engine* e, const uint8_t* frame, int height, int width, int context_xmin, int context_ymin,
                                     int original_sz, const uint8_t* fill, int instance_size, const float* zf, const float* mem_buf,
                                     const int32_t* mem_rows, int nq, const double* window, double target_w, double target_h, double ratio,
                                     double penalty_k, double window_influence, int total_stride, double* result, float* feat_out,
                                     void* stream);

/* Tensor path of USOTTracker.update for one frame, on the device (no host sync): sigmoid, ratio mix, box decode on the
 * stride-8 grid, size/ratio penalty, cosine window, argmax.  cls / cls_mem (1,1,R,R), bbox (1,4,R,R) nchw fp32; window (R,R)
 * float64 (numpy's np.outer(np.hanning, np.hanning)); target_w/h = target size already multiplied by scale_z.
 * result[8] (device, float64) = {r_max, c_max, x1, y1, x2, y2, penalty[r,c], mixed_score[r,c]}. */
USOT_API int usot_tracker_postprocess(const float* cls, const float* cls_mem, const float* bbox, const double* window, int score_size,
                                      int instance_size, double target_w, double target_h, double ratio, double penalty_k,
                                      double window_influence, double* result, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USOT_B200_H */
