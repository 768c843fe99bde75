/*
 * confignet_b200.h - C ABI of libconfignet_b200.so (sm_100a only).
 *
 * The reference (microsoft/ConfigNet) has no FFI: its hot path is Keras layers -> TF 2.1 eager
 * ops -> cuDNN/cuBLAS/Eigen kernels.  Every entry point below replaces one of those implicit
 * kernel call sites; the "replaces" comment names the reference line that triggers it.
 *
 * Conventions
 *   - every function returns 0 on success or a negative CN_ERR_* code; cn_last_error() gives a
 *     thread-local message.  Nothing throws, nothing falls back to the CPU.
 *   - all tensor pointers are DEVICE pointers owned by the caller, fp32, channels-last
 *     (NHWC / NDHWC), Keras kernel layouts ((k..., Cin, Cout), Dense (in, out)).
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it.
 *   - TF "SAME" padding is computed inside from the descriptor.
 */
#ifndef CONFIGNET_B200_H
#define CONFIGNET_B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CN_OK 0
#define CN_ERR_BAD_SHAPE (-1)
#define CN_ERR_BAD_ALIGN (-2)
#define CN_ERR_UNSUPPORTED (-3)
#define CN_ERR_CUDA (-4)

/* activation codes for fused epilogues */
#define CN_ACT_NONE 0
#define CN_ACT_LRELU 1
#define CN_ACT_RELU 2
#define CN_ACT_TANH 3
/* two more codes for the metric networks, accepted by cn_dwconv3x3_fwd and cn_act_ext ONLY (the conv / dense entry points
 * reject them: their epilogues carry the four codes above) */
#define CN_ACT_RELU6 4    /* keras ReLU(6.) of MobileNetV2 */
#define CN_ACT_SIGMOID 5  /* attribute-classifier head (metrics/celeba_attribute_prediction.py:62) */

/* kernel selection for the conv / dense family */
#define CN_IMPL_AUTO 0   /* tcgen05 tensor-core kernel where the shape allows, else CUDA-core */
#define CN_IMPL_FFMA 1   /* force the fp32 CUDA-core implicit GEMM */
#define CN_IMPL_TC 2     /* force tcgen05 (error if the shape is not eligible) */

const char* cn_last_error(void);
int cn_version(void);
/* number of kernels this library has launched since load (or since the last reset != 0) */
long long cn_launch_count(int reset);
/* adjusts the counter (a CUDA-graph capture records launches without executing them; replays are added by the caller) */
long long cn_launch_count_add(long long delta);

/* Parameter buffers.  A caller that keeps its Keras kernels in long-lived device buffers (the ParamGroup flat
 * buffers behind model.get_weights()/set_weights(), confignet_first_stage.py:129-206) registers them once; the conv
 * entry points then keep the tensor-core stage images of those kernels until THAT buffer changes.  The optimizer and
 * EMA entry points below mark the buffers they write themselves; any OTHER writer (set_weights, a memcpy, a replayed
 * CUDA graph that contains optimizer launches) must call cn_params_changed(base) - or cn_weights_changed() for all
 * buffers.  cn_set_params_frozen(base, 1) declares a buffer that only changes through cn_params_changed (VGG19 /
 * VGGFace, perceptual_loss.py:19-41): captured graphs then reuse its images instead of re-packing them per replay.
 * Unregistered weight pointers are re-packed on every call. */
int cn_register_params(const void* base, size_t bytes);
int cn_unregister_params(const void* base);
int cn_params_changed(const void* base);
int cn_set_params_frozen(const void* base, int frozen);
/* the buffer's change counter (-1: not registered): a caller that captured a graph over frozen images re-captures when it moves */
long long cn_params_epoch(const void* base);
int cn_weights_changed(void);
/* call once a CUDA graph has been captured over this library's launches: internal scratch / cache buffers are then
 * never freed (a graph may still reference them), growth abandons the old buffer instead */
int cn_graphs_captured(void);

/* Convolution geometry.  nd = 0 describes a Dense layer (batch rows, cin -> cout). */
typedef struct {
  int nd;          /* spatial dims: 0 (dense), 2 (Conv2D), 3 (Conv3D)                        */
  int batch;
  int in_dims[3];  /* spatial size of x BEFORE the fused upsample; unused entries = 1        */
  int cin, cout;
  int ksize[3];    /* unused entries = 1                                                     */
  int stride;      /* 1 or 2, same on every axis                                             */
  int upsample;    /* 1, or 2 = nearest-neighbour x2 (UpSampling2D/3D) fused on the input    */
  int pad;         /* -1: TF "SAME"; p >= 0: ZeroPadding(p) on every spatial side + "VALID"
                      (keras-applications ResNet50 stem: ZeroPadding2D(3) + 7x7/s2, real_encoder.py:13) */
} cn_conv_desc;

/* which kernel family the last conv call on this thread used: 1 CUDA-core, 2 tcgen05 */
int cn_last_conv_impl(void);
/* y spatial dims for `d` (TF SAME: ceil(in*upsample/stride); explicit pad p: (in + 2p - k)/stride + 1). */
int cn_conv_out_dims(const cn_conv_desc* d, int out_dims[3]);

/* y = act(conv_same(upsample(x), w) + bias).
 * replaces keras.layers.Conv2D/Conv3D/Dense __call__: building_blocks.py:28-30,64-66,91,
 * hologan_generator.py:24,50-56,101, hologan_discriminator.py:20,34,46,77,97,
 * perceptual_loss.py:19-24 (VGG convs + ReLU), UpSampling hologan_generator.py:139-170. */
int cn_conv_fwd(const cn_conv_desc* d, const float* x, const float* w, const float* bias,
                int act, float alpha, float* y, int impl, void* stream);
/* gx = d<gy, conv(x,w)>/dx  (x-shaped, i.e. before the fused upsample).
 * replaces the Conv*BackpropInput ops tf.GradientTape emits (confignet_first_stage.py:472,556). */
int cn_conv_dgrad(const cn_conv_desc* d, const float* gy, const float* w, float* gx,
                  int impl, void* stream);
/* gw = d<gy, conv(x,w)>/dw (overwrites gw, Keras layout); gbias (may be NULL) = column sums of gy.
 * replaces Conv*BackpropFilter / BiasAddGrad. */
int cn_conv_wgrad(const cn_conv_desc* d, const float* x, const float* gy, float* gw,
                  float* gbias, int impl, void* stream);

/* ---- per-(sample,channel) reductions and broadcasts over the spatial axes (NHWC) --------------
 * replaces tf.nn.moments / K.mean / K.std and the broadcast arithmetic inside
 * LayerNormalization (building_blocks.py:132-149), InstanceNormalization
 * (instance_normalization.py:108-131) and get_layer_style (confignet_utils.py:147-159),
 * and every term of their first- and second-order gradients (R1: losses.py:42-43,75-82). */
#define CN_FLAG_LRELU_A 1   /* a := lrelu(a, alpha) before use                    */
#define CN_FLAG_MASK_OUT 2  /* affine result *= lrelu'(a_raw)                     */
#define CN_FLAG_MASK_C 4    /* c := c * lrelu'(a_raw)                             */
/* sums[((z*n + i)*C + ch)*8 + j], j < 7: partial sums of pixel slice z (0 <= z < cn_chan_sums_splits(n, p, ch)) of
 * sample i; the caller provides splits*n*C*8 floats (16-byte aligned), no zero-fill needed; cn_norm_coef adds the slices in a
 * fixed order (bit-reproducible statistics).
 *   j: 0 sum a, 1 sum b, 2 sum c, 3 sum a*a, 4 sum a*b, 5 sum a*c, 6 sum b*c  (b, c may be NULL) */
int cn_chan_sums_splits(int n, int p, int ch);
int cn_chan_sums(const float* a, const float* b, const float* c, int n, int p, int ch,
                 int flags, float alpha, float* sums, void* stream);
/* one pass for the two one-operand statistics of a DiscrBlock's conv output (building_blocks.py:100-106): sums of
 * lrelu(a, alpha) -> sums_act (InstanceNormalization after LeakyReLU), sums of a -> sums_raw (get_layer_style); both buffers
 * as for cn_chan_sums, only j = 0 and j = 3 are written.  CN_ERR_UNSUPPORTED when ch % 4 != 0 or ch > 1024. */
int cn_chan_sums_dual(const float* a, int n, int p, int ch, float alpha, float* sums_act, float* sums_raw, void* stream);
/* out = ka[n,ch]*a + kb[n,ch]*b + kc[n,ch]*c + k0[n,ch]; coef is (n, ch, 4) = (ka,kb,kc,k0). */
int cn_chan_affine(const float* a, const float* b, const float* c, const float* coef,
                   int n, int p, int ch, int flags, float alpha, float* out, void* stream);
/* out = [cn_chan_affine(a, b, c, coef, flags)] + coef2.x[n,ch]*a_raw + coef2.w[n,ch]: the InstanceNorm and layer-style
 * gradients of one DiscrBlock conv output (building_blocks.py:100-106: both consume `c`) in one pass instead of two
 * passes and an add.  CN_ERR_UNSUPPORTED when ch % 4 != 0 or ch > 1024 (use two cn_chan_affine calls). */
int cn_chan_affine2(const float* a, const float* b, const float* c, const float* coef, const float* coef2,
                    int n, int p, int ch, int flags, float alpha, float* out, void* stream);
/* the two second-order results of InstanceNormalization(LeakyReLU(a)) that share the operands (a, gy = b, h = c) in one
 * pass: out_a = cn_chan_affine(a, b, c, coef_a, LRELU_A | MASK_C | MASK_OUT), out_g = cn_chan_affine(a, NULL, c, coef_g,
 * LRELU_A | MASK_C) (R1 penalty: the gradients wrt the conv output and wrt the incoming gradient, losses.py:75-82).
 * CN_ERR_UNSUPPORTED when ch % 4 != 0 or ch > 1024. */
int cn_chan_affine_pair(const float* a, const float* b, const float* c, const float* coef_a, const float* coef_g,
                        int n, int p, int ch, float alpha, float* out_a, float* out_g, void* stream);
/* closed-form coefficients from the 7 sums.  kind:
 *  0 IN fwd      p0=gamma p1=beta          -> coef0
 *  1 IN bwd      sums(a,gy) p0=gamma       -> coef0 (input grad), out0=dgamma[ch], out1=dbeta[ch]
 *  2 IN bwd-bwd  sums(a,gy,h) p0=gamma     -> coef0 (d/da), coef1 (d/dgy), out0=dgamma[ch]
 *  3 style fwd   sums(a)                   -> out0 = style (n,2ch) = concat(mean, sqrt(var+eps))
 *  4 style bwd   sums(a) p0=gstyle(n,2ch)  -> coef0
 *  5 style bwd-bwd sums(a,h) p0=gstyle     -> coef0 (d/da), out0 = d/dgstyle (n,2ch)
 *  6 AdaIN fwd   sums(a) p0=sb (n,2ch)     -> coef0
 *  7 AdaIN bwd   sums(a,gy) p0=sb          -> coef0, out0 = dsb (n,2ch)                         */
int cn_norm_coef(int kind, const float* sums, int nsplit, const float* p0, const float* p1, int n, int ch,
                 int npix, float eps, float* coef0, float* coef1, float* out0, float* out1,
                 void* stream);

/* elementwise helpers on flat fp32 buffers */
int cn_lrelu_fwd(const float* x, float alpha, float* y, int64_t n, void* stream);
/* gx = gy * act'(ref): lrelu (ref = pre- or post-activation), relu (ref = y), tanh (ref = y: 1-y^2) */
int cn_act_bwd(const float* gy, const float* ref, int act, float alpha, float* gx, int64_t n, void* stream);
/* gx = (gy + 2 (y - t) * gloss[0] * k) * act'(y): the activation backward of a layer whose output y is also tapped by a
 * squared-difference loss against t (PerceptualLoss._loss_terms, perceptual_loss.py:61-82); gloss = device scalar (the loss
 * cotangent), gy = gradient from the next layer or NULL.  Replaces SquaredDifference/Mean gradients + AddN + ReluGrad.
 * n % 4 == 0, 16-byte aligned tensors. */
int cn_act_bwd_sqdiff(const float* gy, const float* y, const float* t, const float* gloss, float k, int act,
                      float alpha, float* gx, int64_t n, void* stream);
/* out = a*x + b*y (y may be NULL) */
int cn_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream);

/* 2x2/s2 VALID max-pool, NHWC (VGG pools, perceptual_loss.py:19-24) and its backward
 * (gradient goes to the first maximum in window scan order, as TF's MaxPoolGrad). */
int cn_maxpool2_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream);
int cn_maxpool2_bwd(const float* x, const float* y, const float* gy, int n, int h, int w, int c,
                    float* gx, void* stream);

/* transform_3d_grid_tf (confignet_utils.py:63-120): trilinear resample of (B,S,S,S,C) under a
 * per-sample 3x3 matrix `rot` (B,9) about the volume centre, clamp-to-edge. */
int cn_rotate3d_fwd(const float* grid, const float* rot, int b, int s, int c, float* out, void* stream);
/* gradient wrt the grid: a scatter accumulated in 64-bit fixed point (bit-reproducible), then written whole to ggrid */
int cn_rotate3d_bwd_grid(const float* gout, const float* rot, int b, int s, int c, float* ggrid, void* stream);
/* loss reductions (losses.py:7-18,75-82, perceptual_loss.py:76-80): deterministic two-stage sums.
 * kind 0: sum softplus(sign*x)   kind 1: sum (x-y)^2   kind 2: sum x^2   kind 3: sum x
 * each term optionally multiplied by wgt[i / wdiv] (eye-loss mask broadcast over channels).
 * result[0] = scale * sum.  ws must hold cn_reduce_ws_floats() floats. */
int cn_reduce_ws_floats(void);
int cn_reduce(const float* x, const float* y, const float* wgt, int wdiv, int64_t n, int kind,
              float sign, float scale, float* ws, float* result, void* stream);
/* gx = gscale[0] * k * wgt * d(term)/dx */
int cn_reduce_bwd(const float* x, const float* y, const float* wgt, int wdiv, int64_t n, int kind,
                  float sign, float k, const float* gscale, float* gx, void* stream);

/* images: (x+1)*127.5 clipped to [0,255], truncated to uint8 (confignet_first_stage.py:636-637) */
int cn_to_uint8(const float* x, uint8_t* out, int64_t n, void* stream);
/* uint8 -> float /127.5 - 1 (confignet_first_stage.py:444, confignet_second_stage.py:303) */
int cn_from_uint8(const uint8_t* x, float* out, int64_t n, void* stream);
/* VGG 'caffe' preprocessing of [-1,1] RGB(BGR-flipped) images: out[..., c] = (x[..., 2-c]+1)*127.5 - mean[c]
 * (perceptual_loss.py:50-59); mode 1 maps the gradient back.  Modes 2 / 3: the VGGFace variant
 * (perceptual_loss.py:54-56): out = (x+1)*127.5 - (93.5940, 104.7624, 129.1863), no channel flip, and its backward. */
int cn_vgg_preprocess(const float* x, float* out, int64_t npix, int mode, void* stream);

/* Keras Adam + EMA over a flat parameter buffer (confignet_first_stage.py:393-400,601-602):
 * m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr_t * m / (sqrt(v) + eps);
 * if ema != NULL: ema = ema_alpha*ema + (1-ema_alpha)*p.  g is multiplied by gscale first. */
int cn_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                     float lr_t, float b1, float b2, float eps, float ema_alpha, float gscale,
                     void* stream);
/* the same step with lr_t read from device memory (CUDA-graph friendly: the host updates the scalar between replays) */
int cn_adam_ema_step_dev(float* p, const float* g, float* m, float* v, float* ema, int64_t n,
                         const float* lr_t_dev, float b1, float b2, float eps, float ema_alpha, float gscale,
                         void* stream);

/* dst[dst_off[i] : dst_off[i]+n[i]] = src[i] (zeros when src[i] is NULL) for i < count: packs the
 * per-variable gradients tape.gradient returns (confignet_first_stage.py:472-474,556-558) into the flat
 * buffer the all-reduce and cn_adam_ema_step run on.  src/dst_off/n are HOST arrays. */
#define CN_MULTI_MAX 48
int cn_multi_copy(int count, const float* const* src, const int64_t* dst_off, const int64_t* n,
                  float* dst, void* stream);
/* ema = alpha*ema + (1-alpha)*p (update_smoothed_weights, confignet_first_stage.py:393-400) */
int cn_ema(float* ema, const float* p, int64_t n, float alpha, void* stream);

/* ---- second stage / fine-tuning (confignet_second_stage.py, dnn_models/real_encoder.py) -------------------- */
/* inference-statistics BatchNorm of the ResNet50 encoder (keras BatchNormalization without training=True inside
 * the manual tape, real_encoder.py:27), gamma/beta trainable: scale = gamma/sqrt(var+eps), shift = beta - mean*scale */
int cn_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int c,
               float* scale, float* shift, void* stream);
/* out = relu?(x*scale[c] + shift[c] + residual)   (BN + Add + Activation of a ResNet block; residual may be NULL) */
int cn_bn_act_fwd(const float* x, const float* scale, const float* shift, const float* residual, int relu,
                  float* out, int64_t npix, int c, void* stream);
/* g = relu ? gout*(out>0) : gout; gx = g*scale[c]; gres = g (may be NULL); dgamma/dbeta (overwritten) */
int cn_bn_act_bwd(const float* gout, const float* out, const float* x, const float* scale, const float* mean,
                  const float* var, float eps, int relu, float* gx, float* gres, float* dgamma, float* dbeta,
                  int64_t npix, int c, void* stream);
/* ZeroPadding2D(1) + MaxPool 3x3/s2 VALID (ResNet50 pool1) and its backward (first maximum in scan order) */
int cn_maxpool3s2_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream);
int cn_maxpool3s2_bwd(const float* x, const float* y, const float* gy, int n, int h, int w, int c, float* gx, void* stream);
/* GlobalAveragePooling2D over p pixels (resnet50 pooling="avg") */
int cn_avgpool_fwd(const float* x, int n, int p, int c, float* y, void* stream);
int cn_avgpool_bwd(const float* gy, int n, int p, int c, float* gx, void* stream);
/* out[r][c] = x[r][c] * scale[c]  (rotation_range_multiplier, real_encoder.py:20-21,30) */
int cn_col_scale(const float* x, const float* scale, int rows, int c, float* out, void* stream);
/* euler_angles_to_matrix (confignet_utils.py:122-145) on the device and its backward: angles (b,3) -> rot (b,9) */
int cn_euler_fwd(const float* angles, int b, float* rot, void* stream);
int cn_euler_bwd(const float* angles, const float* grot, int b, float* gangles, void* stream);
/* gradient of cn_rotate3d_fwd wrt the matrix (b,9): needed where the rotation is predicted / optimised
 * (confignet_second_stage.py:169-170, 348, 392) */
int cn_rotate3d_bwd_rot(const float* grid, const float* gout, const float* rot, int b, int s, int c, float* grot, void* stream);
/* compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107): out, labels (b, j); the last nrot
 * columns are not normalised.  bwd writes the gradients wrt both tensors, scaled by gscale[0]. */
int cn_norm_latent_loss_fwd(const float* out, const float* labels, int b, int j, int nrot, float weight, float* loss, void* stream);
int cn_norm_latent_loss_bwd(const float* out, const float* labels, int b, int j, int nrot, float weight, const float* gscale,
                            float* g_out, float* g_labels, void* stream);

/* ---- training-time metric networks (metrics/inception_distance.py, metrics/celeba_attribute_prediction.py) --------
 * Forward only.  The convolutions of InceptionV3 / MobileNetV2 go through cn_conv_fwd with their inference BatchNorm folded
 * into kernel and bias by the host (netspec.fold_batchnorm); these are the remaining layers. */
#define CN_POOL_MAX 0        /* MaxPooling2D: maximum over the in-bounds part of the window                       */
#define CN_POOL_AVG_VALID 1  /* AveragePooling2D(padding="same"): mean over the in-bounds elements (TF leaves the
                                padding out of the count)                                                          */
/* y[n,oy,ox,0:c] = pool over the kh x kw window whose corner is (oy*stride - pad_t, ox*stride - pad_l); output pixel
 * records are ldy >= c floats apart, so a branch of an Inception block can write its slice of the concatenated tensor.
 * replaces MaxPooling2D / AveragePooling2D inside keras.applications.InceptionV3 (inception_distance.py:11). */
int cn_pool2d_fwd(const float* x, int n, int h, int w, int c, int kh, int kw, int stride, int pad_t, int pad_l,
                  int oh, int ow, int mode, float* y, int ldy, void* stream);
/* y = act(depthwise_conv3x3_same(x, wk) + bias): wk (3,3,c) = the Keras DepthwiseConv2D kernel (3,3,c,1), stride 1 or 2,
 * TF SAME geometry.  replaces DepthwiseConv2D + BatchNormalization + ReLU(6.) inside keras.applications.MobileNetV2
 * (celeba_attribute_prediction.py:55). */
int cn_dwconv3x3_fwd(const float* x, const float* wk, const float* bias, int n, int h, int w, int c, int stride,
                     int act, float alpha, float* y, void* stream);
/* y = act(x), act in {CN_ACT_RELU6, CN_ACT_SIGMOID}; x == y allowed.  ReLU(6.) behind a 1x1 conv of MobileNetV2 = the conv
 * with CN_ACT_RELU in its epilogue, then this clamp. */
int cn_act_ext(const float* x, float* y, int64_t n, int act, void* stream);
/* cv2.resize(img, (ow, oh)) with the default INTER_LINEAR on channels-last images: uint8 (is_u8 = 1, OpenCV's fixed-point
 * form, bit-exact) or float32.  replaces celeba_attribute_prediction.py:131-136. */
int cn_resize_bilinear(const void* x, int n, int h, int w, int c, int oh, int ow, int is_u8, void* y, void* stream);
/* y = (float)x for uint8 x (keras-applications preprocess_input's astype, celeba_attribute_prediction.py:138) */
int cn_u8_to_f32(const void* x, float* y, int64_t n, void* stream);
/* mode 0: y = (x + 1) * 127.5 (celeba_attribute_prediction.py:129-130); mode 1: y = x / 127.5 - 1 (preprocess_input
 * mode "tf": inception_distance.py:24, celeba_attribute_prediction.py:138); one rounding per NumPy operation */
int cn_pixel_map(const float* x, float* y, int64_t n, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif
