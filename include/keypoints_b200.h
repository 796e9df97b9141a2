/* keypoints_b200.h — C ABI of the B200-native unsupervised-keypoint training path.
 *
 * The reference (DuaneNielsen/keypoints) has no FFI of its own: its hot path is a chain of
 * ATen calls issued from Python.  This header is the boundary a maintainer binds
 * instead (ctypes stub in INTEGRATION.md); every entry cites the reference call site it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.
 *  - every function returns 0 on success, a negative kp_status otherwise; kp_last_error() returns
 *    a thread-local message.  Nothing throws across the ABI.
 *  - all data pointers are DEVICE pointers borrowed from the caller (the caller's allocator owns
 *    every buffer, including workspaces); every call is asynchronous on `stream`.
 *  - 4-D tensors are passed as kp_view: base pointer + element strides for (n, y, x, c) + dtype,
 *    so NCHW API tensors, NHWC internal activations, channel slices and replicate-padded buffers
 *    (pointer at the padded origin) all use the same kernels.
 *  - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef KEYPOINTS_B200_H
#define KEYPOINTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* the library is built with -fvisibility=hidden: only the entry points declared here are exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef void* kp_stream;                 /* cudaStream_t */

enum kp_status { KP_OK = 0, KP_ERR_ARG = -1, KP_ERR_CUDA = -2, KP_ERR_UNSUPPORTED = -3 };
enum kp_dtype  { KP_F32 = 0, KP_BF16 = 1 };
enum kp_act    { KP_ACT_NONE = 0, KP_ACT_LEAKY = 1, KP_ACT_RELU = 2 };   /* LeakyReLU slope 0.01 */
enum kp_post   { KP_POST_NONE = 0, KP_POST_POOL = 1, KP_POST_UP = 2 };   /* MaxPool2x2 / bilinear x2 */
/* OR-ed into `post` of the forward BatchNorm passes: the caller runs this pass on one stream while tensor-core convolutions
 * run on another, so the small-shared-memory variant that can share an SM with a persistent conv CTA is preferred. */
#define KP_POST_CORESIDENT 0x100

typedef struct kp_view {
    void*   ptr;                         /* element (n=0,y=0,x=0,c=0) */
    int64_t sn, sy, sx, sc;              /* strides in elements */
    int32_t dtype;                       /* kp_dtype */
    int32_t _pad;
} kp_view;

const char* kp_last_error(void);
int kp_version(void);
/* SM count, cc major/minor of the current device; fails loudly if there is no sm_100 GPU. */
int kp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Optional per-device context.  kp_ctx_create binds to the CURRENT CUDA device (fails unless it is sm_100);
 * kp_ctx_set_current makes it the calling thread's context for every later kp_* call (NULL = none).  A context holds
 *  - the SM budget of the persistent kernels (kp_ctx_set_sm_limit: leave SMs free, e.g. for a concurrent collective;
 *    0 = all), and
 *  - a cache of encoded TMA tensor maps keyed by (buffer, shape, box): with the static buffers of a training loop the
 *    tensor-core launchers encode each map once instead of on every call (kp_ctx_info reports hits / misses).
 * Without a current context the calls behave as before (all SMs, maps encoded per call).  One context per thread/GPU;
 * the context itself is internally locked, the entry points stay asynchronous on `stream`. */
typedef struct kp_ctx kp_ctx;
int kp_ctx_create(kp_ctx** ctx);
int kp_ctx_destroy(kp_ctx* ctx);
int kp_ctx_set_current(kp_ctx* ctx);
int kp_ctx_set_sm_limit(kp_ctx* ctx, int sms);
int kp_ctx_info(kp_ctx* ctx, int* device, int* sm_count, int* sm_limit, int64_t* map_hits, int64_t* map_misses,
                int64_t* maps_cached);

/* ---------------------------------------------------------------------------------------------
 * Convolution, direct fp32 (CUDA cores).  Parity-grade path and the path for thin layers
 * (Cin or Cout not a multiple of 64).
 *   out[n,oy,ox,co] = bias[co] + sum_{t,ci} in[n, oy+ty+off, ox+tx+off, ci] * wk[t][ci][co]
 * with zero for reads outside [0,IH)x[0,IW).  fprop on a replicate-padded input: off=0, IH=OH+2;
 * dgrad w.r.t. the padded input: in=dY, taps flipped in wk, off=-(ks-1), IH=OH-(ks-1).
 * stats (nullable): double[2*Cout], += sum and sum of squares of out over (n,oy,ox) — the
 * train-mode BatchNorm batch statistics.
 * Replaces nn.Conv2d forward / convolution_backward(input) (vgg.py:33, knn.py:115,123). */
int kp_conv_simt(kp_stream stream, const kp_view* in, const float* wk, const float* bias,
                 const kp_view* out, double* stats, int N, int OH, int OW, int IH, int IW,
                 int Cin, int Cout, int ks, int off);

/* Weight gradient, direct fp32.  dw_oihw[co][ci][t] += sum_{n,y,x} x[n,y+ty,x+tx,ci] * dy[n,y,x,co]
 * (x = pointer at the padded origin for ks=3, interior for ks=1).  dw must be zeroed by the caller.
 * Replaces convolution_backward(weight). */
int kp_conv_wgrad_simt(kp_stream stream, const kp_view* x, const kp_view* dy, float* dw_oihw,
                       int N, int H, int W, int Cin, int Cout, int ks);

/* ---------------------------------------------------------------------------------------------
 * Convolution on the 5th-gen tensor cores (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM,
 * operands staged by TMA, 128-byte swizzle).  "Flat pixel" formulation: an activation buffer
 * [N][PH][PW][C] (replicate- or zero-bordered, PH=H+2, PW=W+2) is the matrix [Q=N*PH*PW][C]; with
 * input and output on the same pitch a tap is a constant shift of the flat pixel index:
 *     out[q][co] = bias[co] + sum_t sum_ci in[q + shifts[t]][ci] * wt[t][co][ci]       q in [0,Q)
 * reads outside [0,Q) are zero (TMA fill).  Cin % 64 == 0, Cout % 64 == 0, all bf16, out dense
 * [Q][Cout].  stats (nullable): double[2*Cout] += sum / sum of squares over the rows with
 * (q % PW) < VW and (q / PW % PH) < VH (the BatchNorm batch statistics of the valid outputs).
 *   fprop 3x3: in = replicate-padded X, shifts[t] = ty*PW+tx, out top-left aligned (VH=H, VW=W)
 *   fprop 1x1: shifts[0] = PW+1
 *   dgrad    : in = dY interior-aligned with a zero border, wt = tc_d of kp_pack_weights,
 *              shifts[t'] = (t'y-1)*PW + (t'x-1) (1x1: 0); every out row is valid.
 * Replaces nn.Conv2d forward / convolution_backward(input) (vgg.py:33, knn.py:115,123). */
int kp_conv_tc(kp_stream stream, const void* in_bf16, int64_t Q, int Cin, const void* wt_bf16,
               int taps, const int32_t* shifts, const float* bias, void* out_bf16, int Cout,
               double* stats, int PH, int PW, int VH, int VW);

/* Tensor-core weight gradient (both operands MN-major through UMMA shared-memory descriptors):
 *   dw_oihw[co][ci][t] += sum_q dy[q][co] * x[q + shifts[t]][ci]
 * x: bf16 [Q][CinP] (CinP >= Cin channel pitch), dy: bf16 [Q][Cout], zero wherever a row must not
 * count (the zero border of the interior-aligned dY).  Split over pixel ranges, fp32 vector
 * reductions into stg (workspace, taps*Cout*CinP floats), then folded into the OIHW gradient
 * bucket.  3x3: shifts[t] = (ty-1)*PW + (tx-1); 1x1: 0.
 * Replaces convolution_backward(weight). */
int kp_conv_wgrad_tc(kp_stream stream, const void* x_bf16, const void* dy_bf16, int64_t Q, int Cin,
                     int CinP, int Cout, int taps, const int32_t* shifts, float* stg,
                     float* dw_oihw);
/* Deferred folding: kp_conv_wgrad_tc / _img called with dw_oihw == NULL neither zero `stg` nor fold it; the caller zeroes
 * its staging arena once per step and folds all layers with one launch.  table_dev: n_layers rows of 7 int64
 * {stg, dw_oihw, taps, Cout, Cin, CinP, Cout % 128 == 0} in device memory. */
int kp_wgrad_finalize_multi(kp_stream stream, const void* table_dev, int n_layers);
/* Image-aware form of the same gradient: x = replicate-padded input [N][H+2][W+2][CinP], dy = interior-aligned
 * gradient [N][H+2][W+2][Cout]; the contraction runs over the N*H*W valid pixels only (4-D TMA boxes of bx x by = 64
 * pixels), so no MMA work is spent on the pad / zero-border pixels the flat form multiplies (3 % of the pixels at
 * 128x128, 27 % at 16x16).  W must be a power of two >= 16 (W < 64: H % (64 / W) == 0); ks in {1, 3}. */
int kp_conv_wgrad_tc_img(kp_stream stream, const void* x_bf16, const void* dy_bf16, int N, int H, int W,
                         int Cin, int CinP, int Cout, int ks, float* stg, float* dw_oihw);

/* Repack OIHW fp32 master weights for the kernels above (any output may be NULL):
 *  simt_f   f32  [t][ci][co]            simt_d  f32  [t'][co][ci]   (t' = taps flipped)
 *  tc_f     bf16 [t][co][ci_pad]        tc_d    bf16 [t'][ci_pad][co]
 * ci_pad >= Cin pads the input-channel axis with zeros (KeyNet decoder: 74 -> 128). */
int kp_pack_weights(kp_stream stream, const float* w_oihw, int Cout, int Cin, int ks, int ci_pad,
                    float* simt_f, float* simt_d, void* tc_f, void* tc_d);
/* The same for n_layers tensors in ONE launch.  table_dev: device array of n_layers records of nine 64-bit
 * fields {w, simt_f, simt_d, tc_f, tc_d (pointers, 0 = skip), Cout, Cin, ks, ci_pad}. */
int kp_pack_weights_multi(kp_stream stream, const void* table_dev, int n_layers);

/* ---------------------------------------------------------------------------------------------
 * Train-mode BatchNorm statistics -> per-channel affine (nn.BatchNorm2d, vgg.py:35, knn.py:117).
 * stats = double[2*C] (sum, sumsq) over `count` elements per channel.  Writes scale = gamma*invstd,
 * shift = beta - mean*scale, saves mean / invstd for backward, and updates running_mean /
 * running_var (momentum, unbiased variance) and num_batches_tracked (int64, nullable). */
int kp_bn_finalize(kp_stream stream, const double* stats, int C, double count, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean,
                   float* running_var, int64_t* num_batches_tracked, float* scale, float* shift,
                   float* save_mean, float* save_invstd);

/* Fused normalise + activation + (MaxPool2x2 | bilinear x2, align_corners=True) + replicate-pad
 * write.  y: [N,H,W,C] view.  out: view whose ptr is the origin of the (padded if pad=1) output of
 * interior size OHxOW (H,W | H/2,W/2 | 2H,2W).  scale/shift nullable (identity).  Also used as the
 * generic layout / dtype converter (NCHW f32 <-> padded NHWC bf16).
 * Replaces native_batch_norm + leaky_relu_/relu_ + max_pool2d / upsample_bilinear2d +
 * replication_pad2d of the next layer (vgg.py:24-37, knn.py:115-125). */
int kp_bn_act_fwd(kp_stream stream, const kp_view* y, const kp_view* out, const float* scale,
                  const float* shift, int act, int post, int pad, int N, int H, int W, int C);

/* Backward of the above, two passes.  dout: gradient w.r.t. `out` (ptr at the padded origin if pad=1: border
 * gradients are folded into the edge pixels = replication_pad2d backward).
 * Pass 1 (reduce): dz = d/d(BatchNorm output) — fold + pool/upsample backward + activation backward — is written to
 * dy (nullable), and sums[0:C] += sum dz, sums[C:2C] += sum dz * xhat (double).  For layers without BatchNorm
 * (scale/shift/mean/invstd NULL) dy is the final gradient and sums[0:C] the bias gradient.
 * Pass 2 (apply, BatchNorm only), in place: dy = scale * (dy - sums[0]/count - xhat * sums[1]/count);
 * dgamma / dbeta (nullable) receive sums[C:2C] / sums[0:C] (fused kp_bn_grad_finalize). */
int kp_bn_act_bwd_reduce(kp_stream stream, const kp_view* dout, const kp_view* y, const kp_view* dy,
                         const float* scale, const float* shift, const float* mean,
                         const float* invstd, double* sums, int act, int post, int pad, int N,
                         int H, int W, int C);
int kp_bn_act_bwd_apply(kp_stream stream, const kp_view* y, const kp_view* dy, const float* scale,
                        const float* mean, const float* invstd, const double* sums, double count,
                        int N, int H, int W, int C, float* dgamma, float* dbeta);
/* Alternative pass 2 for post = none / pool on the vectorised NHWC path: repeats the gather of pass 1 (called with
 * dy = NULL) and writes dy directly, so dz never round-trips through HBM.  KP_ERR_UNSUPPORTED if not applicable. */
int kp_bn_act_bwd_apply_gather(kp_stream stream, const kp_view* dout, const kp_view* y,
                               const kp_view* dy, const float* scale, const float* shift,
                               const float* mean, const float* invstd, const double* sums,
                               double count, int act, int post, int pad, int N, int H, int W, int C,
                               float* dgamma, float* dbeta);
/* kp_bn_finalize + kp_bn_act_fwd in one launch when the vectorised path applies (otherwise the two
 * kernels are launched back to back): every block derives its channels' affine from the batch
 * statistics, block 0 publishes scale/shift/mean/invstd and updates the running statistics. */
int kp_bn_finalize_act_fwd(kp_stream stream, const double* stats, double count, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float* scale,
                           float* shift, float* save_mean, float* save_invstd, const kp_view* y,
                           const kp_view* out, int act, int post, int pad, int N, int H, int W,
                           int C);
/* dgamma[c] = sums[C+c], dbeta[c] = sums[c] (any nullable; fp32, overwritten). */
int kp_bn_grad_finalize(kp_stream stream, const double* sums, int C, float* dgamma, float* dbeta);

/* ---------------------------------------------------------------------------------------------
 * Bottleneck.  heat / m are fp32 NCHW planes [N*K][h][w]; k is [N*K][2] ordered (y, x).
 * spacial_logsoftmax / spacial_softmax (functional.py:27-44): two 1-D softmaxes of the row means
 * and the column means, expectation against linspace(0,1). */
int kp_spatial_softmax_fwd(kp_stream stream, const float* heat, int planes, int h, int w,
                           float* k, float* p_h, float* p_w);
int kp_spatial_softmax_bwd(kp_stream stream, const float* dk, const float* k, const float* p_h,
                           const float* p_w, int planes, int h, int w, float* dheat);
/* gaussian_like_function (functional.py:56-63): m = exp(-sqrt(dy^2+dx^2+eps)/(2 sigma^2)). */
int kp_gaussian_fwd(kp_stream stream, const float* k, int planes, int h, int w, float sigma,
                    float eps, float* m);
/* dk[p] = sum_ij dm * m * (coord - k)/(2 sigma^2 r).  dm is a view [N,h,w,K] (pad=1: ptr at the
 * padded origin, border folded).  argmax (nullable, int32 [N][h][w]): Transporter 'max' combine —
 * dm has one channel and only pixels whose arg-max keypoint is this plane contribute. */
int kp_gaussian_bwd(kp_stream stream, const kp_view* dm, int pad, const float* k, const int32_t* argmax,
                    int N, int K, int h, int w, float sigma, float eps, float* dk);
/* Transporter feature transport, mode 'max' (models/transporter.py:57-60):
 * M_s = max_k m_s, M_t = max_k m_t, out = phi_s (1-M_s)(1-M_t) + phi_t M_t written replicate-padded.
 * mask_s / mask_t: f32 [N][h][w]; argmax_t: int32 [N][h][w]. */
int kp_transport_fwd(kp_stream stream, const kp_view* phi_s, const kp_view* phi_t, const float* k_s,
                     const float* k_t, const kp_view* out, int pad, float* mask_s, float* mask_t,
                     int32_t* argmax_t, int N, int h, int w, int C, int K, float sigma, float eps);
/* dphi_t = dout * M_t ; dmask_t[n,i,j] = sum_c dout * (phi_t - phi_s (1-M_s)). */
int kp_transport_bwd(kp_stream stream, const kp_view* dout, int pad, const kp_view* phi_s,
                     const kp_view* phi_t, const float* mask_s, const float* mask_t,
                     const kp_view* dphi_t, float* dmask_t, int N, int h, int w, int C);

/* The other combine modes of TransporterNet.forward (models/transporter.py:41-50) and a mode-generic backward.
 * mode: KP_COMBINE_MAX | KP_COMBINE_SUM ('sum_and_clamp') | KP_COMBINE_LOOP ('loop').
 * aux: int32 [N][h][w] (max: arg-max keypoint of the target; sum: 1 where the unclamped target sum lies in [0,1]);
 * coef: f32 [N][h][w][2], loop mode only: phi = phi_s*coef0 + phi_t*coef1.  In loop mode mask_s / mask_t receive the maps
 * of the LAST keypoint (what the reference returns).  The backward writes dphi_t and dm_t (f32 [N][K][h][w], the
 * gradient w.r.t. every rendered target map), to be reduced to d(keypoints) by kp_gaussian_bwd. */
#define KP_COMBINE_MAX 0
#define KP_COMBINE_SUM 1
#define KP_COMBINE_LOOP 2
int kp_transport_mode_fwd(kp_stream stream, int mode, const kp_view* phi_s, const kp_view* phi_t, const float* k_s,
                          const float* k_t, const kp_view* out, int pad, float* mask_s, float* mask_t, int32_t* aux,
                          float* coef, int N, int h, int w, int C, int K, float sigma, float eps);
int kp_transport_mode_bwd(kp_stream stream, int mode, const kp_view* dout, int pad, const kp_view* phi_s,
                          const kp_view* phi_t, const float* k_s, const float* k_t, const float* mask_s,
                          const float* mask_t, const int32_t* aux, const float* coef, const kp_view* dphi_t,
                          float* dm_t, int N, int h, int w, int C, int K, float sigma, float eps);

/* l2_reconstruction_loss (transporter.py:56-60, keypoints.py:54-58) and its gradient:
 * loss[0] += sum (xhat-target)^2 * mask (double, caller zeroes; mean = /numel on the host side or
 * via kp_scale), dxhat = 2 (xhat-target) mask * gscale.  fp32 contiguous, mask nullable. */
int kp_l2_loss(kp_stream stream, const float* xhat, const float* target, const float* mask,
               int64_t numel, float gscale, double* loss_sum, float* dxhat);

/* ---------------------------------------------------------------------------------------------
 * Augmentation (tps.py:10-87,128-131,154-166; data_augments.py:13-16).  fp32 NCHW.
 * TPS: grid = ((x,y) + a0 + a1 x + a2 y + sum_t w_t U(|p-c_t|)) * 2 - 1, U(d) = d^2 log(d+1e-6),
 * evaluated per output pixel in registers (never materialised), then bilinear grid_sample with
 * zeros padding, align_corners=False.  theta: [N][T+3][2] (or [N][T+2][2] if reduced), ctrl [N][T][2].
 * x == NULL samples the constant image 1 (the start of the loss-mask chain, data_augments.py:33). */
int kp_tps_warp(kp_stream stream, const float* x, float* out, const float* theta, const float* ctrl,
                int N, int C, int H, int W, int T, int reduced);
/* Parameters of `draws` TpsAndRotate perturbations per image, drawn on the device (rand_peturb_params,
 * data_augments.py:6-10; tps_sample_params, tps.py:122-125): theta ~ N(0, variance^2) [draws][N][T+3][2],
 * ctrl ~ U[0,1) [draws][N][T][2], rot ~ U(-max_rotate, max_rotate) [draws][N].  Philox4x32-10 keyed by `seed`,
 * counter = (draw*N+image, block, *step_dev): replaying a captured launch draws fresh parameters every step
 * (step_dev is the device step counter kp_adam_step advances; NULL = step 0).  The reference draws on the
 * CPU global RNG; the distribution, not the stream, is what is reproduced here (the parity tests pass the
 * reference's own draws to kp_tps_warp / kp_rotate_warp). */
int kp_aug_draw(kp_stream stream, uint64_t seed, const int32_t* step_dev, int draws, int N, int T,
                float variance, float max_rotate, float* theta, float* ctrl, float* rot);
/* Input conversion on the device (transforms.ToTensor + Normalize, datasets.py:287-300): uint8 NHWC frames as the
 * loader ships them over PCIe -> fp32 NCHW, out = in * scale + bias (grey/color: scale 2/255, bias -1; celeba: 1/255, 0). */
int kp_u8_to_f32(kp_stream stream, const uint8_t* in, float* out, int N, int H, int W, int C, float scale, float bias);
/* Real-data ingestion (datasets.py:237-254 ImageFolder + celeba_transform :297-300), SURVEY.md 8f.4.
 * kp_jpeg_*: nvJPEG decode of one host JPEG stream into an interleaved RGB uint8 device image (libnvjpeg is dlopen'ed at
 * the first kp_jpeg_create; KP_ERR_UNSUPPORTED if it is absent).  A context is not thread-safe: one per decoding thread.
 * kp_resize_to_f32: transforms.Resize((oh,ow)) of a PIL image (Pillow's antialiased BILINEAR: triangle filter of support
 * max(scale,1), horizontal pass then vertical pass, 8-bit rounding after each) + ToTensor: uint8 [H][W][C] ->
 * fp32 [C][oh][ow] in [0,1], i.e. straight into one slot of the NCHW batch.  tmp: uint8 [H][ow][C] scratch. */
int kp_jpeg_create(void** ctx);
int kp_jpeg_destroy(void* ctx);
int kp_jpeg_info(void* ctx, const uint8_t* data, int64_t len, int* width, int* height, int* components);
int kp_jpeg_decode(void* ctx, kp_stream stream, const uint8_t* data, int64_t len, uint8_t* out_rgb, int width, int height);
int kp_resize_to_f32(kp_stream stream, const uint8_t* in_hwc, int in_h, int in_w, int channels, uint8_t* tmp,
                     float* out_chw, int out_h, int out_w);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (gradient / statistics accumulators; graph-capturable). */
int kp_zero(kp_stream stream, void* ptr, int64_t bytes);
/* rotate_affine_grid_multi: affine_grid([[cos,sin,0],[-sin,cos,0]]) + grid_sample. rot: [N]. */
int kp_rotate_warp(kp_stream stream, const float* x, float* out, const float* rot, int N, int C,
                   int H, int W);

/* ---------------------------------------------------------------------------------------------
 * Optimiser: torch.optim.Adam defaults (transporter.py:47) over a flat fp32 bucket.
 * g is multiplied by grad_scale first (1/world_size after the NCCL sum).  step_dev (nullable):
 * device-resident step counter, incremented by the call and used for the bias corrections instead
 * of `step` — lets the whole train step replay as a CUDA graph. */
int kp_adam_step(kp_stream stream, float* p, const float* g, float* m, float* v, int64_t n,
                 double lr, double beta1, double beta2, double eps, int step, float grad_scale,
                 int32_t* step_dev);

/* Data-parallel optimiser step over NVLink peer memory (the reference has no distributed code).
 * Gradient reduce-scatter + Adam + parameter all-gather in one kernel: rank r owns elements
 * [r*ceil(n/4/world)*4, ...) of the flat bucket, sums that slice of every replica's gradient buffer (peer loads, or
 * multimem.ld_reduce through the NVSwitch when use_multicast), updates its slice of m / v (Adam, same arithmetic as
 * kp_adam_step with the device step counter) and writes the new parameters into EVERY replica's buffer (peer stores /
 * multimem.st).  peers->g[j], p[j]: replica j's gradient / parameter buffers mapped into this process (symmetric
 * memory); flag[j]: replica j's int32[KP_DP_MAX_WORLD] barrier flags (zero-initialised); mc_g / mc_p: multicast
 * addresses of the two buffers (NULL without NVLS).  epoch_dev: device int32 barrier epoch (starts at 0, advanced by
 * the call; every replica must make the same sequence of calls).  Bracketed by two flag barriers, so on return (in
 * stream order) all replicas hold identical parameters and the gradient buffers may be overwritten. */
#define KP_DP_MAX_WORLD 8
typedef struct kp_dp_peers {
    float*   g[KP_DP_MAX_WORLD];
    float*   p[KP_DP_MAX_WORLD];
    int32_t* flag[KP_DP_MAX_WORLD];
    float*   mc_g;
    float*   mc_p;
} kp_dp_peers;
int kp_dp_adam_step(kp_stream stream, const kp_dp_peers* peers, int rank, int world, int64_t n, float* m, float* v,
                    double lr, double beta1, double beta2, double eps, float grad_scale, int32_t* step_dev,
                    int32_t* epoch_dev, int use_multicast);

/* Non-blocking loss log (replaces the per-step `loss.item()` host sync of ResultsLogger.log, utils.py:122-130):
 * ring[2*(s % slots)] = s, ring[2*(s % slots)+1] = *loss_sum * scale with s = *step_dev (0 if NULL); ring is
 * double[2*slots] in device memory.  The host copies the ring back every k steps on a side stream and reads the
 * (step, loss) pairs without ever waiting for the step in flight (keypoints_b200/runlog.py). */
int kp_loss_ring_push(kp_stream stream, const double* loss_sum, double scale, const int32_t* step_dev,
                      double* ring, int slots);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* KEYPOINTS_B200_H */
