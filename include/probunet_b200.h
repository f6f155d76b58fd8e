/*
 * probunet_b200 -- C ABI of the B200-native (sm_100a) Probabilistic U-Net hot path.
 *
 * Plain C: pointers + sizes, no torch types, no exceptions across the boundary.
 * Every entry point ENQUEUES work on the given CUDA stream and returns immediately
 * (0 = ok, negative = error; pub_last_error() gives the message).  The caller owns
 * every buffer (inputs, outputs, gradients, workspaces) -- in the Python host layer
 * they are torch allocations passed by data_ptr().
 *
 * The reference (MaryamAlipourH/prob-unet-climate-downscaling) has no FFI of its own:
 * its boundary is the nn.Module API of src/prob_unet.py + src/networks.py.  Each entry
 * point below therefore cites the reference Python function whose arithmetic it
 * replaces (file:line relative to the reference root).
 *
 * Layouts:  "NCHW f32" is the reference's public tensor layout; "NHWC dt" is the
 * engine-internal activation layout, dt = PUB_F32 or PUB_BF16, with an explicit
 * pixel stride `ld` (elements) so that channel slices of a wider buffer are views.
 */
#ifndef PROBUNET_B200_H
#define PROBUNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pub_stream_t; /* == cudaStream_t */

#define PUB_F32 0
#define PUB_BF16 1
#define PUB_TF32 2 /* f32 storage whose values are rounded to tf32; convolutions use tcgen05 kind::tf32 */
/* pub_encoder_create only: PUB_TF32 with the full-resolution stage (its three convs, their gradients) in PUB_BF16; the
 * first max-pool converts.  Rounding in that stage barely moves log sigma, rounding behind it does. */
#define PUB_TF32_BF16S0 3

#define PUB_BACKEND_AUTO 0
#define PUB_BACKEND_SIMT 1    /* fp32-FMA implicit GEMM (parity path, any shape)          */
#define PUB_BACKEND_TCGEN05 2 /* tcgen05.mma + TMEM + TMA implicit GEMM (bf16, C % 32 == 0) */

const char* pub_last_error(void);
int pub_version(void);
/* number of kernels this library has enqueued so far in this process (bench.py: gpu_launches) */
unsigned long long pub_launch_count(void);
/* bring-up / A-B measurement knobs (tools/ only; never needed for correct results).
 *   "conv_halo" 0|1|2 : 3x3 tcgen05 convolutions through the per-tap TMA kernel (0), the halo kernel where it measured
 *                       faster (1, default) or wherever it is legal (2)
 *   "fcomb_fwd_mma" 0|1|2|3 : fcomb forward in bf16 mode with the f32 FMA kernel (0), the tensor-core kernels (1, default:
 *                       bf16 hi + lo split for M <= 32, tf32 for larger ensembles), always tf32 (2), always the split (3)
 *   "wgrad_fused_bias" 0|1 : bias gradient from a separate column-sum pass over dy (0) or from the dy tiles the wgrad
 *                       kernel stages in shared memory anyway (1, default)
 *   "wgrad_box3" 0|1  : 3x3 weight gradient with nine tap boxes (0) or three (8+2) x 8 boxes + row-offset taps (1, default)
 *   "gn_fuse" 0|1     : GroupNorm statistics / backward prologue as separate passes (0) or fused into the epilogues of the
 *                       tcgen05 halo convolutions that produce the tensor / the data gradient (1, default) */
int pub_debug_option(const char* name, int value);
/*   "halo_trace" : device buffer of 3 x 1024 int64 that CTA (0,0) of conv_halo_kernel fills with clock64 stamps
 *                  (producer / MMA issuer / epilogue events; tools/halo_trace.py); NULL switches it off */
/*   "seed_salt"  : device uint32 xor-ed into every dropout key / rsample seed (CUDA-graph replays, see the end of this
 *                  file); NULL (default) switches it off */
int pub_debug_pointer(const char* name, void* p);

/* ------------------------------------------------------------------------------------
 * Convolution primitives.  Replace torch.nn.functional.conv2d at src/networks.py:89
 * (networks.Conv2d, 3x3 / 1x1, stride 1, same padding, bias add :90-91) and
 * nn.Conv2d(k=3,p=1)+ReLU at src/prob_unet.py:41-46, plus what autograd derives from
 * them (convolution_backward: dgrad + wgrad + bias grad).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void* x0; int32_t c0, ld0; /* input, NHWC dt, c0 channels, pixel stride ld0          */
  const void* x1; int32_t c1, ld1; /* optional 2nd input = virtual channel concat (:329)     */
  const void* w;                   /* packed weights [k*k][cout][c0+c1] dt (pub_pack_conv_weight) */
  const float* bias;               /* [cout] or NULL                                          */
  const void* res; int32_t ld_res; /* optional residual added in the epilogue (:178)          */
  const void* mask; int32_t ld_mask; /* optional: y = 0 where mask <= 0 (ReLU backward)        */
  void* y; int32_t ldy;            /* output NHWC dt, cout channels                            */
  int32_t B, H, W, cout, ksize;    /* ksize 1 or 3                                             */
  int32_t relu;                    /* 1: ReLU in the epilogue (src/prob_unet.py:42,46)         */
  int32_t dtype, backend;
} pub_conv_args;
int pub_conv2d_forward(const pub_conv_args* a, pub_stream_t s);

/* The same launch with GroupNorm work fused into its epilogue (bf16, 3x3, shapes pub_conv2d_fused_rows() accepts):
 *   gn_bwd == 0: stat_part rows [B * rows][cout][2] receive per-channel (sum, sum of squares) of the stored output -- the
 *                statistics partials of the GroupNorm that reads y (networks.GroupNorm, src/networks.py:105-107);
 *   gn_bwd == 1: a data-gradient launch whose result g = dL/d(dropout(silu(a x + b))) is stored as
 *                du = g * keep/(1-p) * silu'(a x + b) (the backward of src/networks.py:168,173,177), x = the GroupNorm's
 *                input (gx0 | gx1 virtual concat, gc0 channels in gx0), gcoef = its [B][cout][2] affine table (a, b) as
 *                written by pub_groupnorm_silu_forward; the rows receive (sum du, sum du * x).
 * rows = pub_conv2d_fused_rows(a): partial rows per image (0: this shape / dtype / backend has no fused epilogue). */
typedef struct {
  float* stat_part;
  int32_t gn_bwd;
  const void* gx0; const void* gx1; int32_t gc0, gld0, gld1;
  const float* gcoef;
  float p_drop; uint64_t seed, subseq;
} pub_conv_gn_args;
int pub_conv2d_fused_rows(const pub_conv_args* a);
int pub_conv2d_forward_fused(const pub_conv_args* a, const pub_conv_gn_args* g, pub_stream_t s);

/* OIHW f32 master weight -> packed [tap][cout][cin] dt.  transpose_flip=1 builds the
 * data-gradient operator (taps mirrored, in/out channels swapped): [tap][cin][cout].  */
int pub_pack_conv_weight(const float* w_oihw, void* packed, int cout, int cin, int ksize,
                         int dtype, int transpose_flip, pub_stream_t s);

typedef struct {
  const void* x0; int32_t c0, ld0;
  const void* x1; int32_t c1, ld1;
  const void* dy; int32_t ld_dy;   /* [B,H,W,cout] NHWC dt                                     */
  float* dw;                       /* OIHW f32 [cout][c0+c1][k][k]                             */
  float* dbias;                    /* [cout] or NULL                                           */
  void* workspace; size_t workspace_bytes;
  int32_t B, H, W, cout, ksize;
  int32_t accumulate;              /* 0: overwrite dw/dbias, 1: add                            */
  int32_t dtype, backend;
} pub_wgrad_args;
size_t pub_conv2d_wgrad_workspace(const pub_wgrad_args* a);
int pub_conv2d_wgrad(const pub_wgrad_args* a, pub_stream_t s);

/* layout glue at the module boundary (NCHW f32 <-> NHWC dt), x1 optional (cat([x,target]),
 * src/prob_unet.py:67) */
int pub_nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* y, int ldy,
                     int B, int H, int W, int dtype, pub_stream_t s);
int pub_nhwc_to_nchw(const void* x, int ld, int C, float* y, int B, int H, int W, int dtype,
                     int accumulate, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * GroupNorm (+FiLM) + SiLU (+dropout) (+2x box resample) as ONE op: networks.GroupNorm.forward
 * (F.group_norm, groups = min(32, C/4), eps 1e-5; src/networks.py:97-107) followed by what UNetBlock.forward does
 * with it -- silu(norm0(x)) (:168) or silu(shift + norm1(x) * (scale + 1)) (:170-173), F.dropout (:177) and the
 * depthwise box resample of the following conv (:83-87) -- and its autograd.  x / y / dy / dx are NHWC dt
 * (x with pixel stride ld, the others contiguous); film = [2C] (scale | shift) or NULL; resample 0 none, 1 down
 * (2x2 mean), 2 up (nearest 2x).  stats [B,G,2] and coef [B,C,2] are written by forward and read by backward.
 * ---------------------------------------------------------------------------------- */
size_t pub_groupnorm_scratch_bytes(int B, int C, int H, int W);
int pub_groupnorm_silu_forward(const void* x, int C, int ld, int B, int H, int W, const float* gamma, const float* beta,
                               const float* film, int resample, float p_drop, uint64_t seed, uint64_t subseq,
                               void* y, float* stats, float* coef, void* scratch, size_t scratch_bytes, int dtype,
                               pub_stream_t s);
int pub_groupnorm_silu_backward(const void* x, int C, int ld, int B, int H, int W, const float* gamma, const float* beta,
                                const float* film, int resample, float p_drop, uint64_t seed, uint64_t subseq,
                                const float* stats, const float* coef, const void* dy, void* dx, float* dgamma,
                                float* dbeta, float* dfilm, void* scratch, size_t scratch_bytes, int dtype,
                                pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * U-Net engine: networks.UNet.forward (src/networks.py:299-333) and its autograd.
 * The block list is supplied by the host (it mirrors the ModuleDict order, :321,:327).
 * params/grads are pointer tables in this order:  for every entry -- plain conv:
 * {weight,bias}; UNetBlock: {norm0.w, norm0.b, conv0.w, conv0.b, affine.bias, norm1.w,
 * norm1.b, conv1.w, conv1.b, [skip.w, skip.b if has_skip_conv]} -- then {out_norm.w,
 * out_norm.b, out_conv.w, out_conv.b}.  All f32, reference (OIHW) layouts.
 * ---------------------------------------------------------------------------------- */
typedef struct { int32_t cin, cout, up, down, has_skip_conv, is_conv; } pub_unet_block;
typedef struct pub_unet pub_unet;

int pub_unet_create(const pub_unet_block* enc, int n_enc, const pub_unet_block* dec, int n_dec,
                    int in_channels, int out_channels, float dropout, int dtype, pub_unet** out);
void pub_unet_destroy(pub_unet* u);
int pub_unet_num_params(const pub_unet* u);
size_t pub_unet_workspace_bytes(const pub_unet* u, int B, int H, int W);
/* out: NCHW f32 [B,out,H,W] if out_nchw != 0, else NHWC dt [B,H,W,out] */
int pub_unet_forward(pub_unet* u, int B, int H, int W, const float* x_nchw,
                     const float* const* params, void* out, int out_nchw,
                     void* workspace, size_t workspace_bytes,
                     uint64_t seed, int training, int backend, pub_stream_t s);
/* dout has the layout `out` had; grads are OVERWRITTEN; dx_nchw may be NULL */
int pub_unet_backward(pub_unet* u, int B, int H, int W, const void* dout, int dout_nchw,
                      const float* const* params, float* const* grads, float* dx_nchw,
                      void* workspace, size_t workspace_bytes,
                      uint64_t seed, int training, int backend, pub_stream_t s);
/* test hook: the Bernoulli keep-mask (1/0 bytes, NCHW order) the engine applies in block
 * `block_index` (position in enc+dec order) for (seed); replaces F.dropout's generator
 * (src/networks.py:177) */
int pub_unet_dropout_mask(const pub_unet* u, int block_index, int B, int H, int W, uint64_t seed,
                          uint8_t* mask_nchw, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Axis-aligned Gaussian encoder: AxisAlignedConvGaussian.forward (src/prob_unet.py:56-85).
 * params: {encoder conv k: weight,bias}*(3*n_stages), conv_mu.{w,b}, conv_log_sigma.{w,b}
 * ---------------------------------------------------------------------------------- */
typedef struct pub_encoder pub_encoder;
int pub_encoder_create(int in_channels, const int32_t* filters, int n_stages, int latent_dim,
                       int dtype, pub_encoder** out);
void pub_encoder_destroy(pub_encoder* e);
int pub_encoder_num_params(const pub_encoder* e);
size_t pub_encoder_workspace_bytes(const pub_encoder* e, int B, int H, int W);
int pub_encoder_forward(pub_encoder* e, int B, int H, int W, const float* x_nchw, int cx,
                        const float* t_nchw, int ct, const float* const* params,
                        float* mu, float* sigma, void* workspace, size_t workspace_bytes,
                        int backend, pub_stream_t s);
int pub_encoder_backward(pub_encoder* e, int B, int H, int W, const float* dmu, const float* dsigma,
                         const float* const* params, float* const* grads,
                         void* workspace, size_t workspace_bytes, int backend, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Latent ops: Independent(Normal).rsample / kl_divergence (src/prob_unet.py:215,221,247,255)
 * ---------------------------------------------------------------------------------- */
/* eps[M,B,L] ~ N(0,1) from Philox4x32-10 (seed, offset) unless eps_in given; z = mu + sigma*eps */
int pub_rsample_forward(const float* mu, const float* sigma, const float* eps_in, uint64_t seed,
                        uint64_t offset, int M, int B, int L, float* z, float* eps_out, pub_stream_t s);
int pub_rsample_backward(const float* dz, const float* eps, int M, int B, int L,
                         float* dmu, float* dsigma, pub_stream_t s);
int pub_kl_normal_forward(const float* mu_q, const float* sig_q, const float* mu_p, const float* sig_p,
                          int B, int L, float* kl, pub_stream_t s);
int pub_kl_normal_backward(const float* dkl, const float* mu_q, const float* sig_q, const float* mu_p,
                           const float* sig_p, int B, int L, float* dmu_q, float* dsig_q,
                           float* dmu_p, float* dsig_p, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Fcomb: Fcomb.forward (src/prob_unet.py:120-138) for M latent samples at once.
 * Weights are the f32 OIHW parameters of fcomb.layers.{0,2,4}.  F must be 32.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const void* feat;              /* features: NHWC dt (feat_nchw=0) or strided NCHW f32        */
  int32_t feat_nchw, dtype;
  int64_t stride[4];             /* element strides (b,c,h,w) when feat_nchw=1                 */
  const float* z;                /* [M,B,L]                                                    */
  const float *w0, *b0, *w1, *b1, *w2, *b2;
  float* out;                    /* [B,M,C,H,W] f32                                            */
  int32_t B, H, W, F, L, C, M;
} pub_fcomb_args;
size_t pub_fcomb_forward_workspace(const pub_fcomb_args* a);
int pub_fcomb_forward(const pub_fcomb_args* a, void* workspace, size_t workspace_bytes, pub_stream_t s);
size_t pub_fcomb_backward_workspace(const pub_fcomb_args* a);
/* dfeat: same layout/dtype as feat would have if contiguous (NHWC dt, or NCHW f32); may be NULL */
int pub_fcomb_backward(const pub_fcomb_args* a, const float* dout, void* dfeat, float* dz,
                       float* dw0, float* db0, float* dw1, float* db1, float* dw2, float* db2,
                       void* workspace, size_t workspace_bytes, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Reconstruction losses (src/prob_unet_utils.py:171-268, src/prob_unet.py:357-362)
 * ---------------------------------------------------------------------------------- */
#define PUB_LOSS_AFCRPS 0
#define PUB_LOSS_CRPS 1
size_t pub_loss_workspace(int B, int C, int HW);
/* ens [B,M,C,HW] f32, target [B,C,HW] f32 -> loss (device scalar).  dens (optional, may be
 * NULL) receives d loss / d ens (unscaled; multiply by the upstream scalar gradient). */
int pub_ensemble_loss(const float* ens, const float* target, int B, int M, int C, int HW, int kind,
                      float alpha, float* loss, float* dens, void* workspace, size_t workspace_bytes,
                      pub_stream_t s);
/* out/target [B,C,HW] -> loss[0] = mean |out-target|, loss[1+c] = per-variable means; dout optional */
int pub_l1_loss(const float* out, const float* target, int B, int C, int HW, float* loss, float* dout,
                void* workspace, size_t workspace_bytes, pub_stream_t s);
/* WMSE + (1 - MS-SSIM) of the reference's active elbo (src/prob_unet_utils.py:270-305 + pytorch_msssim.ms_ssim,
 * win_size 7, 5 levels): pred/target [B,C,H,W] f32 NCHW, H and W divisible by 16 and > 96.
 * out3 = { lam*wmse + (1-lam)*(1-msssim), wmse, 1-msssim } (device); dpred (optional) = d out3[0] / d pred.
 * data_range = clamp(max(target) - min(target), 1e-5) is reduced on the device (no host sync). */
size_t pub_msssim_workspace(int B, int C, int H, int W);
int pub_wmse_msssim_loss(const float* pred, const float* target, int B, int C, int H, int W, float alpha, float beta,
                         float lam, float* out3, float* dpred, void* workspace, size_t workspace_bytes, pub_stream_t s);
/* y[i] *= *scale (device scalar): applies the upstream gradient to a stored local gradient */
int pub_scale_by_device_scalar(float* y, const float* scale, int64_t n, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Dataset transform on the GPU: climex2torch.__getitem__ for type "lrinterp_to_residuals"
 * (src/climex_utils.py:197-225) and compute_stats (:255-264), batched.  All tensors f32 NCHW.
 *   pub_climex_stats:     hr [T,C,H,W] -> mean_lr, std_lr [C,H/s,W/s] (unbiased std over T of the s x s cell means)
 *   pub_climex_transform: hr [B,C,H,W] -> inputs, targets, lrinterp [B,C,H,W], lr [B,C,H/s,W/s]
 *                         (lrinterp / lr optional, may be NULL); eps = 1e-10 in the reference (:86)
 * ---------------------------------------------------------------------------------- */
int pub_climex_stats(const float* hr, int T, int C, int H, int W, int lowres_scale, float* mean_lr, float* std_lr,
                     pub_stream_t s);
int pub_climex_transform(const float* hr, const float* mean_lr, const float* std_lr, int B, int C, int H, int W,
                         int lowres_scale, float eps, float* inputs, float* targets, float* lrinterp, float* lr,
                         pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Ensemble metrics: metrics.crps_over_groundtruth / compute_mae (src/metrics.py:11-71) with
 * residual_to_hr + inverse transforms fused (src/climex_utils.py:277-285,42-46).
 * preds [T,M,3,HW] f32; transform=1: preds are standardised residuals, converted with
 * lrinterp [T,3,HW] and std_hr[3] to real units before scoring against hr [T,3,HW].
 * ---------------------------------------------------------------------------------- */
size_t pub_ensemble_metrics_workspace(int T, int C, int HW);
int pub_ensemble_metrics(const float* preds, const float* hr, const float* lrinterp, const float* std_hr,
                         int transform, int T, int M, int C, int HW, float* crps_tc, float* mae_tc,
                         void* workspace, size_t workspace_bytes, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Optimizer: torch.optim.AdamW step (src/train_prob_unet_model.py:139-141) over a device
 * table of tensors.  table: n rows of {param*, grad*, exp_avg*, exp_avg_sq*, numel}.
 * ---------------------------------------------------------------------------------- */
typedef struct { float* p; const float* g; float* m; float* v; int64_t n; } pub_adamw_entry;
int pub_adamw_step(const pub_adamw_entry* device_table, int n_tensors, int64_t max_numel,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                   float grad_scale, pub_stream_t s);

/* ------------------------------------------------------------------------------------
 * Ensemble post-processing diagnostics of results.ipynb, all on the device:
 *   pub_radial_psd  : psd() / compute_psd_tensor() of cell 4 -- torch.fft.fftn power spectrum, azimuthal mean over the
 *                     radial bins [k + 0.5, k + 1.5) (scipy.stats.binned_statistic "mean") times the shell area, for every
 *                     (sample, variable) field of data [N, C, H, H] f32 (H a power of two <= 128), and the mean over
 *                     the N samples.  transfo / units: the inverse variable transforms of that cell fused into the load
 *                     (pr = kgm2sTommday(softplus(x0)), tasmin = KToC(x1), tasmax = KToC(softplus(x2, c=0) + x1)).
 *                     table: bin -> pixel list, built once per H on the host (pub_psd_build_table) and copied to the device.
 *   pub_histogram   : np.histogram(values, bins=edges) of cell 15 (edges f64 ascending, last bin closed); counts must be
 *                     zeroed by the caller and are accumulated (several calls add up).
 * ---------------------------------------------------------------------------------- */
size_t pub_psd_table_ints(int H);
int pub_psd_build_table(int H, int* table_host);
int pub_radial_psd(const float* data, int N, int C, int H, int transfo, int units, const int* table_dev,
                   float* psd_fields, float* psd_mean, pub_stream_t s);
int pub_histogram(const float* values, int64_t n, const double* edges_dev, int nbins, unsigned long long* counts,
                  pub_stream_t s);

/* CUDA-graph variants (the whole training step captured once and replayed): a captured graph
 * replays the kernel ARGUMENTS of the capture, so whatever changes from step to step lives in device memory.
 * counters[0] = optimizer step count, counters[1] = random salt; pub_advance_counters (one tiny launch at the top of
 * the captured step) advances both; pub_debug_pointer("seed_salt", &counters[1]) makes every dropout key and rsample
 * seed of the library depend on the salt; pub_adamw_step_dev reads the step count for its bias corrections. */
int pub_advance_counters(int* counters, pub_stream_t s);
int pub_adamw_step_dev(const pub_adamw_entry* device_table, int n_tensors, int64_t max_numel,
                       float lr, float beta1, float beta2, float eps, float weight_decay,
                       const int* step_dev, float grad_scale, pub_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* PROBUNET_B200_H */
