// Internal launcher interface shared by the engine (.cu) files.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace pub {

struct ConvParams {
  const void* x0; const void* x1; int c0, c1, ld0, ld1;
  const void* w; const float* bias;
  const void* res; int ld_res;
  const void* mask; int ld_mask;
  void* y; int ldy;
  int B, H, W, cout, ks, relu;
  int round_tf32;  // SIMT path only: round the stored output to tf32 (PUB_TF32 tensors feed kind::tf32 MMAs)
  int w_settled;   // set by the engines: w was written by this library's pack kernels (note_weight_pack tracks them),
                   // so the halo kernel may request it before its grid-dependency wait; 0 for caller-supplied weights
  // ---- GroupNorm work fused into the epilogue (tcgen05 halo kernel, bf16 only: conv_fused_rows() > 0)
  // stat_part != nullptr, gn_bwd == 0: rows [B * conv_fused_rows()][cout][2] receive per-channel (sum, sum of squares)
  //   of the stored output = the forward statistics partials of the GroupNorm that reads y (src/networks.py:105-107)
  // gn_bwd == 1 (data-gradient launches): the result g is stored as du = g * keep/(1-p) * silu'(a x + b) -- x (gx0|gx1)
  //   the GroupNorm's input, gcoef its [B][cout][2] affine table -- and the rows receive (sum du, sum du * x)
  float* stat_part;
  int gn_bwd;
  const void* gx0; const void* gx1; int gc0, gld0, gld1;
  const float* gcoef;
  float p_drop; uint64_t seed, subseq;
  const uint32_t* salt;   // optional device word xor-ed into the dropout key (g_seed_salt; set by the engines)
};

struct WgradParams {
  const void* x0; const void* x1; int c0, c1, ld0, ld1;
  const void* dy; int ld_dy;
  float* dw; float* dbias;
  int B, H, W, cout, ks;
};

// ---- conv_simt.cu
int conv_simt(const ConvParams& p, int dtype, cudaStream_t s);
size_t wgrad_simt_workspace(const WgradParams& p);
int wgrad_simt(const WgradParams& p, int dtype, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s);
int wgrad_reduce(const float* part, float* dw, int nsplit, int taps, int cout, int cin, int accumulate, cudaStream_t s);
int colsum(const void* x, int ld, int C, int64_t M, int dtype, float* part, float* out, int accumulate, cudaStream_t s,
           int* nchunk_out = nullptr);
int wgrad_finish(const float* part, float* dw, int nsplit, int taps, int cout, int cin, const float* bpart, int nchunk,
                 float* dbias, int accumulate, cudaStream_t s);
int pack_weight(const float* w, void* out, int cout, int cin, int ks, int dtype, int tflip, cudaStream_t s);
// all weight tensors of a sub-network in ONE launch (a training step packs ~190 of them; one launch each cost
// more in launch gaps than in work)
struct PackEntry { const float* w; void* out; int cout, cin, ks, tflip; };
int pack_weights_batched(const PackEntry* e, int n, int dtype, cudaStream_t s);
int nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* y, int ldy, int B, int H, int W, int dtype,
                 cudaStream_t s);
int nhwc_to_nchw(const void* x, int ld, int C, float* y, int B, int H, int W, int dtype, int accumulate, cudaStream_t s);

// ---- conv_tc.cu (tcgen05 / TMEM / TMA)
bool conv_tc_supported(const ConvParams& p, int dtype);
// partial rows per image a launch with the fused GroupNorm epilogue writes (4 per 8 x 16 output tile), or 0 when this
// launch would not go through the halo kernel in bf16 (the caller then runs the separate statistics pass)
int conv_fused_rows(const ConvParams& p, int dtype, int backend);
int conv_tc(const ConvParams& p, int dtype, cudaStream_t s);
bool wgrad_tc_supported(const WgradParams& p, int dtype);
size_t wgrad_tc_workspace(const WgradParams& p, int dtype);
int wgrad_tc(const WgradParams& p, int dtype, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s);

// dispatchers (api.cu)
int conv_forward(const ConvParams& p, int dtype, int backend, cudaStream_t s);
size_t wgrad_workspace(const WgradParams& p, int dtype, int backend);
int wgrad(const WgradParams& p, int dtype, int backend, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s);

// ---- norm.cu : GroupNorm (+FiLM) + SiLU (+dropout) (+2x resample), forward and backward
struct GnParams {
  const void* x0; const void* x1; int c0, c1, ld0, ld1;  // GN input (virtual concat), at the INPUT resolution
  int B, H, W, groups;                                    // input resolution
  const float* gamma; const float* beta;                  // [C]
  const float* film;                                      // [2C] (scale | shift) or nullptr
  int resample;                                           // 0 none, 1 down (2x2 mean), 2 up (nearest 2x)
  float p_drop; uint64_t seed; uint64_t subseq;           // dropout on the activated output (p_drop = 0: off)
  const uint32_t* salt;                                   // optional device word xor-ed into the dropout key (g_seed_salt)
  float* stats;                                           // [B][G][2] mean, rstd   (saved for backward)
  float* coef;                                            // [B][C][2] scratch: per-(b,c) affine a,b
  float* partial;                                         // scratch for the two-stage reductions
  int rows;                                               // pixels per CTA chunk: filled in by gn_forward / gn_backward
  // statistics partials already emitted by the conv that PRODUCED x0 (and x1) -- ConvParams::stat_part --: pre_rows
  // rows per image of [c0][2] / [c1][2] (sum, sum of squares).  gn_forward then skips its own pass over x.
  const float* pre0; const float* pre1; int pre_rows;
};
size_t gn_partial_floats(int B, int C, int H, int W);
// y [B,H',W',C] NHWC dt (contiguous, ld = C)
int gn_forward(const GnParams& p, void* y, int dtype, cudaStream_t s);
// dy: gradient wrt y (at the OUTPUT resolution, contiguous ld = C).  dx [B,H,W,C] contiguous; addend
// (optional, NHWC dt, ld_add) is added to dx.  dgamma/dbeta [C], dfilm [2C] overwritten.
int gn_backward(const GnParams& p, const void* dy, void* dx, const void* addend, int ld_add, float* dgamma,
                float* dbeta, float* dfilm, int dtype, cudaStream_t s);
// same, when the data-gradient conv already produced du = dy * keep/(1-p) * silu'(a x + b) and its partial sums
// (ConvParams::gn_bwd): du [B,H,W,C] contiguous, du_part = rows_per_image rows per image of [C][2] (sum du, sum du * x).
// No resampling (the fused epilogue only exists for blocks whose conv runs at the GroupNorm's resolution).
int gn_backward_from_du(const GnParams& p, const void* du, const float* du_part, int rows_per_image, void* dx,
                        const void* addend, int ld_add, float* dgamma, float* dbeta, float* dfilm, int dtype,
                        cudaStream_t s);

// ---- elementwise.cu
int resample2x(const void* x, int ld, int C, void* y, int B, int H, int W, int mode, int dtype, cudaStream_t s);
// backward of resample2x: dx (input res [B,H,W,C]) from dy (output res); mode as forward
int resample2x_bwd(const void* dy, int ld, int C, void* dx, int B, int H, int W, int mode, int dtype, cudaStream_t s);
int add_views(const void* a, int lda, const void* b, int ldb, void* y, int ldy, int C, int64_t M, int dtype,
              cudaStream_t s);
// out_dtype >= 0 and of another width than dtype: bf16 in, f32 out (the encoder's bf16 -> tf32 stage boundary)
int maxpool2(const void* x, int C, void* y, int B, int H, int W, int dtype, cudaStream_t s, int out_dtype = -1);
// dx[b,y,x,c] = dy[b,y/2,x/2,c] if x is the (first) argmax of its window and mask_src>0 ... see elementwise.cu
int maxpool2_bwd(const void* x, const void* yp, const void* dy, void* dx, int C, int B, int H, int W, int dtype,
                 cudaStream_t s, int dy_dtype = -1);   // dy_dtype: as out_dtype above (f32 gradient in, bf16 x / dx)
int relu_mask_inplace(void* dy, const void* y, int64_t n, int dtype, cudaStream_t s);
int global_mean(const void* x, int C, int B, int64_t HW, float* out, float* partial, int dtype, cudaStream_t s);
int global_mean_bwd(const float* dmean, const void* y_mask, int C, int B, int64_t HW, void* dx, int dtype,
                    cudaStream_t s);
int fill_zero(void* p, size_t bytes, cudaStream_t s);

}  // namespace pub
