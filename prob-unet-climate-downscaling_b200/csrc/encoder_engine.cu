// Axis-aligned Gaussian encoder executor: AxisAlignedConvGaussian.forward (src/prob_unet.py:56-85)
// -- [MaxPool2d(2)] + 3 x (conv3x3 + ReLU) per stage, global spatial mean, 1x1 mu / log-sigma heads,
// sigma = exp(log sigma) + 1e-7 -- and its reverse-mode gradient.  ReLU is fused in the conv
// epilogue; ReLU' is fused in the data-gradient epilogue (mask on the saved activation); MaxPool
// backward recomputes the argmax from the saved pre-pool activation.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

struct pub_encoder {
  int in_ch = 0, latent = 0, dtype = PUB_BF16;
  int s0_bf16 = 0;   // PUB_TF32_BF16S0: the full-resolution stage (3 convs) in bf16, everything behind its pooling in tf32
  int layer_dtype(int k) const { return (s0_bf16 && k < 3) ? PUB_BF16 : dtype; }
  std::vector<int> filters;
  int nconv = 0, nparams = 0;
};

namespace pub {
namespace {

struct EPlan {
  int B, H, W;
  size_t es;
  void* x_in;
  std::vector<void*> act;     // conv outputs (post-ReLU)
  std::vector<void*> pooled;  // per stage (index 0 unused)
  std::vector<int> ch, hh, ww;  // per conv output
  float *gap, *ls, *dgap, *dls;
  std::vector<void*> wp;      // packed weights per conv (forward OR mirrored layout)
  void *ga, *gb, *wg_ws;
  size_t wg_ws_bytes, total;
};

int build(const pub_encoder* e, int B, int H, int W, void* base, size_t cap, EPlan& pl) {
  Arena ar(base, cap);
  const size_t es = dtype_size(e->dtype);          // widest storage type of the network (scratch / pooled buffers)
  pl.B = B; pl.H = H; pl.W = W; pl.es = es;
  pl.x_in = ar.take((size_t)B * H * W * 8 * dtype_size(e->layer_dtype(0)));
  pl.act.clear(); pl.pooled.clear(); pl.ch.clear(); pl.hh.clear(); pl.ww.clear(); pl.wp.clear();
  int h = H, w = W, cin = e->in_ch;
  size_t max_act = (size_t)B * H * W * 8 * es, max_w = 0, max_wg = 0;
  for (size_t st = 0; st < e->filters.size(); ++st) {
    const int f = e->filters[st];
    if (st > 0) {
      PUB_REQUIRE(h % 2 == 0 && w % 2 == 0, "encoder: odd resolution %dx%d cannot be pooled", h, w);
      h /= 2; w /= 2;
      pl.pooled.push_back(ar.take((size_t)B * h * w * cin * es));
    } else {
      pl.pooled.push_back(nullptr);
    }
    for (int k = 0; k < 3; ++k) {
      PUB_REQUIRE(f % 8 == 0, "encoder: filter counts must be multiples of 8");
      const int ldt = e->layer_dtype((int)st * 3 + k);
      pl.act.push_back(ar.take((size_t)B * h * w * f * dtype_size(ldt)));
      pl.wp.push_back(ar.take((size_t)9 * f * cin * dtype_size(ldt)));
      pl.ch.push_back(f); pl.hh.push_back(h); pl.ww.push_back(w);
      max_act = std::max(max_act, (size_t)B * h * w * std::max(f, cin) * es);
      max_w = std::max(max_w, (size_t)9 * f * std::max(cin, 8) * es);
      WgradParams wp{};
      wp.c0 = cin; wp.cout = f; wp.B = B; wp.H = h; wp.W = w; wp.ks = 3; wp.ld0 = cin; wp.ld_dy = f;
      max_wg = std::max(max_wg, wgrad_workspace(wp, ldt, PUB_BACKEND_AUTO));
      max_wg = std::max(max_wg, wgrad_simt_workspace(wp));
      cin = f;
    }
  }
  // the pre-pool tensors are larger than their pooled versions: scratch must hold the largest activation
  const int F = e->filters.back();
  pl.gap = ar.take_n<float>((size_t)B * F);
  pl.ls = ar.take_n<float>((size_t)B * e->latent);
  pl.dgap = ar.take_n<float>((size_t)B * F);
  pl.dls = ar.take_n<float>((size_t)B * e->latent);
  (void)max_w;
  pl.ga = ar.take(max_act);
  pl.gb = ar.take(max_act);
  pl.wg_ws_bytes = max_wg;
  pl.wg_ws = ar.take(max_wg);
  pl.total = ar.off + 1024;
  PUB_REQUIRE(ar.ok(), "encoder workspace too small: need %zu bytes, have %zu", pl.total, cap);
  return 0;
}

// one warp per (sample, latent): the two 1x1 heads are dot products over the F pooled features (a thread per output
// walking all F serially took 41 us per launch for 2 048 outputs)
__global__ void heads_fwd_kernel(const float* __restrict__ gap, const float* __restrict__ wmu, const float* __restrict__ bmu,
                                 const float* __restrict__ wls, const float* __restrict__ bls, int B, int F, int L,
                                 float* __restrict__ mu, float* __restrict__ sigma, float* __restrict__ ls_out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= B * L) return;
  const int b = i / L, l = i % L;
  float m = 0.f, s = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float g = gap[b * F + f];
    m = fmaf(wmu[l * F + f], g, m);
    s = fmaf(wls[l * F + f], g, s);
  }
  m = warp_sum(m) + bmu[l]; s = warp_sum(s) + bls[l];
  if (lane == 0) { mu[i] = m; ls_out[i] = s; sigma[i] = expf(s) + 1e-7f; }
}

__global__ void heads_bwd_kernel(const float* __restrict__ dmu, const float* __restrict__ dsigma, const float* __restrict__ ls,
                                 const float* __restrict__ gap, const float* __restrict__ wmu, const float* __restrict__ wls,
                                 int B, int F, int L, float* __restrict__ dls, float* __restrict__ dwmu,
                                 float* __restrict__ dbmu, float* __restrict__ dwls, float* __restrict__ dbls,
                                 float* __restrict__ dgap) {
  // phase A: parameter grads (thread per (l, f)); dls recomputed inline
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < L * F; i += nth) {
    const int l = i / F, f = i % F;
    float a = 0.f, c = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = gap[b * F + f];
      a = fmaf(dmu[b * L + l], g, a);
      c = fmaf(dsigma[b * L + l] * expf(ls[b * L + l]), g, c);
    }
    dwmu[i] = a; dwls[i] = c;
  }
  for (int l = tid; l < L; l += nth) {
    float a = 0.f, c = 0.f;
    for (int b = 0; b < B; ++b) { a += dmu[b * L + l]; c += dsigma[b * L + l] * expf(ls[b * L + l]); }
    dbmu[l] = a; dbls[l] = c;
  }
  for (int i = tid; i < B * F; i += nth) {
    const int b = i / F, f = i % F;
    float a = 0.f;
    for (int l = 0; l < L; ++l)
      a += dmu[b * L + l] * wmu[l * F + f] + dsigma[b * L + l] * expf(ls[b * L + l]) * wls[l * F + f];
    dgap[i] = a;
  }
  (void)dls;
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

int pub_encoder_create(int in_channels, const int32_t* filters, int n_stages, int latent_dim, int dtype,
                       pub_encoder** out) {
  PUB_REQUIRE(filters && out && n_stages > 0 && latent_dim > 0, "pub_encoder_create: bad arguments");
  PUB_REQUIRE(in_channels >= 1 && in_channels <= 8, "pub_encoder_create: in_channels must be in [1, 8]");
  pub_encoder* e = new pub_encoder();
  PUB_REQUIRE(dtype == PUB_F32 || dtype == PUB_BF16 || dtype == PUB_TF32 || dtype == PUB_TF32_BF16S0,
              "pub_encoder_create: bad dtype %d", dtype);
  e->in_ch = in_channels; e->latent = latent_dim;
  e->s0_bf16 = dtype == PUB_TF32_BF16S0 && n_stages > 1;
  e->dtype = dtype == PUB_TF32_BF16S0 ? PUB_TF32 : dtype;
  e->filters.assign(filters, filters + n_stages);
  e->nconv = 3 * n_stages;
  e->nparams = 2 * e->nconv + 4;
  *out = e;
  return 0;
}
void pub_encoder_destroy(pub_encoder* e) { delete e; }
int pub_encoder_num_params(const pub_encoder* e) { return e ? e->nparams : 0; }
size_t pub_encoder_workspace_bytes(const pub_encoder* e, int B, int H, int W) {
  EPlan pl;
  if (!e || build(e, B, H, W, nullptr, 0, pl) != 0) return 0;
  return pl.total;
}

int pub_encoder_forward(pub_encoder* e, int B, int H, int W, const float* x_nchw, int cx, const float* t_nchw, int ct,
                        const float* const* P, float* mu, float* sigma, void* ws, size_t ws_bytes, int backend,
                        pub_stream_t s_) {
  PUB_REQUIRE(e && x_nchw && P && mu && sigma && ws, "pub_encoder_forward: null argument");
  PUB_REQUIRE(cx + (t_nchw ? ct : 0) == e->in_ch, "pub_encoder_forward: got %d input channels, encoder expects %d",
              cx + (t_nchw ? ct : 0), e->in_ch);
  cudaStream_t s = (cudaStream_t)s_;
  EPlan pl;
  PUB_TRY(build(e, B, H, W, ws, ws_bytes, pl));
  const int dt = e->dtype;
  PUB_TRY(nchw_to_nhwc(x_nchw, cx, t_nchw, t_nchw ? ct : 0, pl.x_in, 8, B, H, W, e->layer_dtype(0), s));
  {
    std::vector<PackEntry> pe[2];   // [0]: layers stored in e->dtype, [1]: the bf16 first stage of the mixed form
    int ci = e->in_ch;
    for (int k = 0; k < e->nconv; ++k) {
      pe[e->layer_dtype(k) != dt].push_back({P[2 * k], pl.wp[k], pl.ch[k], ci, 3, 0});
      ci = pl.ch[k];
    }
    PUB_TRY(pack_weights_batched(pe[0].data(), (int)pe[0].size(), dt, s));
    if (!pe[1].empty()) PUB_TRY(pack_weights_batched(pe[1].data(), (int)pe[1].size(), PUB_BF16, s));
  }
  const void* cur = pl.x_in;
  int cin = e->in_ch, ld = 8, h = H, w = W;
  for (int k = 0; k < e->nconv; ++k) {
    const int st = k / 3, f = pl.ch[k], ldt = e->layer_dtype(k);
    if (k % 3 == 0 && st > 0) {
      PUB_TRY(maxpool2(cur, cin, pl.pooled[st], B, h, w, e->layer_dtype(k - 1), s, ldt));
      cur = pl.pooled[st]; h /= 2; w /= 2;
    }
    ConvParams c{};
    c.x0 = cur; c.c0 = cin; c.ld0 = ld; c.w = pl.wp[k]; c.bias = P[2 * k + 1];
    c.y = pl.act[k]; c.ldy = f; c.B = B; c.H = h; c.W = w; c.cout = f; c.ks = 3; c.relu = 1;
    c.w_settled = 1;
    PUB_TRY(conv_forward(c, ldt, backend, s));
    cur = pl.act[k]; cin = f; ld = f;
  }
  const int F = e->filters.back(), L = e->latent;
  PUB_TRY(global_mean(cur, F, B, (int64_t)h * w, pl.gap, nullptr, dt, s));
  const float* const* hp = P + 2 * e->nconv;
  heads_fwd_kernel<<<cdiv((int64_t)B * L * 32, 128), 128, 0, s>>>(pl.gap, hp[0], hp[1], hp[2], hp[3], B, F, L, mu, sigma, pl.ls);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_encoder_backward(pub_encoder* e, int B, int H, int W, const float* dmu, const float* dsigma,
                         const float* const* P, float* const* G, void* ws, size_t ws_bytes, int backend,
                         pub_stream_t s_) {
  PUB_REQUIRE(e && dmu && dsigma && P && G && ws, "pub_encoder_backward: null argument");
  cudaStream_t s = (cudaStream_t)s_;
  EPlan pl;
  PUB_TRY(build(e, B, H, W, ws, ws_bytes, pl));
  const int dt = e->dtype, F = e->filters.back(), L = e->latent, n = e->nconv;
  const float* const* hp = P + 2 * n;
  float* const* hg = G + 2 * n;
  heads_bwd_kernel<<<cdiv((int64_t)std::max(L, B) * F, 256), 256, 0, s>>>(dmu, dsigma, pl.ls, pl.gap, hp[0], hp[2], B, F, L, pl.dls, hg[0], hg[1], hg[2],
                                      hg[3], pl.dgap);
  PUB_LAUNCH_CHECK();
  {
    std::vector<PackEntry> pe[2];
    for (int k = 1; k < n; ++k) pe[e->layer_dtype(k) != dt].push_back({P[2 * k], pl.wp[k], pl.ch[k], pl.ch[k - 1], 3, 1});
    PUB_TRY(pack_weights_batched(pe[0].data(), (int)pe[0].size(), dt, s));
    if (!pe[1].empty()) PUB_TRY(pack_weights_batched(pe[1].data(), (int)pe[1].size(), PUB_BF16, s));
  }
  void* ga = pl.ga;
  void* gb = pl.gb;
  PUB_TRY(global_mean_bwd(pl.dgap, pl.act[n - 1], F, B, (int64_t)pl.hh[n - 1] * pl.ww[n - 1], ga, dt, s));
  for (int k = n - 1; k >= 0; --k) {
    const int st = k / 3, f = pl.ch[k], h = pl.hh[k], w = pl.ww[k];
    const int ldt = e->layer_dtype(k);            // type of this layer's tensors, gradients and packed weights
    const bool pooled_in = (k % 3 == 0 && st > 0);
    const void* in_k; int cin, ld;
    if (k == 0) { in_k = pl.x_in; cin = e->in_ch; ld = 8; }
    else if (pooled_in) { in_k = pl.pooled[st]; cin = pl.ch[k - 1]; ld = cin; }
    else { in_k = pl.act[k - 1]; cin = pl.ch[k - 1]; ld = cin; }
    WgradParams wp{};
    wp.x0 = in_k; wp.c0 = cin; wp.ld0 = ld; wp.dy = ga; wp.ld_dy = f; wp.dw = G[2 * k]; wp.dbias = G[2 * k + 1];
    wp.B = B; wp.H = h; wp.W = w; wp.cout = f; wp.ks = 3;
    PUB_TRY(wgrad(wp, ldt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
    if (k == 0) break;
    ConvParams c{};
    c.x0 = ga; c.c0 = f; c.ld0 = f; c.w = pl.wp[k]; c.y = gb; c.ldy = cin; c.B = B; c.H = h; c.W = w; c.cout = cin; c.ks = 3;
    c.w_settled = 1;
    if (!pooled_in) { c.mask = pl.act[k - 1]; c.ld_mask = cin; }
    PUB_TRY(conv_forward(c, ldt, backend, s));
    if (pooled_in) {
      // gb = d pooled (this layer's type)  ->  ga = d pre-activation of conv k-1 at the finer resolution (its type)
      PUB_TRY(maxpool2_bwd(pl.act[k - 1], nullptr, gb, ga, cin, B, pl.hh[k - 1], pl.ww[k - 1], e->layer_dtype(k - 1), s, ldt));
    } else {
      std::swap(ga, gb);
    }
  }
  return 0;
}

}  // extern "C"
