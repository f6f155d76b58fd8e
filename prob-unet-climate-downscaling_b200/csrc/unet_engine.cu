// U-Net executor: networks.UNet.forward (src/networks.py:299-333) and its reverse-mode gradient,
// as a fixed sequence of kernel launches over a caller-provided workspace.  No allocation, no
// synchronisation, no host round trips: one call enqueues the whole network on the stream
// (so a training step can be captured in a CUDA graph).
//
// Data layout in HBM: activations NHWC (bf16 or f32), every saved tensor lives at a deterministic
// offset of the workspace (the same bump-allocation order is replayed by forward and backward).
// Skip connections are never copied: the decoder's conv/GroupNorm kernels read the encoder output
// in place as the second source of a virtual channel concat (src/networks.py:329).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace pub {

struct BlockPlan {
  pub_unet_block d;
  int Hi, Wi, Ho, Wo;       // input / output resolution
  int c0, c1;               // channels of the two input sources (c1 = skip channels of a concat)
  int src0, src1;           // producer entry index of each source (-1: network input)
  int pidx;                 // first index into the param table
  // workspace buffers
  void *a0, *h, *a1, *out, *dx;
  void *wp0, *wp1, *wps;    // packed weights of conv0 / conv1 / skip (plain conv: wp0), forward OR mirrored layout
  float *stats0, *coef0, *stats1, *coef1;
  float *h_part, *out_part;  // GroupNorm statistics partials emitted by the epilogues of conv0 / conv1 (ConvParams::stat_part)
};

}  // namespace pub

struct pub_unet {
  std::vector<pub_unet_block> blocks;  // enc then dec
  int n_enc = 0, in_ch = 0, out_ch = 0, dtype = PUB_BF16;
  float dropout = 0.f;
  int nparams = 0;
  int final_c = 0;
};

namespace pub {
namespace {

inline int groups_of(int c) { int g = c / 4; return g < 32 ? g : 32; }  // networks.GroupNorm: min(32, C // 4)

struct Plan {
  std::vector<BlockPlan> bp;
  int B, H, W;
  size_t es;
  // network-level buffers
  void* x_in;      // NHWC input
  void* a_out;     // silu(out_norm(x))
  void* y_out;     // out_conv output NHWC (when the caller wants NCHW)
  float *stats_o, *coef_o;
  void* dy_in;     // NHWC copy of dout
  void* dx_last;   // grad wrt last decoder block output
  // scratch
  void *wp_out;             // packed out_conv weights
  void *gn_scratch, *skipbuf, *s1, *s2, *sadd, *gsum, *wg_ws;
  float *du_part;           // backward: (sum du, sum du*x) partials emitted by a data-gradient conv (ConvParams::gn_bwd)
  size_t wg_ws_bytes;
  size_t total;
};

int build_plan(const pub_unet* u, int B, int H, int W, void* base, size_t cap, Plan& pl) {
  Arena ar(base, cap);
  const size_t es = dtype_size(u->dtype);
  pl.B = B; pl.H = H; pl.W = W; pl.es = es;
  pl.bp.clear();
  pl.x_in = ar.take((size_t)B * H * W * 8 * es);  // padded to 8 channels
  struct Out { int c, h, w, idx; };
  std::vector<Out> skips;
  int curC = u->in_ch, curH = H, curW = W, cur = -1;
  size_t max_act = 0, max_w = 0, max_gn = 0, max_wg = 0, max_part = 0;
  // rows of statistics partials a fused conv epilogue writes per image at (h, w): 4 per 8 x 16 output tile
  auto part_floats = [&](int h, int w, int c) -> size_t {
    return (h % 16 == 0 && w % 8 == 0) ? (size_t)B * 4 * (w / 8) * (h / 16) * c * 2 : 0;
  };
  int pidx = 0;
  for (size_t i = 0; i < u->blocks.size(); ++i) {
    BlockPlan b{};
    b.d = u->blocks[i];
    const bool is_dec = (int)i >= u->n_enc;
    b.src0 = cur; b.src1 = -1; b.c0 = curC; b.c1 = 0;
    b.Hi = curH; b.Wi = curW;
    if (is_dec && curC != b.d.cin) {
      PUB_REQUIRE(!skips.empty(), "unet plan: skip stack empty at block %zu", i);
      Out s = skips.back(); skips.pop_back();
      PUB_REQUIRE(s.h == curH && s.w == curW && curC + s.c == b.d.cin, "unet plan: concat mismatch at block %zu (%d+%d != %d)", i, curC, s.c, b.d.cin);
      b.c1 = s.c; b.src1 = s.idx;
    }
    PUB_REQUIRE(b.c0 + b.c1 == b.d.cin, "unet plan: channel mismatch at block %zu (%d vs %d)", i, b.c0 + b.c1, b.d.cin);
    b.Ho = b.d.up ? curH * 2 : (b.d.down ? curH / 2 : curH);
    b.Wo = b.d.up ? curW * 2 : (b.d.down ? curW / 2 : curW);
    PUB_REQUIRE(!b.d.down || (curH % 2 == 0 && curW % 2 == 0), "unet plan: odd resolution %dx%d cannot be halved", curH, curW);
    b.pidx = pidx;
    const size_t n_in = (size_t)B * b.Hi * b.Wi, n_out = (size_t)B * b.Ho * b.Wo;
    if (b.d.is_conv) {
      pidx += 2;
      b.out = ar.take(n_out * b.d.cout * es);
      b.dx = nullptr;
      b.wp0 = ar.take((size_t)9 * b.d.cin * b.d.cout * es);
      b.out_part = ar.take_n<float>(part_floats(b.Ho, b.Wo, b.d.cout));
    } else {
      PUB_REQUIRE(b.d.cin % 8 == 0 && b.d.cout % 8 == 0, "unet plan: block channels must be multiples of 8");
      pidx += 9 + (b.d.has_skip_conv ? 2 : 0);
      b.a0 = ar.take(n_out * b.d.cin * es);
      b.h = ar.take(n_out * b.d.cout * es);
      b.a1 = ar.take(n_out * b.d.cout * es);
      b.out = ar.take(n_out * b.d.cout * es);
      b.dx = ar.take(n_in * b.d.cin * es);
      b.wp0 = ar.take((size_t)9 * b.d.cin * b.d.cout * es);
      b.wp1 = ar.take((size_t)9 * b.d.cout * b.d.cout * es);
      b.wps = b.d.has_skip_conv ? ar.take((size_t)b.d.cin * b.d.cout * es) : nullptr;
      b.stats0 = ar.take_n<float>((size_t)B * groups_of(b.d.cin) * 2);
      b.coef0 = ar.take_n<float>((size_t)B * b.d.cin * 2);
      b.stats1 = ar.take_n<float>((size_t)B * groups_of(b.d.cout) * 2);
      b.coef1 = ar.take_n<float>((size_t)B * b.d.cout * 2);
      b.h_part = ar.take_n<float>(part_floats(b.Ho, b.Wo, b.d.cout));
      b.out_part = ar.take_n<float>(part_floats(b.Ho, b.Wo, b.d.cout));
      max_part = std::max(max_part, part_floats(b.Ho, b.Wo, std::max(b.d.cin, b.d.cout)));
      max_gn = std::max(max_gn, gn_partial_floats(B, b.d.cin, b.Hi, b.Wi));
      max_gn = std::max(max_gn, gn_partial_floats(B, b.d.cout, b.Ho, b.Wo));
    }
    max_act = std::max(max_act, std::max(n_in, n_out) * (size_t)std::max(b.d.cin, b.d.cout) * es);
    max_w = std::max(max_w, (size_t)9 * b.d.cin * b.d.cout * es);
    {
      WgradParams wp{};
      wp.c0 = b.d.cin; wp.c1 = 0; wp.cout = b.d.cout; wp.B = B; wp.H = b.Ho; wp.W = b.Wo; wp.ks = 3;
      wp.ld0 = b.d.cin; wp.ld_dy = b.d.cout;
      max_wg = std::max(max_wg, wgrad_workspace(wp, u->dtype, PUB_BACKEND_AUTO));
      max_wg = std::max(max_wg, wgrad_simt_workspace(wp));
    }
    pl.bp.push_back(b);
    cur = (int)i; curC = b.d.cout; curH = b.Ho; curW = b.Wo;
    if (!is_dec) skips.push_back({curC, curH, curW, (int)i});
  }
  PUB_REQUIRE(curH == H && curW == W, "unet plan: output resolution %dx%d != input %dx%d", curH, curW, H, W);
  PUB_REQUIRE(curC % 8 == 0, "unet plan: final channel count must be a multiple of 8");
  const size_t n = (size_t)B * H * W;
  pl.a_out = ar.take(n * curC * es);
  pl.y_out = ar.take(n * u->out_ch * es);
  pl.stats_o = ar.take_n<float>((size_t)B * groups_of(curC) * 2);
  pl.coef_o = ar.take_n<float>((size_t)B * curC * 2);
  pl.dy_in = ar.take(n * u->out_ch * es);
  pl.dx_last = ar.take(n * curC * es);
  max_gn = std::max(max_gn, gn_partial_floats(B, curC, H, W));
  max_w = std::max(max_w, (size_t)9 * curC * u->out_ch * es);
  max_act = std::max(max_act, n * (size_t)std::max(curC, u->out_ch) * es);
  {
    WgradParams wp{};
    wp.c0 = curC; wp.cout = u->out_ch; wp.B = B; wp.H = H; wp.W = W; wp.ks = 3; wp.ld0 = curC; wp.ld_dy = u->out_ch;
    max_wg = std::max(max_wg, wgrad_workspace(wp, u->dtype, PUB_BACKEND_AUTO));
    max_wg = std::max(max_wg, wgrad_simt_workspace(wp));
  }
  pl.wp_out = ar.take((size_t)9 * curC * u->out_ch * es);
  (void)max_w;
  pl.gn_scratch = ar.take(max_gn * sizeof(float));
  pl.skipbuf = ar.take(max_act);
  pl.s1 = ar.take(max_act);
  pl.s2 = ar.take(max_act);
  pl.sadd = ar.take(max_act);
  pl.gsum = ar.take(max_act);
  pl.wg_ws_bytes = max_wg;
  pl.wg_ws = ar.take(max_wg);
  pl.du_part = ar.take_n<float>(std::max(max_part, part_floats(H, W, curC)));
  pl.total = ar.off + 1024;
  PUB_REQUIRE(ar.ok(), "unet workspace too small: need %zu bytes, have %zu", pl.total, cap);
  return 0;
}

// view of a producer's output
struct View { const void* p; int c, ld; };
View src_view(const Plan& pl, int idx, int in_ch_padded_ld) {
  if (idx < 0) return View{pl.x_in, 0, in_ch_padded_ld};
  const BlockPlan& b = pl.bp[idx];
  return View{b.out, b.d.cout, b.d.cout};
}

GnParams gn_params(const Plan& pl, const void* x0, int c0, int ld0, const void* x1, int c1, int ld1, int H, int W,
                   const float* gamma, const float* beta, const float* film, int resample, float p_drop,
                   uint64_t seed, uint64_t subseq, float* stats, float* coef) {
  GnParams g{};
  g.x0 = x0; g.x1 = x1; g.c0 = c0; g.c1 = c1; g.ld0 = ld0; g.ld1 = ld1;
  g.B = pl.B; g.H = H; g.W = W; g.groups = groups_of(c0 + c1);
  g.gamma = gamma; g.beta = beta; g.film = film; g.resample = resample;
  g.p_drop = p_drop; g.seed = seed; g.subseq = subseq; g.salt = g_seed_salt;
  g.stats = stats; g.coef = coef; g.partial = (float*)pl.gn_scratch;
  return g;
}

ConvParams conv_params(const void* x0, int c0, int ld0, const void* x1, int c1, int ld1, const void* w,
                       const float* bias, const void* res, int ld_res, void* y, int ldy, int B, int H, int W, int cout,
                       int ks) {
  ConvParams c{};
  c.x0 = x0; c.x1 = x1; c.c0 = c0; c.c1 = c1; c.ld0 = ld0; c.ld1 = ld1; c.w = w; c.bias = bias;
  c.res = res; c.ld_res = ld_res; c.mask = nullptr; c.ld_mask = 0; c.y = y; c.ldy = ldy;
  c.B = B; c.H = H; c.W = W; c.cout = cout; c.ks = ks; c.relu = 0;
  c.w_settled = 1;   // packed by pack_all() at the start of the pass
  return c;
}

// every conv weight of the network in one launch: forward layout [tap][co][ci] or the mirrored / transposed layout
// of the data gradient
int pack_all(const pub_unet* u, const Plan& pl, const float* const* P, int tflip, int dt, cudaStream_t s) {
  std::vector<PackEntry> e;
  for (const BlockPlan& b : pl.bp) {
    const float* const* p = P + b.pidx;
    if (b.d.is_conv) { e.push_back({p[0], b.wp0, b.d.cout, b.d.cin, 3, tflip}); continue; }
    e.push_back({p[2], b.wp0, b.d.cout, b.d.cin, 3, tflip});
    e.push_back({p[7], b.wp1, b.d.cout, b.d.cout, 3, tflip});
    if (b.d.has_skip_conv) e.push_back({p[9], b.wps, b.d.cout, b.d.cin, 1, tflip});
  }
  const float* const* p = P + (u->nparams - 4);
  e.push_back({p[2], pl.wp_out, u->out_ch, u->final_c, 3, tflip});
  return pack_weights_batched(e.data(), (int)e.size(), dt, s);
}

// keep-mask of the engine's dropout stream, written in NCHW order (test hook)
__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int B, int HW, int C, float p, uint64_t seed,
                                    uint64_t subseq) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // NHWC linear element index
  if (e >= (int64_t)B * HW * C) return;
  bool keep[8];
  dropout_keep8(dropout_key(seed, subseq), e & ~(int64_t)7, drop_thresh(p), keep);   // same stream as norm.cu / conv_tc.cu
  const int c = (int)(e % C);
  const int64_t pix = e / C, b = pix / HW, q = pix % HW;
  mask[(b * C + c) * HW + q] = keep[e & 7] ? 1 : 0;
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

int pub_unet_create(const pub_unet_block* enc, int n_enc, const pub_unet_block* dec, int n_dec, int in_channels,
                    int out_channels, float dropout, int dtype, pub_unet** out) {
  PUB_REQUIRE(enc && dec && out && n_enc > 0 && n_dec > 0, "pub_unet_create: bad arguments");
  PUB_REQUIRE(dtype == PUB_F32 || dtype == PUB_BF16, "pub_unet_create: bad dtype");
  PUB_REQUIRE(in_channels >= 1 && in_channels <= 8, "pub_unet_create: in_channels must be in [1, 8]");
  PUB_REQUIRE(enc[0].is_conv, "pub_unet_create: the first encoder entry must be the plain conv (src/networks.py:269)");
  pub_unet* u = new pub_unet();
  u->blocks.assign(enc, enc + n_enc);
  u->blocks.insert(u->blocks.end(), dec, dec + n_dec);
  u->n_enc = n_enc; u->in_ch = in_channels; u->out_ch = out_channels; u->dropout = dropout; u->dtype = dtype;
  int np = 0;
  for (auto& b : u->blocks) {
    if (b.is_conv) np += 2;
    else {
      np += 9 + (b.has_skip_conv ? 2 : 0);
      if ((b.up || b.down) && b.has_skip_conv) { delete u; set_error("pub_unet_create: resampling blocks with a 1x1 skip conv are not built"); return -1; }
      if (!b.has_skip_conv && !(b.up || b.down) && b.cin != b.cout) { delete u; set_error("pub_unet_create: identity skip needs cin == cout"); return -1; }
    }
  }
  u->nparams = np + 4;
  u->final_c = u->blocks.back().cout;
  *out = u;
  return 0;
}

void pub_unet_destroy(pub_unet* u) { delete u; }
int pub_unet_num_params(const pub_unet* u) { return u ? u->nparams : 0; }

size_t pub_unet_workspace_bytes(const pub_unet* u, int B, int H, int W) {
  Plan pl;
  if (!u || build_plan(u, B, H, W, nullptr, 0, pl) != 0) return 0;
  return pl.total;
}

int pub_unet_forward(pub_unet* u, int B, int H, int W, const float* x_nchw, const float* const* P, void* out,
                     int out_nchw, void* ws, size_t ws_bytes, uint64_t seed, int training, int backend, pub_stream_t s_) {
  PUB_REQUIRE(u && x_nchw && P && out && ws, "pub_unet_forward: null argument");
  cudaStream_t s = (cudaStream_t)s_;
  Plan pl;
  PUB_TRY(build_plan(u, B, H, W, ws, ws_bytes, pl));
  const int dt = u->dtype;
  const float pdrop = training ? u->dropout : 0.f;
  PUB_TRY(nchw_to_nhwc(x_nchw, u->in_ch, nullptr, 0, pl.x_in, 8, B, H, W, dt, s));
  PUB_TRY(pack_all(u, pl, P, 0, dt, s));
  // GroupNorm statistics are produced by the epilogue of the conv that writes the tensor (conv_fused_rows() > 0: the
  // tcgen05 halo kernel in bf16); out_rows[i] = partial rows per image that entry i's output carries (0: none, the
  // consuming GroupNorm runs its own statistics pass)
  std::vector<int> out_rows(pl.bp.size(), 0);
  auto pre_of = [&](GnParams& g, int src0, int src1) {
    if (src0 < 0 || out_rows[src0] == 0) return;
    if (src1 >= 0 && out_rows[src1] != out_rows[src0]) return;
    g.pre0 = pl.bp[src0].out_part; g.pre1 = src1 >= 0 ? pl.bp[src1].out_part : nullptr; g.pre_rows = out_rows[src0];
  };
  for (size_t i = 0; i < pl.bp.size(); ++i) {
    BlockPlan& b = pl.bp[i];
    View v0 = src_view(pl, b.src0, 8);
    if (b.src0 < 0) v0.c = u->in_ch;
    View v1 = b.src1 >= 0 ? src_view(pl, b.src1, 0) : View{nullptr, 0, 0};
    const float* const* p = P + b.pidx;
    if (b.d.is_conv) {
      ConvParams c = conv_params(v0.p, b.c0, v0.ld, nullptr, 0, 0, b.wp0, p[1], nullptr, 0, b.out, b.d.cout, B, b.Ho, b.Wo, b.d.cout, 3);
      out_rows[i] = conv_fused_rows(c, dt, backend);
      if (out_rows[i]) c.stat_part = b.out_part;
      PUB_TRY(conv_forward(c, dt, backend, s));
      continue;
    }
    const int mode = b.d.down ? 1 : (b.d.up ? 2 : 0);
    // a0 = resample(silu(norm0(x)))
    GnParams g0 = gn_params(pl, v0.p, b.c0, v0.ld, v1.p, b.c1, v1.ld, b.Hi, b.Wi, p[0], p[1], nullptr, mode, 0.f, 0, 0, b.stats0, b.coef0);
    pre_of(g0, b.src0, b.src1);
    PUB_TRY(gn_forward(g0, b.a0, dt, s));
    // h = conv0(a0) + bias
    ConvParams c0 = conv_params(b.a0, b.d.cin, b.d.cin, nullptr, 0, 0, b.wp0, p[3], nullptr, 0, b.h, b.d.cout, B, b.Ho, b.Wo, b.d.cout, 3);
    const int h_rows = conv_fused_rows(c0, dt, backend);
    if (h_rows) c0.stat_part = b.h_part;
    PUB_TRY(conv_forward(c0, dt, backend, s));
    // a1 = dropout(silu(shift + norm1(h) * (scale + 1)))
    GnParams g1 = gn_params(pl, b.h, b.d.cout, b.d.cout, nullptr, 0, 0, b.Ho, b.Wo, p[5], p[6], p[4], 0, pdrop, seed, (uint64_t)i, b.stats1, b.coef1);
    if (h_rows) { g1.pre0 = b.h_part; g1.pre_rows = h_rows; }
    PUB_TRY(gn_forward(g1, b.a1, dt, s));
    // residual branch
    const void* res; int ld_res;
    if (b.d.has_skip_conv) {
      ConvParams cs = conv_params(v0.p, b.c0, v0.ld, v1.p, b.c1, v1.ld, b.wps, p[10], nullptr, 0, pl.skipbuf, b.d.cout, B, b.Ho, b.Wo, b.d.cout, 1);
      PUB_TRY(conv_forward(cs, dt, backend, s));
      res = pl.skipbuf; ld_res = b.d.cout;
    } else if (mode) {
      PUB_TRY(resample2x(v0.p, v0.ld, b.d.cin, pl.skipbuf, B, b.Hi, b.Wi, mode, dt, s));
      res = pl.skipbuf; ld_res = b.d.cout;
    } else {
      PUB_REQUIRE(b.c1 == 0, "unet: identity skip on a concatenated input (block %zu)", i);
      res = v0.p; ld_res = v0.ld;
    }
    // out = conv1(a1) + bias + residual
    ConvParams c1 = conv_params(b.a1, b.d.cout, b.d.cout, nullptr, 0, 0, b.wp1, p[8], res, ld_res, b.out, b.d.cout, B, b.Ho, b.Wo, b.d.cout, 3);
    out_rows[i] = conv_fused_rows(c1, dt, backend);
    if (out_rows[i]) c1.stat_part = b.out_part;
    PUB_TRY(conv_forward(c1, dt, backend, s));
  }
  const BlockPlan& last = pl.bp.back();
  const float* const* p = P + (u->nparams - 4);
  const int fc = u->final_c;
  GnParams go = gn_params(pl, last.out, fc, fc, nullptr, 0, 0, H, W, p[0], p[1], nullptr, 0, 0.f, 0, 0, pl.stats_o, pl.coef_o);
  pre_of(go, (int)pl.bp.size() - 1, -1);
  PUB_TRY(gn_forward(go, pl.a_out, dt, s));
  void* y = out_nchw ? pl.y_out : out;
  ConvParams co = conv_params(pl.a_out, fc, fc, nullptr, 0, 0, pl.wp_out, p[3], nullptr, 0, y, u->out_ch, B, H, W, u->out_ch, 3);
  PUB_TRY(conv_forward(co, dt, backend, s));
  if (out_nchw) PUB_TRY(nhwc_to_nchw(pl.y_out, u->out_ch, u->out_ch, (float*)out, B, H, W, dt, 0, s));
  return 0;
}

int pub_unet_backward(pub_unet* u, int B, int H, int W, const void* dout, int dout_nchw, const float* const* P,
                      float* const* G, float* dx_nchw, void* ws, size_t ws_bytes, uint64_t seed, int training,
                      int backend, pub_stream_t s_) {
  PUB_REQUIRE(u && dout && P && G && ws, "pub_unet_backward: null argument");
  cudaStream_t s = (cudaStream_t)s_;
  Plan pl;
  PUB_TRY(build_plan(u, B, H, W, ws, ws_bytes, pl));
  const int dt = u->dtype;
  const float pdrop = training ? u->dropout : 0.f;
  const int fc = u->final_c;
  // ---- output head: y = out_conv(a_out), a_out = silu(out_norm(x_last))
  const void* dy = dout;
  if (dout_nchw) {
    PUB_TRY(nchw_to_nhwc((const float*)dout, u->out_ch, nullptr, 0, pl.dy_in, u->out_ch, B, H, W, dt, s));
    dy = pl.dy_in;
  }
  PUB_TRY(pack_all(u, pl, P, 1, dt, s));
  const BlockPlan& last = pl.bp.back();
  {
    const float* const* p = P + (u->nparams - 4);
    float* const* g = G + (u->nparams - 4);
    WgradParams wp{};
    wp.x0 = pl.a_out; wp.c0 = fc; wp.ld0 = fc; wp.dy = dy; wp.ld_dy = u->out_ch; wp.dw = g[2]; wp.dbias = g[3];
    wp.B = B; wp.H = H; wp.W = W; wp.cout = u->out_ch; wp.ks = 3;
    PUB_TRY(wgrad(wp, dt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
    ConvParams cd = conv_params(dy, u->out_ch, u->out_ch, nullptr, 0, 0, pl.wp_out, nullptr, nullptr, 0, pl.s1, fc, B, H, W, fc, 3);
    GnParams go = gn_params(pl, last.out, fc, fc, nullptr, 0, 0, H, W, p[0], p[1], nullptr, 0, 0.f, 0, 0, pl.stats_o, pl.coef_o);
    // the data-gradient conv stores du = g * silu'(a x + b) and its partial sums (GroupNorm-backward prologue fused)
    const int rows = g_opt_gn_fuse >= 2 ? conv_fused_rows(cd, dt, backend) : 0;
    if (rows) { cd.stat_part = pl.du_part; cd.gn_bwd = 1; cd.gx0 = last.out; cd.gc0 = fc; cd.gld0 = fc; cd.gcoef = pl.coef_o; }
    PUB_TRY(conv_forward(cd, dt, backend, s));
    if (rows) PUB_TRY(gn_backward_from_du(go, pl.s1, pl.du_part, rows, pl.dx_last, nullptr, 0, g[0], g[1], nullptr, dt, s));
    else PUB_TRY(gn_backward(go, pl.s1, pl.dx_last, nullptr, 0, g[0], g[1], nullptr, dt, s));
  }
  // gradient wrt the output of entry i, as a view: consumers are (a) the next entry in execution order
  // (slice [0, cout) of its dx) and (b) for encoder entries the decoder block that concatenated it.
  std::vector<int> skip_consumer(pl.bp.size(), -1);
  for (size_t j = 0; j < pl.bp.size(); ++j)
    if (pl.bp[j].src1 >= 0) skip_consumer[pl.bp[j].src1] = (int)j;
  for (int i = (int)pl.bp.size() - 1; i >= 0; --i) {
    BlockPlan& b = pl.bp[i];
    const float* const* p = P + b.pidx;
    float* const* g = G + b.pidx;
    // ---- assemble g_out
    const void* gp; int gld;
    if (i == (int)pl.bp.size() - 1) { gp = pl.dx_last; gld = fc; }
    else { gp = pl.bp[i + 1].dx; gld = pl.bp[i + 1].d.cin; }
    if (skip_consumer[i] >= 0) {
      const BlockPlan& cns = pl.bp[skip_consumer[i]];
      const char* sp = (const char*)cns.dx + (size_t)cns.c0 * pl.es;
      PUB_TRY(add_views(gp, gld, sp, cns.d.cin, pl.gsum, b.d.cout, b.d.cout, (int64_t)B * b.Ho * b.Wo, dt, s));
      gp = pl.gsum; gld = b.d.cout;
    }
    View v0 = src_view(pl, b.src0, 8);
    if (b.src0 < 0) v0.c = u->in_ch;
    View v1 = b.src1 >= 0 ? src_view(pl, b.src1, 0) : View{nullptr, 0, 0};
    if (b.d.is_conv) {
      WgradParams wp{};
      wp.x0 = v0.p; wp.c0 = b.c0; wp.ld0 = v0.ld; wp.dy = gp; wp.ld_dy = gld; wp.dw = g[0]; wp.dbias = g[1];
      wp.B = B; wp.H = b.Ho; wp.W = b.Wo; wp.cout = b.d.cout; wp.ks = 3;
      PUB_TRY(wgrad(wp, dt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
      if (dx_nchw) {
        ConvParams cd = conv_params(gp, b.d.cout, gld, nullptr, 0, 0, b.wp0, nullptr, nullptr, 0, pl.s1, b.d.cin, B, b.Ho, b.Wo, b.d.cin, 3);
        PUB_TRY(conv_simt(cd, dt, s));
        PUB_TRY(nhwc_to_nchw(pl.s1, b.d.cin, b.d.cin, dx_nchw, B, H, W, dt, 0, s));
      }
      continue;
    }
    const int mode = b.d.down ? 1 : (b.d.up ? 2 : 0);
    // ---- residual branch -> addend for dx
    const void* addp; int add_ld;
    if (b.d.has_skip_conv) {
      WgradParams wp{};
      wp.x0 = v0.p; wp.c0 = b.c0; wp.ld0 = v0.ld; wp.x1 = v1.p; wp.c1 = b.c1; wp.ld1 = v1.ld;
      wp.dy = gp; wp.ld_dy = gld; wp.dw = g[9]; wp.dbias = g[10];
      wp.B = B; wp.H = b.Ho; wp.W = b.Wo; wp.cout = b.d.cout; wp.ks = 1;
      PUB_TRY(wgrad(wp, dt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
      ConvParams cd = conv_params(gp, b.d.cout, gld, nullptr, 0, 0, b.wps, nullptr, nullptr, 0, pl.sadd, b.d.cin, B, b.Ho, b.Wo, b.d.cin, 1);
      PUB_TRY(conv_forward(cd, dt, backend, s));
      addp = pl.sadd; add_ld = b.d.cin;
    } else if (mode) {
      PUB_TRY(resample2x_bwd(gp, gld, b.d.cout, pl.sadd, B, b.Hi, b.Wi, mode, dt, s));
      addp = pl.sadd; add_ld = b.d.cin;
    } else {
      addp = gp; add_ld = gld;
    }
    // ---- conv1
    int rows1 = 0, rows0 = 0;
    {
      WgradParams wp{};
      wp.x0 = b.a1; wp.c0 = b.d.cout; wp.ld0 = b.d.cout; wp.dy = gp; wp.ld_dy = gld; wp.dw = g[7]; wp.dbias = g[8];
      wp.B = B; wp.H = b.Ho; wp.W = b.Wo; wp.cout = b.d.cout; wp.ks = 3;
      PUB_TRY(wgrad(wp, dt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
      ConvParams cd = conv_params(gp, b.d.cout, gld, nullptr, 0, 0, b.wp1, nullptr, nullptr, 0, pl.s1, b.d.cout, B, b.Ho, b.Wo, b.d.cout, 3);
      rows1 = g_opt_gn_fuse >= 2 ? conv_fused_rows(cd, dt, backend) : 0;
      if (rows1) {
        cd.stat_part = pl.du_part; cd.gn_bwd = 1; cd.gx0 = b.h; cd.gc0 = b.d.cout; cd.gld0 = b.d.cout; cd.gcoef = b.coef1;
        cd.p_drop = pdrop; cd.seed = seed; cd.subseq = (uint64_t)i; cd.salt = g_seed_salt;
      }
      PUB_TRY(conv_forward(cd, dt, backend, s));
    }
    // ---- norm1 / FiLM / SiLU / dropout
    GnParams g1 = gn_params(pl, b.h, b.d.cout, b.d.cout, nullptr, 0, 0, b.Ho, b.Wo, p[5], p[6], p[4], 0, pdrop, seed, (uint64_t)i, b.stats1, b.coef1);
    if (rows1) PUB_TRY(gn_backward_from_du(g1, pl.s1, pl.du_part, rows1, pl.s2, nullptr, 0, g[5], g[6], g[4], dt, s));
    else PUB_TRY(gn_backward(g1, pl.s1, pl.s2, nullptr, 0, g[5], g[6], g[4], dt, s));
    // ---- conv0
    {
      WgradParams wp{};
      wp.x0 = b.a0; wp.c0 = b.d.cin; wp.ld0 = b.d.cin; wp.dy = pl.s2; wp.ld_dy = b.d.cout; wp.dw = g[2]; wp.dbias = g[3];
      wp.B = B; wp.H = b.Ho; wp.W = b.Wo; wp.cout = b.d.cout; wp.ks = 3;
      PUB_TRY(wgrad(wp, dt, backend, pl.wg_ws, pl.wg_ws_bytes, 0, s));
      ConvParams cd = conv_params(pl.s2, b.d.cout, b.d.cout, nullptr, 0, 0, b.wp0, nullptr, nullptr, 0, pl.s1, b.d.cin, B, b.Ho, b.Wo, b.d.cin, 3);
      rows0 = (mode == 0 && g_opt_gn_fuse >= 2) ? conv_fused_rows(cd, dt, backend) : 0;   // resampling blocks: GroupNorm at another resolution
      if (rows0) {
        cd.stat_part = pl.du_part; cd.gn_bwd = 1; cd.gx0 = v0.p; cd.gx1 = v1.p; cd.gc0 = b.c0; cd.gld0 = v0.ld; cd.gld1 = v1.ld;
        cd.gcoef = b.coef0;
      }
      PUB_TRY(conv_forward(cd, dt, backend, s));
    }
    // ---- norm0 / SiLU / resample  (+ residual-branch gradient)
    GnParams g0 = gn_params(pl, v0.p, b.c0, v0.ld, v1.p, b.c1, v1.ld, b.Hi, b.Wi, p[0], p[1], nullptr, mode, 0.f, 0, 0, b.stats0, b.coef0);
    if (rows0) PUB_TRY(gn_backward_from_du(g0, pl.s1, pl.du_part, rows0, b.dx, addp, add_ld, g[0], g[1], nullptr, dt, s));
    else PUB_TRY(gn_backward(g0, pl.s1, b.dx, addp, add_ld, g[0], g[1], nullptr, dt, s));
  }
  return 0;
}

int pub_unet_dropout_mask(const pub_unet* u, int block_index, int B, int H, int W, uint64_t seed, uint8_t* mask_nchw,
                          pub_stream_t s) {
  PUB_REQUIRE(u && mask_nchw, "pub_unet_dropout_mask: null argument");
  Plan pl;
  PUB_TRY(build_plan(u, B, H, W, nullptr, 0, pl));
  PUB_REQUIRE(block_index >= 0 && block_index < (int)pl.bp.size() && !pl.bp[block_index].d.is_conv,
              "pub_unet_dropout_mask: block %d is not a UNetBlock", block_index);
  const BlockPlan& b = pl.bp[block_index];
  const int64_t n = (int64_t)B * b.Ho * b.Wo * b.d.cout;
  dropout_mask_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)s>>>(mask_nchw, B, b.Ho * b.Wo, b.d.cout, u->dropout, seed,
                                                                  (uint64_t)block_index);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
