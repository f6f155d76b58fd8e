// C-ABI entry points (include/probunet_b200.h) for the primitive ops + shared host helpers.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace pub {

static thread_local std::string g_err;
unsigned long long g_launch_count = 0;
int g_opt_conv_halo = -1;
int g_opt_wgrad_box3 = 1;
int g_opt_fcomb_fwd_mma = 1;
int g_opt_wgrad_fused_bias = 1;
int g_opt_pdl = 1;
int g_opt_gn_fuse = 1;
int g_opt_metrics_occ = 3;
int g_opt_wgrad_swap = 1;
int g_opt_fcomb_bwd_occ = 2;
cudaStream_t g_pack_stream = nullptr;
unsigned long long g_since_pack = 0;
long long* g_halo_trace = nullptr;
const uint32_t* g_seed_salt = nullptr;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

int conv_forward(const ConvParams& p, int dtype, int backend, cudaStream_t s) {
  if (backend == PUB_BACKEND_TCGEN05) {
    PUB_REQUIRE(conv_tc_supported(p, dtype), "tcgen05 conv backend does not support this shape/dtype "
                "(c0=%d c1=%d cout=%d H=%d W=%d ks=%d dtype=%d)", p.c0, p.c1, p.cout, p.H, p.W, p.ks, dtype);
    return conv_tc(p, dtype, s);
  }
  if (backend == PUB_BACKEND_AUTO && conv_tc_supported(p, dtype)) return conv_tc(p, dtype, s);
  PUB_REQUIRE(p.stat_part == nullptr, "conv_forward: fused GroupNorm epilogue requested for a launch that takes the FMA path");
  ConvParams q = p;
  q.round_tf32 = dtype == PUB_TF32;
  return conv_simt(q, dtype, s);
}

size_t wgrad_workspace(const WgradParams& p, int dtype, int backend) {
  size_t a = wgrad_simt_workspace(p);
  if (backend != PUB_BACKEND_SIMT && wgrad_tc_supported(p, dtype)) {
    size_t b = wgrad_tc_workspace(p, dtype);
    if (b > a) a = b;
  }
  return a;
}

int wgrad(const WgradParams& p, int dtype, int backend, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s) {
  if (backend == PUB_BACKEND_TCGEN05) {
    PUB_REQUIRE(wgrad_tc_supported(p, dtype), "tcgen05 wgrad backend does not support this shape/dtype");
    return wgrad_tc(p, dtype, ws, ws_bytes, accumulate, s);
  }
  if (backend == PUB_BACKEND_AUTO && wgrad_tc_supported(p, dtype)) return wgrad_tc(p, dtype, ws, ws_bytes, accumulate, s);
  return wgrad_simt(p, dtype, ws, ws_bytes, accumulate, s);
}

static ConvParams to_params(const pub_conv_args* a) {
  ConvParams p{};
  p.x0 = a->x0; p.x1 = a->x1; p.c0 = a->c0; p.c1 = a->x1 ? a->c1 : 0; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.w = a->w; p.bias = a->bias; p.res = a->res; p.ld_res = a->ld_res; p.mask = a->mask; p.ld_mask = a->ld_mask;
  p.y = a->y; p.ldy = a->ldy; p.B = a->B; p.H = a->H; p.W = a->W; p.cout = a->cout; p.ks = a->ksize; p.relu = a->relu;
  return p;
}
static WgradParams to_params(const pub_wgrad_args* a) {
  WgradParams p{};
  p.x0 = a->x0; p.x1 = a->x1; p.c0 = a->c0; p.c1 = a->x1 ? a->c1 : 0; p.ld0 = a->ld0; p.ld1 = a->ld1;
  p.dy = a->dy; p.ld_dy = a->ld_dy; p.dw = a->dw; p.dbias = a->dbias;
  p.B = a->B; p.H = a->H; p.W = a->W; p.cout = a->cout; p.ks = a->ksize;
  return p;
}

}  // namespace pub

using namespace pub;

extern "C" {

const char* pub_last_error(void) { return g_err.c_str(); }
int pub_version(void) { return 100; }
unsigned long long pub_launch_count(void) { return g_launch_count; }

int pub_debug_option(const char* name, int value) {
  PUB_REQUIRE(name != nullptr, "pub_debug_option: null name");
  if (strcmp(name, "conv_halo") == 0) { g_opt_conv_halo = value; return 0; }
  if (strcmp(name, "wgrad_box3") == 0) { g_opt_wgrad_box3 = value; return 0; }
  if (strcmp(name, "fcomb_fwd_mma") == 0) { g_opt_fcomb_fwd_mma = value; return 0; }
  if (strcmp(name, "wgrad_fused_bias") == 0) { g_opt_wgrad_fused_bias = value; return 0; }
  if (strcmp(name, "pdl") == 0) { g_opt_pdl = value; return 0; }
  if (strcmp(name, "gn_fuse") == 0) { g_opt_gn_fuse = value; return 0; }
  if (strcmp(name, "metrics_occ") == 0) { g_opt_metrics_occ = value; return 0; }
  if (strcmp(name, "wgrad_swap") == 0) { g_opt_wgrad_swap = value; return 0; }
  if (strcmp(name, "fcomb_bwd_occ") == 0) { g_opt_fcomb_bwd_occ = value; return 0; }
  set_error("pub_debug_option: unknown option '%s'", name);
  return -1;
}

namespace {
// PUB_OPTS="name=value,name=value": the same knobs from the environment (A/B runs of unmodified drivers, e.g. bench.py)
struct EnvOpts {
  EnvOpts() {
    const char* e = getenv("PUB_OPTS");
    if (!e) return;
    std::string s(e);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t end = s.find(',', pos);
      if (end == std::string::npos) end = s.size();
      const std::string kv = s.substr(pos, end - pos);
      const size_t eq = kv.find('=');
      if (eq != std::string::npos) pub_debug_option(kv.substr(0, eq).c_str(), atoi(kv.c_str() + eq + 1));
      pos = end + 1;
    }
  }
} g_env_opts;
}  // namespace

int pub_debug_pointer(const char* name, void* p) {
  PUB_REQUIRE(name != nullptr, "pub_debug_pointer: null name");
  if (strcmp(name, "halo_trace") == 0) { g_halo_trace = (long long*)p; return 0; }
  if (strcmp(name, "seed_salt") == 0) { g_seed_salt = (const uint32_t*)p; return 0; }
  set_error("pub_debug_pointer: unknown name '%s'", name);
  return -1;
}

int pub_conv2d_forward(const pub_conv_args* a, pub_stream_t s) {
  PUB_REQUIRE(a && a->x0 && a->w && a->y, "pub_conv2d_forward: null argument");
  PUB_REQUIRE(a->ksize == 1 || a->ksize == 3, "pub_conv2d_forward: ksize must be 1 or 3");
  PUB_REQUIRE(a->dtype == PUB_F32 || a->dtype == PUB_BF16 || a->dtype == PUB_TF32, "pub_conv2d_forward: bad dtype");
  return conv_forward(to_params(a), a->dtype, a->backend, (cudaStream_t)s);
}

int pub_conv2d_fused_rows(const pub_conv_args* a) {
  if (!a || !a->x0 || !a->w || !a->y) return 0;
  return conv_fused_rows(to_params(a), a->dtype, a->backend);
}

int pub_conv2d_forward_fused(const pub_conv_args* a, const pub_conv_gn_args* g, pub_stream_t s) {
  PUB_REQUIRE(a && g && a->x0 && a->w && a->y && g->stat_part, "pub_conv2d_forward_fused: null argument");
  ConvParams p = to_params(a);
  PUB_REQUIRE(conv_fused_rows(p, a->dtype, a->backend) > 0, "pub_conv2d_forward_fused: no fused epilogue for this launch "
              "(bf16 3x3, W %% 8 == 0, H %% 16 == 0, channels %% 32 == 0 through the tcgen05 halo kernel only)");
  p.stat_part = g->stat_part; p.gn_bwd = g->gn_bwd;
  p.gx0 = g->gx0; p.gx1 = g->gx1; p.gc0 = g->gc0; p.gld0 = g->gld0; p.gld1 = g->gld1; p.gcoef = g->gcoef;
  p.p_drop = g->p_drop; p.seed = g->seed; p.subseq = g->subseq;
  return conv_forward(p, a->dtype, a->backend, (cudaStream_t)s);
}

int pub_pack_conv_weight(const float* w, void* packed, int cout, int cin, int ksize, int dtype, int tflip,
                         pub_stream_t s) {
  PUB_REQUIRE(w && packed, "pub_pack_conv_weight: null argument");
  return pack_weight(w, packed, cout, cin, ksize, dtype, tflip, (cudaStream_t)s);
}

size_t pub_conv2d_wgrad_workspace(const pub_wgrad_args* a) { return wgrad_workspace(to_params(a), a->dtype, a->backend); }

int pub_conv2d_wgrad(const pub_wgrad_args* a, pub_stream_t s) {
  PUB_REQUIRE(a && a->x0 && a->dy && a->dw && a->workspace, "pub_conv2d_wgrad: null argument");
  return wgrad(to_params(a), a->dtype, a->backend, a->workspace, a->workspace_bytes, a->accumulate, (cudaStream_t)s);
}

int pub_nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* y, int ldy, int B, int H, int W, int dtype,
                     pub_stream_t s) {
  return nchw_to_nhwc(x0, c0, x1, x1 ? c1 : 0, y, ldy, B, H, W, dtype, (cudaStream_t)s);
}
int pub_nhwc_to_nchw(const void* x, int ld, int C, float* y, int B, int H, int W, int dtype, int accumulate,
                     pub_stream_t s) {
  return nhwc_to_nchw(x, ld, C, y, B, H, W, dtype, accumulate, (cudaStream_t)s);
}

}  // extern "C"
