// WMSE + (1 - MS-SSIM) reconstruction term of the reference's ACTIVE elbo, value and gradient in one call.
//
//   src/prob_unet_utils.py:270-305  wmse_ms_ssim_loss:  L = lam * WMSE + (1 - lam) * (1 - MS-SSIM)
//       WMSE      = mean( min(alpha * exp(beta * y), 1) * (x - y)^2 )
//       data_range = clamp(max(y) - min(y), 1e-5)       (batch-global; a device scalar here, no host sync)
//   pytorch_msssim.ms_ssim v1.0.0 (third party, restated -- oracle/probunet_oracle.py::ms_ssim, parity unpinned):
//       5 levels, 7-tap Gaussian (sigma 1.5) VALID window, 2x2 mean pooling between levels,
//       per (sample, channel):  prod_l relu(v_l)^w_l,  v_l = mean cs_l (l < 4), v_4 = mean ssim_4;  mean over (b, c).
//
// Kernels (all f32, NCHW planes, deterministic two-stage reductions):
//   minmax -> pyramid (pool + WMSE sum) -> level_stats x5 -> finalize -> [level_coef -> level_grad] x5, coarse to fine.
// level_coef stores, per VALID window p, the three numbers the chain rule needs:
//   dL/dX[q] = sum_p G[q-p] * (a_p + 2 X[q] b_p + Y[q] c_p)      (a: via mu_x, b: via E[x^2], c: via E[xy])
// and level_grad gathers them over the <= 49 windows that contain pixel q, adds the pooled gradient of the
// next-coarser level (/4) and, at level 0, the WMSE term.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.h"

namespace pub {
namespace {

constexpr int LEVELS = 5, WIN = 7;
__constant__ float c_g[WIN];              // normalised 1-D Gaussian
__constant__ float c_w[LEVELS] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};

struct Lvl { int h, w, vh, vw; };         // image size and VALID output size (h - 6, w - 6)

struct Window { float mu1, mu2, e11, e22, e12; };
__device__ __forceinline__ Window window_stats(const float* __restrict__ X, const float* __restrict__ Y, int w, int y, int x) {
  Window s{0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < WIN; ++i) {
    float r1 = 0.f, r2 = 0.f, r11 = 0.f, r22 = 0.f, r12 = 0.f;   // separable: rows first (as the reference filters)
    const float* xr = X + (int64_t)(y + i) * w + x;
    const float* yr = Y + (int64_t)(y + i) * w + x;
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
      const float a = xr[j], b = yr[j], g = c_g[j];
      r1 = fmaf(g, a, r1); r2 = fmaf(g, b, r2);
      r11 = fmaf(g, a * a, r11); r22 = fmaf(g, b * b, r22); r12 = fmaf(g, a * b, r12);
    }
    const float g = c_g[i];
    s.mu1 = fmaf(g, r1, s.mu1); s.mu2 = fmaf(g, r2, s.mu2);
    s.e11 = fmaf(g, r11, s.e11); s.e22 = fmaf(g, r22, s.e22); s.e12 = fmaf(g, r12, s.e12);
  }
  return s;
}

// scal[0] = min(y), scal[1] = max(y) partial results -> scal[2] = C1, scal[3] = C2
__global__ void minmax_kernel(const float* __restrict__ y, int64_t n, float* __restrict__ part) {
  __shared__ float smn[8], smx[8];
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = y[i];
    mn = fminf(mn, v); mx = fmaxf(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x / 32; ++i) { mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    part[2 * blockIdx.x] = mn; part[2 * blockIdx.x + 1] = mx;
  }
}
__global__ void range_kernel(const float* __restrict__ part, int nblk, float* __restrict__ scal) {
  float mn = INFINITY, mx = -INFINITY;
  for (int i = 0; i < nblk; ++i) { mn = fminf(mn, part[2 * i]); mx = fmaxf(mx, part[2 * i + 1]); }
  const float R = fmaxf(mx - mn, 1e-5f);
  scal[0] = R; scal[2] = (0.01f * R) * (0.01f * R); scal[3] = (0.03f * R) * (0.03f * R);
}

// 2x2 mean pooling of X and Y (even sizes); at level 0 also the WMSE partial sum of the 4 pixels each thread reads
__global__ void pool_kernel(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ Xo,
                            float* __restrict__ Yo, int planes, int h, int w, int with_wmse, float alpha, float beta,
                            float* __restrict__ wpart) {
  __shared__ float red[8];
  const int ho = h / 2, wo = w / 2;
  const int64_t n = (int64_t)planes * ho * wo;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % wo), yo = (int)((i / wo) % ho);
    const int64_t pl = i / ((int64_t)wo * ho);
    const float* xp = X + (pl * h + 2 * yo) * w + 2 * xo;
    const float* yp = Y + (pl * h + 2 * yo) * w + 2 * xo;
    const float x0 = xp[0], x1 = xp[1], x2 = xp[w], x3 = xp[w + 1];
    const float y0 = yp[0], y1 = yp[1], y2 = yp[w], y3 = yp[w + 1];
    Xo[i] = 0.25f * ((x0 + x1) + (x2 + x3));
    Yo[i] = 0.25f * ((y0 + y1) + (y2 + y3));
    if (with_wmse) {
      acc += fminf(alpha * __expf(beta * y0), 1.f) * (x0 - y0) * (x0 - y0) + fminf(alpha * __expf(beta * y1), 1.f) * (x1 - y1) * (x1 - y1) +
             fminf(alpha * __expf(beta * y2), 1.f) * (x2 - y2) * (x2 - y2) + fminf(alpha * __expf(beta * y3), 1.f) * (x3 - y3) * (x3 - y3);
    }
  }
  if (with_wmse) {
    const float s = block_sum<256>(acc, red);
    if (threadIdx.x == 0) wpart[blockIdx.x] = s;
  }
}

// per (plane, chunk): partial sums of cs and ssim over the VALID windows
__global__ void level_stats_kernel(const float* __restrict__ X, const float* __restrict__ Y, Lvl L,
                                   const float* __restrict__ scal, float* __restrict__ part /* [plane][chunk][2] */) {
  __shared__ float red[8];
  const int plane = blockIdx.y, nv = L.vh * L.vw;
  const float C1 = scal[2], C2 = scal[3];
  const float* Xp = X + (int64_t)plane * L.h * L.w;
  const float* Yp = Y + (int64_t)plane * L.h * L.w;
  float scs = 0.f, sss = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
    const Window s = window_stats(Xp, Yp, L.w, i / L.vw, i % L.vw);
    const float s11 = s.e11 - s.mu1 * s.mu1, s22 = s.e22 - s.mu2 * s.mu2, s12 = s.e12 - s.mu1 * s.mu2;
    const float cs = (2.f * s12 + C2) / (s11 + s22 + C2);
    const float lum = (2.f * s.mu1 * s.mu2 + C1) / (s.mu1 * s.mu1 + s.mu2 * s.mu2 + C1);
    scs += cs; sss += lum * cs;
  }
  const float a = block_sum<256>(scs, red);
  const float b = block_sum<256>(sss, red);
  if (threadIdx.x == 0) {
    part[((int64_t)plane * gridDim.x + blockIdx.x) * 2] = a;
    part[((int64_t)plane * gridDim.x + blockIdx.x) * 2 + 1] = b;
  }
}

struct FinArgs {
  const float* part[LEVELS]; int nchunk[LEVELS]; int nvalid[LEVELS];
  const float* wpart; int nwpart; int64_t nelem;
  int planes; float lam;
};
// one CTA: per-plane MS-SSIM, the three output scalars, and the upstream coefficient of every (level, plane):
//   gco[l][plane] = dL/d(mean map_l) / nvalid_l
__global__ void finalize_kernel(FinArgs f, float* __restrict__ out3, float* __restrict__ gco) {
  __shared__ double sred[256];
  double acc = 0.0;
  for (int p = threadIdx.x; p < f.planes; p += blockDim.x) {
    double v[LEVELS];
    for (int l = 0; l < LEVELS; ++l) {
      double s = 0.0;
      for (int k = 0; k < f.nchunk[l]; ++k) s += (double)f.part[l][((int64_t)p * f.nchunk[l] + k) * 2 + (l == LEVELS - 1 ? 1 : 0)];
      v[l] = s / f.nvalid[l];
      if (v[l] < 0.0) v[l] = 0.0;                      // relu
    }
    double P = 1.0;
    for (int l = 0; l < LEVELS; ++l) P *= pow(v[l], (double)c_w[l]);
    acc += P;
    // L = lam*wmse + (1-lam)*(1 - mean_p P)  ->  dL/dv_l = -(1-lam)/planes * w_l * P / v_l
    for (int l = 0; l < LEVELS; ++l) {
      const double g = v[l] > 0.0 ? -(1.0 - f.lam) / f.planes * c_w[l] * P / v[l] : 0.0;
      gco[l * f.planes + p] = (float)(g / f.nvalid[l]);
    }
  }
  sred[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) s += sred[i];
    const double ms = 1.0 - s / f.planes;
    double w = 0.0;
    for (int i = 0; i < f.nwpart; ++i) w += (double)f.wpart[i];
    w /= (double)f.nelem;
    out3[0] = (float)(f.lam * w + (1.0 - f.lam) * ms);
    out3[1] = (float)w;
    out3[2] = (float)ms;
  }
}

// per VALID window: (a, b, c) scaled by the upstream coefficient of this (level, plane)
__global__ void level_coef_kernel(const float* __restrict__ X, const float* __restrict__ Y, Lvl L, int last,
                                  const float* __restrict__ scal, const float* __restrict__ gco /* [plane] */,
                                  float* __restrict__ abc /* [plane][3][nv] */) {
  const int plane = blockIdx.y, nv = L.vh * L.vw;
  const float C1 = scal[2], C2 = scal[3], up = gco[plane];
  const float* Xp = X + (int64_t)plane * L.h * L.w;
  const float* Yp = Y + (int64_t)plane * L.h * L.w;
  float* o = abc + (int64_t)plane * 3 * nv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
    const Window s = window_stats(Xp, Yp, L.w, i / L.vw, i % L.vw);
    const float s11 = s.e11 - s.mu1 * s.mu1, s22 = s.e22 - s.mu2 * s.mu2, s12 = s.e12 - s.mu1 * s.mu2;
    const float A2 = 2.f * s12 + C2, B2 = s11 + s22 + C2;
    const float cs = A2 / B2;
    float dcs = up, dmu = 0.f;                         // d/dcs, direct d/dmu1 (through the luminance term)
    if (last) {
      const float A1 = 2.f * s.mu1 * s.mu2 + C1, B1 = s.mu1 * s.mu1 + s.mu2 * s.mu2 + C1;
      const float lum = A1 / B1;
      dcs = up * lum;
      dmu = up * cs * (2.f * s.mu2 * B1 - A1 * 2.f * s.mu1) / (B1 * B1);
    }
    const float ds12 = dcs * 2.f / B2, ds11 = -dcs * A2 / (B2 * B2);
    o[i] = dmu - ds12 * s.mu2 - 2.f * ds11 * s.mu1;   // a: total d/dmu1
    o[nv + i] = ds11;                                  // b: d/dE[x^2]
    o[2 * nv + i] = ds12;                              // c: d/dE[xy]
  }
}

// dX_l[q] = sum_p G[q-p] (a_p + 2 X[q] b_p + Y[q] c_p) + dX_{l+1}[q/2] / 4 (+ lam * 2 w (x - y) / N at level 0)
__global__ void level_grad_kernel(const float* __restrict__ X, const float* __restrict__ Y, Lvl L,
                                  const float* __restrict__ abc, const float* __restrict__ dcoarse /* or null */,
                                  float* __restrict__ dX, int planes, int wmse, float alpha, float beta, float lamN) {
  const int64_t n = (int64_t)planes * L.h * L.w;
  const int nv = L.vh * L.vw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % L.w), y = (int)((i / L.w) % L.h);
    const int64_t pl = i / ((int64_t)L.w * L.h);
    const float* o = abc + pl * 3 * nv;
    float ga = 0.f, gb = 0.f, gc = 0.f;
#pragma unroll
    for (int dy = 0; dy < WIN; ++dy) {
      const int py = y - dy;
      if (py < 0 || py >= L.vh) continue;
      float ra = 0.f, rb = 0.f, rc = 0.f;
#pragma unroll
      for (int dx = 0; dx < WIN; ++dx) {
        const int px = x - dx;
        if (px < 0 || px >= L.vw) continue;
        const int k = py * L.vw + px;
        const float g = c_g[dx];
        ra = fmaf(g, o[k], ra); rb = fmaf(g, o[nv + k], rb); rc = fmaf(g, o[2 * nv + k], rc);
      }
      const float g = c_g[dy];
      ga = fmaf(g, ra, ga); gb = fmaf(g, rb, gb); gc = fmaf(g, rc, gc);
    }
    const float xv = X[i], yv = Y[i];
    float d = ga + 2.f * xv * gb + yv * gc;
    if (dcoarse) d += 0.25f * dcoarse[(pl * (L.h / 2) + y / 2) * (L.w / 2) + x / 2];
    if (wmse) d += lamN * 2.f * fminf(alpha * __expf(beta * yv), 1.f) * (xv - yv);
    dX[i] = d;
  }
}

struct Plan {
  Lvl lv[LEVELS];
  size_t off_x[LEVELS], off_y[LEVELS], off_dx[LEVELS];   // float offsets (level 0 = caller tensors)
  size_t off_abc, off_part[LEVELS], off_wpart, off_mm, off_scal, off_gco, total;
  int nchunk[LEVELS];
};
constexpr int NWPART = 512, NMM = 256;

int make_plan(int planes, int H, int W, Plan& p) {
  PUB_REQUIRE(H % 16 == 0 && W % 16 == 0, "ms-ssim kernel needs H and W divisible by 16 (got %dx%d)", H, W);
  PUB_REQUIRE(H > (WIN - 1) * 16 && W > (WIN - 1) * 16, "ms_ssim: image side must exceed %d (pytorch_msssim assert; got %dx%d)",
              (WIN - 1) * 16, H, W);
  size_t off = 0;
  auto take = [&](size_t n) { size_t r = off; off += (n + 63) / 64 * 64; return r; };
  for (int l = 0; l < LEVELS; ++l) {
    p.lv[l] = Lvl{H >> l, W >> l, (H >> l) - (WIN - 1), (W >> l) - (WIN - 1)};
    const size_t n = (size_t)planes * p.lv[l].h * p.lv[l].w;
    if (l) { p.off_x[l] = take(n); p.off_y[l] = take(n); p.off_dx[l] = take(n); }
    p.nchunk[l] = cdiv((int64_t)p.lv[l].vh * p.lv[l].vw, 256 * 4);
    p.off_part[l] = take((size_t)planes * p.nchunk[l] * 2);
  }
  p.off_abc = take((size_t)planes * 3 * p.lv[0].vh * p.lv[0].vw);
  p.off_wpart = take(NWPART); p.off_mm = take(2 * NMM); p.off_scal = take(8); p.off_gco = take((size_t)LEVELS * planes);
  p.total = off * sizeof(float);
  return 0;
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

size_t pub_msssim_workspace(int B, int C, int H, int W) {
  Plan p;
  if (make_plan(B * C, H, W, p) != 0) return 0;
  return p.total + 256;
}

int pub_wmse_msssim_loss(const float* pred, const float* target, int B, int C, int H, int W, float alpha, float beta,
                         float lam, float* out3, float* dpred, void* ws, size_t ws_bytes, pub_stream_t s_) {
  PUB_REQUIRE(pred && target && out3 && ws, "pub_wmse_msssim_loss: null argument");
  const int planes = B * C;
  Plan p;
  PUB_TRY(make_plan(planes, H, W, p));
  PUB_REQUIRE(ws_bytes >= p.total, "pub_wmse_msssim_loss: workspace too small (%zu < %zu)", ws_bytes, p.total);
  cudaStream_t s = (cudaStream_t)s_;
  static bool init = false;
  if (!init) {
    double g[WIN], sum = 0.0;
    for (int i = 0; i < WIN; ++i) { g[i] = exp(-(double)((i - WIN / 2) * (i - WIN / 2)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
    float gf[WIN];
    for (int i = 0; i < WIN; ++i) gf[i] = (float)(g[i] / sum);
    PUB_CUDA(cudaMemcpyToSymbol(c_g, gf, sizeof(gf)));
    init = true;
  }
  float* w = (float*)ws;
  const float* X[LEVELS]; const float* Y[LEVELS]; float* dX[LEVELS];
  X[0] = pred; Y[0] = target; dX[0] = dpred;
  for (int l = 1; l < LEVELS; ++l) { X[l] = w + p.off_x[l]; Y[l] = w + p.off_y[l]; dX[l] = w + p.off_dx[l]; }
  float* scal = w + p.off_scal;
  const int64_t n0 = (int64_t)planes * H * W;
  minmax_kernel<<<NMM, 256, 0, s>>>(target, n0, w + p.off_mm);
  PUB_LAUNCH_CHECK();
  range_kernel<<<1, 1, 0, s>>>(w + p.off_mm, NMM, scal);
  PUB_LAUNCH_CHECK();
  for (int l = 0; l + 1 < LEVELS; ++l) {
    const int64_t n = (int64_t)planes * p.lv[l + 1].h * p.lv[l + 1].w;
    const int grid = l == 0 ? NWPART : (int)std::min<int64_t>(NWPART, (n + 255) / 256);
    pool_kernel<<<grid, 256, 0, s>>>(X[l], Y[l], (float*)X[l + 1], (float*)Y[l + 1], planes, p.lv[l].h, p.lv[l].w, l == 0,
                                     alpha, beta, w + p.off_wpart);
    PUB_LAUNCH_CHECK();
  }
  FinArgs f{};
  for (int l = 0; l < LEVELS; ++l) {
    dim3 grid(p.nchunk[l], planes);
    level_stats_kernel<<<grid, 256, 0, s>>>(X[l], Y[l], p.lv[l], scal, w + p.off_part[l]);
    PUB_LAUNCH_CHECK();
    f.part[l] = w + p.off_part[l]; f.nchunk[l] = p.nchunk[l]; f.nvalid[l] = p.lv[l].vh * p.lv[l].vw;
  }
  f.wpart = w + p.off_wpart; f.nwpart = NWPART; f.nelem = n0; f.planes = planes; f.lam = lam;
  finalize_kernel<<<1, 256, 0, s>>>(f, out3, w + p.off_gco);
  PUB_LAUNCH_CHECK();
  if (!dpred) return 0;
  for (int l = LEVELS - 1; l >= 0; --l) {
    dim3 grid(p.nchunk[l], planes);
    level_coef_kernel<<<grid, 256, 0, s>>>(X[l], Y[l], p.lv[l], l == LEVELS - 1, scal, w + p.off_gco + (size_t)l * planes,
                                           w + p.off_abc);
    PUB_LAUNCH_CHECK();
    const int64_t n = (int64_t)planes * p.lv[l].h * p.lv[l].w;
    level_grad_kernel<<<(int)std::min<int64_t>(num_sms() * 16, (n + 255) / 256), 256, 0, s>>>(
        X[l], Y[l], p.lv[l], w + p.off_abc, l + 1 < LEVELS ? dX[l + 1] : nullptr, dX[l], planes, l == 0 && lam != 0.f, alpha,
        beta, lam / (float)n0);
    PUB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
