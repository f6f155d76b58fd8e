// Latent-space ops, reconstruction losses, ensemble metrics and the AdamW step.
//   rsample / KL         : torch.distributions Normal.rsample, kl_divergence (src/prob_unet.py:215,247,255)
//   afCRPS / CRPS / L1   : src/prob_unet_utils.py:171-268, src/prob_unet.py:357-362
//   ensemble CRPS + MAE  : src/metrics.py:11-71 (+ residual_to_hr / inverse transforms, src/climex_utils.py)
//   AdamW                : torch.optim.AdamW (src/train_prob_unet_model.py:139-141)
// All reductions: warp shuffle -> block -> per-block partial -> fixed-order final sum (deterministic).
#include "common.cuh"
#include "kernels.h"

namespace pub {
namespace {

constexpr int NT = 256;

// ------------------------------------------------------------------ rsample
__global__ void rsample_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                   const float* __restrict__ eps_in, uint64_t seed, uint64_t offset, int M, int BL,
                                   float* __restrict__ z, float* __restrict__ eps_out, const uint32_t* __restrict__ salt) {
  if (salt) seed ^= (uint64_t)__ldg(salt) << 20;   // per-step device word (CUDA-graph replays reuse the arguments)
  const int64_t n = (int64_t)M * BL;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float e;
  if (eps_in) {
    e = eps_in[i];
  } else {  // Box-Muller on two Philox uniforms; one counter per element
    const uint4 r = Philox::gen(seed, offset, (uint64_t)i);
    const float u1 = ((r.x >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = Philox::u01(r.y);
    e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
  const int bl = (int)(i % BL);
  z[i] = mu[bl] + sigma[bl] * e;
  if (eps_out) eps_out[i] = e;
}

__global__ void rsample_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ eps, int M, int BL,
                                   float* __restrict__ dmu, float* __restrict__ dsigma) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BL) return;
  float a = 0.f, b = 0.f;
  for (int m = 0; m < M; ++m) {
    const float g = dz[(int64_t)m * BL + i];
    a += g; b += g * eps[(int64_t)m * BL + i];
  }
  dmu[i] = a; dsigma[i] = b;
}

// ------------------------------------------------------------------ KL(q || p), one warp per sample
__global__ void kl_fwd_kernel(const float* __restrict__ mq, const float* __restrict__ sq, const float* __restrict__ mp,
                              const float* __restrict__ sp, int B, int L, float* __restrict__ kl) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int l = lane; l < L; l += 32) {
    const int i = b * L + l;
    const float r = sq[i] / sp[i], vr = r * r, d = (mq[i] - mp[i]) / sp[i];
    s += 0.5f * (vr + d * d - 1.f - logf(vr));
  }
  s = warp_sum(s);
  if (lane == 0) kl[b] = s;
}

__global__ void kl_bwd_kernel(const float* __restrict__ dkl, const float* __restrict__ mq, const float* __restrict__ sq,
                              const float* __restrict__ mp, const float* __restrict__ sp, int B, int L,
                              float* __restrict__ dmq, float* __restrict__ dsq, float* __restrict__ dmp,
                              float* __restrict__ dsp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const float g = dkl[i / L];
  const float d = mq[i] - mp[i], ip = 1.f / sp[i], ip2 = ip * ip;
  const float a = g * d * ip2;
  if (dmq) dmq[i] = a;
  if (dmp) dmp[i] = -a;
  if (dsq) dsq[i] = g * (sq[i] * ip2 - 1.f / sq[i]);
  if (dsp) dsp[i] = g * (-(sq[i] * sq[i] + d * d) * ip2 * ip + ip);
}

// ------------------------------------------------------------------ ensemble losses
// one thread per (b, c, pixel) column of M members held in registers
template <int MAXM>
__global__ void __launch_bounds__(NT) ens_loss_kernel(const float* __restrict__ ens, const float* __restrict__ tgt,
                                                      int B, int M, int C, int HW, float c_pair, float inv_n,
                                                      float* __restrict__ part, float* __restrict__ dens) {
  __shared__ float red[NT / 32];
  const int64_t n = (int64_t)B * C * HW;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    const int64_t b = i / ((int64_t)C * HW), r = i % ((int64_t)C * HW);
    const float* e = ens + b * M * C * HW + r;
    const float y = tgt[i];
    float x[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j) x[j] = j < M ? e[(int64_t)j * C * HW] : 0.f;
    // pairwise terms once per unordered pair: sum_{j,k} |x_j - x_k| = 2 sum_{j<k} |d|, and sign(x_j - x_k) enters the
    // gradient of member j with + and of member k with - (ties: sign 0, like torch.abs' backward)
    float s1 = 0.f, s2 = 0.f, sg[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      sg[j] = 0.f;
      if (j < M) s1 += fabsf(x[j] - y);
    }
#pragma unroll
    for (int j = 0; j < MAXM; ++j)
#pragma unroll
      for (int k = j + 1; k < MAXM; ++k)
        if (k < M) {
          const float d = x[j] - x[k];
          s2 += fabsf(d);
          const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
          sg[j] += sgn;
          sg[k] -= sgn;
        }
    s2 *= 2.f;
    if (dens) {
#pragma unroll
      for (int j = 0; j < MAXM; ++j)
        if (j < M) {
          const float dy = x[j] - y;
          const float sy = (dy > 0.f) ? 1.f : ((dy < 0.f) ? -1.f : 0.f);
          dens[b * M * C * HW + (int64_t)j * C * HW + r] = (sy / M - 2.f * c_pair * sg[j]) * inv_n;
        }
    }
    acc += s1 / M - c_pair * s2;
  }
  const float tot = block_sum<NT>(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

__global__ void final_sum_kernel(const float* __restrict__ part, int n, float scale, float* __restrict__ out) {
  // single warp, fixed order
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += (double)part[i];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) out[0] = (float)(s * scale);
}

// L1: part[block][1 + C]
__global__ void __launch_bounds__(NT) l1_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int B,
                                                int C, int HW, float inv_n, float* __restrict__ part,
                                                float* __restrict__ dout) {
  __shared__ float red[NT / 32];
  // blockIdx.y = channel
  const int c = blockIdx.y;
  const int64_t n = (int64_t)B * HW;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    const int64_t b = i / HW, p = i % HW;
    const int64_t idx = (b * C + c) * HW + p;
    const float d = out[idx] - tgt[idx];
    acc += fabsf(d);
    if (dout) dout[idx] = ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) * inv_n;
  }
  const float tot = block_sum<NT>(acc, red);
  if (threadIdx.x == 0) part[(int64_t)c * gridDim.x + blockIdx.x] = tot;
}
__global__ void l1_final_kernel(const float* __restrict__ part, int nblk, int C, double n_per_c, float* __restrict__ loss) {
  // one warp per channel (blockDim = 32*C)
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = lane; i < nblk; i += 32) s += (double)part[(int64_t)c * nblk + i];
  s = warp_sum_d(s);
  __shared__ double tot[32];
  if (lane == 0) { tot[c] = s; loss[1 + c] = (float)(s / n_per_c); }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < C; ++k) t += tot[k];
    loss[0] = (float)(t / (n_per_c * C));
  }
}

__global__ void scale_kernel(float* __restrict__ y, const float* __restrict__ scale, int64_t n) {
  const float sc = *scale;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) y[i] *= sc;
}

// ------------------------------------------------------------------ ensemble metrics
__device__ __forceinline__ float softplus_ref(float x, float c) { return x > 20.f ? x : logf(expf(x) + 1.f) - c; }

// grid (blocks, C, T): per (t, c) CRPS = mean_p [ mean_j |x_j - y| - 1/(2 M^2) sum_{j,k} |x_j - x_k| ] and
// MAE = mean_p | mean_j x_j - y |.  Members are transformed to real units on load when transform != 0
// (residual_to_hr, src/climex_utils.py:277-285, then invert_transfo_3vars, results.ipynb cell 2).
//
// A thread owns one pixel of one variable: its M members live in REGISTERS (NP = M padded to a power of two with
// +inf), are sorted by a fully unrolled bitonic network (compile-time indices only: no local memory), and the pair
// term is the sorted form  sum_{j<k} |x_j - x_k| = sum_i (2 i - M + 1) x_(i)  -- O(M log^2 M) min/max instead of the
// M (M - 1) / 2 = 4 950 subtract-abs-add triples per pixel at M = 100 (and a local-memory array) of the pairwise
// loop.  Loads are coalesced: consecutive lanes read consecutive pixels of one member plane.
constexpr int MET_MAXM = 128;
// softplus without a branch (a data-dependent branch between the member loads keeps the compiler from batching them)
__device__ __forceinline__ float softplus_sel(float x, float c) {
  const float sp = logf(expf(x) + 1.f) - c;
  return x > 20.f ? x : sp;
}

template <int NP, int MINB>
__global__ void __launch_bounds__(128, MINB) metrics_kernel(const float* __restrict__ preds, const float* __restrict__ hr,
                                                         const float* __restrict__ lrinterp, const float* __restrict__ std_hr,
                                                         int transform, int M, int C, int HW, float* __restrict__ part) {
  __shared__ float red[4];
  const int t = blockIdx.z, c = blockIdx.y;
  float a_crps = 0.f, a_mae = 0.f;
  const float inv_m = 1.f / (float)M;
  const float INF = __int_as_float(0x7f800000);
  const int64_t mstride = (int64_t)C * HW;                 // floats between two members of one (field, variable, pixel)
  for (int p = blockIdx.x * 128 + threadIdx.x; p < HW; p += gridDim.x * 128) {
    float x[NP];
    const float y = hr[((int64_t)t * C + c) * HW + p];
    const float* pc = preds + ((int64_t)t * M * C + c) * HW + p;
    // phase 1: the M raw loads, nothing else -- independent, so they are all in flight together (with the transform
    // interleaved, every load waited for the previous member's branchy arithmetic: 6.3 ms per 64 x 100 members
    // instead of < 1 ms, profiles/r02_ensemble_pass_cupti.txt)
#pragma unroll
    for (int j = 0; j < NP; ++j) x[j] = j < M ? __ldg(pc + j * mstride) : INF;   // +inf sorts behind the real members
    if (transform) {
      // residual_to_hr (src/climex_utils.py:277-285) + invert_transfo_3vars (results.ipynb cell 2)
      const float li = lrinterp[((int64_t)t * C + c) * HW + p], sc = std_hr[c] + 1e-10f;
      if (c == 0) {
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (j < M) x[j] = softplus_sel(fmaf(x[j], sc, li), 1e-7f) * 24.f * 60.f * 60.f;   // kgm2sTommday's own order
      } else if (c == 1) {
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (j < M) x[j] = fmaf(x[j], sc, li) - 273.15f;
      } else {
        const float li1 = lrinterp[((int64_t)t * C + 1) * HW + p], s1c = std_hr[1] + 1e-10f;
        const float* p1 = preds + ((int64_t)t * M * C + 1) * HW + p;
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (j < M) x[j] = softplus_sel(fmaf(x[j], sc, li), 1e-7f);
#pragma unroll
        for (int j = 0; j < NP; ++j)                      // second batch of independent loads: the tasmin member
          if (j < M) x[j] = (x[j] + fmaf(__ldg(p1 + j * mstride), s1c, li1)) - 273.15f;
      }
    }
    float mean = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (j < M) { mean += x[j]; s1 += fabsf(x[j] - y); }
    // bitonic sorting network, ascending
#pragma unroll
    for (int k = 2; k <= NP; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int l = i ^ j;
          if (l > i) {
            const float lo = fminf(x[i], x[l]), hi = fmaxf(x[i], x[l]);
            if ((i & k) == 0) { x[i] = lo; x[l] = hi; } else { x[i] = hi; x[l] = lo; }
          }
        }
      }
    }
    float s2 = 0.f;                                    // sum_{j<k} |x_j - x_k|
#pragma unroll
    for (int i = 0; i < NP; ++i)
      if (i < M) s2 = fmaf((float)(2 * i + 1 - M), x[i], s2);
    a_crps += s1 * inv_m - s2 * inv_m * inv_m;
    a_mae += fabsf(mean * inv_m - y);
  }
  float tot = block_sum<128>(a_crps, red);
  const int64_t o = (((int64_t)t * C + c) * gridDim.x + blockIdx.x) * 2;
  if (threadIdx.x == 0) part[o] = tot;
  tot = block_sum<128>(a_mae, red);
  if (threadIdx.x == 0) part[o + 1] = tot;
}
__global__ void metrics_final_kernel(const float* __restrict__ part, int nblk, int HW, int TC, float* __restrict__ crps,
                                     float* __restrict__ mae) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TC) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblk; ++k) { a += (double)part[((int64_t)i * nblk + k) * 2]; b += (double)part[((int64_t)i * nblk + k) * 2 + 1]; }
  crps[i] = (float)(a / HW); mae[i] = (float)(b / HW);
}

// ------------------------------------------------------------------ AdamW over a device table
__global__ void __launch_bounds__(NT) adamw_kernel(const pub_adamw_entry* __restrict__ tab, float lr, float b1, float b2,
                                                   float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  const pub_adamw_entry e = tab[blockIdx.y];
  const float step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * NT) {
    const float g = e.g[i] * gscale;
    float p = e.p[i] * (1.f - lr * wd);
    const float m = b1 * e.m[i] + (1.f - b1) * g;
    const float v = b2 * e.v[i] + (1.f - b2) * g * g;
    p -= step * m / (sqrtf(v) / bc2_sqrt + eps);
    e.p[i] = p; e.m[i] = m; e.v[i] = v;
  }
}

// the same update with the step count read from device memory (a captured graph replays fixed arguments)
__global__ void __launch_bounds__(NT) adamw_dev_kernel(const pub_adamw_entry* __restrict__ tab, float lr, float b1, float b2,
                                                       float eps, float wd, const int* __restrict__ step_dev, float gscale) {
  const float st = (float)__ldg(step_dev);
  const float bc1 = 1.f - powf(b1, st), bc2_sqrt = sqrtf(1.f - powf(b2, st));
  const pub_adamw_entry e = tab[blockIdx.y];
  const float step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * NT) {
    const float g = e.g[i] * gscale;
    float p = e.p[i] * (1.f - lr * wd);
    const float m = b1 * e.m[i] + (1.f - b1) * g;
    const float v = b2 * e.v[i] + (1.f - b2) * g * g;
    p -= step * m / (sqrtf(v) / bc2_sqrt + eps);
    e.p[i] = p; e.m[i] = m; e.v[i] = v;
  }
}
// counters[0]: optimizer step (+1), counters[1]: random salt (next value of a 32-bit mixer sequence)
__global__ void advance_counters_kernel(int* __restrict__ counters) {
  if (threadIdx.x == 0) {
    counters[0] += 1;
    counters[1] = (int)mix32((uint32_t)counters[1] * 0x9E3779B1u + 0x7F4A7C15u);
  }
}

inline int grid_for(int64_t n) {
  int64_t g = (n + NT - 1) / NT;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(g < cap ? g : cap);
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

int pub_rsample_forward(const float* mu, const float* sigma, const float* eps_in, uint64_t seed, uint64_t offset, int M,
                        int B, int L, float* z, float* eps_out, pub_stream_t s) {
  PUB_REQUIRE(mu && sigma && z, "pub_rsample_forward: null argument");
  const int64_t n = (int64_t)M * B * L;
  rsample_fwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)s>>>(mu, sigma, eps_in, seed, offset, M, B * L, z, eps_out, g_seed_salt);
  PUB_LAUNCH_CHECK();
  return 0;
}
int pub_rsample_backward(const float* dz, const float* eps, int M, int B, int L, float* dmu, float* dsigma, pub_stream_t s) {
  rsample_bwd_kernel<<<cdiv(B * L, 256), 256, 0, (cudaStream_t)s>>>(dz, eps, M, B * L, dmu, dsigma);
  PUB_LAUNCH_CHECK();
  return 0;
}
int pub_kl_normal_forward(const float* mq, const float* sq, const float* mp, const float* sp, int B, int L, float* kl,
                          pub_stream_t s) {
  kl_fwd_kernel<<<cdiv(B, 4), 128, 0, (cudaStream_t)s>>>(mq, sq, mp, sp, B, L, kl);
  PUB_LAUNCH_CHECK();
  return 0;
}
int pub_kl_normal_backward(const float* dkl, const float* mq, const float* sq, const float* mp, const float* sp, int B,
                           int L, float* dmq, float* dsq, float* dmp, float* dsp, pub_stream_t s) {
  kl_bwd_kernel<<<cdiv(B * L, 256), 256, 0, (cudaStream_t)s>>>(dkl, mq, sq, mp, sp, B, L, dmq, dsq, dmp, dsp);
  PUB_LAUNCH_CHECK();
  return 0;
}

size_t pub_loss_workspace(int B, int C, int HW) { return (size_t)(num_sms() * 8 * (C + 1) + 64) * sizeof(float); }

int pub_ensemble_loss(const float* ens, const float* target, int B, int M, int C, int HW, int kind, float alpha,
                      float* loss, float* dens, void* ws, size_t ws_bytes, pub_stream_t s) {
  PUB_REQUIRE(ens && target && loss && ws, "pub_ensemble_loss: null argument");
  PUB_REQUIRE(M >= 1 && M <= 64, "pub_ensemble_loss: M must be in [1, 64] (got %d)", M);
  PUB_REQUIRE(kind == PUB_LOSS_CRPS || M >= 2, "afCRPS needs M >= 2");
  PUB_REQUIRE(ws_bytes >= pub_loss_workspace(B, C, HW), "pub_ensemble_loss: workspace too small");
  const int64_t n = (int64_t)B * C * HW;
  float c_pair;
  if (kind == PUB_LOSS_AFCRPS) c_pair = (1.f - (1.f - alpha) / M) / (2.f * M * (M - 1));
  else c_pair = 1.f / (2.f * M * M);
  const int grid = grid_for(n);
  float* part = (float*)ws;
  const float inv_n = 1.f / (float)n;
  cudaStream_t st = (cudaStream_t)s;
  if (M <= 8) ens_loss_kernel<8><<<grid, NT, 0, st>>>(ens, target, B, M, C, HW, c_pair, inv_n, part, dens);
  else if (M <= 16) ens_loss_kernel<16><<<grid, NT, 0, st>>>(ens, target, B, M, C, HW, c_pair, inv_n, part, dens);
  else if (M <= 32) ens_loss_kernel<32><<<grid, NT, 0, st>>>(ens, target, B, M, C, HW, c_pair, inv_n, part, dens);
  else ens_loss_kernel<64><<<grid, NT, 0, st>>>(ens, target, B, M, C, HW, c_pair, inv_n, part, dens);
  PUB_LAUNCH_CHECK();
  final_sum_kernel<<<1, 32, 0, st>>>(part, grid, 1.0f / (float)n, loss);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_l1_loss(const float* out, const float* target, int B, int C, int HW, float* loss, float* dout, void* ws,
                size_t ws_bytes, pub_stream_t s) {
  PUB_REQUIRE(out && target && loss && ws, "pub_l1_loss: null argument");
  PUB_REQUIRE(C >= 1 && C <= 32, "pub_l1_loss: C must be <= 32");
  PUB_REQUIRE(ws_bytes >= pub_loss_workspace(B, C, HW), "pub_l1_loss: workspace too small");
  const int64_t n = (int64_t)B * HW;
  const int gx = grid_for(n);
  dim3 grid(gx, C);
  l1_kernel<<<grid, NT, 0, (cudaStream_t)s>>>(out, target, B, C, HW, 1.f / (float)(n * C), (float*)ws, dout);
  PUB_LAUNCH_CHECK();
  l1_final_kernel<<<1, 32 * C, 0, (cudaStream_t)s>>>((const float*)ws, gx, C, (double)n, loss);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_scale_by_device_scalar(float* y, const float* scale, int64_t n, pub_stream_t s) {
  scale_kernel<<<grid_for(n), NT, 0, (cudaStream_t)s>>>(y, scale, n);
  PUB_LAUNCH_CHECK();
  return 0;
}

size_t pub_ensemble_metrics_workspace(int T, int C, int HW) {
  const int nblk = cdiv(HW, 128) < 32 ? cdiv(HW, 128) : 32;
  return (size_t)T * C * nblk * 2 * sizeof(float) + 256;
}

int pub_ensemble_metrics(const float* preds, const float* hr, const float* lrinterp, const float* std_hr, int transform,
                         int T, int M, int C, int HW, float* crps_tc, float* mae_tc, void* ws, size_t ws_bytes,
                         pub_stream_t s) {
  PUB_REQUIRE(preds && hr && crps_tc && mae_tc, "pub_ensemble_metrics: null argument");
  PUB_REQUIRE(M >= 1 && M <= MET_MAXM, "pub_ensemble_metrics: M must be in [1, %d]", MET_MAXM);
  PUB_REQUIRE(!transform || (C == 3 && lrinterp && std_hr), "pub_ensemble_metrics: transform needs C == 3, lrinterp, std_hr");
  const int nblk = cdiv(HW, 128) < 32 ? cdiv(HW, 128) : 32;
  PUB_REQUIRE(ws && ws_bytes >= pub_ensemble_metrics_workspace(T, C, HW), "pub_ensemble_metrics: workspace too small");
  float* part = (float*)ws;
  dim3 grid(nblk, C, T);
  cudaStream_t st = (cudaStream_t)s;
  if (M <= 8) metrics_kernel<8, 3><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  else if (M <= 16) metrics_kernel<16, 3><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  else if (M <= 32) metrics_kernel<32, 3><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  else if (M <= 64) metrics_kernel<64, 3><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  else if (g_opt_metrics_occ >= 4) metrics_kernel<128, 4><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  else metrics_kernel<128, 3><<<grid, 128, 0, st>>>(preds, hr, lrinterp, std_hr, transform, M, C, HW, part);
  PUB_LAUNCH_CHECK();
  metrics_final_kernel<<<cdiv(T * C, 128), 128, 0, (cudaStream_t)s>>>(part, nblk, HW, T * C, crps_tc, mae_tc);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_adamw_step(const pub_adamw_entry* table, int n_tensors, int64_t max_numel, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, pub_stream_t s) {
  PUB_REQUIRE(table && n_tensors > 0 && step >= 1, "pub_adamw_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  int gx = cdiv(max_numel, (int64_t)NT * 4);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, n_tensors);
  adamw_kernel<<<grid, NT, 0, (cudaStream_t)s>>>(table, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_adamw_step_dev(const pub_adamw_entry* table, int n_tensors, int64_t max_numel, float lr, float beta1, float beta2,
                       float eps, float weight_decay, const int* step_dev, float grad_scale, pub_stream_t s) {
  PUB_REQUIRE(table && n_tensors > 0 && step_dev, "pub_adamw_step_dev: bad arguments");
  int gx = cdiv(max_numel, (int64_t)NT * 4);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, n_tensors);
  adamw_dev_kernel<<<grid, NT, 0, (cudaStream_t)s>>>(table, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_advance_counters(int* counters, pub_stream_t s) {
  PUB_REQUIRE(counters, "pub_advance_counters: null argument");
  advance_counters_kernel<<<1, 32, 0, (cudaStream_t)s>>>(counters);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
