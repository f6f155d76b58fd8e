// Fused Fcomb (src/prob_unet.py:87-138): broadcast z over H x W, concat with the U-Net features and run
// the per-pixel MLP  (F+L) -> F -> F -> C  for M latent samples in ONE pass over the features.
//
// Algebra used: W0 [f ; z] = W0f f + (W0z z + b0).  The feature half (W0f f) is computed once per
// pixel and reused by all M members; the latent half is a per-(member, sample) bias computed by a
// tiny pre-kernel.  Nothing of shape [B, F+L, H, W] is ever materialised.
//
// Backward recomputes the hidden activations, back-propagates per pixel and accumulates the weight
// gradients as smem-tiled outer products; per-CTA partials are reduced in a fixed order.
#include "common.cuh"
#include "kernels.h"

namespace pub {
namespace {

constexpr int F = 32;    // unet_output_channels (num_filters[0])
constexpr int CO = 3;    // num_classes
constexpr int NT = 128;  // one pixel per thread

struct FcombDev {
  const void* feat; int feat_nchw; int64_t sb, sc, sh, sw;
  const float* zb;  // [M][B][F]  = b0 + W0z z
  const float *w0, *w1, *b1, *w2, *b2;
  int B, H, W, L, M;
};

// zb[m][b][j] = b0[j] + sum_l w0[j][F+l] z[m][b][l]
__global__ void fcomb_zbias_kernel(const float* __restrict__ z, const float* __restrict__ w0, const float* __restrict__ b0,
                                   int MB, int L, float* __restrict__ zb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MB * F) return;
  const int j = i % F, mb = i / F;
  float s = b0[j];
  for (int l = 0; l < L; ++l) s = fmaf(w0[j * (F + L) + F + l], z[(int64_t)mb * L + l], s);
  zb[i] = s;
}

template <typename T>
__device__ __forceinline__ void load_feat(const FcombDev& a, int b, int pix, float (&f)[F]) {
  if (a.feat_nchw) {
    const float* p = (const float*)a.feat + b * a.sb + (pix / a.W) * a.sh + (pix % a.W) * a.sw;
#pragma unroll
    for (int i = 0; i < F; ++i) f[i] = p[i * a.sc];
  } else {
    const T* p = (const T*)a.feat + ((int64_t)b * a.H * a.W + pix) * F;
#pragma unroll
    for (int g = 0; g < F / 8; ++g) {
      float v[8];
      Vec8<T>::load(p + g * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[g * 8 + i] = v[i];
    }
  }
}

struct FcombSmem {
  float w0f[F][F];  // [j][i]
  float w1[F][F];   // [k][j]
  float w2[CO][F];
  float b1[F];
  float b2[4];
};

__device__ __forceinline__ void load_weights(FcombSmem& s, const FcombDev& a) {
  for (int i = threadIdx.x; i < F * F; i += NT) {
    s.w0f[i / F][i % F] = a.w0[(i / F) * (F + a.L) + i % F];
    s.w1[i / F][i % F] = a.w1[i];
  }
  for (int i = threadIdx.x; i < CO * F; i += NT) s.w2[i / F][i % F] = a.w2[i];
  if (threadIdx.x < F) s.b1[threadIdx.x] = a.b1[threadIdx.x];
  if (threadIdx.x < CO) s.b2[threadIdx.x] = a.b2[threadIdx.x];
}

template <typename T>
__global__ void __launch_bounds__(NT, 4) fcomb_fwd_kernel(FcombDev a, float* __restrict__ out) {
  __shared__ FcombSmem s;
  extern __shared__ float zbs[];  // [M][F] for this sample
  const int b = blockIdx.y, HW = a.H * a.W;
  load_weights(s, a);
  for (int i = threadIdx.x; i < a.M * F; i += NT) zbs[i] = a.zb[((int64_t)(i / F) * a.B + b) * F + i % F];
  __syncthreads();
  for (int pix = blockIdx.x * NT + threadIdx.x; pix < HW; pix += gridDim.x * NT) {
    float f[F], base[F];
    load_feat<T>(a, b, pix, f);
#pragma unroll
    for (int j = 0; j < F; ++j) {
      asm volatile("" ::: "memory");  // keep the weight loads of row j next to their FMAs (no 1024-value hoist)
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < F; i += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&s.w0f[j][i]);
        acc = fmaf(w.x, f[i], acc); acc = fmaf(w.y, f[i + 1], acc); acc = fmaf(w.z, f[i + 2], acc); acc = fmaf(w.w, f[i + 3], acc);
      }
      base[j] = acc;
    }
    for (int m = 0; m < a.M; ++m) {
      float h1[F];
#pragma unroll
      for (int j = 0; j < F; ++j) h1[j] = fmaxf(base[j] + zbs[m * F + j], 0.f);
      float o0 = s.b2[0], o1 = s.b2[1], o2 = s.b2[2];
      // k stays a rolled loop: fully unrolled, the compiler hoists all 1024 (member-invariant) weight loads out
      // of the member loop and spills them (measured: 30 KB of spills, 32 registers)
#pragma unroll 2
      for (int k = 0; k < F; ++k) {
        asm volatile("" ::: "memory");
        float acc = s.b1[k];
#pragma unroll
        for (int j = 0; j < F; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&s.w1[k][j]);
          acc = fmaf(w.x, h1[j], acc); acc = fmaf(w.y, h1[j + 1], acc); acc = fmaf(w.z, h1[j + 2], acc); acc = fmaf(w.w, h1[j + 3], acc);
        }
        acc = fmaxf(acc, 0.f);
        o0 = fmaf(s.w2[0][k], acc, o0); o1 = fmaf(s.w2[1][k], acc, o1); o2 = fmaf(s.w2[2][k], acc, o2);
      }
      float* op = out + (((int64_t)b * a.M + m) * CO) * HW + pix;
      op[0] = o0; op[(int64_t)HW] = o1; op[2 * (int64_t)HW] = o2;
    }
  }
}

// ---- backward
// per-CTA partial layout (floats): dW1[F*F] | dW0f[F*F] | dW2[CO*F] | db1[F] | db2[4] | S[M][F]
__host__ __device__ inline int part_stride(int M) { return 2 * F * F + CO * F + F + 4 + M * F; }

constexpr int TS = F + 4;  // smem tile row stride (floats), keeps rows 16-byte aligned

template <typename T>
__global__ void __launch_bounds__(NT) fcomb_bwd_kernel(FcombDev a, const float* __restrict__ dout, void* __restrict__ dfeat,
                                                       float* __restrict__ part) {
  __shared__ FcombSmem s;
  __shared__ __align__(16) float tA[NT][TS];  // "row" operand  (dp2 / dbase)
  __shared__ __align__(16) float tB[NT][TS];  // "col" operand  (h1 / f / h2)
  extern __shared__ float dyn[];              // zbs[M][F] | Sacc[M][F]
  float* zbs = dyn;
  float* Sacc = dyn + a.M * F;
  const int b = blockIdx.y, HW = a.H * a.W, t = threadIdx.x;
  load_weights(s, a);
  for (int i = t; i < a.M * F; i += NT) { zbs[i] = a.zb[((int64_t)(i / F) * a.B + b) * F + i % F]; Sacc[i] = 0.f; }
  __syncthreads();
  // this thread's slice of the 32x32 outer-product accumulators: row r0, columns c0..c0+7
  const int r0 = t >> 2, c0 = (t & 3) * 8;
  float aW1[8], aW0[8], aW2 = 0.f, ab1 = 0.f, ab2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) aW1[i] = aW0[i] = 0.f;
  const int ntile = (HW + NT - 1) / NT;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int pix = tile * NT + t;
    const bool valid = pix < HW;
    float f[F], base[F], dbase[F];
    if (valid) load_feat<T>(a, b, pix, f);
    else {
#pragma unroll
      for (int i = 0; i < F; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < F; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < F; i += 4) {
        const float4 w = *reinterpret_cast<const float4*>(&s.w0f[j][i]);
        acc = fmaf(w.x, f[i], acc); acc = fmaf(w.y, f[i + 1], acc); acc = fmaf(w.z, f[i + 2], acc); acc = fmaf(w.w, f[i + 3], acc);
      }
      base[j] = acc; dbase[j] = 0.f;
    }
    for (int m = 0; m < a.M; ++m) {
      float h1[F], h2[F], dh1[F];
#pragma unroll
      for (int j = 0; j < F; ++j) { h1[j] = fmaxf(base[j] + zbs[m * F + j], 0.f); dh1[j] = 0.f; }
      float d0 = 0.f, d1 = 0.f, d2 = 0.f;
      if (valid) {
        const float* dp = dout + (((int64_t)b * a.M + m) * CO) * HW + pix;
        d0 = dp[0]; d1 = dp[(int64_t)HW]; d2 = dp[2 * (int64_t)HW];
      }
#pragma unroll
      for (int k = 0; k < F; ++k) {
        float acc = s.b1[k];
#pragma unroll
        for (int j = 0; j < F; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&s.w1[k][j]);
          acc = fmaf(w.x, h1[j], acc); acc = fmaf(w.y, h1[j + 1], acc); acc = fmaf(w.z, h1[j + 2], acc); acc = fmaf(w.w, h1[j + 3], acc);
        }
        h2[k] = fmaxf(acc, 0.f);
        // dp2[k] = relu'(pre2) * (W2^T dout)[k]
        const float g = acc > 0.f ? (s.w2[0][k] * d0 + s.w2[1][k] * d1 + s.w2[2][k] * d2) : 0.f;
#pragma unroll
        for (int j = 0; j < F; j += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&s.w1[k][j]);
          dh1[j] = fmaf(w.x, g, dh1[j]); dh1[j + 1] = fmaf(w.y, g, dh1[j + 1]);
          dh1[j + 2] = fmaf(w.z, g, dh1[j + 2]); dh1[j + 3] = fmaf(w.w, g, dh1[j + 3]);
        }
        tA[t][k] = g;  // dp2
      }
      // ---- dW1[k][j] += dp2[k] * h1[j]   (tA = dp2, tB = h1)
#pragma unroll
      for (int j = 0; j < F; ++j) tB[t][j] = h1[j];
      __syncthreads();
      for (int px = 0; px < NT; ++px) {
        const float av = tA[px][r0];
        const float4 b0 = *reinterpret_cast<const float4*>(&tB[px][c0]);
        const float4 b1 = *reinterpret_cast<const float4*>(&tB[px][c0 + 4]);
        aW1[0] = fmaf(av, b0.x, aW1[0]); aW1[1] = fmaf(av, b0.y, aW1[1]); aW1[2] = fmaf(av, b0.z, aW1[2]); aW1[3] = fmaf(av, b0.w, aW1[3]);
        aW1[4] = fmaf(av, b1.x, aW1[4]); aW1[5] = fmaf(av, b1.y, aW1[5]); aW1[6] = fmaf(av, b1.z, aW1[6]); aW1[7] = fmaf(av, b1.w, aW1[7]);
      }
      if (t < F) {  // db1[k] += sum_px dp2[px][k]
        float sb = 0.f;
        for (int px = 0; px < NT; ++px) sb += tA[px][t];
        ab1 += sb;
      }
      __syncthreads();
      // ---- dW2[c][k] += dout[c] * h2[k], db2[c] += dout[c]   (tA cols 0..2 = dout, tB = h2)
      tA[t][0] = d0; tA[t][1] = d1; tA[t][2] = d2;
#pragma unroll
      for (int k = 0; k < F; ++k) tB[t][k] = h2[k];
      // dp1 = relu'(pre1) * dh1  -> dbase, and the per-sample sums S[m][j]
#pragma unroll
      for (int j = 0; j < F; ++j) { const float g = h1[j] > 0.f ? dh1[j] : 0.f; dbase[j] += g; dh1[j] = g; }
      __syncthreads();
      if (t < CO * F) {
        const int c = t / F, k = t % F;
        float sw = 0.f;
        for (int px = 0; px < NT; ++px) sw = fmaf(tA[px][c], tB[px][k], sw);
        aW2 += sw;
      } else if (t < CO * F + CO) {
        const int c = t - CO * F;
        float sb = 0.f;
        for (int px = 0; px < NT; ++px) sb += tA[px][c];
        ab2 += sb;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < F; ++j) tA[t][j] = dh1[j];  // dp1
      __syncthreads();
      if (t < F) {
        float sj = 0.f;
        for (int px = 0; px < NT; ++px) sj += tA[px][t];
        Sacc[m * F + t] += sj;
      }
      __syncthreads();
    }
    // ---- dW0f[j][i] += dbase[j] * f[i]; dfeat[i] = sum_j W0f[j][i] dbase[j]
#pragma unroll
    for (int j = 0; j < F; ++j) { tA[t][j] = dbase[j]; tB[t][j] = f[j]; }
    __syncthreads();
    for (int px = 0; px < NT; ++px) {
      const float av = tA[px][r0];
      const float4 b0 = *reinterpret_cast<const float4*>(&tB[px][c0]);
      const float4 b1 = *reinterpret_cast<const float4*>(&tB[px][c0 + 4]);
      aW0[0] = fmaf(av, b0.x, aW0[0]); aW0[1] = fmaf(av, b0.y, aW0[1]); aW0[2] = fmaf(av, b0.z, aW0[2]); aW0[3] = fmaf(av, b0.w, aW0[3]);
      aW0[4] = fmaf(av, b1.x, aW0[4]); aW0[5] = fmaf(av, b1.y, aW0[5]); aW0[6] = fmaf(av, b1.z, aW0[6]); aW0[7] = fmaf(av, b1.w, aW0[7]);
    }
    __syncthreads();
    if (dfeat && valid) {
      float df[F];
#pragma unroll
      for (int i = 0; i < F; ++i) df[i] = 0.f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
#pragma unroll
        for (int i = 0; i < F; i += 4) {
          const float4 w = *reinterpret_cast<const float4*>(&s.w0f[j][i]);
          df[i] = fmaf(w.x, dbase[j], df[i]); df[i + 1] = fmaf(w.y, dbase[j], df[i + 1]);
          df[i + 2] = fmaf(w.z, dbase[j], df[i + 2]); df[i + 3] = fmaf(w.w, dbase[j], df[i + 3]);
        }
      }
      if (a.feat_nchw) {
        float* p = (float*)dfeat + ((int64_t)b * F) * HW + pix;
#pragma unroll
        for (int i = 0; i < F; ++i) p[(int64_t)i * HW] = df[i];
      } else {
        T* p = (T*)dfeat + ((int64_t)b * HW + pix) * F;
#pragma unroll
        for (int g = 0; g < F / 8; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = df[g * 8 + i];
          Vec8<T>::store(p + g * 8, v);
        }
      }
    }
  }
  // ---- write this CTA's partials
  __syncthreads();
  float* o = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * part_stride(a.M);
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[r0 * F + c0 + i] = aW1[i]; o[F * F + r0 * F + c0 + i] = aW0[i]; }
  if (t < CO * F) o[2 * F * F + t] = aW2;
  else if (t < CO * F + CO) o[2 * F * F + CO * F + F + (t - CO * F)] = ab2;
  if (t < F) o[2 * F * F + CO * F + t] = ab1;
  for (int i = t; i < a.M * F; i += NT) o[2 * F * F + CO * F + F + 4 + i] = Sacc[i];
}

// final reduction over CTAs (fixed order) + the latent-half gradients
__global__ void fcomb_bwd_final_kernel(const float* __restrict__ part, int nx, int B, int M, int L,
                                       const float* __restrict__ z, const float* __restrict__ w0,
                                       float* __restrict__ dz, float* __restrict__ dw0, float* __restrict__ db0,
                                       float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                       float* __restrict__ db2, float* __restrict__ Stot /* [M][B][F] scratch */) {
  const int ps = part_stride(M);
  const int nfixed = 2 * F * F + CO * F + F + CO;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // phase 1 (this kernel is launched twice: phase via L sign trick is avoided -> two kernels below)
  for (int i = tid; i < nfixed; i += nth) {
    int src;
    if (i < 2 * F * F + CO * F + F) src = i;
    else src = 2 * F * F + CO * F + F + (i - (2 * F * F + CO * F + F));
    double sum = 0.0;
    for (int c = 0; c < nx * B; ++c) sum += (double)part[(int64_t)c * ps + src];
    const float v = (float)sum;
    if (i < F * F) dw1[i] = v;
    else if (i < 2 * F * F) { const int k = i - F * F; dw0[(k / F) * (F + L) + k % F] = v; }
    else if (i < 2 * F * F + CO * F) dw2[i - 2 * F * F] = v;
    else if (i < 2 * F * F + CO * F + F) db1[i - 2 * F * F - CO * F] = v;
    else db2[i - 2 * F * F - CO * F - F] = v;
  }
  for (int i = tid; i < M * B * F; i += nth) {
    const int j = i % F, b = (i / F) % B, m = i / (F * B);
    double sum = 0.0;
    for (int c = 0; c < nx; ++c) sum += (double)part[((int64_t)b * nx + c) * ps + 2 * F * F + CO * F + F + 4 + m * F + j];
    Stot[i] = (float)sum;
  }
}
__global__ void fcomb_bwd_latent_kernel(const float* __restrict__ Stot, int B, int M, int L, const float* __restrict__ z,
                                        const float* __restrict__ w0, float* __restrict__ dz, float* __restrict__ dw0,
                                        float* __restrict__ db0) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int MB = M * B;
  for (int i = tid; i < F; i += nth) {  // db0[j]
    double s = 0.0;
    for (int mb = 0; mb < MB; ++mb) s += (double)Stot[mb * F + i];
    db0[i] = (float)s;
  }
  for (int i = tid; i < F * L; i += nth) {  // dW0z[j][l] = sum_mb S[mb][j] z[mb][l]
    const int j = i / L, l = i % L;
    double s = 0.0;
    for (int mb = 0; mb < MB; ++mb) s += (double)Stot[mb * F + j] * (double)z[(int64_t)mb * L + l];
    dw0[j * (F + L) + F + l] = (float)s;
  }
  if (dz) {
    for (int i = tid; i < MB * L; i += nth) {  // dz[mb][l] = sum_j W0z[j][l] S[mb][j]
      const int mb = i / L, l = i % L;
      float s = 0.f;
      for (int j = 0; j < F; ++j) s = fmaf(w0[j * (F + L) + F + l], Stot[mb * F + j], s);
      dz[i] = s;
    }
  }
}

int bwd_grid_x(int B, int HW) {
  int gx = cdiv(4 * num_sms(), B);
  const int ntile = cdiv(HW, NT);
  if (gx > ntile) gx = ntile;
  if (gx < 1) gx = 1;
  return gx;
}

int validate(const pub_fcomb_args* a) {
  PUB_REQUIRE(a && a->feat && a->z && a->w0 && a->b0 && a->w1 && a->b1 && a->w2 && a->b2, "pub_fcomb: null argument");
  PUB_REQUIRE(a->F == F && a->C == CO, "pub_fcomb: only unet_output_channels=32, num_classes=3 are built (got F=%d C=%d)", a->F, a->C);
  PUB_REQUIRE(a->M >= 1 && a->M <= 256 && a->L >= 1, "pub_fcomb: bad M/L");
  return 0;
}

FcombDev make_dev(const pub_fcomb_args* a, const float* zb) {
  FcombDev d{};
  d.feat = a->feat; d.feat_nchw = a->feat_nchw;
  d.sb = a->stride[0]; d.sc = a->stride[1]; d.sh = a->stride[2]; d.sw = a->stride[3];
  d.zb = zb; d.w0 = a->w0; d.w1 = a->w1; d.b1 = a->b1; d.w2 = a->w2; d.b2 = a->b2;
  d.B = a->B; d.H = a->H; d.W = a->W; d.L = a->L; d.M = a->M;
  return d;
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

size_t pub_fcomb_backward_workspace(const pub_fcomb_args* a) {
  const int gx = bwd_grid_x(a->B, a->H * a->W);
  return align_up((size_t)a->M * a->B * F * 4, 256) * 2 + align_up((size_t)gx * a->B * part_stride(a->M) * 4, 256);
}

size_t pub_fcomb_forward_workspace(const pub_fcomb_args* a) { return align_up((size_t)a->M * a->B * F * 4, 256); }

int pub_fcomb_forward(const pub_fcomb_args* a, void* ws, size_t ws_bytes, pub_stream_t s) {
  PUB_TRY(validate(a));
  PUB_REQUIRE(a->out && ws, "pub_fcomb_forward: null out / workspace");
  PUB_REQUIRE(ws_bytes >= pub_fcomb_forward_workspace(a), "pub_fcomb_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)s;
  float* zb = (float*)ws;  // caller-owned: stream-ordered cudaMallocAsync/FreeAsync pairs cost 2-80 ms after every sync
  fcomb_zbias_kernel<<<cdiv(a->M * a->B * F, 256), 256, 0, st>>>(a->z, a->w0, a->b0, a->M * a->B, a->L, zb);
  PUB_LAUNCH_CHECK();
  const FcombDev d = make_dev(a, zb);
  const int HW = a->H * a->W;
  int gx = cdiv(HW, NT);
  dim3 grid(gx, a->B);
  const size_t dyn = (size_t)a->M * F * 4;
  if (!a->feat_nchw && a->dtype == PUB_BF16) fcomb_fwd_kernel<bf16><<<grid, NT, dyn, st>>>(d, a->out);
  else fcomb_fwd_kernel<float><<<grid, NT, dyn, st>>>(d, a->out);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_fcomb_backward(const pub_fcomb_args* a, const float* dout, void* dfeat, float* dz, float* dw0, float* db0,
                       float* dw1, float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, pub_stream_t s) {
  PUB_TRY(validate(a));
  PUB_REQUIRE(dout && dw0 && db0 && dw1 && db1 && dw2 && db2 && ws, "pub_fcomb_backward: null argument");
  PUB_REQUIRE(ws_bytes >= pub_fcomb_backward_workspace(a), "pub_fcomb_backward: workspace too small");
  PUB_REQUIRE(a->M <= 96, "pub_fcomb_backward: M <= 96");
  cudaStream_t st = (cudaStream_t)s;
  const size_t zbytes = align_up((size_t)a->M * a->B * F * 4, 256);
  float* zb = (float*)ws;
  float* Stot = (float*)((char*)ws + zbytes);
  float* part = (float*)((char*)ws + 2 * zbytes);
  fcomb_zbias_kernel<<<cdiv(a->M * a->B * F, 256), 256, 0, st>>>(a->z, a->w0, a->b0, a->M * a->B, a->L, zb);
  PUB_LAUNCH_CHECK();
  const FcombDev d = make_dev(a, zb);
  const int gx = bwd_grid_x(a->B, a->H * a->W);
  dim3 grid(gx, a->B);
  const size_t dyn = (size_t)a->M * F * 4 * 2;
  static bool attr = false;
  if (!attr) {  // static (tiles + weights, ~45 KB) + dynamic smem exceeds the 48 KB default limit for M > 8
    PUB_CUDA(cudaFuncSetAttribute(fcomb_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    PUB_CUDA(cudaFuncSetAttribute(fcomb_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  if (!a->feat_nchw && a->dtype == PUB_BF16) fcomb_bwd_kernel<bf16><<<grid, NT, dyn, st>>>(d, dout, dfeat, part);
  else fcomb_bwd_kernel<float><<<grid, NT, dyn, st>>>(d, dout, dfeat, part);
  PUB_LAUNCH_CHECK();
  fcomb_bwd_final_kernel<<<16, 256, 0, st>>>(part, gx, a->B, a->M, a->L, a->z, a->w0, dz, dw0, db0, dw1, db1, dw2, db2, Stot);
  PUB_LAUNCH_CHECK();
  fcomb_bwd_latent_kernel<<<8, 256, 0, st>>>(Stot, a->B, a->M, a->L, a->z, a->w0, dz, dw0, db0);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
