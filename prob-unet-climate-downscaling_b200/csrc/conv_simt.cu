// fp32-FMA implicit-GEMM convolution family (any shape, f32 or bf16 storage, f32 accumulate).
// This is the PARITY path (fp32 rel-err <= 1e-4 vs the reference) and the fallback for the
// shapes the tcgen05 kernels do not take (Cin = 3/6 first layers).  Replaces F.conv2d at
// src/networks.py:89 / src/prob_unet.py:41-46 and convolution_backward's weight gradient.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace pub {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
__device__ __forceinline__ float load2src(const T* x0, const T* x1, int c0, int cin, int ld0, int ld1,
                                          int64_t pix, int c) {
  if (c < c0) return to_f<T>(x0[pix * ld0 + c]);
  if (c < cin) return to_f<T>(x1[pix * ld1 + (c - c0)]);
  return 0.f;
}

template <typename T>
__global__ void __launch_bounds__(NT) conv_simt_kernel(ConvParams p) {
  pdl_enter();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* x0 = (const T*)p.x0;
  const T* x1 = (const T*)p.x1;
  const T* w = (const T*)p.w;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cin = p.c0 + p.c1;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // this thread's load coordinates
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int64_t lm = m0 + lr;
  int lb = 0, ly = 0, lx = 0;
  const bool lvalid = lm < M;
  if (lvalid) { lx = (int)(lm % p.W); ly = (int)((lm / p.W) % p.H); lb = (int)(lm / ((int64_t)p.W * p.H)); }
  const int ln = n0 + lr;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int taps = p.ks * p.ks, half = p.ks / 2;
  const int kchunks = (cin + BK - 1) / BK;
  for (int tap = 0; tap < taps; ++tap) {
    const int yy = ly + tap / p.ks - half, xx = lx + tap % p.ks - half;
    const bool inb = lvalid && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
    const int64_t sp = ((int64_t)lb * p.H + yy) * p.W + xx;
    for (int kc = 0; kc < kchunks; ++kc) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = kc * BK + lk + j;
        As[lk + j][lr] = inb ? load2src<T>(x0, x1, p.c0, cin, p.ld0, p.ld1, sp, c) : 0.f;
        Bs[lk + j][lr] = (ln < p.cout && c < cin) ? to_f<T>(w[((int64_t)tap * p.cout + ln) * cin + c]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  T* y = (T*)p.y;
  const T* res = (const T*)p.res;
  const T* mask = (const T*)p.mask;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.cout) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (res) v += to_f<T>(res[m * p.ld_res + n]);
      if (p.relu) v = fmaxf(v, 0.f);
      if (mask && !(to_f<T>(mask[m * p.ld_mask + n]) > 0.f)) v = 0.f;
      if (p.round_tf32) v = round_tf32_f(v);
      y[m * p.ldy + n] = from_f<T>(v);
    }
  }
}

// First `CIN` channels of the 8-channel padded input pixel (16 B of bf16 / 32 B of f32): loads only what is used.
template <typename T, int CIN> struct PixLoad;
template <int CIN> struct PixLoad<float, CIN> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[CIN]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float t[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int i = 0; i < (CIN < 4 ? CIN : 4); ++i) v[i] = t[i];
    if (CIN > 4) {
      const float4 b = *reinterpret_cast<const float4*>(p + 4);
      const float u[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 4; i < CIN; ++i) v[i] = u[i - 4];
    }
  }
};
template <int CIN> struct PixLoad<bf16, CIN> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[CIN]) {
    if (CIN <= 4) {
      const uint2 u = *reinterpret_cast<const uint2*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
      const float t[4] = {f0.x, f0.y, f1.x, f1.y};
#pragma unroll
      for (int i = 0; i < CIN; ++i) v[i] = t[i];
    } else {
      float t[8];
      Vec8<bf16>::load(p, t);
#pragma unroll
      for (int i = 0; i < CIN; ++i) v[i] = t[i];
    }
  }
};

// Forward conv of the Cin <= 8 first layers (3 / 6 climate variables -> 32 channels).  Thread = output pixel with
// its 32 output channels in registers; the [tap][ci][co] weights are smem broadcasts (one LDS.128 feeds 4 FMAs), the
// 9 input pixels are 8/16/32-byte loads coalesced across the warp.  27 (Cin = 3) FMA instructions per pixel and
// warp instead of 72 + 72 shared loads of the lane-per-channel version it replaces (1.1 ms -> HBM-bound output write).
template <typename T, int CIN>
__global__ void __launch_bounds__(128) conv_smallc_kernel(ConvParams p) {
  pdl_enter();
  __shared__ __align__(16) float ws[9 * CIN][32];  // [tap*CIN + ci][co]
  __shared__ float bs[32];
  const T* x0 = (const T*)p.x0;
  const T* w = (const T*)p.w;
  const int cin = p.c0, taps = p.ks * p.ks, half = p.ks / 2;
  const int n0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < 9 * CIN * 32; i += 128) {
    const int co = i & 31, tc = i >> 5, t = tc / CIN, c = tc % CIN;
    ws[tc][co] = (n0 + co < p.cout && t < taps && c < cin) ? to_f<T>(w[((int64_t)t * p.cout + n0 + co) * cin + c]) : 0.f;
  }
  if (threadIdx.x < 32) bs[threadIdx.x] = (p.bias && n0 + threadIdx.x < p.cout) ? p.bias[n0 + threadIdx.x] : 0.f;
  __syncthreads();
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t m = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (m >= M) return;
  const int x = (int)(m % p.W), y = (int)((m / p.W) % p.H), b = (int)(m / ((int64_t)p.W * p.H));
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = bs[j];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (t < taps) {
      const int yy = y + t / p.ks - half, xx = x + t % p.ks - half;
      const bool ok = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
      const int yc = min(max(yy, 0), p.H - 1), xc = min(max(xx, 0), p.W - 1);  // always a valid address
      float xv[CIN];
      PixLoad<T, CIN>::load(x0 + (((int64_t)b * p.H + yc) * p.W + xc) * p.ld0, xv);
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float xs = (ok && c < cin) ? xv[c] : 0.f;  // padding channels may hold anything
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[t * CIN + c][j]);
          acc[j] = fmaf(wv.x, xs, acc[j]); acc[j + 1] = fmaf(wv.y, xs, acc[j + 1]);
          acc[j + 2] = fmaf(wv.z, xs, acc[j + 2]); acc[j + 3] = fmaf(wv.w, xs, acc[j + 3]);
        }
      }
    }
  }
  const T* res = (const T*)p.res;
  const T* mask = (const T*)p.mask;
  T* yo = (T*)p.y + m * p.ldy + n0;
  const bool fast = n0 + 32 <= p.cout && p.ldy % 8 == 0 && (((uintptr_t)p.y) & 31) == 0 &&
                    (!res || (p.ld_res % 8 == 0 && (((uintptr_t)p.res) & 31) == 0)) &&
                    (!mask || (p.ld_mask % 8 == 0 && (((uintptr_t)p.mask) & 31) == 0));
  if (fast) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = acc[g * 8 + j];
      if (res) {
        float f[8];
        Vec8<T>::load(res + m * p.ld_res + n0 + g * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += f[j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (mask) {
        float f[8];
        Vec8<T>::load(mask + m * p.ld_mask + n0 + g * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f[j] > 0.f ? v[j] : 0.f;
      }
      if (p.round_tf32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = round_tf32_f(v[j]);
      }
      Vec8<T>::store(yo + g * 8, v);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n >= p.cout) continue;
      float v = acc[j];
      if (res) v += to_f<T>(res[m * p.ld_res + n]);
      if (p.relu) v = fmaxf(v, 0.f);
      if (mask && !(to_f<T>(mask[m * p.ld_mask + n]) > 0.f)) v = 0.f;
      if (p.round_tf32) v = round_tf32_f(v);
      yo[j] = from_f<T>(v);
    }
  }
}

// dW[tap][co][ci] partial over a pixel range:  sum_p dy[p][co] * x[p + tap][ci]
template <typename T>
__global__ void __launch_bounds__(NT) wgrad_simt_kernel(WgradParams p, float* __restrict__ part, int nsplit,
                                                        int pix_per_split) {
  pdl_enter();
  __shared__ float As[BK][BM + 4];  // [pixel][co]
  __shared__ float Bs[BK][BN + 4];  // [pixel][ci]
  const T* x0 = (const T*)p.x0;
  const T* x1 = (const T*)p.x1;
  const T* dy = (const T*)p.dy;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cin = p.c0 + p.c1;
  const int taps = p.ks * p.ks, half = p.ks / 2;
  const int tap = blockIdx.z / nsplit, split = blockIdx.z % nsplit;
  const int ddy = tap / p.ks - half, ddx = tap % p.ks - half;
  const int co0 = blockIdx.y * BM, ci0 = blockIdx.x * BN;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t pbeg = (int64_t)split * pix_per_split;
  const int64_t pend = min(M, pbeg + pix_per_split);
  const int lp = tid >> 4, lc = (tid & 15) * 4;  // load: pixel lp, channels lc..lc+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t pb = pbeg; pb < pend; pb += BK) {
    const int64_t m = pb + lp;
    const bool v = m < pend;
    int x = 0, y = 0, b = 0;
    if (v) { x = (int)(m % p.W); y = (int)((m / p.W) % p.H); b = (int)(m / ((int64_t)p.W * p.H)); }
    const int yy = y + ddy, xx = x + ddx;
    const bool inb = v && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
    const int64_t sp = ((int64_t)b * p.H + yy) * p.W + xx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + lc + j, ci = ci0 + lc + j;
      As[lp][lc + j] = (v && co < p.cout) ? to_f<T>(dy[m * p.ld_dy + co]) : 0.f;
      Bs[lp][lc + j] = inb ? load2src<T>(x0, x1, p.c0, cin, p.ld0, p.ld1, sp, ci) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= p.cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci >= cin) continue;
      part[(((int64_t)split * taps + tap) * p.cout + co) * cin + ci] = acc[i][j];
    }
  }
}

// sum the split partials in a fixed order -> OIHW f32 gradient (deterministic split-K).  One thread per
// element (coalesced across elements); the loads of consecutive splits are independent, so the unrolled loop
// keeps 8 of them in flight while the additions stay in split order.
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                           int nsplit, int taps, int cout, int cin, int accumulate) {
  pdl_enter();
  const int64_t n = (int64_t)taps * cout * cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  int k = 0;
  for (; k + 8 <= nsplit; k += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(part + (int64_t)(k + j) * n + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  for (; k < nsplit; ++k) s += __ldg(part + (int64_t)k * n + i);
  const int ci = (int)(i % cin), co = (int)((i / cin) % cout), tap = (int)(i / ((int64_t)cin * cout));
  const int64_t o = ((int64_t)co * cin + ci) * taps + tap;  // OIHW with (ky,kx) == tap
  dw[o] = accumulate ? dw[o] + s : s;
}

// split-K reduction of dW AND the final reduction of the bias-gradient column sums in one launch
// (blocks [0, nbw): the dW sums; blocks [nbw, ...): colsum_final_kernel's work, one warp per channel).
// KL = 1: a thread owns one dW element and walks all splits (8 loads in flight).  KL = 4 (many splits, few outputs:
// e.g. 148 splits of a 32 -> 32 layer, where 72 blocks of serial 148-term sums took ~12 us): the four warps of a block
// share 32 consecutive elements, warp w sums splits w, w+4, ..., and warp 0 adds the four partials in fixed order.
template <int KL>
__global__ void __launch_bounds__(128) wgrad_finish_kernel(const float* __restrict__ part, float* __restrict__ dw,
                                                           int nsplit, int taps, int cout, int cin, int nbw,
                                                           const float* __restrict__ bpart, int nchunk,
                                                           float* __restrict__ dbias, int accumulate) {
  pdl_enter();
  __shared__ float red[KL][32];
  if ((int)blockIdx.x < nbw) {
    const int64_t n = (int64_t)taps * cout * cin;
    const int kl = KL == 1 ? 0 : (int)(threadIdx.x >> 5);
    const int64_t i = KL == 1 ? (int64_t)blockIdx.x * blockDim.x + threadIdx.x
                              : (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    float s = 0.f;
    if (i < n) {
      int k = kl;
      for (; k + 7 * KL < nsplit; k += 8 * KL) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(part + (int64_t)(k + j * KL) * n + i);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
      }
      for (; k < nsplit; k += KL) s += __ldg(part + (int64_t)k * n + i);
    }
    if (KL > 1) {
      red[kl][threadIdx.x & 31] = s;
      __syncthreads();
      if (kl != 0) return;
      s = red[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < KL; ++w) s += red[w][threadIdx.x];
    }
    if (i >= n) return;
    const int ci = (int)(i % cin), co = (int)((i / cin) % cout), tap = (int)(i / ((int64_t)cin * cout));
    const int64_t o = ((int64_t)co * cin + ci) * taps + tap;  // OIHW with (ky,kx) == tap
    dw[o] = accumulate ? dw[o] + s : s;
  } else {
    const int c = ((int)blockIdx.x - nbw) * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= cout) return;
    double s = 0.0;
    for (int k = lane; k < nchunk; k += 32) s += (double)bpart[(int64_t)k * cout + c];
    s = warp_sum_d(s);
    if (lane == 0) dbias[c] = accumulate ? dbias[c] + (float)s : (float)s;
  }
}

// Weight gradient of the Cin <= 8 first layers (3 / 6 input variables): lane = output channel, the 9 x CIN
// accumulators live in registers.  A warp walks its pixel range one pixel per iteration; the operands of a pixel --
// the nine 8-channel padded input pixels of its 3x3 neighbourhood (16 / 32 B each) and the 32-channel dy row -- are
// fetched EIGHT pixels ahead by ONE warp-wide cp.async (lane l owns 16-byte chunk l of the slot: 9 or 18 x chunks,
// then 4 or 8 dy chunks) into a per-warp ring in shared memory and read back as broadcast LDS.128.  Out-of-image
// taps and channels >= Cout are zero-filled by cp.async (src-size 0), so the inner loop is 9 x CIN plain FMAs.
// The previous version (direct global loads, one or two pixels in flight per warp) sat on load latency at
// IPC ~0.17: 0.55 / 0.42 / 0.37 ms for the three first layers of a training step.  part[cta][tap][co][ci].
template <typename T, int CIN, int KS>
__global__ void __launch_bounds__(256, 2) wgrad_smallc_kernel(WgradParams p, float* __restrict__ part, int pix_per_cta) {
  pdl_enter();
  constexpr int ES = sizeof(T), EPC = 16 / ES;   // elements per 16-byte chunk
  constexpr int CPT = 8 / EPC;                   // chunks per padded input pixel: 1 (bf16) / 2 (f32)
  constexpr int NXC = 9 * CPT, NDC = 32 / EPC;   // x / dy chunks per pixel
  constexpr int SLOT = (NXC + NDC) * 16;         // 208 / 416 bytes
  constexpr int D = 8;                           // pixels in flight per warp (16 measured no faster)
  constexpr int TAPS = KS * KS;
  __shared__ __align__(16) uint8_t ring_raw[8 * D * SLOT];
  __shared__ float red[9 * CIN][32];
  const T* x0 = (const T*)p.x0;
  const T* dy = (const T*)p.dy;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cin = p.c0, taps = TAPS;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t pbeg = (int64_t)blockIdx.x * pix_per_cta, pend = min(M, pbeg + pix_per_cta);
  float acc[9][CIN];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < CIN; ++c) acc[t][c] = 0.f;
  // each warp owns a contiguous pixel range
  const int64_t per_warp = (pend - pbeg + 7) / 8;
  const int64_t wbeg = min(pend, pbeg + warp * per_warp), wend = min(pend, wbeg + per_warp);
  // This lane's chunk of every slot.  NHWC pixels are contiguous, so the source of the chunk for the next pixel is the
  // previous one plus a constant stride; only the border test needs the (x, y) of the pixel (warp-uniform counters).
  const bool is_x = lane < NXC, is_d = lane >= NXC && lane < NXC + NDC;
  const int tap = is_x ? lane / CPT : 0, hf = lane % CPT;
  const int ty = KS == 3 ? tap / 3 - 1 : 0, tx = KS == 3 ? tap % 3 - 1 : 0;   // KS = 1: tap 0 is the centre
  const bool x_live = is_x && tap < TAPS;
  const int d_el = blockIdx.y * 32 + (lane - NXC) * EPC;                       // first dy channel of the chunk
  const bool d_live = is_d && d_el < p.cout;                                   // cout % 8 == 0 (smallc_ok)
  const bool active = x_live || is_d;
  const bool nt = x_live && ty < 0, nb = x_live && ty > 0, nl = x_live && tx < 0, nr = x_live && tx > 0;
  const bool never = is_d && !d_live;                                          // always zero-filled
  const char* src = is_x ? (const char*)(x0 + wbeg * p.ld0 + ((ty * p.W + tx) * p.ld0 + hf * EPC))
                         : (const char*)(dy + wbeg * p.ld_dy + (d_live ? d_el : 0));
  const int64_t step = (int64_t)(is_x ? p.ld0 : p.ld_dy) * ES;
  const char* safe = is_x ? (const char*)x0 : (const char*)dy;                 // valid address for zero-size copies
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_raw) + (uint32_t)(warp * D * SLOT);
  const uint32_t lane_dst = ring + (uint32_t)(lane * 16);
  int64_t mi = wbeg;                                                           // issue stream
  int xi = (int)(wbeg % p.W), yi = (int)((wbeg / p.W) % p.H);
  const int xl = p.W - 1, yl = p.H - 1;
  uint32_t ioff = 0;                                                           // byte offset of the slot being filled
  auto issue = [&]() {
    if (mi < wend) {
      const bool oob = never || (nt && yi == 0) || (nb && yi == yl) || (nl && xi == 0) || (nr && xi == xl);
      if (active)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(lane_dst + ioff), "l"(oob ? safe : src), "r"(oob ? 0 : 16) : "memory");
      src += step;
      ++mi;
      if (xi == xl) { xi = 0; yi = yi == yl ? 0 : yi + 1; } else ++xi;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    ioff = ioff == (uint32_t)((D - 1) * SLOT) ? 0u : ioff + SLOT;
  };
#pragma unroll
  for (int k = 0; k < D - 1; ++k) issue();
  uint32_t coff = 0;                                                           // byte offset of the slot being consumed
  for (int64_t m = wbeg; m < wend; ++m) {
    __syncwarp();                                   // every lane has read the slot that is refilled now
    issue();
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
    __syncwarp();                                   // ... and every lane's chunk of pixel m has landed
    const uint32_t sb = ring + coff;
    coff = coff == (uint32_t)((D - 1) * SLOT) ? 0u : coff + SLOT;
    float g;
    if (ES == 4) {
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g) : "r"(sb + NXC * 16 + lane * 4));
    } else {
      unsigned short h;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(sb + NXC * 16 + lane * 2));
      g = __uint_as_float((uint32_t)h << 16);
    }
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      float xv[8];
      uint32_t w[4 * CPT];
#pragma unroll
      for (int h2 = 0; h2 < CPT; ++h2)
        if (h2 == 0 || CIN > 4)
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w[4 * h2]), "=r"(w[4 * h2 + 1]), "=r"(w[4 * h2 + 2]), "=r"(w[4 * h2 + 3])
                       : "r"(sb + (uint32_t)((t * CPT + h2) * 16)));
      if (ES == 4) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[c] = __uint_as_float(w[c % (4 * CPT)]);
      } else {
#pragma unroll
        for (int c = 0; c < CIN; ++c) xv[c] = __uint_as_float((c & 1) ? (w[(c >> 1) % (4 * CPT)] & 0xFFFF0000u) : (w[(c >> 1) % (4 * CPT)] << 16));
      }
#pragma unroll
      for (int c = 0; c < CIN; ++c) acc[t][c] = fmaf(g, xv[c], acc[t][c]);
    }
  }
  // sum the 8 warps in a fixed order
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < CIN; ++c) red[t * CIN + c][lane] = (w == 0 ? 0.f : red[t * CIN + c][lane]) + acc[t][c];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < taps * 32 * cin; i += 256) {
    const int ci = i % cin, l = (i / cin) % 32, t = i / (cin * 32);
    const int c2 = blockIdx.y * 32 + l;
    if (c2 < p.cout) part[(((int64_t)blockIdx.x * taps + t) * p.cout + c2) * cin + ci] = red[t * CIN + ci][l];
  }
}

// column sums over pixels: partial[chunk][C]  (vectorised: 8 channels per thread, several rows in flight)
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, int ld, int C, int64_t M,
                                                         int rows_per_chunk, float* __restrict__ part) {
  pdl_enter();
  extern __shared__ float sm[];  // [RY][V][8]
  const int V = C / 8, RY = 256 / V;
  const int v = threadIdx.x % V, ry = threadIdx.x / V;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  if (ry < RY) {
    int64_t r = r0 + ry;
    for (; r + 3 * (int64_t)RY < r1; r += 4 * (int64_t)RY) {
      float f0[8], f1[8], f2[8], f3[8];
      Vec8<T>::load(x + r * ld + v * 8, f0);
      Vec8<T>::load(x + (r + RY) * ld + v * 8, f1);
      Vec8<T>::load(x + (r + 2 * (int64_t)RY) * ld + v * 8, f2);
      Vec8<T>::load(x + (r + 3 * (int64_t)RY) * ld + v * 8, f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
    }
    for (; r < r1; r += RY) {
      float f0[8];
      Vec8<T>::load(x + r * ld + v * 8, f0);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += f0[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[(ry * V + v) * 8 + j] = a[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * 8; i += 256) {
    float s = 0.f;
    for (int q = 0; q < RY; ++q) s += sm[q * V * 8 + i];
    part[(int64_t)blockIdx.x * C + i] = s;
  }
}

template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, int ld, int C, int64_t M, int rows_per_chunk,
                                      float* __restrict__ part) {
  pdl_enter();
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += to_f<T>(x[r * ld + c]);
  part[(int64_t)blockIdx.x * C + c] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int nchunk, int C, float* __restrict__ out,
                                    int accumulate) {
  pdl_enter();
  // one warp per channel, lanes stride over the chunk partials, fixed shuffle tree
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s = 0.0;
  for (int k = lane; k < nchunk; k += 32) s += (double)part[(int64_t)k * C + c];
  s = warp_sum_d(s);
  if (lane == 0) out[c] = accumulate ? out[c] + (float)s : (float)s;
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int ks,
                                   int tflip, int rtf32) {
  const int taps = ks * ks;
  const int64_t n = (int64_t)taps * cout * cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!tflip) {  // out[tap][co][ci]
    const int ci = (int)(i % cin), co = (int)((i / cin) % cout), tap = (int)(i / ((int64_t)cin * cout));
    const float v = w[((int64_t)co * cin + ci) * taps + tap];
    out[i] = from_f<T>(rtf32 ? round_tf32_f(v) : v);
  } else {  // out[tap][ci][co] = w[co][ci][mirror(tap)]
    const int co = (int)(i % cout), ci = (int)((i / cout) % cin), tap = (int)(i / ((int64_t)cin * cout));
    const float v = w[((int64_t)co * cin + ci) * taps + (taps - 1 - tap)];
    out[i] = from_f<T>(rtf32 ? round_tf32_f(v) : v);
  }
}

constexpr int PACK_MAX = 96;
struct PackTable { PackEntry e[PACK_MAX]; int start[PACK_MAX + 1]; int n; };   // start: element offsets / 256 (blocks)

// one CTA = 256 consecutive elements of one entry (entries are padded to whole CTAs)
template <typename T>
__global__ void pack_weights_batched_kernel(const __grid_constant__ PackTable t, int rtf32) {
  int lo = 0, hi = t.n;                       // largest k with start[k] <= blockIdx.x
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (t.start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
  const PackEntry& e = t.e[lo];
  const int taps = e.ks * e.ks;
  const int64_t n = (int64_t)taps * e.cout * e.cin;
  const int64_t i = (int64_t)(blockIdx.x - t.start[lo]) * 256 + threadIdx.x;
  if (i >= n) return;
  T* out = (T*)e.out;
  if (!e.tflip) {  // out[tap][co][ci]
    const int ci = (int)(i % e.cin), co = (int)((i / e.cin) % e.cout), tap = (int)(i / ((int64_t)e.cin * e.cout));
    const float v = e.w[((int64_t)co * e.cin + ci) * taps + tap];
    out[i] = from_f<T>(rtf32 ? round_tf32_f(v) : v);
  } else {  // out[tap][ci][co] = w[co][ci][mirror(tap)]
    const int co = (int)(i % e.cout), ci = (int)((i / e.cout) % e.cin), tap = (int)(i / ((int64_t)e.cin * e.cout));
    const float v = e.w[((int64_t)co * e.cin + ci) * taps + (taps - 1 - tap)];
    out[i] = from_f<T>(rtf32 ? round_tf32_f(v) : v);
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
                                    T* __restrict__ y, int ldy, int B, int64_t HW) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * HW) return;
  const int64_t b = i / HW, p = i % HW;
  for (int c = 0; c < c0; ++c) y[i * ldy + c] = from_f<T>(x0[(b * c0 + c) * HW + p]);
  for (int c = 0; c < c1; ++c) y[i * ldy + c0 + c] = from_f<T>(x1[(b * c1 + c) * HW + p]);
  for (int c = c0 + c1; c < ldy; ++c) y[i * ldy + c] = from_f<T>(0.f);  // padding channels: finite, never NaN
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int ld, int C, float* __restrict__ y, int B, int64_t HW,
                                    int accumulate) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * HW) return;
  const int64_t b = i / HW, p = i % HW;
  for (int c = 0; c < C; ++c) {
    const float v = to_f<T>(x[i * ld + c]);
    float* o = y + (b * C + c) * HW + p;
    *o = accumulate ? *o + v : v;
  }
}

}  // namespace

int conv_simt(const ConvParams& p, int dtype, cudaStream_t s) {
  const int64_t M = (int64_t)p.B * p.H * p.W;
  // Cin <= 8 with an 8-channel padded, 32-byte aligned input buffer (how the engines stage the network input)
  if (p.c1 == 0 && p.c0 <= 8 && p.ld0 % 8 == 0 && (((uintptr_t)p.x0) & 31) == 0 && (p.ks == 1 || p.ks == 3)) {
    dim3 grid(cdiv(M, 128), cdiv(p.cout, 32));
    const int cc = p.c0 <= 3 ? 3 : (p.c0 <= 6 ? 6 : 8);
#define PUB_SMALLC(TT, CC) launch_pdl(conv_smallc_kernel<TT, CC>, grid, 128, 0, s, p)
    if (dtype == PUB_BF16) { if (cc == 3) PUB_SMALLC(bf16, 3); else if (cc == 6) PUB_SMALLC(bf16, 6); else PUB_SMALLC(bf16, 8); }
    else { if (cc == 3) PUB_SMALLC(float, 3); else if (cc == 6) PUB_SMALLC(float, 6); else PUB_SMALLC(float, 8); }
#undef PUB_SMALLC
    PUB_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(cdiv(M, BM), cdiv(p.cout, BN));
  if (dtype == PUB_BF16) launch_pdl(conv_simt_kernel<bf16>, grid, NT, 0, s, p);
  else launch_pdl(conv_simt_kernel<float>, grid, NT, 0, s, p);
  PUB_LAUNCH_CHECK();
  return 0;
}

static bool smallc_ok(const WgradParams& p) {
  // the 8-channel padded, 32-byte aligned staging buffer of the network input (PixLoad reads 8..32 B per pixel)
  // ... and 16-byte chunks of the dy rows for the cp.async ring of wgrad_smallc_kernel
  return p.c1 == 0 && p.c0 <= 8 && (p.ks == 1 || p.ks == 3) && p.ld0 % 8 == 0 && (((uintptr_t)p.x0) & 31) == 0 &&
         p.cout % 8 == 0 && p.ld_dy % 8 == 0 && (((uintptr_t)p.dy) & 15) == 0;
}
static void smallc_plan(const WgradParams& p, int& nctas, int& ppc) {
  const int64_t M = (int64_t)p.B * p.H * p.W;
  int want = 4 * num_sms() / cdiv(p.cout, 32);
  if (want < 1) want = 1;
  ppc = (int)((M + want - 1) / want);
  if (ppc < 256) ppc = 256;
  nctas = cdiv(M, ppc);
}

static void wgrad_simt_plan(const WgradParams& p, int& nsplit, int& pps) {
  const int cin = p.c0 + p.c1, taps = p.ks * p.ks;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int tiles = cdiv(cin, BN) * cdiv(p.cout, BM) * taps;
  int want = cdiv(4 * num_sms(), tiles);
  const int maxs = (int)((M + 511) / 512);
  nsplit = want < 1 ? 1 : (want > maxs ? maxs : want);
  if (nsplit < 1) nsplit = 1;
  pps = (int)(((M + nsplit - 1) / nsplit + BK - 1) / BK * BK);
  nsplit = cdiv(M, pps);
}

size_t wgrad_simt_workspace(const WgradParams& p) {
  int nsplit, pps;
  wgrad_simt_plan(p, nsplit, pps);
  if (smallc_ok(p)) {
    int nctas, ppc;
    smallc_plan(p, nctas, ppc);
    if (nctas > nsplit) nsplit = nctas;
  }
  const int cin = p.c0 + p.c1, taps = p.ks * p.ks;
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const size_t a = (size_t)nsplit * taps * p.cout * cin * sizeof(float);
  const size_t b = (size_t)cdiv(M, 1024) * p.cout * sizeof(float);
  return align_up(a, 256) + align_up(b, 256);
}

int wgrad_finish(const float* part, float* dw, int nsplit, int taps, int cout, int cin, const float* bpart, int nchunk,
                 float* dbias, int accumulate, cudaStream_t s) {
  const int64_t n = (int64_t)taps * cout * cin;
  const int nbb = dbias ? cdiv(cout, 4) : 0;
  if (nsplit >= 32 && n <= 128 * 1024) {
    const int nbw = cdiv(n, 32);
    launch_pdl(wgrad_finish_kernel<4>, nbw + nbb, 128, 0, s, part, dw, nsplit, taps, cout, cin, nbw, bpart, nchunk, dbias, accumulate);
  } else {
    const int nbw = cdiv(n, 128);
    launch_pdl(wgrad_finish_kernel<1>, nbw + nbb, 128, 0, s, part, dw, nsplit, taps, cout, cin, nbw, bpart, nchunk, dbias, accumulate);
  }
  PUB_LAUNCH_CHECK();
  return 0;
}

// out == nullptr: partial sums only (part[chunk][C]); returns the chunk count through *nchunk_out
int colsum(const void* x, int ld, int C, int64_t M, int dtype, float* part, float* out, int accumulate,
           cudaStream_t s, int* nchunk_out) {
  const int rows = 1024;
  const int nchunk = cdiv(M, rows);
  if (nchunk_out) *nchunk_out = nchunk;
  if (C % 8 == 0 && ld % 8 == 0 && C <= 2048 && (((uintptr_t)x) & 31) == 0) {
    const size_t smem = (size_t)(256 / (C / 8)) * (C / 8) * 8 * sizeof(float);
    if (dtype == PUB_BF16) launch_pdl(colsum_vec_kernel<bf16>, nchunk, 256, smem, s, (const bf16*)x, ld, C, M, rows, part);
    else launch_pdl(colsum_vec_kernel<float>, nchunk, 256, smem, s, (const float*)x, ld, C, M, rows, part);
  } else {
    dim3 grid(nchunk, cdiv(C, 64));
    if (dtype == PUB_BF16) launch_pdl(colsum_partial_kernel<bf16>, grid, 64, 0, s, (const bf16*)x, ld, C, M, rows, part);
    else launch_pdl(colsum_partial_kernel<float>, grid, 64, 0, s, (const float*)x, ld, C, M, rows, part);
  }
  PUB_LAUNCH_CHECK();
  if (!out) return 0;
  launch_pdl(colsum_final_kernel, cdiv((int64_t)C * 32, 256), 256, 0, s, part, nchunk, C, out, accumulate);
  PUB_LAUNCH_CHECK();
  return 0;
}

int wgrad_reduce(const float* part, float* dw, int nsplit, int taps, int cout, int cin, int accumulate, cudaStream_t s) {
  const int64_t n = (int64_t)taps * cout * cin;
  launch_pdl(wgrad_reduce_kernel, cdiv(n, 128), 128, 0, s, part, dw, nsplit, taps, cout, cin, accumulate);
  PUB_LAUNCH_CHECK();
  return 0;
}

int wgrad_simt(const WgradParams& p, int dtype, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s) {
  int nsplit, pps;
  wgrad_simt_plan(p, nsplit, pps);
  const int cin = p.c0 + p.c1, taps = p.ks * p.ks;
  PUB_REQUIRE(ws_bytes >= wgrad_simt_workspace(p), "wgrad workspace too small (%zu < %zu)", ws_bytes,
              wgrad_simt_workspace(p));
  float* part = (float*)ws;
  if (smallc_ok(p)) {
    int nctas, ppc;
    smallc_plan(p, nctas, ppc);
    dim3 grid(nctas, cdiv(p.cout, 32));
    const int cc = p.c0 <= 3 ? 3 : (p.c0 <= 6 ? 6 : 8);
#define PUB_SMALLC(TT, CC)                                                                   \
  do {                                                                                       \
    if (p.ks == 3) launch_pdl(wgrad_smallc_kernel<TT, CC, 3>, grid, 256, 0, s, p, part, ppc); \
    else launch_pdl(wgrad_smallc_kernel<TT, CC, 1>, grid, 256, 0, s, p, part, ppc);           \
  } while (0)
    if (dtype == PUB_BF16) { if (cc == 3) PUB_SMALLC(bf16, 3); else if (cc == 6) PUB_SMALLC(bf16, 6); else PUB_SMALLC(bf16, 8); }
    else { if (cc == 3) PUB_SMALLC(float, 3); else if (cc == 6) PUB_SMALLC(float, 6); else PUB_SMALLC(float, 8); }
#undef PUB_SMALLC
    nsplit = nctas;
  } else {
    dim3 grid(cdiv(cin, BN), cdiv(p.cout, BM), taps * nsplit);
    if (dtype == PUB_BF16) launch_pdl(wgrad_simt_kernel<bf16>, grid, NT, 0, s, p, part, nsplit, pps);
    else launch_pdl(wgrad_simt_kernel<float>, grid, NT, 0, s, p, part, nsplit, pps);
  }
  PUB_LAUNCH_CHECK();
  const int64_t n = (int64_t)taps * p.cout * cin;
  float* bpart = (float*)((char*)ws + align_up((size_t)nsplit * n * sizeof(float), 256));
  int nchunk = 0;
  if (p.dbias) PUB_TRY(colsum(p.dy, p.ld_dy, p.cout, (int64_t)p.B * p.H * p.W, dtype, bpart, nullptr, 0, s, &nchunk));
  return wgrad_finish(part, p.dw, nsplit, taps, p.cout, cin, bpart, nchunk, p.dbias, accumulate, s);
}

int pack_weight(const float* w, void* out, int cout, int cin, int ks, int dtype, int tflip, cudaStream_t s) {
  const int64_t n = (int64_t)ks * ks * cout * cin;
  if (dtype == PUB_BF16) pack_weight_kernel<bf16><<<cdiv(n, 256), 256, 0, s>>>(w, (bf16*)out, cout, cin, ks, tflip, 0);
  else pack_weight_kernel<float><<<cdiv(n, 256), 256, 0, s>>>(w, (float*)out, cout, cin, ks, tflip, dtype == PUB_TF32);
  PUB_LAUNCH_CHECK();
  note_weight_pack(s);
  return 0;
}

int pack_weights_batched(const PackEntry* e, int n, int dtype, cudaStream_t s) {
  for (int b0 = 0; b0 < n; b0 += PACK_MAX) {
    PackTable t;
    t.n = std::min(PACK_MAX, n - b0);
    int blocks = 0;
    for (int k = 0; k < t.n; ++k) {
      t.e[k] = e[b0 + k];
      t.start[k] = blocks;
      blocks += cdiv((int64_t)e[b0 + k].ks * e[b0 + k].ks * e[b0 + k].cout * e[b0 + k].cin, 256);
    }
    t.start[t.n] = blocks;
    if (blocks == 0) continue;
    if (dtype == PUB_BF16) pack_weights_batched_kernel<bf16><<<blocks, 256, 0, s>>>(t, 0);
    else pack_weights_batched_kernel<float><<<blocks, 256, 0, s>>>(t, dtype == PUB_TF32);
    PUB_LAUNCH_CHECK();
    note_weight_pack(s);
  }
  return 0;
}

int nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* y, int ldy, int B, int H, int W, int dtype,
                 cudaStream_t s) {
  const int64_t n = (int64_t)B * H * W;
  if (dtype == PUB_BF16)
    launch_pdl(nchw_to_nhwc_kernel<bf16>, cdiv(n, 256), 256, 0, s, x0, c0, x1, c1, (bf16*)y, ldy, B, (int64_t)H * W);
  else
    launch_pdl(nchw_to_nhwc_kernel<float>, cdiv(n, 256), 256, 0, s, x0, c0, x1, c1, (float*)y, ldy, B, (int64_t)H * W);
  PUB_LAUNCH_CHECK();
  return 0;
}

int nhwc_to_nchw(const void* x, int ld, int C, float* y, int B, int H, int W, int dtype, int accumulate,
                 cudaStream_t s) {
  const int64_t n = (int64_t)B * H * W;
  if (dtype == PUB_BF16)
    launch_pdl(nhwc_to_nchw_kernel<bf16>, cdiv(n, 256), 256, 0, s, (const bf16*)x, ld, C, y, B, (int64_t)H * W, accumulate);
  else
    launch_pdl(nhwc_to_nchw_kernel<float>, cdiv(n, 256), 256, 0, s, (const float*)x, ld, C, y, B, (int64_t)H * W, accumulate);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // namespace pub
