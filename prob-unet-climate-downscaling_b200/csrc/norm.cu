// GroupNorm (+FiLM scale/shift) + SiLU (+dropout) (+2x box resample), forward and backward, NHWC.
// Replaces F.group_norm / silu / addcmul / F.dropout / the depthwise resample convs of
// src/networks.py:105-107,166-177,83-87 and what autograd derives from them.
//
// All reductions are two-stage with a fixed summation order (per-chunk partials -> finalize in
// double), so results are run-to-run deterministic (no float atomics).
//
// HBM-bound by construction: the apply/backward passes read each element once as 16 B (bf16) or
// 32 B (f32) vectors and write once; statistics add one extra read of x.
#include "common.cuh"
#include "kernels.h"

namespace pub {

namespace {

constexpr int GN_NT = 256;
// Pixels per chunk (one CTA), GnParams::rows, picked by gn_pick_rows() on the host.  These kernels are bound by the
// latency chain of a CTA's few load -> compute -> store iterations, not by HBM or issue rate, and a partly filled
// wave takes as long as a full one (ncu, B = 64: 256 CTAs 23 us, 512 CTAs = 1.15 waves 41 us).  So the chunk count per
// image is the largest that keeps B * chunks within ONE wave of resident CTAs (3 per SM); beyond one wave (large
// batches) the old rule -- about 8 vectors per thread -- applies.
inline int gn_pick_rows(int B, int HW, int C) {
  const int V = C / 8;
  const int slots = 3 * num_sms();
  int n = slots / (B > 0 ? B : 1);                       // chunks per image that still fit one wave
  const int min_rows = std::max(8, 2 * GN_NT / V);       // >= 2 vectors per thread
  if (n >= 1) {
    int rows = (HW + n - 1) / n;
    if (rows < min_rows) rows = min_rows;
    // large maps: several waves of the classic 8-vectors-per-thread chunks beat one wave of very long CTAs only
    // when the tail is small; one wave wins or ties in every case measured, so keep it
    return rows < HW ? rows : HW;
  }
  const int rows = 8 * GN_NT / V;
  return rows < HW ? rows : HW;
}

// Every heavy kernel uses the same decomposition: grid = (pixel chunks, batch); inside a CTA thread t owns the
// 8-channel vector v = t % V for the rows pr, pr + ppi, ... of the chunk (ppi = 256 / V).  The per-channel
// constants of that vector (affine a/b, backward c1/c2/c3) are loaded ONCE into registers, so the inner loop is
// 16/32-byte vector traffic (through the cp.async prefetch ring below) + ~10 FLOP per element.

template <typename T>
__device__ __forceinline__ void load_vec(const GnParams& p, int64_t pix, int v, float (&f)[8]) {
  const int c = v * 8;
  if (c < p.c0) Vec8<T>::load((const T*)p.x0 + pix * p.ld0 + c, f);
  else Vec8<T>::load((const T*)p.x1 + pix * p.ld1 + (c - p.c0), f);
}

template <typename T>
__device__ __forceinline__ const T* vec_ptr(const GnParams& p, int64_t pix, int v) {
  const int c = v * 8;
  return c < p.c0 ? (const T*)p.x0 + pix * p.ld0 + c : (const T*)p.x1 + pix * p.ld1 + (c - p.c0);
}
// ---- per-thread prefetch ring in shared memory (cp.async)
// A thread's rows r, r + ppi, ... are requested D - 1 iterations ahead into cells only that thread reads back, so
// there is no block-level synchronisation: cp.async.wait_group orders a thread's own copies.  The loads of the next
// rows are in flight while the current row is being computed, at no register cost (raw-vector registers held the
// look-ahead before and capped it at two rows).  Everything per row is incremental -- source pointers advance by a
// constant byte stride, the stage offsets wrap -- because ncu showed these kernels as much issue-bound as
// latency-bound.  A stage holds NTEN tensors as planes of GN_NT 16-byte cells (conflict-free 128-bit accesses; two
// planes per tensor for f32).  (Tried on top of this: the forward pass storing the dropout keep bits, one byte per
// vector, for the backward kernels to read through the ring instead of re-running Philox -- 108 of their ~300
// instructions per vector.  Backward kernel times did not move, so they are not issue-bound any more; removed.)
template <typename T> struct RingCfg {
  static constexpr int HV = sizeof(T) / 2;                  // 16-byte halves per 8-channel vector
  static constexpr int D = sizeof(T) == 2 ? 4 : 3;          // stages
};
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int NTEN> struct RowRing {
  static constexpr int HV = RingCfg<T>::HV, D = RingCfg<T>::D;
  static constexpr uint32_t PLANE = GN_NT * 16;
  static constexpr uint32_t STAGE = NTEN * HV * PLANE;
  static constexpr size_t BYTES = (size_t)D * STAGE;
  uint32_t base16;               // this thread's cell in the planes of stage 0
  uint32_t ioff, coff;           // byte offsets of the stage being filled / consumed
  int left;                      // rows still to request
  const char* src[NTEN];         // next row of each tensor (nullptr: tensor absent)
  int step[NTEN];                // bytes between this thread's consecutive rows
  __device__ __forceinline__ void init(uint32_t ring, int t, int rows) {
    base16 = ring + (uint32_t)(t << 4);
    ioff = coff = 0; left = rows;
#pragma unroll
    for (int i = 0; i < NTEN; ++i) { src[i] = nullptr; step[i] = 0; }
  }
  __device__ __forceinline__ void issue() {
    if (left > 0) {
#pragma unroll
      for (int i = 0; i < NTEN; ++i)
        if (src[i]) {
#pragma unroll
          for (int h = 0; h < HV; ++h) cp_async16(base16 + ioff + (uint32_t)((i * HV + h) * PLANE), src[i] + 16 * h);
          src[i] += step[i];
        }
      --left;
    }
    cp_async_commit();
    ioff = ioff == (D - 1) * STAGE ? 0u : ioff + STAGE;
  }
  __device__ __forceinline__ void prologue() {
#pragma unroll
    for (int k = 0; k < D - 1; ++k) issue();
  }
  // requests the row D - 1 ahead, waits for the current one and returns its stage offset
  __device__ __forceinline__ uint32_t next() {
    issue();
    cp_async_wait<D - 1>();
    const uint32_t c = coff;
    coff = coff == (D - 1) * STAGE ? 0u : coff + STAGE;
    return c;
  }
  __device__ __forceinline__ void read(uint32_t stage_off, int ten, float (&v)[8]) const {
    uint32_t w[4 * HV];
#pragma unroll
    for (int h = 0; h < HV; ++h)
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(w[4 * h]), "=r"(w[4 * h + 1]), "=r"(w[4 * h + 2]), "=r"(w[4 * h + 3])
                   : "r"(base16 + stage_off + (uint32_t)((ten * HV + h) * PLANE)));
    if (sizeof(T) == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(w[i % (4 * HV)]);
    }
  }
};
template <typename T, int NTEN> constexpr size_t ring_bytes() { return RowRing<T, NTEN>::BYTES; }

// coef[b][c][2] -> a[8], bb[8] for channels v*8 .. v*8+7
__device__ __forceinline__ void load_affine(const float* coef, int b, int C, int v, float (&a)[8], float (&bb)[8]) {
  const float4* q = reinterpret_cast<const float4*>(coef + ((int64_t)b * C + v * 8) * 2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = q[i];
    a[2 * i] = t.x; bb[2 * i] = t.y; a[2 * i + 1] = t.z; bb[2 * i + 1] = t.w;
  }
}

// gradient wrt the activated output at INPUT resolution, rebuilt from dy (output resolution) + dropout
template <typename T>
__device__ __forceinline__ void load_gy(const GnParams& p, const T* __restrict__ dy, int b, int r, int64_t pix, int C,
                                        int v, uint32_t thresh, uint32_t dkey, float inv_keep, float (&g)[8]) {
  if (p.resample == 0) {
    Vec8<T>::load(dy + pix * C + v * 8, g);
  } else {
    const int yy = r / p.W, xx = r % p.W;
    if (p.resample == 1) {  // forward was a 2x2 mean
      const int64_t q = ((int64_t)b * (p.H / 2) + yy / 2) * (p.W / 2) + xx / 2;
      Vec8<T>::load(dy + q * C + v * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
    } else {  // forward was a nearest 2x upsample
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t q = ((int64_t)b * (p.H * 2) + yy * 2 + (d >> 1)) * (p.W * 2) + xx * 2 + (d & 1);
        float h[8];
        Vec8<T>::load(dy + q * C + v * 8, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += h[j];
      }
    }
  }
  if (p.p_drop > 0.f) {
    bool keep[8];
    dropout_keep8(dkey, pix * C + v * 8, thresh, keep);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = keep[j] ? g[j] * inv_keep : 0.f;
  }
}

// ---------------------------------------------------------------- per-channel two-value partial sums
// MODE 0: (x, x^2)                 -- forward statistics
// MODE 1: (du, du * (x - mean_g))  -- backward; du = g_y * silu'(a x + b); sum du*xhat = rstd * second sum
//         (accumulated as sum du*x and corrected by -mean * sum du when the partial is written: 8 fewer live
//         registers and one fewer FLOP per element in the hot loop)
// Rows arrive through the per-thread cp.async ring (three rows ahead); with <= 64 (MODE 0) / <= 80 (MODE 1)
// registers three to four CTAs share an SM.
template <typename T, int MODE>
__global__ void __launch_bounds__(GN_NT, MODE == 0 ? 4 : 3) gn_partial_kernel(GnParams p, const T* __restrict__ dy,
                                                                               float* __restrict__ part) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];  // [ppi][V][16] reduction buffer, then the prefetch ring
  const int C = p.c0 + p.c1, V = C / 8;
  const int ppi = GN_NT / V;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int HW = p.H * p.W;
  const int rows = p.rows;
  const int r0 = chunk * rows, r1 = min(HW, r0 + rows);
  const int t = threadIdx.x;
  const int v = t % V, pr = t / V;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (pr < ppi) {
    float a[8], bb[8];
    if (MODE == 1) load_affine(p.coef, b, C, v, a, bb);
    const uint32_t thresh = drop_thresh(p.p_drop);
    const uint32_t dkey = dropout_key(p.seed, p.subseq) ^ (p.salt ? __ldg(p.salt) : 0u);
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    if (MODE == 0 || p.resample == 0) {
      // this thread's vector of pixel (b, 0) and its pixel pitch (the two sources of a virtual concat differ)
      const int c = v * 8;
      const T* xb = c < p.c0 ? (const T*)p.x0 + (int64_t)b * HW * p.ld0 + c
                             : (const T*)p.x1 + (int64_t)b * HW * p.ld1 + (c - p.c0);
      const int64_t ldx = c < p.c0 ? p.ld0 : p.ld1;
      const T* gb = MODE == 1 ? dy + (int64_t)b * HW * C + c : nullptr;
      constexpr int NTEN = MODE == 0 ? 1 : 2;
      const int K = r0 + pr < r1 ? (r1 - r0 - pr + ppi - 1) / ppi : 0;               // rows of this thread
      RowRing<T, NTEN> rr;
      rr.init(smem_addr(sm) + (uint32_t)(GN_NT * 16 * sizeof(float)), t, K);          // behind the reduction buffer
      rr.src[0] = (const char*)(xb + (r0 + pr) * ldx); rr.step[0] = (int)(ppi * ldx * (int64_t)sizeof(T));
      int64_t e = ((int64_t)b * HW + r0 + pr) * C + c;                                // first element of the row vector
      const int64_t estep = (int64_t)ppi * C;
      if (MODE == 1) { rr.src[1] = (const char*)(gb + (int64_t)(r0 + pr) * C); rr.step[1] = (int)(estep * (int64_t)sizeof(T)); }
      rr.prologue();
      for (int k = 0; k < K; ++k) {
        const uint32_t st = rr.next();
        float x[8];
        rr.read(st, 0, x);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1[j] += x[j]; s2[j] = fmaf(x[j], x[j], s2[j]); }
        } else {
          float g[8];
          rr.read(st, 1, g);
          if (p.p_drop > 0.f) {
            bool keep[8];
            dropout_keep8(dkey, e, thresh, keep);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = keep[j] ? g[j] * inv_keep : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float du = g[j] * silu_grad_t<T>(fmaf(a[j], x[j], bb[j]));
            s1[j] += du;
            s2[j] = fmaf(du, x[j], s2[j]);
          }
        }
        e += estep;
      }
    } else {
      for (int r = r0 + pr; r < r1; r += ppi) {
        const int64_t pix = (int64_t)b * HW + r;
        float x[8], g[8];
        load_vec<T>(p, pix, v, x);
        load_gy<T>(p, dy, b, r, pix, C, v, thresh, dkey, inv_keep, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float du = g[j] * silu_grad_t<T>(fmaf(a[j], x[j], bb[j]));
          s1[j] += du;
          s2[j] = fmaf(du, x[j], s2[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { sm[(pr * V + v) * 16 + j] = s1[j]; sm[(pr * V + v) * 16 + 8 + j] = s2[j]; }
  }
  __syncthreads();
  // fixed-order sum over the row slots: thread i owns channel i (both sums)
  for (int c = t; c < C; c += GN_NT) {
    const int vv = c >> 3, jj = c & 7;
    float q1 = 0.f, q2 = 0.f;
    for (int q = 0; q < ppi; ++q) { q1 += sm[(q * V + vv) * 16 + jj]; q2 += sm[(q * V + vv) * 16 + 8 + jj]; }
    if (MODE == 1) q2 -= p.stats[((int64_t)b * p.groups + c / (C / p.groups)) * 2] * q1;
    *reinterpret_cast<float2*>(part + (((int64_t)b * gridDim.x + chunk) * C + c) * 2) = make_float2(q1, q2);
  }
}

// one warp per (b, g): mean / rstd, then the per-channel affine  y = silu(a x + b)
// partial rows: part0 holds channels [0, w0) in rows of w0 channels, part1 (optional) channels [w0, C) in rows of C - w0
// (one buffer of full-width rows from gn_partial_kernel, or one buffer per source of a virtual concat when the
// producing convs emitted them)
__global__ void gn_finalize_kernel(GnParams p, const float* __restrict__ part0, int w0, const float* __restrict__ part1,
                                   int nchunk) {
  pdl_enter();
  const int C = p.c0 + p.c1, cpg = C / p.groups;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= p.B * p.groups) return;
  const int b = wid / p.groups, g = wid % p.groups;
  double s = 0.0, ss = 0.0;
  // four independent loads in flight per lane: with conv-emitted partials there are hundreds of rows per image and
  // the loop is a chain of L2 round trips otherwise (fixed order: deterministic)
  const int total = nchunk * cpg;
  for (int i0 = lane; i0 < total; i0 += 128) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      v[u] = make_float2(0.f, 0.f);
      if (i < total) {
        const int k = i / cpg, c = g * cpg + i % cpg;
        const float* row = c < w0 ? part0 + (((int64_t)b * nchunk + k) * w0 + c) * 2
                                  : part1 + (((int64_t)b * nchunk + k) * (C - w0) + (c - w0)) * 2;
        v[u] = *reinterpret_cast<const float2*>(row);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { s += (double)v[u].x; ss += (double)v[u].y; }
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss);
  const double n = (double)cpg * p.H * p.W;
  const double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float meanf = (float)mean;
  if (lane == 0) { p.stats[((int64_t)b * p.groups + g) * 2] = meanf; p.stats[((int64_t)b * p.groups + g) * 2 + 1] = rstd; }
  for (int i = lane; i < cpg; i += 32) {
    const int c = g * cpg + i;
    const float sc = p.film ? 1.f + p.film[c] : 1.f, sh = p.film ? p.film[C + c] : 0.f;
    const float ga = p.gamma[c], be = p.beta[c];
    p.coef[((int64_t)b * C + c) * 2] = rstd * ga * sc;
    p.coef[((int64_t)b * C + c) * 2 + 1] = (be - meanf * rstd * ga) * sc + sh;
  }
}

// Same result, one CTA of 128 threads per (b, g): conv-emitted partials come as hundreds of rows per image (four per
// 8 x 16 output tile); a single warp walking them is a chain of L2 round trips (13.5 us per launch measured, against
// 3 us for the few rows of gn_partial_kernel).  Fixed thread -> element assignment and a fixed-order final sum in
// double: deterministic.
__global__ void __launch_bounds__(128) gn_finalize_wide_kernel(GnParams p, const float* __restrict__ part0, int w0,
                                                               const float* __restrict__ part1, int nchunk) {
  pdl_enter();
  __shared__ double red[2][4];
  const int C = p.c0 + p.c1, cpg = C / p.groups;
  const int b = blockIdx.x / p.groups, g = blockIdx.x % p.groups;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  double s = 0.0, ss = 0.0;
  const int total = nchunk * cpg;
  for (int i0 = t; i0 < total; i0 += 512) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 128 * u;
      v[u] = make_float2(0.f, 0.f);
      if (i < total) {
        const int k = i / cpg, c = g * cpg + i % cpg;
        const float* row = c < w0 ? part0 + (((int64_t)b * nchunk + k) * w0 + c) * 2
                                  : part1 + (((int64_t)b * nchunk + k) * (C - w0) + (c - w0)) * 2;
        v[u] = *reinterpret_cast<const float2*>(row);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { s += (double)v[u].x; ss += (double)v[u].y; }
  }
  s = warp_sum_d(s); ss = warp_sum_d(ss);
  if (lane == 0) { red[0][w] = s; red[1][w] = ss; }
  __syncthreads();
  s = ((red[0][0] + red[0][1]) + red[0][2]) + red[0][3];
  ss = ((red[1][0] + red[1][1]) + red[1][2]) + red[1][3];
  const double n = (double)cpg * p.H * p.W;
  const double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float meanf = (float)mean;
  if (t == 0) { p.stats[((int64_t)b * p.groups + g) * 2] = meanf; p.stats[((int64_t)b * p.groups + g) * 2 + 1] = rstd; }
  for (int i = t; i < cpg; i += 128) {
    const int c = g * cpg + i;
    const float sc = p.film ? 1.f + p.film[c] : 1.f, sh = p.film ? p.film[C + c] : 0.f;
    const float ga = p.gamma[c], be = p.beta[c];
    p.coef[((int64_t)b * C + c) * 2] = rstd * ga * sc;
    p.coef[((int64_t)b * C + c) * 2 + 1] = (be - meanf * rstd * ga) * sc + sh;
  }
}

// y = resample(dropout(silu(a x + b))); the chunk index runs over INPUT pixels (none / up) or OUTPUT pixels (down)
template <typename T>
__global__ void __launch_bounds__(GN_NT, 3) gn_apply_kernel(GnParams p, T* __restrict__ y) {
  pdl_enter();
  const int C = p.c0 + p.c1, V = C / 8, ppi = GN_NT / V;
  const int b = blockIdx.y, t = threadIdx.x, v = t % V, pr = t / V;
  if (pr >= ppi) return;
  float a[8], bb[8];
  load_affine(p.coef, b, C, v, a, bb);
  const int HW = p.H * p.W;
  if (p.resample != 1) {
    const uint32_t thresh = drop_thresh(p.p_drop);
    const uint32_t dkey = dropout_key(p.seed, p.subseq) ^ (p.salt ? __ldg(p.salt) : 0u);
    const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
    const int rows = p.rows;
    const int r0 = blockIdx.x * rows, r1 = min(HW, r0 + rows);
    const int c = v * 8;
    const T* xb = c < p.c0 ? (const T*)p.x0 + (int64_t)b * HW * p.ld0 + c
                           : (const T*)p.x1 + (int64_t)b * HW * p.ld1 + (c - p.c0);
    const int64_t ldx = c < p.c0 ? p.ld0 : p.ld1;
    extern __shared__ __align__(16) uint8_t gn_ring_raw[];
    const int K = r0 + pr < r1 ? (r1 - r0 - pr + ppi - 1) / ppi : 0;    // rows of this thread
    RowRing<T, 1> rr;
    rr.init(smem_addr(gn_ring_raw), t, K);
    rr.src[0] = (const char*)(xb + (r0 + pr) * ldx); rr.step[0] = (int)(ppi * ldx * (int64_t)sizeof(T));
    rr.prologue();
    int r = r0 + pr;
    int64_t e = ((int64_t)b * HW + r) * C + c;                          // first element of the row vector
    const int64_t estep = (int64_t)ppi * C;
    for (int k = 0; k < K; ++k, r += ppi, e += estep) {
      const uint32_t st = rr.next();
      float x[8], o[8];
      rr.read(st, 0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = silu_t<T>(fmaf(a[j], x[j], bb[j]));
      if (p.p_drop > 0.f) {
        bool keep[8];
        dropout_keep8(dkey, e, thresh, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = keep[j] ? o[j] * inv_keep : 0.f;
      }
      if (p.resample == 0) {
        Vec8<T>::store(y + e, o);
      } else {  // nearest 2x upsample: write the 2x2 children
        const int yy = r / p.W, xx = r % p.W;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const int64_t q = ((int64_t)b * (p.H * 2) + yy * 2 + (d >> 1)) * (p.W * 2) + xx * 2 + (d & 1);
          Vec8<T>::store(y + q * C + c, o);
        }
      }
    }
  } else {  // 2x2 mean of the activated values
    const int Ho = p.H / 2, Wo = p.W / 2, HWo = Ho * Wo;
    const int rows = p.rows;
    const int r0 = blockIdx.x * rows, r1 = min(HWo, r0 + rows);
    for (int r = r0 + pr; r < r1; r += ppi) {
      const int yo = r / Wo, xo = r % Wo;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int64_t pix = ((int64_t)b * p.H + yo * 2 + (d >> 1)) * p.W + xo * 2 + (d & 1);
        float x[8];
        load_vec<T>(p, pix, v, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += silu_t<T>(fmaf(a[j], x[j], bb[j]));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
      Vec8<T>::store(y + ((int64_t)b * HWo + r) * C + v * 8, o);
    }
  }
}

// backward finalize 1: per (b, g) -> bcoef[b][c] = (c2, c3):  dx = a*du + c2*x + c3  (a = forward coefficient)
// partials hold P1 = sum du, P2 = sum du*(x - mean);  sum du*xhat = rstd * P2
// raw2 != 0: the second partial is sum du * x (conv epilogue, ConvParams::gn_bwd) and is centred here, row by row
__global__ void gn_bwd_group_kernel(GnParams p, const float* __restrict__ part, int nchunk, int raw2,
                                    float* __restrict__ bcoef, float* __restrict__ bsum) {
  pdl_enter();
  const int C = p.c0 + p.c1, cpg = C / p.groups;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= p.B * p.groups) return;
  const int b = wid / p.groups, g = wid % p.groups;
  // per channel: sum the chunk partials (lanes stride over chunks, fixed tree), keep (S1, rstd*S2) per (b, c)
  // for the parameter-gradient kernel, and fold them into the group sums
  const float rstd_g = p.stats[((int64_t)b * p.groups + g) * 2 + 1];
  const float mean_g = raw2 ? p.stats[((int64_t)b * p.groups + g) * 2] : 0.f;
  double q1 = 0.0, q2 = 0.0;
  for (int ci = 0; ci < cpg; ++ci) {
    const int c = g * cpg + ci;
    float a1 = 0.f, a2 = 0.f;
    for (int k0 = lane; k0 < nchunk; k0 += 128) {        // four independent loads in flight per lane
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + 32 * u;
        v[u] = k < nchunk ? *reinterpret_cast<const float2*>(part + (((int64_t)b * nchunk + k) * C + c) * 2) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a1 += v[u].x; a2 += fmaf(-mean_g, v[u].x, v[u].y); }
    }
    a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { bsum[((int64_t)b * C + c) * 2] = a1; bsum[((int64_t)b * C + c) * 2 + 1] = rstd_g * a2; }
    const float sc = p.film ? 1.f + p.film[c] : 1.f;
    const double gp = (double)p.gamma[c] * sc;
    q1 += gp * (double)a1; q2 += gp * (double)a2;
  }
  const float mean = p.stats[((int64_t)b * p.groups + g) * 2], rstd = p.stats[((int64_t)b * p.groups + g) * 2 + 1];
  const double n = (double)cpg * p.H * p.W;
  const double m1 = q1 / n;
  const double m2 = (double)rstd * q2 / n;
  const float c2 = (float)(-(double)rstd * rstd * m2);
  const float c3 = (float)(-(double)rstd * m1 + (double)rstd * rstd * m2 * mean);
  for (int i = lane; i < cpg; i += 32) {
    const int c = g * cpg + i;
    *reinterpret_cast<float2*>(bcoef + ((int64_t)b * C + c) * 2) = make_float2(c2, c3);
  }
}

// backward finalize 2: per channel parameter gradients = sum over the batch of the per-(b, c) sums written by
// gn_bwd_group_kernel.  One warp per channel: lanes stride over the batch, fixed shuffle tree (a thread per
// channel walking the batch serially took 50 us per launch -- 3 ms of a training step).
__global__ void gn_bwd_param_kernel(GnParams p, const float* __restrict__ bsum, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, float* __restrict__ dfilm) {
  pdl_enter();
  const int C = p.c0 + p.c1;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = lane; b < p.B; b += 32) {
    const float2 v = *reinterpret_cast<const float2*>(bsum + ((int64_t)b * C + c) * 2);
    s1 += (double)v.x; s2 += (double)v.y;
  }
  s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
  if (lane != 0) return;
  const float sc = p.film ? 1.f + p.film[c] : 1.f;
  dgamma[c] = (float)(s2 * sc);
  dbeta[c] = (float)(s1 * sc);
  if (dfilm) {
    dfilm[c] = (float)((double)p.gamma[c] * s2 + (double)p.beta[c] * s1);  // d scale
    dfilm[C + c] = (float)s1;                                              // d shift
  }
}

// dx = a*du + c2*x + c3 (+ addend);  a = rstd*gamma*(1+scale) is the forward coefficient, (c2, c3) come from the
// group sums (bcoef[b][c] = (c2, c3)).  x, dy and the addend arrive through the cp.async ring; <= 80 registers.
template <typename T, bool FROM_DU>
__global__ void __launch_bounds__(GN_NT, 3) gn_bwd_apply_kernel(GnParams p, const T* __restrict__ dy,
                                                                const float* __restrict__ bcoef, T* __restrict__ dx,
                                                                const T* __restrict__ addend, int ld_add) {
  pdl_enter();
  const int C = p.c0 + p.c1, V = C / 8, ppi = GN_NT / V;
  const int b = blockIdx.y, t = threadIdx.x, v = t % V, pr = t / V;
  if (pr >= ppi) return;
  float a[8], bb[8], c2[8], c3[8];
  load_affine(p.coef, b, C, v, a, bb);
  load_affine(bcoef, b, C, v, c2, c3);
  const uint32_t thresh = drop_thresh(p.p_drop);
  const uint32_t dkey = dropout_key(p.seed, p.subseq) ^ (p.salt ? __ldg(p.salt) : 0u);
  const float inv_keep = p.p_drop > 0.f ? 1.f / (1.f - p.p_drop) : 1.f;
  const int HW = p.H * p.W;
  const int rows = p.rows;
  const int r0 = blockIdx.x * rows, r1 = min(HW, r0 + rows);
  const int c = v * 8;
  if (p.resample == 0) {
    const T* xb = c < p.c0 ? (const T*)p.x0 + (int64_t)b * HW * p.ld0 + c
                           : (const T*)p.x1 + (int64_t)b * HW * p.ld1 + (c - p.c0);
    const int64_t ldx = c < p.c0 ? p.ld0 : p.ld1;
    const T* gb = dy + (int64_t)b * HW * C + c;
    const T* ab = addend ? addend + (int64_t)b * HW * ld_add + c : nullptr;
    T* ob = dx + (int64_t)b * HW * C + c;
    extern __shared__ __align__(16) uint8_t gn_ring_raw[];
    const int K = r0 + pr < r1 ? (r1 - r0 - pr + ppi - 1) / ppi : 0;    // rows of this thread
    int64_t e = ((int64_t)b * HW + r0 + pr) * C + c;                    // first element of the row vector
    const int64_t estep = (int64_t)ppi * C;
    RowRing<T, 3> rr;
    rr.init(smem_addr(gn_ring_raw), t, K);
    rr.src[0] = (const char*)(xb + (r0 + pr) * ldx); rr.step[0] = (int)(ppi * ldx * (int64_t)sizeof(T));
    rr.src[1] = (const char*)(gb + (int64_t)(r0 + pr) * C); rr.step[1] = (int)(estep * (int64_t)sizeof(T));
    if (ab) { rr.src[2] = (const char*)(ab + (int64_t)(r0 + pr) * ld_add); rr.step[2] = (int)((int64_t)ppi * ld_add * (int64_t)sizeof(T)); }
    rr.prologue();
    T* op = ob + (int64_t)(r0 + pr) * C;
    for (int k = 0; k < K; ++k, e += estep, op += estep) {
      const uint32_t st = rr.next();
      float x[8], g[8], o[8];
      rr.read(st, 0, x);
      rr.read(st, 1, g);
      if (!FROM_DU && p.p_drop > 0.f) {
        bool keep[8];
        dropout_keep8(dkey, e, thresh, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = keep[j] ? g[j] * inv_keep : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float du = FROM_DU ? g[j] : g[j] * silu_grad_t<T>(fmaf(a[j], x[j], bb[j]));
        o[j] = fmaf(a[j], du, fmaf(c2[j], x[j], c3[j]));
      }
      if (ab) {
        float ad[8];
        rr.read(st, 2, ad);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += ad[j];
      }
      Vec8<T>::store(op, o);
    }
    return;
  }
  for (int r = r0 + pr; r < r1; r += ppi) {
    const int64_t pix = (int64_t)b * HW + r;
    float x[8], g[8], o[8];
    load_vec<T>(p, pix, v, x);
    load_gy<T>(p, dy, b, r, pix, C, v, thresh, dkey, inv_keep, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float du = g[j] * silu_grad_t<T>(fmaf(a[j], x[j], bb[j]));
      o[j] = fmaf(a[j], du, fmaf(c2[j], x[j], c3[j]));
    }
    if (addend) {
      float ad[8];
      Vec8<T>::load(addend + pix * ld_add + c, ad);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += ad[j];
    }
    Vec8<T>::store(dx + pix * C + c, o);
  }
}

int check(const GnParams& p) {
  const int C = p.c0 + p.c1;
  PUB_REQUIRE(C % 8 == 0 && p.c0 % 8 == 0 && C / 8 <= GN_NT, "GroupNorm kernel needs C %% 8 == 0 and C <= 2048 (C=%d c0=%d)", C, p.c0);
  PUB_REQUIRE(p.groups > 0 && C % p.groups == 0, "GroupNorm: C=%d not divisible by groups=%d", C, p.groups);
  PUB_REQUIRE(p.ld0 % 8 == 0 && (p.c1 == 0 || p.ld1 % 8 == 0), "GroupNorm: pixel strides must be multiples of 8");
  PUB_REQUIRE(p.resample != 1 || (p.H % 2 == 0 && p.W % 2 == 0), "GroupNorm: 2x down needs even H, W");
  return 0;
}

inline int nchunks(const GnParams& p) { return cdiv((int64_t)p.H * p.W, p.rows); }
inline int grid_for(int64_t n) {
  int64_t g = (n + GN_NT - 1) / GN_NT;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < cap ? g : cap);
}

}  // namespace

// dynamic shared memory of the ring kernels (reduction buffer + prefetch ring) goes past the 48 KB default
static int set_gn_smem_attrs() {
  static bool done = false;
  if (done) return 0;
  const int red = GN_NT * 16 * (int)sizeof(float);
  const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
  PUB_CUDA(cudaFuncSetAttribute(gn_partial_kernel<bf16, 0>, at, red + (int)ring_bytes<bf16, 1>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_partial_kernel<float, 0>, at, red + (int)ring_bytes<float, 1>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_partial_kernel<bf16, 1>, at, red + (int)ring_bytes<bf16, 2>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_partial_kernel<float, 1>, at, red + (int)ring_bytes<float, 2>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_apply_kernel<bf16>, at, (int)ring_bytes<bf16, 1>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_apply_kernel<float>, at, (int)ring_bytes<float, 1>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel<bf16, false>, at, (int)ring_bytes<bf16, 3>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel<float, false>, at, (int)ring_bytes<float, 3>()));
  PUB_CUDA(cudaFuncSetAttribute(gn_bwd_apply_kernel<bf16, true>, at, (int)ring_bytes<bf16, 3>()));
  done = true;
  return 0;
}

size_t gn_partial_floats(int B, int C, int H, int W) {
  return (size_t)B * cdiv((int64_t)H * W, gn_pick_rows(B, H * W, C)) * C * 2 + (size_t)B * C * 6 + 64;
}

int gn_forward(const GnParams& p_, void* y, int dtype, cudaStream_t s) {
  PUB_TRY(check(p_));
  GnParams p = p_;
  p.rows = gn_pick_rows(p.B, p.H * p.W, p.c0 + p.c1);
  const int C = p.c0 + p.c1, V = C / 8, nc = nchunks(p);
  const size_t red = (size_t)GN_NT * 16 * sizeof(float);
  dim3 grid(nc, p.B);
  PUB_TRY(set_gn_smem_attrs());
  if (p.pre0 && (p.c1 == 0 || p.pre1)) {
    // the producing conv(s) already emitted the (sum, sum of squares) partials: no pass over x
    if (p.pre_rows * (C / p.groups) > 256)
      launch_pdl(gn_finalize_wide_kernel, p.B * p.groups, 128, 0, s, p, p.pre0, p.c0, p.pre1, p.pre_rows);
    else
      launch_pdl(gn_finalize_kernel, cdiv((int64_t)p.B * p.groups * 32, 128), 128, 0, s, p, p.pre0, p.c0, p.pre1, p.pre_rows);
    PUB_LAUNCH_CHECK();
  } else {
    if (dtype == PUB_BF16) launch_pdl(gn_partial_kernel<bf16, 0>, grid, GN_NT, red + ring_bytes<bf16, 1>(), s, p, nullptr, p.partial);
    else launch_pdl(gn_partial_kernel<float, 0>, grid, GN_NT, red + ring_bytes<float, 1>(), s, p, nullptr, p.partial);
    PUB_LAUNCH_CHECK();
    launch_pdl(gn_finalize_kernel, cdiv((int64_t)p.B * p.groups * 32, 128), 128, 0, s, p, (const float*)p.partial, C,
               (const float*)nullptr, nc);
    PUB_LAUNCH_CHECK();
  }
  dim3 agrid(cdiv((int64_t)p.H * p.W / (p.resample == 1 ? 4 : 1), p.rows), p.B);
  if (dtype == PUB_BF16) launch_pdl(gn_apply_kernel<bf16>, agrid, GN_NT, ring_bytes<bf16, 1>(), s, p, (bf16*)y);
  else launch_pdl(gn_apply_kernel<float>, agrid, GN_NT, ring_bytes<float, 1>(), s, p, (float*)y);
  PUB_LAUNCH_CHECK();
  return 0;
}

int gn_backward(const GnParams& p_, const void* dy, void* dx, const void* addend, int ld_add, float* dgamma,
                float* dbeta, float* dfilm, int dtype, cudaStream_t s) {
  PUB_TRY(check(p_));
  GnParams p = p_;
  p.rows = gn_pick_rows(p.B, p.H * p.W, p.c0 + p.c1);
  PUB_TRY(set_gn_smem_attrs());
  const int C = p.c0 + p.c1, V = C / 8, nc = nchunks(p);
  const size_t red = (size_t)GN_NT * 16 * sizeof(float);
  float* bcoef = p.partial + align_up((size_t)p.B * nc * C * 2, 4);  // 16-byte aligned rows of 4 floats
  dim3 grid(nc, p.B);
  if (dtype == PUB_BF16) launch_pdl(gn_partial_kernel<bf16, 1>, grid, GN_NT, red + ring_bytes<bf16, 2>(), s, p, (const bf16*)dy, p.partial);
  else launch_pdl(gn_partial_kernel<float, 1>, grid, GN_NT, red + ring_bytes<float, 2>(), s, p, (const float*)dy, p.partial);
  PUB_LAUNCH_CHECK();
  float* bsum = bcoef + (size_t)p.B * C * 2;
  launch_pdl(gn_bwd_group_kernel, cdiv((int64_t)p.B * p.groups * 32, 128), 128, 0, s, p, (const float*)p.partial, nc, 0, bcoef, bsum);
  PUB_LAUNCH_CHECK();
  launch_pdl(gn_bwd_param_kernel, cdiv((int64_t)C * 32, 256), 256, 0, s, p, bsum, dgamma, dbeta, dfilm);
  PUB_LAUNCH_CHECK();
  if (dx) {
    if (dtype == PUB_BF16)
      launch_pdl(gn_bwd_apply_kernel<bf16, false>, grid, GN_NT, ring_bytes<bf16, 3>(), s, p, (const bf16*)dy, bcoef, (bf16*)dx, (const bf16*)addend, ld_add);
    else
      launch_pdl(gn_bwd_apply_kernel<float, false>, grid, GN_NT, ring_bytes<float, 3>(), s, p, (const float*)dy, bcoef, (float*)dx, (const float*)addend, ld_add);
    PUB_LAUNCH_CHECK();
  }
  return 0;
}

int gn_backward_from_du(const GnParams& p_, const void* du, const float* du_part, int rows_per_image, void* dx,
                        const void* addend, int ld_add, float* dgamma, float* dbeta, float* dfilm, int dtype,
                        cudaStream_t s) {
  PUB_TRY(check(p_));
  PUB_REQUIRE(dtype == PUB_BF16 && p_.resample == 0 && du && du_part && dx, "gn_backward_from_du: bf16, no resampling");
  GnParams p = p_;
  p.rows = gn_pick_rows(p.B, p.H * p.W, p.c0 + p.c1);
  p.p_drop = 0.f;                                    // the dropout mask is already inside du
  PUB_TRY(set_gn_smem_attrs());
  const int C = p.c0 + p.c1, nc = nchunks(p);
  float* bcoef = p.partial;                          // scratch: [B][C][2] coefficients, then [B][C][2] per-(b,c) sums
  float* bsum = bcoef + (size_t)p.B * C * 2;
  launch_pdl(gn_bwd_group_kernel, cdiv((int64_t)p.B * p.groups * 32, 128), 128, 0, s, p, du_part, rows_per_image, 1, bcoef, bsum);
  PUB_LAUNCH_CHECK();
  launch_pdl(gn_bwd_param_kernel, cdiv((int64_t)C * 32, 256), 256, 0, s, p, (const float*)bsum, dgamma, dbeta, dfilm);
  PUB_LAUNCH_CHECK();
  dim3 grid(nc, p.B);
  launch_pdl(gn_bwd_apply_kernel<bf16, true>, grid, GN_NT, ring_bytes<bf16, 3>(), s, p, (const bf16*)du, (const float*)bcoef,
             (bf16*)dx, (const bf16*)addend, ld_add);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // namespace pub

using namespace pub;

extern "C" {

size_t pub_groupnorm_scratch_bytes(int B, int C, int H, int W) { return gn_partial_floats(B, C, H, W) * sizeof(float); }

static GnParams gn_abi_params(const void* x, int C, int ld, int B, int H, int W, const float* gamma, const float* beta,
                              const float* film, int resample, float p_drop, uint64_t seed, uint64_t subseq,
                              const float* stats, const float* coef, void* scratch) {
  GnParams g{};
  g.x0 = x; g.c0 = C; g.ld0 = ld; g.B = B; g.H = H; g.W = W;
  g.groups = C / 4 < 32 ? C / 4 : 32;                       // networks.GroupNorm: min(32, C // 4)
  g.gamma = gamma; g.beta = beta; g.film = film; g.resample = resample;
  g.p_drop = p_drop; g.seed = seed; g.subseq = subseq;
  g.stats = const_cast<float*>(stats); g.coef = const_cast<float*>(coef); g.partial = (float*)scratch;
  return g;
}

int pub_groupnorm_silu_forward(const void* x, int C, int ld, int B, int H, int W, const float* gamma, const float* beta,
                               const float* film, int resample, float p_drop, uint64_t seed, uint64_t subseq, void* y,
                               float* stats, float* coef, void* scratch, size_t scratch_bytes, int dtype, pub_stream_t s) {
  PUB_REQUIRE(x && gamma && beta && y && stats && coef && scratch, "pub_groupnorm_silu_forward: null argument");
  PUB_REQUIRE(dtype == PUB_F32 || dtype == PUB_BF16, "pub_groupnorm_silu_forward: dtype must be PUB_F32 or PUB_BF16");
  PUB_REQUIRE(C >= 4 && scratch_bytes >= pub_groupnorm_scratch_bytes(B, C, H, W), "pub_groupnorm_silu_forward: scratch too small");
  return gn_forward(gn_abi_params(x, C, ld, B, H, W, gamma, beta, film, resample, p_drop, seed, subseq, stats, coef, scratch),
                    y, dtype, (cudaStream_t)s);
}

int pub_groupnorm_silu_backward(const void* x, int C, int ld, int B, int H, int W, const float* gamma, const float* beta,
                                const float* film, int resample, float p_drop, uint64_t seed, uint64_t subseq,
                                const float* stats, const float* coef, const void* dy, void* dx, float* dgamma,
                                float* dbeta, float* dfilm, void* scratch, size_t scratch_bytes, int dtype, pub_stream_t s) {
  PUB_REQUIRE(x && gamma && beta && stats && coef && dy && dx && dgamma && dbeta && scratch, "pub_groupnorm_silu_backward: null argument");
  PUB_REQUIRE(dtype == PUB_F32 || dtype == PUB_BF16, "pub_groupnorm_silu_backward: dtype must be PUB_F32 or PUB_BF16");
  PUB_REQUIRE(C >= 4 && scratch_bytes >= pub_groupnorm_scratch_bytes(B, C, H, W), "pub_groupnorm_silu_backward: scratch too small");
  PUB_REQUIRE(film == nullptr || dfilm != nullptr, "pub_groupnorm_silu_backward: dfilm is required when film is given");
  return gn_backward(gn_abi_params(x, C, ld, B, H, W, gamma, beta, film, resample, p_drop, seed, subseq, stats, coef, scratch),
                     dy, dx, nullptr, 0, dgamma, dbeta, film ? dfilm : nullptr, dtype, (cudaStream_t)s);
}

}  // extern "C"
