// Ensemble post-processing diagnostics of results.ipynb on the GPU:
//   * psd / compute_psd_tensor (cell 4): azimuthally averaged 2-D power spectral density of every (sample, variable)
//     field, mean over samples; the inverse variable transforms of that cell (softplus / kgm2sTommday / KToC -- tasmax
//     with softplus(c = 0) THERE) are fused into the load
//   * the value histograms of cell 15 (np.histogram over given bin edges)
// The reference loops over T x M samples in Python (torch.fft.fftn + scipy.stats.binned_statistic per field, on the
// host).  Here one CTA owns one field: the field lives in shared memory as complex f32, a radix-2 FFT runs over the
// rows and then the columns (twiddles from a shared table computed in double), and the radial bins are summed from a
// precomputed pixel list (CSR by bin) in a fixed order -- no cuFFT, no atomics on floats, run-to-run deterministic.
#include "common.cuh"
#include "kernels.h"

namespace pub {
namespace {

constexpr int PSD_NT = 256;

__device__ __forceinline__ float softplus_nb(float x, float c) {
  const float sp = logf(expf(x) + 1.f) - c;
  return x > 20.f ? x : sp;
}

// in-place radix-2 decimation-in-time FFT of `count` lines of length n = 2^lg stored with element stride `es` and line
// stride `ls` in d; tw[k] = exp(-2 pi i k / n), k < n / 2.  All threads of the CTA participate.
__device__ void fft_lines(float2* d, const float2* tw, int n, int lg, int count, int es, int ls) {
  const int t = threadIdx.x;
  // bit reversal
  for (int i = t; i < count * n; i += PSD_NT) {
    const int line = i / n, k = i % n;
    const int r = (int)(__brev((unsigned)k) >> (32 - lg));
    if (r > k) {
      float2* a = d + line * ls + k * es;
      float2* b = d + line * ls + r * es;
      const float2 tmp = *a; *a = *b; *b = tmp;
    }
  }
  __syncthreads();
  for (int s = 1; s <= lg; ++s) {
    const int half = 1 << (s - 1), step = n >> s;       // twiddle stride
    for (int i = t; i < count * (n / 2); i += PSD_NT) {
      const int line = i / (n / 2), j = i % (n / 2);
      const int grp = j / half, pos = j % half;
      const int i0 = grp * 2 * half + pos, i1 = i0 + half;
      const float2 w = tw[pos * step];
      float2* a = d + line * ls + i0 * es;
      float2* b = d + line * ls + i1 * es;
      const float2 u = *a, v = *b;
      const float2 vw = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
      *a = make_float2(u.x + vw.x, u.y + vw.y);
      *b = make_float2(u.x - vw.x, u.y - vw.y);
    }
    __syncthreads();
  }
}

// grid (C, N): field (n, c) of data [N, C, H, W] -> psd [N, C, H/2]
// transfo: 0 none, 1 stored-transform domain (cell 4: pr = softplus(s0), tasmax = softplus(s2, c=0) + s1), then real
// units in both cases when units != 0 (pr * 86400 in kgm2sTommday's order, temperatures - 273.15)
__global__ void __launch_bounds__(PSD_NT) psd_kernel(const float* __restrict__ data, int C, int H, int lg, int transfo,
                                                     int units, const int* __restrict__ bin_ptr,
                                                     const int* __restrict__ bin_pix, float* __restrict__ psd) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  float2* d = reinterpret_cast<float2*>(sm_raw);                      // [H][H]
  float2* tw = d + H * H;                                             // [H/2]
  float* red = reinterpret_cast<float*>(tw + H / 2);                  // [nb][4]
  const int c = blockIdx.x, n = blockIdx.y, t = threadIdx.x, nb = H / 2;
  const float* f = data + ((int64_t)n * C + c) * H * H;
  const float* f1 = data + ((int64_t)n * C + 1) * H * H;              // tasmin (needed by tasmax)
  for (int i = t; i < H * H; i += PSD_NT) {
    float v = f[i];
    if (C == 3) {
      if (transfo) {
        if (c == 0) v = softplus_nb(v, 1e-7f);
        else if (c == 2) v = softplus_nb(v, 0.f) + f1[i];
      }
      if (units) v = c == 0 ? v * 24.f * 60.f * 60.f : v - 273.15f;
    }
    d[i] = make_float2(v, 0.f);
  }
  for (int k = t; k < H / 2; k += PSD_NT) {
    double sn, cs;
    sincospi(-2.0 * (double)k / (double)H, &sn, &cs);
    tw[k] = make_float2((float)cs, (float)sn);
  }
  __syncthreads();
  fft_lines(d, tw, H, lg, H, 1, H);        // rows
  fft_lines(d, tw, H, lg, H, H, 1);        // columns
  // radial bins: thread (b, q) sums quarter q of bin b's pixel list in list order, thread (b, 0) combines
  for (int i = t; i < nb * 4; i += PSD_NT) {
    const int b = i >> 2, q = i & 3;
    const int p0 = bin_ptr[b], p1 = bin_ptr[b + 1], len = p1 - p0;
    const int s0 = p0 + (int)(((int64_t)len * q) / 4), s1 = p0 + (int)(((int64_t)len * (q + 1)) / 4);
    float acc = 0.f;
    for (int j = s0; j < s1; ++j) { const float2 z = d[bin_pix[j]]; acc += z.x * z.x + z.y * z.y; }
    red[i] = acc;
  }
  __syncthreads();
  for (int b = t; b < nb; b += PSD_NT) {
    const int cnt = bin_ptr[b + 1] - bin_ptr[b];
    const float s = (red[4 * b] + red[4 * b + 1]) + (red[4 * b + 2] + red[4 * b + 3]);
    const double lo = b + 0.5, hi = b + 1.5;
    const float area = (float)(3.14159265358979323846 * (hi * hi - lo * lo));
    psd[((int64_t)n * C + c) * nb + b] = cnt > 0 ? (s / (float)cnt) * area : nanf("");   // empty bin: NaN like scipy
  }
}

// mean over the N samples of x [N, K] -> out [K]; one warp per column, lanes stride over N, double, fixed tree
__global__ void colmean_kernel(const float* __restrict__ x, int N, int K, float* __restrict__ out) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= K) return;
  double s = 0.0;
  for (int n = lane; n < N; n += 32) s += (double)x[(int64_t)n * K + k];
  s = warp_sum_d(s);
  if (lane == 0) out[k] = (float)(s / N);
}

// np.histogram(values, bins=edges): bin i = [e_i, e_{i+1}), the last bin closed; values outside are dropped.
// Per-CTA integer counts in shared memory (integer atomics: order-independent), then one atomicAdd per bin.
__global__ void __launch_bounds__(256) histogram_kernel(const float* __restrict__ v, int64_t n, const double* __restrict__ edges,
                                                        int nbins, unsigned long long* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char hs_raw[];
  double* e = reinterpret_cast<double*>(hs_raw);                       // [nbins + 1]
  unsigned int* h = reinterpret_cast<unsigned int*>(e + nbins + 1);    // [nbins]
  for (int i = threadIdx.x; i <= nbins; i += 256) e[i] = edges[i];
  for (int i = threadIdx.x; i < nbins; i += 256) h[i] = 0u;
  __syncthreads();
  const double lo = e[0], hi = e[nbins];
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const double x = (double)v[i];
    if (!(x >= lo && x <= hi)) continue;                               // also drops NaN
    int a = 0, b = nbins;                                              // largest a with e[a] <= x  (a < nbins)
    while (b - a > 1) { const int m = (a + b) >> 1; if (e[m] <= x) a = m; else b = m; }
    atomicAdd(&h[a], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += 256)
    if (h[i]) atomicAdd(&counts[i], (unsigned long long)h[i]);
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

// number of ints the caller-provided CSR table needs: (H/2 + 1) bin offsets + at most H*H pixel indices
size_t pub_psd_table_ints(int H) { return (size_t)(H / 2 + 1) + (size_t)H * H; }

// fills the host table: pixels (row-major index ky*H + kx of the FFT output) grouped by radial bin
// [b + 0.5, b + 1.5), b < H/2 (last edge closed), wavenumbers k = fftfreq(H) * H as in results.ipynb cell 4
int pub_psd_build_table(int H, int* table_host) {
  PUB_REQUIRE(table_host && H >= 8 && (H & (H - 1)) == 0, "pub_psd_build_table: H must be a power of two >= 8");
  const int nb = H / 2;
  int* ptr = table_host;
  int* pix = table_host + nb + 1;
  int n = 0;
  for (int b = 0; b < nb; ++b) {
    ptr[b] = n;
    const double lo = b + 0.5, hi = b + 1.5;
    for (int i = 0; i < H; ++i) {
      const int ki = i < H / 2 ? i : i - H;
      for (int j = 0; j < H; ++j) {
        const int kj = j < H / 2 ? j : j - H;
        // the reference takes sqrt in float32 (torch.sqrt of a float32 grid) and bins it as float64
        const double kr = (double)sqrtf((float)(ki * ki + kj * kj));
        const bool in = b == nb - 1 ? (kr >= lo && kr <= hi) : (kr >= lo && kr < hi);
        if (in) pix[n++] = i * H + j;
      }
    }
  }
  ptr[nb] = n;
  return n;
}

// data [N, C, H, H] f32 on the device -> psd_fields [N, C, H/2] and psd_mean [C, H/2] (mean over N)
int pub_radial_psd(const float* data, int N, int C, int H, int transfo, int units, const int* table_dev,
                   float* psd_fields, float* psd_mean, pub_stream_t s_) {
  PUB_REQUIRE(data && table_dev && psd_fields, "pub_radial_psd: null argument");
  PUB_REQUIRE(H >= 8 && H <= 128 && (H & (H - 1)) == 0, "pub_radial_psd: square fields with H a power of two in [8, 128] "
              "(one field lives in shared memory as complex f32)");
  PUB_REQUIRE(!(transfo || units) || C == 3, "pub_radial_psd: the variable transforms need the 3 ClimEx variables");
  cudaStream_t s = (cudaStream_t)s_;
  int lg = 0;
  while ((1 << lg) < H) ++lg;
  const size_t smem = (size_t)H * H * sizeof(float2) + (size_t)(H / 2) * sizeof(float2) + (size_t)(H / 2) * 4 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    PUB_CUDA(cudaFuncSetAttribute(psd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    attr = true;
  }
  const int nb = H / 2;
  psd_kernel<<<dim3(C, N), PSD_NT, smem, s>>>(data, C, H, lg, transfo, units, table_dev, table_dev + nb + 1, psd_fields);
  PUB_LAUNCH_CHECK();
  if (psd_mean) {
    colmean_kernel<<<cdiv((int64_t)C * nb * 32, 256), 256, 0, s>>>(psd_fields, N, C * nb, psd_mean);
    PUB_LAUNCH_CHECK();
  }
  return 0;
}

// counts [nbins] (uint64) += histogram of values over edges [nbins + 1] (f64, ascending); caller zeroes counts
int pub_histogram(const float* values, int64_t n, const double* edges_dev, int nbins, unsigned long long* counts,
                  pub_stream_t s_) {
  PUB_REQUIRE(values && edges_dev && counts && nbins >= 1 && nbins <= 4096, "pub_histogram: bad arguments");
  int grid = (int)((n + 255) / 256);
  const int cap = num_sms() * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  const size_t smem = (size_t)(nbins + 1) * sizeof(double) + (size_t)nbins * sizeof(unsigned int);
  histogram_kernel<<<grid, 256, smem, (cudaStream_t)s_>>>(values, n, edges_dev, nbins, counts);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
