// GPU side of the reference's dataset transform: climex2torch.__getitem__ for
// type = "lrinterp_to_residuals" (src/climex_utils.py:197-225) and compute_stats (:255-264), batched.
//
//   lr        = AvgPool2d(s)(hr)                              [.,C,H/s,W/s]
//   lrinterp  = nearest-upsample(lr, s)                       [.,C,H,W]
//   (mean, std) of lr over the time axis (unbiased std), expanded to the HR grid (block-constant)
//   inputs    = (lrinterp - mean) / (std + eps)
//   targets   = (hr - mean) / (std + eps) - inputs
//
// The reference does this per sample on the DataLoader's main thread (num_workers = 0): ~185 samples/s on its
// workstation, i.e. about its training speed.  Here one CTA owns one s x s cell of one (sample, variable): a block
// reduction gives the cell mean, every thread then writes its pixel of the three outputs.  HBM-bound: hr is read
// once, three tensors are written once.
#include "common.cuh"
#include "kernels.h"

namespace pub {
namespace {

// sum over the CTA (<= 1024 threads), result broadcast to all threads; fixed shuffle tree + fixed order
__device__ __forceinline__ float cta_sum_bcast(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += sm[i];
  return r;
}

// grid = (cells_x, cells_y, B*C); block = s*s threads (one per pixel of the cell)
__global__ void climex_transform_kernel(const float* __restrict__ hr, const float* __restrict__ mean_lr,
                                        const float* __restrict__ std_lr, int C, int H, int W, int s, float eps,
                                        float* __restrict__ inputs, float* __restrict__ targets,
                                        float* __restrict__ lrinterp, float* __restrict__ lr) {
  __shared__ float sm[32];
  const int cx = blockIdx.x, cy = blockIdx.y, bc = blockIdx.z, c = bc % C;
  const int hs = H / s, ws = W / s;
  const int px = threadIdx.x % s, py = threadIdx.x / s;
  const int64_t o = ((int64_t)bc * H + cy * s + py) * W + cx * s + px;
  const float v = hr[o];
  const float cell = cta_sum_bcast(v, sm) / (float)(s * s);          // AvgPool2d(s)
  const float m = mean_lr[((int64_t)c * hs + cy) * ws + cx], sd = std_lr[((int64_t)c * hs + cy) * ws + cx] + eps;
  const float in = (cell - m) / sd;                                   // lrinterp_stand
  inputs[o] = in;
  targets[o] = (v - m) / sd - in;                                     // hr_stand - lrinterp_stand (same op order)
  if (lrinterp) lrinterp[o] = cell;
  if (lr && threadIdx.x == 0) lr[((int64_t)bc * hs + cy) * ws + cx] = cell;
}

// grid = (cells_x, cells_y, C); the CTA walks the T time steps of its cell: Welford over the cell means
__global__ void climex_stats_kernel(const float* __restrict__ hr, int T, int C, int H, int W, int s,
                                    float* __restrict__ mean_lr, float* __restrict__ std_lr) {
  __shared__ float sm[32];
  const int cx = blockIdx.x, cy = blockIdx.y, c = blockIdx.z;
  const int hs = H / s, ws = W / s;
  const int px = threadIdx.x % s, py = threadIdx.x / s;
  double mean = 0.0, m2 = 0.0;
  for (int t = 0; t < T; ++t) {
    const float v = hr[(((int64_t)t * C + c) * H + cy * s + py) * W + cx * s + px];
    const double cell = (double)(cta_sum_bcast(v, sm) / (float)(s * s));
    const double d = cell - mean;
    mean += d / (double)(t + 1);
    m2 += d * (cell - mean);
  }
  if (threadIdx.x == 0) {
    mean_lr[((int64_t)c * hs + cy) * ws + cx] = (float)mean;
    std_lr[((int64_t)c * hs + cy) * ws + cx] = T > 1 ? (float)sqrt(m2 / (double)(T - 1)) : 0.f;   // torch.std: unbiased
  }
}

int check_shape(int C, int H, int W, int s) {
  PUB_REQUIRE(s >= 1 && s <= 32 && H % s == 0 && W % s == 0 && C >= 1,
              "climex transform: lowres_scale %d must be in [1, 32] and divide %dx%d", s, H, W);
  return 0;
}

}  // namespace
}  // namespace pub

using namespace pub;

extern "C" {

int pub_climex_stats(const float* hr, int T, int C, int H, int W, int lowres_scale, float* mean_lr, float* std_lr,
                     pub_stream_t s_) {
  PUB_REQUIRE(hr && mean_lr && std_lr && T >= 1, "pub_climex_stats: bad arguments");
  PUB_TRY(check_shape(C, H, W, lowres_scale));
  dim3 grid(W / lowres_scale, H / lowres_scale, C);
  climex_stats_kernel<<<grid, lowres_scale * lowres_scale, 0, (cudaStream_t)s_>>>(hr, T, C, H, W, lowres_scale, mean_lr, std_lr);
  PUB_LAUNCH_CHECK();
  return 0;
}

int pub_climex_transform(const float* hr, const float* mean_lr, const float* std_lr, int B, int C, int H, int W,
                         int lowres_scale, float eps, float* inputs, float* targets, float* lrinterp, float* lr,
                         pub_stream_t s_) {
  PUB_REQUIRE(hr && mean_lr && std_lr && inputs && targets && B >= 1, "pub_climex_transform: bad arguments");
  PUB_TRY(check_shape(C, H, W, lowres_scale));
  PUB_REQUIRE((int64_t)B * C <= 65535, "pub_climex_transform: B*C = %lld exceeds the grid z limit", (long long)B * C);
  dim3 grid(W / lowres_scale, H / lowres_scale, B * C);
  climex_transform_kernel<<<grid, lowres_scale * lowres_scale, 0, (cudaStream_t)s_>>>(hr, mean_lr, std_lr, C, H, W, lowres_scale,
                                                                                      eps, inputs, targets, lrinterp, lr);
  PUB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
