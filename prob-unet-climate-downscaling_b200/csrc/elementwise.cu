// Small HBM-bound NHWC kernels: 2x box resample (the kernel=0 "skip" Conv2d of up/down blocks,
// src/networks.py:83-87,159), strided add, MaxPool2d(2)+ReLU backward and the global spatial
// mean of the Gaussian encoders (src/prob_unet.py:39,73).  16/32-byte vector accesses, grid-stride.
#include "common.cuh"
#include "kernels.h"

namespace pub {

namespace {

constexpr int NT = 256;
inline int grid_for(int64_t n) {
  int64_t g = (n + NT - 1) / NT;
  const int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < cap ? g : cap);
}

// mode 1: y[B,H/2,W/2,C] = mean2x2(x[B,H,W,C]);  mode 2: y[B,2H,2W,C] = nearest(x)
template <typename T>
__global__ void resample_kernel(const T* __restrict__ x, int ld, int C, T* __restrict__ y, int B, int H, int W,
                                int mode) {
  pdl_enter();
  const int V = C / 8;
  const int Ho = mode == 1 ? H / 2 : H * 2, Wo = mode == 1 ? W / 2 : W * 2;
  const int64_t total = (int64_t)B * Ho * Wo * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t q = i / V;
    const int xo = (int)(q % Wo), yo = (int)((q / Wo) % Ho), b = (int)(q / ((int64_t)Wo * Ho));
    float o[8];
    if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float f[8];
        Vec8<T>::load(x + (((int64_t)b * H + yo * 2 + (d >> 1)) * W + xo * 2 + (d & 1)) * ld + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
    } else {
      Vec8<T>::load(x + (((int64_t)b * H + yo / 2) * W + xo / 2) * ld + v * 8, o);
    }
    Vec8<T>::store(y + q * C + v * 8, o);
  }
}

// backward of the above: dx at the forward-input resolution [B,H,W,C]
template <typename T>
__global__ void resample_bwd_kernel(const T* __restrict__ dy, int ld, int C, T* __restrict__ dx, int B, int H, int W,
                                    int mode) {
  pdl_enter();
  const int V = C / 8;
  const int64_t total = (int64_t)B * H * W * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t q = i / V;
    const int xx = (int)(q % W), yy = (int)((q / W) % H), b = (int)(q / ((int64_t)W * H));
    float o[8];
    if (mode == 1) {
      Vec8<T>::load(dy + (((int64_t)b * (H / 2) + yy / 2) * (W / 2) + xx / 2) * ld + v * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= 0.25f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        float f[8];
        Vec8<T>::load(dy + (((int64_t)b * (H * 2) + yy * 2 + (d >> 1)) * (W * 2) + xx * 2 + (d & 1)) * ld + v * 8, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += f[j];
      }
    }
    Vec8<T>::store(dx + q * C + v * 8, o);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, int lda, const T* __restrict__ b, int ldb, T* __restrict__ y,
                           int ldy, int C, int64_t M) {
  pdl_enter();
  const int V = C / 8;
  const int64_t total = M * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t m = i / V;
    float fa[8], fb[8];
    Vec8<T>::load(a + m * lda + v * 8, fa);
    Vec8<T>::load(b + m * ldb + v * 8, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    Vec8<T>::store(y + m * ldy + v * 8, fa);
  }
}

// TI / TO differ where the Gaussian encoder's bf16 first stage hands over to its tf32 stages (bf16 values are tf32 values)
template <typename T, typename TO = T>
__global__ void maxpool_kernel(const T* __restrict__ x, int C, TO* __restrict__ y, int B, int H, int W) {
  pdl_enter();
  const int V = C / 8, Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t q = i / V;
    const int xo = (int)(q % Wo), yo = (int)((q / Wo) % Ho), b = (int)(q / ((int64_t)Wo * Ho));
    float o[8];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float f[8];
      Vec8<T>::load(x + (((int64_t)b * H + yo * 2 + (d >> 1)) * W + xo * 2 + (d & 1)) * C + v * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = d == 0 ? f[j] : fmaxf(o[j], f[j]);
    }
    Vec8<TO>::store(y + q * C + v * 8, o);
  }
}

// x is the post-ReLU pre-pool activation.  dx = dy routed to the first maximum of each 2x2 window
// (scan order (0,0),(0,1),(1,0),(1,1) like ATen) and gated by the ReLU (x > 0).
template <typename T, typename TG = T>     // TG: type of the incoming (pooled-resolution) gradient
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const TG* __restrict__ dy, T* __restrict__ dx, int C,
                                   int B, int H, int W) {
  pdl_enter();
  const int V = C / 8, Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t q = i / V;
    const int xo = (int)(q % Wo), yo = (int)((q / Wo) % Ho), b = (int)(q / ((int64_t)Wo * Ho));
    float f[4][8], g[8];
    Vec8<TG>::load(dy + q * C + v * 8, g);
#pragma unroll
    for (int d = 0; d < 4; ++d)
      Vec8<T>::load(x + (((int64_t)b * H + yo * 2 + (d >> 1)) * W + xo * 2 + (d & 1)) * C + v * 8, f[d]);
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int a = 0;
      float m = f[0][j];
#pragma unroll
      for (int d = 1; d < 4; ++d)
        if (f[d][j] > m) { m = f[d][j]; a = d; }
      arg[j] = m > 0.f ? a : -1;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = arg[j] == d ? g[j] : 0.f;
      Vec8<T>::store(dx + (((int64_t)b * H + yo * 2 + (d >> 1)) * W + xo * 2 + (d & 1)) * C + v * 8, o);
    }
  }
}

template <typename T>
__global__ void relu_mask_kernel(T* __restrict__ dy, const T* __restrict__ y, int64_t nvec) {
  pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * NT) {
    float g[8], a[8];
    Vec8<T>::load(dy + i * 8, g);
    Vec8<T>::load(y + i * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
    Vec8<T>::store(dy + i * 8, g);
  }
}

// out[b][c] = mean over HW of x[b][p][c]; one thread per (b, c), fixed order
template <typename T>
__global__ void global_mean_kernel(const T* __restrict__ x, int C, int64_t HW, float* __restrict__ out) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  float s = 0.f;
  const T* p = x + (int64_t)b * HW * C + c;
  for (int64_t r = 0; r < HW; ++r) s += to_f<T>(p[r * C]);
  out[(int64_t)b * C + c] = s / (float)HW;
}

// dx[b][p][c] = (mask[b][p][c] > 0) ? dmean[b][c] / HW : 0
template <typename T>
__global__ void global_mean_bwd_kernel(const float* __restrict__ dmean, const T* __restrict__ mask, int C, int64_t HW,
                                       int64_t total_vec, T* __restrict__ dx) {
  pdl_enter();
  const int V = C / 8;
  const float inv = 1.f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * NT) {
    const int v = (int)(i % V);
    const int64_t pix = i / V;
    const int64_t b = pix / HW;
    float m[8], o[8];
    Vec8<T>::load(mask + pix * C + v * 8, m);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = m[j] > 0.f ? dmean[b * C + v * 8 + j] * inv : 0.f;
    Vec8<T>::store(dx + pix * C + v * 8, o);
  }
}

}  // namespace

int resample2x(const void* x, int ld, int C, void* y, int B, int H, int W, int mode, int dtype, cudaStream_t s) {
  PUB_REQUIRE(C % 8 == 0 && ld % 8 == 0, "resample2x: C and ld must be multiples of 8");
  const int64_t n = (int64_t)B * H * W * (C / 8) * (mode == 1 ? 1 : 4) / (mode == 1 ? 4 : 1);
  if (dtype == PUB_BF16) launch_pdl(resample_kernel<bf16>, grid_for(n), NT, 0, s, (const bf16*)x, ld, C, (bf16*)y, B, H, W, mode);
  else launch_pdl(resample_kernel<float>, grid_for(n), NT, 0, s, (const float*)x, ld, C, (float*)y, B, H, W, mode);
  PUB_LAUNCH_CHECK();
  return 0;
}

int resample2x_bwd(const void* dy, int ld, int C, void* dx, int B, int H, int W, int mode, int dtype, cudaStream_t s) {
  PUB_REQUIRE(C % 8 == 0 && ld % 8 == 0, "resample2x_bwd: C and ld must be multiples of 8");
  const int64_t n = (int64_t)B * H * W * (C / 8);
  if (dtype == PUB_BF16) launch_pdl(resample_bwd_kernel<bf16>, grid_for(n), NT, 0, s, (const bf16*)dy, ld, C, (bf16*)dx, B, H, W, mode);
  else launch_pdl(resample_bwd_kernel<float>, grid_for(n), NT, 0, s, (const float*)dy, ld, C, (float*)dx, B, H, W, mode);
  PUB_LAUNCH_CHECK();
  return 0;
}

int add_views(const void* a, int lda, const void* b, int ldb, void* y, int ldy, int C, int64_t M, int dtype,
              cudaStream_t s) {
  PUB_REQUIRE(C % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldy % 8 == 0, "add_views: strides must be multiples of 8");
  const int64_t n = M * (C / 8);
  if (dtype == PUB_BF16) launch_pdl(add_kernel<bf16>, grid_for(n), NT, 0, s, (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)y, ldy, C, M);
  else launch_pdl(add_kernel<float>, grid_for(n), NT, 0, s, (const float*)a, lda, (const float*)b, ldb, (float*)y, ldy, C, M);
  PUB_LAUNCH_CHECK();
  return 0;
}

int maxpool2(const void* x, int C, void* y, int B, int H, int W, int dtype, cudaStream_t s, int out_dtype) {
  PUB_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2: C %% 8 and even H, W required");
  const int64_t n = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  if (out_dtype >= 0 && (out_dtype == PUB_BF16) != (dtype == PUB_BF16)) {
    PUB_REQUIRE(dtype == PUB_BF16, "maxpool2: the only mixed form is bf16 in, f32 out");
    launch_pdl(maxpool_kernel<bf16, float>, grid_for(n), NT, 0, s, (const bf16*)x, C, (float*)y, B, H, W);
  } else if (dtype == PUB_BF16) launch_pdl(maxpool_kernel<bf16>, grid_for(n), NT, 0, s, (const bf16*)x, C, (bf16*)y, B, H, W);
  else launch_pdl(maxpool_kernel<float>, grid_for(n), NT, 0, s, (const float*)x, C, (float*)y, B, H, W);
  PUB_LAUNCH_CHECK();
  return 0;
}

int maxpool2_bwd(const void* x, const void* /*yp*/, const void* dy, void* dx, int C, int B, int H, int W, int dtype,
                 cudaStream_t s, int dy_dtype) {
  const int64_t n = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  if (dy_dtype >= 0 && (dy_dtype == PUB_BF16) != (dtype == PUB_BF16)) {
    PUB_REQUIRE(dtype == PUB_BF16, "maxpool2_bwd: the only mixed form is bf16 activations / dx with an f32 incoming gradient");
    launch_pdl(maxpool_bwd_kernel<bf16, float>, grid_for(n), NT, 0, s, (const bf16*)x, (const float*)dy, (bf16*)dx, C, B, H, W);
  } else if (dtype == PUB_BF16) launch_pdl(maxpool_bwd_kernel<bf16>, grid_for(n), NT, 0, s, (const bf16*)x, (const bf16*)dy, (bf16*)dx, C, B, H, W);
  else launch_pdl(maxpool_bwd_kernel<float>, grid_for(n), NT, 0, s, (const float*)x, (const float*)dy, (float*)dx, C, B, H, W);
  PUB_LAUNCH_CHECK();
  return 0;
}

int relu_mask_inplace(void* dy, const void* y, int64_t n, int dtype, cudaStream_t s) {
  PUB_REQUIRE(n % 8 == 0, "relu_mask_inplace: n %% 8");
  if (dtype == PUB_BF16) launch_pdl(relu_mask_kernel<bf16>, grid_for(n / 8), NT, 0, s, (bf16*)dy, (const bf16*)y, n / 8);
  else launch_pdl(relu_mask_kernel<float>, grid_for(n / 8), NT, 0, s, (float*)dy, (const float*)y, n / 8);
  PUB_LAUNCH_CHECK();
  return 0;
}

int global_mean(const void* x, int C, int B, int64_t HW, float* out, float* /*partial*/, int dtype, cudaStream_t s) {
  dim3 grid(cdiv(C, 128), B);
  if (dtype == PUB_BF16) launch_pdl(global_mean_kernel<bf16>, grid, 128, 0, s, (const bf16*)x, C, HW, out);
  else launch_pdl(global_mean_kernel<float>, grid, 128, 0, s, (const float*)x, C, HW, out);
  PUB_LAUNCH_CHECK();
  return 0;
}

int global_mean_bwd(const float* dmean, const void* mask, int C, int B, int64_t HW, void* dx, int dtype, cudaStream_t s) {
  const int64_t n = (int64_t)B * HW * (C / 8);
  if (dtype == PUB_BF16) launch_pdl(global_mean_bwd_kernel<bf16>, grid_for(n), NT, 0, s, dmean, (const bf16*)mask, C, HW, n, (bf16*)dx);
  else launch_pdl(global_mean_bwd_kernel<float>, grid_for(n), NT, 0, s, dmean, (const float*)mask, C, HW, n, (float*)dx);
  PUB_LAUNCH_CHECK();
  return 0;
}

int fill_zero(void* p, size_t bytes, cudaStream_t s) {
  PUB_CUDA(cudaMemsetAsync(p, 0, bytes, s));
  return 0;
}

}  // namespace pub
