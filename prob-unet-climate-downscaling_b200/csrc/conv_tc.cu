// tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (bf16 in, f32 accumulate).
//
//  conv_tc   : y[p][n] = sum_{tap,c} x[p+tap][c] * w[tap][n][c]   (3x3 / 1x1, stride 1, same pad)
//              M = 128 output pixels (a TW x TH x TB patch), N = BN output channels, K = taps*Cin.
//              A tiles are fetched by 4-D *tiled* TMA loads with the tap offset folded into the box
//              coordinates -- out-of-bounds rows/cols are zero-filled by the TMA unit, which IS the
//              same-padding; no im2col buffer exists.  B tiles ([tap][N][K], K-major) by 3-D TMA.
//              Also runs the data gradient (weights packed transposed+mirrored) and the virtual
//              channel concat (two A tensor maps, src/networks.py:329).
//  wgrad_tc  : dW[tap][co][ci] = sum_p dy[p][co] * x[p+tap][ci]; K = pixels, both operands MN-major
//              (channels contiguous) straight from the same NHWC TMA boxes; 9 tap accumulators live
//              in TMEM side by side; split-K partials are reduced in a fixed order (deterministic).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2..5 =
// epilogue (tcgen05.ld -> registers -> global).  mbarrier ring between producer and issuer,
// tcgen05.commit releases stages and publishes the accumulator.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace pub {

namespace {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("probunet_b200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, kind::tf32: A, B are 32-bit floats whose low 13 mantissa bits are ignored (K = 8 per instruction)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int ES>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (ES == 2) umma_bf16(d, a, b, idesc, acc);
  else umma_tf32(d, a, b, idesc, acc);
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// 32 lanes x 32 consecutive f32 columns -> 32 registers per thread (thread t <-> lane base+t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 consecutive f32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ descriptors
// smem matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), majors (15,16), N>>3 (17..22),
// M>>4 (24..28)
// fmt: 1 = BF16, 2 = TF32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int NTHREADS = 192;
constexpr int MAX_STAGES = 8;

struct TcConvArgs {
  int c0, c1;                  // channels of source 0 / 1
  int taps, ks;                // 1 or 9
  int B, H, W;
  int TW, TH, TB;              // 128-pixel patch
  int tiles_x, tiles_y;        // patches per image row / column
  int m_tiles;                 // total number of 128-pixel patches
  int BN;                      // output channels per CTA (multiple of 32, <= 256)
  int cout;                    // total output channels
  int stages;
  int resident;                // 1: the CTA's whole weight slab [taps*Cin/KC][BN] stays in smem for all its tiles
  const float* bias;
  const void* res; int ld_res;
  const void* mask; int ld_mask;
  void* y; int ldy;
  int relu;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One lane of a CONVERGED warp (the CUTLASS idiom): the whole warp runs the issuing loop, so every descriptor is
// computed once on the uniform datapath, and only the tcgen05 instruction itself is predicated.  Issuing from inside
// `if (lane == 0)` instead makes the compiler treat all operands as per-thread values: ~60 dependent instructions
// (R2UR, ELECT/BRA.U.ANY loops) per tap, measured at ~200 cycles per MMA -- the single issuing thread, not the
// tensor pipe, was the bottleneck.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ forward / dgrad kernel
// ROWB: bytes per smem row (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B); ES: operand element size
// (2 = bf16 / kind::f16, 4 = tf32-rounded f32 / kind::tf32).  Channels per K block = ROWB / ES.
//
// PERSISTENT: grid.x CTAs stride over the 128-pixel patches of one N tile (blockIdx.y).  The accumulator is
// double-buffered in TMEM (2 x BN columns) so the epilogue of patch i overlaps the MMAs of patch i+1, and the
// per-CTA set-up (TMEM alloc, barrier init, descriptor prefetch) is paid once.  When the weight slab of the
// N tile fits (<= 80 KB) it is loaded once and stays resident; the ring then carries activation tiles only.
template <int ROWB_, int ES>
__global__ void __launch_bounds__(NTHREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                           const __grid_constant__ CUtensorMap tmA1,
                                                           const __grid_constant__ CUtensorMap tmW, TcConvArgs a) {
  typedef typename std::conditional<ES == 2, bf16, float>::type T;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t ROWB = ROWB_;
  constexpr int KC = ROWB_ / ES;                          // channels per K block
  constexpr uint32_t LAYOUT = (ROWB_ == 128) ? 2u : 4u;   // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t SBO = 8 * ROWB;                      // 8-row core-matrix group pitch
  constexpr uint32_t A_BYTES = 128 * ROWB;
  const uint32_t B_BYTES = (uint32_t)a.BN * ROWB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk0 = a.c0 / KC, cblk = (a.c0 + a.c1) / KC;
  const int num_kb = a.taps * cblk;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wres = base;                                            // resident weights (if any)
  const uint32_t ring = base + (a.resident ? (uint32_t)num_kb * B_BYTES : 0u);
  const uint32_t STAGE_BYTES = A_BYTES + (a.resident ? 0u : B_BYTES);

  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 5];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]);
  const uint32_t accfull0 = smem_u32(&bars[2 * MAX_STAGES]), accempty0 = smem_u32(&bars[2 * MAX_STAGES + 2]);
  const uint32_t wbar = smem_u32(&bars[2 * MAX_STAGES + 4]);

  uint32_t ncols = 32;
  while ((int)ncols < 2 * a.BN) ncols <<= 1;
  const int n0 = blockIdx.y * a.BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    mbar_init(accfull0, 1); mbar_init(accfull0 + 8, 1);
    mbar_init(accempty0, 4); mbar_init(accempty0 + 8, 4);
    mbar_init(wbar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmW);
    if (a.c1) prefetch_tmap(&tmA1);
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_enter();   // everything above overlapped the previous kernel's tail; no global memory touched yet

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- TMA producer
      const int half = a.ks / 2;
      if (a.resident) {
        mbar_expect_tx(wbar, (uint32_t)num_kb * B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_3d(wres + kb * B_BYTES, &tmW, wbar, (kb % cblk) * KC, n0, kb / cblk);
      }
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % a.tiles_x; t /= a.tiles_x;
        const int ty = t % a.tiles_y; t /= a.tiles_y;
        const int b0 = t * a.TB, y0 = ty * a.TH, x0 = tx * a.TW;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / cblk, cb = kb % cblk;
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = ring + stage * STAGE_BYTES;
          mbar_expect_tx(full0 + 8 * stage, STAGE_BYTES);
          const int dy = tap / a.ks - half, dx = tap % a.ks - half;
          if (cb < cblk0) tma_load_4d(sa, &tmA0, full0 + 8 * stage, cb * KC, x0 + dx, y0 + dy, b0);
          else tma_load_4d(sa, &tmA1, full0 + 8 * stage, (cb - cblk0) * KC, x0 + dx, y0 + dy, b0);
          if (!a.resident) tma_load_3d(sa + A_BYTES, &tmW, full0 + 8 * stage, cb * KC, n0, tap);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer: descriptors are base + byte offset >> 4 (one add per operand)
      const uint32_t idesc = make_idesc(128, a.BN, 0, 0, ES == 2 ? 1u : 2u);
      const uint64_t adesc0 = make_desc(ring, 16, SBO, LAYOUT);
      const uint64_t bdesc0 = make_desc(a.resident ? wres : ring + A_BYTES, 16, SBO, LAYOUT);
      const uint32_t bstep = a.resident ? (B_BYTES >> 4) : 0u;
      const uint32_t sstep = STAGE_BYTES >> 4;
      if (a.resident) mbar_wait(wbar, 0);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(accempty0 + 8 * buf, (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this buffer
        tc_fence_after();
        const uint32_t dcol = tmem_base + (uint32_t)(buf * a.BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)(stage * sstep);
          const uint64_t bd = bdesc0 + (uint64_t)(a.resident ? kb * bstep : stage * sstep);
#pragma unroll
          for (int k = 0; k < (int)ROWB / 32; ++k)  // one MMA consumes 32 B of K; +32 B inside the swizzled row
            umma<ES>(dcol, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit(empty0 + 8 * stage);  // frees the smem stage when these MMAs retire
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(accfull0 + 8 * buf);  // accumulator of this patch complete
      }
    }
  } else {
    // ---------------- epilogue: warps 2..5 own TMEM lane quarters (warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int xl = row % a.TW, yl = (row / a.TW) % a.TH, bl = row / (a.TW * a.TH);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++it) {
      int t = tile;
      const int tx = t % a.tiles_x; t /= a.tiles_x;
      const int ty = t % a.tiles_y; t /= a.tiles_y;
      const int bb = t * a.TB + bl, yy = ty * a.TH + yl, xx = tx * a.TW + xl;
      const bool valid = bb < a.B && yy < a.H && xx < a.W;
      const int64_t pix = ((int64_t)bb * a.H + yy) * a.W + xx;
      const int buf = it & 1;
      mbar_wait(accfull0 + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      for (int cb = 0; cb < a.BN; cb += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * a.BN + cb), r);
        if (cb + 32 >= a.BN) {  // last chunk read: hand the TMEM buffer back to the MMA warp before storing
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
        }
        if (valid) {
          const int n = n0 + cb;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (a.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __ldg(a.bias + n + j);
          }
          if (a.res) {
            const T* rp = (const T*)a.res + pix * a.ld_res + n;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
              Vec8<T>::load(rp + g * 8, f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
            }
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (a.mask) {
            const T* mp = (const T*)a.mask + pix * a.ld_mask + n;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float f[8];
              Vec8<T>::load(mp + g * 8, f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] = f[j] > 0.f ? v[g * 8 + j] : 0.f;
            }
          }
          T* yp = (T*)a.y + pix * a.ldy + n;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = ES == 4 ? round_tf32(v[g * 8 + j]) : v[g * 8 + j];
            Vec8<T>::store(yp + g * 8, f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------ halo forward / dgrad kernel (3x3)
// The per-tap TMA im2col above fetches every activation 9 times from L2 (one box per tap), which bounds all
// layers with N <= 128 by L2 -> smem bandwidth (and by the TMA's rate of 64-byte rows for C = 32).  Here each
// activation is fetched ONCE per tile: one TMA box per K block brings the (16+2) x (8+2) pixel halo of an 8-wide x
// 16-tall output tile into shared memory as ordinary K-major swizzled rows
//      A[halo pixel = hy*10 + hx][ROWB bytes of channels]       (SWIZZLE_128B for 128-byte rows, 64B for 64)
// (out-of-image pixels are zero-filled by the TMA unit = same padding).  The M = 128 operand rows of a tile are 16
// groups of 8 consecutive halo pixels, 10 pixels apart (SBO = 10 rows), and a filter tap (dy, dx) is nothing but a
// start-address offset of ((dy+1)*10 + (dx+1)) rows: the nine taps of a K block are nine descriptors over the same
// bytes.  The swizzle XOR is a function of the absolute smem address (descriptor base-offset field 0 -- setting the
// row phase there gives wrong results, measured), so rows keep their meaning under any row shift.
// Bring-up notes (tools/halo_trace.py, ncu): an unswizzled "plane" layout, where a shift is also just an offset, ran
// at ~200 cycles per MMA; software producers (LDG + STS, then cp.async) were correct but their per-stage
// fence.proxy.async lowers to MEMBAR.ALL.CTA, which waits for the prefetches in flight and serialises the pipeline.
//
// Warp roles (352 threads): 0 = halo + resident-weight TMA, 1 = MMA issuer, 2 = weight-ring TMA, 3..10 = epilogue in
// TWO groups of four warps: group 0 drains TMEM accumulator buffer 0 (even tiles of the CTA), group 1 buffer 1 (odd
// tiles).  The epilogue of a tile is one long dependency chain per warp (TMEM load -> residual / GroupNorm-input loads
// from L2 -> math -> shared-memory transposition for the fused statistics -> stores); with a single group it was the
// critical path as soon as any GroupNorm work was fused into it (profiles/r02_gn_fuse_v1_nopdl.txt).  Two groups give
// every tile two tile periods and put four epilogue warps (two CTAs per SM) on each scheduler.
constexpr int HALO_W = 10, HALO_H = 18, HALO_PX = HALO_W * HALO_H;
constexpr int HTHREADS = 352;

struct HaloArgs {
  int c0, c1;
  int B, H, W;
  int tiles_x, tiles_y, m_tiles;
  int BN, cout;
  int astages, bstages, resident;
  int w_early;        // weights were packed >= 2 launches ago: safe to load before the grid-dependency wait
  const float* bias;
  const void* res; int ld_res;
  const void* mask; int ld_mask;
  void* y; int ldy;
  int relu;
  long long* trace;   // optional event trace of CTA (0,0) (tools/halo_trace.py): [role][1024] clock64 stamps
  // ---- fused GroupNorm work in the epilogue (template parameter EPI of conv_halo_kernel)
  uint32_t epi_off;   // byte offset (from the aligned smem base) of the 8 x [16][33] f32 transposition buffers
  // EPI_STATS: per-(tile, epilogue warp) per-channel (sum, sum of squares) of the values as STORED (rounded to the
  // storage type): rows [(tile * 4 + warp)][cout][2] -- the forward statistics of the GroupNorm that reads y
  float* stat_part;
  // EPI_GNBWD (data-gradient launches): the conv result g = dL/d(activated GroupNorm output) is turned into
  // du = g * keep / (1 - p) * silu'(a x + b) before it is stored, and (sum du, sum du * x) are emitted like above;
  // x = the GroupNorm's input (virtual concat gx0 | gx1), (a, b) = its per-(sample, channel) affine table
  const void* gx0; const void* gx1; int gc0, gld0, gld1;
  const float* gcoef;          // [B][cout][2]
  float p_drop; uint64_t seed, subseq;
  const uint32_t* salt;
};
enum { EPI_PLAIN = 0, EPI_STATS = 1, EPI_GNBWD = 2 };

// role r, event counter n: stamps clock64 into trace[r*1024 + n]
#define HALO_TR(r, n)                                                                   \
  do {                                                                                  \
    if (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && (n) < 1024) a.trace[(r) * 1024 + (n)++] = clock64(); \
  } while (0)

template <int ROWB_, int ES, int EPI>
__global__ void __launch_bounds__(HTHREADS, 2) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                             const __grid_constant__ CUtensorMap tmA1,
                                                             const __grid_constant__ CUtensorMap tmW, HaloArgs a) {
  typedef typename std::conditional<ES == 2, bf16, float>::type T;
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t ROWB = ROWB_;
  constexpr int KC = ROWB_ / ES;                                   // channels per K block
  constexpr uint32_t LAYOUT = (ROWB_ == 128) ? 2u : 4u;            // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t A_TX = HALO_PX * ROWB;                        // bytes one halo box delivers
  constexpr uint32_t A_STAGE = (A_TX + 1023u) & ~1023u;            // 23552 / 12288 B (swizzle pattern aligned)
  const uint32_t B_BYTES = (uint32_t)a.BN * ROWB;

  // warp index through a shuffle: provably warp-uniform, so the role branches and everything computed inside the
  // MMA-issuer warp stay on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int cblk0 = a.c0 / KC, cblk = (a.c0 + a.c1) / KC;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbase = base;   // resident slab [cb][tap] or ring of single (cb, tap) tiles
  const uint32_t abase = base + (uint32_t)(a.resident ? cblk * 9 : a.bstages) * B_BYTES;

  __shared__ __align__(8) uint64_t bars[4 * MAX_STAGES + 5];
  __shared__ uint32_t tmem_slot;
  const uint32_t afull0 = smem_u32(&bars[0]), aempty0 = smem_u32(&bars[MAX_STAGES]);
  const uint32_t bfull0 = smem_u32(&bars[2 * MAX_STAGES]), bempty0 = smem_u32(&bars[3 * MAX_STAGES]);
  const uint32_t accfull0 = smem_u32(&bars[4 * MAX_STAGES]), accempty0 = smem_u32(&bars[4 * MAX_STAGES + 2]);
  const uint32_t wbar = smem_u32(&bars[4 * MAX_STAGES + 4]);

  uint32_t ncols = 32;
  while ((int)ncols < 2 * a.BN) ncols <<= 1;
  const int n0 = blockIdx.y * a.BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1);
      mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1);
    }
    mbar_init(accfull0, 1); mbar_init(accfull0 + 8, 1);
    mbar_init(accempty0, 4); mbar_init(accempty0 + 8, 4);
    mbar_init(wbar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmW);
    if (a.c1) prefetch_tmap(&tmA1);
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // Resident weights are requested BEFORE the grid-dependency wait when the host knows they were packed at least two
  // launches ago (a.w_early; see launch_pdl in common.cuh): the up-to-147 KB slab then lands while the previous
  // kernel is still draining.  Everything else -- activations, residual, mask, the output -- is touched after it.
  auto load_resident = [&]() {
    mbar_expect_tx(wbar, (uint32_t)(cblk * 9) * B_BYTES);
    for (int cb = 0; cb < cblk; ++cb)
      for (int tap = 0; tap < 9; ++tap)
        tma_load_3d(wbase + (uint32_t)(cb * 9 + tap) * B_BYTES, &tmW, wbar, cb * KC, n0, tap);
  };
  if (threadIdx.x == 0 && a.resident && a.w_early) load_resident();
  pdl_enter();

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- resident weights once, then one halo box per (tile, K block)
      if (a.resident && !a.w_early) load_resident();
      int stage = 0; uint32_t phase = 0;
      int trn = 0;
      for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % a.tiles_x; t /= a.tiles_x;
        const int ty = t % a.tiles_y; t /= a.tiles_y;
        for (int cb = 0; cb < cblk; ++cb) {
          mbar_wait(aempty0 + 8 * stage, phase ^ 1);
          HALO_TR(0, trn);
          mbar_expect_tx(afull0 + 8 * stage, A_TX);
          const uint32_t sa = abase + (uint32_t)stage * A_STAGE;
          if (cb < cblk0) tma_load_4d(sa, &tmA0, afull0 + 8 * stage, cb * KC, tx * 8 - 1, ty * 16 - 1, t);
          else tma_load_4d(sa, &tmA1, afull0 + 8 * stage, (cb - cblk0) * KC, tx * 8 - 1, ty * 16 - 1, t);
          if (++stage == a.astages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0 && !a.resident) {
      // ---------------- weight ring: one [BN][KC] tile per (K block, tap), in the order the MMA warp consumes them
      int slot = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x)
        for (int cb = 0; cb < cblk; ++cb)
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bempty0 + 8 * slot, phase ^ 1);
            mbar_expect_tx(bfull0 + 8 * slot, B_BYTES);
            tma_load_3d(wbase + (uint32_t)slot * B_BYTES, &tmW, bfull0 + 8 * slot, cb * KC, n0, tap);
            if (++slot == a.bstages) { slot = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the loop converged, one elected lane issues (see elect_one)
    const uint32_t idesc = make_idesc(128, a.BN, 0, 0, ES == 2 ? 1u : 2u);
    const uint64_t adesc0 = make_desc(abase, 16, HALO_W * ROWB, LAYOUT);
    const uint64_t bdesc0 = make_desc(wbase, 16, 8 * ROWB, LAYOUT);
    const uint32_t bstep = B_BYTES >> 4;
    if (a.resident) mbar_wait(wbar, 0);
    int astage = 0; uint32_t aphase = 0;
    int bslot = 0; uint32_t bphase = 0;
    int it = 0;
    int trn = 0;
    for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(accempty0 + 8 * buf, (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      if (lane == 0) HALO_TR(1, trn);
      const uint32_t dcol = tmem_base + (uint32_t)(buf * a.BN);
      for (int cb = 0; cb < cblk; ++cb) {
        mbar_wait(afull0 + 8 * astage, aphase);
        tc_fence_after();
        if (lane == 0) HALO_TR(1, trn);
        const uint64_t ad_stage = adesc0 + (uint64_t)((uint32_t)astage * (A_STAGE >> 4));
        uint64_t bd_cb = bdesc0 + (uint64_t)((uint32_t)(cb * 9) * bstep);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          uint64_t bd;
          if (a.resident) {
            bd = bd_cb + (uint64_t)((uint32_t)tap * bstep);
          } else {
            mbar_wait(bfull0 + 8 * bslot, bphase);
            tc_fence_after();
            bd = bdesc0 + (uint64_t)((uint32_t)bslot * bstep);
          }
          // tap (dy, dx): the operand starts (dy+1)*10 + (dx+1) rows into the halo tile
          const uint64_t ad = ad_stage + (uint64_t)((uint32_t)((tap / 3) * HALO_W + tap % 3) * (ROWB >> 4));
          const uint32_t first = (uint32_t)(cb | tap);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < (int)ROWB / 32; ++k)   // one MMA consumes 32 B of K; +32 B inside the swizzled row
              umma<ES>(dcol, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (first | (uint32_t)k) != 0u);
            if (!a.resident) umma_commit(bempty0 + 8 * bslot);
          }
          __syncwarp();
          if (!a.resident) { if (++bslot == a.bstages) { bslot = 0; bphase ^= 1; } }
        }
        if (elect_one()) umma_commit(aempty0 + 8 * astage);
        __syncwarp();
        if (lane == 0) HALO_TR(1, trn);
        if (++astage == a.astages) { astage = 0; aphase ^= 1; }
      }
      if (elect_one()) umma_commit(accfull0 + 8 * buf);
      __syncwarp();
    }
  } else {
    // ---------------- epilogue: warps 3..10; TMEM lane quarter = warp % 4, group (accumulator buffer) = (warp - 3) / 4
    // The accumulator is drained in 16-column pieces: 32-column pieces plus the fused GroupNorm arithmetic needed more
    // than the 80 registers two resident CTAs leave each thread, and with ~216 KB of the SM configured as shared memory
    // the spills missed L1 (hit rate 4.5 %) -- ~8 000 cycles per tile in the GroupNorm-backward epilogue (ncu,
    // profiles/r02_gnbwd_epilogue_spills_ncu.txt).
    constexpr int CW = 16;
    const int q = warp & 3;
    const int grp = (warp - 3) >> 2;
    const int row = q * 32 + lane;
    const int xl = row & 7, yl = row >> 3;
    int it = 0;
    int trn = 0;
    const bool trw = warp == 3 && lane == 0;
    float* tsm = nullptr;
    if (EPI != EPI_PLAIN)
      tsm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + a.epi_off) + (grp * 4 + q) * (CW * 33);
    const uint32_t thresh = EPI == EPI_GNBWD ? drop_thresh(a.p_drop) : 0u;
    const uint32_t dkey = EPI == EPI_GNBWD ? (dropout_key(a.seed, a.subseq) ^ (a.salt ? __ldg(a.salt) : 0u)) : 0u;
    const float inv_keep = (EPI == EPI_GNBWD && a.p_drop > 0.f) ? 1.f / (1.f - a.p_drop) : 1.f;
    const int cl = lane & (CW - 1), half = lane >> 4;      // column sums: lane (cl, half) adds half a column
    for (int tile = blockIdx.x; tile < a.m_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      if (buf != grp) continue;                        // the other group's tile
      int t = tile;
      const int tx = t % a.tiles_x; t /= a.tiles_x;
      const int ty = t % a.tiles_y; t /= a.tiles_y;
      const int64_t pix = ((int64_t)t * a.H + ty * 16 + yl) * a.W + tx * 8 + xl;
      mbar_wait(accfull0 + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (trw) HALO_TR(2, trn);
      for (int cb = 0; cb < a.BN; cb += CW) {
        float v[CW];
        {
          uint32_t r[CW];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * a.BN + cb), r);
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (cb + CW >= a.BN) {                           // last piece read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty0 + 8 * buf);
        }
        const int n = n0 + cb;
        if (a.bias) {
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (a.res) {
          const T* rp = (const T*)a.res + pix * a.ld_res + n;
#pragma unroll
          for (int g = 0; g < CW / 8; ++g) {
            float f[8];
            Vec8<T>::load(rp + g * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] += f[j];
          }
        }
        if (a.relu) {
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (a.mask) {
          const T* mp = (const T*)a.mask + pix * a.ld_mask + n;
#pragma unroll
          for (int g = 0; g < CW / 8; ++g) {
            float f[8];
            Vec8<T>::load(mp + g * 8, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g * 8 + j] = f[j] > 0.f ? v[g * 8 + j] : 0.f;
          }
        }
        uint4 xr[EPI == EPI_GNBWD ? CW / 8 : 1];          // the GroupNorm input of this pixel, raw bf16 (kept packed)
        if (EPI == EPI_GNBWD) {
          // v = dL/d(dropout(silu(a x + b)))  ->  du = v * keep / (1 - p) * silu'(a x + b)
          // The (a, b) pairs of this piece's channels are fetched by ONE coalesced load (lane = channel) and broadcast
          // through shared memory (a chain of dependent uniform loads per thread costs an L2 round trip each).  The
          // first 2 * CW floats of the warp's transposition buffer hold them until the sums start.
          const T* xp = n < a.gc0 ? (const T*)a.gx0 + pix * a.gld0 + n : (const T*)a.gx1 + pix * a.gld1 + (n - a.gc0);
          float2 ab = make_float2(0.f, 0.f);
          if (lane < CW) ab = __ldg(reinterpret_cast<const float2*>(a.gcoef + ((int64_t)t * a.cout + n + lane) * 2));
#pragma unroll
          for (int g = 0; g < CW / 8; ++g) xr[g] = __ldg(reinterpret_cast<const uint4*>(xp + g * 8));
          if (lane < CW) reinterpret_cast<float2*>(tsm)[lane] = ab;
          __syncwarp();
#pragma unroll
          for (int g = 0; g < CW / 8; ++g) {
            float f[8];
            {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xr[g]);
#pragma unroll
              for (int i = 0; i < 4; ++i) { const float2 q2 = __bfloat1622float2(h2[i]); f[2 * i] = q2.x; f[2 * i + 1] = q2.y; }
            }
            if (a.p_drop > 0.f) {
              bool keep[8];
              dropout_keep8(dkey, pix * a.cout + n + g * 8, thresh, keep);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[g * 8 + j] = keep[j] ? v[g * 8 + j] * inv_keep : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const float4 c4 = reinterpret_cast<const float4*>(tsm)[(g * 8 + j) / 2];      // (a, b) of two channels
              v[g * 8 + j] *= silu_grad_t<T>(fmaf(c4.x, f[j], c4.y));
              v[g * 8 + j + 1] *= silu_grad_t<T>(fmaf(c4.z, f[j + 1], c4.w));
            }
          }
          __syncwarp();                                  // the coefficient slots are reused by the column sums below
        }
        T* yp = (T*)a.y + pix * a.ldy + n;
#pragma unroll
        for (int g = 0; g < CW / 8; ++g) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = ES == 4 ? round_tf32(v[g * 8 + j]) : v[g * 8 + j];
          Vec8<T>::store(yp + g * 8, f);
        }
        if (EPI != EPI_PLAIN) {
          // per-channel sums over the warp's 32 pixels of the values as the consumer will read them (rounded to the
          // storage type): every lane writes its CW values as a column of a [CW][33] f32 buffer (conflict-free both
          // ways), then lane (cl, half) adds 16 pixels of channel cl in a fixed order and the halves are combined
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = to_f<T>(from_f<T>(v[j]));
#pragma unroll
          for (int j = 0; j < CW; ++j) tsm[j * 33 + lane] = v[j];
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = tsm[cl * 33 + half * 16 + i];
            s1 += x;
            if (EPI == EPI_STATS) s2 = fmaf(x, x, s2);
          }
          __syncwarp();
          if (EPI == EPI_GNBWD) {
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&xr[g]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 q2 = __bfloat1622float2(h2[i]);
                tsm[(g * 8 + 2 * i) * 33 + lane] = v[g * 8 + 2 * i] * q2.x;
                tsm[(g * 8 + 2 * i + 1) * 33 + lane] = v[g * 8 + 2 * i + 1] * q2.y;
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) s2 += tsm[cl * 33 + half * 16 + i];
            __syncwarp();
          }
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          if (half == 0)
            *reinterpret_cast<float2*>(a.stat_part + ((int64_t)(tile * 4 + q) * a.cout + n + cl) * 2) = make_float2(s1, s2);
        }
      }
      if (trw) HALO_TR(2, trn);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------ weight-gradient kernel
struct TcWgradArgs {
  int c0, c1, cout;
  int taps, ks;
  int B, H, W;
  int TW, TH;             // K tile = TW*TH = 64 pixels of one image
  int tiles_x, tiles_y;   // per image
  int tiles_total, tiles_per_split;
  int stages;
  float* part;            // [split][tap][cout][cin] f32
  int box3;               // 3x3, 8x8 patches: x as one (8+2) x (8+2) halo box, taps as row offsets (see wgrad_tc_kernel)
  int swap;               // box3 with cout <= 64: operand roles swapped, D[(dy, ci)][co] (see wgrad_tc_kernel)
  float* bias_part;       // [split][cout] column sums of dy (bias gradient partials) or nullptr
};

constexpr int WG_P = 64;  // pixels (K) per stage

// ES = 2: bf16 operands, 32-channel groups are 64 B rows (SWIZZLE_64B, 8-row swizzle atoms), 16 pixels per MMA
// ES = 4: tf32 operands, 32-channel groups are 128 B rows; MN-major tf32 only exists with the
//         SWIZZLE_128B_BASE32B layout (32-byte chunks XOR row%4, 4-row atoms; TMA mode 128B_ATOM_32B), 8 pixels per MMA
// box3 mode (3x3, H and W multiples of 8): the K tile is an 8 x 8 pixel patch and x arrives as ONE box of
// (8+2) x (8+2) pixels (its halo); every tap is a start-address offset of dy*10 + dx rows (64 / 128 B each) into the
// same box -- the swizzle XOR is a function of the absolute smem address, so a row shift keeps every byte where the
// MMA expects it.  K groups (8 pixels = one patch row) are 10 rows apart (SBO); the three vertical taps are the
// N groups of ONE MMA, also 10 rows apart (LBO = SBO, overlapping reads): 3 MMAs of N = 96 per K step instead of 9
// tap boxes.  The producer is bound by the TMA's rate of rows (~4.7 cycles per 64-byte row): 164 rows per K tile
// (100 halo + 64 dy) instead of 640.
constexpr int WG_BOXROWS = 100;   // (8 + 2) x (8 + 2) pixels

template <int ES>
__global__ void __launch_bounds__(NTHREADS) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX0,
                                                            const __grid_constant__ CUtensorMap tmX1,
                                                            const __grid_constant__ CUtensorMap tmDY,
                                                            TcWgradArgs a) {
  constexpr uint32_t ROW = 32 * ES;                      // bytes of one 32-channel row
  constexpr uint32_t WG_GROUP_BYTES = WG_P * ROW;        // one 32-channel x 64-pixel box
  constexpr uint32_t WG_BOX_BYTES = (WG_BOXROWS * ROW + 1023u) & ~1023u;   // one dy box (box3 mode), pattern aligned
  constexpr uint32_t WG_A_BYTES = 4 * WG_GROUP_BYTES;
  constexpr uint32_t LAYOUT = ES == 2 ? 4u : 1u;         // SWIZZLE_64B : SWIZZLE_128B_BASE32B
  constexpr uint32_t SBO_WG = ES == 2 ? 8 * ROW : 4 * ROW;  // pitch between swizzle atoms along K (8 rows / 4 rows)
  constexpr int KROWS = 32 / ES;                         // pixels per MMA (UMMA_K)
  constexpr uint32_t KSTEP_BYTES = KROWS * ROW;          // = 1024 for both element sizes
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[MAX_STAGES]),
                 accbar = smem_u32(&bars[2 * MAX_STAGES]);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const bool box3 = a.box3 != 0;
  const uint32_t B_BYTES = box3 ? WG_BOX_BYTES : (uint32_t)a.taps * WG_GROUP_BYTES;
  const uint32_t STAGE_BYTES = WG_A_BYTES + B_BYTES;
  uint32_t ncols = 32;
  while ((int)ncols < (a.swap ? 3 * a.cout : a.taps * 32)) ncols <<= 1;

  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 128, split = blockIdx.z;
  const int cin = a.c0 + a.c1;
  const int ngroups = min(4, (a.cout - co0) / 32);  // real 32-channel groups of dy in this co block
  const int t_beg = split * a.tiles_per_split;
  const int t_end = min(a.tiles_total, t_beg + a.tiles_per_split);

  // bias gradient: the CTAs of the first ci block also sum the dy tiles they stage (the four epilogue warps are idle
  // during the main loop); a stage is then released by the MMA commit AND all 128 threads of those four warps.
  // Every LANE arrives, after its own loads: the lanes of a warp can leave the polling loop of mbar_wait in different
  // iterations, ptxas places no reconvergence point after it (and drops a __syncwarp() there), so a lane-0-only
  // arrive released the stage while other lanes had not read it yet -- seen as run-to-run differences of one warp's
  // 16 bias channels when another grid kept the SMs busy (profiles/r02_wgrad_rows128.txt, tools/wgrad_race.py).
  const bool do_bias = a.bias_part != nullptr && blockIdx.x == 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < a.stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, do_bias ? 1 + 4 * 32 : 1); }
    mbar_init(accbar, 1);
    fence_barrier_init();
    prefetch_tmap(&tmX0);
    prefetch_tmap(&tmDY);
    if (a.c1) prefetch_tmap(&tmX1);
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_enter();   // prologue overlapped the previous kernel's tail; no global memory touched yet

  if (warp == 0) {
    if (lane == 0) {
      const int half = a.ks / 2;
      const bool src1 = ci0 >= a.c0;
      const CUtensorMap* tmX = src1 ? &tmX1 : &tmX0;
      const int cx = src1 ? ci0 - a.c0 : ci0;
      int stage = 0; uint32_t phase = 0;
      for (int t = t_beg; t < t_end; ++t) {
        int r = t;
        const int tx = r % a.tiles_x; r /= a.tiles_x;
        const int ty = r % a.tiles_y; r /= a.tiles_y;
        const int b = r, y0 = ty * a.TH, x0 = tx * a.TW;
        mbar_wait(empty0 + 8 * stage, phase ^ 1);
        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + WG_A_BYTES;
        if (box3) {
          mbar_expect_tx(full0 + 8 * stage, (uint32_t)ngroups * WG_GROUP_BYTES + (uint32_t)WG_BOXROWS * ROW);
          for (int g = 0; g < ngroups; ++g)
            tma_load_4d(sa + g * WG_GROUP_BYTES, &tmDY, full0 + 8 * stage, co0 + g * 32, x0, y0, b);
          tma_load_4d(sb, tmX, full0 + 8 * stage, cx, x0 - 1, y0 - 1, b);   // the (8+2) x (8+2) halo of the patch
        } else {
          mbar_expect_tx(full0 + 8 * stage, (uint32_t)(ngroups + a.taps) * WG_GROUP_BYTES);
          for (int g = 0; g < ngroups; ++g)
            tma_load_4d(sa + g * WG_GROUP_BYTES, &tmDY, full0 + 8 * stage, co0 + g * 32, x0, y0, b);
          for (int tap = 0; tap < a.taps; ++tap)
            tma_load_4d(sb + tap * WG_GROUP_BYTES, tmX, full0 + 8 * stage, cx, x0 + tap % a.ks - half,
                        y0 + tap / a.ks - half, b);
        }
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // A = dy  [K=64 px][M=128 co]  MN-major: 4 groups of 32 channels, LBO = group pitch
    // B = x   [K=64 px][N = taps x 32 ci]  MN-major.  9-box mode: the tap tiles lie WG_GROUP_BYTES apart in smem, which
    //         is exactly the LBO stride between 32-element N groups -> all taps of a K step are ONE wide-N MMA (two
    //         for 3x3: N = 160 + 128).  box3 mode: per dx one MMA of N = 96 whose N groups are the three dy boxes.
    // The whole warp runs the loop converged; one elected lane issues (see elect_one).
    const uint32_t fmt = ES == 2 ? 1u : 2u;
    const int n1 = a.taps > 5 ? 5 : a.taps, n2 = a.taps - n1;  // taps per MMA (9-box mode)
    const uint32_t idesc1 = make_idesc(128, n1 * 32, 1, 1, fmt);
    const uint32_t idesc2 = n2 ? make_idesc(128, n2 * 32, 1, 1, fmt) : 0u;
    const uint32_t idesc3 = make_idesc(128, 96, 1, 1, fmt);
    const uint32_t idesc_sw = make_idesc(128, a.cout, 1, 1, fmt);
    const uint64_t adesc0 = make_desc(base, WG_GROUP_BYTES, SBO_WG, LAYOUT);
    const uint64_t bdesc0 = make_desc(base + WG_A_BYTES, WG_GROUP_BYTES, SBO_WG, LAYOUT);
    // box3: K groups (patch rows) and N groups (vertical taps) are both 10 rows apart; in the tf32 layout an atom holds
    // 4 rows, i.e. half a patch row.  The descriptor starts at halo row 1 (tap dy = -1, dx = 0); dx adds -1/0/+1 rows.
    const uint64_t bdesc3 = make_desc(base + WG_A_BYTES + ROW, 10 * ROW, ES == 2 ? 10 * ROW : SBO_WG, LAYOUT);
    const uint32_t sstep = STAGE_BYTES >> 4;
    int stage = 0; uint32_t phase = 0;
    uint32_t first = 1;
    for (int t = t_beg; t < t_end; ++t) {
      mbar_wait(full0 + 8 * stage, phase);
      tc_fence_after();
      const uint64_t ad = adesc0 + (uint64_t)((uint32_t)stage * sstep);
      if (elect_one()) {
        if (a.swap) {
          // Cout <= 64: with dy as the M operand three quarters (half) of every M = 128 instruction are padding, and an
          // MMA costs max(M, 128) * N / 256 tensor cycles whatever M is (ncu: tensor pipe 63 % busy at 32 -> 32).
          // Swapped: A = the x halo box (M = 4 groups of 32 input channels: the three vertical taps + one group of
          // padding rows nobody reads back), B = dy (N = Cout): N / 2 = 16 or 32 cycles per MMA instead of 48.
          // D[(dy, ci)][co] per dx at TMEM column dx * Cout; the descriptors are the same two, exchanged.
          const uint64_t xd = bdesc3 + (uint64_t)((uint32_t)stage * sstep);
#pragma unroll
          for (int k = 0; k < WG_P / KROWS; ++k) {
            const uint32_t acc = (!first || k != 0) ? 1u : 0u;
            const uint64_t ka = (uint64_t)(k * (KSTEP_BYTES >> 4));
            const uint64_t kb = (uint64_t)(k * ((KROWS / 8) * 10 * ROW >> 4));
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              umma<ES>(tmem_base + (uint32_t)(dx * a.cout), xd + kb + (uint64_t)((dx - 1) * (int)(ROW >> 4)), ad + ka, idesc_sw, acc);
          }
        } else if (box3) {
          const uint64_t bd = bdesc3 + (uint64_t)((uint32_t)stage * sstep);
#pragma unroll
          for (int k = 0; k < WG_P / KROWS; ++k) {
            const uint32_t acc = (!first || k != 0) ? 1u : 0u;
            const uint64_t ka = (uint64_t)(k * (KSTEP_BYTES >> 4));
            const uint64_t kb = (uint64_t)(k * ((KROWS / 8) * 10 * ROW >> 4));   // KROWS/8 patch rows per K step
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)   // columns [dx*96, dx*96+96) = taps (dy = 0..2, dx)
              umma<ES>(tmem_base + (uint32_t)(dx * 96), ad + ka, bd + kb + (uint64_t)((dx - 1) * (int)(ROW >> 4)), idesc3, acc);
          }
        } else {
          const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)stage * sstep);
#pragma unroll
          for (int k = 0; k < WG_P / KROWS; ++k) {
            const uint32_t acc = (!first || k != 0) ? 1u : 0u;
            const uint64_t ko = (uint64_t)(k * (KSTEP_BYTES >> 4));
            umma<ES>(tmem_base, ad + ko, bd + ko, idesc1, acc);
            if (n2) umma<ES>(tmem_base + (uint32_t)(n1 * 32), ad + ko, bd + ko + (uint64_t)((n1 * WG_GROUP_BYTES) >> 4), idesc2, acc);
          }
        }
        umma_commit(empty0 + 8 * stage);
      }
      __syncwarp();
      first = 0;
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(accbar);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    const bool has_work = t_end > t_beg;
    if (do_bias) {
      // column sums of the staged dy tiles (group q of this warp): a lane owns one LOGICAL 16-byte chunk of the row
      // (8 bf16 / 4 f32 channels) and walks the rows; its physical position follows the TMA swizzle (bf16: 16-byte
      // chunk ^= (row >> 1) & 3; tf32: 32-byte chunk ^= row & 3).  16-byte loads: half the shared-memory wavefronts
      // of scalar loads -- they compete with the tensor core's operand reads.
      constexpr int CPRW = ROW / 16;            // 16-byte chunks per row: 4 / 8
      constexpr int RPI = 32 / CPRW;            // rows per warp instruction: 8 / 4
      constexpr int NV = 16 / ES;               // channels per chunk: 8 / 4
      const int c16 = lane % CPRW, r0 = lane / CPRW;
      float bs[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) bs[j] = 0.f;
      int stage = 0; uint32_t phase = 0;
      for (int t = t_beg; t < t_end; ++t) {
        mbar_wait(full0 + 8 * stage, phase);
        if (q < ngroups) {
          const uint32_t sg = base + (uint32_t)stage * STAGE_BYTES + (uint32_t)q * WG_GROUP_BYTES;
#pragma unroll
          for (int k = 0; k < WG_P / RPI; ++k) {
            const uint32_t r = (uint32_t)(r0 + k * RPI);
            const uint32_t pos = ES == 2 ? (((uint32_t)c16 ^ ((r >> 1) & 3u)) << 4)
                                         : (((((uint32_t)c16 >> 1) ^ (r & 3u)) << 5) | (((uint32_t)c16 & 1u) << 4));
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(sg + r * ROW + pos));
            if (ES == 2) {
              bs[0] += __uint_as_float(w0 << 16); bs[1] += __uint_as_float(w0 & 0xFFFF0000u);
              bs[2] += __uint_as_float(w1 << 16); bs[3] += __uint_as_float(w1 & 0xFFFF0000u);
              bs[4 % NV] += __uint_as_float(w2 << 16); bs[5 % NV] += __uint_as_float(w2 & 0xFFFF0000u);
              bs[6 % NV] += __uint_as_float(w3 << 16); bs[7 % NV] += __uint_as_float(w3 & 0xFFFF0000u);
            } else {
              bs[0] += __uint_as_float(w0); bs[1] += __uint_as_float(w1); bs[2] += __uint_as_float(w2); bs[3] += __uint_as_float(w3);
            }
          }
        }
        mbar_arrive(empty0 + 8 * stage);     // EVERY lane, after its own loads: see the note at do_bias
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      // rows were dealt over lane / CPRW: fixed xor tree over those lane bits, then lanes 0 .. CPRW-1 write their chunk
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int o = CPRW; o < 32; o <<= 1) bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], o);
      if (lane < CPRW) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const int cc = co0 + q * 32 + c16 * NV + j;
          if (cc < a.cout) a.bias_part[(int64_t)split * a.cout + cc] = bs[j];
        }
      }
    }
    mbar_wait(accbar, 0);
    tc_fence_after();
    if (a.swap) {
      // TMEM lane = (vertical tap q, input channel ci0 + lane); columns dx * Cout + co.  For a fixed co the 32 lanes
      // of a warp write 32 consecutive ci: coalesced 128-byte stores.  Lane quarter 3 holds the padding group.
      if (q < 3) {
        const int ci = ci0 + lane;
        for (int dx = 0; dx < 3; ++dx) {
          const int tap = q * 3 + dx;
          for (int cb = 0; cb < a.cout; cb += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * a.cout + cb), r);
            float* o = a.part + (((int64_t)split * a.taps + tap) * a.cout + cb) * cin + ci;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(int64_t)j * cin] = has_work ? __uint_as_float(r[j]) : 0.f;
          }
        }
      }
    } else {
    for (int tap = 0; tap < a.taps; ++tap) {
      uint32_t r[32];
      // box3 keeps the accumulators ordered [dx][dy]
      const int col = box3 ? ((tap % 3) * 3 + tap / 3) * 32 : tap * 32;
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, r);
      if (co < a.cout) {
        float* o = a.part + (((int64_t)split * a.taps + tap) * a.cout + co) * cin + ci0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = has_work ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(o + j) = v;
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------ host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// NHWC activation view [B][H][W][C] (es-byte elements) with pixel stride ld; box = (cbox, tw, th, tb)
int make_act_map(CUtensorMap* tm, const void* ptr, int es, int C, int ld, int B, int H, int W, int cbox, int tw, int th,
                 int tb, CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  PUB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es};
  cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PUB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation es=%d C=%d ld=%d B=%d H=%d W=%d box=%d,%d,%d,%d) -> %d",
              es, C, ld, B, H, W, cbox, tw, th, tb, (int)r);
  return 0;
}

int make_weight_map(CUtensorMap* tm, const void* ptr, int es, int K, int N, int taps, int kbox, int nbox,
                    CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode();
  PUB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)K * es, (cuuint64_t)N * K * es};
  cuuint32_t box[3] = {(cuuint32_t)kbox, (cuuint32_t)nbox, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PUB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights es=%d K=%d N=%d taps=%d box=%d,%d) -> %d", es, K, N, taps,
              kbox, nbox, (int)r);
  return 0;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int pick_bn(int n) {
  for (int bn : {256, 128, 64, 32})
    if (n % bn == 0) return bn;
  return 0;
}

// a `target`-pixel patch (tw x th x tb) that tiles [B][H][W] exactly in y and x
bool pick_patch(int H, int W, int target, int& tw, int& th, int& tb) {
  tw = W < 16 ? W : 16;
  if (W % tw) return false;
  th = target / tw;
  if (th > H) th = H;
  if (H % th) return false;
  tb = target / (tw * th);
  return tw * th * tb == target && tb >= 1;
}

inline int esize(int dtype) { return dtype == PUB_BF16 ? 2 : 4; }

template <typename K>
int set_smem_attr(K kernel, int bytes) {
  PUB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return 0;
}

}  // namespace

namespace {
bool halo_enabled() {
  if (g_opt_conv_halo < 0) { const char* e = getenv("PUB_CONV_HALO"); g_opt_conv_halo = (e && e[0] == '0') ? 0 : 1; }
  return g_opt_conv_halo != 0;
}
// Which 3x3 layers take the halo kernel: every shape it can tile (8 x 16 output tiles) -- with the halo arriving as
// one TMA box it measured faster than the per-tap kernel on all 36 bf16 and 10 tf32 layer shapes of a training step
// (tools/conv_ab.py, B = 64: profiles/r01b_conv_ab_halo_vs_tap.txt; sum over a step 8.7 -> 5.9 ms and 2.8 -> 2.3 ms).
// g_opt_conv_halo = 0 (pub_debug_option "conv_halo") routes everything through the per-tap kernel for A/B runs.
bool conv_halo_ok(const ConvParams& p, int dtype) {
  return halo_enabled() && p.ks == 3 && p.W % 8 == 0 && p.H % 16 == 0 && (dtype == PUB_BF16 || dtype == PUB_TF32);
}

}  // namespace

bool conv_tc_supported(const ConvParams& p, int dtype) {
  if (dtype != PUB_BF16 && dtype != PUB_TF32) return false;
  const int es = esize(dtype), al = 16 / es;
  if (p.ks != 1 && p.ks != 3) return false;
  const int cin = p.c0 + p.c1;
  if (p.c0 % 32 || p.c1 % 32 || cin < 32 || pick_bn(p.cout) == 0) return false;
  if (p.ld0 % al || (p.x1 && p.ld1 % al) || p.ldy % al) return false;
  if (!aligned16(p.x0) || !aligned16(p.x1) || !aligned16(p.w) || !aligned16(p.y)) return false;
  if ((p.res && (p.ld_res % al || !aligned16(p.res))) || (p.mask && (p.ld_mask % al || !aligned16(p.mask))))
    return false;
  if (conv_halo_ok(p, dtype)) return true;      // 8 x 16 halo tiles: any W % 8 == 0, H % 16 == 0
  int tw, th, tb;
  return pick_patch(p.H, p.W, 128, tw, th, tb);
}

int conv_fused_rows(const ConvParams& p, int dtype, int backend) {
  if (backend == PUB_BACKEND_SIMT || dtype != PUB_BF16 || g_opt_gn_fuse == 0) return 0;
  if (!conv_tc_supported(p, dtype) || !conv_halo_ok(p, dtype)) return 0;
  return 4 * (p.W / 8) * (p.H / 16);
}

namespace {
template <int EPI>
void launch_halo(int es, int rowb, dim3 grid, size_t smem, cudaStream_t s, const CUtensorMap& tmA0, const CUtensorMap& tmA1,
                 const CUtensorMap& tmW, const HaloArgs& a) {
  if (es == 4) launch_pdl(conv_halo_kernel<128, 4, EPI_PLAIN>, grid, HTHREADS, smem, s, tmA0, tmA1, tmW, a);
  else if (rowb == 128) launch_pdl(conv_halo_kernel<128, 2, EPI>, grid, HTHREADS, smem, s, tmA0, tmA1, tmW, a);
  else launch_pdl(conv_halo_kernel<64, 2, EPI>, grid, HTHREADS, smem, s, tmA0, tmA1, tmW, a);
}

int conv_halo(const ConvParams& p, int dtype, cudaStream_t s) {
  const int cin = p.c0 + p.c1, es = esize(dtype);
  const int KC = es == 4 ? 32 : ((p.c0 % 64 == 0 && p.c1 % 64 == 0) ? 64 : 32);
  const int rowb = KC * es;
  const CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const int epi = p.stat_part ? (p.gn_bwd ? EPI_GNBWD : EPI_STATS) : EPI_PLAIN;
  PUB_REQUIRE(epi == EPI_PLAIN || es == 2, "conv_halo: the fused GroupNorm epilogue exists for bf16 only");
  PUB_REQUIRE(epi != EPI_GNBWD || (p.gx0 && p.gcoef && p.gc0 % 32 == 0), "conv_halo: incomplete GroupNorm-backward epilogue arguments");
  HaloArgs a{};
  a.c0 = p.c0; a.c1 = p.c1;
  a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_x = p.W / 8; a.tiles_y = p.H / 16; a.m_tiles = p.B * a.tiles_x * a.tiles_y;
  a.BN = pick_bn(p.cout); a.cout = p.cout;
  a.bias = p.bias; a.res = p.res; a.ld_res = p.ld_res; a.mask = p.mask; a.ld_mask = p.ld_mask;
  a.y = p.y; a.ldy = p.ldy; a.relu = p.relu;
  a.trace = g_halo_trace;
  a.w_early = p.w_settled && weights_settled_on(s);
  a.stat_part = p.stat_part;
  a.gx0 = p.gx0; a.gx1 = p.gx1; a.gc0 = p.gx1 ? p.gc0 : p.cout; a.gld0 = p.gld0; a.gld1 = p.gld1;
  a.gcoef = p.gcoef; a.p_drop = p.p_drop; a.seed = p.seed; a.subseq = p.subseq; a.salt = p.salt;
  const int n_tiles = p.cout / a.BN;
  const int cblk = cin / KC;
  const size_t b_bytes = (size_t)a.BN * rowb, a_stage = align_up((size_t)HALO_PX * rowb, 1024);
  const size_t wres_bytes = (size_t)cblk * 9 * b_bytes;
  const size_t epi_bytes = epi == EPI_PLAIN ? 0 : (size_t)8 * 16 * 33 * sizeof(float);   // transposition buffers (8 epilogue warps)
  const bool small_tmem = 2 * a.BN <= 256;            // two CTAs per SM are possible
  const size_t budget2 = 110 * 1024 - 2048 - epi_bytes, budget1 = 222 * 1024 - 2048 - epi_bytes;
  const int min_a = 3;
  bool two;
  if (small_tmem && wres_bytes + min_a * a_stage <= budget2) {
    a.resident = 1; two = true;
    a.astages = (int)((budget2 - wres_bytes) / a_stage);
  } else if (wres_bytes + min_a * a_stage <= budget1) {
    a.resident = 1; two = false;
    a.astages = (int)((budget1 - wres_bytes) / a_stage);
  } else {
    a.resident = 0; two = false;
    a.astages = 4;
    a.bstages = (int)((budget1 - a.astages * a_stage) / b_bytes);
    if (a.bstages > MAX_STAGES) a.bstages = MAX_STAGES;
    PUB_REQUIRE(a.bstages >= 2, "conv_halo: weight ring does not fit");
  }
  if (a.astages > 6) a.astages = 6;
  const size_t ring_bytes = (a.resident ? wres_bytes : (size_t)a.bstages * b_bytes) + (size_t)a.astages * a_stage;
  a.epi_off = (uint32_t)ring_bytes;
  const size_t smem = 1024 + ring_bytes + epi_bytes;
  int gx = (two ? 2 : 1) * num_sms() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;
  CUtensorMap tmA0, tmA1, tmW;
  PUB_TRY(make_act_map(&tmA0, p.x0, es, p.c0, p.ld0, p.B, p.H, p.W, KC, HALO_W, HALO_H, 1, sw));
  if (p.c1) PUB_TRY(make_act_map(&tmA1, p.x1, es, p.c1, p.ld1, p.B, p.H, p.W, KC, HALO_W, HALO_H, 1, sw));
  else tmA1 = tmA0;
  PUB_TRY(make_weight_map(&tmW, p.w, es, cin, p.cout, 9, KC, a.BN, sw));
  static bool attr = false;
  if (!attr) {
    PUB_TRY(set_smem_attr(conv_halo_kernel<128, 2, EPI_PLAIN>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<64, 2, EPI_PLAIN>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<128, 4, EPI_PLAIN>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<128, 2, EPI_STATS>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<64, 2, EPI_STATS>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<128, 2, EPI_GNBWD>, 225 * 1024));
    PUB_TRY(set_smem_attr(conv_halo_kernel<64, 2, EPI_GNBWD>, 225 * 1024));
    attr = true;
  }
  dim3 grid(gx, n_tiles);
  if (epi == EPI_STATS) launch_halo<EPI_STATS>(es, rowb, grid, smem, s, tmA0, tmA1, tmW, a);
  else if (epi == EPI_GNBWD) launch_halo<EPI_GNBWD>(es, rowb, grid, smem, s, tmA0, tmA1, tmW, a);
  else launch_halo<EPI_PLAIN>(es, rowb, grid, smem, s, tmA0, tmA1, tmW, a);
  PUB_LAUNCH_CHECK();
  return 0;
}
}  // namespace

int conv_tc(const ConvParams& p, int dtype, cudaStream_t s) {
  PUB_REQUIRE(conv_tc_supported(p, dtype), "conv_tc: unsupported shape (c0=%d c1=%d cout=%d H=%d W=%d ks=%d dtype=%d)",
              p.c0, p.c1, p.cout, p.H, p.W, p.ks, dtype);
  if (conv_halo_ok(p, dtype)) return conv_halo(p, dtype, s);
  PUB_REQUIRE(p.stat_part == nullptr, "conv_tc: the fused GroupNorm epilogue needs the halo kernel (check conv_fused_rows first)");
  const int cin = p.c0 + p.c1, es = esize(dtype);
  // channels per K block: bf16 -> 64 (128 B rows) when both sources allow it, else 32 (64 B rows); tf32 -> 32 (128 B rows)
  const int KC = es == 4 ? 32 : ((p.c0 % 64 == 0 && p.c1 % 64 == 0) ? 64 : 32);
  const int rowb = KC * es;
  const CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  TcConvArgs a{};
  a.c0 = p.c0; a.c1 = p.c1; a.ks = p.ks; a.taps = p.ks * p.ks;
  a.B = p.B; a.H = p.H; a.W = p.W;
  pick_patch(p.H, p.W, 128, a.TW, a.TH, a.TB);
  a.tiles_x = p.W / a.TW; a.tiles_y = p.H / a.TH;
  a.BN = pick_bn(p.cout); a.cout = p.cout;
  a.bias = p.bias; a.res = p.res; a.ld_res = p.ld_res;
  a.mask = p.mask; a.ld_mask = p.ld_mask;
  a.y = p.y; a.ldy = p.ldy; a.relu = p.relu;
  a.m_tiles = cdiv(p.B, a.TB) * a.tiles_x * a.tiles_y;
  const int n_tiles = p.cout / a.BN;
  const int num_kb = a.taps * (cin / KC);
  const size_t a_bytes = (size_t)128 * rowb, b_bytes = (size_t)a.BN * rowb;
  const size_t wres_bytes = (size_t)num_kb * b_bytes;
  a.resident = wres_bytes <= 80 * 1024;
  // TMEM: 2 x BN columns (double-buffered accumulator).  <= 256 columns -> two CTAs per SM (more epilogue warps in
  // flight for the HBM-bound small-C layers), so keep each CTA under ~110 KB of smem there.
  const bool two_per_sm = 2 * a.BN <= 256;
  const size_t budget = (two_per_sm ? 110 : 200) * 1024 - 1024;
  if (a.resident && wres_bytes + 2 * a_bytes > budget) a.resident = 0;
  const size_t stage_bytes = a_bytes + (a.resident ? 0 : b_bytes);
  int stages = (int)((budget - (a.resident ? wres_bytes : 0)) / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  a.stages = stages;
  const size_t smem = 1024 + (a.resident ? wres_bytes : 0) + stages * stage_bytes;
  int gx = (two_per_sm ? 2 : 1) * num_sms() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > a.m_tiles) gx = a.m_tiles;

  CUtensorMap tmA0, tmA1, tmW;
  PUB_TRY(make_act_map(&tmA0, p.x0, es, p.c0, p.ld0, p.B, p.H, p.W, KC, a.TW, a.TH, a.TB, sw));
  if (p.c1) PUB_TRY(make_act_map(&tmA1, p.x1, es, p.c1, p.ld1, p.B, p.H, p.W, KC, a.TW, a.TH, a.TB, sw));
  else tmA1 = tmA0;
  PUB_TRY(make_weight_map(&tmW, p.w, es, cin, p.cout, a.taps, KC, a.BN, sw));

  dim3 grid(gx, n_tiles);
  static bool attr = false;
  if (!attr) {
    PUB_TRY(set_smem_attr(conv_tc_kernel<128, 2>, 201 * 1024));
    PUB_TRY(set_smem_attr(conv_tc_kernel<64, 2>, 201 * 1024));
    PUB_TRY(set_smem_attr(conv_tc_kernel<128, 4>, 201 * 1024));
    attr = true;
  }
  if (es == 4) launch_pdl(conv_tc_kernel<128, 4>, grid, NTHREADS, smem, s, tmA0, tmA1, tmW, a);
  else if (rowb == 128) launch_pdl(conv_tc_kernel<128, 2>, grid, NTHREADS, smem, s, tmA0, tmA1, tmW, a);
  else launch_pdl(conv_tc_kernel<64, 2>, grid, NTHREADS, smem, s, tmA0, tmA1, tmW, a);
  PUB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ wgrad host side
namespace {
struct WgPlan { int tw, th, tiles_x, tiles_y, tiles_total, nsplit, tiles_per_split, stages, box3; size_t smem; };
bool wgrad_plan(const WgradParams& p, int es, WgPlan& pl) {
  int tb;
  pl.box3 = g_opt_wgrad_box3 != 0 && p.ks == 3 && p.H % 8 == 0 && p.W % 8 == 0;
  if (pl.box3) { pl.tw = 8; pl.th = 8; }
  else if (!pick_patch(p.H, p.W, WG_P, pl.tw, pl.th, tb) || tb != 1) return false;
  pl.tiles_x = p.W / pl.tw; pl.tiles_y = p.H / pl.th;
  pl.tiles_total = p.B * pl.tiles_x * pl.tiles_y;
  const int cin = p.c0 + p.c1;
  const int ctas = (cin / 32) * cdiv(p.cout, 128);
  const size_t group = (size_t)WG_P * 32 * es;
  const size_t stage = pl.box3 ? 4 * group + align_up((size_t)WG_BOXROWS * 32 * es, 1024)
                               : (4 + (size_t)p.ks * p.ks) * group;
  // two CTAs per SM when three stages fit in half the shared memory (TMEM: 512 columns for 3x3 -> one CTA; the
  // allocation blocks, so co-residency only helps 1x1), otherwise one CTA per SM and a single wave of splits
  pl.stages = (int)((212 * 1024) / stage);
  if (pl.stages > 6) pl.stages = 6;
  pl.smem = pl.stages * stage + 1024;
  // one CTA per SM: the split count is rounded DOWN so that ctas * nsplit <= SMs -- rounding up (e.g. 16 x 10 = 160
  // CTAs on 148 SMs) leaves a second wave of a few CTAs that doubles the kernel time
  int want = num_sms() / ctas;
  if (want < 1) want = 1;
  int max_split = pl.tiles_total / 4;  // at least 4 K tiles per CTA
  if (max_split < 1) max_split = 1;
  if (want > max_split) want = max_split;
  pl.tiles_per_split = cdiv(pl.tiles_total, want);
  pl.nsplit = cdiv(pl.tiles_total, pl.tiles_per_split);
  return pl.stages >= 2;
}
}  // namespace

bool wgrad_tc_supported(const WgradParams& p, int dtype) {
  if (dtype != PUB_BF16 && dtype != PUB_TF32) return false;
  const int es = esize(dtype), al = 16 / es;
  if (p.ks != 1 && p.ks != 3) return false;
  if (p.c0 % 32 || p.c1 % 32 || p.c0 + p.c1 < 32 || p.cout % 32) return false;
  if (p.ld0 % al || (p.x1 && p.ld1 % al) || p.ld_dy % al) return false;
  if (!aligned16(p.x0) || !aligned16(p.x1) || !aligned16(p.dy)) return false;
  WgPlan pl;
  return wgrad_plan(p, es, pl);
}

size_t wgrad_tc_workspace(const WgradParams& p, int dtype) {
  WgPlan pl;
  if (!wgrad_plan(p, esize(dtype), pl)) return 0;
  const int64_t n = (int64_t)p.ks * p.ks * p.cout * (p.c0 + p.c1);
  const int64_t M = (int64_t)p.B * p.H * p.W;
  return align_up((size_t)pl.nsplit * n * sizeof(float), 256) +
         align_up((size_t)std::max(cdiv(M, 1024), pl.nsplit) * p.cout * 4, 256);
}

int wgrad_tc(const WgradParams& p, int dtype, void* ws, size_t ws_bytes, int accumulate, cudaStream_t s) {
  WgPlan pl;
  const int es = esize(dtype);
  PUB_REQUIRE(wgrad_tc_supported(p, dtype) && wgrad_plan(p, es, pl), "wgrad_tc: unsupported shape");
  PUB_REQUIRE(ws_bytes >= wgrad_tc_workspace(p, dtype), "wgrad_tc: workspace too small");
  const int cin = p.c0 + p.c1, taps = p.ks * p.ks;
  TcWgradArgs a{};
  a.c0 = p.c0; a.c1 = p.c1; a.cout = p.cout; a.taps = taps; a.ks = p.ks;
  a.B = p.B; a.H = p.H; a.W = p.W; a.TW = pl.tw; a.TH = pl.th;
  a.tiles_x = pl.tiles_x; a.tiles_y = pl.tiles_y; a.tiles_total = pl.tiles_total;
  a.tiles_per_split = pl.tiles_per_split; a.stages = pl.stages; a.box3 = pl.box3;
  a.swap = (pl.box3 && es == 2 && p.cout <= 64 && g_opt_wgrad_swap) ? 1 : 0;   // bf16 only: the tf32 MN-major atom layout (BASE32B) did not survive the exchange
  a.part = (float*)ws;
  float* bpart = (float*)((char*)ws + align_up((size_t)pl.nsplit * taps * p.cout * cin * sizeof(float), 256));
  a.bias_part = (p.dbias && g_opt_wgrad_fused_bias) ? bpart : nullptr;
  const CUtensorMapSwizzle sw = es == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  CUtensorMap tmX0, tmX1, tmDY;
  const int xbw = pl.box3 ? pl.tw + 2 : pl.tw, xbh = pl.box3 ? pl.th + 2 : pl.th;   // box3: the patch plus its halo
  PUB_TRY(make_act_map(&tmX0, p.x0, es, p.c0, p.ld0, p.B, p.H, p.W, 32, xbw, xbh, 1, sw));
  if (p.c1) PUB_TRY(make_act_map(&tmX1, p.x1, es, p.c1, p.ld1, p.B, p.H, p.W, 32, xbw, xbh, 1, sw));
  else tmX1 = tmX0;
  PUB_TRY(make_act_map(&tmDY, p.dy, es, p.cout, p.ld_dy, p.B, p.H, p.W, 32, pl.tw, pl.th, 1, sw));
  static bool attr = false;
  if (!attr) {
    PUB_TRY(set_smem_attr(wgrad_tc_kernel<2>, 220 * 1024));
    PUB_TRY(set_smem_attr(wgrad_tc_kernel<4>, 220 * 1024));
    attr = true;
  }
  dim3 grid(cin / 32, cdiv(p.cout, 128), pl.nsplit);
  if (es == 2) launch_pdl(wgrad_tc_kernel<2>, grid, NTHREADS, pl.smem, s, tmX0, tmX1, tmDY, a);
  else launch_pdl(wgrad_tc_kernel<4>, grid, NTHREADS, pl.smem, s, tmX0, tmX1, tmDY, a);
  PUB_LAUNCH_CHECK();
  int nchunk = pl.nsplit;     // bias partials: one row per split (summed inside the wgrad kernel) ...
  if (p.dbias && !a.bias_part)  // ... or per 1024-pixel chunk from the separate column-sum pass
    PUB_TRY(colsum(p.dy, p.ld_dy, p.cout, (int64_t)p.B * p.H * p.W, dtype, bpart, nullptr, 0, s, &nchunk));
  return wgrad_finish(a.part, p.dw, pl.nsplit, taps, p.cout, cin, bpart, nchunk, p.dbias, accumulate, s);
}

}  // namespace pub
