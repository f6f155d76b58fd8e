// Shared device/host helpers for the probunet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/probunet_b200.h"

namespace pub {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);

#define PUB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::pub::set_error(__VA_ARGS__);      \
      return -1;                          \
    }                                     \
  } while (0)

#define PUB_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::pub::set_error("%s -> %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return -2;                                                                             \
    }                                                                                        \
  } while (0)

extern int g_opt_conv_halo;   // -1: from PUB_CONV_HALO (default on); 0/1 forced by pub_debug_option("conv_halo", v)
extern int g_opt_wgrad_box3;  // 1 (default): 3x3 wgrad loads x as three (8+2) x 8 boxes; 0: nine tap boxes (A/B: pub_debug_option("wgrad_box3", v))
extern int g_opt_fcomb_fwd_mma;  // fcomb forward in bf16 mode: 1 = tensor-core kernel (bf16 operands), 0 = f32 FMA kernel
extern int g_opt_wgrad_fused_bias;  // 1 (default): bias gradients summed from the dy tiles staged by the wgrad kernel; 0: separate pass
extern int g_opt_gn_fuse;  // GroupNorm work fused into the halo conv epilogues: 0 none, 1 forward statistics, 2 + backward prologue
// CUDA-graph support: a device word mixed into every dropout key / rsample seed, advanced once per step by
// pub_advance_counters (a graph replays the same kernel ARGUMENTS every step, so per-step randomness has to come from
// device memory).  nullptr (default): unused, the (seed, offset) arguments alone decide -- what the parity tests run.
extern int g_opt_wgrad_swap;      // 1 (default): 3x3 weight gradients with Cout <= 64 put the x halo on the M side (see wgrad_tc_kernel)
extern int g_opt_metrics_occ;     // CTAs per SM the ensemble metric kernel is compiled for: 3 (168 registers) or 4 (128, some spills)
extern int g_opt_fcomb_bwd_occ;   // CTAs per SM of fcomb_bwd_mma_kernel: 2 (237 registers) or 3 (168)
extern const uint32_t* g_seed_salt;
extern long long* g_halo_trace;  // device buffer for conv_halo_kernel event stamps (pub_debug_pointer("halo_trace", p)), else null
extern unsigned long long g_launch_count;  // kernels enqueued by this library (bench.py's gpu_launches)
#define PUB_LAUNCH_CHECK()          \
  do {                              \
    ++::pub::g_launch_count;        \
    PUB_CUDA(cudaGetLastError());   \
  } while (0)

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Kernels whose first global-memory access comes after pdl_wait() are launched through launch_pdl(): the next kernel
// of the stream may be scheduled once every CTA of this one has passed its own pdl_wait() and called pdl_trigger(),
// so its CTAs occupy SMs as ours drain and its prologue (barrier init, TMEM allocation, tensor-map prefetch,
// resident-weight TMA) overlaps our tail instead of following it.  Trigger-after-wait bounds the look-ahead to one
// kernel: when a kernel starts, everything up to its predecessor's predecessor has completed and is visible.
// Kernels launched the ordinary way (<<< >>>) in between keep full stream serialisation; in them, and with
// g_opt_pdl = 0 (pub_debug_option("pdl", 0)), both instructions are no-ops.
extern int g_opt_pdl;
// Which stream the last weight-pack kernel ran on and how many launch_pdl() launches followed it ON THAT STREAM
// (ordinary <<< >>> launches are not counted: under-counting only delays the early weight loads).  The library is
// driven from one host thread per process (one process per GPU); these are plain globals.
extern cudaStream_t g_pack_stream;
extern unsigned long long g_since_pack;
inline void note_weight_pack(cudaStream_t s) { g_pack_stream = s; g_since_pack = 0; }
// true when weights packed by this library are complete and visible to a kernel that is launched next on `s` even
// BEFORE that kernel's grid-dependency wait: at least two trigger-after-wait kernels separate it from the pack
inline bool weights_settled_on(cudaStream_t s) { return g_opt_pdl && s == g_pack_stream && g_since_pack >= 2; }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_opt_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);   // errors surface in PUB_LAUNCH_CHECK
  if (s == g_pack_stream) ++g_since_pack;
}
#endif

#define PUB_TRY(expr)        \
  do {                       \
    int r__ = (expr);        \
    if (r__ != 0) return r__; \
  } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t dtype_size(int dt) { return dt == PUB_BF16 ? 2 : 4; }
int num_sms();

// ---------------------------------------------------------------- scalar conversion
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels of an NHWC tensor <-> 8 floats (16 B of bf16 / 32 B of f32)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in thread 0 (deterministic: fixed shuffle tree + fixed order)
template <int NT> __device__ __forceinline__ float block_sum(float v, float* smem /* >= NT/32 floats */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) smem[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (l < NT / 32) ? smem[l] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
  static __device__ __forceinline__ uint4 gen(uint64_t seed, uint64_t subseq, uint64_t ctr) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)subseq, c3 = (uint32_t)(subseq >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  static __device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
};

__device__ __forceinline__ float round_tf32_f(float x) {  // round-to-nearest onto the 10-bit-mantissa tf32 grid
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = 1.f / (1.f + __expf(-x));
  return s * (1.f + x * (1.f - s));
}

// ---------------------------------------------------------------- GroupNorm / SiLU / dropout helpers
// (shared by norm.cu and by the conv epilogues that fuse the GroupNorm statistics / backward prologue, conv_tc.cu)
#ifdef __CUDACC__
// MUFU-only sigmoid (ex2 + rcp, no IEEE-division subroutine): rel. error ~1e-6, far inside the 1e-4 parity budget.
// The kernels below are otherwise issue-bound on the division slow path rather than HBM-bound.
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_fast(1.f + __expf(-x)); }
// bf16 storage: one MUFU (tanh.approx, rel. error 2^-11 -- below the 2^-9 of the bf16 value it is multiplied into)
// instead of ex2 + rcp + 2 FP32 ops; f32 storage keeps the exact-to-1e-6 form
__device__ __forceinline__ float sigmoid_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
template <typename T> __device__ __forceinline__ float sigmoid_t(float x) {
  return sizeof(T) == 2 ? sigmoid_tanh(x) : sigmoid_fast(x);
}
template <typename T> __device__ __forceinline__ float silu_t(float x) { return x * sigmoid_t<T>(x); }
template <typename T> __device__ __forceinline__ float silu_grad_t(float x) {
  const float s = sigmoid_t<T>(x);
  return fmaf(x * s, 1.f - s, s);
}

// Dropout mask (replaces F.dropout's generator, src/networks.py:177; distributional parity only -- for exact parity the
// masks are exported, pub_unet_dropout_mask).  Counter-based: 16 random bits per element from a 32-bit integer mixer
// (two multiply-xorshift rounds, the "lowbias32" constants) of (key ^ word index), key = mix of (seed, block
// subsequence); keep <=> u16 >= round(p * 65536).  ~35 instructions per 8 elements -- the Philox4x32-10 stream used
// before cost ~110 and made the fused GroupNorm-backward conv epilogue issue-bound (rsample keeps Philox).
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint64_t subseq) {
  return mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) ^ mix32((uint32_t)subseq * 0x9E3779B9u + 0x85EBCA6Bu)));
}
// keep-mask of 8 consecutive NHWC elements starting at linear element index e (e % 8 == 0)
__device__ __forceinline__ void dropout_keep8(uint32_t key, int64_t e, uint32_t thresh, bool (&keep)[8]) {
  const uint32_t w = (uint32_t)(e >> 1);            // index of the first of four 32-bit words (two elements each)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t r = mix32(key ^ ((w + (uint32_t)q) * 0x9E3779B1u));
    keep[2 * q] = (r & 0xFFFFu) >= thresh; keep[2 * q + 1] = (r >> 16) >= thresh;
  }
}
__device__ __forceinline__ uint32_t drop_thresh(float p) { return (uint32_t)(p * 65536.f + 0.5f); }

#endif

// ---------------------------------------------------------------- bump allocator over a caller workspace
struct Arena {
  char* base = nullptr;
  size_t off = 0, cap = 0;
  bool dry = true;  // dry run: only measure
  Arena() {}
  Arena(void* p, size_t c) : base((char*)p), cap(c), dry(p == nullptr) {}
  void* take(size_t bytes) {
    off = align_up(off, 1024);
    void* r = dry ? nullptr : (void*)(base + off);
    off += bytes;
    return r;
  }
  template <typename U> U* take_n(size_t n) { return (U*)take(n * sizeof(U)); }
  bool ok() const { return dry || off <= cap; }
};

}  // namespace pub
