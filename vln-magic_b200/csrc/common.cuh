// Shared device/host helpers for libmagic_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define MAGIC_F32 0
#define MAGIC_BF16 1

#define MAGIC_OK 0
#define MAGIC_ERR_ARG 1
#define MAGIC_ERR_CUDA 2
#define MAGIC_ERR_UNSUPPORTED 3

void magic_set_error(const char* fmt, ...);

#define MAGIC_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      magic_set_error(__VA_ARGS__);           \
      return MAGIC_ERR_ARG;                   \
    }                                         \
  } while (0)

#define MAGIC_CHECK_LAUNCH(name)                                                      \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      magic_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));   \
      return MAGIC_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define MAGIC_CUDA(call, name)                                                        \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      magic_set_error("%s: %s", name, cudaGetErrorString(e__));                       \
      return MAGIC_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

static inline int magic_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through magic_launch() may be scheduled while its
// predecessor in the stream is still draining: it calls pdl_trigger() first (so ITS successor can be staged as
// early as possible) and pdl_wait() before its first access to global memory, which blocks until the
// predecessor grid has completed and its writes are visible.  Everything before pdl_wait() (barrier init, TMEM
// allocation, tensor-map prefetch, index math) overlaps the predecessor's tail.  The step is a chain of
// several hundred small dependent kernels, so the launch gap is a first-order cost.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int magic_pdl_enabled();  // api.cu: on unless env MAGIC_PDL=0

template <typename... KArgs, typename... Args>
static inline cudaError_t magic_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = magic_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, with a thread-block cluster of `cluster` CTAs (distributed shared memory between them)
template <typename... KArgs, typename... Args>
static inline cudaError_t magic_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                               cudaStream_t st, dim3 cluster, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster.x;
  attr[0].val.clusterDim.y = cluster.y;
  attr[0].val.clusterDim.z = cluster.z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = magic_pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// element access: activations are f32 or bf16; math is always fp32
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p, size_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void stf(float* p, size_t i, float v) { p[i] = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, size_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// round-trip through the storage type (so fwd and bwd recomputation see identical values)
__device__ __forceinline__ float rt(const float*, float v) { return v; }
__device__ __forceinline__ float rt(const __nv_bfloat16*, float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// ---------------------------------------------------------------------------------------------
// stateless dropout RNG: keep(idx) is a pure function of (seed, salt, idx) so backward regenerates
// the forward mask instead of storing it.
// ---------------------------------------------------------------------------------------------
// 32-bit avalanche mix ("lowbias32"): 2 multiplies + 3 xor-shifts.  The per-element cost matters: the hash runs inside
// GEMM / LayerNorm / attention epilogues (one call per element), where a 64-bit splitmix was ~3x the instructions.
__device__ __forceinline__ uint32_t magic_mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// (seed, salt) -> 32-bit stream key, once per thread
__device__ __forceinline__ uint32_t magic_key(uint64_t seed, uint32_t salt) {
  return magic_mix32((uint32_t)seed ^ magic_mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (salt + 1u)));
}
// element index is taken modulo 2^32 (every tensor on this path has < 2^32 elements)
__device__ __forceinline__ uint32_t magic_hash(uint32_t key, uint64_t idx) { return magic_mix32((uint32_t)idx ^ key); }
struct Dropout {
  float p;           // drop probability (0 = disabled)
  float inv_keep;    // 1 / (1 - p)
  uint32_t thresh;   // drop iff hash < thresh
  uint32_t key;      // magic_key(seed, salt)
  __device__ __forceinline__ float scale(uint64_t idx) const {
    if (p <= 0.f) return 1.f;
    return magic_hash(key, idx) < thresh ? 0.f : inv_keep;
  }
};
// seed_ptr is DEVICE memory (so a captured CUDA graph sees fresh seeds)
__device__ __forceinline__ Dropout make_dropout(float p, const unsigned long long* seed_ptr, uint32_t salt) {
  Dropout d;
  d.p = p;
  d.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.thresh = p > 0.f ? (uint32_t)fminf(4294967295.f, p * 4294967296.f) : 0u;
  d.key = magic_key((p > 0.f && seed_ptr) ? *seed_ptr : 0ull, salt);
  return d;
}

// GELU through the normal tail: Phi(-a) = 2^Q(a) with Q a degree-5 polynomial fitted on a = |x| in [0, 6] (weighted
// minimax of the GELU value, scripts/fit_gelu.py: |abs err| < 5e-7 over all x), so
//   gelu(x) = max(x, 0) - |x| * 2^Q(min(|x|, 6))
// costs 8 FMA-pipe ops and ONE exp2 (an Abramowitz-Stegun 7.1.26 erf: 16 + 2 MUFU, the version this replaced).  A 128 x 256 accumulator slab is 128 elements
// per epilogue thread, so the FFN-up epilogue was longer than its 12-k-block main loop.
__device__ __forceinline__ float gelu_tail(float a) {  // Phi(-a), a >= 0
  a = fminf(a, 6.f);
  float q = fmaf(a, -4.733715079e-04f, 7.084977951e-03f);
  q = fmaf(q, a, -5.182837537e-02f);
  q = fmaf(q, a, -4.599914417e-01f);
  q = fmaf(q, a, -1.150788262e+00f);
  q = fmaf(q, a, -1.000037571e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return e;
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float a = fabsf(x);
  return fmaf(-a, gelu_tail(a), fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_fast_f(float x) {  // Phi(x) + x * phi(x)
  const float t = gelu_tail(fabsf(x));
  float g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(-0.72134752044448170f * x * x));  // exp(-x^2 / 2)
  const float cdf = x >= 0.f ? 1.f - t : t;
  return fmaf(x * 0.3989422804014327f, g, cdf);
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// 16-byte vector access for the vocabulary-sized rows (bf16: 8 elements, fp32: 4 elements)
template <typename T>
struct RowVec;
template <>
struct RowVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  // streaming (evict-first) variant for data that is read exactly once
  static __device__ __forceinline__ void load_cs(const float* p, float (&v)[4]) {
    const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct RowVec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void load_cs(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};

// online log-sum-exp update of (m, s) with value x
__device__ __forceinline__ void lse_push(float& m, float& s, float x) {
  if (x > m) {
    s = s * expf(m - x) + 1.f;  // m = -inf: s is 0 and expf(-inf) = 0
    m = x;
  } else if (x > -INFINITY) {
    s += expf(x - m);
  }
}
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) return;
  s = s * expf(m - mm) + s2 * expf(m2 - mm);
  m = mm;
}

