// Fused flat-arena optimizer kernels: global grad-norm (sum of squares), AdamW with the reference's
// exact update order (pretrain_src/optim/adamw.py:84-110: moments, bias-corrected step, THEN decoupled
// weight decay with lr), gradient clipping folded in (clip coefficient read from device memory), and
// the bf16 shadow copy of the parameters written in the same pass.
// Algorithmic bytes: 16 B/param read (p, g, m, v) + 12 B/param write (p, m, v) [+2 B bf16 shadow].
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

// Sum of squares in TWO deterministic stages: every CTA writes its partial to `partials[blockIdx.x]` (fixed grid,
// fixed per-thread element order), then one CTA adds the partials in index order.  No floating-point atomics: the
// clip coefficient is a function of the gradient values alone, so data-parallel replicas that hold bit-identical
// gradients after the exchange stay bit-identical after the step.
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                            float* __restrict__ partials) {
  __shared__ float red[32];
  float a = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    a += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    a += g[i] * g[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = a;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(const float* __restrict__ partials, int np,
                                                          float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < np; i += blockDim.x) a += partials[i];  // fixed assignment, fixed order
  a = block_sum(a, red);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + a : a;
}

// hyper (device, fp32): [0] lr  [1] step_size = lr*sqrt(1-b2^t)/(1-b1^t)  [2] beta1  [3] beta2  [4] eps
//                        [5] max_grad_norm (<=0: no clipping)
__global__ void __launch_bounds__(256)
    adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 __nv_bfloat16* __restrict__ shadow, long long n, const float* __restrict__ hyper, float weight_decay,
                 const float* __restrict__ sumsq) {
  const float lr = hyper[0], step = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4], maxn = hyper[5];
  float clip = 1.f;
  if (sumsq != nullptr && maxn > 0.f) {
    const float c = maxn / (sqrtf(sumsq[0]) + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    clip = c < 1.f ? c : 1.f;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = m[i] * b1 + (1.f - b1) * gi;
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) + eps;
    float pi = p[i] + (-step) * (mi / denom);
    if (weight_decay > 0.f) pi = pi + (-lr * weight_decay) * pi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

// Same update for an arena whose parameters are not all active in this step (pretrain_src/optim/adamw.py:66-67 skips
// parameters with `p.grad is None`, and :86 keeps a per-parameter state['step']): the arena range [0, n) is cut into
// `nseg` segments [bounds[s], bounds[s+1]) with a code each -- code < 0: the segment is left untouched (no moment decay,
// no weight decay, no step); code >= 0: the segment uses hyper slot `code` (hyper + 8 * code: its own bias-corrected
// step size, i.e. its own step count).  The task heads and KD projections of the task that did not run this step are
// such inactive segments.  bounds / codes live in device memory (one table per task: graph-replayable).
__global__ void __launch_bounds__(256)
    adamw_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     __nv_bfloat16* __restrict__ shadow, long long n, const float* __restrict__ hyper, float weight_decay,
                     const float* __restrict__ sumsq, const int* __restrict__ bounds, const int* __restrict__ codes,
                     int nseg) {
  __shared__ int s_b[MAGIC_ADAMW_MAX_SEGS + 1];
  __shared__ int s_c[MAGIC_ADAMW_MAX_SEGS];
  for (int i = threadIdx.x; i <= nseg; i += blockDim.x) s_b[i] = bounds[i];
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) s_c[i] = codes[i];
  __syncthreads();
  const float maxn = hyper[5];
  float clip = 1.f;
  if (sumsq != nullptr && maxn > 0.f) {
    const float c = maxn / (sqrtf(sumsq[0]) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  // segment by segment (a handful of them: heads and projections are contiguous in module order), grid-strided inside
  for (int sg = 0; sg < nseg; sg++) {
    const int code = s_c[sg];
    if (code < 0) continue;
    const float* h = hyper + 8 * code;
    const float lr = h[0], step = h[1], b1 = h[2], b2 = h[3], eps = h[4];
    const long long end = min((long long)s_b[sg + 1], n);
    for (long long i = (long long)s_b[sg] + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end;
         i += (long long)gridDim.x * blockDim.x) {
      const float gi = g[i] * clip;
      const float mi = m[i] * b1 + (1.f - b1) * gi;
      const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) + eps;
      float pi = p[i] + (-step) * (mi / denom);
      if (weight_decay > 0.f) pi = pi + (-lr * weight_decay) * pi;
      m[i] = mi;
      v[i] = vi;
      p[i] = pi;
      if (shadow) shadow[i] = __float2bfloat16_rn(pi);
    }
  }
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float s) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= s;
}

__global__ void delay_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

}  // namespace

extern "C" {

int magic_delay(long long cycles, cudaStream_t st) {
  delay_kernel<<<1, 1, 0, st>>>(cycles);
  MAGIC_CHECK_LAUNCH("magic_delay");
  return MAGIC_OK;
}


int magic_sumsq(const float* g, long long n, float* out, int zero_first, cudaStream_t st) {
  if (n <= 0) {
    if (zero_first) MAGIC_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st), "magic_sumsq");
    return MAGIC_OK;
  }
  MAGIC_CHECK_ARG(((uintptr_t)g % 16) == 0, "magic_sumsq: pointer must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (blocks > MAGIC_SUMSQ_SCRATCH) blocks = MAGIC_SUMSQ_SCRATCH;
  float* partials = out + 1;  // the caller's buffer carries the scratch: out[1 .. MAGIC_SUMSQ_SCRATCH]
  sumsq_partial_kernel<<<(int)blocks, 256, 0, st>>>(g, n, partials);
  MAGIC_CHECK_LAUNCH("magic_sumsq");
  sumsq_final_kernel<<<1, 256, 0, st>>>(partials, (int)blocks, out, zero_first ? 0 : 1);
  MAGIC_CHECK_LAUNCH("magic_sumsq(final)");
  return MAGIC_OK;
}

int magic_adamw(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, const float* hyper,
                float weight_decay, const float* sumsq, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = 16LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)bf16_shadow, n, hyper, weight_decay, sumsq);
  MAGIC_CHECK_LAUNCH("magic_adamw");
  return MAGIC_OK;
}

int magic_adamw_seg(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, const float* hyper,
                    float weight_decay, const float* sumsq, const int* bounds, const int* codes, int nseg,
                    cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  MAGIC_CHECK_ARG(nseg >= 1 && nseg <= MAGIC_ADAMW_MAX_SEGS && bounds && codes, "magic_adamw_seg: bad segment table (%d)",
                  nseg);
  MAGIC_CHECK_ARG(n < (1LL << 31), "magic_adamw_seg: n too large for 32-bit segment bounds");
  long long blocks = (n + 255) / 256;
  const long long cap = 16LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  adamw_seg_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, (__nv_bfloat16*)bf16_shadow, n, hyper, weight_decay, sumsq,
                                                bounds, codes, nseg);
  MAGIC_CHECK_LAUNCH("magic_adamw_seg");
  return MAGIC_OK;
}

int magic_scale(float* x, long long n, float s, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = 16LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  scale_kernel<<<(int)blocks, 256, 0, st>>>(x, n, s);
  MAGIC_CHECK_LAUNCH("magic_scale");
  return MAGIC_OK;
}

}  // extern "C"
