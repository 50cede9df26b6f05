// Epilogue shared by the SIMT and tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

#define MAGIC_ACT_NONE 0
#define MAGIC_ACT_GELU 1
#define MAGIC_ACT_RELU 2

struct GemmEpi {
  const float* bias;                   // [N] fp32 or null
  int act;                             // MAGIC_ACT_*
  void* pre_out;                       // forward: optional copy of (acc + bias) before the activation (dtype/ld of C)
  const void* dact_pre;                // backward: multiply by act'(pre[m,n]) (and the dropout scale) instead
  int dact_dt;                         // dtype of dact_pre
  long dact_ld;
  const void* residual;                // forward: optional residual[m,n] (dtype of C) added after act/dropout
  long res_ld;
  float alpha, beta;                   // C = epi(alpha*acc) + residual + beta*C
  int atomic;                          // split-K: accumulate with atomicAdd (fp32 C, linear epilogue only)
  float drop_p;                        // dropout applied to the activation OUTPUT (forward) / its gradient (backward)
  const unsigned long long* seed_ptr;  // device pointer
  uint32_t salt;
  float* rowsum;                       // optional [M] fp32: rowsum[m] += alpha * sum_k A(m,k)  (bias gradient of a wgrad GEMM)
};

__device__ __forceinline__ float act_fwd(int act, float v) {
  if (act == MAGIC_ACT_GELU) return gelu_f(v);
  if (act == MAGIC_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}
__device__ __forceinline__ float act_bwd(int act, float pre) {
  if (act == MAGIC_ACT_GELU) return gelu_grad_f(pre);
  if (act == MAGIC_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  return 1.f;
}

// bias_n = bias[n] hoisted by the caller (column-invariant)
template <typename TC>
__device__ __forceinline__ void epi_store(const GemmEpi& epi, const Dropout& dr, TC* C, float acc, int m, int n,
                                          long ldc, float bias_n) {
  float v = epi.alpha * acc;
  const size_t off = (size_t)m * ldc + n;
  if (epi.dact_pre != nullptr) {  // backward: dz = dh * dropscale * act'(z)
    const size_t poff = (size_t)m * epi.dact_ld + n;
    const float pre = epi.dact_dt == MAGIC_BF16 ? ldf((const __nv_bfloat16*)epi.dact_pre, poff)
                                                : ldf((const float*)epi.dact_pre, poff);
    v *= dr.scale(poff) * act_bwd(epi.act, pre);
  } else {
    v += bias_n;
    if (epi.pre_out) {
      stf((TC*)epi.pre_out, off, v);
      v = rt((const TC*)nullptr, v);  // activation sees what backward will re-read
    }
    v = act_fwd(epi.act, v);
    v *= dr.scale(off);
    if (epi.residual) v += ldf((const TC*)epi.residual, (size_t)m * epi.res_ld + n);
  }
  if (epi.atomic) {
    atomicAdd(reinterpret_cast<float*>(C) + off, v);
    return;
  }
  if (epi.beta != 0.f) v += epi.beta * ldf(C, off);
  stf(C, off, v);
}
