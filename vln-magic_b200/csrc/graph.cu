// Graph / head kernels: gmap node-feature aggregation (CSR segment-mean), SAP logit masking + fusion,
// row-wise cross-entropy.  Index/mask work is exact integer logic; the host builds the
// index tables from the viewpoint-id strings once per batch (vln-magic_b200/graph_index.py).
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

// entries e >= 0 : row e of `tokens` ([ΣT*V, h]);  e < 0 : row -(e+1) of `fused` ([ΣT, h])
template <typename T>
__global__ void __launch_bounds__(256)
    gmap_agg_fwd_kernel(const T* __restrict__ tokens, const T* __restrict__ fused, const int* __restrict__ node_ptr,
                        const int* __restrict__ entries, T* __restrict__ out, int n_nodes, int h) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int n = blockIdx.x * 8 + w; n < n_nodes; n += gridDim.x * 8) {
    const int e0 = node_ptr[n], e1 = node_ptr[n + 1];
    const float cnt = (float)(e1 - e0);
    for (int c = lane; c < h; c += 32) {
      float a = 0.f;
      for (int e = e0; e < e1; e++) {
        const int s = entries[e];
        a += s >= 0 ? ldf(tokens, (size_t)s * h + c) : ldf(fused, (size_t)(-(s + 1)) * h + c);
      }
      // single-entry nodes copy the source exactly (the oracle's mean over one element / direct use)
      stf(out, (size_t)n * h + c, (e1 - e0 <= 1) ? a : a / cnt);
    }
  }
}

// reverse CSR: for unique source s (encoded as above), d_src[s] = sum_e w[e] * dout[nodes[e]]
template <typename T>
__global__ void __launch_bounds__(256)
    gmap_agg_bwd_kernel(const T* __restrict__ dout, const int* __restrict__ src_ids, const int* __restrict__ src_ptr,
                        const int* __restrict__ src_nodes, const float* __restrict__ src_w, T* __restrict__ dtokens,
                        T* __restrict__ dfused, int n_src, int h) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = blockIdx.x * 8 + w; i < n_src; i += gridDim.x * 8) {
    const int s = src_ids[i];
    const int e0 = src_ptr[i], e1 = src_ptr[i + 1];
    if (e1 <= e0) continue;  // padding entry (fixed-capacity tables for CUDA-graph replay)
    for (int c = lane; c < h; c += 32) {
      float a = 0.f;
      for (int e = e0; e < e1; e++) a = fmaf(src_w[e], ldf(dout, (size_t)src_nodes[e] * h + c), a);
      if (s >= 0) stf(dtokens, (size_t)s * h + c, a);
      else stf(dfused, (size_t)(-(s + 1)) * h + c, a);
    }
  }
}

// one CTA (64 threads) per sample
__global__ void __launch_bounds__(64)
    sap_fuse_fwd_kernel(const float* __restrict__ g_raw, const float* __restrict__ l_raw,
                        const float* __restrict__ gate_raw, const unsigned char* __restrict__ g_valid,
                        const unsigned char* __restrict__ l_valid, const int* __restrict__ node2cand,
                        const unsigned char* __restrict__ bw_mask, float* __restrict__ gl, float* __restrict__ ll,
                        float* __restrict__ fl, int G, int Vp) {
  extern __shared__ float sm[];  // [Vp] local logits, [1] bw
  const int b = blockIdx.x;
  const float fw = gate_raw ? 1.f / (1.f + expf(-gate_raw[b])) : 0.5f;
  for (int j = threadIdx.x; j < Vp; j += blockDim.x) {
    const float v = l_valid[(size_t)b * Vp + j] ? l_raw[(size_t)b * Vp + j] * (1.f - fw) : -INFINITY;
    sm[j] = v;
    ll[(size_t)b * Vp + j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bw = 0.f;  // sequential, in candidate order, like the reference python loop
    for (int j = 1; j < Vp; j++)
      if (bw_mask[(size_t)b * Vp + j]) bw += sm[j];
    sm[Vp] = bw;
  }
  __syncthreads();
  const float bw = sm[Vp];
  for (int n = threadIdx.x; n < G; n += blockDim.x) {
    const bool ok = g_valid[(size_t)b * G + n];
    const float g = ok ? g_raw[(size_t)b * G + n] * fw : -INFINITY;
    gl[(size_t)b * G + n] = g;
    float f = g;
    if (n == 0) f += sm[0];
    else if (ok) {
      const int c = node2cand[(size_t)b * G + n];
      f += (c >= 0) ? sm[c] : bw;
    }
    fl[(size_t)b * G + n] = f;
  }
}

__global__ void __launch_bounds__(64)
    sap_fuse_bwd_kernel(const float* __restrict__ dgl, const float* __restrict__ dll, const float* __restrict__ dfl,
                        const float* __restrict__ g_raw, const float* __restrict__ l_raw,
                        const float* __restrict__ gate_raw, const unsigned char* __restrict__ g_valid,
                        const unsigned char* __restrict__ l_valid, const int* __restrict__ node2cand,
                        const unsigned char* __restrict__ bw_mask, float* __restrict__ dg_raw,
                        float* __restrict__ dl_raw, float* __restrict__ dgate_raw, int G, int Vp) {
  extern __shared__ float sm[];  // [Vp] d(ll total), [1] d(bw)
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float fw = gate_raw ? 1.f / (1.f + expf(-gate_raw[b])) : 0.5f;
  for (int j = threadIdx.x; j <= Vp; j += blockDim.x) sm[j] = (j < Vp) ? dll[(size_t)b * Vp + j] : 0.f;
  __syncthreads();
  float dfw = 0.f;
  for (int n = threadIdx.x; n < G; n += blockDim.x) {
    const bool ok = g_valid[(size_t)b * G + n];
    const float df = dfl[(size_t)b * G + n];
    const float dg = ok ? (dgl[(size_t)b * G + n] + df) : 0.f;
    dg_raw[(size_t)b * G + n] = dg * fw;
    dfw += dg * (ok ? g_raw[(size_t)b * G + n] : 0.f);
    if (n == 0) atomicAdd(&sm[0], df);
    else if (ok) {
      const int c = node2cand[(size_t)b * G + n];
      atomicAdd(c >= 0 ? &sm[c] : &sm[Vp], df);
    }
  }
  __syncthreads();
  const float dbw = sm[Vp];
  for (int j = threadIdx.x; j < Vp; j += blockDim.x) {
    const bool ok = l_valid[(size_t)b * Vp + j];
    float d = sm[j];
    if (j >= 1 && bw_mask[(size_t)b * Vp + j]) d += dbw;
    d = ok ? d : 0.f;
    dl_raw[(size_t)b * Vp + j] = d * (1.f - fw);
    dfw -= d * (ok ? l_raw[(size_t)b * Vp + j] : 0.f);
  }
  dfw = block_sum(dfw, red);
  if (threadIdx.x == 0 && dgate_raw) dgate_raw[b] = gate_raw ? dfw * fw * (1.f - fw) : 0.f;
}

// ---- cross-entropy, one CTA per row ---------------------------------------------------------------
// one CTA per row; a single pass over the logits (online softmax), 16-byte loads when `vec` (aligned rows).
// Works in the log2 domain with the running maximum updated once per 16-byte vector (vector max first): one FMUL +
// one FADD + one MUFU.EX2 per logit and no per-element branch (the per-element lse_push form was compute-bound:
// 40 us for the 77 MB of MLM logits).
__device__ __forceinline__ float ce_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int VN>
__device__ __forceinline__ void ce_push(float& m, float& s, const float (&v)[VN]) {
  constexpr float LOG2E = 1.4426950408889634f;
  float u[VN];
  float vm = -INFINITY;
#pragma unroll
  for (int i = 0; i < VN; i++) {
    u[i] = v[i] * LOG2E;
    vm = fmaxf(vm, u[i]);
  }
  if (vm > m) {
    s *= ce_ex2(m - vm);  // m = -inf: s is 0 and 2^-inf = 0
    m = vm;
  }
  if (m > -INFINITY) {  // (a thread that has only seen -inf logits contributes nothing)
#pragma unroll
    for (int i = 0; i < VN; i++) s += ce_ex2(u[i] - m);
  }
}
__device__ __forceinline__ void ce_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  if (mm == -INFINITY) return;
  s = s * ce_ex2(m - mm) + s2 * ce_ex2(m2 - mm);
  m = mm;
}

template <typename T>
__global__ void __launch_bounds__(256)
    ce_fwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels, float* __restrict__ loss,
                  float* __restrict__ lse_out, int C, long ld, long long ignore_index, int vec) {
  __shared__ float red_m[8], red_s[8];
  const int r = blockIdx.x;
  const T* row = logits + (size_t)r * ld;
  float m = -INFINITY, s = 0.f;  // m in log2 units
  constexpr int VN = RowVec<T>::N;
  const int cv = vec ? (C / VN) * VN : 0;
  const int step = blockDim.x * VN;
  int c = threadIdx.x * VN;
  for (; c + step < cv; c += 2 * step) {  // two independent 16-byte loads in flight
    float v0[VN], v1[VN];
    RowVec<T>::load(row + c, v0);
    RowVec<T>::load(row + c + step, v1);
    ce_push<VN>(m, s, v0);
    ce_push<VN>(m, s, v1);
  }
  for (; c < cv; c += step) {
    float v0[VN];
    RowVec<T>::load(row + c, v0);
    ce_push<VN>(m, s, v0);
  }
  for (int cc = cv + threadIdx.x; cc < C; cc += blockDim.x) {
    const float v1[1] = {ldf(row, cc)};
    ce_push<1>(m, s, v1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    ce_merge(m, s, m2, s2);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    red_m[w] = m;
    red_s[w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; i++) ce_merge(m, s, red_m[i], red_s[i]);
    const float lse = (m + log2f(s)) * 0.6931471805599453f;
    lse_out[r] = lse;
    const long long y = labels[r];
    loss[r] = (y == ignore_index) ? 0.f : lse - ldf(row, (size_t)y);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    ce_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ labels, const float* __restrict__ lse,
                  const float* __restrict__ dloss, T* __restrict__ dlogits, int C, long ld, long long ignore_index,
                  int vec) {
  const int r = blockIdx.x;
  const long long y = labels[r];
  const float g = (y == ignore_index) ? 0.f : dloss[r];
  const float l2 = lse[r] * 1.4426950408889634f;  // softmax as 2^(v * log2 e - lse * log2 e): one FFMA + one MUFU.EX2
  const T* row = logits + (size_t)r * ld;
  T* drow = dlogits + (size_t)r * ld;
  constexpr int VN = RowVec<T>::N;
  const int cv = vec ? (C / VN) * VN : 0;
  if (y == ignore_index) {  // exact zeros, also for a row whose logits are all -inf (lse = -inf would give NaN)
    float z[VN];
#pragma unroll
    for (int i = 0; i < VN; i++) z[i] = 0.f;
    for (int c = threadIdx.x * VN; c < cv; c += blockDim.x * VN) RowVec<T>::store(drow + c, z);
    for (int c = cv + threadIdx.x; c < C; c += blockDim.x) stf(drow, c, 0.f);
    return;
  }
  for (int c = threadIdx.x * VN; c < cv; c += blockDim.x * VN) {
    float v[VN];
    RowVec<T>::load(row + c, v);
#pragma unroll
    for (int i = 0; i < VN; i++)
      v[i] = g * (ce_ex2(fmaf(v[i], 1.4426950408889634f, -l2)) - ((long long)(c + i) == y ? 1.f : 0.f));
    RowVec<T>::store(drow + c, v);
  }
  for (int c = cv + threadIdx.x; c < C; c += blockDim.x) {
    const float p = ce_ex2(fmaf(ldf(row, c), 1.4426950408889634f, -l2));
    stf(drow, c, g * (p - ((long long)c == y ? 1.f : 0.f)));
  }
}

}  // namespace

extern "C" {

int magic_gmap_aggregate_fwd(const void* tokens, const void* fused, const int* node_ptr, const int* entries,
                             void* out, int n_nodes, int h, int dtype, cudaStream_t st) {
  if (n_nodes <= 0) return MAGIC_OK;
  int grid = (n_nodes + 7) / 8;
  if (dtype == MAGIC_F32)
    gmap_agg_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)tokens, (const float*)fused, node_ptr, entries,
                                                    (float*)out, n_nodes, h);
  else if (dtype == MAGIC_BF16)
    gmap_agg_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)tokens,
                                                            (const __nv_bfloat16*)fused, node_ptr, entries,
                                                            (__nv_bfloat16*)out, n_nodes, h);
  else {
    magic_set_error("magic_gmap_aggregate_fwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_gmap_aggregate_fwd");
  return MAGIC_OK;
}

int magic_gmap_aggregate_bwd(const void* dout, const int* src_ids, const int* src_ptr, const int* src_nodes,
                             const float* src_w, void* dtokens, long long n_token_rows, void* dfused,
                             long long n_fused_rows, int n_src, int h, int dtype, cudaStream_t st) {
  const size_t esz = dtype == MAGIC_BF16 ? 2 : 4;
  MAGIC_CUDA(cudaMemsetAsync(dtokens, 0, (size_t)n_token_rows * h * esz, st), "magic_gmap_aggregate_bwd");
  MAGIC_CUDA(cudaMemsetAsync(dfused, 0, (size_t)n_fused_rows * h * esz, st), "magic_gmap_aggregate_bwd");
  if (n_src <= 0) return MAGIC_OK;
  int grid = (n_src + 7) / 8;
  if (dtype == MAGIC_F32)
    gmap_agg_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dout, src_ids, src_ptr, src_nodes, src_w,
                                                    (float*)dtokens, (float*)dfused, n_src, h);
  else if (dtype == MAGIC_BF16)
    gmap_agg_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dout, src_ids, src_ptr, src_nodes,
                                                            src_w, (__nv_bfloat16*)dtokens, (__nv_bfloat16*)dfused,
                                                            n_src, h);
  else {
    magic_set_error("magic_gmap_aggregate_bwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_gmap_aggregate_bwd");
  return MAGIC_OK;
}

int magic_sap_fuse_fwd(const float* g_raw, const float* l_raw, const float* gate_raw, const unsigned char* g_valid,
                       const unsigned char* l_valid, const int* node2cand, const unsigned char* bw_mask, float* gl,
                       float* ll, float* fl, int B, int G, int Vp, cudaStream_t st) {
  if (B <= 0) return MAGIC_OK;
  sap_fuse_fwd_kernel<<<B, 64, (Vp + 1) * sizeof(float), st>>>(g_raw, l_raw, gate_raw, g_valid, l_valid, node2cand,
                                                               bw_mask, gl, ll, fl, G, Vp);
  MAGIC_CHECK_LAUNCH("magic_sap_fuse_fwd");
  return MAGIC_OK;
}

int magic_sap_fuse_bwd(const float* dgl, const float* dll, const float* dfl, const float* g_raw, const float* l_raw,
                       const float* gate_raw, const unsigned char* g_valid, const unsigned char* l_valid,
                       const int* node2cand, const unsigned char* bw_mask, float* dg_raw, float* dl_raw,
                       float* dgate_raw, int B, int G, int Vp, cudaStream_t st) {
  if (B <= 0) return MAGIC_OK;
  sap_fuse_bwd_kernel<<<B, 64, (Vp + 1) * sizeof(float), st>>>(dgl, dll, dfl, g_raw, l_raw, gate_raw, g_valid,
                                                               l_valid, node2cand, bw_mask, dg_raw, dl_raw, dgate_raw,
                                                               G, Vp);
  MAGIC_CHECK_LAUNCH("magic_sap_fuse_bwd");
  return MAGIC_OK;
}

int magic_ce_fwd(const void* logits, const long long* labels, float* loss, float* lse, int R, int C, long ld,
                 long long ignore_index, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  const int esz = dtype == MAGIC_BF16 ? 2 : 4;
  const int vec = (((uintptr_t)logits & 15) == 0 && (ld * esz) % 16 == 0) ? 1 : 0;
  if (dtype == MAGIC_F32)
    ce_fwd_kernel<float><<<R, 256, 0, st>>>((const float*)logits, labels, loss, lse, C, ld, ignore_index, vec);
  else if (dtype == MAGIC_BF16)
    ce_fwd_kernel<__nv_bfloat16><<<R, 256, 0, st>>>((const __nv_bfloat16*)logits, labels, loss, lse, C, ld,
                                                   ignore_index, vec);
  else {
    magic_set_error("magic_ce_fwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_ce_fwd");
  return MAGIC_OK;
}

int magic_ce_bwd(const void* logits, const long long* labels, const float* lse, const float* dloss, void* dlogits,
                 int R, int C, long ld, long long ignore_index, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  const int esz = dtype == MAGIC_BF16 ? 2 : 4;
  const int vec = (((uintptr_t)logits & 15) == 0 && ((uintptr_t)dlogits & 15) == 0 && (ld * esz) % 16 == 0) ? 1 : 0;
  if (dtype == MAGIC_F32)
    ce_bwd_kernel<float><<<R, 256, 0, st>>>((const float*)logits, labels, lse, dloss, (float*)dlogits, C, ld,
                                            ignore_index, vec);
  else if (dtype == MAGIC_BF16)
    ce_bwd_kernel<__nv_bfloat16><<<R, 256, 0, st>>>((const __nv_bfloat16*)logits, labels, lse, dloss,
                                                   (__nv_bfloat16*)dlogits, C, ld, ignore_index, vec);
  else {
    magic_set_error("magic_ce_bwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_ce_bwd");
  return MAGIC_OK;
}

}  // extern "C"
