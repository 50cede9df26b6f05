// Row-wise fused kernels (HBM-bound): LayerNorm(+residual,+dropout), text embedding+LN, positional
// "posfuse" (small-K linear + LN + embedding adds), row gather/scatter, adaptive panorama pooling,
// N=1 heads (rowdot), bias-gradient column sums, dtype casts.
// One warp owns one row of h <= 768 elements held in registers (lane i owns columns i, i+32, ...):
// every global access of a warp is a contiguous 128 B (fp32) / 64 B (bf16) segment.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int MAXE = 24;        // 24 * 32 = 768 columns max per row
constexpr int ROW_WARPS = 8;    // warps per CTA for the row kernels

__host__ int row_grid(int M) {
  int g = (M + ROW_WARPS - 1) / ROW_WARPS;
  int cap = magic_num_sms() * 4;
  return g < cap ? (g > 0 ? g : 1) : cap;
}

// ---- LayerNorm math on a register-resident row ------------------------------------------------
template <int NE>
__device__ __forceinline__ void row_stats(const float (&v)[NE], int h, int lane, float eps, float& mean,
                                          float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NE; i++)
    if (lane + 32 * i < h) s += v[i];
  mean = warp_sum(s) / (float)h;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NE; i++)
    if (lane + 32 * i < h) {
      const float d = v[i] - mean;
      q += d * d;
    }
  rstd = rsqrtf(warp_sum(q) / (float)h + eps);
}

// given xhat (in v) and dxhat = g*gamma (in d): dv = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))
template <int NE>
__device__ __forceinline__ void ln_bwd_row(const float (&xhat)[NE], float (&d)[NE], int h, int lane, float rstd) {
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < NE; i++)
    if (lane + 32 * i < h) {
      c1 += d[i];
      c2 += d[i] * xhat[i];
    }
  c1 = warp_sum(c1) / (float)h;
  c2 = warp_sum(c2) / (float)h;
#pragma unroll
  for (int i = 0; i < NE; i++)
    if (lane + 32 * i < h) d[i] = rstd * (d[i] - c1 - xhat[i] * c2);
}

// CTA-level accumulation of per-column partial sums into global fp32 (one atomic per column per CTA)
__device__ __forceinline__ void flush_cols(float* smem_acc, float* gout, int h) {
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    const float v = smem_acc[c];
    if (v != 0.f) atomicAdd(gout + c, v);
  }
}

// =================================================================================================
// LayerNorm forward:  y = drop_out( LN( drop_in(x) + res ) * gamma + beta )
// =================================================================================================
template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                  const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ stats, int M, int h,
                  float eps, float p_in, uint32_t salt_in, float p_out, uint32_t salt_out,
                  const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Dropout din = make_dropout(p_in, seed_ptr, salt_in), dout = make_dropout(p_out, seed_ptr, salt_out);
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    float v[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      v[i] = 0.f;
      if (c < h) {
        float t = ldf(x, base + c) * din.scale(base + c);
        if (res) t += ldf(res, base + c);
        v[i] = t;
      }
    }
    float mean, rstd;
    row_stats(v, h, lane, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) stf(y, base + c, ((v[i] - mean) * rstd * gamma[c] + beta[c]) * dout.scale(base + c));
    }
    if (lane == 0 && stats) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                  const float* __restrict__ gamma, const float* __restrict__ stats, T* __restrict__ dx,
                  T* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int h,
                  float p_in, uint32_t salt_in, float p_out, uint32_t salt_out,
                  const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];  // [2*h] : dgamma | dbeta partials
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < 2 * h; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const Dropout din = make_dropout(p_in, seed_ptr, salt_in), dout = make_dropout(p_out, seed_ptr, salt_out);
  float pg[NE], pb[NE];
#pragma unroll
  for (int i = 0; i < NE; i++) pg[i] = pb[i] = 0.f;
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float xh[NE], d[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      xh[i] = d[i] = 0.f;
      if (c < h) {
        float t = ldf(x, base + c) * din.scale(base + c);
        if (res) t += ldf(res, base + c);
        xh[i] = (t - mean) * rstd;
        const float g = ldf(dy, base + c) * dout.scale(base + c);
        pg[i] += g * xh[i];
        pb[i] += g;
        d[i] = g * gamma[c];
      }
    }
    ln_bwd_row(xh, d, h, lane, rstd);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) {
        if (dres) stf(dres, base + c, d[i]);
        stf(dx, base + c, d[i] * din.scale(base + c));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int c = lane + 32 * i;
    if (c < h) {
      atomicAdd(&sm[c], pg[i]);
      atomicAdd(&sm[h + c], pb[i]);
    }
  }
  flush_cols(sm, dgamma, h);
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    const float v = sm[h + c];
    if (v != 0.f) atomicAdd(dbeta + c, v);
  }
}

// =================================================================================================
// LayerNorm, vectorised variant for h % 128 == 0 (every model width of the reference: 128/256/384/768):
// a lane owns 4 CONSECUTIVE columns of each 128-column chunk, so a warp moves one contiguous 256 B (bf16) or
// 512 B (fp32) segment per instruction instead of 32 two-byte elements, and gamma/beta sit in registers for the
// whole row loop.  Same math, same dropout indices (mask = hash(seed, salt, row*h + col)) as the scalar kernels.
// =================================================================================================
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 t;
    *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    ln_fwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                      const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ stats, int M, float eps,
                      float p_in, uint32_t salt_in, float p_out, uint32_t salt_out,
                      const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  constexpr int h = NCH * 128;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Dropout din = make_dropout(p_in, seed_ptr, salt_in), dout = make_dropout(p_out, seed_ptr, salt_out);
  float g[NCH][4], bt[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    Vec4<float>::load(gamma + c * 128 + lane * 4, g[c]);
    Vec4<float>::load(beta + c * 128 + lane * 4, bt[c]);
  }
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h + lane * 4;
    float v[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; c++) Vec4<T>::load(x + base + c * 128, v[c]);
    if (p_in > 0.f) {
#pragma unroll
      for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int e = 0; e < 4; e++) v[c][e] *= din.scale(base + c * 128 + e);
    }
    if (res) {
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        float t[4];
        Vec4<T>::load(res + base + c * 128, t);
#pragma unroll
        for (int e = 0; e < 4; e++) v[c][e] += t[e];
      }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) s += (v[c][0] + v[c][1]) + (v[c][2] + v[c][3]);
    const float mean = warp_sum(s) * (1.f / (float)h);
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float d = v[c][e] - mean;
        q += d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / (float)h) + eps);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; e++) o[e] = (v[c][e] - mean) * rstd * g[c][e] + bt[c][e];
      if (p_out > 0.f) {
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] *= dout.scale(base + c * 128 + e);
      }
      Vec4<T>::store(y + base + c * 128, o);
    }
    if (lane == 0 && stats) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    ln_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                      const float* __restrict__ gamma, const float* __restrict__ stats, T* __restrict__ dx,
                      T* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, float p_in,
                      uint32_t salt_in, float p_out, uint32_t salt_out, const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  constexpr int h = NCH * 128;
  extern __shared__ float sm[];  // [ROW_WARPS][2*h] : per-warp dgamma | dbeta partials
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Dropout din = make_dropout(p_in, seed_ptr, salt_in), dout = make_dropout(p_out, seed_ptr, salt_out);
  float g[NCH][4], pg[NCH][4], pb[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    Vec4<float>::load(gamma + c * 128 + lane * 4, g[c]);
#pragma unroll
    for (int e = 0; e < 4; e++) pg[c][e] = pb[c][e] = 0.f;
  }
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h + lane * 4;
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float xh[NCH][4], d[NCH][4];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      float t[4], gy[4];
      Vec4<T>::load(x + base + c * 128, t);
      Vec4<T>::load(dy + base + c * 128, gy);
      if (p_in > 0.f) {
#pragma unroll
        for (int e = 0; e < 4; e++) t[e] *= din.scale(base + c * 128 + e);
      }
      if (res) {
        float rr[4];
        Vec4<T>::load(res + base + c * 128, rr);
#pragma unroll
        for (int e = 0; e < 4; e++) t[e] += rr[e];
      }
      if (p_out > 0.f) {
#pragma unroll
        for (int e = 0; e < 4; e++) gy[e] *= dout.scale(base + c * 128 + e);
      }
#pragma unroll
      for (int e = 0; e < 4; e++) {
        xh[c][e] = (t[e] - mean) * rstd;
        pg[c][e] += gy[e] * xh[c][e];
        pb[c][e] += gy[e];
        d[c][e] = gy[e] * g[c][e];
        c1 += d[c][e];
        c2 += d[c][e] * xh[c][e];
      }
    }
    c1 = warp_sum(c1) * (1.f / (float)h);
    c2 = warp_sum(c2) * (1.f / (float)h);
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; e++) o[e] = rstd * (d[c][e] - c1 - xh[c][e] * c2);
      if (dres) Vec4<T>::store(dres + base + c * 128, o);
      if (p_in > 0.f) {
#pragma unroll
        for (int e = 0; e < 4; e++) o[e] *= din.scale(base + c * 128 + e);
      }
      Vec4<T>::store(dx + base + c * 128, o);
    }
  }
  // column sums: every warp parks its register partials in its own smem row (128-bit stores, no atomics: shared
  // fp32 atomics are CAS loops and 8 warps hitting the same 2h words made this tail the longest part of the
  // kernel), then each thread adds the ROW_WARPS rows of its columns and issues ONE global reduction per column
  float* mine = sm + (size_t)w * 2 * h;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    *reinterpret_cast<float4*>(mine + c * 128 + lane * 4) = make_float4(pg[c][0], pg[c][1], pg[c][2], pg[c][3]);
    *reinterpret_cast<float4*>(mine + h + c * 128 + lane * 4) = make_float4(pb[c][0], pb[c][1], pb[c][2], pb[c][3]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * h; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int ww = 0; ww < ROW_WARPS; ww++) v += sm[(size_t)ww * 2 * h + c];
    if (v != 0.f) atomicAdd(c < h ? dgamma + c : dbeta + (c - h), v);
  }
}

// =================================================================================================
// text embedding + LayerNorm:  y[r] = drop_out( LN(word[ids[r]] + pos[r % L] + type0) )
// =================================================================================================
template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    embed_ln_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                        const float* __restrict__ pos, const float* __restrict__ type0,
                        const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
                        float* __restrict__ stats, int M, int L, int h, float eps, float p_out, uint32_t salt_out,
                        const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Dropout dout = make_dropout(p_out, seed_ptr, salt_out);
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    const size_t wb = (size_t)ids[r] * h, pb = (size_t)(r % L) * h;
    float v[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      v[i] = (c < h) ? (word[wb + c] + pos[pb + c] + type0[c]) : 0.f;
    }
    float mean, rstd;
    row_stats(v, h, lane, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) stf(y, base + c, ((v[i] - mean) * rstd * gamma[c] + beta[c]) * dout.scale(base + c));
    }
    if (lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    embed_ln_bwd_kernel(const T* __restrict__ dy, const long long* __restrict__ ids, const float* __restrict__ word,
                        const float* __restrict__ pos, const float* __restrict__ type0,
                        const float* __restrict__ gamma, const float* __restrict__ stats, float* __restrict__ dword,
                        float* __restrict__ dpos, float* __restrict__ dtype0, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int M, int L, int h, float p_out, uint32_t salt_out,
                        const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];  // [3*h] : dgamma | dbeta | dtype0
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < 3 * h; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const Dropout dout = make_dropout(p_out, seed_ptr, salt_out);
  float pg[NE], pbt[NE], pt[NE];
#pragma unroll
  for (int i = 0; i < NE; i++) pg[i] = pbt[i] = pt[i] = 0.f;
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    const size_t wb = (size_t)ids[r] * h, pb = (size_t)(r % L) * h;
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float xh[NE], d[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      xh[i] = d[i] = 0.f;
      if (c < h) {
        xh[i] = (word[wb + c] + pos[pb + c] + type0[c] - mean) * rstd;
        const float g = ldf(dy, base + c) * dout.scale(base + c);
        pg[i] += g * xh[i];
        pbt[i] += g;
        d[i] = g * gamma[c];
      }
    }
    ln_bwd_row(xh, d, h, lane, rstd);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) {
        atomicAdd(dword + wb + c, d[i]);
        atomicAdd(dpos + pb + c, d[i]);
        pt[i] += d[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int c = lane + 32 * i;
    if (c < h) {
      atomicAdd(&sm[c], pg[i]);
      atomicAdd(&sm[h + c], pbt[i]);
      atomicAdd(&sm[2 * h + c], pt[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    atomicAdd(dgamma + c, sm[c]);
    atomicAdd(dbeta + c, sm[h + c]);
    atomicAdd(dtype0 + c, sm[2 * h + c]);
  }
}

// =================================================================================================
// posfuse:  y[r] = xin[r] + emb[idx[r]] + cst + LN(W f[r] + b; gamma, beta)      (K <= 16)
// =================================================================================================
constexpr int MAXK = 16;

template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    posfuse_fwd_kernel(const T* __restrict__ xin, const long long* __restrict__ idx, const float* __restrict__ emb,
                       const float* __restrict__ cst, const float* __restrict__ f, const float* __restrict__ W,
                       const float* __restrict__ b, const float* __restrict__ gamma, const float* __restrict__ beta,
                       T* __restrict__ y, float* __restrict__ stats, int M, int h, int K, float eps) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    float fr[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; k++) fr[k] = (k < K) ? f[(size_t)r * K + k] : 0.f;
    float v[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      v[i] = 0.f;
      if (c < h) {
        float u = b[c];
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) u = fmaf(W[(size_t)c * K + k], fr[k], u);
        v[i] = u;
      }
    }
    float mean, rstd;
    row_stats(v, h, lane, eps, mean, rstd);
    const size_t eb = idx ? (size_t)idx[r] * h : 0;
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) {
        float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
        if (xin) o += ldf(xin, base + c);
        if (emb) o += emb[eb + c];
        if (cst) o += cst[c];
        stf(y, base + c, o);
      }
    }
    if (lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
  }
}

// backward: dxin == dy (handled by the caller).  smem: [h*K] dW | [h] db | [h] dgamma | [h] dbeta | [h] dcst
template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    posfuse_bwd_kernel(const T* __restrict__ dy, const long long* __restrict__ idx, const float* __restrict__ f,
                       const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
                       const float* __restrict__ stats, float* __restrict__ demb, float* __restrict__ dcst,
                       float* __restrict__ dW, float* __restrict__ db, float* __restrict__ dgamma,
                       float* __restrict__ dbeta, int M, int h, int K) {
  extern __shared__ float sm[];
  float* s_dW = sm;
  float* s_db = sm + (size_t)h * K;
  float* s_dg = s_db + h;
  float* s_dbt = s_dg + h;
  float* s_dc = s_dbt + h;
  const int tot = h * K + 4 * h;
  for (int c = threadIdx.x; c < tot; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float pg[NE], pbt[NE], pdb[NE];
  // dW partials live in registers across the row loop when they fit (h <= 128): one shared-memory atomic per
  // (warp, weight) at the end instead of one per (row, weight)
  constexpr bool REG_DW = NE * MAXK <= 64;
  float pw[REG_DW ? NE : 1][REG_DW ? MAXK : 1];
#pragma unroll
  for (int i = 0; i < NE; i++) pg[i] = pbt[i] = pdb[i] = 0.f;
  if (REG_DW) {
#pragma unroll
    for (int i = 0; i < (REG_DW ? NE : 1); i++)
#pragma unroll
      for (int k = 0; k < (REG_DW ? MAXK : 1); k++) pw[i][k] = 0.f;
  }
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const size_t base = (size_t)r * h;
    float fr[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; k++) fr[k] = (k < K) ? f[(size_t)r * K + k] : 0.f;
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    const size_t eb = idx ? (size_t)idx[r] * h : 0;
    float xh[NE], d[NE];
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      xh[i] = d[i] = 0.f;
      if (c < h) {
        float u = b[c];
#pragma unroll
        for (int k = 0; k < MAXK; k++)
          if (k < K) u = fmaf(W[(size_t)c * K + k], fr[k], u);
        xh[i] = (u - mean) * rstd;
        const float g = ldf(dy, base + c);
        pg[i] += g * xh[i];
        pbt[i] += g;
        d[i] = g * gamma[c];
        if (demb) atomicAdd(demb + eb + c, g);
      }
    }
    ln_bwd_row(xh, d, h, lane, rstd);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) {
        pdb[i] += d[i];
        if (REG_DW) {
#pragma unroll
          for (int k = 0; k < MAXK; k++) pw[REG_DW ? i : 0][REG_DW ? k : 0] = fmaf(d[i], fr[k], pw[REG_DW ? i : 0][REG_DW ? k : 0]);
        } else {
#pragma unroll
          for (int k = 0; k < MAXK; k++)
            if (k < K) atomicAdd(&s_dW[(size_t)c * K + k], d[i] * fr[k]);
        }
      }
    }
  }
  if (REG_DW) {
#pragma unroll
    for (int i = 0; i < (REG_DW ? NE : 1); i++) {
      const int c = lane + 32 * i;
      if (c < h) {
#pragma unroll
        for (int k = 0; k < (REG_DW ? MAXK : 1); k++)
          if (k < K) atomicAdd(&s_dW[(size_t)c * K + k], pw[i][k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int c = lane + 32 * i;
    if (c < h) {
      atomicAdd(&s_dg[c], pg[i]);
      atomicAdd(&s_dbt[c], pbt[i]);
      atomicAdd(&s_db[c], pdb[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h * K; c += blockDim.x) atomicAdd(dW + c, s_dW[c]);
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    atomicAdd(db + c, s_db[c]);
    atomicAdd(dgamma + c, s_dg[c]);
    atomicAdd(dbeta + c, s_dbt[c]);
    if (dcst) atomicAdd(dcst + c, s_dbt[c]);  // d(cst) = sum_r dy = same sum as dbeta
  }
  (void)s_dc;
}

// =================================================================================================
// row gather / scatter
// =================================================================================================
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, const long long* __restrict__ idx, T* __restrict__ out,
                                   int R, int h) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * ROW_WARPS + w; r < R; r += gridDim.x * ROW_WARPS) {
    const long long s = idx[r];
    for (int c = lane; c < h; c += 32) stf(out, (size_t)r * h + c, s >= 0 ? ldf(src, (size_t)s * h + c) : 0.f);
  }
}
template <typename T>
__global__ void scatter_rows_kernel(const T* __restrict__ dout, const long long* __restrict__ idx,
                                    T* __restrict__ dsrc, int R, int h) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * ROW_WARPS + w; r < R; r += gridDim.x * ROW_WARPS) {
    const long long s = idx[r];
    if (s < 0) continue;
    for (int c = lane; c < h; c += 32) stf(dsrc, (size_t)s * h + c, ldf(dout, (size_t)r * h + c));
  }
}

// =================================================================================================
// adaptive panorama pooling:  fused[r] = sum_v softmax_v(x[r,v].w + b | v < len[r]) x[r,v]
// (w == nullptr -> masked mean).  One CTA (128 threads) per panorama.
// =================================================================================================
constexpr int PF_MAXV = 128;

template <typename T>
__global__ void __launch_bounds__(128)
    pano_fuse_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wv, const float* __restrict__ bias,
                         const long long* __restrict__ lens, T* __restrict__ fused, float* __restrict__ probs, int V,
                         int h) {
  __shared__ float s[PF_MAXV];
  const int r = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int len = (int)min((long long)V, lens[r]);
  const T* xr = x + (size_t)r * V * h;
  for (int v = w; v < V; v += 4) {
    float sc = -INFINITY;
    if (v < len) {
      if (wv) {
        float a = 0.f;
        for (int c = lane; c < h; c += 32) a = fmaf(ldf(xr, (size_t)v * h + c), wv[c], a);
        sc = warp_sum(a) + bias[0];
      } else {
        sc = 0.f;
      }
    }
    if (lane == 0) s[v] = sc;
  }
  __syncthreads();
  if (w == 0) {
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, s[v]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int v = lane; v < V; v += 32) {
      const float e = (v < len) ? __expf(s[v] - mx) : 0.f;
      s[v] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int v = lane; v < V; v += 32) {
      s[v] *= inv;
      probs[(size_t)r * V + v] = s[v];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    float a = 0.f;
    for (int v = 0; v < len; v++) a = fmaf(s[v], ldf(xr, (size_t)v * h + c), a);
    stf(fused, (size_t)r * h + c, a);
  }
}

template <typename T>
__global__ void __launch_bounds__(128)
    pano_fuse_bwd_kernel(const T* __restrict__ dfused, const T* __restrict__ x, const float* __restrict__ wv,
                         const long long* __restrict__ lens, const float* __restrict__ probs, T* __restrict__ dx,
                         float* __restrict__ dw, float* __restrict__ dbias, int V, int h) {
  __shared__ float s_dp[PF_MAXV];
  __shared__ float s_ds[PF_MAXV];
  __shared__ float s_D;
  const int r = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int len = (int)min((long long)V, lens[r]);
  const T* xr = x + (size_t)r * V * h;
  const T* df = dfused + (size_t)r * h;
  const float* pr = probs + (size_t)r * V;
  for (int v = w; v < V; v += 4) {
    float a = 0.f;
    if (v < len)
      for (int c = lane; c < h; c += 32) a = fmaf(ldf(df, c), ldf(xr, (size_t)v * h + c), a);
    a = warp_sum(a);
    if (lane == 0) s_dp[v] = a;
  }
  __syncthreads();
  if (w == 0) {
    float D = 0.f;
    for (int v = lane; v < len; v += 32) D += pr[v] * s_dp[v];
    D = warp_sum(D);
    float dsum = 0.f;
    for (int v = lane; v < V; v += 32) {
      const float ds = (v < len && wv) ? pr[v] * (s_dp[v] - D) : 0.f;
      s_ds[v] = ds;
      dsum += ds;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) {
      s_D = D;
      if (dbias && wv) atomicAdd(dbias, dsum);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) {
    const float dfc = ldf(df, c);
    const float wc = wv ? wv[c] : 0.f;
    float dwc = 0.f;
    for (int v = 0; v < V; v++) {
      float o = 0.f;
      if (v < len) {
        o = pr[v] * dfc + s_ds[v] * wc;
        dwc = fmaf(s_ds[v], ldf(xr, (size_t)v * h + c), dwc);
      }
      stf(dx, ((size_t)r * V + v) * h + c, o);
    }
    if (dw && wv) atomicAdd(dw + c, dwc);
  }
  (void)s_D;
}

// =================================================================================================
// rowdot: y[m] = x[m,:].w + b    (N = 1 heads; fp32 outputs)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    rowdot_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wv, const float* __restrict__ bias,
                      float* __restrict__ y, int M, int h) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    float a = 0.f;
    for (int c = lane; c < h; c += 32) a = fmaf(ldf(x, (size_t)r * h + c), wv[c], a);
    a = warp_sum(a);
    if (lane == 0) y[r] = a + (bias ? bias[0] : 0.f);
  }
}
template <typename T, int NE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
    rowdot_bwd_kernel(const float* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ wv,
                      T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ dbias, int M, int h) {
  extern __shared__ float sm[];  // [h + 1]
  for (int c = threadIdx.x; c < h + 1; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float pw[NE];
#pragma unroll
  for (int i = 0; i < NE; i++) pw[i] = 0.f;
  float pb = 0.f;
  for (int r = blockIdx.x * ROW_WARPS + w; r < M; r += gridDim.x * ROW_WARPS) {
    const float g = dy[r];
    pb += g;
#pragma unroll
    for (int i = 0; i < NE; i++) {
      const int c = lane + 32 * i;
      if (c < h) {
        pw[i] = fmaf(g, ldf(x, (size_t)r * h + c), pw[i]);
        stf(dx, (size_t)r * h + c, g * wv[c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int c = lane + 32 * i;
    if (c < h) atomicAdd(&sm[c], pw[i]);
  }
  if (lane == 0) atomicAdd(&sm[h], pb);
  __syncthreads();
  for (int c = threadIdx.x; c < h; c += blockDim.x) atomicAdd(dw + c, sm[c]);
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, sm[h]);
}

// =================================================================================================
// colsum: out[n] += sum_m x[m,n]   (bias gradients)
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int M, int N,
                                                     long ld, int rows_per_cta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per_cta;
  const int m1 = min(M, m0 + rows_per_cta);
  float a = 0.f;
  if (n < N)
    for (int m = m0 + ty; m < m1; m += 8) a += ldf(x, (size_t)m * ld + n);
  sm[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sm[i][tx];
    atomicAdd(out + n, t);
  }
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    stf(out, i, ldf(in, i));
}

template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ pre, T* __restrict__ dz, size_t n,
                               int act, float p, uint32_t salt, const unsigned long long* seed_ptr) {
  pdl_trigger();
  pdl_wait();
  const Dropout dr = make_dropout(p, seed_ptr, salt);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float z = ldf(pre, i);
    float d = 1.f;
    if (act == 1) d = gelu_grad_f(z);
    else if (act == 2) d = z > 0.f ? 1.f : 0.f;
    stf(dz, i, ldf(dy, i) * dr.scale(i) * d);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                           T* __restrict__ out, size_t n, float scale) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = ldf(a, i) + ldf(b, i);
    if (c) v += ldf(c, i);
    stf(out, i, v * scale);
  }
}

template <typename T>
__global__ void copy2d_kernel(const T* __restrict__ src, long src_ld, T* __restrict__ dst, long dst_ld, int rows,
                              int cols) {
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / cols, c = i % cols;
    dst[r * dst_ld + c] = src[r * src_ld + c];
  }
}

__global__ void segsum_kernel(const float* __restrict__ vals, const long long* __restrict__ seg,
                              const float* __restrict__ seg_scale, float* __restrict__ out, int R) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
    const long long s = seg[i];
    atomicAdd(out + s, vals[i] * (seg_scale ? seg_scale[s] : 1.f));
  }
}

// out[i] = (src ? src[idx ? idx[i] : i] : 1) * (scale ? scale[i] : 1): per-row KD weights (MKTD sample weight of
// the row's sample x the padding mask / count correction of graph_index.pad_batch)
__global__ void row_weights_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                   const float* __restrict__ scale, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 1.f;
  if (src) {
    const long long j = idx ? idx[i] : (long long)i;
    v = j >= 0 ? src[j] : 0.f;
  }
  out[i] = v * (scale ? scale[i] : 1.f);
}

// out[i] = valid[i] ? x[i] : fill   (object-grounding logits: -inf on padded object slots; backward: fill = 0)
__global__ void mask_fill_kernel(const float* __restrict__ x, const unsigned char* __restrict__ valid,
                                 float* __restrict__ out, int n, float fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = valid[i] ? x[i] : fill;
}

__global__ void exp_decay_kernel(const float* __restrict__ in, float* __restrict__ out, int n, float rate) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = expf(-rate * in[i]);
}

// out = 1 - (x - min) / (max - min), single CTA
__global__ void __launch_bounds__(256) invert_norm_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  __shared__ float red[32];
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    mn = fminf(mn, in[i]);
    mx = fmaxf(mx, in[i]);
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = 1.f - (in[i] - mn) / (mx - mn);
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
#define DISPATCH_T(dt, ...)                                    \
  if ((dt) == MAGIC_F32) {                                     \
    typedef float T;                                           \
    __VA_ARGS__;                                               \
  } else if ((dt) == MAGIC_BF16) {                             \
    typedef __nv_bfloat16 T;                                   \
    __VA_ARGS__;                                               \
  } else {                                                     \
    magic_set_error("bad dtype %d", (int)(dt));                \
    return MAGIC_ERR_ARG;                                      \
  }

#define DISPATCH_NCH(h, ...)           \
  if ((h) == 128) {                    \
    constexpr int NCH = 1;             \
    __VA_ARGS__;                       \
  } else if ((h) == 256) {             \
    constexpr int NCH = 2;             \
    __VA_ARGS__;                       \
  } else if ((h) == 384) {             \
    constexpr int NCH = 3;             \
    __VA_ARGS__;                       \
  } else if ((h) == 512) {             \
    constexpr int NCH = 4;             \
    __VA_ARGS__;                       \
  } else if ((h) == 768) {             \
    constexpr int NCH = 6;             \
    __VA_ARGS__;                       \
  } else {                             \
    magic_set_error("layer norm: vector path has no instance for h=%d", (int)(h)); \
    return MAGIC_ERR_ARG;              \
  }

// every pointer of a vectorised row kernel must be 16-byte aligned (NULL = absent operand)
static inline bool vec_ok(const void* a, const void* b, const void* c, const void* d, const void* e) {
  return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c) | ((uintptr_t)d) | ((uintptr_t)e)) & 15) == 0;
}
static inline bool vec_h(int h) { return h == 128 || h == 256 || h == 384 || h == 512 || h == 768; }

#define DISPATCH_NE(h, ...)            \
  if ((h) <= 64) {                     \
    constexpr int NE = 2;              \
    __VA_ARGS__;                       \
  } else if ((h) <= 128) {             \
    constexpr int NE = 4;              \
    __VA_ARGS__;                       \
  } else if ((h) <= 256) {             \
    constexpr int NE = 8;              \
    __VA_ARGS__;                       \
  } else if ((h) <= 384) {             \
    constexpr int NE = 12;             \
    __VA_ARGS__;                       \
  } else {                             \
    constexpr int NE = 24;             \
    __VA_ARGS__;                       \
  }

template <typename T, int NE>
int launch_posfuse_bwd(int grid, size_t smem, cudaStream_t st, const void* dy, const long long* idx, const float* f,
                       const float* W, const float* b, const float* gamma, const float* stats, float* demb,
                       float* dcst, float* dW, float* db, float* dgamma, float* dbeta, int M, int h, int K) {
  if (smem > 48 * 1024)
    MAGIC_CUDA(cudaFuncSetAttribute(posfuse_bwd_kernel<T, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "magic_posfuse_bwd");
  posfuse_bwd_kernel<T, NE><<<grid, ROW_WARPS * 32, smem, st>>>((const T*)dy, idx, f, W, b, gamma, stats, demb, dcst, dW,
                                                                 db, dgamma, dbeta, M, h, K);
  return MAGIC_OK;
}

extern "C" {

int magic_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, float* stats,
                 int M, int h, float eps, int dtype, float p_in, unsigned salt_in, float p_out, unsigned salt_out,
                 const unsigned long long* seed_ptr, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32, "magic_ln_fwd: hidden size %d unsupported (max %d)", h, MAXE * 32);
  if (M <= 0) return MAGIC_OK;
  if (vec_h(h) && vec_ok(x, res, y, gamma, beta)) {
    DISPATCH_T(dtype, DISPATCH_NCH(h, (magic_launch(ln_fwd_vec_kernel<T, NCH>, dim3(row_grid(M)), dim3(ROW_WARPS * 32),
                          0, st, (const T*)x, (const T*)res, gamma, beta, (T*)y, stats, M, eps, p_in, salt_in, p_out,
                          salt_out, seed_ptr))));
    MAGIC_CHECK_LAUNCH("magic_ln_fwd");
    return MAGIC_OK;
  }
  DISPATCH_T(dtype, DISPATCH_NE(h, (magic_launch(ln_fwd_kernel<T, NE>, dim3(row_grid(M)), dim3(ROW_WARPS * 32), 0, st, 
                        (const T*)x, (const T*)res, gamma, beta, (T*)y, stats, M, h, eps, p_in, salt_in, p_out,
                        salt_out, seed_ptr))));
  MAGIC_CHECK_LAUNCH("magic_ln_fwd");
  return MAGIC_OK;
}

int magic_ln_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* stats, void* dx,
                 void* dres, float* dgamma, float* dbeta, int M, int h, int dtype, float p_in, unsigned salt_in,
                 float p_out, unsigned salt_out, const unsigned long long* seed_ptr, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32, "magic_ln_bwd: hidden size %d unsupported", h);
  if (M <= 0) return MAGIC_OK;
  const size_t smem = 2 * (size_t)h * sizeof(float);
  if (vec_h(h) && vec_ok(dy, x, res, dx, dres) && vec_ok(gamma, nullptr, nullptr, nullptr, nullptr)) {
    // several rows per warp at the narrow widths: fewer CTAs -> fewer column reductions into dgamma / dbeta
    static int rpw_env = -1;
    if (rpw_env < 0) {
      const char* e = getenv("MAGIC_LN_RPW");  // measurement aid
      rpw_env = e ? atoi(e) : 0;
    }
    const int rpw = rpw_env > 0 ? rpw_env : (h <= 256 ? 2 : 1);  // measured at h = 128: 2 rows/warp 5.9 us, 1: 7.3, 4: 8.2
    int grid = (M + ROW_WARPS * rpw - 1) / (ROW_WARPS * rpw);
    const int cap = magic_num_sms() * 4;
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    const size_t smem_v = (size_t)ROW_WARPS * 2 * h * sizeof(float);  // <= 48 KB at h = 768
    DISPATCH_T(dtype, DISPATCH_NCH(h, (magic_launch(ln_bwd_vec_kernel<T, NCH>, dim3(grid), dim3(ROW_WARPS * 32),
                          smem_v, st, (const T*)dy, (const T*)x, (const T*)res, gamma, stats, (T*)dx, (T*)dres, dgamma,
                          dbeta, M, p_in, salt_in, p_out, salt_out, seed_ptr))));
    MAGIC_CHECK_LAUNCH("magic_ln_bwd");
    return MAGIC_OK;
  }
  DISPATCH_T(dtype, DISPATCH_NE(h, (magic_launch(ln_bwd_kernel<T, NE>, dim3(row_grid(M)), dim3(ROW_WARPS * 32), smem, st, 
                        (const T*)dy, (const T*)x, (const T*)res, gamma, stats, (T*)dx, (T*)dres, dgamma, dbeta, M, h,
                        p_in, salt_in, p_out, salt_out, seed_ptr))));
  MAGIC_CHECK_LAUNCH("magic_ln_bwd");
  return MAGIC_OK;
}

int magic_embed_ln_fwd(const long long* ids, const float* word, const float* pos, const float* type0,
                       const float* gamma, const float* beta, void* y, float* stats, int M, int L, int h, float eps,
                       int dtype, float p_out, unsigned salt_out, const unsigned long long* seed_ptr,
                       cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32, "magic_embed_ln_fwd: hidden size %d unsupported", h);
  if (M <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, DISPATCH_NE(h, (magic_launch(embed_ln_fwd_kernel<T, NE>, dim3(row_grid(M)), dim3(ROW_WARPS * 32), 0, st, 
                        ids, word, pos, type0, gamma, beta, (T*)y, stats, M, L, h, eps, p_out, salt_out, seed_ptr))));
  MAGIC_CHECK_LAUNCH("magic_embed_ln_fwd");
  return MAGIC_OK;
}

int magic_embed_ln_bwd(const void* dy, const long long* ids, const float* word, const float* pos,
                       const float* type0, const float* gamma, const float* stats, float* dword, float* dpos,
                       float* dtype0, float* dgamma, float* dbeta, int M, int L, int h, int dtype, float p_out,
                       unsigned salt_out, const unsigned long long* seed_ptr, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32, "magic_embed_ln_bwd: hidden size %d unsupported", h);
  if (M <= 0) return MAGIC_OK;
  const size_t smem = 3 * (size_t)h * sizeof(float);
  DISPATCH_T(dtype, DISPATCH_NE(h, (magic_launch(embed_ln_bwd_kernel<T, NE>, dim3(row_grid(M)), dim3(ROW_WARPS * 32), smem, st, 
                        (const T*)dy, ids, word, pos, type0, gamma, stats, dword, dpos, dtype0, dgamma, dbeta, M, L, h,
                        p_out, salt_out, seed_ptr))));
  MAGIC_CHECK_LAUNCH("magic_embed_ln_bwd");
  return MAGIC_OK;
}

int magic_posfuse_fwd(const void* xin, const long long* idx, const float* emb, const float* cst, const float* f,
                      const float* W, const float* b, const float* gamma, const float* beta, void* y, float* stats,
                      int M, int h, int K, float eps, int dtype, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32 && K > 0 && K <= MAXK, "magic_posfuse_fwd: h=%d K=%d unsupported", h, K);
  if (M <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, DISPATCH_NE(h, (posfuse_fwd_kernel<T, NE><<<row_grid(M), ROW_WARPS * 32, 0, st>>>(
                        (const T*)xin, idx, emb, cst, f, W, b, gamma, beta, (T*)y, stats, M, h, K, eps))));
  MAGIC_CHECK_LAUNCH("magic_posfuse_fwd");
  return MAGIC_OK;
}

int magic_posfuse_bwd(const void* dy, const long long* idx, const float* f, const float* W, const float* b,
                      const float* gamma, const float* stats, float* demb, float* dcst, float* dW, float* db,
                      float* dgamma, float* dbeta, int M, int h, int K, int dtype, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32 && K > 0 && K <= MAXK, "magic_posfuse_bwd: h=%d K=%d unsupported", h, K);
  if (M <= 0) return MAGIC_OK;
  const size_t smem = ((size_t)h * K + 4 * (size_t)h) * sizeof(float);
  int grid = row_grid(M);
  if (grid > magic_num_sms()) grid = magic_num_sms();
  int rc = MAGIC_OK;
  DISPATCH_T(dtype, DISPATCH_NE(h, (rc = launch_posfuse_bwd<T, NE>(grid, smem, st, dy, idx, f, W, b, gamma, stats, demb,
                                                                  dcst, dW, db, dgamma, dbeta, M, h, K))));
  if (rc) return rc;
  MAGIC_CHECK_LAUNCH("magic_posfuse_bwd");
  return MAGIC_OK;
}

int magic_gather_rows(const void* src, const long long* idx, void* out, int R, int h, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, (magic_launch(gather_rows_kernel<T>, dim3(row_grid(R)), dim3(ROW_WARPS * 32), 0, st, (const T*)src, idx, (T*)out, R, h)));
  MAGIC_CHECK_LAUNCH("magic_gather_rows");
  return MAGIC_OK;
}

int magic_scatter_rows(const void* dout, const long long* idx, void* dsrc, int R, int n_src_rows, int h, int dtype,
                       cudaStream_t st) {
  const size_t esz = dtype == MAGIC_BF16 ? 2 : 4;
  MAGIC_CUDA(cudaMemsetAsync(dsrc, 0, (size_t)n_src_rows * h * esz, st), "magic_scatter_rows(memset)");
  if (R <= 0) return MAGIC_OK;
  DISPATCH_T(dtype,
             (magic_launch(scatter_rows_kernel<T>, dim3(row_grid(R)), dim3(ROW_WARPS * 32), 0, st, (const T*)dout, idx, (T*)dsrc, R, h)));
  MAGIC_CHECK_LAUNCH("magic_scatter_rows");
  return MAGIC_OK;
}

int magic_pano_fuse_fwd(const void* x, const float* w, const float* bias, const long long* lens, void* fused,
                        float* probs, int R, int V, int h, int dtype, cudaStream_t st) {
  MAGIC_CHECK_ARG(V > 0 && V <= PF_MAXV, "magic_pano_fuse_fwd: V=%d unsupported (max %d)", V, PF_MAXV);
  if (R <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, (pano_fuse_fwd_kernel<T><<<R, 128, 0, st>>>((const T*)x, w, bias, lens, (T*)fused, probs, V, h)));
  MAGIC_CHECK_LAUNCH("magic_pano_fuse_fwd");
  return MAGIC_OK;
}

int magic_pano_fuse_bwd(const void* dfused, const void* x, const float* w, const long long* lens, const float* probs,
                        void* dx, float* dw, float* dbias, int R, int V, int h, int dtype, cudaStream_t st) {
  MAGIC_CHECK_ARG(V > 0 && V <= PF_MAXV, "magic_pano_fuse_bwd: V=%d unsupported", V);
  if (R <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, (pano_fuse_bwd_kernel<T><<<R, 128, 0, st>>>((const T*)dfused, (const T*)x, w, lens, probs, (T*)dx,
                                                               dw, dbias, V, h)));
  MAGIC_CHECK_LAUNCH("magic_pano_fuse_bwd");
  return MAGIC_OK;
}

int magic_rowdot_fwd(const void* x, const float* w, const float* bias, float* y, int M, int h, int dtype,
                     cudaStream_t st) {
  if (M <= 0) return MAGIC_OK;
  DISPATCH_T(dtype, (rowdot_fwd_kernel<T><<<row_grid(M), ROW_WARPS * 32, 0, st>>>((const T*)x, w, bias, y, M, h)));
  MAGIC_CHECK_LAUNCH("magic_rowdot_fwd");
  return MAGIC_OK;
}

int magic_rowdot_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int M, int h,
                     int dtype, cudaStream_t st) {
  MAGIC_CHECK_ARG(h > 0 && h <= MAXE * 32, "magic_rowdot_bwd: hidden size %d unsupported", h);
  if (M <= 0) return MAGIC_OK;
  const size_t smem = ((size_t)h + 1) * sizeof(float);
  DISPATCH_T(dtype, DISPATCH_NE(h, (rowdot_bwd_kernel<T, NE><<<row_grid(M), ROW_WARPS * 32, smem, st>>>(dy, (const T*)x, w, (T*)dx, dw,
                                                                                   dbias, M, h))));
  MAGIC_CHECK_LAUNCH("magic_rowdot_bwd");
  return MAGIC_OK;
}

int magic_colsum(const void* x, float* out, int M, int N, long ld, int dtype, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MAGIC_OK;
  int chunks = (M + 255) / 256;
  const int max_chunks = 4 * magic_num_sms();
  if (chunks > max_chunks) chunks = max_chunks;
  const int rows_per_cta = (M + chunks - 1) / chunks;
  dim3 grid((N + 31) / 32, (M + rows_per_cta - 1) / rows_per_cta);
  DISPATCH_T(dtype, (magic_launch(colsum_kernel<T>, dim3(grid), dim3(256), 0, st, (const T*)x, out, M, N, ld, rows_per_cta)));
  MAGIC_CHECK_LAUNCH("magic_colsum");
  return MAGIC_OK;
}

int magic_act_bwd(const void* dy, const void* pre, void* dz, long long n, int act, int dtype, float drop_p,
                  unsigned salt, const unsigned long long* seed_ptr, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  long long blocks = (n + 1023) / 1024;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  DISPATCH_T(dtype, (magic_launch(act_bwd_kernel<T>, dim3((int)blocks), dim3(256), 0, st, (const T*)dy, (const T*)pre, (T*)dz, (size_t)n, act,
                                                                    drop_p, salt, seed_ptr)));
  MAGIC_CHECK_LAUNCH("magic_act_bwd");
  return MAGIC_OK;
}

int magic_add(const void* a, const void* b, const void* c, void* out, long long n, float scale, int dtype,
              cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  long long blocks = (n + 1023) / 1024;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  DISPATCH_T(dtype, (magic_launch(add_kernel<T>, dim3((int)blocks), dim3(256), 0, st, (const T*)a, (const T*)b, (const T*)c, (T*)out,
                                                                (size_t)n, scale)));
  MAGIC_CHECK_LAUNCH("magic_add");
  return MAGIC_OK;
}

int magic_copy2d(const void* src, long src_ld, void* dst, long dst_ld, int rows, int cols, int dtype,
                 cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return MAGIC_OK;
  long long blocks = ((long long)rows * cols + 255) / 256;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  DISPATCH_T(dtype, (copy2d_kernel<T><<<(int)blocks, 256, 0, st>>>((const T*)src, src_ld, (T*)dst, dst_ld, rows, cols)));
  MAGIC_CHECK_LAUNCH("magic_copy2d");
  return MAGIC_OK;
}

int magic_segsum(const float* vals, const long long* seg, const float* seg_scale, float* out, int R, int n_seg,
                 cudaStream_t st) {
  MAGIC_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_seg, st), "magic_segsum");
  if (R <= 0) return MAGIC_OK;
  segsum_kernel<<<(R + 255) / 256, 256, 0, st>>>(vals, seg, seg_scale, out, R);
  MAGIC_CHECK_LAUNCH("magic_segsum");
  return MAGIC_OK;
}

int magic_exp_decay(const float* in, float* out, int n, float rate, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  exp_decay_kernel<<<(n + 255) / 256, 256, 0, st>>>(in, out, n, rate);
  MAGIC_CHECK_LAUNCH("magic_exp_decay");
  return MAGIC_OK;
}

int magic_row_weights(const float* src, const long long* idx, const float* scale, float* out, int n,
                      cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  row_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, idx, scale, out, n);
  MAGIC_CHECK_LAUNCH("magic_row_weights");
  return MAGIC_OK;
}

int magic_mask_fill(const float* x, const unsigned char* valid, float* out, int n, float fill, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  mask_fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, valid, out, n, fill);
  MAGIC_CHECK_LAUNCH("magic_mask_fill");
  return MAGIC_OK;
}

int magic_invert_norm(const float* in, float* out, int n, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  invert_norm_kernel<<<1, 256, 0, st>>>(in, out, n);
  MAGIC_CHECK_LAUNCH("magic_invert_norm");
  return MAGIC_OK;
}

int magic_cast(const void* in, int in_dt, void* out, int out_dt, long long n, cudaStream_t st) {
  if (n <= 0) return MAGIC_OK;
  long long blocks = (n + 1023) / 1024;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  if (in_dt == MAGIC_F32 && out_dt == MAGIC_BF16)
    cast_kernel<float, __nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, (size_t)n);
  else if (in_dt == MAGIC_BF16 && out_dt == MAGIC_F32)
    cast_kernel<__nv_bfloat16, float><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, (size_t)n);
  else if (in_dt == MAGIC_F32 && out_dt == MAGIC_F32)
    cast_kernel<float, float><<<(int)blocks, 256, 0, st>>>((const float*)in, (float*)out, (size_t)n);
  else {
    magic_set_error("magic_cast: unsupported dtype pair %d -> %d", in_dt, out_dt);
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_cast");
  return MAGIC_OK;
}

}  // extern "C"
