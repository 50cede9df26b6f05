// Fused multi-head attention for the MAGIC encoders (head_dim = 64, Lq/Lk <= 320):
//   S = scale * Q K^T + (w_sprel * dist + b_sprel)      graph-distance bias added inside the softmax tile
//   keys j >= key_len[b] are excluded (== the reference's -10000 / -inf additive masks in fp32)
//   P = softmax(S);  O = dropout(P) V;  Pbar = mean_heads(P)  (KD attention map, optional, fp32)
// All K/V of one (batch, head) live in shared memory (fp32, rows padded to 68 floats: 16-byte aligned and
// bank-conflict-free for 128-bit loads when lanes sweep rows).  One warp owns TWO query (or key) rows and
// shares every K/V shared-memory load between them; lanes sweep keys for the score/softmax phase and pairs of
// head-dim columns for the P*V phase (0.375 LDS per FMA).  One CTA walks all heads of its rows so the
// head-mean map is accumulated in registers and written once (no atomics).
// Backward = two passes with the same structure (dQ by query rows, dK/dV by key rows); P is recomputed from
// the saved log-sum-exp, the softmax row term delta is produced by pass 1.
#include "common.cuh"
#include "attention.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int D = 64;
constexpr int DP = 68;       // padded smem row (floats)
constexpr int NW = 8;        // warps per CTA
constexpr int RPW = 2;       // rows per warp (processed together)
constexpr int ROWS = NW * RPW;
constexpr int MAXJ = 10;     // ceil(320 / 32)
constexpr int MAXL = 32 * MAXJ;


__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// dst[rows_pad][DP] <- src[(row0 + r) * ld + head*64 + d]; rows in [rows, rows_pad) are zero-filled
template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, long ld, int head, int rows, int rows_pad,
                                          int row0) {
  for (int e = threadIdx.x; e < rows_pad * (D / 2); e += blockDim.x) {
    const int r = e >> 5, d = (e & 31) * 2;
    float2 v = make_float2(0.f, 0.f);
    if (r < rows) {
      const size_t g = (size_t)(row0 + r) * ld + head * D + d;
      v.x = ldf(src, g);
      v.y = ldf(src, g + 1);
    }
    *reinterpret_cast<float2*>(dst + r * DP + d) = v;
  }
}

// one row of 64 values -> per-warp smem vector (lane writes 2 consecutive floats)
template <typename T>
__device__ __forceinline__ void load_vec(float* dst, const T* src, int lane) {
  *reinterpret_cast<float2*>(dst + 2 * lane) = make_float2(ldf(src, 2 * lane), ldf(src, 2 * lane + 1));
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// a0 = x0 . row, a1 = x1 . row   (x0/x1: per-warp broadcast vectors, row: lane-varying smem row)
__device__ __forceinline__ void dot2_row(const float* x0, const float* x1, const float* row, float& a0, float& a1) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const float4* p0 = reinterpret_cast<const float4*>(x0);
  const float4* p1 = reinterpret_cast<const float4*>(x1);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < D / 4; i++) {
    const float4 kv = r4[i];
    s0 += dot4(p0[i], kv);
    s1 += dot4(p1[i], kv);
  }
  a0 = s0;
  a1 = s1;
}

// o0 += sum_j c0[j] * M[j][2*lane..], o1 likewise with c1  (c0/c1: per-warp coefficient vectors, n % 4 == 0)
__device__ __forceinline__ void accum2(const float* c0, const float* c1, const float* M, int n, int lane, float2& o0,
                                       float2& o1) {
  const float4* q0 = reinterpret_cast<const float4*>(c0);
  const float4* q1 = reinterpret_cast<const float4*>(c1);
  for (int j = 0; j < n; j += 4) {
    const float4 a = q0[j >> 2], b = q1[j >> 2];
    const float2 m0 = *reinterpret_cast<const float2*>(M + (size_t)(j + 0) * DP + 2 * lane);
    const float2 m1 = *reinterpret_cast<const float2*>(M + (size_t)(j + 1) * DP + 2 * lane);
    const float2 m2 = *reinterpret_cast<const float2*>(M + (size_t)(j + 2) * DP + 2 * lane);
    const float2 m3 = *reinterpret_cast<const float2*>(M + (size_t)(j + 3) * DP + 2 * lane);
    o0.x = fmaf(a.x, m0.x, fmaf(a.y, m1.x, fmaf(a.z, m2.x, fmaf(a.w, m3.x, o0.x))));
    o0.y = fmaf(a.x, m0.y, fmaf(a.y, m1.y, fmaf(a.z, m2.y, fmaf(a.w, m3.y, o0.y))));
    o1.x = fmaf(b.x, m0.x, fmaf(b.y, m1.x, fmaf(b.z, m2.x, fmaf(b.w, m3.x, o1.x))));
    o1.y = fmaf(b.x, m0.y, fmaf(b.y, m1.y, fmaf(b.z, m2.y, fmaf(b.w, m3.y, o1.y))));
  }
}

template <typename T>
__device__ __forceinline__ void store2(T* row, int lane, float2 v) {
  stf(row, 2 * lane, v.x);
  stf(row, 2 * lane + 1, v.y);
}

template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_fwd_kernel(AttnParams P) {
  extern __shared__ __align__(16) float sm[];
  const int Lkp = round4(P.Lk);
  float* Ks = sm;                            // [Lkp][DP]
  float* Vs = Ks + (size_t)Lkp * DP;         // [Lkp][DP]
  float* qs = Vs + (size_t)Lkp * DP;         // [NW][RPW][D]
  float* ps = qs + NW * RPW * D;             // [NW][RPW][MAXL]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const int klen4 = round4(klen);
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;
  const int i0 = row0 + w * RPW;             // first row of this warp
  const bool act0 = i0 < P.Lq, act1 = i0 + 1 < P.Lq;
  const int ia = min(i0, P.Lq - 1), ib = min(i0 + 1, P.Lq - 1);  // clamped (duplicates are never stored)
  float* q0 = qs + (w * RPW + 0) * D;
  float* q1 = qs + (w * RPW + 1) * D;
  float* p0 = ps + (w * RPW + 0) * MAXL;
  float* p1 = ps + (w * RPW + 1) * MAXL;
  float pb0[MAXJ], pb1[MAXJ];
#pragma unroll
  for (int jj = 0; jj < MAXJ; jj++) pb0[jj] = pb1[jj] = 0.f;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Ks, (const T*)P.k, P.k_ld, hd, P.Lk, Lkp, b * P.Lk);
    load_tile(Vs, (const T*)P.v, P.v_ld, hd, P.Lk, Lkp, b * P.Lk);
    __syncthreads();
    if (!act0) continue;  // warp-uniform; barriers above are reached by every warp on the next iteration
    const size_t qa = (size_t)(b * P.Lq + ia), qb = (size_t)(b * P.Lq + ib);
    __syncwarp();
    load_vec(q0, (const T*)P.q + qa * P.q_ld + hd * D, lane);
    load_vec(q1, (const T*)P.q + qb * P.q_ld + hd * D, lane);
    __syncwarp();
    float s0[MAXJ], s1[MAXJ];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = jj * 32 + lane;
      s0[jj] = s1[jj] = -INFINITY;
      if (jj * 32 < P.Lk && j < klen && j != P.key_skip) {
        float a0, a1;
        dot2_row(q0, q1, Ks + (size_t)j * DP, a0, a1);
        a0 *= P.scale;
        a1 *= P.scale;
        if (P.dists) {
          a0 += sw * P.dists[((size_t)b * P.Lq + ia) * P.Lk + j] + sb;
          a1 += sw * P.dists[((size_t)b * P.Lq + ib) * P.Lk + j] + sb;
        }
        s0[jj] = a0;
        s1[jj] = a1;
        mx0 = fmaxf(mx0, a0);
        mx1 = fmaxf(mx1, a1);
      }
    }
    mx0 = warp_max(mx0);
    mx1 = warp_max(mx1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const float e0 = (s0[jj] == -INFINITY) ? 0.f : __expf(s0[jj] - mx0);
      const float e1 = (s1[jj] == -INFINITY) ? 0.f : __expf(s1[jj] - mx1);
      s0[jj] = e0;
      s1[jj] = e1;
      sum0 += e0;
      sum1 += e1;
    }
    sum0 = warp_sum(sum0);
    sum1 = warp_sum(sum1);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    if (lane == 0) {
      P.lse[((size_t)b * P.H + hd) * P.Lq + ia] = mx0 + __logf(sum0);
      if (act1) P.lse[((size_t)b * P.H + hd) * P.Lq + ib] = mx1 + __logf(sum1);
    }
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = jj * 32 + lane;
      if (jj * 32 < klen4 && j < klen4) {
        const float pa = s0[jj] * inv0, pc = s1[jj] * inv1;  // exactly 0 for masked / padded keys
        pb0[jj] += pa * invH;
        pb1[jj] += pc * invH;
        const size_t di = (((size_t)b * P.H + hd) * P.Lq) * P.Lk + j;
        p0[j] = pa * dr.scale(di + (size_t)ia * P.Lk);
        p1[j] = pc * dr.scale(di + (size_t)ib * P.Lk);
      }
    }
    __syncwarp();
    float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
    accum2(p0, p1, Vs, klen4, lane, o0, o1);
    store2((T*)P.out + qa * (size_t)(P.H * D) + hd * D, lane, o0);
    if (act1) store2((T*)P.out + qb * (size_t)(P.H * D) + hd * D, lane, o1);
  }
  if (P.pbar && act0) {
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = jj * 32 + lane;
      if (j < P.Lk) {
        P.pbar[(size_t)b * P.pbar_bs + (size_t)ia * P.pbar_rs + j] = pb0[jj];
        if (act1) P.pbar[(size_t)b * P.pbar_bs + (size_t)ib * P.pbar_rs + j] = pb1[jj];
      }
    }
  }
}

// ---- backward pass 1: dQ, delta, d(sprel) ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_bwd_q_kernel(AttnParams P) {
  extern __shared__ __align__(16) float sm[];
  const int Lkp = round4(P.Lk);
  float* Ks = sm;
  float* Vs = Ks + (size_t)Lkp * DP;
  float* qs = Vs + (size_t)Lkp * DP;     // [NW][RPW][D]
  float* gs = qs + NW * RPW * D;         // [NW][RPW][D]   dO rows
  float* ps = gs + NW * RPW * D;         // [NW][RPW][MAXL] dS rows
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const int klen4 = round4(klen);
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;
  const int i0 = row0 + w * RPW;
  const bool act0 = i0 < P.Lq, act1 = i0 + 1 < P.Lq;
  const int ia = min(i0, P.Lq - 1), ib = min(i0 + 1, P.Lq - 1);
  float* q0 = qs + (w * RPW + 0) * D;
  float* q1 = qs + (w * RPW + 1) * D;
  float* g0 = gs + (w * RPW + 0) * D;
  float* g1 = gs + (w * RPW + 1) * D;
  float* d0 = ps + (w * RPW + 0) * MAXL;
  float* d1 = ps + (w * RPW + 1) * MAXL;
  float acc_dw = 0.f, acc_db = 0.f;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Ks, (const T*)P.k, P.k_ld, hd, P.Lk, Lkp, b * P.Lk);
    load_tile(Vs, (const T*)P.v, P.v_ld, hd, P.Lk, Lkp, b * P.Lk);
    __syncthreads();
    if (!act0) continue;
    const size_t qa = (size_t)(b * P.Lq + ia), qb = (size_t)(b * P.Lq + ib);
    __syncwarp();
    load_vec(q0, (const T*)P.q + qa * P.q_ld + hd * D, lane);
    load_vec(q1, (const T*)P.q + qb * P.q_ld + hd * D, lane);
    load_vec(g0, (const T*)P.dout + qa * (size_t)(P.H * D) + hd * D, lane);
    load_vec(g1, (const T*)P.dout + qb * (size_t)(P.H * D) + hd * D, lane);
    __syncwarp();
    const float lse0 = P.lse[((size_t)b * P.H + hd) * P.Lq + ia];
    const float lse1 = P.lse[((size_t)b * P.H + hd) * P.Lq + ib];
    float pa[MAXJ], pc[MAXJ], ga[MAXJ], gc[MAXJ];
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = jj * 32 + lane;
      pa[jj] = pc[jj] = ga[jj] = gc[jj] = 0.f;
      if (jj * 32 < P.Lk && j < klen && j != P.key_skip) {
        float a0, a1, e0, e1;
        dot2_row(q0, q1, Ks + (size_t)j * DP, a0, a1);
        dot2_row(g0, g1, Vs + (size_t)j * DP, e0, e1);
        a0 *= P.scale;
        a1 *= P.scale;
        if (P.dists) {
          a0 += sw * P.dists[((size_t)b * P.Lq + ia) * P.Lk + j] + sb;
          a1 += sw * P.dists[((size_t)b * P.Lq + ib) * P.Lk + j] + sb;
        }
        pa[jj] = __expf(a0 - lse0);
        pc[jj] = __expf(a1 - lse1);
        const size_t di = (((size_t)b * P.H + hd) * P.Lq) * P.Lk + j;
        e0 *= dr.scale(di + (size_t)ia * P.Lk);
        e1 *= dr.scale(di + (size_t)ib * P.Lk);
        if (P.dpbar) {
          e0 += P.dpbar[(size_t)b * P.pbar_bs + (size_t)ia * P.pbar_rs + j] * invH;
          e1 += P.dpbar[(size_t)b * P.pbar_bs + (size_t)ib * P.pbar_rs + j] * invH;
        }
        ga[jj] = e0;
        gc[jj] = e1;
        dl0 = fmaf(pa[jj], e0, dl0);
        dl1 = fmaf(pc[jj], e1, dl1);
      }
    }
    dl0 = warp_sum(dl0);
    dl1 = warp_sum(dl1);
    if (lane == 0) {
      P.delta[((size_t)b * P.H + hd) * P.Lq + ia] = dl0;
      if (act1) P.delta[((size_t)b * P.H + hd) * P.Lq + ib] = dl1;
    }
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) {
      const int j = jj * 32 + lane;
      if (jj * 32 < klen4 && j < klen4) {
        const float ds0 = pa[jj] * (ga[jj] - dl0), ds1 = pc[jj] * (gc[jj] - dl1);
        d0[j] = ds0;
        d1[j] = ds1;
        if (P.dists && j < klen && j != P.key_skip) {
          acc_dw = fmaf(ds0, P.dists[((size_t)b * P.Lq + ia) * P.Lk + j], acc_dw);
          acc_db += ds0;
          if (act1) {
            acc_dw = fmaf(ds1, P.dists[((size_t)b * P.Lq + ib) * P.Lk + j], acc_dw);
            acc_db += ds1;
          }
        }
      }
    }
    __syncwarp();
    float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
    accum2(d0, d1, Ks, klen4, lane, o0, o1);
    o0.x *= P.scale; o0.y *= P.scale; o1.x *= P.scale; o1.y *= P.scale;
    store2((T*)P.dq + qa * P.dq_ld + hd * D, lane, o0);
    if (act1) store2((T*)P.dq + qb * P.dq_ld + hd * D, lane, o1);
  }
  if (P.dists && P.dsprel) {
    const float tw = block_sum(acc_dw, red);
    const float tb = block_sum(acc_db, red);
    if (threadIdx.x == 0) {
      atomicAdd(P.dsprel, tw);
      atomicAdd(P.dsprel + 1, tb);
    }
  }
}

// ---- backward pass 2: dK, dV (one warp per two key rows, lanes sweep queries) -------------------------
template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_bwd_kv_kernel(AttnParams P) {
  extern __shared__ __align__(16) float sm[];
  const int Lqp = round4(P.Lq);
  float* Qs = sm;                            // [Lqp][DP]
  float* Gs = Qs + (size_t)Lqp * DP;         // [Lqp][DP]  dO
  float* ls = Gs + (size_t)Lqp * DP;         // [Lqp] lse
  float* dls = ls + Lqp;                     // [Lqp] delta
  float* ks = dls + Lqp;                     // [NW][RPW][D]
  float* vs = ks + NW * RPW * D;             // [NW][RPW][D]
  float* pp = vs + NW * RPW * D;             // [NW][RPW][MAXL]  dropped probs
  float* pd = pp + NW * RPW * MAXL;          // [NW][RPW][MAXL]  dS
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;
  const int j0 = row0 + w * RPW;
  const bool act0 = j0 < P.Lk, act1 = j0 + 1 < P.Lk;
  const int ja = min(j0, P.Lk - 1), jb = min(j0 + 1, P.Lk - 1);
  const bool m0 = ja < klen && ja != P.key_skip, m1 = jb < klen && jb != P.key_skip;  // unmasked keys
  float* k0 = ks + (w * RPW + 0) * D;
  float* k1 = ks + (w * RPW + 1) * D;
  float* v0 = vs + (w * RPW + 0) * D;
  float* v1 = vs + (w * RPW + 1) * D;
  float* pp0 = pp + (w * RPW + 0) * MAXL;
  float* pp1 = pp + (w * RPW + 1) * MAXL;
  float* pd0 = pd + (w * RPW + 0) * MAXL;
  float* pd1 = pd + (w * RPW + 1) * MAXL;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Qs, (const T*)P.q, P.q_ld, hd, P.Lq, Lqp, b * P.Lq);
    load_tile(Gs, (const T*)P.dout, (long)P.H * D, hd, P.Lq, Lqp, b * P.Lq);
    for (int i = threadIdx.x; i < Lqp; i += blockDim.x) {
      ls[i] = i < P.Lq ? P.lse[((size_t)b * P.H + hd) * P.Lq + i] : 0.f;
      dls[i] = i < P.Lq ? P.delta[((size_t)b * P.H + hd) * P.Lq + i] : 0.f;
    }
    __syncthreads();
    if (!act0) continue;
    const size_t ka = (size_t)(b * P.Lk + ja), kb = (size_t)(b * P.Lk + jb);
    T* dk0 = (T*)P.dk + ka * P.dk_ld + hd * D;
    T* dk1 = (T*)P.dk + kb * P.dk_ld + hd * D;
    T* dv0 = (T*)P.dv + ka * P.dv_ld + hd * D;
    T* dv1 = (T*)P.dv + kb * P.dv_ld + hd * D;
    if (!m0 && !m1) {  // both keys masked: no gradient
      const float2 z = make_float2(0.f, 0.f);
      store2(dk0, lane, z);
      store2(dv0, lane, z);
      if (act1) {
        store2(dk1, lane, z);
        store2(dv1, lane, z);
      }
      continue;
    }
    __syncwarp();
    load_vec(k0, (const T*)P.k + ka * P.k_ld + hd * D, lane);
    load_vec(k1, (const T*)P.k + kb * P.k_ld + hd * D, lane);
    load_vec(v0, (const T*)P.v + ka * P.v_ld + hd * D, lane);
    load_vec(v1, (const T*)P.v + kb * P.v_ld + hd * D, lane);
    __syncwarp();
#pragma unroll
    for (int ii = 0; ii < MAXJ; ii++) {
      const int i = ii * 32 + lane;
      if (ii * 32 < Lqp && i < Lqp) {
        float c0 = 0.f, c1 = 0.f, e0 = 0.f, e1 = 0.f;  // dropped probs / dS for (i, ja), (i, jb)
        if (i < P.Lq) {
          float a0, a1, g0, g1;
          dot2_row(k0, k1, Qs + (size_t)i * DP, a0, a1);
          dot2_row(v0, v1, Gs + (size_t)i * DP, g0, g1);
          a0 *= P.scale;
          a1 *= P.scale;
          if (P.dists) {
            a0 += sw * P.dists[((size_t)b * P.Lq + i) * P.Lk + ja] + sb;
            a1 += sw * P.dists[((size_t)b * P.Lq + i) * P.Lk + jb] + sb;
          }
          const float pa = m0 ? __expf(a0 - ls[i]) : 0.f, pc = m1 ? __expf(a1 - ls[i]) : 0.f;
          const size_t di = ((((size_t)b * P.H + hd) * P.Lq) + i) * P.Lk;
          const float sc0 = dr.scale(di + ja), sc1 = dr.scale(di + jb);
          g0 *= sc0;
          g1 *= sc1;
          if (P.dpbar) {
            g0 += P.dpbar[(size_t)b * P.pbar_bs + (size_t)i * P.pbar_rs + ja] * invH;
            g1 += P.dpbar[(size_t)b * P.pbar_bs + (size_t)i * P.pbar_rs + jb] * invH;
          }
          c0 = pa * sc0;
          c1 = pc * sc1;
          e0 = pa * (g0 - dls[i]);
          e1 = pc * (g1 - dls[i]);
        }
        pp0[i] = c0;
        pp1[i] = c1;
        pd0[i] = e0;
        pd1[i] = e1;
      }
    }
    __syncwarp();
    float2 ov0 = make_float2(0.f, 0.f), ov1 = make_float2(0.f, 0.f);
    float2 ok0 = make_float2(0.f, 0.f), ok1 = make_float2(0.f, 0.f);
    accum2(pp0, pp1, Gs, Lqp, lane, ov0, ov1);
    accum2(pd0, pd1, Qs, Lqp, lane, ok0, ok1);
    ok0.x *= P.scale; ok0.y *= P.scale; ok1.x *= P.scale; ok1.y *= P.scale;
    store2(dk0, lane, ok0);
    store2(dv0, lane, ov0);
    if (act1) {  // jb masked -> c1 = e1 = 0 -> exact zeros
      store2(dk1, lane, ok1);
      store2(dv1, lane, ov1);
    }
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 227 * 1024) {
    magic_set_error("%s: needs %zu bytes of shared memory (> 227 KB)", name, bytes);
    return MAGIC_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024)
    MAGIC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), name);
  return MAGIC_OK;
}

// call-site state (like magic_gemm_set_sm_budget): key index masked in the launches that follow, -1 = none
int g_key_skip = -1;

AttnParams make_params(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, int B, int H,
                       int Lq, int Lk, const int* key_lens, const float* dists, const float* sprel_w,
                       const float* sprel_b, float scale, float drop_p, unsigned salt,
                       const unsigned long long* seed_ptr) {
  AttnParams P;
  memset(&P, 0, sizeof(P));
  P.q = q; P.k = k; P.v = v;
  P.q_ld = q_ld; P.k_ld = k_ld; P.v_ld = v_ld;
  P.B = B; P.H = H; P.Lq = Lq; P.Lk = Lk;
  P.key_lens = key_lens; P.dists = dists; P.sprel_w = sprel_w; P.sprel_b = sprel_b;
  P.scale = scale; P.drop_p = drop_p; P.salt = salt; P.seed_ptr = seed_ptr;
  P.key_skip = g_key_skip;
  return P;
}

}  // namespace

extern "C" {

int magic_attn_set_key_skip(int key_index) {
  g_key_skip = key_index < 0 ? -1 : key_index;
  return MAGIC_OK;
}

int magic_attn_fwd(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, void* out,
                   float* lse, float* pbar, long pbar_bs, long pbar_rs, int B, int H, int Lq, int Lk,
                   const int* key_lens, const float* dists, const float* sprel_w, const float* sprel_b, float scale,
                   int dtype, float drop_p, unsigned salt, const unsigned long long* seed_ptr, cudaStream_t st) {
  MAGIC_CHECK_ARG(Lq > 0 && Lk > 0 && Lq <= MAXL && Lk <= MAXL, "magic_attn_fwd: Lq=%d Lk=%d unsupported (max %d)",
                  Lq, Lk, MAXL);
  MAGIC_CHECK_ARG(!dists || (sprel_w && sprel_b), "magic_attn_fwd: dists given without sprel_w/sprel_b");
  if (B <= 0) return MAGIC_OK;
  AttnParams P = make_params(q, k, v, q_ld, k_ld, v_ld, B, H, Lq, Lk, key_lens, dists, sprel_w, sprel_b, scale, drop_p,
                             salt, seed_ptr);
  P.out = out; P.lse = lse; P.pbar = pbar; P.pbar_bs = pbar_bs; P.pbar_rs = pbar_rs;
  if (dtype == MAGIC_BF16) {  // tensor-core path (attention_mma.cu) when the shape / alignment allows
    const int rc = attn_mma_fwd(P, st);
    if (rc != MAGIC_ERR_UNSUPPORTED) return rc;
  }
  const size_t smem = ((size_t)2 * round4(Lk) * DP + NW * RPW * D + NW * RPW * MAXL) * sizeof(float);
  dim3 grid((Lq + ROWS - 1) / ROWS, B);
  if (dtype == MAGIC_F32) {
    int rc = set_smem(attn_fwd_kernel<float>, smem, "magic_attn_fwd");
    if (rc) return rc;
    attn_fwd_kernel<float><<<grid, NW * 32, smem, st>>>(P);
  } else if (dtype == MAGIC_BF16) {
    int rc = set_smem(attn_fwd_kernel<__nv_bfloat16>, smem, "magic_attn_fwd");
    if (rc) return rc;
    attn_fwd_kernel<__nv_bfloat16><<<grid, NW * 32, smem, st>>>(P);
  } else {
    magic_set_error("magic_attn_fwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_attn_fwd");
  return MAGIC_OK;
}

int magic_attn_bwd(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, const void* dout,
                   const float* lse, const float* dpbar, long pbar_bs, long pbar_rs, float* delta, void* dq,
                   void* dk, void* dv, long dq_ld, long dk_ld, long dv_ld, float* dsprel, int B, int H, int Lq,
                   int Lk, const int* key_lens, const float* dists, const float* sprel_w, const float* sprel_b,
                   float scale, int dtype, float drop_p, unsigned salt, const unsigned long long* seed_ptr,
                   cudaStream_t st) {
  MAGIC_CHECK_ARG(Lq > 0 && Lk > 0 && Lq <= MAXL && Lk <= MAXL, "magic_attn_bwd: Lq=%d Lk=%d unsupported (max %d)",
                  Lq, Lk, MAXL);
  if (B <= 0) return MAGIC_OK;
  AttnParams P = make_params(q, k, v, q_ld, k_ld, v_ld, B, H, Lq, Lk, key_lens, dists, sprel_w, sprel_b, scale, drop_p,
                             salt, seed_ptr);
  P.lse = const_cast<float*>(lse); P.dout = dout; P.dpbar = dpbar; P.pbar_bs = pbar_bs; P.pbar_rs = pbar_rs;
  P.delta = delta; P.dq = dq; P.dk = dk; P.dv = dv; P.dq_ld = dq_ld; P.dk_ld = dk_ld; P.dv_ld = dv_ld;
  P.dsprel = dsprel;
  if (dtype == MAGIC_BF16) {
    const int rc = attn_mma_bwd(P, st);
    if (rc != MAGIC_ERR_UNSUPPORTED) return rc;
  }
  const size_t smem1 = ((size_t)2 * round4(Lk) * DP + 2 * NW * RPW * D + NW * RPW * MAXL) * sizeof(float);
  const size_t smem2 =
      ((size_t)2 * round4(Lq) * DP + 2 * round4(Lq) + 2 * NW * RPW * D + 2 * NW * RPW * MAXL) * sizeof(float);
  dim3 grid1((Lq + ROWS - 1) / ROWS, B), grid2((Lk + ROWS - 1) / ROWS, B);
  if (dtype == MAGIC_F32) {
    int rc = set_smem(attn_bwd_q_kernel<float>, smem1, "magic_attn_bwd");
    if (rc) return rc;
    rc = set_smem(attn_bwd_kv_kernel<float>, smem2, "magic_attn_bwd");
    if (rc) return rc;
    attn_bwd_q_kernel<float><<<grid1, NW * 32, smem1, st>>>(P);
    attn_bwd_kv_kernel<float><<<grid2, NW * 32, smem2, st>>>(P);
  } else if (dtype == MAGIC_BF16) {
    int rc = set_smem(attn_bwd_q_kernel<__nv_bfloat16>, smem1, "magic_attn_bwd");
    if (rc) return rc;
    rc = set_smem(attn_bwd_kv_kernel<__nv_bfloat16>, smem2, "magic_attn_bwd");
    if (rc) return rc;
    attn_bwd_q_kernel<__nv_bfloat16><<<grid1, NW * 32, smem1, st>>>(P);
    attn_bwd_kv_kernel<__nv_bfloat16><<<grid2, NW * 32, smem2, st>>>(P);
  } else {
    magic_set_error("magic_attn_bwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_attn_bwd");
  return MAGIC_OK;
}


/* One half of the bf16 tensor-core backward, so the caller can run the two halves on different streams: part 1 =
 * query-major kernel (delta, dQ, d sprel), part 2 = key-major kernel (dK, dV) with delta recomputed as dO . O from the
 * forward output `out`.  Only without a KD-map gradient; MAGIC_ERR_UNSUPPORTED (no launch) when the shape / dtype is
 * not covered -- the caller then uses magic_attn_bwd. */
int magic_attn_bwd_part(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, const void* dout,
                        const void* out, const float* lse, float* delta, void* dq, void* dk, void* dv, long dq_ld,
                        long dk_ld, long dv_ld, float* dsprel, int B, int H, int Lq, int Lk, const int* key_lens,
                        const float* dists, const float* sprel_w, const float* sprel_b, float scale, int dtype,
                        float drop_p, unsigned salt, const unsigned long long* seed_ptr, int part, cudaStream_t st) {
  if (dtype != MAGIC_BF16 || Lq <= 0 || Lk <= 0 || Lq > MAXL || Lk > MAXL || (part != 1 && part != 2))
    return MAGIC_ERR_UNSUPPORTED;
  if (B <= 0) return MAGIC_OK;
  AttnParams P = make_params(q, k, v, q_ld, k_ld, v_ld, B, H, Lq, Lk, key_lens, dists, sprel_w, sprel_b, scale, drop_p,
                             salt, seed_ptr);
  P.lse = const_cast<float*>(lse); P.dout = dout; P.dpbar = nullptr; P.pbar_bs = 0; P.pbar_rs = 0;
  P.out = const_cast<void*>(out);
  P.delta = delta; P.dq = dq; P.dk = dk; P.dv = dv; P.dq_ld = dq_ld; P.dk_ld = dk_ld; P.dv_ld = dv_ld;
  P.dsprel = dsprel;
  return attn_mma_bwd_part(P, part, st);
}

}  // extern "C"
