// bf16 tensor-core attention for the MAGIC encoders (head_dim 64, Lq / Lk <= 160), forward and backward.
//
// The sequences are short (text 80/160 tokens, 36-view panoramas, 20/50-node graphs), so a whole key range fits
// one register-resident score tile: one warp owns a 16-row slab of S = Q K^T (mma.sync m16n8k16, bf16 in, fp32
// accumulate), adds the graph-distance bias w*dist+b and the key-padding mask INSIDE the tile, does the softmax
// with quad shuffles, and feeds P straight back into the P V product as the A operand (no shared-memory round
// trip).  K / V / Q tiles are staged by cp.async into 144-byte-pitch shared rows (conflict-free ldmatrix).
//
//   forward     grid (ceil(Lq/64), heads, B); with the KD attention map requested the heads of one query tile are
//               split over a thread-block CLUSTER (2 or 4 CTAs along grid.y, each walking H/cluster heads and
//               accumulating its partial mean_heads(P) in shared memory); after a cluster barrier every CTA sums
//               its slice of the tile over the peers' partials through distributed shared memory in rank order
//               (no atomics, no zero-fill, deterministic) and writes the map once
//   backward 1  (query-major) recomputes P from the saved log-sum-exp, dP = dO V^T, the softmax row term
//               delta = sum_j P (dP*drop + dPbar/H), dS, dQ = dS K, and the sprel affine's gradients
//   backward 2  (key-major) works on the transposed tile S^T = K Q^T so that P^T and dS^T are already the A
//               operands of dV = P^T dO and dK = dS^T Q; no atomics, each dK/dV row is written exactly once
//
// Dropout uses the same stateless hash and element index as the SIMT kernels (attention.cu), so both paths
// generate identical masks.  fp32 activations, longer sequences and unaligned views stay on the SIMT path.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "attention.cuh"
#include "../../include/magic_b200.h"

namespace {

typedef __nv_bfloat16 bf16;

constexpr int D = 64;
constexpr int PITCH = 72;   // bf16 elements per shared row (144 B: 16-byte aligned, ldmatrix conflict-free)
constexpr int TILE = 64;    // rows (queries or keys) per CTA: 4 warps x 16
constexpr int NTHREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void ldsm_x4(const bf16* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(const bf16* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// dst[rows_pad][PITCH] <- src[r * ld + 0..63] for r < rows_valid (>= 1); remaining rows are zero-filled
__device__ __forceinline__ void load_rows(bf16* dst, const bf16* src, long ld, int rows_valid, int rows_pad) {
  for (int e = threadIdx.x; e < rows_pad * 8; e += NTHREADS) {
    const int r = e >> 3, c = e & 7;
    const int rs = r < rows_valid ? r : rows_valid - 1;  // keep the address legal for the zero-fill form
    cp_async16(dst + r * PITCH + c * 8, src + (size_t)rs * ld + c * 8, r < rows_valid ? 16 : 0);
  }
}

// A fragments (16 rows x 64 k) of a row-major tile
__device__ __forceinline__ void load_a_frags(const bf16* tile, int row0, int lane, uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; kk++)
    ldsm_x4(tile + (row0 + (lane & 15)) * PITCH + kk * 16 + (lane >> 4) * 8, a[kk][0], a[kk][1], a[kk][2], a[kk][3]);
}

// acc[NT][4] (16 x NT*8) = A(16 x 64) * Y^T, Y = row-major [NT*8][64] tile ("B col-major": B[k][n] = Y[n][k])
template <int NT>
__device__ __forceinline__ void mma_a_yt(float (&acc)[NT][4], const uint32_t (&a)[4][4], const bf16* Y, int lane) {
#pragma unroll
  for (int j = 0; j < NT; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const bf16* base = Y + ((lane & 7) + ((lane >> 4) & 1) * 8) * PITCH + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int j = 0; j < NT; j += 2) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(base + j * 8 * PITCH + kk * 16, b0, b1, b2, b3);
      mma16816(acc[j], a[kk], b0, b1);
      mma16816(acc[j + 1], a[kk], b2, b3);
    }
  }
}

// o[8][4] (16 x 64) = C(16 x NT*8, fp32 fragments rounded to bf16) * Y, Y = row-major [NT*8][64] tile; k-steps
// at or beyond `kmax` rows are skipped (their coefficients are exactly zero)
template <int NT>
__device__ __forceinline__ void mma_c_y(float (&o)[8][4], const float (&c)[NT][4], const bf16* Y, int lane, int kmax) {
#pragma unroll
  for (int dn = 0; dn < 8; dn++) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
  const bf16* base = Y + ((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + ((lane >> 4) & 1) * 8;
#pragma unroll
  for (int k2 = 0; k2 < NT / 2; k2++) {
    if (k2 * 16 < kmax) {
      uint32_t a[4];
      a[0] = pack2(c[2 * k2][0], c[2 * k2][1]);
      a[1] = pack2(c[2 * k2][2], c[2 * k2][3]);
      a[2] = pack2(c[2 * k2 + 1][0], c[2 * k2 + 1][1]);
      a[3] = pack2(c[2 * k2 + 1][2], c[2 * k2 + 1][3]);
#pragma unroll
      for (int dn = 0; dn < 8; dn += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(base + k2 * 16 * PITCH + dn * 8, b0, b1, b2, b3);
        mma16816(o[dn], a, b0, b1);
        mma16816(o[dn + 1], a, b2, b3);
      }
    }
  }
}

// store a 16 x 64 fp32 fragment tile (scaled) as bf16 rows; row pointers may be null (row out of range)
__device__ __forceinline__ void store_rows(bf16* ra, bf16* rb, const float (&o)[8][4], float mul, int t) {
#pragma unroll
  for (int dn = 0; dn < 8; dn++) {
    if (ra) *reinterpret_cast<uint32_t*>(ra + dn * 8 + 2 * t) = pack2(o[dn][0] * mul, o[dn][1] * mul);
    if (rb) *reinterpret_cast<uint32_t*>(rb + dn * 8 + 2 * t) = pack2(o[dn][2] * mul, o[dn][3] * mul);
  }
}

// ===================================================================================================
// forward
// ===================================================================================================
template <int NT>
__global__ void __launch_bounds__(NTHREADS) attn_mma_fwd_kernel(AttnParams P, int hc, int csize) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_trigger();
  pdl_wait();
  constexpr int LKP = NT * 8;
  bf16* Ks = reinterpret_cast<bf16*>(smraw);
  bf16* Vs = Ks + LKP * PITCH;
  bf16* Qs = Vs + LKP * PITCH;
  float* pb = reinterpret_cast<float*>(Qs + TILE * PITCH);  // [TILE][LKP + 8], only when P.pbar
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, q0 = blockIdx.x * TILE, h_begin = blockIdx.y * hc;
  const int Lq = P.Lq, Lk = P.Lk, H = P.H;
  const int klen = P.key_lens ? min(Lk, P.key_lens[b]) : Lk;
  const int rows_q = min(TILE, Lq - q0);
  const bool wact = w * 16 < rows_q;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)H;
  const int ia = q0 + w * 16 + g, ib = ia + 8;
  const bool va = ia < Lq, vb = ib < Lq;
  // head-mean of P over this CTA's heads: a thread owns the same (row, column) cells of the score tile for every
  // head, so it accumulates them in its own shared-memory cells (8-byte accesses, pitch LKP + 8 floats: no bank
  // conflicts, no synchronisation, and the first head stores instead of adding so nothing is zero-filled)
  constexpr int PBP = LKP + 8;
  float* pba = pb + (w * 16 + g) * PBP + 2 * t;
  float* pbb = pba + 8 * PBP;

  for (int hd = h_begin; hd < h_begin + hc; hd++) {
    __syncthreads();
    load_rows(Ks, (const bf16*)P.k + (size_t)b * Lk * P.k_ld + hd * D, P.k_ld, Lk, LKP);
    load_rows(Vs, (const bf16*)P.v + (size_t)b * Lk * P.v_ld + hd * D, P.v_ld, Lk, LKP);
    load_rows(Qs, (const bf16*)P.q + ((size_t)b * Lq + q0) * P.q_ld + hd * D, P.q_ld, rows_q, TILE);
    cp_async_wait_all();
    __syncthreads();
    if (!wact) continue;  // warp-uniform; every warp reaches the barriers at the top of the next iteration

    float s[NT][4];
    {
      uint32_t qa[4][4];
      load_a_frags(Qs, w * 16, lane, qa);
      mma_a_yt<NT>(s, qa, Ks, lane);
    }
    const size_t da = ((size_t)b * Lq + (va ? ia : 0)) * Lk, db = ((size_t)b * Lq + (vb ? ib : 0)) * Lk;
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = j * 8 + 2 * t + e;
        float x0 = s[j][e] * P.scale, x1 = s[j][2 + e] * P.scale;
        if (c < klen && c != P.key_skip) {
          if (P.dists) {
            x0 += sw * P.dists[da + c] + sb;
            x1 += sw * P.dists[db + c] + sb;
          }
        } else {
          x0 = x1 = -INFINITY;
        }
        s[j][e] = x0;
        s[j][2 + e] = x1;
        mxa = fmaxf(mxa, x0);
        mxb = fmaxf(mxb, x1);
      }
    }
    mxa = quad_max(mxa);
    mxb = quad_max(mxb);
    float suma = 0.f, sumb = 0.f;
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float e0 = (s[j][e] == -INFINITY) ? 0.f : __expf(s[j][e] - mxa);
        const float e1 = (s[j][2 + e] == -INFINITY) ? 0.f : __expf(s[j][2 + e] - mxb);
        s[j][e] = e0;
        s[j][2 + e] = e1;
        suma += e0;
        sumb += e1;
      }
    }
    suma = quad_sum(suma);
    sumb = quad_sum(sumb);
    const float inva = 1.f / suma, invb = 1.f / sumb;
    if (t == 0) {
      if (va) P.lse[((size_t)b * H + hd) * Lq + ia] = mxa + __logf(suma);
      if (vb) P.lse[((size_t)b * H + hd) * Lq + ib] = mxb + __logf(sumb);
    }
    const size_t dia = (((size_t)b * H + hd) * Lq + ia) * Lk, dib = (((size_t)b * H + hd) * Lq + ib) * Lk;
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = j * 8 + 2 * t + e;
        const float pa = s[j][e] * inva, pc = s[j][2 + e] * invb;  // exactly 0 for masked / padded keys
        s[j][e] = pa;  // normalised probabilities (the head-mean below reads them before dropout scales them)
        s[j][2 + e] = pc;
      }
    }
    if (P.pbar) {
      if (hd == h_begin) {
#pragma unroll
        for (int j = 0; j < NT; j++) {
          *reinterpret_cast<float2*>(pba + j * 8) = make_float2(s[j][0] * invH, s[j][1] * invH);
          *reinterpret_cast<float2*>(pbb + j * 8) = make_float2(s[j][2] * invH, s[j][3] * invH);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NT; j++) {
          float2 a = *reinterpret_cast<float2*>(pba + j * 8), c2 = *reinterpret_cast<float2*>(pbb + j * 8);
          a.x += s[j][0] * invH; a.y += s[j][1] * invH;
          c2.x += s[j][2] * invH; c2.y += s[j][3] * invH;
          *reinterpret_cast<float2*>(pba + j * 8) = a;
          *reinterpret_cast<float2*>(pbb + j * 8) = c2;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = j * 8 + 2 * t + e;
        const float pa = s[j][e], pc = s[j][2 + e];
        s[j][e] = pa * dr.scale(dia + c);
        s[j][2 + e] = pc * dr.scale(dib + c);
      }
    }
    float o[8][4];
    mma_c_y<NT>(o, s, Vs, lane, klen);
    bf16* ob = (bf16*)P.out + hd * D;
    store_rows(va ? ob + ((size_t)b * Lq + ia) * (size_t)(H * D) : nullptr,
               vb ? ob + ((size_t)b * Lq + ib) * (size_t)(H * D) : nullptr, o, 1.f, t);
  }
  if (P.pbar) {
    if (csize == 1) {
      __syncthreads();
      for (int e = threadIdx.x; e < rows_q * Lk; e += NTHREADS) {
        const int r = e / Lk, c = e - r * Lk;
        P.pbar[(size_t)b * P.pbar_bs + (size_t)(q0 + r) * P.pbar_rs + c] = pb[r * PBP + c];
      }
    } else {
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();  // every peer's partial head-mean is complete and visible cluster-wide
      const int rank = (int)cluster.block_rank();
      const float* peer[4];
      for (int r = 0; r < csize; r++) peer[r] = cluster.map_shared_rank(pb, r);
      const int total = rows_q * Lk, per = (total + csize - 1) / csize;
      const int e0 = rank * per, e1 = min(total, e0 + per);
      for (int e = e0 + threadIdx.x; e < e1; e += NTHREADS) {
        const int r = e / Lk, c = e - r * Lk;
        float acc = 0.f;
        for (int k = 0; k < csize; k++) acc += peer[k][r * PBP + c];  // fixed rank order: deterministic
        P.pbar[(size_t)b * P.pbar_bs + (size_t)(q0 + r) * P.pbar_rs + c] = acc;
      }
      cluster.sync();  // keep this CTA's shared memory alive until the peers have read it
    }
  }
}

// ===================================================================================================
// forward, version 2: one CTA owns ALL query rows of a (batch, head) item (ceil(Lq / 16) warps, at most 10), so K and V
// are staged once per item instead of once per 64-row query tile; the CTA walks a list of items -- the heads of its
// cluster slice when the KD map is requested, a grid-strided range of (batch, head) pairs otherwise -- and, when two
// operand sets fit in shared memory, cp.async-prefetches the next item's Q / K / V while the current one is in the
// tensor cores.  Each warp stages its 16 output rows in the shared rows its Q slab came from (dead once the A
// fragments are in registers) and writes them with 16-byte row-contiguous stores.
// ===================================================================================================
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// like load_rows, for any block size
__device__ __forceinline__ void load_rows_n(bf16* dst, const bf16* src, long ld, int rows_valid, int rows_pad) {
  for (int e = threadIdx.x; e < rows_pad * 8; e += blockDim.x) {
    const int r = e >> 3, c = e & 7;
    const int rs = r < rows_valid ? r : rows_valid - 1;
    cp_async16(dst + r * PITCH + c * 8, src + (size_t)rs * ld + c * 8, r < rows_valid ? 16 : 0);
  }
}

template <int NT, bool DB>
__global__ void __launch_bounds__(320) attn_fwd2_kernel(AttnParams P, int hc, int csize) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_trigger();
  pdl_wait();
  constexpr int LKP = NT * 8;
  const int nw = blockDim.x >> 5, QP = nw * 16;
  const int set_elems = (2 * LKP + QP) * PITCH;
  bf16* sets = reinterpret_cast<bf16*>(smraw);
  float* pb = reinterpret_cast<float*>(sets + (DB ? 2 : 1) * set_elems);  // [QP][LKP + 8], only when P.pbar
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int Lq = P.Lq, Lk = P.Lk, H = P.H;
  const bool kd = P.pbar != nullptr;
  // item list of this CTA
  const int b_fixed = blockIdx.z, h_begin = blockIdx.y * hc;
  const int n_items = kd ? hc : ((P.B * H - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
  auto item = [&](int k, int& b, int& hd) {
    if (kd) { b = b_fixed; hd = h_begin + k; }
    else { const int idx = blockIdx.x + k * gridDim.x; b = idx / H; hd = idx - b * H; }
  };
  auto load_item = [&](int set, int k) {
    int b, hd;
    item(k, b, hd);
    bf16* Ks = sets + set * set_elems;
    bf16* Vs = Ks + LKP * PITCH;
    bf16* Qs = Vs + LKP * PITCH;
    load_rows_n(Ks, (const bf16*)P.k + (size_t)b * Lk * P.k_ld + hd * D, P.k_ld, Lk, LKP);
    load_rows_n(Vs, (const bf16*)P.v + (size_t)b * Lk * P.v_ld + hd * D, P.v_ld, Lk, LKP);
    load_rows_n(Qs, (const bf16*)P.q + (size_t)b * Lq * P.q_ld + hd * D, P.q_ld, Lq, QP);
    cp_async_commit();
  };
  const bool wact = w * 16 < Lq;
  const float LOG2E = 1.4426950408889634f;
  const float sc2 = P.scale * LOG2E;  // exp2-domain softmax, as in version 3 below
  const float sw2 = P.dists ? P.sprel_w[0] * LOG2E : 0.f, sb2 = P.dists ? P.sprel_b[0] * LOG2E : 0.f;
  const int skip = P.key_skip;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)H;
  const int ia = w * 16 + g, ib = ia + 8;
  const bool va = ia < Lq, vb = ib < Lq;
  constexpr int PBP = LKP + 8;
  float* pba = pb + (w * 16 + g) * PBP + 2 * t;
  float* pbb = pba + 8 * PBP;

  if (n_items > 0) load_item(0, 0);
  for (int k = 0; k < n_items; k++) {
    const int set = DB ? (k & 1) : 0;
    if (DB && k + 1 < n_items) {
      load_item((k + 1) & 1, k + 1);  // the other set was released by the barrier that closed iteration k - 1
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    int b, hd;
    item(k, b, hd);
    bf16* Ks = sets + set * set_elems;
    bf16* Vs = Ks + LKP * PITCH;
    bf16* Qs = Vs + LKP * PITCH;
    const int klen = P.key_lens ? min(Lk, P.key_lens[b]) : Lk;
    if (wact) {
      float s[NT][4];
      {
        uint32_t qa[4][4];
        load_a_frags(Qs, w * 16, lane, qa);
        mma_a_yt<NT>(s, qa, Ks, lane);
      }
      const float* da = P.dists ? P.dists + ((size_t)b * Lq + (va ? ia : 0)) * Lk : nullptr;
      const float* db = P.dists ? P.dists + ((size_t)b * Lq + (vb ? ib : 0)) * Lk : nullptr;
      float mxa[2] = {-INFINITY, -INFINITY}, mxb[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const int c0 = j * 8 + 2 * t;
        float x0 = s[j][0] * sc2, x1 = s[j][1] * sc2, x2 = s[j][2] * sc2, x3 = s[j][3] * sc2;
        if (j * 8 < klen) {  // (tile-uniform) at least one live key in this n-tile
          if (da) {
            const bool k0 = c0 < klen, k1 = c0 + 1 < klen;
            if (k0) { x0 = fmaf(sw2, da[c0], x0 + sb2); x2 = fmaf(sw2, db[c0], x2 + sb2); }
            if (k1) { x1 = fmaf(sw2, da[c0 + 1], x1 + sb2); x3 = fmaf(sw2, db[c0 + 1], x3 + sb2); }
          }
          if (j * 8 + 8 > klen || (skip >= j * 8 && skip < j * 8 + 8)) {  // straddles the length / holds the hole
            if (c0 >= klen || c0 == skip) x0 = x2 = -INFINITY;
            if (c0 + 1 >= klen || c0 + 1 == skip) x1 = x3 = -INFINITY;
          }
        } else {
          x0 = x1 = x2 = x3 = -INFINITY;
        }
        s[j][0] = x0; s[j][1] = x1; s[j][2] = x2; s[j][3] = x3;
        mxa[j & 1] = fmaxf(mxa[j & 1], fmaxf(x0, x1));
        mxb[j & 1] = fmaxf(mxb[j & 1], fmaxf(x2, x3));
      }
      float ma = quad_max(fmaxf(mxa[0], mxa[1])), mb = quad_max(fmaxf(mxb[0], mxb[1]));
      if (ma == -INFINITY) ma = 0.f;  // (a fully masked row: keeps -inf - -inf out of the exponent)
      if (mb == -INFINITY) mb = 0.f;
      float sma[2] = {0.f, 0.f}, smb[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < NT; j++) {  // ex2(-inf) = 0: masked keys need no test (a row always has a live key)
        const float e0 = ex2f(s[j][0] - ma), e1 = ex2f(s[j][1] - ma), e2 = ex2f(s[j][2] - mb), e3 = ex2f(s[j][3] - mb);
        s[j][0] = e0; s[j][1] = e1; s[j][2] = e2; s[j][3] = e3;
        sma[j & 1] += e0 + e1;
        smb[j & 1] += e2 + e3;
      }
      const float suma = quad_sum(sma[0] + sma[1]), sumb = quad_sum(smb[0] + smb[1]);
      const float inva = 1.f / suma, invb = 1.f / sumb;
      if (t == 0) {  // natural-log log-sum-exp (what the backward kernels subtract)
        if (va) P.lse[((size_t)b * H + hd) * Lq + ia] = (ma + __log2f(suma)) * 0.6931471805599453f;
        if (vb) P.lse[((size_t)b * H + hd) * Lq + ib] = (mb + __log2f(sumb)) * 0.6931471805599453f;
      }
#pragma unroll
      for (int j = 0; j < NT; j++) {
        s[j][0] *= inva; s[j][1] *= inva;
        s[j][2] *= invb; s[j][3] *= invb;
      }
      if (kd) {
        if (k == 0) {  // the first head stores, later heads add: nothing is zero-filled
#pragma unroll
          for (int j = 0; j < NT; j++) {
            *reinterpret_cast<float2*>(pba + j * 8) = make_float2(s[j][0] * invH, s[j][1] * invH);
            *reinterpret_cast<float2*>(pbb + j * 8) = make_float2(s[j][2] * invH, s[j][3] * invH);
          }
        } else {
#pragma unroll
          for (int j = 0; j < NT; j++) {
            float2 a = *reinterpret_cast<float2*>(pba + j * 8), c2 = *reinterpret_cast<float2*>(pbb + j * 8);
            a.x += s[j][0] * invH; a.y += s[j][1] * invH;
            c2.x += s[j][2] * invH; c2.y += s[j][3] * invH;
            *reinterpret_cast<float2*>(pba + j * 8) = a;
            *reinterpret_cast<float2*>(pbb + j * 8) = c2;
          }
        }
      }
      if (dr.p > 0.f) {
        const size_t dia = (((size_t)b * H + hd) * Lq + ia) * Lk, dib = (((size_t)b * H + hd) * Lq + ib) * Lk;
#pragma unroll
        for (int j = 0; j < NT; j++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int c = j * 8 + 2 * t + e;
            s[j][e] *= dr.scale(dia + c);
            s[j][2 + e] *= dr.scale(dib + c);
          }
        }
      }
      float o[8][4];
      mma_c_y<NT>(o, s, Vs, lane, klen);
      // stage this warp's 16 x 64 output slab in its own (now dead) Q rows, then 16-byte row-contiguous stores
      __syncwarp();
      bf16* stg = Qs + w * 16 * PITCH;
      store_rows(stg + g * PITCH, stg + (g + 8) * PITCH, o, 1.f, t);
      __syncwarp();
      bf16* ob = (bf16*)P.out + ((size_t)b * Lq + w * 16) * (size_t)(H * D) + hd * D;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = i * 4 + (lane >> 3), c = lane & 7;
        if (w * 16 + r < Lq)
          *reinterpret_cast<uint4*>(ob + (size_t)r * (H * D) + c * 8) = *reinterpret_cast<const uint4*>(stg + r * PITCH + c * 8);
      }
    }
    __syncthreads();  // every warp is done with this set before the loads of iteration k + 1 / k + 2 overwrite it
    if (!DB && k + 1 < n_items) load_item(0, k + 1);
  }
  if (kd) {
    const int b = b_fixed;
    if (csize == 1) {
      for (int e = threadIdx.x; e < Lq * Lk; e += blockDim.x) {
        const int r = e / Lk, c = e - r * Lk;
        P.pbar[(size_t)b * P.pbar_bs + (size_t)r * P.pbar_rs + c] = pb[r * PBP + c];
      }
    } else {
      namespace cg = cooperative_groups;
      cg::cluster_group cluster = cg::this_cluster();
      cluster.sync();  // every peer's partial head-mean is complete and visible cluster-wide
      const int rank = (int)cluster.block_rank();
      const float* peer[4];
      for (int r = 0; r < csize; r++) peer[r] = cluster.map_shared_rank(pb, r);
      const int total = Lq * Lk, per = (total + csize - 1) / csize;
      const int e0 = rank * per, e1 = min(total, e0 + per);
      for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int r = e / Lk, c = e - r * Lk;
        float acc = 0.f;
        for (int kk = 0; kk < csize; kk++) acc += peer[kk][r * PBP + c];  // fixed rank order: deterministic
        P.pbar[(size_t)b * P.pbar_bs + (size_t)r * P.pbar_rs + c] = acc;
      }
      cluster.sync();  // keep this CTA's shared memory alive until the peers have read it
    }
  }
}

// ===================================================================================================
// forward, version 3 (no KD map requested): ONE (batch, head, query tile) per CTA and nothing else -- no head loop,
// no cluster.  The GPU holds every CTA of a launch at once (B x H x tiles CTAs of <= 5 warps, 35-60 KB of shared
// memory, <= 96 registers at Lk <= 80), so one CTA's load phase hides behind the others' tensor-core / softmax phases:
// the measured bound of versions 1 / 2 was latency (issue slots 34 % busy, 2.5 warps per scheduler:
// profiles/r02_ncu_full_summary_v1.md).  Softmax in the exp2 domain (scores pre-multiplied by log2 e: one FMUL less
// per element), tile-uniform mask fast path (only the n-tiles that straddle the key length or hold the skipped key
// test per element), two independent max / sum chains per row.  Measured (graph-replayed, dropout 0.1): B64 H12 80x80
// 20.1 -> 14.5 us, B128 H12 160x160 115 -> 101 us, B128 H12 50x160 55 -> 45 us, B320 H12 36x36 31 -> 26 us.
//
// With the KD map the head mean needs a cross-head reduction.  Tried here and dropped: summing the heads in global
// memory in 2^-30 fixed point with 64-bit integer `red.add` (order-independent, so bit-reproducible) -- the L2 atomic
// units retire ~0.16 G cells/us, +10 us at B64 H12 80x80 and +40 us at B128 H2 160x160, worse than the cluster /
// distributed-shared-memory reduction of versions 1 / 2 (+8 us), which therefore keep that case.
// ===================================================================================================

template <int NT, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) attn_fwd3_kernel(AttnParams P) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_trigger();
  constexpr int LKP = NT * 8, QT = NW * 16;
  bf16* Ks = reinterpret_cast<bf16*>(smraw);
  bf16* Vs = Ks + LKP * PITCH;
  bf16* Qs = Vs + LKP * PITCH;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * QT;
  const int Lq = P.Lq, Lk = P.Lk, H = P.H;
  const int rows_q = min(QT, Lq - q0);
  pdl_wait();
  const int klen = P.key_lens ? min(Lk, P.key_lens[b]) : Lk;
  load_rows_n(Ks, (const bf16*)P.k + (size_t)b * Lk * P.k_ld + hd * D, P.k_ld, Lk, LKP);
  load_rows_n(Vs, (const bf16*)P.v + (size_t)b * Lk * P.v_ld + hd * D, P.v_ld, Lk, LKP);
  load_rows_n(Qs, (const bf16*)P.q + ((size_t)b * Lq + q0) * P.q_ld + hd * D, P.q_ld, rows_q, QT);
  cp_async_commit();
  const float LOG2E = 1.4426950408889634f;
  const float sc2 = P.scale * LOG2E;
  const float sw2 = P.dists ? P.sprel_w[0] * LOG2E : 0.f, sb2 = P.dists ? P.sprel_b[0] * LOG2E : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const int ia = q0 + w * 16 + g, ib = ia + 8;
  const bool va = ia < Lq, vb = ib < Lq;
  const int skip = P.key_skip;
  cp_async_wait<0>();
  __syncthreads();
  if (w * 16 < rows_q) {
    float s[NT][4];
    {
      uint32_t qa[4][4];
      load_a_frags(Qs, w * 16, lane, qa);
      mma_a_yt<NT>(s, qa, Ks, lane);
    }
    const float* da = P.dists ? P.dists + ((size_t)b * Lq + (va ? ia : 0)) * Lk : nullptr;
    const float* db = P.dists ? P.dists + ((size_t)b * Lq + (vb ? ib : 0)) * Lk : nullptr;
    float mxa[2] = {-INFINITY, -INFINITY}, mxb[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int c0 = j * 8 + 2 * t;
      float x0 = s[j][0] * sc2, x1 = s[j][1] * sc2, x2 = s[j][2] * sc2, x3 = s[j][3] * sc2;
      if (j * 8 < klen) {  // (tile-uniform) at least one live key in this n-tile
        if (da) {
          const bool k0 = c0 < klen, k1 = c0 + 1 < klen;
          if (k0) { x0 = fmaf(sw2, da[c0], x0 + sb2); x2 = fmaf(sw2, db[c0], x2 + sb2); }
          if (k1) { x1 = fmaf(sw2, da[c0 + 1], x1 + sb2); x3 = fmaf(sw2, db[c0 + 1], x3 + sb2); }
        }
        if (j * 8 + 8 > klen || (skip >= j * 8 && skip < j * 8 + 8)) {  // straddles the length / holds the hole
          if (c0 >= klen || c0 == skip) x0 = x2 = -INFINITY;
          if (c0 + 1 >= klen || c0 + 1 == skip) x1 = x3 = -INFINITY;
        }
      } else {
        x0 = x1 = x2 = x3 = -INFINITY;
      }
      s[j][0] = x0; s[j][1] = x1; s[j][2] = x2; s[j][3] = x3;
      mxa[j & 1] = fmaxf(mxa[j & 1], fmaxf(x0, x1));
      mxb[j & 1] = fmaxf(mxb[j & 1], fmaxf(x2, x3));
    }
    float ma = quad_max(fmaxf(mxa[0], mxa[1])), mb = quad_max(fmaxf(mxb[0], mxb[1]));
    if (ma == -INFINITY) ma = 0.f;  // (a fully masked row: keeps -inf - -inf out of the exponent)
    if (mb == -INFINITY) mb = 0.f;
    float sma[2] = {0.f, 0.f}, smb[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; j++) {  // ex2(-inf) = 0: masked keys need no test (a row always has a live key)
      const float e0 = ex2f(s[j][0] - ma), e1 = ex2f(s[j][1] - ma), e2 = ex2f(s[j][2] - mb), e3 = ex2f(s[j][3] - mb);
      s[j][0] = e0; s[j][1] = e1; s[j][2] = e2; s[j][3] = e3;
      sma[j & 1] += e0 + e1;
      smb[j & 1] += e2 + e3;
    }
    const float suma = quad_sum(sma[0] + sma[1]), sumb = quad_sum(smb[0] + smb[1]);
    const float inva = 1.f / suma, invb = 1.f / sumb;
    if (t == 0) {  // natural-log log-sum-exp (what the backward kernels subtract)
      if (va) P.lse[((size_t)b * H + hd) * Lq + ia] = (ma + __log2f(suma)) * 0.6931471805599453f;
      if (vb) P.lse[((size_t)b * H + hd) * Lq + ib] = (mb + __log2f(sumb)) * 0.6931471805599453f;
    }
#pragma unroll
    for (int j = 0; j < NT; j++) {
      s[j][0] *= inva; s[j][1] *= inva;
      s[j][2] *= invb; s[j][3] *= invb;
    }
    if (dr.p > 0.f) {
      const size_t dia = (((size_t)b * H + hd) * Lq + ia) * Lk, dib = (((size_t)b * H + hd) * Lq + ib) * Lk;
#pragma unroll
      for (int j = 0; j < NT; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int c = j * 8 + 2 * t + e;
          s[j][e] *= dr.scale(dia + c);
          s[j][2 + e] *= dr.scale(dib + c);
        }
      }
    }
    float o[8][4];
    mma_c_y<NT>(o, s, Vs, lane, klen);
    __syncwarp();
    bf16* stg = Qs + w * 16 * PITCH;  // this warp's Q rows are dead: stage the output slab there
    store_rows(stg + g * PITCH, stg + (g + 8) * PITCH, o, 1.f, t);
    __syncwarp();
    bf16* ob = (bf16*)P.out + ((size_t)b * Lq + q0 + w * 16) * (size_t)(H * D) + hd * D;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = i * 4 + (lane >> 3), c = lane & 7;
      if (q0 + w * 16 + r < Lq)
        *reinterpret_cast<uint4*>(ob + (size_t)r * (H * D) + c * 8) = *reinterpret_cast<const uint4*>(stg + r * PITCH + c * 8);
    }
  }
}

// ===================================================================================================
// backward pass 1 (query-major): delta, dQ, d(sprel)
// ===================================================================================================
template <int NT>
__global__ void __launch_bounds__(NTHREADS) attn_mma_bwd_q_kernel(AttnParams P) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_trigger();
  pdl_wait();
  constexpr int LKP = NT * 8;
  bf16* Ks = reinterpret_cast<bf16*>(smraw);
  bf16* Vs = Ks + LKP * PITCH;
  bf16* Qs = Vs + LKP * PITCH;
  bf16* Gs = Qs + TILE * PITCH;
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * TILE;
  const int Lq = P.Lq, Lk = P.Lk, H = P.H;
  const int klen = P.key_lens ? min(Lk, P.key_lens[b]) : Lk;
  const int rows_q = min(TILE, Lq - q0);
  const bool wact = w * 16 < rows_q;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)H;
  const int ia = q0 + w * 16 + g, ib = ia + 8;
  const bool va = ia < Lq, vb = ib < Lq;

  load_rows(Ks, (const bf16*)P.k + (size_t)b * Lk * P.k_ld + hd * D, P.k_ld, Lk, LKP);
  load_rows(Vs, (const bf16*)P.v + (size_t)b * Lk * P.v_ld + hd * D, P.v_ld, Lk, LKP);
  load_rows(Qs, (const bf16*)P.q + ((size_t)b * Lq + q0) * P.q_ld + hd * D, P.q_ld, rows_q, TILE);
  load_rows(Gs, (const bf16*)P.dout + ((size_t)b * Lq + q0) * (size_t)(H * D) + hd * D, (long)H * D, rows_q, TILE);
  cp_async_wait_all();
  __syncthreads();

  float acc_dw = 0.f, acc_db = 0.f;
  if (wact) {
    float p[NT][4], dp[NT][4];
    {
      uint32_t qa[4][4];
      load_a_frags(Qs, w * 16, lane, qa);
      mma_a_yt<NT>(p, qa, Ks, lane);
    }
    {
      uint32_t ga[4][4];
      load_a_frags(Gs, w * 16, lane, ga);
      mma_a_yt<NT>(dp, ga, Vs, lane);
    }
    const float lsa = va ? P.lse[((size_t)b * H + hd) * Lq + ia] : 0.f;
    const float lsb = vb ? P.lse[((size_t)b * H + hd) * Lq + ib] : 0.f;
    const size_t da = ((size_t)b * Lq + (va ? ia : 0)) * Lk, db = ((size_t)b * Lq + (vb ? ib : 0)) * Lk;
    const size_t dia = (((size_t)b * H + hd) * Lq + ia) * Lk, dib = (((size_t)b * H + hd) * Lq + ib) * Lk;
    const float* dpa = P.dpbar ? P.dpbar + (size_t)b * P.pbar_bs + (size_t)(va ? ia : 0) * P.pbar_rs : nullptr;
    const float* dpb = P.dpbar ? P.dpbar + (size_t)b * P.pbar_bs + (size_t)(vb ? ib : 0) * P.pbar_rs : nullptr;
    float dla = 0.f, dlb = 0.f;
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = j * 8 + 2 * t + e;
        float p0 = 0.f, p1 = 0.f, t0 = 0.f, t1 = 0.f;
        if (c < klen && c != P.key_skip) {
          float x0 = p[j][e] * P.scale, x1 = p[j][2 + e] * P.scale;
          if (P.dists) {
            x0 += sw * P.dists[da + c] + sb;
            x1 += sw * P.dists[db + c] + sb;
          }
          p0 = va ? __expf(x0 - lsa) : 0.f;
          p1 = vb ? __expf(x1 - lsb) : 0.f;
          t0 = dp[j][e] * dr.scale(dia + c);
          t1 = dp[j][2 + e] * dr.scale(dib + c);
          if (P.dpbar) {
            t0 += dpa[c] * invH;
            t1 += dpb[c] * invH;
          }
        }
        p[j][e] = p0;
        p[j][2 + e] = p1;
        dp[j][e] = t0;
        dp[j][2 + e] = t1;
        dla = fmaf(p0, t0, dla);
        dlb = fmaf(p1, t1, dlb);
      }
    }
    dla = quad_sum(dla);
    dlb = quad_sum(dlb);
    if (t == 0) {
      if (va) P.delta[((size_t)b * H + hd) * Lq + ia] = dla;
      if (vb) P.delta[((size_t)b * H + hd) * Lq + ib] = dlb;
    }
#pragma unroll
    for (int j = 0; j < NT; j++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = j * 8 + 2 * t + e;
        const float ds0 = p[j][e] * (dp[j][e] - dla), ds1 = p[j][2 + e] * (dp[j][2 + e] - dlb);
        p[j][e] = ds0;
        p[j][2 + e] = ds1;
        if (P.dists && c < klen && c != P.key_skip) {
          acc_dw = fmaf(ds0, P.dists[da + c], acc_dw);
          acc_dw = fmaf(ds1, P.dists[db + c], acc_dw);
          acc_db += ds0 + ds1;
        }
      }
    }
    float o[8][4];
    mma_c_y<NT>(o, p, Ks, lane, klen);
    bf16* qb = (bf16*)P.dq + hd * D;
    store_rows(va ? qb + ((size_t)b * Lq + ia) * P.dq_ld : nullptr, vb ? qb + ((size_t)b * Lq + ib) * P.dq_ld : nullptr,
               o, P.scale, t);
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

// ===================================================================================================
// backward pass 2 (key-major): dK, dV from the transposed tile
// ===================================================================================================
template <int NTQ>
__global__ void __launch_bounds__(NTHREADS) attn_mma_bwd_kv_kernel(AttnParams P) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_trigger();
  pdl_wait();
  constexpr int LQP = NTQ * 8;
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Gs = Qs + LQP * PITCH;
  bf16* Ks = Gs + LQP * PITCH;
  bf16* Vs = Ks + TILE * PITCH;
  float* ls = reinterpret_cast<float*>(Vs + TILE * PITCH);  // [LQP] lse
  float* dl = ls + LQP;                                     // [LQP] delta
  bf16* Os = reinterpret_cast<bf16*>(dl + LQP);             // [LQP][PITCH] forward output, only when delta_from_out
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z, hd = blockIdx.y, k0 = blockIdx.x * TILE;
  const int Lq = P.Lq, Lk = P.Lk, H = P.H;
  const int klen = P.key_lens ? min(Lk, P.key_lens[b]) : Lk;
  const int rows_k = min(TILE, Lk - k0);
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)H;

  load_rows(Qs, (const bf16*)P.q + (size_t)b * Lq * P.q_ld + hd * D, P.q_ld, Lq, LQP);
  load_rows(Gs, (const bf16*)P.dout + (size_t)b * Lq * (size_t)(H * D) + hd * D, (long)H * D, Lq, LQP);
  load_rows(Ks, (const bf16*)P.k + ((size_t)b * Lk + k0) * P.k_ld + hd * D, P.k_ld, rows_k, TILE);
  load_rows(Vs, (const bf16*)P.v + ((size_t)b * Lk + k0) * P.v_ld + hd * D, P.v_ld, rows_k, TILE);
  if (P.delta_from_out)
    load_rows(Os, (const bf16*)P.out + (size_t)b * Lq * (size_t)(H * D) + hd * D, (long)H * D, Lq, LQP);
  for (int i = threadIdx.x; i < LQP; i += NTHREADS) {
    ls[i] = i < Lq ? P.lse[((size_t)b * H + hd) * Lq + i] : 0.f;
    if (!P.delta_from_out) dl[i] = i < Lq ? P.delta[((size_t)b * H + hd) * Lq + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  if (P.delta_from_out) {
    // delta_i = sum_j P_ij dP_ij = dO_i . O_i (dropout included: O was formed with the same mask), so this kernel
    // does not wait for the query-major one and the two run as concurrent graph branches
    for (int i = threadIdx.x; i < LQP; i += NTHREADS) {
      float acc = 0.f;
      if (i < Lq) {
        const uint4* gp = reinterpret_cast<const uint4*>(Gs + i * PITCH);
        const uint4* op = reinterpret_cast<const uint4*>(Os + i * PITCH);
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const uint4 gv = gp[c], ov = op[c];
          const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gv);
          const __nv_bfloat162* oh = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const float2 a = __bfloat1622float2(gh[e]), o2 = __bfloat1622float2(oh[e]);
            acc = fmaf(a.x, o2.x, acc);
            acc = fmaf(a.y, o2.y, acc);
          }
        }
      }
      dl[i] = acc;
    }
    __syncthreads();
  }
  if (w * 16 >= rows_k) return;  // no barriers below

  const int ja = k0 + w * 16 + g, jb = ja + 8;
  const bool va = ja < Lk, vb = jb < Lk;
  bf16* dka = va ? (bf16*)P.dk + ((size_t)b * Lk + ja) * P.dk_ld + hd * D : nullptr;
  bf16* dkb = vb ? (bf16*)P.dk + ((size_t)b * Lk + jb) * P.dk_ld + hd * D : nullptr;
  bf16* dva = va ? (bf16*)P.dv + ((size_t)b * Lk + ja) * P.dv_ld + hd * D : nullptr;
  bf16* dvb = vb ? (bf16*)P.dv + ((size_t)b * Lk + jb) * P.dv_ld + hd * D : nullptr;
  float o[8][4];
  if (k0 + w * 16 >= klen) {  // every key of this slab is masked: exact zero gradients
#pragma unroll
    for (int dn = 0; dn < 8; dn++) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    store_rows(dka, dkb, o, 0.f, t);
    store_rows(dva, dvb, o, 0.f, t);
    return;
  }
  const bool ma = ja < klen && ja != P.key_skip, mb = jb < klen && jb != P.key_skip;
  const size_t dbase = ((size_t)b * H + hd) * Lq;

  float p[NTQ][4];
  {
    uint32_t ka[4][4];
    load_a_frags(Ks, w * 16, lane, ka);
    mma_a_yt<NTQ>(p, ka, Qs, lane);
  }
#pragma unroll
  for (int n = 0; n < NTQ; n++) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int c = n * 8 + 2 * t + e;  // query index
      float p0 = 0.f, p1 = 0.f;
      if (c < Lq) {
        float x0 = p[n][e] * P.scale, x1 = p[n][2 + e] * P.scale;
        if (P.dists) {
          const size_t di = ((size_t)b * Lq + c) * Lk;
          if (ma) x0 += sw * P.dists[di + ja] + sb;
          if (mb) x1 += sw * P.dists[di + jb] + sb;
        }
        p0 = ma ? __expf(x0 - ls[c]) : 0.f;
        p1 = mb ? __expf(x1 - ls[c]) : 0.f;
      }
      p[n][e] = p0;
      p[n][2 + e] = p1;
    }
  }
  // dV = (P^T o dropout) dO
  {
    float pd[NTQ][4];
#pragma unroll
    for (int n = 0; n < NTQ; n++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = n * 8 + 2 * t + e;
        const size_t di = (dbase + c) * Lk;
        pd[n][e] = p[n][e] * dr.scale(di + ja);
        pd[n][2 + e] = p[n][2 + e] * dr.scale(di + jb);
      }
    }
    mma_c_y<NTQ>(o, pd, Gs, lane, Lq);
    store_rows(dva, dvb, o, 1.f, t);
  }
  // dS^T = P^T o (dP^T o dropout + dPbar^T / H - delta)
  {
    float dp[NTQ][4];
    {
      uint32_t vf[4][4];
      load_a_frags(Vs, w * 16, lane, vf);
      mma_a_yt<NTQ>(dp, vf, Gs, lane);
    }
#pragma unroll
    for (int n = 0; n < NTQ; n++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = n * 8 + 2 * t + e;
        const size_t di = (dbase + c) * Lk;
        float t0 = dp[n][e] * dr.scale(di + ja), t1 = dp[n][2 + e] * dr.scale(di + jb);
        if (P.dpbar && c < Lq) {
          const float* dr_ = P.dpbar + (size_t)b * P.pbar_bs + (size_t)c * P.pbar_rs;
          if (ma) t0 += dr_[ja] * invH;
          if (mb) t1 += dr_[jb] * invH;
        }
        p[n][e] *= (t0 - dl[c]);
        p[n][2 + e] *= (t1 - dl[c]);
      }
    }
  }
  mma_c_y<NTQ>(o, p, Qs, lane, Lq);
  store_rows(dka, dkb, o, P.scale, t);
}

// ---- host side ----------------------------------------------------------------------------------------
bool mma_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_ATTN_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int pick_nt(int L) { return L <= 32 ? 4 : L <= 48 ? 6 : L <= 80 ? 10 : L <= 160 ? 20 : 0; }

bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }
bool ld_ok(long ld) { return ld >= D && (ld % 8) == 0; }

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 48 * 1024)
    MAGIC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), name);
  return MAGIC_OK;
}

bool fwd_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_ATTN_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int NT>
int launch_fwd_v1(const AttnParams& P, cudaStream_t st) {
  // KD attention map: the heads of a query tile are shared out over a cluster of 4 / 2 CTAs when each still gets
  // >= 2 heads (below that the cluster barrier costs more than the serial head walk it replaces: measured)
  int csize = 1;
  if (P.pbar && P.H >= 4) csize = (P.H % 4 == 0 && P.H >= 8) ? 4 : (P.H % 2 == 0) ? 2 : 1;
  const size_t smem = (size_t)(2 * NT * 8 + TILE) * PITCH * 2 + (P.pbar ? (size_t)TILE * (NT * 8 + 8) * 4 : 0);
  int rc = set_smem(attn_mma_fwd_kernel<NT>, smem, "magic_attn_fwd");
  if (rc) return rc;
  const int hc = P.pbar ? P.H / csize : 1;
  dim3 grid((P.Lq + TILE - 1) / TILE, P.H / hc, P.B);
  if (csize > 1) {
    MAGIC_CUDA(magic_launch_cluster(attn_mma_fwd_kernel<NT>, grid, dim3(NTHREADS), smem, st, dim3(1, csize, 1), P, hc,
                                    csize),
               "magic_attn_fwd(mma, cluster)");
  } else {
    MAGIC_CUDA(magic_launch(attn_mma_fwd_kernel<NT>, grid, dim3(NTHREADS), smem, st, P, hc, 1), "magic_attn_fwd(mma)");
  }
  return MAGIC_OK;
}

template <int NT, bool DB>
int launch_fwd2_k(const AttnParams& P, int nw, size_t smem, int hc, int csize, cudaStream_t st) {
  int rc = set_smem(attn_fwd2_kernel<NT, DB>, smem, "magic_attn_fwd");
  if (rc) return rc;
  if (P.pbar) {
    dim3 grid(1, P.H / hc, P.B);
    if (csize > 1) {
      MAGIC_CUDA(magic_launch_cluster(attn_fwd2_kernel<NT, DB>, grid, dim3(nw * 32), smem, st, dim3(1, csize, 1), P, hc,
                                      csize),
                 "magic_attn_fwd(v2, cluster)");
    } else {
      MAGIC_CUDA(magic_launch(attn_fwd2_kernel<NT, DB>, grid, dim3(nw * 32), smem, st, P, hc, 1), "magic_attn_fwd(v2)");
    }
    return MAGIC_OK;
  }
  // persistent walk over the (batch, head) pairs: about as many CTAs as fit on the GPU at once
  int per_sm = (int)((200 * 1024) / smem);
  const int by_threads = 1536 / (nw * 32);
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  const long items = (long)P.B * P.H;
  long ctas = (long)magic_num_sms() * per_sm;
  if (ctas > items) ctas = items;
  MAGIC_CUDA(magic_launch(attn_fwd2_kernel<NT, DB>, dim3((unsigned)ctas), dim3(nw * 32), smem, st, P, 1, 1),
             "magic_attn_fwd(v2)");
  return MAGIC_OK;
}

// MAGIC_ATTN_FWD: 0 (default) = by the measured heuristic, 1 / 2 / 3 = force that forward version where it applies
int fwd_force() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_ATTN_FWD");
    v = e ? atoi(e) : 0;
    if (fwd_v1()) v = 1;
  }
  return v;
}

template <int NT, int NW>
int launch_fwd3_k(const AttnParams& P, int tiles, cudaStream_t st) {
  constexpr int MINB = NT >= 20 ? 2 : 4;
  const size_t smem = (size_t)(2 * NT * 8 + NW * 16) * PITCH * 2;
  int rc = set_smem(attn_fwd3_kernel<NT, NW, MINB>, smem, "magic_attn_fwd");
  if (rc) return rc;
  MAGIC_CUDA(magic_launch(attn_fwd3_kernel<NT, NW, MINB>, dim3(tiles, P.H, P.B), dim3(NW * 32), smem, st, P),
             "magic_attn_fwd(v3)");
  return MAGIC_OK;
}

template <int NT>
int launch_fwd3(const AttnParams& P, cudaStream_t st) {
  // query tiles of at most 80 rows (5 warps), balanced: Lq = 100 -> 2 tiles of 64 rows, Lq = 160 -> 2 x 80
  const int tiles = (P.Lq + 79) / 80;
  const int nw = ((P.Lq + tiles - 1) / tiles + 15) / 16;
  switch (nw) {
    case 1:
    case 2: return launch_fwd3_k<NT, 2>(P, (P.Lq + 31) / 32, st);
    case 3: return launch_fwd3_k<NT, 3>(P, (P.Lq + 47) / 48, st);
    case 4: return launch_fwd3_k<NT, 4>(P, (P.Lq + 63) / 64, st);
    default: return launch_fwd3_k<NT, 5>(P, (P.Lq + 79) / 80, st);
  }
}

template <int NT>
int launch_fwd2(const AttnParams& P, cudaStream_t st) {
  const int nw = (P.Lq + 15) / 16;
  int csize = 1;
  if (P.pbar && P.H >= 4) csize = (P.H % 4 == 0 && P.H >= 8) ? 4 : (P.H % 2 == 0) ? 2 : 1;
  const int hc = P.pbar ? P.H / csize : 1;
  const size_t set_bytes = (size_t)(2 * NT * 8 + nw * 16) * PITCH * 2;
  const size_t pb_bytes = P.pbar ? (size_t)nw * 16 * (NT * 8 + 8) * 4 : 0;
  const long my_items = P.pbar ? hc : 2;  // plain mode: a CTA usually walks several pairs
  // two operand sets (prefetch of the next item) when they fit beside the KD-map accumulator and pay off
  const bool db = my_items > 1 && 2 * set_bytes + pb_bytes <= 200 * 1024;
  if (db) return launch_fwd2_k<NT, true>(P, nw, 2 * set_bytes + pb_bytes, hc, csize, st);
  if (set_bytes + pb_bytes > 220 * 1024) return launch_fwd_v1<NT>(P, st);
  return launch_fwd2_k<NT, false>(P, nw, set_bytes + pb_bytes, hc, csize, st);
}

template <int NT>
int launch_fwd(const AttnParams& P, cudaStream_t st) {
  const int force = fwd_force();
  const bool v3_ok = P.Lq <= 160 && !P.pbar;
  if (force == 3 && v3_ok) return launch_fwd3<NT>(P, st);
  if (force == 1 || P.Lq > 160) return launch_fwd_v1<NT>(P, st);
  if (force == 2) return P.Lq <= 64 ? launch_fwd_v1<NT>(P, st) : launch_fwd2<NT>(P, st);
  // default: version 3 (one CTA per (batch, head, query tile)) whenever no KD map is requested
  if (v3_ok) return launch_fwd3<NT>(P, st);
  // measured (B200, graph-replayed, scripts/graph_micro.py attn_l): the one-CTA-per-item kernel wins where a 64-row
  // tiling would stage K / V twice or more (Lq 80: 29.7 -> 24.8 us, Lq 160: 206 -> 166 us at H = 12 with the KD map;
  // 43.8 -> 29.8 us at H = 2) and loses on short query ranges (36 x 36: 31 -> 38 us, 50 x 160: 55 -> 90 us), where its
  // 2 - 4 warp CTAs leave the SM under-occupied
  if (P.Lq <= 64) return launch_fwd_v1<NT>(P, st);
  return launch_fwd2<NT>(P, st);
}

template <int NT>
int launch_bwd_q(const AttnParams& P, cudaStream_t st) {
  const size_t smem = (size_t)(2 * NT * 8 + 2 * TILE) * PITCH * 2;
  int rc = set_smem(attn_mma_bwd_q_kernel<NT>, smem, "magic_attn_bwd");
  if (rc) return rc;
  dim3 grid((P.Lq + TILE - 1) / TILE, P.H, P.B);
  MAGIC_CUDA(magic_launch(attn_mma_bwd_q_kernel<NT>, grid, dim3(NTHREADS), smem, st, P), "magic_attn_bwd(mma q)");
  return MAGIC_OK;
}

template <int NTQ>
int launch_bwd_kv(const AttnParams& P, cudaStream_t st) {
  const size_t smem = (size_t)(2 * NTQ * 8 + 2 * TILE) * PITCH * 2 + (size_t)2 * NTQ * 8 * 4 +
                      (P.delta_from_out ? (size_t)NTQ * 8 * PITCH * 2 : 0);
  int rc = set_smem(attn_mma_bwd_kv_kernel<NTQ>, smem, "magic_attn_bwd");
  if (rc) return rc;
  dim3 grid((P.Lk + TILE - 1) / TILE, P.H, P.B);
  MAGIC_CUDA(magic_launch(attn_mma_bwd_kv_kernel<NTQ>, grid, dim3(NTHREADS), smem, st, P), "magic_attn_bwd(mma kv)");
  return MAGIC_OK;
}

bool operands_ok(const AttnParams& P) {
  return al16(P.q) && al16(P.k) && al16(P.v) && ld_ok(P.q_ld) && ld_ok(P.k_ld) && ld_ok(P.v_ld);
}

}  // namespace

#define DISPATCH_NT(nt, fn, ...)                     \
  switch (nt) {                                      \
    case 4: return fn<4>(__VA_ARGS__);               \
    case 6: return fn<6>(__VA_ARGS__);               \
    case 10: return fn<10>(__VA_ARGS__);             \
    default: return fn<20>(__VA_ARGS__);             \
  }

int attn_mma_fwd(const AttnParams& P, cudaStream_t st) {
  const int nt = pick_nt(P.Lk);
  if (mma_disabled() || nt == 0 || !operands_ok(P) || ((uintptr_t)P.out & 3)) return MAGIC_ERR_UNSUPPORTED;
  DISPATCH_NT(nt, launch_fwd, P, st);
}

int attn_mma_bwd_part(const AttnParams& P0, int part, cudaStream_t st) {
  AttnParams P = P0;
  const int nt = pick_nt(P.Lk), ntq = pick_nt(P.Lq);
  if (mma_disabled() || nt == 0 || ntq == 0 || !operands_ok(P) || !al16(P.dout) || P.dpbar != nullptr)
    return MAGIC_ERR_UNSUPPORTED;
  if (((uintptr_t)P.dq & 3) || ((uintptr_t)P.dk & 3) || ((uintptr_t)P.dv & 3) || (P.dq_ld & 1) || (P.dk_ld & 1) ||
      (P.dv_ld & 1))
    return MAGIC_ERR_UNSUPPORTED;
  if (part == 1) {
    P.delta_from_out = 0;
    switch (nt) {
      case 4: return launch_bwd_q<4>(P, st);
      case 6: return launch_bwd_q<6>(P, st);
      case 10: return launch_bwd_q<10>(P, st);
      default: return launch_bwd_q<20>(P, st);
    }
  }
  if (P.out == nullptr || !al16(P.out)) return MAGIC_ERR_UNSUPPORTED;
  P.delta_from_out = 1;
  DISPATCH_NT(ntq, launch_bwd_kv, P, st);
}

int attn_mma_bwd(const AttnParams& P, cudaStream_t st) {
  const int nt = pick_nt(P.Lk), ntq = pick_nt(P.Lq);
  if (mma_disabled() || nt == 0 || ntq == 0 || !operands_ok(P) || !al16(P.dout)) return MAGIC_ERR_UNSUPPORTED;
  if (((uintptr_t)P.dq & 3) || ((uintptr_t)P.dk & 3) || ((uintptr_t)P.dv & 3) || (P.dq_ld & 1) || (P.dk_ld & 1) ||
      (P.dv_ld & 1))
    return MAGIC_ERR_UNSUPPORTED;
  int rc;
  switch (nt) {
    case 4: rc = launch_bwd_q<4>(P, st); break;
    case 6: rc = launch_bwd_q<6>(P, st); break;
    case 10: rc = launch_bwd_q<10>(P, st); break;
    default: rc = launch_bwd_q<20>(P, st); break;
  }
  if (rc) return rc;
  DISPATCH_NT(ntq, launch_bwd_kv, P, st);
}
