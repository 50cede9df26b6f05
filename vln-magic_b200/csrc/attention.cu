// Fused multi-head attention for the MAGIC encoders (head_dim = 64, Lq/Lk <= 320):
//   S = scale * Q K^T + (w_sprel * dist + b_sprel)      graph-distance bias added inside the softmax tile
//   keys j >= key_len[b] are excluded (== the reference's -10000 / -inf additive masks in fp32)
//   P = softmax(S);  O = dropout(P) V;  Pbar = mean_heads(P)  (KD attention map, optional, fp32)
// All K/V of one (batch, head) live in shared memory; one warp owns one query (or key) row; lanes
// sweep keys for the score/softmax phase and head-dim columns for the P*V phase.  One CTA walks all
// heads of its rows so the head-mean map is accumulated in registers and written once (no atomics).
// Backward = two passes with the same structure (dQ by query rows, dK/dV by key rows); P is
// recomputed from the saved log-sum-exp, the softmax row term delta is produced by pass 1.
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int D = 64;
constexpr int DP = 65;       // padded row (bank-conflict-free when lanes sweep rows)
constexpr int NW = 8;        // warps per CTA
constexpr int RPW = 2;       // rows per warp
constexpr int ROWS = NW * RPW;
constexpr int MAXJ = 10;     // ceil(320 / 32)
constexpr int MAXL = 32 * MAXJ;

struct AttnParams {
  const void *q, *k, *v;
  long q_ld, k_ld, v_ld;        // elements between consecutive tokens
  void* out;                    // [B*Lq, H*64]
  float* lse;                   // [B, H, Lq]
  float* pbar;                  // optional head-mean probs
  long pbar_bs, pbar_rs;        // batch / row strides (elements)
  const int* key_lens;          // [B] or null
  const float* dists;           // [B, Lq, Lk] or null
  const float *sprel_w, *sprel_b;
  int B, H, Lq, Lk;
  float scale;
  float drop_p;
  uint32_t salt;
  const unsigned long long* seed_ptr;
  // backward
  const void* dout;
  const float* dpbar;
  float* delta;                 // [B, H, Lq]
  void *dq, *dk, *dv;
  long dq_ld, dk_ld, dv_ld;
  float* dsprel;                // [2] : dw, db
};

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, const T* src, long ld, int head, int rows, int b_row0) {
  // dst[rows][DP] <- src[(b_row0 + r) * ld + head*64 + d]
  for (int e = threadIdx.x; e < rows * D; e += blockDim.x) {
    const int r = e >> 6, d = e & 63;
    dst[r * DP + d] = ldf(src, (size_t)(b_row0 + r) * ld + head * D + d);
  }
}

template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_fwd_kernel(AttnParams P) {
  extern __shared__ float sm[];
  float* Ks = sm;                          // [Lk][DP]
  float* Vs = Ks + (size_t)P.Lk * DP;      // [Lk][DP]
  float* qs = Vs + (size_t)P.Lk * DP;      // [NW][D]
  float* ps = qs + NW * D;                 // [NW][MAXL]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;
  float pb[RPW][MAXJ];
#pragma unroll
  for (int r = 0; r < RPW; r++)
#pragma unroll
    for (int jj = 0; jj < MAXJ; jj++) pb[r][jj] = 0.f;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Ks, (const T*)P.k, P.k_ld, hd, P.Lk, b * P.Lk);
    load_tile(Vs, (const T*)P.v, P.v_ld, hd, P.Lk, b * P.Lk);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; r++) {
      const int i = row0 + w * RPW + r;
      if (i >= P.Lq) continue;  // warp-uniform
      const size_t qrow = (size_t)(b * P.Lq + i);
      __syncwarp();
      qs[w * D + lane] = ldf((const T*)P.q, qrow * P.q_ld + hd * D + lane);
      qs[w * D + lane + 32] = ldf((const T*)P.q, qrow * P.q_ld + hd * D + lane + 32);
      __syncwarp();
      float s[MAXJ];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const int j = jj * 32 + lane;
        s[jj] = -INFINITY;
        if (jj * 32 < P.Lk && j < klen) {
          float a = 0.f;
          const float* kr = Ks + (size_t)j * DP;
#pragma unroll 16
          for (int d = 0; d < D; d++) a = fmaf(qs[w * D + d], kr[d], a);
          a *= P.scale;
          if (P.dists) a += sw * P.dists[((size_t)b * P.Lq + i) * P.Lk + j] + sb;
          s[jj] = a;
          mx = fmaxf(mx, a);
        }
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const float e = (s[jj] == -INFINITY) ? 0.f : __expf(s[jj] - mx);
        s[jj] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      if (lane == 0) P.lse[((size_t)b * P.H + hd) * P.Lq + i] = mx + __logf(sum);
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const int j = jj * 32 + lane;
        if (jj * 32 < P.Lk && j < P.Lk) {
          const float p = s[jj] * inv;
          pb[r][jj] += p * invH;
          ps[w * MAXL + j] = p * dr.scale((((size_t)b * P.H + hd) * P.Lq + i) * P.Lk + j);
        }
      }
      __syncwarp();
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < klen; j++) {
        const float p = ps[w * MAXL + j];
        o0 = fmaf(p, Vs[(size_t)j * DP + lane], o0);
        o1 = fmaf(p, Vs[(size_t)j * DP + lane + 32], o1);
      }
      T* orow = (T*)P.out + qrow * (size_t)(P.H * D) + hd * D;
      stf(orow, lane, o0);
      stf(orow, lane + 32, o1);
    }
  }
  if (P.pbar) {
#pragma unroll
    for (int r = 0; r < RPW; r++) {
      const int i = row0 + w * RPW + r;
      if (i >= P.Lq) continue;
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const int j = jj * 32 + lane;
        if (j < P.Lk) P.pbar[(size_t)b * P.pbar_bs + (size_t)i * P.pbar_rs + j] = pb[r][jj];
      }
    }
  }
}

// ---- backward pass 1: dQ, delta, d(sprel) ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_bwd_q_kernel(AttnParams P) {
  extern __shared__ float sm[];
  float* Ks = sm;
  float* Vs = Ks + (size_t)P.Lk * DP;
  float* qs = Vs + (size_t)P.Lk * DP;   // [NW][D]
  float* gs = qs + NW * D;              // [NW][D]   dO row
  float* ps = gs + NW * D;              // [NW][MAXL] dS row
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;
  float acc_dw = 0.f, acc_db = 0.f;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Ks, (const T*)P.k, P.k_ld, hd, P.Lk, b * P.Lk);
    load_tile(Vs, (const T*)P.v, P.v_ld, hd, P.Lk, b * P.Lk);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; r++) {
      const int i = row0 + w * RPW + r;
      if (i >= P.Lq) continue;
      const size_t qrow = (size_t)(b * P.Lq + i);
      __syncwarp();
      qs[w * D + lane] = ldf((const T*)P.q, qrow * P.q_ld + hd * D + lane);
      qs[w * D + lane + 32] = ldf((const T*)P.q, qrow * P.q_ld + hd * D + lane + 32);
      gs[w * D + lane] = ldf((const T*)P.dout, qrow * (size_t)(P.H * D) + hd * D + lane);
      gs[w * D + lane + 32] = ldf((const T*)P.dout, qrow * (size_t)(P.H * D) + hd * D + lane + 32);
      __syncwarp();
      const float lse = P.lse[((size_t)b * P.H + hd) * P.Lq + i];
      float p[MAXJ], dp[MAXJ];
      float dl = 0.f;
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const int j = jj * 32 + lane;
        p[jj] = 0.f;
        dp[jj] = 0.f;
        if (jj * 32 < P.Lk && j < klen) {
          float a = 0.f, g = 0.f;
          const float* kr = Ks + (size_t)j * DP;
          const float* vr = Vs + (size_t)j * DP;
#pragma unroll 16
          for (int d = 0; d < D; d++) {
            a = fmaf(qs[w * D + d], kr[d], a);
            g = fmaf(gs[w * D + d], vr[d], g);
          }
          a *= P.scale;
          if (P.dists) a += sw * P.dists[((size_t)b * P.Lq + i) * P.Lk + j] + sb;
          p[jj] = __expf(a - lse);
          g *= dr.scale((((size_t)b * P.H + hd) * P.Lq + i) * P.Lk + j);
          if (P.dpbar) g += P.dpbar[(size_t)b * P.pbar_bs + (size_t)i * P.pbar_rs + j] * invH;
          dp[jj] = g;
          dl = fmaf(p[jj], g, dl);
        }
      }
      dl = warp_sum(dl);
      if (lane == 0) P.delta[((size_t)b * P.H + hd) * P.Lq + i] = dl;
#pragma unroll
      for (int jj = 0; jj < MAXJ; jj++) {
        const int j = jj * 32 + lane;
        if (jj * 32 < P.Lk && j < P.Lk) {
          const float ds = p[jj] * (dp[jj] - dl);
          ps[w * MAXL + j] = ds;
          if (P.dists && j < klen) {
            acc_dw = fmaf(ds, P.dists[((size_t)b * P.Lq + i) * P.Lk + j], acc_dw);
            acc_db += ds;
          }
        }
      }
      __syncwarp();
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < klen; j++) {
        const float ds = ps[w * MAXL + j];
        o0 = fmaf(ds, Ks[(size_t)j * DP + lane], o0);
        o1 = fmaf(ds, Ks[(size_t)j * DP + lane + 32], o1);
      }
      T* drow = (T*)P.dq + qrow * P.dq_ld + hd * D;
      stf(drow, lane, o0 * P.scale);
      stf(drow, lane + 32, o1 * P.scale);
    }
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

// ---- backward pass 2: dK, dV (one warp per key row, lanes sweep queries) ---------------------------
template <typename T>
__global__ void __launch_bounds__(NW * 32) attn_bwd_kv_kernel(AttnParams P) {
  extern __shared__ float sm[];
  float* Qs = sm;                          // [Lq][DP]
  float* Gs = Qs + (size_t)P.Lq * DP;      // [Lq][DP]  dO
  float* ls = Gs + (size_t)P.Lq * DP;      // [Lq] lse
  float* dls = ls + P.Lq;                  // [Lq] delta
  float* ks = dls + P.Lq;                  // [NW][D]
  float* vs = ks + NW * D;                 // [NW][D]
  float* pp = vs + NW * D;                 // [NW][MAXL]  P~ (dropped probs)
  float* pd = pp + NW * MAXL;              // [NW][MAXL]  dS
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.y, row0 = blockIdx.x * ROWS;
  const int klen = P.key_lens ? min(P.Lk, P.key_lens[b]) : P.Lk;
  const float sw = P.dists ? P.sprel_w[0] : 0.f, sb = P.dists ? P.sprel_b[0] : 0.f;
  const Dropout dr = make_dropout(P.drop_p, P.seed_ptr, P.salt);
  const float invH = 1.f / (float)P.H;

  for (int hd = 0; hd < P.H; hd++) {
    __syncthreads();
    load_tile(Qs, (const T*)P.q, P.q_ld, hd, P.Lq, b * P.Lq);
    load_tile(Gs, (const T*)P.dout, (long)P.H * D, hd, P.Lq, b * P.Lq);
    for (int i = threadIdx.x; i < P.Lq; i += blockDim.x) {
      ls[i] = P.lse[((size_t)b * P.H + hd) * P.Lq + i];
      dls[i] = P.delta[((size_t)b * P.H + hd) * P.Lq + i];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; r++) {
      const int j = row0 + w * RPW + r;
      if (j >= P.Lk) continue;
      const size_t krow = (size_t)(b * P.Lk + j);
      T* dkrow = (T*)P.dk + krow * P.dk_ld + hd * D;
      T* dvrow = (T*)P.dv + krow * P.dv_ld + hd * D;
      if (j >= klen) {  // masked key: no gradient
        stf(dkrow, lane, 0.f);
        stf(dkrow, lane + 32, 0.f);
        stf(dvrow, lane, 0.f);
        stf(dvrow, lane + 32, 0.f);
        continue;
      }
      __syncwarp();
      ks[w * D + lane] = ldf((const T*)P.k, krow * P.k_ld + hd * D + lane);
      ks[w * D + lane + 32] = ldf((const T*)P.k, krow * P.k_ld + hd * D + lane + 32);
      vs[w * D + lane] = ldf((const T*)P.v, krow * P.v_ld + hd * D + lane);
      vs[w * D + lane + 32] = ldf((const T*)P.v, krow * P.v_ld + hd * D + lane + 32);
      __syncwarp();
#pragma unroll
      for (int ii = 0; ii < MAXJ; ii++) {
        const int i = ii * 32 + lane;
        if (ii * 32 < P.Lq && i < P.Lq) {
          float a = 0.f, g = 0.f;
          const float* qr = Qs + (size_t)i * DP;
          const float* gr = Gs + (size_t)i * DP;
#pragma unroll 16
          for (int d = 0; d < D; d++) {
            a = fmaf(qr[d], ks[w * D + d], a);
            g = fmaf(gr[d], vs[w * D + d], g);
          }
          a *= P.scale;
          if (P.dists) a += sw * P.dists[((size_t)b * P.Lq + i) * P.Lk + j] + sb;
          const float p = __expf(a - ls[i]);
          const float dsc = dr.scale((((size_t)b * P.H + hd) * P.Lq + i) * P.Lk + j);
          g *= dsc;
          if (P.dpbar) g += P.dpbar[(size_t)b * P.pbar_bs + (size_t)i * P.pbar_rs + j] * invH;
          pp[w * MAXL + i] = p * dsc;
          pd[w * MAXL + i] = p * (g - dls[i]);
        }
      }
      __syncwarp();
      float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
      for (int i = 0; i < P.Lq; i++) {
        const float p = pp[w * MAXL + i], ds = pd[w * MAXL + i];
        dv0 = fmaf(p, Gs[(size_t)i * DP + lane], dv0);
        dv1 = fmaf(p, Gs[(size_t)i * DP + lane + 32], dv1);
        dk0 = fmaf(ds, Qs[(size_t)i * DP + lane], dk0);
        dk1 = fmaf(ds, Qs[(size_t)i * DP + lane + 32], dk1);
      }
      stf(dkrow, lane, dk0 * P.scale);
      stf(dkrow, lane + 32, dk1 * P.scale);
      stf(dvrow, lane, dv0);
      stf(dvrow, lane + 32, dv1);
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
  return P;
}

}  // namespace

extern "C" {

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
  const size_t smem = ((size_t)2 * Lk * DP + NW * D + NW * MAXL) * sizeof(float);
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
  const size_t smem1 = ((size_t)2 * Lk * DP + 2 * NW * D + NW * MAXL) * sizeof(float);
  const size_t smem2 = ((size_t)2 * Lq * DP + 2 * Lq + 2 * NW * D + 2 * NW * MAXL) * sizeof(float);
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

}  // extern "C"
