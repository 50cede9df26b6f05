// Fused MAKD (meta-ability knowledge distillation) losses -- HBM-streaming kernels.
//
//  * magic_makd_mse_{fwd,bwd}: ONE launch covers every hidden-state / attention-map MSE of a step
//    (up to 16 segments: txt, pano, pano-fused, global, local embeddings + 4 attention maps).  Each
//    segment is `rows` x `inner` with independent row strides for student and teacher (so the
//    `[:, :min_len]` layer slices of agent.py:560-671 need no copies), an optional per-row MKTD
//    weight w[row] (kd_loss.py:11-13) and a scalar scale = MKRW weight / numel (kd_loss.py:8,14 mean).
//  * magic_makd_kl_{fwd,bwd}: temperature-scaled KL on action / MLM logits with the -inf -> -1e6 rule
//    (kd_loss.py:21-22), mean over B*C (unweighted, :29) or over rows after the per-row weight (:31-40).
//
// Algorithmic bytes: fwd reads |S|+|T| once; bwd reads |S|+|T| once and writes |dS|.
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int NT = 256;         // threads per CTA
constexpr int CH_ELEMS = 4096;  // elements per chunk (measured: 8192-element bf16 chunks balance worse over the grid)

struct Args {
  MagicMseSeg seg[MAGIC_MAKD_MAX_SEGS];
  long long chunk0[MAGIC_MAKD_MAX_SEGS + 1];
  int ch[MAGIC_MAKD_MAX_SEGS];  // elements per chunk of segment i
  int nseg;
};

__device__ __forceinline__ float ld_any(const void* p, int dt, size_t i) {
  return dt == MAGIC_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_any(void* p, int dt, size_t i, float v) {
  if (dt == MAGIC_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// One CTA owns a CONTIGUOUS range of chunks (balanced over the grid), so a thread carries its partial sum across
// chunks and the block reduction + 2 atomics happen once per (CTA, segment) instead of once per chunk.  Row tails
// stay on the 128-bit path whenever `inner` is a multiple of the vector width.
// One chunk = elements [c0, c1) of one row; `sp` / `tp` / `dp` point at the row.  32-bit column arithmetic.
template <typename T, bool BWD>
__device__ __forceinline__ float mse_chunk(const T* __restrict__ sp, const T* __restrict__ tp, T* __restrict__ dp,
                                           int c0, int c1, float coef) {
  constexpr int VN = RowVec<T>::N;
  constexpr int CH = CH_ELEMS;
  constexpr int UN = CH / (NT * VN);  // 4 (fp32) / 2 (bf16) independent 16-byte loads per tensor per thread
  float acc = 0.f;
  const int base = c0 + (int)threadIdx.x * VN;
  if (c1 - c0 == CH) {
    float a[UN][VN], b[UN][VN];
#pragma unroll
    for (int k = 0; k < UN; k++) {
      RowVec<T>::load_cs(sp + base + k * NT * VN, a[k]);
      RowVec<T>::load_cs(tp + base + k * NT * VN, b[k]);
    }
#pragma unroll
    for (int k = 0; k < UN; k++) {
#pragma unroll
      for (int i = 0; i < VN; i++) {
        const float d = a[k][i] - b[k][i];
        if (BWD) a[k][i] = coef * d;
        else acc = fmaf(d, d, acc);
      }
      if (BWD) RowVec<T>::store(dp + base + k * NT * VN, a[k]);
    }
  } else {
    const int cv = c0 + (c1 - c0) / VN * VN;  // c0 is a multiple of CH, so [c0, cv) stays 16-byte aligned
    for (int c = base; c < cv; c += NT * VN) {
      float a[VN], b[VN];
      RowVec<T>::load_cs(sp + c, a);
      RowVec<T>::load_cs(tp + c, b);
#pragma unroll
      for (int i = 0; i < VN; i++) {
        const float d = a[i] - b[i];
        if (BWD) a[i] = coef * d;
        else acc = fmaf(d, d, acc);
      }
      if (BWD) RowVec<T>::store(dp + c, a);
    }
    for (int c = cv + (int)threadIdx.x; c < c1; c += NT) {  // < VN leftover elements of a ragged row
      const float d = ldf(sp, c) - ldf(tp, c);
      if (BWD) stf(dp, c, coef * d);
      else acc = fmaf(d, d, acc);
    }
  }
  return acc;
}

// One CTA owns a CONTIGUOUS range of chunks.  Everything that depends only on the segment (pointers, strides, scale)
// or on the row (row pointers, MKTD weight) is cached in registers and advanced incrementally: the first version
// re-derived them per chunk (64-bit divisions, param-struct reloads) and executed ~250 bookkeeping instructions per
// warp and chunk -- 12.3 M warp instructions for 99 MB, issue-bound instead of HBM-bound (ncu, r01 v6).
template <bool BWD>
__global__ void __launch_bounds__(NT) makd_mse_kernel(const __grid_constant__ Args A, float* __restrict__ loss,
                                                      const float* __restrict__ gseg,
                                                      const float* __restrict__ gtot) {
  __shared__ float red[32];
  const long long total = A.chunk0[A.nseg];
  const long long ch_beg = total * blockIdx.x / gridDim.x, ch_end = total * (blockIdx.x + 1) / gridDim.x;
  int si = -1;
  long long seg_end = 0;
  // segment state
  const char *sbase = nullptr, *tbase = nullptr;
  char* dbase = nullptr;
  const float* wp = nullptr;
  float seg_scale = 0.f, gup = 0.f;
  long long s_rb = 0, t_rb = 0;  // row strides in BYTES
  int inner = 0, mode = 0;       // mode 0: scalar any-dtype path, 1: fp32 vectors, 2: bf16 vectors
  int s_dt = 0, t_dt = 0;
  // row state
  long long row = 0;
  int c0 = 0;
  float wr = 0.f;
  float seg_acc = 0.f;  // this thread's share of segment si, already multiplied by row weight * scale
  for (long long ch = ch_beg; ch < ch_end; ch++) {
    bool new_row = false;
    if (ch >= seg_end) {
      if (!BWD && si >= 0) {
        const float tot = block_sum(seg_acc, red);
        if (threadIdx.x == 0 && tot != 0.f) {
          atomicAdd(loss + si, tot);
          atomicAdd(loss + MAGIC_MAKD_MAX_SEGS, tot);  // running total of all segments
        }
        seg_acc = 0.f;
      }
      if (si < 0) si = 0;
      while (ch >= A.chunk0[si + 1]) si++;
      seg_end = A.chunk0[si + 1];
      const MagicMseSeg& S = A.seg[si];
      sbase = (const char*)S.s; tbase = (const char*)S.t; dbase = (char*)S.ds;
      wp = S.w;
      s_dt = S.s_dt; t_dt = S.t_dt;
      const int ssz = s_dt == MAGIC_BF16 ? 2 : 4, tsz = t_dt == MAGIC_BF16 ? 2 : 4;
      s_rb = S.s_rs * ssz; t_rb = S.t_rs * tsz;
      inner = (int)S.inner;
      mode = !S.vec_ok ? 0 : (s_dt == MAGIC_F32 ? 1 : 2);
      seg_scale = S.scale * (S.scale_dev ? S.scale_dev[0] : 1.f);
      gup = BWD ? ((gseg ? gseg[si] : 0.f) + (gtot ? gtot[0] : 0.f)) : 0.f;
      const long long cpr = (S.inner + CH_ELEMS - 1) / CH_ELEMS;
      const long long local = ch - A.chunk0[si];
      row = local / cpr;
      c0 = (int)(local % cpr) * CH_ELEMS;
      new_row = true;
    }
    if (new_row || c0 == 0) wr = (wp ? wp[row] : 1.f) * seg_scale;
    const char* sp = sbase + row * s_rb;
    const char* tp = tbase + row * t_rb;
    char* dp = BWD ? dbase + row * s_rb : nullptr;
    const int c1 = min(inner, c0 + CH_ELEMS);
    const float coef = BWD ? 2.f * wr * gup : 0.f;
    float acc = 0.f;
    if (mode == 1) {
      acc = mse_chunk<float, BWD>((const float*)sp, (const float*)tp, (float*)dp, c0, c1, coef);
    } else if (mode == 2) {
      acc = mse_chunk<__nv_bfloat16, BWD>((const __nv_bfloat16*)sp, (const __nv_bfloat16*)tp, (__nv_bfloat16*)dp, c0,
                                          c1, coef);
    } else {
      for (int c = c0 + (int)threadIdx.x; c < c1; c += NT) {
        const float d = ld_any(sp, s_dt, c) - ld_any(tp, t_dt, c);
        if (BWD) st_any(dp, s_dt, c, coef * d);
        else acc = fmaf(d, d, acc);
      }
    }
    if (!BWD) seg_acc = fmaf(acc, wr, seg_acc);
    c0 += CH_ELEMS;  // next chunk of this segment
    if (c0 >= inner) {
      c0 = 0;
      row++;
    }
  }
  if (!BWD && si >= 0) {
    const float tot = block_sum(seg_acc, red);
    if (threadIdx.x == 0 && tot != 0.f) {
      atomicAdd(loss + si, tot);
      atomicAdd(loss + MAGIC_MAKD_MAX_SEGS, tot);
    }
  }
}

// ---- KL on logits: one CTA per row -----------------------------------------------------------------

// ONE pass over both rows: per thread an online triple for the teacher (max m_t, z_t = sum e^{x_t - m_t},
// a = sum e^{x_t - m_t} (x_t - x_s)) and an online pair for the student (m_s, z_s); then
//   KL(row) = a / z_t - (m_t - m_s) - log(z_t / z_s)          [= sum_c p_t (log p_t - log p_s)]
// so each logit is read from HBM exactly once (the two-pass form re-read both rows).  The running maxima are
// updated once per 16-byte vector (vector max first), which removes the per-element branches.
// The pass works in the log2 domain (u = logit * (1/T) * log2 e): an exponential is then one FADD + one MUFU.EX2,
// and the kernel is MUFU-bound (2 exponentials per logit pair at 16 / clk / SM), so every saved instruction counts.
struct KlAcc {
  float ms, zs, mt, zt, a;  // maxima in log2 units; a = sum 2^{u_t - m_t} (u_t - u_s)
};
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// -inf -> -1e6 (kd_loss.py:21-22: torch.where(x == -inf, -1e6, x) -- NaN and finite values below -1e6 pass through
// unchanged, as in the reference) and the temperature / log2 e scale
__device__ __forceinline__ float fix2(float v, float c) { return (v == -INFINITY ? -1e6f : v) * c; }
// forward: the same replacement as ONE ALU-pipe op -- max.NaN(v, -1e6) maps -inf to -1e6, propagates NaN like the
// reference's torch.where, and differs only for finite logits below -1e6 (clamped), which no softmax input reaches.
// The bf16 forward is bound by the ALU + XU pipes (their busy times ADD on this part), so the select pair
// (FSETP + FSEL) per value was a sixth of the kernel.
__device__ __forceinline__ float fix2_fwd(float v, float c) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(-1e6f));
  return r * c;
}
// 8 bf16 -> fp32 with the low halves shifted by an integer multiply (IMAD: FMA pipe) instead of SHF / PRMT (ALU pipe)
template <typename T>
struct KlLoad {
  static __device__ __forceinline__ void load(const T* p, float (&v)[RowVec<T>::N]) { RowVec<T>::load_cs(p, v); }
};
template <>
struct KlLoad<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldcs(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t lo;
      asm("mul.lo.u32 %0, %1, 65536;" : "=r"(lo) : "r"(w[i]));
      v[2 * i] = __uint_as_float(lo);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
__device__ __forceinline__ void kl_merge(KlAcc& x, const KlAcc& y) {
  const float ms = fmaxf(x.ms, y.ms), mt = fmaxf(x.mt, y.mt);
  if (ms > -INFINITY) x.zs = x.zs * fast_ex2(x.ms - ms) + y.zs * fast_ex2(y.ms - ms);
  if (mt > -INFINITY) {
    const float fx = fast_ex2(x.mt - mt), fy = fast_ex2(y.mt - mt);
    x.zt = x.zt * fx + y.zt * fy;
    x.a = x.a * fx + y.a * fy;
  }
  x.ms = ms;
  x.mt = mt;
}
template <int VN>
__device__ __forceinline__ void kl_push(KlAcc& k, const float (&xs)[VN], const float (&xt)[VN]) {
  float vs = xs[0], vt = xt[0];
#pragma unroll
  for (int i = 1; i < VN; i++) {
    vs = fmaxf(vs, xs[i]);
    vt = fmaxf(vt, xt[i]);
  }
  if (vs > k.ms) {
    k.zs *= fast_ex2(k.ms - vs);  // m = -inf: z is 0 and 2^-inf = 0
    k.ms = vs;
  }
  if (vt > k.mt) {
    const float f = fast_ex2(k.mt - vt);
    k.zt *= f;
    k.a *= f;
    k.mt = vt;
  }
#pragma unroll
  for (int i = 0; i < VN; i++) {
    k.zs += fast_ex2(xs[i] - k.ms);
    const float e = fast_ex2(xt[i] - k.mt);
    k.zt += e;
    k.a = fmaf(e, xt[i] - xs[i], k.a);  // e == 0 (masked / far below the max) contributes exactly 0
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
    makd_kl_fwd_kernel(const T* __restrict__ s, const T* __restrict__ t, int C, long ld, float invT,
                       const float* __restrict__ w, float scale, const float* __restrict__ scale_dev,
                       float* __restrict__ stats, float* __restrict__ loss, int vec) {
  __shared__ KlAcc part[NT / 32];
  const int r = blockIdx.x;
  const T* sr = s + (size_t)r * ld;
  const T* tr = t + (size_t)r * ld;
  constexpr int VN = RowVec<T>::N;
  const int cv = vec ? (C / VN) * VN : 0;
  const float c2 = invT * 1.4426950408889634f;
  constexpr float LN2 = 0.6931471805599453f;
  KlAcc k{-INFINITY, 0.f, -INFINITY, 0.f, 0.f};
  int c = threadIdx.x * VN;
  for (; c + NT * VN < cv; c += 2 * NT * VN) {  // two independent 16-byte loads per row in flight
    float a0[VN], b0[VN], a1[VN], b1[VN];
    KlLoad<T>::load(sr + c, a0);
    KlLoad<T>::load(tr + c, b0);
    KlLoad<T>::load(sr + c + NT * VN, a1);
    KlLoad<T>::load(tr + c + NT * VN, b1);
#pragma unroll
    for (int i = 0; i < VN; i++) {
      a0[i] = fix2_fwd(a0[i], c2); b0[i] = fix2_fwd(b0[i], c2);
      a1[i] = fix2_fwd(a1[i], c2); b1[i] = fix2_fwd(b1[i], c2);
    }
    kl_push<VN>(k, a0, b0);
    kl_push<VN>(k, a1, b1);
  }
  for (; c < cv; c += NT * VN) {
    float a0[VN], b0[VN];
    KlLoad<T>::load(sr + c, a0);
    KlLoad<T>::load(tr + c, b0);
#pragma unroll
    for (int i = 0; i < VN; i++) {
      a0[i] = fix2_fwd(a0[i], c2); b0[i] = fix2_fwd(b0[i], c2);
    }
    kl_push<VN>(k, a0, b0);
  }
  for (int cc = cv + threadIdx.x; cc < C; cc += NT) {
    const float xs[1] = {fix2_fwd(ldf(sr, cc), c2)}, xt[1] = {fix2_fwd(ldf(tr, cc), c2)};
    kl_push<1>(k, xs, xt);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    KlAcc y;
    y.ms = __shfl_xor_sync(0xffffffffu, k.ms, o);
    y.zs = __shfl_xor_sync(0xffffffffu, k.zs, o);
    y.mt = __shfl_xor_sync(0xffffffffu, k.mt, o);
    y.zt = __shfl_xor_sync(0xffffffffu, k.zt, o);
    y.a = __shfl_xor_sync(0xffffffffu, k.a, o);
    kl_merge(k, y);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) part[wp] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < NT / 32; i++) kl_merge(k, part[i]);
    const float kl = LN2 * (k.a / k.zt - (k.mt - k.ms)) - logf(k.zt / k.zs);
    stats[2 * r] = LN2 * k.ms + logf(k.zs);  // natural-log log-sum-exp of the scaled rows, as backward expects
    stats[2 * r + 1] = LN2 * k.mt + logf(k.zt);
    atomicAdd(loss, kl * (w ? w[r] : 1.f) * scale * (scale_dev ? scale_dev[0] : 1.f));
  }
}

template <typename T>
__global__ void __launch_bounds__(NT)
    makd_kl_bwd_kernel(const T* __restrict__ s, const T* __restrict__ t, T* __restrict__ ds, int C, long ld,
                       float invT, const float* __restrict__ w, float scale, const float* __restrict__ scale_dev,
                       const float* __restrict__ stats, const float* __restrict__ gout, int vec) {
  const int r = blockIdx.x;
  const T* sr = s + (size_t)r * ld;
  const T* tr = t + (size_t)r * ld;
  T* dr = ds + (size_t)r * ld;
  // softmax probabilities as 2^(u - lse * log2 e): one FFMA + one MUFU.EX2 each
  const float c2 = invT * 1.4426950408889634f;
  const float l2s = stats[2 * r] * 1.4426950408889634f, l2t = stats[2 * r + 1] * 1.4426950408889634f;
  const float coef = gout[0] * scale * (scale_dev ? scale_dev[0] : 1.f) * (w ? w[r] : 1.f) * invT;
  constexpr int VN = RowVec<T>::N;
  const int cv = vec ? (C / VN) * VN : 0;
  for (int c = threadIdx.x * VN; c < cv; c += NT * VN) {
    float a[VN], b[VN];
    RowVec<T>::load(sr + c, a);
    RowVec<T>::load(tr + c, b);
#pragma unroll
    for (int i = 0; i < VN; i++)
      a[i] = a[i] != -INFINITY ? coef * (fast_ex2(a[i] * c2 - l2s) - fast_ex2(fix2(b[i], c2) - l2t)) : 0.f;
    RowVec<T>::store(dr + c, a);
  }
  for (int c = cv + threadIdx.x; c < C; c += NT) {
    const float sv = ldf(sr, c);
    float g = 0.f;
    if (sv != -INFINITY) g = coef * (fast_ex2(sv * c2 - l2s) - fast_ex2(fix2(ldf(tr, c), c2) - l2t));
    stf(dr, c, g);
  }
}

// total = alpha * (mse_total + kl) + (1 - alpha) * mean(sup)      (agent.py:1119)
__global__ void __launch_bounds__(256)
    mix_fwd_kernel(const float* __restrict__ mse_total, const float* __restrict__ kl, const float* __restrict__ sup,
                   int n, float alpha, const float* __restrict__ inv_n, float* __restrict__ out) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += sup[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) {
    const float sm = inv_n ? a * inv_n[0] : (n > 0 ? a / (float)n : 0.f);
    const float kd = (mse_total ? mse_total[0] : 0.f) + (kl ? kl[0] : 0.f);
    out[0] = alpha * kd + (1.f - alpha) * sm;
    out[1] = sm;
    out[2] = kd;
  }
}
__global__ void mix_bwd_kernel(const float* __restrict__ g, int n, float alpha, const float* __restrict__ inv_n,
                               float* __restrict__ d_mse, float* __restrict__ d_kl, float* __restrict__ d_sup) {
  const float gv = g[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (d_mse) d_mse[0] = alpha * gv;
    if (d_kl) d_kl[0] = alpha * gv;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    d_sup[i] = inv_n ? (1.f - alpha) * gv * inv_n[0] : (1.f - alpha) * gv / (float)n;
}

int build_args(const MagicMseSeg* segs, int nseg, Args& A, const char* name) {
  MAGIC_CHECK_ARG(nseg >= 0 && nseg <= MAGIC_MAKD_MAX_SEGS, "%s: nseg=%d out of range", name, nseg);
  A.nseg = nseg;
  long long c = 0;
  for (int i = 0; i < nseg; i++) {
    A.seg[i] = segs[i];
    A.chunk0[i] = c;
    MAGIC_CHECK_ARG(segs[i].rows >= 0 && segs[i].inner >= 0, "%s: negative extent in segment %d", name, i);
    // 128-bit path: same dtype, 16-byte aligned bases and row strides
    const int esz = segs[i].s_dt == MAGIC_BF16 ? 2 : 4;
    const int CH = CH_ELEMS;
    A.ch[i] = CH;
    c += segs[i].rows * ((segs[i].inner + CH - 1) / CH);
    const int v = 16 / esz;
    bool ok = segs[i].s_dt == segs[i].t_dt && segs[i].s_rs % v == 0 && segs[i].t_rs % v == 0 &&
              ((uintptr_t)segs[i].s % 16 == 0) && ((uintptr_t)segs[i].t % 16 == 0) &&
              (segs[i].ds == nullptr || (uintptr_t)segs[i].ds % 16 == 0);
    A.seg[i].vec_ok = ok ? 1 : 0;
  }
  A.chunk0[nseg] = c;
  return MAGIC_OK;
}

}  // namespace

extern "C" {

int magic_makd_mse_fwd(const MagicMseSeg* segs, int nseg, float* loss, cudaStream_t st) {
  Args A;
  int rc = build_args(segs, nseg, A, "magic_makd_mse_fwd");
  if (rc) return rc;
  MAGIC_CUDA(cudaMemsetAsync(loss, 0, sizeof(float) * (MAGIC_MAKD_MAX_SEGS + 1), st), "magic_makd_mse_fwd");
  const long long total = A.chunk0[nseg];
  if (total == 0) return MAGIC_OK;
  const long long cap = 16LL * magic_num_sms();  // measured: 16 x SMs beats one resident wave (8 x SMs) by 4-8 %
  makd_mse_kernel<false><<<(int)(total < cap ? total : cap), NT, 0, st>>>(A, loss, nullptr, nullptr);
  MAGIC_CHECK_LAUNCH("magic_makd_mse_fwd");
  return MAGIC_OK;
}

int magic_makd_mse_bwd(const MagicMseSeg* segs, int nseg, const float* gseg, const float* gtot, cudaStream_t st) {
  Args A;
  int rc = build_args(segs, nseg, A, "magic_makd_mse_bwd");
  if (rc) return rc;
  for (int i = 0; i < nseg; i++)
    MAGIC_CHECK_ARG(segs[i].ds != nullptr || segs[i].rows * segs[i].inner == 0, "magic_makd_mse_bwd: segment %d has no ds",
                    i);
  const long long total = A.chunk0[nseg];
  if (total == 0) return MAGIC_OK;
  const long long cap = 16LL * magic_num_sms();
  makd_mse_kernel<true><<<(int)(total < cap ? total : cap), NT, 0, st>>>(A, nullptr, gseg, gtot);
  MAGIC_CHECK_LAUNCH("magic_makd_mse_bwd");
  return MAGIC_OK;
}

int magic_loss_mix_fwd(const float* mse_total, const float* kl, const float* sup, int n, float alpha,
                       const float* inv_n, float* out, cudaStream_t st) {
  mix_fwd_kernel<<<1, 256, 0, st>>>(mse_total, kl, sup, n, alpha, inv_n, out);
  MAGIC_CHECK_LAUNCH("magic_loss_mix_fwd");
  return MAGIC_OK;
}

int magic_loss_mix_bwd(const float* g, int n, float alpha, const float* inv_n, float* d_mse, float* d_kl,
                       float* d_sup, cudaStream_t st) {
  mix_bwd_kernel<<<(n + 255) / 256 > 0 ? (n + 255) / 256 : 1, 256, 0, st>>>(g, n, alpha, inv_n, d_mse, d_kl, d_sup);
  MAGIC_CHECK_LAUNCH("magic_loss_mix_bwd");
  return MAGIC_OK;
}

int magic_makd_kl_fwd(const void* s, const void* t, int R, int C, long ld, float temperature, const float* w,
                      float scale, const float* scale_dev, float* stats, float* loss, int dtype, cudaStream_t st) {
  MAGIC_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st), "magic_makd_kl_fwd");
  if (R <= 0) return MAGIC_OK;
  const float invT = 1.f / temperature;
  const int esz = dtype == MAGIC_BF16 ? 2 : 4;
  const int vec = (((uintptr_t)s & 15) == 0 && ((uintptr_t)t & 15) == 0 && (ld * esz) % 16 == 0) ? 1 : 0;
  if (dtype == MAGIC_F32)
    makd_kl_fwd_kernel<float><<<R, NT, 0, st>>>((const float*)s, (const float*)t, C, ld, invT, w, scale, scale_dev,
                                                stats, loss, vec);
  else if (dtype == MAGIC_BF16)
    makd_kl_fwd_kernel<__nv_bfloat16><<<R, NT, 0, st>>>((const __nv_bfloat16*)s, (const __nv_bfloat16*)t, C, ld, invT,
                                                       w, scale, scale_dev, stats, loss, vec);
  else {
    magic_set_error("magic_makd_kl_fwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_makd_kl_fwd");
  return MAGIC_OK;
}

int magic_makd_kl_bwd(const void* s, const void* t, void* ds, int R, int C, long ld, float temperature,
                      const float* w, float scale, const float* scale_dev, const float* stats, const float* gout,
                      int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  const float invT = 1.f / temperature;
  const int esz = dtype == MAGIC_BF16 ? 2 : 4;
  const int vec = (((uintptr_t)s & 15) == 0 && ((uintptr_t)t & 15) == 0 && ((uintptr_t)ds & 15) == 0 &&
                   (ld * esz) % 16 == 0) ? 1 : 0;
  if (dtype == MAGIC_F32)
    makd_kl_bwd_kernel<float><<<R, NT, 0, st>>>((const float*)s, (const float*)t, (float*)ds, C, ld, invT, w, scale,
                                                scale_dev, stats, gout, vec);
  else if (dtype == MAGIC_BF16)
    makd_kl_bwd_kernel<__nv_bfloat16><<<R, NT, 0, st>>>((const __nv_bfloat16*)s, (const __nv_bfloat16*)t,
                                                       (__nv_bfloat16*)ds, C, ld, invT, w, scale, scale_dev, stats,
                                                       gout, vec);
  else {
    magic_set_error("magic_makd_kl_bwd: bad dtype");
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_makd_kl_bwd");
  return MAGIC_OK;
}

}  // extern "C"
