// Kernels of the MRC (masked region classification) and CFP (cross-modal feature pooling, InfoNCE) task heads:
//   * magic_zero_rows        zero the masked views of the last-step panorama in place (data/tasks.py:178-181)
//   * magic_soft_ce_{fwd,bwd} KL(target || softmax(logits)) per row against soft labels over 1000 classes
//                             (validate_mrc, pretrain_src/train_r2r_magic.py:483-486)
//   * magic_l2norm_{fwd,bwd}  row-wise x / max(||x||, eps) for the contrastive features (validate_cfp, :545-560)
// One warp per row; all reductions fp32.
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int WARPS = 8;

int grid_for(int rows) {
  int g = (rows + WARPS - 1) / WARPS;
  const int cap = magic_num_sms() * 8;
  return g < 1 ? 1 : (g < cap ? g : cap);
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
    zero_rows_kernel(T* __restrict__ x, const long long* __restrict__ rows, int n, int h) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * WARPS + w; r < n; r += gridDim.x * WARPS) {
    const long long row = rows[r];
    if (row < 0) continue;  // padding entry
    T* p = x + (size_t)row * h;
    for (int c = lane; c < h; c += 32) stf(p, c, 0.f);
  }
}

// loss[r] = sum_c t*log(t) - t*x + lse * sum_c t      (t == 0 contributes exactly 0, like F.kl_div)
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
    soft_ce_fwd_kernel(const T* __restrict__ logits, const float* __restrict__ tgt, float* __restrict__ loss,
                       float* __restrict__ stats, int R, int C, long ld, long tld) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * WARPS + w; r < R; r += gridDim.x * WARPS) {
    const T* x = logits + (size_t)r * ld;
    const float* t = tgt + (size_t)r * tld;
    float m = -INFINITY, s = 0.f, tsum = 0.f, acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xv = ldf(x, c), tv = t[c];
      lse_push(m, s, xv);
      tsum += tv;
      if (tv > 0.f) acc += tv * (logf(tv) - xv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    tsum = warp_sum(tsum);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float lse = m + logf(s);
      loss[r] = acc + lse * tsum;
      stats[2 * r] = lse;
      stats[2 * r + 1] = tsum;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
    soft_ce_bwd_kernel(const T* __restrict__ logits, const float* __restrict__ tgt, const float* __restrict__ stats,
                       const float* __restrict__ dloss, T* __restrict__ dlogits, int R, int C, long ld, long tld) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * WARPS + w; r < R; r += gridDim.x * WARPS) {
    const T* x = logits + (size_t)r * ld;
    const float* t = tgt + (size_t)r * tld;
    T* d = dlogits + (size_t)r * ld;
    const float lse = stats[2 * r], tsum = stats[2 * r + 1], g = dloss[r];
    for (int c = lane; c < C; c += 32) stf(d, c, g * (expf(ldf(x, c) - lse) * tsum - t[c]));
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
    l2norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ inv, int R, int h, float eps) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * WARPS + w; r < R; r += gridDim.x * WARPS) {
    const T* p = x + (size_t)r * h;
    float q = 0.f;
    for (int c = lane; c < h; c += 32) {
      const float v = ldf(p, c);
      q = fmaf(v, v, q);
    }
    q = warp_sum(q);
    const float iv = 1.f / fmaxf(sqrtf(q), eps);
    for (int c = lane; c < h; c += 32) stf(y + (size_t)r * h, c, ldf(p, c) * iv);
    if (lane == 0) inv[r] = iv;
  }
}

// dx = inv * (dy - y * (y . dy))   (exact for ||x|| > eps, which holds for any non-degenerate feature)
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
    l2norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ inv,
                      T* __restrict__ dx, int R, int h) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * WARPS + w; r < R; r += gridDim.x * WARPS) {
    const size_t b = (size_t)r * h;
    float dot = 0.f;
    for (int c = lane; c < h; c += 32) dot = fmaf(ldf(y, b + c), ldf(dy, b + c), dot);
    dot = warp_sum(dot);
    const float iv = inv[r];
    for (int c = lane; c < h; c += 32) stf(dx, b + c, iv * (ldf(dy, b + c) - ldf(y, b + c) * dot));
  }
}

}  // namespace

#define HEADS_DISPATCH(dt, name, ...)                         \
  if ((dt) == MAGIC_F32) {                                    \
    typedef float T;                                          \
    __VA_ARGS__;                                              \
  } else if ((dt) == MAGIC_BF16) {                            \
    typedef __nv_bfloat16 T;                                  \
    __VA_ARGS__;                                              \
  } else {                                                    \
    magic_set_error("%s: bad dtype %d", name, (int)(dt));     \
    return MAGIC_ERR_ARG;                                     \
  }

extern "C" {

int magic_zero_rows(void* x, const long long* rows, int n, int h, int dtype, cudaStream_t st) {
  if (n <= 0 || h <= 0) return MAGIC_OK;
  HEADS_DISPATCH(dtype, "magic_zero_rows", (zero_rows_kernel<T><<<grid_for(n), WARPS * 32, 0, st>>>((T*)x, rows, n, h)));
  MAGIC_CHECK_LAUNCH("magic_zero_rows");
  return MAGIC_OK;
}

int magic_soft_ce_fwd(const void* logits, const float* targets, float* loss, float* stats, int R, int C, long ld,
                      long tld, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  MAGIC_CHECK_ARG(C > 0 && ld >= C && tld >= C, "magic_soft_ce_fwd: bad extents C=%d ld=%ld tld=%ld", C, ld, tld);
  HEADS_DISPATCH(dtype, "magic_soft_ce_fwd",
                 (soft_ce_fwd_kernel<T><<<grid_for(R), WARPS * 32, 0, st>>>((const T*)logits, targets, loss, stats, R,
                                                                           C, ld, tld)));
  MAGIC_CHECK_LAUNCH("magic_soft_ce_fwd");
  return MAGIC_OK;
}

int magic_soft_ce_bwd(const void* logits, const float* targets, const float* stats, const float* dloss,
                      void* dlogits, int R, int C, long ld, long tld, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  HEADS_DISPATCH(dtype, "magic_soft_ce_bwd",
                 (soft_ce_bwd_kernel<T><<<grid_for(R), WARPS * 32, 0, st>>>((const T*)logits, targets, stats, dloss,
                                                                           (T*)dlogits, R, C, ld, tld)));
  MAGIC_CHECK_LAUNCH("magic_soft_ce_bwd");
  return MAGIC_OK;
}

int magic_l2norm_fwd(const void* x, void* y, float* inv_norm, int R, int h, float eps, int dtype, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  HEADS_DISPATCH(dtype, "magic_l2norm_fwd",
                 (l2norm_fwd_kernel<T><<<grid_for(R), WARPS * 32, 0, st>>>((const T*)x, (T*)y, inv_norm, R, h, eps)));
  MAGIC_CHECK_LAUNCH("magic_l2norm_fwd");
  return MAGIC_OK;
}

int magic_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int R, int h, int dtype,
                     cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  HEADS_DISPATCH(dtype, "magic_l2norm_bwd",
                 (l2norm_bwd_kernel<T><<<grid_for(R), WARPS * 32, 0, st>>>((const T*)dy, (const T*)y, inv_norm, (T*)dx,
                                                                          R, h)));
  MAGIC_CHECK_LAUNCH("magic_l2norm_bwd");
  return MAGIC_OK;
}

}  // extern "C"
