// GPU batch featuriser, feature half: the panorama view features live in HBM once (the reference
// keeps them in a host-side cache read per sample, pretrain_src/data/dataset.py:210-244, and ships 36 x 768 floats per
// step over PCIe); a batch then only carries, per trajectory step, the index of its panorama and the order of its
// views (candidate views first, then the remaining views ascending: dataset.py:742-756), and these kernels build
// `traj_view_img_fts` and `gmap_pair_dists` directly in device memory.
//
//   magic_gather_views       out[r, j, :] = perm[r, j] >= 0 ? store[vp[r], perm[r, j], :] : 0      (HBM-bound copy)
//   magic_gather_pair_dists  out[b, i, j] = (v_i, v_j >= 0) ? dist[v_i * N + v_j] : 0,  v = node_vp[b, :]
//                            (the all-pairs shortest-path matrix of dataset.py:545-549, resident on the device)
#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

// one warp per output row of D elements; 16-byte accesses (D * esz is a multiple of 16)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
    gather_views_kernel(const TI* __restrict__ store, const long long* __restrict__ vp, const int* __restrict__ perm,
                        TO* __restrict__ out, long long rows, int V, int D, long long n_store) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    const long long r = row / V;
    const int j = (int)(row - r * V);
    const int pv = perm[row];
    const long long v = vp[r];
    TO* dst = out + (size_t)row * D;
    if (pv < 0 || v < 0 || v >= n_store) {
      for (int c = lane; c < D; c += 32) dst[c] = (TO)0.f;
      continue;
    }
    const TI* src = store + ((size_t)v * V + pv) * D;
    if (sizeof(TI) == sizeof(TO)) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      const int n4 = D * (int)sizeof(TI) / 16;
      for (int c = lane; c < n4; c += 32) d4[c] = __ldg(s4 + c);
    } else {
      for (int c = lane; c < D; c += 32) dst[c] = (TO)(float)src[c];
    }
    (void)j;
  }
}

__global__ void pair_dists_kernel(const float* __restrict__ dist, long long N, const long long* __restrict__ node_vp,
                                  float* __restrict__ out, int B, int G) {
  const long long total = (long long)B * G * G;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e % G);
    const int i = (int)((e / G) % G);
    const long long b = e / ((long long)G * G);
    const long long vi = node_vp[b * G + i], vj = node_vp[b * G + j];
    out[e] = (vi >= 0 && vj >= 0 && vi < N && vj < N) ? dist[vi * N + vj] : 0.f;
  }
}

}  // namespace

extern "C" {

int magic_gather_views(const void* store, int store_dt, long long n_store, const long long* vp, const int* perm,
                       void* out, int out_dt, long long R, int V, int D, cudaStream_t st) {
  if (R <= 0) return MAGIC_OK;
  MAGIC_CHECK_ARG(V > 0 && D > 0 && D % 8 == 0, "magic_gather_views: bad V=%d D=%d", V, D);
  MAGIC_CHECK_ARG(((uintptr_t)store % 16) == 0 && ((uintptr_t)out % 16) == 0, "magic_gather_views: unaligned pointer");
  const long long rows = R * V;
  long long blocks = (rows + 7) / 8;
  const long long cap = 16LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  typedef __nv_bfloat16 bf;
  if (store_dt == MAGIC_BF16 && out_dt == MAGIC_BF16)
    gather_views_kernel<bf, bf><<<(int)blocks, 256, 0, st>>>((const bf*)store, vp, perm, (bf*)out, rows, V, D, n_store);
  else if (store_dt == MAGIC_F32 && out_dt == MAGIC_F32)
    gather_views_kernel<float, float><<<(int)blocks, 256, 0, st>>>((const float*)store, vp, perm, (float*)out, rows, V, D,
                                                                  n_store);
  else if (store_dt == MAGIC_BF16 && out_dt == MAGIC_F32)
    gather_views_kernel<bf, float><<<(int)blocks, 256, 0, st>>>((const bf*)store, vp, perm, (float*)out, rows, V, D,
                                                               n_store);
  else if (store_dt == MAGIC_F32 && out_dt == MAGIC_BF16)
    gather_views_kernel<float, bf><<<(int)blocks, 256, 0, st>>>((const float*)store, vp, perm, (bf*)out, rows, V, D,
                                                               n_store);
  else {
    magic_set_error("magic_gather_views: bad dtype pair %d -> %d", store_dt, out_dt);
    return MAGIC_ERR_ARG;
  }
  MAGIC_CHECK_LAUNCH("magic_gather_views");
  return MAGIC_OK;
}

int magic_gather_pair_dists(const float* dist, long long N, const long long* node_vp, float* out, int B, int G,
                            cudaStream_t st) {
  if (B <= 0 || G <= 0) return MAGIC_OK;
  const long long total = (long long)B * G * G;
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * magic_num_sms();
  if (blocks > cap) blocks = cap;
  pair_dists_kernel<<<(int)blocks, 256, 0, st>>>(dist, N, node_vp, out, B, G);
  MAGIC_CHECK_LAUNCH("magic_gather_pair_dists");
  return MAGIC_OK;
}

}  // extern "C"
