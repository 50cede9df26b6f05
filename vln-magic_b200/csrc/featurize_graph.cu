// GPU batch featuriser, graph half: everything of a pretraining batch that derives from the
// navigation graph is built on the device from a resident copy of the world and a few integers per sample.
//
// The reference builds it per sample in DataLoader workers with python dict / list loops
// (pretrain_src/data/dataset.py: get_cur_angle :433-443, get_traj_pano_fts :733-772, get_gmap_inputs :513-549,
// get_gmap_pos_fts :553-575, get_vp_pos_fts :577-586, get_act_labels :622-640) and collates with data/tasks.py:110-166;
// the model's string loops then run on the collated viewpoint ids.  Here:
//
//   world tables (HBM, once)   positions [N,3], all-pairs shortest distance / hop count [N,N], per-viewpoint candidate
//                              table (next viewpoint, view index, heading / elevation offsets, in the scanvp_cands dict
//                              order), the 36 view angles
//   per sample (host -> device) the path as store rows (<= Tmax ints), its length, the start heading, the next ground-
//                              truth viewpoint, the first panorama row of the sample
//
//   featurize_sample_kernel    one CTA per sample.  Warp 0 replays the ordered-dict logic of get_gmap_inputs (visited in
//                              first-visit order with the LAST step id, unvisited in insertion order with deletions, a
//                              re-inserted key moves to the end) with warp-cooperative membership tests; then all warps
//                              emit step ids, masks, position features (fp64 arcsin / sqrt like numpy, rounded to fp32),
//                              pair distances, candidate-first view orders, location features, nav types, the local
//                              position features, action labels, and the integer tables the model needs (CSR of gmap
//                              sources, SAP masks, local -> global scatter table) into a per-sample slab
//   featurize_compact_kernel   one CTA: exclusive scan of the per-sample entry counts; writes the batch-wide CSR
//                              (node_ptr / entries) and its reverse (src_ids ascending / src_ptr / src_nodes / src_w) --
//                              every source feeds exactly one node, so the reverse CSR is a permutation and the global
//                              sorted order is known from per-sample ranks: [visited rows of the last sample ... first
//                              sample | candidate tokens of the first sample ... last sample]
//
// Integer outputs are bit-exact against the reference loops; float features agree to 1-2 ulp (numpy evaluates sin /
// cos in fp32 with its own polynomial; here they are evaluated in fp64 and rounded).
#include <math.h>

#include "common.cuh"
#include "../../include/magic_b200.h"

namespace {

constexpr int FT = 128;       // threads per sample CTA
constexpr int MAXN = 256;     // ordered-set capacity per sample (visited + ever-inserted unvisited)
constexpr int V36 = 36;
constexpr double PI_D = 3.14159265358979323846;

// index of `key` among the live entries of list[0..n) (alive may be null), or -1; cooperative over one warp
__device__ __forceinline__ int warp_find(const int* list, const unsigned char* alive, int n, int key, int lane) {
  int found = -1;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    const bool hit = i < n && list[i] == key && (alive == nullptr || alive[i]);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      found = base + __ffs(m) - 1;
      break;
    }
  }
  return found;
}

struct RelPos {
  float sh, ch, se, ce, d30;
};
// pretrain_src/data/common.py:142-160 in fp64 on fp64 positions, then the fp32 rounding of dataset.py:570-573
// (rel_angles / rel_dists `.astype(np.float32)`; numpy evaluates sin / cos of the fp32 angle in fp32)
__device__ __forceinline__ RelPos rel_pos(const double* pa, const double* pb, double base_h, double base_e) {
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
  const double xy = fmax(sqrt(dx * dx + dy * dy), 1e-8), xyz = fmax(sqrt(dx * dx + dy * dy + dz * dz), 1e-8);
  double h = asin(dx / xy);
  if (pb[1] < pa[1]) h = PI_D - h;
  h -= base_h;
  const double e = asin(dz / xyz) - base_e;
  const float hf = (float)h, ef = (float)e;
  RelPos r;
  r.sh = (float)sin((double)hf); r.ch = (float)cos((double)hf);
  r.se = (float)sin((double)ef); r.ce = (float)cos((double)ef);
  r.d30 = (float)(xyz / 30.0);
  return r;
}
// float32(d / MAX_DIST), float32(steps / MAX_STEP) as numpy rounds them (division in fp64)
__device__ __forceinline__ float norm_dist(float d) { return (float)((double)d / 30.0); }
__device__ __forceinline__ float norm_hops(int n) { return (float)((double)n / 10.0); }

__global__ void __launch_bounds__(FT) featurize_sample_kernel(const MagicFeatArgs A) {
  __shared__ int s_vis[MAXN], s_step[MAXN], s_unv[MAXN], s_node[MAXN], s_last_t[MAXN], s_cnt[MAXN], s_ptr[MAXN + 1];
  __shared__ unsigned char s_alive[MAXN], s_visited[MAXN];
  __shared__ int s_nv, s_nu, s_g, s_err;
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int T = A.path_len[b], G = A.G, C = A.C, Vp = A.Vp, Tmax = A.Tmax;
  const int* path = A.path + (size_t)b * Tmax;
  const int row0 = A.row0[b];
  if (threadIdx.x == 0) s_err = 0;
  __syncthreads();

  // ---- ordered-dict replay (warp 0) ------------------------------------------------------------------
  if (w == 0) {
    int nv = 0, nu = 0;
    for (int t = 0; t < T; t++) {
      const int vp = path[t];
      int k = warp_find(s_vis, nullptr, min(nv, MAXN), vp, lane);
      if (k >= 0) {
        if (lane == 0) { s_step[k] = t + 1; s_last_t[k] = t; }
      } else {
        if (nv < MAXN && lane == 0) { s_vis[nv] = vp; s_step[nv] = t + 1; s_last_t[nv] = t; }
        nv++;
      }
      __syncwarp();
      k = warp_find(s_unv, s_alive, min(nu, MAXN), vp, lane);
      if (k >= 0 && lane == 0) s_alive[k] = 0;  // `del unvisited_vpids[vp]`
      __syncwarp();
      const int nc = A.n_cand[vp];
      for (int j = 0; j < nc; j++) {
        const int nx = A.cand_vp[(size_t)vp * C + j];
        if (warp_find(s_vis, nullptr, min(nv, MAXN), nx, lane) >= 0) continue;
        if (warp_find(s_unv, s_alive, min(nu, MAXN), nx, lane) >= 0) continue;  // re-assigning a live key keeps its position
        if (nu < MAXN && lane == 0) { s_unv[nu] = nx; s_alive[nu] = 1; }
        nu++;
        __syncwarp();
      }
      if (nv > MAXN || nu > MAXN) break;
    }
    if (lane == 0) {
      int g = 1;
      s_node[0] = -1; s_visited[0] = 0;
      if (nv > MAXN || nu > MAXN) { s_err = 1; nv = min(nv, MAXN); nu = min(nu, MAXN); }
      for (int k = 0; k < nv && g < MAXN; k++, g++) { s_node[g] = s_vis[k]; s_visited[g] = 1; }
      const int gv = g;
      for (int k = 0; k < nu && g < MAXN; k++)
        if (s_alive[k]) { s_node[g] = s_unv[k]; s_visited[g] = 0; s_step[g - 1] = 0; g++; }
      // (s_step is indexed by node - 1: visited nodes keep their step id, unvisited ones get 0)
      (void)gv;
      if (g > G) { s_err = 2; g = G; }
      s_nv = nv; s_nu = nu; s_g = g;
    }
  }
  __syncthreads();
  const int g_len = s_g, nv = s_nv;
  if (threadIdx.x == 0 && s_err) atomicMax(A.status, s_err);

  // ---- current heading / elevation (get_cur_angle) ---------------------------------------------------
  const int cur = path[T - 1];
  double cur_h = (double)A.start_heading[b], cur_e = 0.0;
  const int prev_given = A.prev_vp != nullptr ? A.prev_vp[b] : -1;
  if (T >= 2 || prev_given >= 0) {
    const int prev = prev_given >= 0 ? prev_given : path[T - 2];
    int view = 0;
    for (int j = 0; j < A.n_cand[prev]; j++)
      if (A.cand_vp[(size_t)prev * C + j] == cur) view = A.cand_view[(size_t)prev * C + j];  // dict: the last key wins
    cur_h = (double)(view % 12) * (30.0 * PI_D / 180.0);
    cur_e = (double)(view / 12 - 1) * (30.0 * PI_D / 180.0);
  }
  const double* pcur = A.pos + (size_t)cur * 3;

  // ---- global map tensors ----------------------------------------------------------------------------
  for (int n = threadIdx.x; n < G; n += FT) {
    const size_t o = (size_t)b * G + n;
    const bool live = n < g_len;
    const int vp = live ? s_node[n] : -1;
    A.gmap_node_vp[o] = vp;
    A.gmap_step_ids[o] = (live && n > 0) ? s_step[n - 1] : 0;
    A.gmap_visited_masks[o] = live ? s_visited[n] : 0;
    float* pf = A.gmap_pos_fts + o * 7;
    if (vp >= 0) {
      const RelPos r = rel_pos(pcur, A.pos + (size_t)vp * 3, cur_h, cur_e);
      pf[0] = r.sh; pf[1] = r.ch; pf[2] = r.se; pf[3] = r.ce;
      pf[4] = r.d30;
      pf[5] = norm_dist(A.dist[(size_t)cur * A.N + vp]);
      pf[6] = norm_hops(A.hops[(size_t)cur * A.N + vp]);
    } else if (live) {  // [stop]: zero angles -> (sin, cos) = (0, 1), zero distances
      pf[0] = 0.f; pf[1] = 1.f; pf[2] = 0.f; pf[3] = 1.f; pf[4] = pf[5] = pf[6] = 0.f;
    } else {
      for (int k = 0; k < 7; k++) pf[k] = 0.f;
    }
  }
  if (threadIdx.x == 0) A.gmap_lens[b] = g_len;
  for (int e = threadIdx.x; e < G * G; e += FT) {
    const int i = e / G, j = e - i * G;
    float d = 0.f;
    // dataset.py:545-549 reads shortest_distances[vp_i][vp_j] for i < j and mirrors it
    if (i >= 1 && j >= 1 && i != j && i < g_len && j < g_len) d = A.dist[(size_t)s_node[min(i, j)] * A.N + s_node[max(i, j)]];
    A.gmap_pair_dists[(size_t)b * G * G + e] = d;
  }

  // ---- panoramas of the trajectory: candidate-first view order, location features, nav types ---------
  for (int t = w; t < T; t += FT / 32) {  // one warp per step
    const int vp = path[t], nc = A.n_cand[vp];
    const size_t row = (size_t)row0 + t;
    unsigned long long used = 0ull;
    for (int j = 0; j < nc; j++) used |= 1ull << A.cand_view[(size_t)vp * C + j];
    const int n_used = __popcll(used);
    if (nc + (V36 - n_used) > V36 && lane == 0) atomicMax(A.status, 3);  // two candidates share a view: > 36 tokens
    if (lane == 0) {
      A.traj_vp_index[row] = vp;
      A.traj_vp_view_lens[row] = min(V36, nc + V36 - n_used);
    }
    for (int j = lane; j < V36; j += 32) {
      int view;
      double h, e;
      long long nav;
      if (j < nc) {
        view = A.cand_view[(size_t)vp * C + j];
        const double vh = (double)A.view_ang[view * 2], ve = (double)A.view_ang[view * 2 + 1];
        const double oh = (double)A.cand_ang[((size_t)vp * C + j) * 2], oe = (double)A.cand_ang[((size_t)vp * C + j) * 2 + 1];
        if (A.correct_heading) { h = cur_h - vh + oh; e = cur_e - ve + oe; }
        else { h = vh + oh; e = ve + oe; }
        nav = 1;
      } else {
        // the (j - nc)-th view index that no candidate uses, ascending
        int k = j - nc;
        view = -1;
        for (int v = 0; v < V36; v++)
          if (!((used >> v) & 1ull)) {
            if (k == 0) { view = v; break; }
            k--;
          }
        h = view >= 0 ? (double)A.view_ang[view * 2] : 0.0;
        e = view >= 0 ? (double)A.view_ang[view * 2 + 1] : 0.0;
        nav = 0;
      }
      A.traj_view_perm[row * V36 + j] = view;
      A.traj_nav_types[row * V36 + j] = view >= 0 ? nav : 0;
      float* lf = A.traj_loc_fts + (row * V36 + j) * 7;
      if (view >= 0) {
        const float hf = (float)h, ef = (float)e;  // np.stack(view_angles) is float32 (all_point_rel_angles dtype)
        lf[0] = (float)sin((double)hf); lf[1] = (float)cos((double)hf);
        lf[2] = (float)sin((double)ef); lf[3] = (float)cos((double)ef);
        lf[4] = lf[5] = lf[6] = 1.f;
      } else {
        for (int k = 0; k < 7; k++) lf[k] = 0.f;
      }
    }
  }

  // ---- local branch: position features of [stop] + the last panorama's tokens -------------------------
  const int nc_last = A.n_cand[cur];
  {
    const RelPos rs = rel_pos(pcur, A.pos + (size_t)path[0] * 3, cur_h, cur_e);
    const float sd = norm_dist(A.dist[(size_t)cur * A.N + path[0]]), sh = norm_hops(A.hops[(size_t)cur * A.N + path[0]]);
    for (int j = threadIdx.x; j < Vp; j += FT) {
      float* pf = A.vp_pos_fts + ((size_t)b * Vp + j) * 14;
      pf[0] = rs.sh; pf[1] = rs.ch; pf[2] = rs.se; pf[3] = rs.ce; pf[4] = rs.d30; pf[5] = sd; pf[6] = sh;
      if (j >= 1 && j - 1 < nc_last) {
        const int c = A.cand_vp[(size_t)cur * C + (j - 1)];
        const RelPos r = rel_pos(pcur, A.pos + (size_t)c * 3, cur_h, cur_e);
        pf[7] = r.sh; pf[8] = r.ch; pf[9] = r.se; pf[10] = r.ce; pf[11] = r.d30;
        pf[12] = norm_dist(A.dist[(size_t)cur * A.N + c]);
        pf[13] = norm_hops(A.hops[(size_t)cur * A.N + c]);
      } else {
        for (int k = 7; k < 14; k++) pf[k] = 0.f;
      }
    }
  }

  // ---- action labels (R2R get_act_labels) --------------------------------------------------------------
  if (threadIdx.x == 0 && A.next_vp != nullptr) {
    const int nx = A.next_vp[b];
    long long gl = -100, ll = -100;
    if (nx == -1) gl = ll = 0;  // stop
    else if (nx >= 0) {
      for (int n = 1; n < g_len; n++)
        if (s_node[n] == nx) { gl = n; break; }
      for (int j = 0; j < nc_last; j++)
        if (A.cand_vp[(size_t)cur * C + j] == nx) { ll = j + 1; break; }
    }
    A.global_act_labels[b] = gl;
    A.local_act_labels[b] = ll;
  }

  // ---- integer tables of the model (graph_index.build_index) ------------------------------------------
  // SAP masks and the local -> global scatter table
  for (int n = threadIdx.x; n < G; n += FT) {
    const size_t o = (size_t)b * G + n;
    const bool live = n < g_len;
    A.g_valid[o] = (live && !s_visited[n]) ? 1 : 0;
    int c2 = -1;
    if (live && n > 0 && !s_visited[n]) {
      const int vp = s_node[n];
      for (int j = 0; j < nc_last && j + 1 < Vp; j++)
        if (A.cand_vp[(size_t)cur * C + j] == vp) c2 = j + 1;  // the last candidate with this id wins
    }
    A.node2cand[o] = c2;
  }
  unsigned long long used_last = 0ull;
  for (int j = 0; j < nc_last; j++) used_last |= 1ull << A.cand_view[(size_t)cur * C + j];
  const int vp_len = min(V36, nc_last + V36 - __popcll(used_last)) + 1;  // tokens of the last panorama + [stop]
  for (int j = threadIdx.x; j < Vp; j += FT) {
    const size_t o = (size_t)b * Vp + j;
    unsigned char lv = 0, bw = 0;
    long long src = -1;
    if (j == 0) lv = 1;
    else {
      if (j - 1 < V36) src = ((long long)row0 + T - 1) * V36 + (j - 1);
      if (j < vp_len && j - 1 < nc_last) {
        lv = 1;
        const int c = A.cand_vp[(size_t)cur * C + (j - 1)];
        for (int k = 0; k < nv; k++)
          if (s_vis[k] == c) bw = 1;  // leads back to a visited node
      }
    }
    A.l_valid[o] = lv;
    A.bw_mask[o] = bw;
    A.vp_gather[o] = src;
  }
  if (threadIdx.x == 0) {
    A.key_lens_gmap[b] = g_len;
    A.key_lens_vp[b] = vp_len;
    A.last_rows[b] = (long long)row0 + T - 1;
  }

  // gmap sources (CSR per node, local): visited node -> -(row of its LAST visit + 1); unvisited node -> every
  // (step, candidate slot) that names it, in (step, slot) order
  for (int n = threadIdx.x; n < G; n += FT) {
    int cnt = 0;
    if (n >= 1 && n < g_len) {
      if (s_visited[n]) cnt = 1;
      else {
        const int vp = s_node[n];
        for (int t = 0; t < T; t++) {
          const int pv = path[t];
          for (int j = 0; j < A.n_cand[pv]; j++) cnt += A.cand_vp[(size_t)pv * C + j] == vp;
        }
      }
    }
    s_cnt[n] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int n = 0; n < G; n++) { s_ptr[n] = acc; acc += s_cnt[n]; }
    s_ptr[G] = acc;
    if (acc > A.E_s) { atomicMax(A.status, 4); }
    A.slab_total[b] = min(acc, A.E_s);
    A.slab_nvis[b] = nv;
  }
  __syncthreads();
  int* slab_e = A.slab_entries + (size_t)b * A.E_s;
  int* slab_n = A.slab_nodes + (size_t)b * A.E_s;
  for (int n = threadIdx.x; n < G; n += FT) {
    A.slab_ptr[(size_t)b * (G + 1) + n] = s_ptr[n];
    if (n == 0) A.slab_ptr[(size_t)b * (G + 1) + G] = s_ptr[G];
    if (n < 1 || n >= g_len) continue;
    int o = s_ptr[n];
    if (s_visited[n]) {
      if (o < A.E_s) { slab_e[o] = -(row0 + s_last_t[n - 1] + 1); slab_n[o] = n; }
    } else {
      const int vp = s_node[n];
      for (int t = 0; t < T; t++) {
        const int pv = path[t];
        for (int j = 0; j < A.n_cand[pv]; j++)
          if (A.cand_vp[(size_t)pv * C + j] == vp) {
            if (o < A.E_s) { slab_e[o] = (row0 + t) * V36 + j; slab_n[o] = n; }
            o++;
          }
      }
    }
  }
  __syncthreads();
  // rank of every local entry in ascending order (rank sort: <= E_s elements, all distinct)
  const int tot = min(s_ptr[G], A.E_s);
  for (int i = threadIdx.x; i < tot; i += FT) {
    const int v = slab_e[i];
    int r = 0;
    for (int k = 0; k < tot; k++) r += slab_e[k] < v;
    A.slab_rank[(size_t)b * A.E_s + i] = r;
  }
}

// one CTA: batch-wide CSR and reverse CSR from the per-sample slabs
__global__ void __launch_bounds__(1024) featurize_compact_kernel(const MagicFeatArgs A) {
  extern __shared__ int sm[];  // [B] entry offset, [B] visited-offset (from the end), [B] token offset
  int* off = sm;
  int* neg0 = sm + A.B;
  int* tok0 = sm + 2 * A.B;
  __shared__ int s_total, s_negs;
  const int B = A.B, G = A.G;
  if (threadIdx.x == 0) {
    int acc = 0, negs = 0;
    for (int b = 0; b < B; b++) { off[b] = acc; acc += A.slab_total[b]; negs += min(A.slab_nvis[b], A.slab_total[b]); }
    s_total = acc; s_negs = negs;
    // ascending order of the signed ids: visited rows of the LAST sample first (most negative), tokens after them
    int a = 0;
    for (int b = B - 1; b >= 0; b--) { neg0[b] = a; a += min(A.slab_nvis[b], A.slab_total[b]); }
    int t = negs;
    for (int b = 0; b < B; b++) { tok0[b] = t; t += A.slab_total[b] - min(A.slab_nvis[b], A.slab_total[b]); }
    if (acc > A.E_cap || acc > A.S_cap) atomicMax(A.status, 5);
    A.n_src[0] = acc;
  }
  __syncthreads();
  const int total = min(s_total, min(A.E_cap, A.S_cap));
  for (int i = threadIdx.x; i < B * G; i += blockDim.x) {
    const int b = i / G, n = i - b * G;
    A.node_ptr[i] = off[b] + A.slab_ptr[(size_t)b * (G + 1) + n];
  }
  if (threadIdx.x == 0) A.node_ptr[B * G] = s_total;
  for (int idx = threadIdx.x; idx < B * A.E_s; idx += blockDim.x) {
    const int b = idx / A.E_s, i = idx - b * A.E_s;
    const int tot = A.slab_total[b], nneg = min(A.slab_nvis[b], tot);
    if (i >= tot) continue;
    const int e = A.slab_entries[idx];
    const int n = A.slab_nodes[idx];
    const int r = A.slab_rank[idx];  // local rank: negatives (more negative = later row) first
    if (off[b] + i < A.E_cap) A.entries[off[b] + i] = e;
    const int pos = r < nneg ? neg0[b] + r : tok0[b] + (r - nneg);
    if (pos < total) {
      A.src_ids[pos] = e;
      A.src_nodes[pos] = b * G + n;
      const int cnt = A.slab_ptr[(size_t)b * (G + 1) + n + 1] - A.slab_ptr[(size_t)b * (G + 1) + n];
      A.src_w[pos] = 1.f / (float)cnt;
    }
  }
  // panorama rows beyond the real ones (capacity padding): one zero view, referenced by no table
  for (long long i = (long long)A.R * V36 + threadIdx.x; i < (long long)A.R_cap * V36; i += blockDim.x) {
    A.traj_view_perm[i] = -1;
    A.traj_nav_types[i] = 0;
    for (int k = 0; k < 7; k++) A.traj_loc_fts[i * 7 + k] = 0.f;
  }
  for (int r = A.R + threadIdx.x; r < A.R_cap; r += blockDim.x) {
    A.traj_vp_index[r] = 0;
    A.traj_vp_view_lens[r] = 1;
  }
  // capacity padding, as graph_index.pad_batch lays it out: entries / src_nodes / src_w / src_ids = 0, src_ptr = total
  for (int i = threadIdx.x; i < A.E_cap; i += blockDim.x)
    if (i >= total) { A.entries[i] = 0; A.src_nodes[i] = 0; A.src_w[i] = 0.f; }
  for (int i = threadIdx.x; i <= A.S_cap; i += blockDim.x) {
    A.src_ptr[i] = min(i, total);
    if (i < A.S_cap && i >= total) A.src_ids[i] = 0;
  }
}

}  // namespace

extern "C" int magic_feat_args_size(void) { return (int)sizeof(MagicFeatArgs); }

extern "C" int magic_featurize_graph(const MagicFeatArgs* args, cudaStream_t st) {
  const MagicFeatArgs& A = *args;
  MAGIC_CHECK_ARG(A.B > 0 && A.G > 0 && A.G <= 256 && A.C > 0 && A.Tmax > 0 && A.Vp > 1 && A.E_s > 0,
                  "magic_featurize_graph: bad sizes B=%d G=%d C=%d Tmax=%d Vp=%d E_s=%d", A.B, A.G, A.C, A.Tmax, A.Vp, A.E_s);
  MAGIC_CHECK_ARG(A.B <= 4096, "magic_featurize_graph: B=%d > 4096", A.B);
  featurize_sample_kernel<<<A.B, FT, 0, st>>>(A);
  MAGIC_CHECK_LAUNCH("magic_featurize_graph(sample)");
  featurize_compact_kernel<<<1, 1024, 3 * A.B * sizeof(int), st>>>(A);
  MAGIC_CHECK_LAUNCH("magic_featurize_graph(compact)");
  return MAGIC_OK;
}
