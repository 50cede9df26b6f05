/* libmagic_b200 -- C ABI of the B200-native MAGIC pretraining / distillation hot path.
 *
 * The reference (CrystalSixone/VLN-MAGIC) is pure PyTorch on this path and ships no FFI; its model file
 * is absent (readme.md:75).  The entry points below are the operators its missing
 * `model/pretrain_goat.py::GlocalTextPathCMTPreTraining.forward(batch, task, compute_loss)`
 * (call sites pretrain_src/train_r2r_magic.py:448,483,510-512,545-546) decomposes into, plus the KD
 * primitives of pretrain_src/optim/kd_loss.py and the optimizer of pretrain_src/optim/adamw.py.
 * Each declaration cites the reference interface it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: raw DEVICE pointers, explicit sizes, an explicit cudaStream_t, no internal
 * synchronisation, `int` status return (0 = ok) with the message in magic_last_error().
 * dtype: MAGIC_F32 = 0, MAGIC_BF16 = 1 (storage type of activations; all math is fp32).
 * Dropout: p = 0 disables; `seed_ptr` is a DEVICE pointer to a 64-bit seed (CUDA-graph friendly),
 * `salt` distinguishes call sites; backward regenerates the forward mask from (seed, salt, index).
 */
#ifndef MAGIC_B200_H_
#define MAGIC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define MAGIC_F32 0
#define MAGIC_BF16 1
#define MAGIC_ACT_NONE 0
#define MAGIC_ACT_GELU 1
#define MAGIC_ACT_RELU 2
#define MAGIC_MAKD_MAX_SEGS 32
#define MAGIC_SUMSQ_SCRATCH 2048

const char* magic_last_error(void);
int magic_version(void);
/* programmatic dependent launch for subsequent launches: 1 on, 0 off, -1 follow env MAGIC_PDL (default on) */
int magic_set_pdl(int on);
/* 1 if the tcgen05/TMA GEMM path is compiled in and usable for (M,N,K) bf16 operands */
int magic_gemm_tc_supported(int M, int N, int K);

/* ---- dense contractions: nn.Linear forward / dgrad / wgrad of every encoder layer ------------------
 * C[m,n] = epi(alpha * sum_k A(m,k) B(k,n) + bias[n]) + residual[m,n] + beta*C[m,n]
 * A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].  act/pre_out/dact_pre/dropout: see gemm_epi.cuh.
 * Replaces torch F.linear + F.gelu + dropout + residual in BertLayer / BertCrossLayer /
 * TransformerEncoderLayer (names pinned at train_r2r_magic.py:189-208).
 * bf16 operands with 16-byte aligned, unit-stride layouts run on tcgen05 tensor cores (TMA-fed, TMEM
 * accumulators); everything else runs the fp32 FFMA kernel. */
int magic_gemm(const void* A, int a_dt, long sam, long sak, const void* B, int b_dt, long sbk, long sbn, void* C,
               int c_dt, long ldc, int M, int N, int K, const float* bias, int act, void* pre_out,
               const void* dact_pre, int dact_dt, long dact_ld, const void* residual, long res_ld, float alpha,
               float beta, float drop_p, unsigned salt, const unsigned long long* seed_ptr, int allow_tc,
               cudaStream_t st);

/* Weight and bias gradient of y = x W^T + b in one launch (the backward of every nn.Linear named at
 * train_r2r_magic.py:189-208; replaces autograd's addmm-backward + sum kernels):
 *   dw[n*dw_ld + k] = beta*dw + sum_m dy[m*dy_ld + n] * x[m*x_ld + k]    (fp32)
 *   dbias[n]       += sum_m dy[m*dy_ld + n]                               (fp32, may be NULL)
 * M = tokens (the reduction), N = out features, K = in features. */
int magic_gemm_wgrad(const void* dy, int dy_dt, long dy_ld, const void* x, int x_dt, long x_ld, float* dw, long dw_ld,
                     float* dbias, int M, int N, int K, float beta, int allow_tc, cudaStream_t st);

/* ---- fused attention with graph-distance bias (GlobalMapEncoder: sprel_linear(gmap_pair_dists)) ----
 * q/k/v: rows are tokens, head hd occupies columns [hd*64, hd*64+64) of a row; *_ld = row stride.
 * out [B*Lq, H*64]; lse [B,H,Lq]; pbar (optional) = head-mean probabilities at
 * pbar[b*pbar_bs + i*pbar_rs + j] (the KD attention map of agent.py:579,628,654,671). */
int magic_attn_fwd(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, void* out,
                   float* lse, float* pbar, long pbar_bs, long pbar_rs, int B, int H, int Lq, int Lk,
                   const int* key_lens, const float* dists, const float* sprel_w, const float* sprel_b, float scale,
                   int dtype, float drop_p, unsigned salt, const unsigned long long* seed_ptr, cudaStream_t st);
int magic_attn_bwd(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, const void* dout,
                   const float* lse, const float* dpbar, long pbar_bs, long pbar_rs, float* delta, void* dq,
                   void* dk, void* dv, long dq_ld, long dk_ld, long dv_ld, float* dsprel, int B, int H, int Lq,
                   int Lk, const int* key_lens, const float* dists, const float* sprel_w, const float* sprel_b,
                   float scale, int dtype, float drop_p, unsigned salt, const unsigned long long* seed_ptr,
                   cudaStream_t st);
/* Call-site state for the attention launches that follow (forward and backward): key `key_index` of every sequence is
 * never attended (probability exactly 0, zero dK / dV), on top of the key-length mask; -1 switches it off.  The
 * navigation graph's [MEM] slot sits INSIDE the valid prefix but is not a key (map_nav_src/r2r/agent.py:228,
 * `batch_gmap_masks[:,1] = False`). */
int magic_attn_set_key_skip(int key_index);
/* One half of the bf16 tensor-core attention backward (run the halves on two streams): part 1 = query-major kernel
 * (delta, dQ, d sprel), part 2 = key-major kernel (dK, dV) with delta = dO . O from the forward output `out`.  Without a
 * KD-map gradient only.  Returns MAGIC_ERR_UNSUPPORTED (nothing launched) when not covered: use magic_attn_bwd then. */
int magic_attn_bwd_part(const void* q, const void* k, const void* v, long q_ld, long k_ld, long v_ld, const void* dout,
                        const void* out, const float* lse, float* delta, void* dq, void* dk, void* dv, long dq_ld,
                        long dk_ld, long dv_ld, float* dsprel, int B, int H, int Lq, int Lk, const int* key_lens,
                        const float* dists, const float* sprel_w, const float* sprel_b, float scale, int dtype,
                        float drop_p, unsigned salt, const unsigned long long* seed_ptr, int part, cudaStream_t st);

/* ---- LayerNorm (+residual, +dropout): BertSelfOutput / BertOutput / norm1,norm2 / head LNs ---------
 * y = drop_out(LN(drop_in(x) + res) * gamma + beta); stats[r] = (mean, rstd). Parameter grads ACCUMULATE. */
int magic_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y, float* stats,
                 int M, int h, float eps, int dtype, float p_in, unsigned salt_in, float p_out, unsigned salt_out,
                 const unsigned long long* seed_ptr, cudaStream_t st);
int magic_ln_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* stats, void* dx,
                 void* dres, float* dgamma, float* dbeta, int M, int h, int dtype, float p_in, unsigned salt_in,
                 float p_out, unsigned salt_out, const unsigned long long* seed_ptr, cudaStream_t st);

/* ---- text embeddings (bert.embeddings.*: word + position + token_type -> LayerNorm -> dropout) ----- */
int magic_embed_ln_fwd(const long long* ids, const float* word, const float* pos, const float* type0,
                       const float* gamma, const float* beta, void* y, float* stats, int M, int L, int h, float eps,
                       int dtype, float p_out, unsigned salt_out, const unsigned long long* seed_ptr,
                       cudaStream_t st);
int magic_embed_ln_bwd(const void* dy, const long long* ids, const float* word, const float* pos,
                       const float* type0, const float* gamma, const float* stats, float* dword, float* dpos,
                       float* dtype0, float* dgamma, float* dbeta, int M, int L, int h, int dtype, float p_out,
                       unsigned salt_out, const unsigned long long* seed_ptr, cudaStream_t st);

/* ---- positional fusion: y = xin + emb[idx] + cst + LN(W f + b)  (K <= 16) --------------------------
 * loc_linear+loc_layer_norm+nav_type_embedding (img_embeddings), gmap_pos_embeddings+gmap_step_embeddings,
 * vp_pos_embeddings.  backward: d(xin) == dy (not written); parameter grads ACCUMULATE. */
int magic_posfuse_fwd(const void* xin, const long long* idx, const float* emb, const float* cst, const float* f,
                      const float* W, const float* b, const float* gamma, const float* beta, void* y, float* stats,
                      int M, int h, int K, float eps, int dtype, cudaStream_t st);
int magic_posfuse_bwd(const void* dy, const long long* idx, const float* f, const float* W, const float* b,
                      const float* gamma, const float* stats, float* demb, float* dcst, float* dW, float* db,
                      float* dgamma, float* dbeta, int M, int h, int K, int dtype, cudaStream_t st);

/* ---- row gather (idx < 0 -> zero row) / its adjoint scatter (unique idx; dsrc zero-filled first) ----
 * MLM masked-token gather (train_r2r_magic.py:450-452 order) and the [stop] + last-step view assembly
 * of the local branch (data/dataset.py:581-584). */
int magic_gather_rows(const void* src, const long long* idx, void* out, int R, int h, int dtype, cudaStream_t st);
int magic_scatter_rows(const void* dout, const long long* idx, void* dsrc, int R, int n_src_rows, int h, int dtype,
                       cudaStream_t st);

/* ---- adaptive panorama pooling (adaptive_pano_fusion, r2r_magic_model_config.json:57) -------------- */
int magic_pano_fuse_fwd(const void* x, const float* w, const float* bias, const long long* lens, void* fused,
                        float* probs, int R, int V, int h, int dtype, cudaStream_t st);
int magic_pano_fuse_bwd(const void* dfused, const void* x, const float* w, const long long* lens, const float* probs,
                        void* dx, float* dw, float* dbias, int R, int V, int h, int dtype, cudaStream_t st);

/* ---- N = 1 heads (ClsPrediction.net.3) and bias gradients ----------------------------------------- */
int magic_rowdot_fwd(const void* x, const float* w, const float* bias, float* y, int M, int h, int dtype,
                     cudaStream_t st);
int magic_rowdot_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int M, int h,
                     int dtype, cudaStream_t st);
int magic_colsum(const void* x, float* out, int M, int N, long ld, int dtype, cudaStream_t st); /* out += */
int magic_cast(const void* in, int in_dt, void* out, int out_dt, long long n, cudaStream_t st);
/* glue: out = a + b (+ c); strided 2-D copy; segment sums (per-sample means of masked-token losses) */
int magic_add(const void* a, const void* b, const void* c, void* out, long long n, float scale, int dtype,
              cudaStream_t st); /* out = scale * (a + b (+ c)) */
int magic_copy2d(const void* src, long src_ld, void* dst, long dst_ld, int rows, int cols, int dtype,
                 cudaStream_t st);
int magic_segsum(const float* vals, const long long* seg, const float* seg_scale, float* out, int R, int n_seg,
                 cudaStream_t st);
/* MKTD sample weights: exponential_decay (kd_loss.py:43-44), invert_normalized_losses (kd_loss.py:46-54) */
int magic_exp_decay(const float* in, float* out, int n, float rate, cudaStream_t st);
int magic_invert_norm(const float* in, float* out, int n, cudaStream_t st);
/* per-row KD weights: out[i] = (src ? src[idx ? idx[i] : i] : 1) * (scale ? scale[i] : 1); idx < 0 -> 0.  The
   gather is the per-row form of t_sample_weights (kd_loss.py:31-40, agent.py:1019); `scale` masks padded rows */
int magic_row_weights(const float* src, const long long* idx, const float* scale, float* out, int n,
                      cudaStream_t st);
/* out[i] = valid[i] ? x[i] : fill -- object-grounding logits (`obj_logits.masked_fill_(obj_masks.logical_not(), -inf)`,
   the OG head of the DUET lineage; batch schema data/tasks.py:503-559); its backward uses fill = 0 */
int magic_mask_fill(const float* x, const unsigned char* valid, float* out, int n, float fill, cudaStream_t st);
/* dz = dy * dropscale * act'(pre): backward of a stand-alone Linear+activation (ClsPrediction, MLM transform) */
int magic_act_bwd(const void* dy, const void* pre, void* dz, long long n, int act, int dtype, float drop_p,
                  unsigned salt, const unsigned long long* seed_ptr, cudaStream_t st);

/* ---- gmap node features (DUET-lineage _aggregate_gmap_features) -------------------- */
int magic_gmap_aggregate_fwd(const void* tokens, const void* fused, const int* node_ptr, const int* entries,
                             void* out, int n_nodes, int h, int dtype, cudaStream_t st);
int magic_gmap_aggregate_bwd(const void* dout, const int* src_ids, const int* src_ptr, const int* src_nodes,
                             const float* src_w, void* dtokens, long long n_token_rows, void* dfused,
                             long long n_fused_rows, int n_src, int h, int dtype, cudaStream_t st);

/* ---- SAP logits: gate, -inf masks, local->global scatter (outputs pinned train_r2r_magic.py:510-518) */
int magic_sap_fuse_fwd(const float* g_raw, const float* l_raw, const float* gate_raw, const unsigned char* g_valid,
                       const unsigned char* l_valid, const int* node2cand, const unsigned char* bw_mask, float* gl,
                       float* ll, float* fl, int B, int G, int Vp, cudaStream_t st);
int magic_sap_fuse_bwd(const float* dgl, const float* dll, const float* dfl, const float* g_raw, const float* l_raw,
                       const float* gate_raw, const unsigned char* g_valid, const unsigned char* l_valid,
                       const int* node2cand, const unsigned char* bw_mask, float* dg_raw, float* dl_raw,
                       float* dgate_raw, int B, int G, int Vp, cudaStream_t st);

/* ---- cross-entropy rows (F.cross_entropy(reduction='none', ignore_index)) ------------------------- */
int magic_ce_fwd(const void* logits, const long long* labels, float* loss, float* lse, int R, int C, long ld,
                 long long ignore_index, int dtype, cudaStream_t st);
int magic_ce_bwd(const void* logits, const long long* labels, const float* lse, const float* dloss, void* dlogits,
                 int R, int C, long ld, long long ignore_index, int dtype, cudaStream_t st);

/* ---- MRC / CFP task heads (outputs pinned at pretrain_src/train_r2r_magic.py:483-488 and :545-560) ----
 * zero_rows: x[rows[i], :] = 0 in place (masked last-step views, data/tasks.py:178-181; rows[i] < 0 skipped).
 * soft_ce: loss[r] = KL(targets[r] || softmax(logits[r])) summed over C classes (F.kl_div 'none' .sum(1));
 *          stats[r] = (logsumexp, sum of targets) saved for backward.
 * l2norm:  y = x / max(||x||_2, eps) per row (F.normalize); inv_norm[r] saved for backward. */
int magic_zero_rows(void* x, const long long* rows, int n, int h, int dtype, cudaStream_t st);
int magic_soft_ce_fwd(const void* logits, const float* targets, float* loss, float* stats /* [R,2] */, int R, int C,
                      long ld, long tld, int dtype, cudaStream_t st);
int magic_soft_ce_bwd(const void* logits, const float* targets, const float* stats, const float* dloss,
                      void* dlogits, int R, int C, long ld, long tld, int dtype, cudaStream_t st);
int magic_l2norm_fwd(const void* x, void* y, float* inv_norm, int R, int h, float eps, int dtype, cudaStream_t st);
int magic_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int R, int h, int dtype,
                     cudaStream_t st);

/* ---- MAKD losses (pretrain_src/optim/kd_loss.py:5-41; aggregation map_nav_src/r2r/agent.py:546-719) */
typedef struct MagicMseSeg {
  const void* s;      /* student (already projected to the teacher width) */
  const void* t;      /* teacher */
  void* ds;           /* backward: gradient wrt s, same layout as s (may be NULL in forward) */
  const float* w;     /* per-row MKTD weights (kd_loss.py:11-13) or NULL */
  const float* scale_dev; /* optional DEVICE scalar multiplied into `scale` (device-resident MKRW weight, so a
                             captured CUDA graph sees the weights drawn for the current step) or NULL */
  long long rows, inner, s_rs, t_rs; /* rows x inner elements, row strides in elements */
  float scale;        /* MKRW weight / (rows*inner)  (mean reduction, kd_loss.py:8,14) */
  int s_dt, t_dt;
  int vec_ok;         /* filled by the library */
} MagicMseSeg;
/* loss[0..nseg) per-segment, loss[MAGIC_MAKD_MAX_SEGS] = sum of all segments (buffer of MAX_SEGS+1, zeroed here) */
int magic_makd_mse_fwd(const MagicMseSeg* segs, int nseg, float* loss, cudaStream_t st);
/* upstream gradient of segment i = gseg[i] (nullable) + gtot[0] (nullable) */
int magic_makd_mse_bwd(const MagicMseSeg* segs, int nseg, const float* gseg, const float* gtot, cudaStream_t st);
/* total = alpha*(mse_total + kl) + (1-alpha)*mean(sup[0..n))  (agent.py:1119); out = [total, sup_mean, kd].
 * inv_n (nullable, device): 1/(number of real rows) when sup is padded with zero rows for CUDA-graph replay */
int magic_loss_mix_fwd(const float* mse_total, const float* kl, const float* sup, int n, float alpha,
                       const float* inv_n, float* out, cudaStream_t st);
int magic_loss_mix_bwd(const float* g, int n, float alpha, const float* inv_n, float* d_mse, float* d_kl,
                       float* d_sup, cudaStream_t st);
/* measurement aid: when dev_buf (32 x u64, device) is non-NULL, CTA 0 of every tensor-core GEMM launched afterwards
 * stamps clock64() at its phase boundaries into it (scripts/gemm_trace.py); NULL switches tracing off */
int magic_gemm_set_trace(unsigned long long* dev_buf);
/* scheduling hint: SMs the persistent tcgen05 GEMM may occupy from now on (0 = all).  The stepper lowers it while it
 * captures a graph that runs beside another one, so both make progress instead of queueing SM by SM. */
int magic_gemm_set_sm_budget(int sms);
/* measurement aid: keeps the stream busy for ~cycles SM clocks so the host can queue launches ahead of the GPU */
int magic_delay(long long cycles, cudaStream_t st);
/* measurement aid (no reference counterpart): events recordable inside a captured CUDA graph (external event-record
 * nodes), so bench.py times every kernel of the REPLAYED step with CUDA events on the launching stream */
int magic_event_create(void** ev);
int magic_event_destroy(void* ev);
int magic_event_record(void* ev, cudaStream_t st);
int magic_event_elapsed_ms(void* e0, void* e1, float* ms);
/* make `st` (a stream OUTSIDE the graph, e.g. the one NCCL is issued from) wait for an event-record node of the graph
 * launched before this call: the gradient exchange of a finished layer group starts while the graph still runs */
int magic_stream_wait_event(cudaStream_t st, void* ev);
/* scale_dev: optional device scalar multiplied into `scale` (see MagicMseSeg.scale_dev) */
int magic_makd_kl_fwd(const void* s, const void* t, int R, int C, long ld, float temperature, const float* w,
                      float scale, const float* scale_dev, float* stats /* [R,2] */,
                      float* loss /* [1], zeroed here */, int dtype, cudaStream_t st);
int magic_makd_kl_bwd(const void* s, const void* t, void* ds, int R, int C, long ld, float temperature,
                      const float* w, float scale, const float* scale_dev, const float* stats, const float* gout,
                      int dtype, cudaStream_t st);

/* ---- GPU batch featuriser (feature half): replaces the host-side per-sample feature read + collate + 36 x 768 fp32
 * H2D copy of pretrain_src/data/dataset.py:210-244,742-756 / data/tasks.py:121-133 with a gather from a device-
 * resident store.  out[r, j, :] = perm[r, j] >= 0 ? store[vp[r], perm[r, j], :] : 0  (store [n_store, V, D]);
 * gmap_pair_dists[b, i, j] = dist[v_i, v_j] from the resident all-pairs matrix (dataset.py:545-549), 0 for [stop]/pad */
int magic_gather_views(const void* store, int store_dt, long long n_store, const long long* vp, const int* perm,
                       void* out, int out_dt, long long R, int V, int D, cudaStream_t st);
int magic_gather_pair_dists(const float* dist, long long N, const long long* node_vp, float* out, int B, int G,
                            cudaStream_t st);

/* ---- GPU batch featuriser (graph half): builds, on the device, every tensor of a pretraining batch that derives from
 * the navigation graph -- what pretrain_src/data/dataset.py builds per sample in DataLoader workers (get_cur_angle
 * :433-443, get_traj_pano_fts :733-772, get_gmap_inputs :513-549, get_gmap_pos_fts :553-575, get_vp_pos_fts :577-586,
 * get_act_labels :622-640), what data/tasks.py:110-166 collates, and the integer tables the model's viewpoint-id string
 * loops reduce to (gmap source CSR + reverse, SAP masks, local -> global scatter).  Integer outputs are bit-exact
 * against those loops; float features agree to 1-2 ulp.  All pointers are device memory; `status` receives the largest
 * error code seen (0 = ok, 1 = more than 256 distinct viewpoints in a sample, 2 = graph larger than G, 3 = two candidates
 * share a view, 4 / 5 = entry capacities E_s / E_cap / S_cap too small). */
typedef struct MagicFeatArgs {
  /* world (resident): N viewpoints, at most C candidates each */
  const double* pos;       /* [N, 3] (fp64, like the connectivity files) */
  const float* dist;       /* [N, N] shortest distances */
  const int* hops;         /* [N, N] len(shortest path) - 1 */
  const int* cand_vp;      /* [N, C] candidate viewpoint rows in scanvp_cands dict order */
  const int* cand_view;    /* [N, C] view index of the candidate (v[0]) */
  const float* cand_ang;   /* [N, C, 2] heading / elevation offsets (v[2], v[3]) */
  const int* n_cand;       /* [N] */
  const float* view_ang;   /* [36, 2] all_point_rel_angles[12] */
  /* batch (per sample) */
  const int* path;          /* [B, Tmax] viewpoint rows */
  const int* path_len;      /* [B] */
  const float* start_heading; /* [B] */
  const int* next_vp;       /* [B] ground-truth next viewpoint, -1 = stop, -2 = unknown; NULL: no labels */
  const int* prev_vp;       /* [B] viewpoint the agent came from (the heading's source, dataset.py:433-443); -1 or NULL:
                               path[len - 2].  Differs when the loader cut a long path to 20 viewpoints + the end one */
  const int* row0;          /* [B] first panorama row of the sample (exclusive prefix sum of path_len) */
  /* outputs: panoramas [R_cap rows] */
  long long* traj_vp_index; int* traj_view_perm; float* traj_loc_fts; long long* traj_nav_types;
  long long* traj_vp_view_lens;
  /* outputs: graph / local branch / labels */
  long long* gmap_node_vp; long long* gmap_step_ids; unsigned char* gmap_visited_masks; long long* gmap_lens;
  float* gmap_pos_fts; float* gmap_pair_dists; float* vp_pos_fts; long long* global_act_labels;
  long long* local_act_labels;
  /* outputs: index tables (graph_index.build_index + pad_batch layout) */
  int* node_ptr; int* entries; int* src_ids; int* src_ptr; int* src_nodes; float* src_w; int* n_src;
  unsigned char* g_valid; unsigned char* l_valid; int* node2cand; unsigned char* bw_mask; long long* vp_gather;
  int* key_lens_gmap; int* key_lens_vp; long long* last_rows;
  /* scratch: per-sample slabs */
  int* slab_entries; int* slab_nodes; int* slab_rank; int* slab_ptr; int* slab_total; int* slab_nvis;
  int* status;
  int N, C, B, Tmax, G, Vp, R, R_cap, E_s, E_cap, S_cap, correct_heading;
} MagicFeatArgs;
int magic_featurize_graph(const MagicFeatArgs* args, cudaStream_t st);
int magic_feat_args_size(void); /* sizeof(MagicFeatArgs): lets a binding check its mirror of the struct */

/* ---- optimizer (pretrain_src/optim/adamw.py:53-112, clip grad_norm r2r_magic_pretrain.json:22) ----- */
/* out[0] (+)= sum g^2, deterministic (two-stage, no floating-point atomics: data-parallel replicas with identical
 * gradients compute the identical clip coefficient).  `out` must hold 1 + MAGIC_SUMSQ_SCRATCH floats: out[1..] is the
 * scratch of the per-CTA partials. */
int magic_sumsq(const float* g, long long n, float* out, int zero_first, cudaStream_t st);
int magic_adamw(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, const float* hyper,
                float weight_decay, const float* sumsq, cudaStream_t st);
/* The same update when not every parameter took part in the step: adamw.py:66-67 skips parameters whose grad is None
 * and :86 counts steps per parameter.  [0, n) is cut into nseg <= MAGIC_ADAMW_MAX_SEGS segments [bounds[s], bounds[s+1]);
 * codes[s] < 0 leaves the segment untouched, codes[s] >= 0 selects the hyper slot (hyper + 8 * code) whose step size
 * carries that segment's own step count.  bounds (nseg + 1 ints) / codes (nseg ints) are device memory. */
#define MAGIC_ADAMW_MAX_SEGS 512
int magic_adamw_seg(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, const float* hyper,
                    float weight_decay, const float* sumsq, const int* bounds, const int* codes, int nseg,
                    cudaStream_t st);
int magic_scale(float* x, long long n, float s, cudaStream_t st);

#ifdef __cplusplus
}
#endif
#endif /* MAGIC_B200_H_ */
