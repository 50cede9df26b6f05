// Shared parameter block of the attention kernels (SIMT fp32/bf16 path in attention.cu, bf16 tensor-core
// path in attention_mma.cu).
#pragma once
#include "common.cuh"

struct AttnParams {
  const void *q, *k, *v;
  long q_ld, k_ld, v_ld;        // elements between consecutive tokens
  void* out;                    // [B*Lq, H*64]
  float* lse;                   // [B, H, Lq]
  float* pbar;                  // optional head-mean probs
  long pbar_bs, pbar_rs;        // batch / row strides (elements)
  const int* key_lens;          // [B] or null
  int key_skip;                 // one key index that is never attended (the [MEM] slot of the navigation graph), -1 = none
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
  int delta_from_out;           // key-major backward: delta[i] = dO_i . O_i from `out` (no KD-map gradient) instead of
                                // reading `delta` written by the query-major kernel, so the two kernels are independent
};

// bf16 mma.sync path (attention_mma.cu).  Return MAGIC_ERR_UNSUPPORTED when the shape / alignment is outside
// what the tensor-core kernels cover; the caller then uses the SIMT kernels.
int attn_mma_fwd(const AttnParams& P, cudaStream_t st);
int attn_mma_bwd(const AttnParams& P, cudaStream_t st);
// part 1: query-major kernel only (delta, dQ, d sprel); part 2: key-major kernel only (dK, dV), delta from dO . O
int attn_mma_bwd_part(const AttnParams& P, int part, cudaStream_t st);
