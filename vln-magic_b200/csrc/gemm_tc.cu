// bf16 GEMM on the 5th-generation tensor cores (sm_100a), persistent and warp-specialised:
//
//   warp 0      TMA producer   cp.async.bulk.tensor (128B swizzle) -> smem ring (BK = 64, `stages` deep)
//   warp 1      MMA issuer     tcgen05.mma cta_group::1, UMMA 128 x BN x 16, fp32 accumulators in TMEM;
//                              TWO accumulator buffers (2*BN columns) so tile i+1 is multiplied while tile i
//                              is still being drained
//   warps 2..9  epilogue       tcgen05.ld (one TMEM lane quarter per warp, two warps per quarter splitting the
//                              columns) -> fused epilogue in registers -> 128B-swizzled smem staging ->
//                              TMA store (cp.async.bulk.tensor ... global.shared::cta), or TMA reduce-add for
//                              fp32 accumulation into the gradient arena (beta = 1, split-K)
//
// Each CTA walks work items (m tile, n tile, k split) round-robin; the three pipelines (smem full/empty,
// TMEM full/empty, per-warp staging buffers tracked by bulk async-groups) never meet at a CTA-wide barrier.
//
// Fused epilogue: alpha, bias, pre-activation copy-out, GELU / ReLU, activation-derivative (backward),
// dropout (stateless hash), residual.  M/N edges need no predication on stores: TMA clips the box.
//
// Both operands may be K-major (reduce dim contiguous) or MN-major (M / N contiguous), so the SAME kernel
// serves forward (x W^T), dgrad (dy W) and wgrad (dy^T x) of every nn.Linear without transposed copies:
//   K-major  : TMA box {64 k, rows}, canonical layout SBO = 1024 B, k-step = +32 B inside the swizzle atom
//   MN-major : TMA boxes {64 mn, 64 k} per 64-wide MN block, SBO = 1024 B, LBO = 8192 B, k-step = +2048 B
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_epi.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 64 + NUM_EPI_WARPS * 32;  // 10 warps
constexpr int STG_BYTES = 4096;                    // one staging buffer: 32 rows x 128 B
constexpr int STG_BUFS = 2;                        // per epilogue warp
constexpr int MAX_STAGES = 8;
constexpr int RS_COLS = 16;                        // accumulator columns of the row-sum (bias gradient) UMMA
constexpr int ONES_BYTES = RS_COLS * 128;          // 16 rows x 64 bf16 ones, K-major SWIZZLE_128B tile
constexpr int BIAS_BYTES = 2 * 256 * 4;            // bias of the current / next tile (BN <= 256 fp32 each)

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// One lane of a CONVERGED warp.  The producer and MMA warps run their loops with all 32 lanes (uniform control flow)
// and elect only around the issue instructions: TMA / tcgen05 take their operands from uniform registers, and
// inside an `if (lane == 0)` region ptxas cannot prove uniformity, so every UTMALDG / UTCHMMA was wrapped in an
// ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall loop (~170 cycles per issued MMA, measured).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC share the operands of a 256 x BN tile -----------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by EITHER CTA of the pair; the transaction bytes are credited to the LEADER's (rank 0) mbarrier:
// clearing bit 24 of the shared::cluster address selects the even CTA's copy of the barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP_C:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE_C;\n"
      "bra WAIT_LOOP_C;\n"
      "DONE_C:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive columns -> 32 registers per thread (no wait: pair with tmem_wait_ld)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 32 lanes x 1 column -> one register per thread
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
  return r;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;  // layout_type = SWIZZLE_128B
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- epilogue helpers: one thread = one accumulator row, 32 consecutive columns per call -------------------
__device__ __forceinline__ void ld_row32(const float* p, float (&v)[32]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float4 t = __ldg(q + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ld_row32(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 t = __ldg(q + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float2 f = __bfloat1622float2(h[e]);
      v[8 * i + 2 * e] = f.x; v[8 * i + 2 * e + 1] = f.y;
    }
  }
}
// guarded (edge) variant: columns >= ncols read as 0
template <typename T>
__device__ __forceinline__ void ld_row32_edge(const T* p, float (&v)[32], int ncols) {
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = j < ncols ? ldf(p, j) : 0.f;
}

// 32 fp32 values of row `lane` -> staging buffer in the TMA SWIZZLE_128B layout
// (row r at r*128 B, 16-byte chunk c at position c ^ (r & 7)); bf16: `half` selects columns 0..31 / 32..63
__device__ __forceinline__ void stage_row32(uint8_t* buf, int lane, int half, const float (&v)[32],
                                            const __nv_bfloat16*) {
  uint8_t* row = buf + lane * 128;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(v[8 * i + 2 * e], v[8 * i + 2 * e + 1]);
    const int c = half * 4 + i;
    *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) = t;
  }
}
__device__ __forceinline__ void stage_row32(uint8_t* buf, int lane, int, const float (&v)[32], const float*) {
  uint8_t* row = buf + lane * 128;
#pragma unroll
  for (int c = 0; c < 8; c++)
    *reinterpret_cast<float4*>(row + ((c ^ (lane & 7)) << 4)) =
        make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
template <typename T>
__device__ __forceinline__ void round_trip32(float (&v)[32]) {
  if (sizeof(T) == 2) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
  }
}

struct TcParams {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int total_kb, kb_per_split;
  int stages;
  int a_mn, b_mn;  // 1 = MN-major operand
  long ldc;
  int reduce;      // 1: accumulate into C with TMA reduce-add (fp32 C; beta = 1 and/or split-K)
  unsigned long long* trace;  // measurement aid (magic_gemm_set_trace): CTA 0 stamps clock64() at phase boundaries
  GemmEpi epi;
};

__device__ __forceinline__ void stamp(const TcParams& P, int i) {
  if (P.trace != nullptr && blockIdx.x == 0) {
    P.trace[i] = (unsigned long long)clock64();
    if (i == 0 || i == 10) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
      P.trace[16 + i] = g;
    }
  }
}

// activation on the tensor-core path: bf16 storage takes the branch-free one-MUFU forms (common.cuh gelu_tail)
template <typename TC, int ACT>
__device__ __forceinline__ float tc_act_fwd(float v) {
  if (ACT == MAGIC_ACT_GELU) return sizeof(TC) == 2 ? gelu_fast_f(v) : gelu_f(v);
  if (ACT == MAGIC_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}
template <typename TC, int ACT>
__device__ __forceinline__ float tc_act_bwd(float z) {
  if (ACT == MAGIC_ACT_GELU) return sizeof(TC) == 2 ? gelu_grad_fast_f(z) : gelu_grad_f(z);
  if (ACT == MAGIC_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

// One epilogue "unit" = UNIT_COLS columns of one accumulator row (128 bytes of C).  The side operand of the unit
// (residual in the forward pass, pre-activation in the backward pass; storage type of C) is fetched into registers
// BEFORE the thread waits for the accumulator, so its global-memory latency hides behind the MMA main loop.
template <typename TC>
struct SideRegs {
  uint4 r[8];  // 128 bytes: 64 bf16 or 32 fp32
};
// -> true when `sr` holds the unit (full-width unit, or a row beyond M whose values are never stored); false on
// the N edge, where epi_math32 falls back to guarded element loads
template <typename TC>
__device__ __forceinline__ bool side_load(SideRegs<TC>& sr, const TC* p, bool row_ok, int ncols) {
  constexpr int UC = 128 / (int)sizeof(TC);
  if (row_ok && ncols >= UC) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) sr.r[i] = __ldg(q + i);
    return true;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) sr.r[i] = make_uint4(0u, 0u, 0u, 0u);
  return !row_ok;
}
// columns [32*half, 32*half+32) of the unit as floats
// (`half` selects between two statically indexed copies so the registers never become a local-memory array)
template <int HALF>
__device__ __forceinline__ void side_get32_s(const SideRegs<__nv_bfloat16>& sr, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 t = sr.r[HALF * 4 + i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float2 f = __bfloat1622float2(h[e]);
      v[8 * i + 2 * e] = f.x; v[8 * i + 2 * e + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void side_get32(const SideRegs<__nv_bfloat16>& sr, int half, float (&v)[32]) {
  if (half == 0) side_get32_s<0>(sr, v);
  else side_get32_s<1>(sr, v);
}
__device__ __forceinline__ void side_get32(const SideRegs<float>& sr, int, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    v[4 * i] = __uint_as_float(sr.r[i].x); v[4 * i + 1] = __uint_as_float(sr.r[i].y);
    v[4 * i + 2] = __uint_as_float(sr.r[i].z); v[4 * i + 3] = __uint_as_float(sr.r[i].w);
  }
}

// fused epilogue math on 32 columns [nb, nb+32) of row m.  `v` in: raw accumulators; out: final values.
// If epi.pre_out is set, `pre` receives the pre-activation values (already in storage precision).
// `sbias`: this tile's bias staged in shared memory (zero beyond N), indexed from the tile's first column.
// `side`: prefetched residual (forward) or pre-activation (backward, when `side_is_dact`), else unused.
template <typename TC, int ACT, bool DACT>
__device__ __forceinline__ void epi_math32(const TcParams& P, const Dropout& dr, float (&v)[32], float (&pre)[32],
                                           int m, int nb, bool row_ok, const float* sbias, const SideRegs<TC>& side,
                                           int half, bool side_ok) {
  const GemmEpi& epi = P.epi;
  const int ncols = P.N - nb;  // may be < 32 on the N edge (or <= 0: whole chunk clipped by the store)
  const bool full = ncols >= 32;
  if (epi.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] *= epi.alpha;
  }
  if constexpr (DACT) {
    float z[32];
    const size_t poff = (size_t)m * epi.dact_ld + nb;
    if (side_ok) {
      side_get32(side, half, z);
    } else if (!row_ok || ncols <= 0) {
#pragma unroll
      for (int j = 0; j < 32; j++) z[j] = 0.f;
    } else if (epi.dact_dt == MAGIC_BF16) {
      if (full) ld_row32((const __nv_bfloat16*)epi.dact_pre + poff, z);
      else ld_row32_edge((const __nv_bfloat16*)epi.dact_pre + poff, z, ncols);
    } else {
      if (full) ld_row32((const float*)epi.dact_pre + poff, z);
      else ld_row32_edge((const float*)epi.dact_pre + poff, z, ncols);
    }
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] *= tc_act_bwd<TC, ACT>(z[j]);
    if (dr.p > 0.f) {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] *= dr.scale(poff + j);
    }
    return;
  }
  if (epi.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(sbias + j);  // warp-uniform address: broadcast
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (epi.pre_out) {
#pragma unroll
    for (int j = 0; j < 32; j++) pre[j] = v[j];
    round_trip32<TC>(v);  // the activation sees what backward will re-read
  }
  if (ACT != 0) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = tc_act_fwd<TC, ACT>(v[j]);
  }
  if (dr.p > 0.f) {
    const size_t off = (size_t)m * P.ldc + nb;
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] *= dr.scale(off + j);
  }
  if (epi.residual) {
    float rs[32];
    if (side_ok) {
      side_get32(side, half, rs);
    } else if (row_ok && ncols > 0) {
      const TC* rp = (const TC*)epi.residual + (size_t)m * epi.res_ld + nb;
      if (full) ld_row32(rp, rs);
      else ld_row32_edge(rp, rs, ncols);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j++) rs[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] += rs[j];
  }
}

// One kernel instantiation per (tile width, C type, activation, forward / backward-derivative epilogue): each carries
// only its own epilogue code.  These launches execute every instruction once or twice, so instruction fetch of a
// do-everything epilogue (3 activations x 2 directions, unrolled) was a first-order cost of the small GEMMs.
// CTAS = 2: a cluster of two CTAs owns a 256 x BN tile.  Each CTA loads ITS 128 rows of A and HALF of the B tile
// (BN/2 rows), the leader (cluster rank 0) issues tcgen05.mma.cta_group::2 (UMMA 256 x BN x 16) that reads both
// CTAs' shared memory and writes each CTA's 128 accumulator rows into its own TMEM, and both CTAs drain their
// rows.  Per SM and k-block the operand traffic drops from 16 KB + BN*128 B to 16 KB + BN*64 B for the same
// number of MMA cycles, which is what the L2 -> SM path (the bound of the 128 x BN single-CTA main loop) needs.
template <int BN, typename TC, int ACT, bool DACT, int CTAS>
__global__ void __launch_bounds__(NTHREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_pre,
                   const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_ROWS = BN / CTAS;  // rows of the B tile this CTA stages
  constexpr int B_BYTES = B_ROWS * BK * 2;
  constexpr int UNIT_COLS = 128 / (int)sizeof(TC);  // columns of one 128-byte staging row: 64 (bf16) / 32 (fp32)
  constexpr int UNITS = BN / UNIT_COLS;             // store units per lane quarter per tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stages = P.stages;
  uint8_t* sA = smem;
  uint8_t* sB = sA + stages * A_BYTES;
  uint8_t* sStage = sB + stages * B_BYTES;  // [NUM_EPI_WARPS][STG_BUFS][4096], 1024-aligned
  uint8_t* sOnes = sStage + NUM_EPI_WARPS * STG_BUFS * STG_BYTES;  // 1024-aligned (all staging sizes are)
  float* sBias = (float*)(sOnes + ONES_BYTES);  // [2][256]: this tile's bias (double-buffered over tiles)
  uint64_t* full = (uint64_t*)(sOnes + ONES_BYTES + BIAS_BYTES);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tmem_full = empty + MAX_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_work = P.m_tiles * P.n_tiles * P.splits;  // CTAS = 2: m_tiles counts 256-row pair tiles
  const int crank = CTAS == 2 ? (int)cluster_ctarank() : 0;
  const int wfirst = (int)blockIdx.x / CTAS, wstride = (int)gridDim.x / CTAS;  // work walk of this CTA (pair)
  const bool leader = crank == 0;
  pdl_trigger();
  if (threadIdx.x == 0) stamp(P, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    if (P.epi.pre_out) prefetch_tmap(&tmap_pre);
    for (int s = 0; s < stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], NUM_EPI_WARPS * CTAS);  // pair: both CTAs' epilogue warps release the leader's copy
    }
    fence_barrier_init();
  }
  const bool rowsum = P.epi.rowsum != nullptr;  // host guarantees BN <= 128 then
  const uint32_t tmem_cols = rowsum ? 4 * BN : 2 * BN;
  if (rowsum) {
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_proxy_async();  // the tensor core reads this tile through the async proxy
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_pair(tmem_ptr, tmem_cols);
    else tmem_alloc(tmem_ptr, tmem_cols);
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail.  griddepcontrol.wait is
  // executed per ROLE, as late as possible: the producer walks to its first TMA issue (tile coordinates, first
  // empty-slot wait -- which also pulls that code into the instruction cache) before it waits, the MMA warp never
  // touches global memory and does not wait at all, the epilogue warps wait before their first global access.
  if (threadIdx.x == 0) stamp(P, 1);

  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    bool pdl_done = false;
    for (int work = wfirst; work < total_work; work += wstride) {
      const int tile = work % (P.m_tiles * P.n_tiles), ks = work / (P.m_tiles * P.n_tiles);
      const int m0 = (tile / P.n_tiles) * (BM * CTAS) + crank * BM, n0 = (tile % P.n_tiles) * BN + crank * B_ROWS;
      const int kb0 = ks * P.kb_per_split, kb1 = min(P.total_kb, kb0 + P.kb_per_split);
      for (int kb = kb0; kb < kb1; kb++) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        const int k0 = kb * BK;
        if (!pdl_done) {  // first k-block of this CTA: the operands are the predecessor's outputs
          pdl_wait();
          pdl_done = true;
        }
        if (elect_one()) {
          if (work == 0 && kb == kb0) stamp(P, 14);
          if (CTAS == 2) {
            // both CTAs' bytes land on the leader's barrier; only the leader posts the expected count
            if (leader) mbar_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
            if (!P.a_mn) {
              tma_load_2d_pair(&tmap_a, &full[s], a_dst, k0, m0);
            } else {
              tma_load_2d_pair(&tmap_a, &full[s], a_dst, m0, k0);
              tma_load_2d_pair(&tmap_a, &full[s], a_dst + 64 * BK * 2, m0 + 64, k0);
            }
            if (!P.b_mn) {
              tma_load_2d_pair(&tmap_b, &full[s], b_dst, k0, n0);  // box {64 k, BN/2 rows}
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; j++)
                tma_load_2d_pair(&tmap_b, &full[s], b_dst + j * 64 * BK * 2, n0 + 64 * j, k0);
            }
          } else {
            mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
            if (!P.a_mn) {
              tma_load_2d(&tmap_a, &full[s], a_dst, k0, m0);  // box {64 k, 128 rows}
            } else {
              tma_load_2d(&tmap_a, &full[s], a_dst, m0, k0);  // box {64 m, 64 k} x 2
              tma_load_2d(&tmap_a, &full[s], a_dst + 64 * BK * 2, m0 + 64, k0);
            }
            if (!P.b_mn) {
              tma_load_2d(&tmap_b, &full[s], b_dst, k0, n0);  // box {64 k, BN rows}
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; j++)
                tma_load_2d(&tmap_b, &full[s], b_dst + j * 64 * BK * 2, n0 + 64 * j, k0);
            }
          }
          if (work == 0 && kb == kb0) stamp(P, 2);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (pair: the leader CTA only; whole warp walks the loop, one elected lane issues) =====
      const uint32_t idesc = make_idesc(BM * CTAS, BN, P.a_mn, P.b_mn);
      const uint32_t rs_idesc = make_idesc(BM, RS_COLS, P.a_mn, 0);
      const uint64_t ones_desc = make_desc(smem_u32(sOnes), 16, 1024);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int work = wfirst; work < total_work; work += wstride, it++) {
        const int ks = work / (P.m_tiles * P.n_tiles);
        const int kb0 = ks * P.kb_per_split, kb1 = min(P.total_kb, kb0 + P.kb_per_split);
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        // bias gradient: the CTAs of the first tile column also multiply A by a tile of ones
        const bool rs_tile = rowsum && (work % (P.m_tiles * P.n_tiles)) % P.n_tiles == 0;
        const uint32_t tmem_rs = tmem_base + (uint32_t)(2 * BN + as * RS_COLS);
        if (CTAS == 2) mbar_wait_cluster(&tmem_empty[as], aph ^ 1);  // both CTAs have drained this buffer
        else mbar_wait(&tmem_empty[as], aph ^ 1);                   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
          if (elect_one()) {
            if (work == 0 && kb == kb0) stamp(P, 3);
            if (work == 0 && kb == kb0 + 1) stamp(P, 13);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) {
              // K-major: +32 B per UMMA_K inside the 128 B swizzle row; MN-major: +16 k-rows * 128 B
              const uint64_t adesc =
                  P.a_mn ? make_desc(a_base + k * 2048, 8192, 1024) : make_desc(a_base + k * 32, 16, 1024);
              const uint64_t bdesc =
                  P.b_mn ? make_desc(b_base + k * 2048, 8192, 1024) : make_desc(b_base + k * 32, 16, 1024);
              if (CTAS == 2) {
                umma_bf16_pair(tmem_c, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              } else {
                umma_bf16(tmem_c, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                if (rs_tile) umma_bf16(tmem_rs, adesc, ones_desc, rs_idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              }
            }
            // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
            if (CTAS == 2) umma_commit_pair(&empty[s]);
            else umma_commit(&empty[s]);
            if (work == 0 && kb == kb0) stamp(P, 12);
            if (kb == kb1 - 1) {  // accumulator complete (pair: both CTAs' epilogues wake)
              if (CTAS == 2) umma_commit_pair(&tmem_full[as]);
              else umma_commit(&tmem_full[as]);
              if (work == 0) stamp(P, 4);
            }
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int et = (int)threadIdx.x - 64;  // 0..255 within the epilogue warps
    uint8_t* stg = sStage + (warp - 2) * (STG_BUFS * STG_BYTES);
    pdl_wait();  // bias / residual / seed loads and the stores below touch global memory
    const Dropout dr = make_dropout(P.epi.drop_p, P.epi.seed_ptr, P.epi.salt);
    const bool has_pre = P.epi.pre_out != nullptr;
    const bool use_bias = !DACT && P.epi.bias != nullptr;
    // side operand prefetched into registers: pre-activation (backward) or residual (forward), storage type of C
    const bool side_dact = DACT && P.epi.dact_dt == (sizeof(TC) == 2 ? MAGIC_BF16 : MAGIC_F32);
    const bool side_res = !DACT && P.epi.residual != nullptr;
    const bool has_side = side_dact || side_res;
    const TC* side_base = side_dact ? (const TC*)P.epi.dact_pre : (const TC*)P.epi.residual;
    const long side_ld = side_dact ? P.epi.dact_ld : P.epi.res_ld;
    int nstore = 0;  // staging buffers used so far by this warp (buffer = nstore & 1)
    int it = 0;
    for (int work = wfirst; work < total_work; work += wstride, it++) {
      const int tile = work % (P.m_tiles * P.n_tiles);
      const int m0 = (tile / P.n_tiles) * (BM * CTAS) + crank * BM, n0 = (tile % P.n_tiles) * BN;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < P.M;
      // everything that does not depend on the accumulator happens BEFORE the wait: the tile's bias goes to shared
      // memory (one global load per thread instead of 64 on the critical path) and the first unit's side operand
      // to registers; both latencies hide behind the TMA / MMA main loop
      float* sb = sBias + (it & 1) * 256;
      if (use_bias) {
        if (et < BN) sb[et] = (n0 + et < P.N) ? __ldg(P.epi.bias + n0 + et) : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
      }
      SideRegs<TC> side;
      bool side_ok = false;
      if (has_side && hf < UNITS) {
        const int nu = n0 + hf * UNIT_COLS;
        side_ok = side_load(side, side_base + (size_t)m * side_ld + nu, row_ok, P.N - nu);
      }
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const bool tr = work == 0 && warp == 2 && lane == 0;
      if (tr) stamp(P, 5);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      if (rowsum && hf == 0 && n0 == 0) {  // column 0 of the ones product = sum_k A(m, k)
        const uint32_t r = tmem_ld1(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * BN + as * RS_COLS));
        tmem_wait_ld();
        if (row_ok) atomicAdd(P.epi.rowsum + m, P.epi.alpha * __uint_as_float(r));
      }
      bool released = false;
      if (UNITS <= hf) {  // nothing to drain for this warp (BN = 64 with bf16 output): just release
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
        released = true;
      }
#pragma unroll 1
      for (int u = hf; u < UNITS; u += 2) {
        const int nu = n0 + u * UNIT_COLS;
        if (has_side && u != hf)  // later units: issued before the TMEM read so the two latencies overlap
          side_ok = side_load(side, side_base + (size_t)m * side_ld + nu, row_ok, P.N - nu);
        if (nu >= P.N) {  // unit entirely beyond the N edge (warp-uniform): nothing to read or store
          if (u + 2 >= UNITS && !released) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
            released = true;
          }
          continue;
        }
        // staging buffers: at most one store group may still be reading (the other buffer)
        if (elect_one()) {  // (elect.sync is deterministic per member mask: the same lane owns the bulk groups)
          if (has_pre) bulk_wait_read<0>();
          else bulk_wait_read<1>();
        }
        __syncwarp();
        uint8_t* bufc = stg + (nstore & 1) * STG_BYTES;
        nstore++;
        uint8_t* bufp = nullptr;
        if (has_pre) {
          bufp = stg + (nstore & 1) * STG_BYTES;
          nstore++;
        }
        constexpr int HALVES = 4 / (int)sizeof(TC);  // 32-column TMEM reads per unit: 2 (bf16 C) / 1 (fp32 C)
#pragma unroll 1
        for (int half = 0; half < HALVES; half++) {
          uint32_t r[32];
          tmem_ld32(t_row + (uint32_t)(u * UNIT_COLS + 32 * half), r);
          tmem_wait_ld();
          if (tr && u == hf && half == 0) stamp(P, 6);
          if (half == HALVES - 1 && u + 2 >= UNITS && !released) {  // last TMEM read of this warp for this tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]); }
            released = true;
          }
          float v[32], pre[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
          epi_math32<TC, ACT, DACT>(P, dr, v, pre, m, nu + 32 * half, row_ok, sb + (nu - n0) + 32 * half, side, half,
                                    side_ok);
          stage_row32(bufc, lane, half, v, (const TC*)nullptr);
          if (has_pre) stage_row32(bufp, lane, half, pre, (const TC*)nullptr);
        }
        if (tr && u == hf) stamp(P, 7);
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
          if (P.reduce) tma_reduce_add_2d(&tmap_c, bufc, nu, m0 + q * 32);
          else tma_store_2d(&tmap_c, bufc, nu, m0 + q * 32);
          if (has_pre) tma_store_2d(&tmap_pre, bufp, nu, m0 + q * 32);
          bulk_commit();
        }
        __syncwarp();
        if (tr && u == hf) stamp(P, 8);
      }
    }
    if (elect_one()) bulk_wait_read<0>();  // smem must outlive the stores' reads; the writes land before grid completion
  }
  if (warp == 2 && lane == 0) stamp(P, 9);
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();  // the leader's MMAs read the peer's shared memory: leave together
  else __syncthreads();
  if (threadIdx.x == 0) stamp(P, 10);
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
    if (lane == 0) stamp(P, 11);
  }
}

// ---- host side: tensor-map cache ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, stride;
  uint32_t box_inner, box_outer, esz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
           box_outer == o.box_outer && esz == o.esz;
  }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.stride;
    h = h * 1000003u ^ ((size_t)k.box_inner << 20 | (size_t)k.box_outer << 4 | k.esz);
    return h;
  }
};

std::mutex g_mu;
std::unordered_map<MapKey, CUtensorMap, MapHash> g_maps;

// 2-D tensor of bf16 (esz 2) or fp32 (esz 4): `inner` contiguous elements, `outer` rows of `stride` elements
int get_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride, uint32_t box_inner, uint32_t box_outer,
            uint32_t esz, CUtensorMap* out) {
  MapKey key{ptr, inner, outer, stride, box_inner, box_outer, esz};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return MAGIC_OK;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    magic_set_error("magic_gemm(tc): cuTensorMapEncodeTiled entry point unavailable");
    return MAGIC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {stride * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    magic_set_error("magic_gemm(tc): cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu stride=%llu esz=%u", (int)r,
                    (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride, esz);
    return MAGIC_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_maps.size() > 8192) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return MAGIC_OK;
}

// SMs the persistent GEMM may occupy (MAGIC_TC_MAX_SMS, default all; magic_gemm_set_sm_budget overrides per call site)
int g_sm_budget = 0;
int tc_sm_budget() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("MAGIC_TC_MAX_SMS");
    env = e ? atoi(e) : 0;
  }
  int v = g_sm_budget > 0 ? g_sm_budget : env;
  const int sms = magic_num_sms();
  if (v <= 0 || v > sms) v = sms;
  return v < 2 ? 2 : v;
}

constexpr size_t SMEM_MAX = 227 * 1024;
constexpr size_t SMEM_FIXED = 1024 /*align slack*/ + NUM_EPI_WARPS * STG_BUFS * STG_BYTES + ONES_BYTES + BIAS_BYTES +
                              (2 * MAX_STAGES + 4) * 8 + 16;

template <int BN, typename TC, int ACT, bool DACT, int CTAS>
int launch_tc_k(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp, TcParams& P,
                cudaStream_t st) {
  constexpr size_t stage_bytes = (size_t)BM * BK * 2 + (size_t)(BN / CTAS) * BK * 2;
  int stages = (int)((SMEM_MAX - SMEM_FIXED) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  // never more stages than k-blocks a CTA will ever load
  const long work = (long)P.m_tiles * P.n_tiles * P.splits;
  const int slots = tc_sm_budget() / CTAS;  // CTAs (CTAS = 1) or CTA pairs (CTAS = 2) that run at once
  // balanced waves: the SMALLEST grid that still finishes in ceil(work / slots) rounds.  5120-row GEMMs of the
  // h = 768 encoders make 60 / 180 / 240 pair tiles: 60 CTA pairs need 1 / 3 / 4 rounds, exactly like 74 would, and
  // the 28 SMs left alone run the kernels of the concurrent branches (student, panorama encoder, next teacher) that
  // would otherwise queue behind a persistent CTA holding an SM's whole shared memory and TMEM
  const long waves = (work + slots - 1) / slots;
  const int groups = (int)((work + waves - 1) / waves);
  const long per_cta_kb = ((work + groups - 1) / groups) * (long)P.kb_per_split;
  if (stages > per_cta_kb) stages = (int)(per_cta_kb < 2 ? 2 : per_cta_kb);
  P.stages = stages;
  const size_t smem = SMEM_FIXED + (size_t)stages * stage_bytes;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MAGIC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, TC, ACT, DACT, CTAS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX),
               "magic_gemm(tc)");
    attr_smem = SMEM_MAX;
  }
  if (CTAS == 2) {
    MAGIC_CUDA(magic_launch_cluster(gemm_tc_kernel<BN, TC, ACT, DACT, CTAS>, dim3(groups * 2), dim3(NTHREADS), smem, st,
                                    dim3(2, 1, 1), ta, tb, tc, tp, P),
               "magic_gemm(tc, cta pair)");
  } else {
    MAGIC_CUDA(magic_launch(gemm_tc_kernel<BN, TC, ACT, DACT, CTAS>, dim3(groups), dim3(NTHREADS), smem, st, ta, tb, tc,
                            tp, P),
               "magic_gemm(tc)");
  }
  return MAGIC_OK;
}

// epilogue variant from the call: fp32 C (weight gradients, fp32 heads) only ever takes the linear epilogue
template <int BN, typename TC, int CTAS>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tp, TcParams& P,
              cudaStream_t st) {
  const bool dact = P.epi.dact_pre != nullptr;
  const int act = P.epi.act;
  if (!dact && act == MAGIC_ACT_NONE) return launch_tc_k<BN, TC, MAGIC_ACT_NONE, false, CTAS>(ta, tb, tc, tp, P, st);
  if (sizeof(TC) == 2) {
    typedef __nv_bfloat16 bf;
    if (!dact && act == MAGIC_ACT_GELU) return launch_tc_k<BN, bf, MAGIC_ACT_GELU, false, CTAS>(ta, tb, tc, tp, P, st);
    if (!dact && act == MAGIC_ACT_RELU) return launch_tc_k<BN, bf, MAGIC_ACT_RELU, false, CTAS>(ta, tb, tc, tp, P, st);
    if (dact && act == MAGIC_ACT_GELU) return launch_tc_k<BN, bf, MAGIC_ACT_GELU, true, CTAS>(ta, tb, tc, tp, P, st);
    if (dact && act == MAGIC_ACT_RELU) return launch_tc_k<BN, bf, MAGIC_ACT_RELU, true, CTAS>(ta, tb, tc, tp, P, st);
    if (dact) return launch_tc_k<BN, bf, MAGIC_ACT_NONE, true, CTAS>(ta, tb, tc, tp, P, st);
  }
  return MAGIC_ERR_UNSUPPORTED;  // fp32 C with an activation epilogue: the caller falls back to the FFMA kernel
}

bool tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_DISABLE_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

unsigned long long* g_trace = nullptr;

// MAGIC_TC_PAIR: 0 = never use CTA pairs, 1 (default) = by the heuristic, 2 = whenever the shape allows
int pair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_TC_PAIR");
    v = e ? atoi(e) : 1;
  }
  return v;
}

int force_bn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_TC_BN");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace

extern "C" int magic_gemm_set_sm_budget(int sms) {
  g_sm_budget = sms;
  return MAGIC_OK;
}

extern "C" int magic_gemm_set_trace(unsigned long long* dev_buf) {
  g_trace = dev_buf;
  return MAGIC_OK;
}

int gemm_tc_shape_ok(int M, int N, int K) { return (!tc_disabled() && M > 0 && N > 0 && K > 0) ? 1 : 0; }

int gemm_tc_dispatch(const void* A, const void* B, void* C, int c_dt, int M, int N, int K, long sam, long sak, long sbk,
                     long sbn, long ldc, const GemmEpi& epi, cudaStream_t st) {
  if (tc_disabled()) return MAGIC_ERR_UNSUPPORTED;
  int a_mn, b_mn;
  long lda, ldb;
  if (sak == 1) { a_mn = 0; lda = sam; } else if (sam == 1) { a_mn = 1; lda = sak; } else return MAGIC_ERR_UNSUPPORTED;
  if (sbk == 1) { b_mn = 0; ldb = sbn; } else if (sbn == 1) { b_mn = 1; ldb = sbk; } else return MAGIC_ERR_UNSUPPORTED;
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!al16(A) || !al16(B) || (lda % 8) || (ldb % 8) || lda <= 0 || ldb <= 0) return MAGIC_ERR_UNSUPPORTED;
  if (a_mn == 0 && lda < K) return MAGIC_ERR_UNSUPPORTED;
  if (a_mn == 1 && lda < M) return MAGIC_ERR_UNSUPPORTED;
  if (b_mn == 0 && ldb < K) return MAGIC_ERR_UNSUPPORTED;
  if (b_mn == 1 && ldb < N) return MAGIC_ERR_UNSUPPORTED;
  // epilogue operands must be legal for TMA stores / 128-bit loads
  const int cesz = c_dt == MAGIC_BF16 ? 2 : 4;
  if (!al16(C) || (ldc * cesz) % 16 != 0 || ldc < N) return MAGIC_ERR_UNSUPPORTED;
  if (epi.bias && !al16(epi.bias)) return MAGIC_ERR_UNSUPPORTED;
  if (epi.pre_out && !al16(epi.pre_out)) return MAGIC_ERR_UNSUPPORTED;
  if (epi.residual && (!al16(epi.residual) || (epi.res_ld * cesz) % 16 != 0)) return MAGIC_ERR_UNSUPPORTED;
  if (epi.dact_pre && (!al16(epi.dact_pre) || (epi.dact_ld * (epi.dact_dt == MAGIC_BF16 ? 2 : 4)) % 16 != 0))
    return MAGIC_ERR_UNSUPPORTED;
  const bool linear_epi = c_dt == MAGIC_F32 && !epi.bias && epi.act == 0 && !epi.pre_out && !epi.dact_pre &&
                          !epi.residual && epi.drop_p == 0.f;
  if (epi.beta != 0.f && !(epi.beta == 1.f && linear_epi)) return MAGIC_ERR_UNSUPPORTED;

  TcParams P;
  P.M = M; P.N = N; P.K = K; P.a_mn = a_mn; P.b_mn = b_mn; P.ldc = ldc; P.epi = epi;
  P.epi.atomic = 0;
  P.trace = g_trace;
  P.m_tiles = (M + BM - 1) / BM;
  P.total_kb = (K + BK - 1) / BK;
  // tile width: the widest tile that still gives every SM work (wider tiles re-read less of A)
  const int sms = magic_num_sms();
  int BN = 64;
  if (N > 64) {
    BN = 128;
    const long t128 = (long)P.m_tiles * ((N + 127) / 128);
    const long t256 = (long)P.m_tiles * ((N + 255) / 256);
    if (N > 128 && t256 >= sms && P.total_kb >= 4) BN = 256;
    else if (t128 < sms / 2) BN = 64;
    // (tried: 64-wide tiles where 128-wide ones spill into a mostly empty extra wave, e.g. M = 5120, N = 512 --
    // slower, 12.7 vs 12.1 us: at BN = 64 only half of the epilogue warps have a 128-byte unit to drain)
  }
  if (force_bn() == 64 || force_bn() == 128 || force_bn() == 256) BN = force_bn();
  if (epi.rowsum && BN > 128) BN = 128;  // the row-sum accumulators need TMEM columns beyond the two tile buffers
  // CTA pairs (256 x 256 tiles, cta_group::2) for the wide GEMMs with a real reduction depth (K >= 512): bf16 C, no
  // bias-gradient rider, and enough pair tiles to occupy at least half of the 74 pairs.  Measured crossover
  // (scripts/pair_check.py small): below ~2.4 M outputs or at K <= 384 the single-CTA tiles are 0.3-1.4 us faster
  bool pair = false;
  if (pair_mode() != 0 && c_dt == MAGIC_BF16 && !epi.rowsum && epi.beta == 0.f && N >= 256 && M >= 512 && P.total_kb >= 8) {
    const long pt = (long)((M + 255) / 256) * ((N + 255) / 256);
    pair = pt * 4 >= sms || pair_mode() == 2;
  }
  if (pair) {
    BN = 256;
    P.m_tiles = (M + 2 * BM - 1) / (2 * BM);
  }
  P.n_tiles = (N + BN - 1) / BN;
  // split-K for skinny outputs with a long reduction (weight gradients): fp32 C, purely linear epilogue
  P.splits = 1;
  P.kb_per_split = P.total_kb;
  P.reduce = epi.beta == 1.f ? 1 : 0;
  const long tiles = (long)P.m_tiles * P.n_tiles;
  if (!pair && linear_epi && tiles * 2 <= sms && P.total_kb >= 8) {
    int splits = (int)(sms / tiles);
    if (splits > P.total_kb / 4) splits = P.total_kb / 4;
    if (splits > 1) {
      P.kb_per_split = (P.total_kb + splits - 1) / splits;
      P.splits = (P.total_kb + P.kb_per_split - 1) / P.kb_per_split;
      P.reduce = 1;
      if (epi.beta == 0.f) {
        if (ldc == N) {
          MAGIC_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st), "magic_gemm(tc) split-K memset");
        } else {
          MAGIC_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st),
                     "magic_gemm(tc) split-K memset");
        }
      }
    }
  }
  CUtensorMap ta, tb, tc, tp;
  int rc;
  if (!a_mn) rc = get_map(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM, 2, &ta);
  else rc = get_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64, 2, &ta);
  if (rc) return rc;
  if (!b_mn) rc = get_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)(pair ? BN / 2 : BN), 2, &tb);
  else rc = get_map(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64, 2, &tb);
  if (rc) return rc;
  const uint32_t unit_cols = 128 / cesz;
  rc = get_map(C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, unit_cols, 32, cesz, &tc);
  if (rc) return rc;
  tp = tc;
  if (epi.pre_out) {
    rc = get_map(epi.pre_out, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, unit_cols, 32, cesz, &tp);
    if (rc) return rc;
  }
  typedef __nv_bfloat16 bf;
  if (pair) return launch_tc<256, bf, 2>(ta, tb, tc, tp, P, st);
  if (c_dt == MAGIC_BF16) {
    if (BN == 64) return launch_tc<64, bf, 1>(ta, tb, tc, tp, P, st);
    if (BN == 128) return launch_tc<128, bf, 1>(ta, tb, tc, tp, P, st);
    return launch_tc<256, bf, 1>(ta, tb, tc, tp, P, st);
  }
  if (BN == 64) return launch_tc<64, float, 1>(ta, tb, tc, tp, P, st);
  if (BN == 128) return launch_tc<128, float, 1>(ta, tb, tc, tp, P, st);
  return launch_tc<256, float, 1>(ta, tb, tc, tp, P, st);
}
