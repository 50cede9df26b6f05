// bf16 GEMM on the 5th-generation tensor cores (sm_100a): TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// memory -> tcgen05.mma (one elected thread, cta_group::1, UMMA 128 x BN x 16) -> fp32 accumulators in TMEM
// -> tcgen05.ld -> fused epilogue (bias / GELU / ReLU / activation-derivative / dropout / residual / beta)
// -> global.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (one TMEM lane quarter each).  4-stage mbarrier ring, BK = 64 (one 128-byte swizzle atom of bf16).
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
constexpr int STAGES = 4;
constexpr int NTHREADS = 192;  // 6 warps

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

// ---- vectorised row-per-thread epilogue (the common, dropout-free modes) ---------------------------------
// Each epilogue thread owns one accumulator row and 32 consecutive columns: 64 B (bf16) / 128 B (fp32) of
// contiguous output per chunk, moved with 128-bit loads/stores (full 32 B sectors, ~20x fewer instructions
// than an element-wise loop -- the epilogue runs one warp per scheduler, so instruction count is latency).
__device__ __forceinline__ void ld_row32(const float* p, float (&v)[32]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float4 t = q[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ld_row32(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 t = q[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const float2 f = __bfloat1622float2(h[e]);
      v[8 * i + 2 * e] = f.x; v[8 * i + 2 * e + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void st_row32(float* p, const float (&v)[32]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void st_row32(__nv_bfloat16* p, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; e++) h[e] = __floats2bfloat162_rn(v[8 * i + 2 * e], v[8 * i + 2 * e + 1]);
    q[i] = t;
  }
}
template <typename T>
__device__ __forceinline__ void rt_row32(float (&v)[32]) {
  if (sizeof(T) == 2) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
  }
}

template <typename TC, int ACT>
__device__ __forceinline__ void epi_fast_chunk(const GemmEpi& epi, const Dropout& dr, TC* C, const uint32_t (&r)[32],
                                               int m, int nb, long ldc) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = epi.alpha * __uint_as_float(r[j]);
  const size_t off = (size_t)m * ldc + nb;
  if (epi.dact_pre != nullptr) {
    float pre[32];
    const size_t poff = (size_t)m * epi.dact_ld + nb;
    if (epi.dact_dt == MAGIC_BF16) ld_row32((const __nv_bfloat16*)epi.dact_pre + poff, pre);
    else ld_row32((const float*)epi.dact_pre + poff, pre);
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] *= act_bwd(ACT, pre[j]);
    if (dr.p > 0.f) {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] *= dr.scale(poff + j);
    }
  } else {
    if (epi.bias) {
      float b[32];
      ld_row32(epi.bias + nb, b);
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] += b[j];
    }
    if (epi.pre_out) {
      st_row32((TC*)epi.pre_out + off, v);
      rt_row32<TC>(v);
    }
    if (ACT != 0) {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = act_fwd(ACT, v[j]);
    }
    if (dr.p > 0.f) {
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] *= dr.scale(off + j);
    }
    if (epi.residual) {
      float rs[32];
      ld_row32((const TC*)epi.residual + (size_t)m * epi.res_ld + nb, rs);
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] += rs[j];
    }
  }
  if (epi.atomic) {
#pragma unroll
    for (int j = 0; j < 32; j++) atomicAdd(reinterpret_cast<float*>(C) + off + j, v[j]);
    return;
  }
  if (epi.beta != 0.f) {
    float c[32];
    ld_row32(C + off, c);
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] += epi.beta * c[j];
  }
  st_row32(C + off, v);
}

struct TcParams {
  int M, N, K;
  int kb_per_split;  // k-blocks handled by one CTA along gridDim.z (split-K)
  int a_mn, b_mn;  // 1 = MN-major operand
  long ldc;
  GemmEpi epi;
  int fast_epi;             // 1: vectorised row-per-thread epilogue is legal (alignment, no dropout)
  unsigned long long* dbg;  // optional per-CTA timestamps (MAGIC_TC_DEBUG), else null
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define DBG(slot)                                                                            \
  if (P.dbg && (threadIdx.x & 31) == 0)                                                      \
    P.dbg[((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot))] = gtime();

template <int BN, typename TC>
__global__ void __launch_bounds__(NTHREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   TC* __restrict__ C, const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [STAGES][A 16 KB][B BN*128 B] (1024-aligned), then barriers
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = (uint64_t*)(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = (uint32_t*)(tmem_full + 1);
  float* stage = (float*)(tmem_ptr + 4);  // [4 epilogue warps][32][33] fp32 staging for coalesced stores

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (threadIdx.x == 0) { DBG(0) }
  const int total_kb = (P.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * P.kb_per_split;
  const int num_kb = min(P.kb_per_split, total_kb - kb0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) { DBG(1) }

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        const int k0 = (kb0 + kb) * BK;
        if (!P.a_mn) {
          tma_load_2d(&tmap_a, &full[s], a_dst, k0, m0);                       // box {64 k, 128 rows}
        } else {
          tma_load_2d(&tmap_a, &full[s], a_dst, m0, k0);                       // box {64 m, 64 k} x 2
          tma_load_2d(&tmap_a, &full[s], a_dst + 64 * BK * 2, m0 + 64, k0);
        }
        if (!P.b_mn) {
          tma_load_2d(&tmap_b, &full[s], b_dst, k0, n0);                       // box {64 k, BN rows}
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; j++) tma_load_2d(&tmap_b, &full[s], b_dst + j * 64 * BK * 2, n0 + 64 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(BM, BN, P.a_mn, P.b_mn);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (kb == 0) { DBG(2) }
        const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; k++) {
          // K-major: +32 B per UMMA_K inside the 128 B swizzle row; MN-major: +16 k-rows * 128 B
          const uint64_t adesc = P.a_mn ? make_desc(a_base + k * 2048, 8192, 1024) : make_desc(a_base + k * 32, 16, 1024);
          const uint64_t bdesc = P.b_mn ? make_desc(b_base + k * 2048, 8192, 1024) : make_desc(b_base + k * 32, 16, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0);
        }
        umma_commit(&empty[s]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    // TMEM -> registers (row = lane) -> padded smem -> rolled loop with lane = column, so the code stays
    // small (instruction cache) and every global access of a warp is one contiguous row segment.
    const int q = warp & 3;
    float* st = stage + q * (32 * 33);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (warp == 2) { DBG(3) }
    const Dropout dr = make_dropout(P.epi.drop_p, P.epi.seed_ptr, P.epi.salt);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (P.fast_epi && n0 + c0 + 32 <= P.N) {  // warp-uniform
        const int m = m0 + q * 32 + lane;
        if (m < P.M) {
          if (P.epi.act == MAGIC_ACT_GELU) epi_fast_chunk<TC, MAGIC_ACT_GELU>(P.epi, dr, C, r, m, n0 + c0, P.ldc);
          else if (P.epi.act == MAGIC_ACT_RELU) epi_fast_chunk<TC, MAGIC_ACT_RELU>(P.epi, dr, C, r, m, n0 + c0, P.ldc);
          else epi_fast_chunk<TC, MAGIC_ACT_NONE>(P.epi, dr, C, r, m, n0 + c0, P.ldc);
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; j++) st[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      const int n = n0 + c0 + lane;
      if (n < P.N) {
        const float bias_n = P.epi.bias ? __ldg(P.epi.bias + n) : 0.f;
        const int rows = min(32, P.M - (m0 + q * 32));
#pragma unroll 4
        for (int rr = 0; rr < rows; rr++)
          epi_store<TC>(P.epi, dr, C, st[rr * 33 + lane], m0 + q * 32 + rr, n, P.ldc, bias_n);
      }
      __syncwarp();
    }
    if (warp == 2) { DBG(4) }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
    DBG(5)
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
  uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && stride == o.stride && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct MapHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.ptr;
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.stride;
    h = h * 1000003u ^ ((size_t)k.box_inner << 16 | k.box_outer);
    return h;
  }
};

unsigned long long* g_dbg_buf = nullptr;
std::mutex g_mu;
std::unordered_map<MapKey, CUtensorMap, MapHash> g_maps;

// 2-D bf16 tensor: `inner` contiguous elements, `outer` rows of `stride` elements
int get_map(const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride, uint32_t box_inner, uint32_t box_outer,
            CUtensorMap* out) {
  MapKey key{ptr, inner, outer, stride, box_inner, box_outer};
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
  cuuint64_t strides[1] = {stride * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    magic_set_error("magic_gemm(tc): cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu stride=%llu", (int)r,
                    (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride);
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

template <int BN, typename TC>
int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, void* C, const TcParams& P, cudaStream_t st) {
  constexpr size_t smem = 1024 + (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 +
                          4 * 32 * 33 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    MAGIC_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
               "magic_gemm(tc)");
    attr_set = true;
  }
  const int total_kb = (P.K + BK - 1) / BK;
  dim3 grid((P.N + BN - 1) / BN, (P.M + BM - 1) / BM, (total_kb + P.kb_per_split - 1) / P.kb_per_split);
  gemm_tc_kernel<BN, TC><<<grid, NTHREADS, smem, st>>>(ta, tb, (TC*)C, P);
  MAGIC_CHECK_LAUNCH("magic_gemm(tc)");
  return MAGIC_OK;
}

bool tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_DISABLE_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

extern "C" unsigned long long* magic_tc_debug_buffer(void) { return g_dbg_buf; }

int gemm_tc_shape_ok(int M, int N, int K) { return (!tc_disabled() && M > 0 && N > 0 && K > 0) ? 1 : 0; }

int gemm_tc_dispatch(const void* A, const void* B, void* C, int c_dt, int M, int N, int K, long sam, long sak, long sbk,
                     long sbn, long ldc, const GemmEpi& epi, cudaStream_t st) {
  if (tc_disabled()) return MAGIC_ERR_UNSUPPORTED;
  int a_mn, b_mn;
  long lda, ldb;
  if (sak == 1) { a_mn = 0; lda = sam; } else if (sam == 1) { a_mn = 1; lda = sak; } else return MAGIC_ERR_UNSUPPORTED;
  if (sbk == 1) { b_mn = 0; ldb = sbn; } else if (sbn == 1) { b_mn = 1; ldb = sbk; } else return MAGIC_ERR_UNSUPPORTED;
  // degenerate unit extents make both strides "1"-compatible; pick by the other stride being a valid ld
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda % 8) || (ldb % 8) || lda <= 0 || ldb <= 0)
    return MAGIC_ERR_UNSUPPORTED;
  if (a_mn == 0 && lda < K) return MAGIC_ERR_UNSUPPORTED;
  if (a_mn == 1 && lda < M) return MAGIC_ERR_UNSUPPORTED;
  if (b_mn == 0 && ldb < K) return MAGIC_ERR_UNSUPPORTED;
  if (b_mn == 1 && ldb < N) return MAGIC_ERR_UNSUPPORTED;
  const int BN = (N <= 64) ? 64 : 128;
  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = get_map(A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, BM, &ta);
  else rc = get_map(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, 64, &ta);
  if (rc) return rc;
  if (!b_mn) rc = get_map(B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, (uint32_t)BN, &tb);
  else rc = get_map(B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 64, &tb);
  if (rc) return rc;
  TcParams P;
  {
    static unsigned long long* dbg_buf = nullptr;
    static int dbg_on = -1;
    if (dbg_on < 0) {
      const char* e = getenv("MAGIC_TC_DEBUG");
      dbg_on = (e && e[0] == '1') ? 1 : 0;
      if (dbg_on) cudaMalloc(&dbg_buf, 8 * 8 * 65536);
    }
    P.dbg = dbg_on ? dbg_buf : nullptr;
    if (dbg_on) g_dbg_buf = dbg_buf;
  }
  P.M = M; P.N = N; P.K = K; P.a_mn = a_mn; P.b_mn = b_mn; P.ldc = ldc; P.epi = epi;
  {
    const int cesz = c_dt == MAGIC_BF16 ? 2 : 4;
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    bool ok = al16(C) && (ldc * cesz) % 16 == 0;
    if (epi.bias) ok = ok && al16(epi.bias);
    if (epi.pre_out) ok = ok && al16(epi.pre_out);
    if (epi.residual) ok = ok && al16(epi.residual) && (epi.res_ld * cesz) % 16 == 0;
    if (epi.dact_pre) ok = ok && al16(epi.dact_pre) && (epi.dact_ld * (epi.dact_dt == MAGIC_BF16 ? 2 : 4)) % 16 == 0;
    P.fast_epi = ok ? 1 : 0;
  }
  // split-K for skinny outputs with a long reduction (weight gradients): fp32 C, purely linear epilogue
  const int total_kb = (K + BK - 1) / BK;
  P.kb_per_split = total_kb;
  const long tiles = (long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const bool linear_epi = c_dt == MAGIC_F32 && !epi.bias && epi.act == 0 && !epi.pre_out && !epi.dact_pre &&
                          !epi.residual && epi.drop_p == 0.f && (epi.beta == 0.f || epi.beta == 1.f);
  if (linear_epi && tiles * 2 <= magic_num_sms() && total_kb >= 8) {
    int splits = (int)(magic_num_sms() / tiles);
    if (splits > total_kb / 4) splits = total_kb / 4;
    if (splits > 1) {
      P.kb_per_split = (total_kb + splits - 1) / splits;
      P.epi.atomic = 1;
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
  typedef __nv_bfloat16 bf;
  if (BN == 64) {
    if (c_dt == MAGIC_BF16) return launch_tc<64, bf>(ta, tb, C, P, st);
    return launch_tc<64, float>(ta, tb, C, P, st);
  }
  if (c_dt == MAGIC_BF16) return launch_tc<128, bf>(ta, tb, C, P, st);
  return launch_tc<128, float>(ta, tb, C, P, st);
}
