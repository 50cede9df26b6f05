// Error channel, version and the GEMM dispatcher (tcgen05 path vs fp32 FFMA path).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_epi.cuh"
#include "../../include/magic_b200.h"

static thread_local char g_err[512] = "";

void magic_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_pdl_override = -1;  // magic_set_pdl: -1 = follow the environment

int magic_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MAGIC_PDL");
    v = (e && e[0] == '0') ? 0 : 1;  // on by default (MAGIC_PDL=0 disables)
  }
  return g_pdl_override >= 0 ? g_pdl_override : v;
}

int gemm_simt_dispatch(const void* A, int a_dt, const void* B, int b_dt, void* C, int c_dt, int M, int N, int K,
                       long sam, long sak, long sbk, long sbn, long ldc, const GemmEpi& epi, cudaStream_t st);
// returns MAGIC_ERR_UNSUPPORTED when the operand layout is not one the tcgen05 kernel takes
int gemm_tc_dispatch(const void* A, const void* B, void* C, int c_dt, int M, int N, int K, long sam, long sak,
                     long sbk, long sbn, long ldc, const GemmEpi& epi, cudaStream_t st);
int gemm_tc_shape_ok(int M, int N, int K);

extern "C" {

const char* magic_last_error(void) { return g_err; }
int magic_version(void) { return 200; }
/* Programmatic dependent launch for the kernels launched from now on (1 / 0; -1 = follow MAGIC_PDL).  A launch
 * attribute, so inside a captured graph it sticks to the nodes captured while it was set.  PDL shortens a lone chain
 * of small kernels (-5 % on the MAGIC-S step) but an early-launched dependent grid occupies SM slots while it waits,
 * which costs more than it saves when several graph branches compete for the SMs (+4 % on the distillation step). */
int magic_set_pdl(int on) {
  g_pdl_override = on;
  return MAGIC_OK;
}
int magic_gemm_tc_supported(int M, int N, int K) { return gemm_tc_shape_ok(M, N, K); }

int magic_gemm(const void* A, int a_dt, long sam, long sak, const void* B, int b_dt, long sbk, long sbn, void* C,
               int c_dt, long ldc, int M, int N, int K, const float* bias, int act, void* pre_out,
               const void* dact_pre, int dact_dt, long dact_ld, const void* residual, long res_ld, float alpha,
               float beta, float drop_p, unsigned salt, const unsigned long long* seed_ptr, int allow_tc,
               cudaStream_t st) {
  MAGIC_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "magic_gemm: negative shape");
  if (M == 0 || N == 0) return MAGIC_OK;
  GemmEpi epi;
  epi.bias = bias;
  epi.act = act;
  epi.pre_out = pre_out;
  epi.dact_pre = dact_pre;
  epi.dact_dt = dact_dt;
  epi.dact_ld = dact_ld;
  epi.residual = residual;
  epi.res_ld = res_ld;
  epi.alpha = alpha;
  epi.beta = beta;
  epi.atomic = 0;
  epi.drop_p = drop_p;
  epi.seed_ptr = seed_ptr;
  epi.salt = salt;
  epi.rowsum = nullptr;
  if (allow_tc && a_dt == MAGIC_BF16 && b_dt == MAGIC_BF16) {
    const int rc = gemm_tc_dispatch(A, B, C, c_dt, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    if (rc != MAGIC_ERR_UNSUPPORTED) return rc;
  }
  return gemm_simt_dispatch(A, a_dt, B, b_dt, C, c_dt, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
}

/* weight + bias gradient of y = x W^T + b in ONE launch on the tensor-core path: the bias gradient is an extra
 * 16-column UMMA against a tile of ones (see gemm_tc.cu), so no separate column-sum kernel runs. */
int magic_gemm_wgrad(const void* dy, int dy_dt, long dy_ld, const void* x, int x_dt, long x_ld, float* dw, long dw_ld,
                     float* dbias, int M, int N, int K, float beta, int allow_tc, cudaStream_t st) {
  MAGIC_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "magic_gemm_wgrad: negative shape");
  if (N == 0 || K == 0 || M == 0) return MAGIC_OK;
  GemmEpi epi;
  memset(&epi, 0, sizeof(epi));
  epi.alpha = 1.f;
  epi.beta = beta;
  epi.rowsum = dbias;
  if (allow_tc && dy_dt == MAGIC_BF16 && x_dt == MAGIC_BF16) {
    // A(m=n_out, k=token) = dy[token*dy_ld + n_out]; B(k=token, n=k_in) = x[token*x_ld + k_in]
    const int rc = gemm_tc_dispatch(dy, x, dw, MAGIC_F32, N, K, M, 1, dy_ld, x_ld, 1, dw_ld, epi, st);
    if (rc != MAGIC_ERR_UNSUPPORTED) return rc;
  }
  epi.rowsum = nullptr;
  const int rc = gemm_simt_dispatch(dy, dy_dt, x, x_dt, dw, MAGIC_F32, N, K, M, 1, dy_ld, x_ld, 1, dw_ld, epi, st);
  if (rc != MAGIC_OK || dbias == nullptr) return rc;
  return magic_colsum(dy, dbias, M, N, dy_ld, dy_dt, st);
}

/* ---- measurement aid: CUDA events that can be recorded INSIDE a captured graph ---------------------------------
 * cudaEventRecordWithFlags(..., cudaEventRecordExternal) turns the record into a graph node during stream capture, so
 * every replay re-stamps the event and the host reads kernel durations of the REPLAYED graph (bench.py). */
int magic_event_create(void** ev) {
  cudaEvent_t e;
  MAGIC_CUDA(cudaEventCreate(&e), "magic_event_create");
  *ev = (void*)e;
  return MAGIC_OK;
}
int magic_event_destroy(void* ev) {
  MAGIC_CUDA(cudaEventDestroy((cudaEvent_t)ev), "magic_event_destroy");
  return MAGIC_OK;
}
int magic_event_record(void* ev, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  MAGIC_CUDA(cudaStreamIsCapturing(st, &cs), "magic_event_record");
  if (cs == cudaStreamCaptureStatusActive)
    MAGIC_CUDA(cudaEventRecordWithFlags((cudaEvent_t)ev, st, cudaEventRecordExternal), "magic_event_record(external)");
  else
    MAGIC_CUDA(cudaEventRecord((cudaEvent_t)ev, st), "magic_event_record");
  return MAGIC_OK;
}
int magic_stream_wait_event(cudaStream_t st, void* ev) {
  MAGIC_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ev, 0), "magic_stream_wait_event");
  return MAGIC_OK;
}
int magic_event_elapsed_ms(void* e0, void* e1, float* ms) {
  MAGIC_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1), "magic_event_elapsed_ms");
  return MAGIC_OK;
}

}  // extern "C"
