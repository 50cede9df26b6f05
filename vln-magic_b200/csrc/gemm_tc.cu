// tcgen05 / TMA GEMM (placeholder until the kernel lands in this file).
#include "common.cuh"
#include "gemm_epi.cuh"

int gemm_tc_shape_ok(int, int, int) { return 0; }
int gemm_tc_dispatch(const void*, const void*, void*, int, int, int, int, long, long, long, long, long,
                     const GemmEpi&, cudaStream_t) {
  return MAGIC_ERR_UNSUPPORTED;
}
