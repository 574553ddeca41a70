// gemm_simt.cu -- instantiates the CUDA-core GEMM (gemm_simt.cuh) for float and double.
#include "gemm_simt.cuh"

namespace modl {

template int gemm_simt<float>(modl_ctx *, int, int, int64_t, int64_t, int64_t, float, const float *, int64_t, const float *,
                              int64_t, float, float *, int64_t, cudaStream_t, WsSlot);
template int gemm_simt<double>(modl_ctx *, int, int, int64_t, int64_t, int64_t, double, const double *, int64_t,
                               const double *, int64_t, double, double *, int64_t, cudaStream_t, WsSlot);

}  // namespace modl
