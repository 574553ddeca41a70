// ridge_launch.cu -- launch geometry of the Cholesky ridge kernels (ridge_kernels.cuh).
#include "launch.h"
#include "ridge_kernels.cuh"

namespace modl {

// ---------------------------------------------------------------------------------------
// ridge launch
// ---------------------------------------------------------------------------------------
template <typename T>
int ridge_solve(modl_ctx *ctx, const T *G, int64_t g_stride, T *Dx, T *code, const int64_t *indices,
                       T *code_batch, int64_t b, int64_t k, T alpha, cudaStream_t st)
{
    if (b <= 0) return MODL_OK;
    MODL_REQUIRE(k >= 1 && k <= 1024, "n_components must be in [1, 1024]");
    const int64_t nfac = g_stride == 0 ? 1 : b;
    T *F = nullptr;
    MODL_TRY(ws<T>(ctx, WS_CHOL, (size_t)(nfac * k * k), &F));
    int *info = static_cast<int *>(ctx->slot_ptr[WS_INFO]);
    chol_factor_kernel<T><<<(unsigned)nfac, 1024, 0, st>>>(G, g_stride, alpha, F, (int)k, info);
    MODL_LAUNCH_CHECK(ctx);
    const int tiles = (int)ceil_div(k, 32);
    const int warps = 4;
    int grid = (int)ceil_div(b, warps);
    if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
    const int64_t fs = g_stride == 0 ? 0 : k * k;
#define MODL_RS_CASE(TL) \
    chol_solve_kernel<T, TL><<<grid, warps * 32, 0, st>>>(F, fs, Dx, code, indices, code_batch, (int)b, (int)k)
    if (tiles <= 1) MODL_RS_CASE(1);
    else if (tiles <= 2) MODL_RS_CASE(2);
    else if (tiles <= 4) MODL_RS_CASE(4);
    else if (tiles <= 8) MODL_RS_CASE(8);
    else if (tiles <= 16) MODL_RS_CASE(16);
    else MODL_RS_CASE(32);
#undef MODL_RS_CASE
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}


template int ridge_solve<float>(modl_ctx *, const float *, int64_t, float *, float *, const int64_t *, float *, int64_t,
                                int64_t, float, cudaStream_t);
template int ridge_solve<double>(modl_ctx *, const double *, int64_t, double *, double *, const int64_t *, double *, int64_t,
                                 int64_t, double, cudaStream_t);

}  // namespace modl
