// cd_launch.cu -- launch geometry of the coordinate-descent kernels (cd_kernels.cuh).
#include "cd_kernels.cuh"
#include "launch.h"

namespace modl {

// ---------------------------------------------------------------------------------------
// coordinate descent launch
// ---------------------------------------------------------------------------------------
template <typename T, int TILES, bool PACKED>
static int cd_launch_inst(modl_ctx *ctx, const T *G, int64_t g_stride, const T *Dx, const T *xnorm2, T *code,
                          const int64_t *indices, T *code_batch, int64_t b, int64_t k, T alpha, T beta, T tol,
                          int max_iter, int positive, int32_t *sweeps, cudaStream_t st)
{
    auto kern = cd_regression_kernel<T, TILES, PACKED>;
    size_t smem = PACKED ? cd_packed_elems(TILES) * sizeof(T) : 0;
    int warps, grid;
    if (PACKED) {
        warps = ctx->opt_cd_warps > 0 ? ctx->opt_cd_warps : (int)ceil_div(b, ctx->sm_count);
        if (warps < 4) warps = 4;
        if (warps > 16) warps = 16;
        grid = (int)ceil_div(b, warps);
        if (grid > ctx->sm_count) grid = ctx->sm_count;
        static bool configured[64] = {};    // per instantiation and device; the attribute is per function, not per call
        if (ctx->device < 0 || ctx->device >= 64 || !configured[ctx->device]) {
            MODL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (ctx->device >= 0 && ctx->device < 64) configured[ctx->device] = true;
        }
        // tile-packed lower triangle of G in global memory, pulled by TMA bulk copies
        T *packed = nullptr;
        MODL_TRY(ws<T>(ctx, WS_GPACK, cd_packed_elems(TILES) + (size_t)TILES * CD_TILE, &packed));    // image | H of the all-ones start
        cd_pack_gram_kernel<T><<<cd_tri(TILES) + 1, 256, 0, st>>>(G, (int)k, TILES, packed);
        MODL_LAUNCH_CHECK(ctx);
        G = packed;
    } else {
        warps = ctx->opt_cd_warps > 0 ? ctx->opt_cd_warps : 4;
        if (warps > 8) warps = 8;
        grid = (int)ceil_div(b, warps);
        if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
    }
    kern<<<grid, warps * 32, smem, st>>>(G, g_stride, Dx, xnorm2, code, indices, code_batch, (int)b, (int)k,
                                         alpha, beta, tol, max_iter, positive, sweeps);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
int cd_launch(modl_ctx *ctx, const T *G, int64_t g_stride, const T *Dx, const T *xnorm2, T *code,
                     const int64_t *indices, T *code_batch, int64_t b, int64_t k, T alpha, T beta, T tol,
                     int max_iter, int positive, int32_t *sweeps, cudaStream_t st)
{
    if (b <= 0) return MODL_OK;
    MODL_REQUIRE(k >= 1 && k <= 1024, "n_components must be in [1, 1024]");
    const int tiles = (int)ceil_div(k, 32);
    const bool packed = g_stride == 0 && !ctx->opt_force_global_gram && tiles <= 10 &&
                        cd_packed_elems(tiles) * sizeof(T) + 1024 <= (size_t)ctx->max_smem_optin;
#define MODL_CD_CASE(TL, PK)                                                                              \
    return cd_launch_inst<T, TL, PK>(ctx, G, g_stride, Dx, xnorm2, code, indices, code_batch, b, k, alpha, \
                                     beta, tol, max_iter, positive, sweeps, st)
    if (packed) {
        switch (tiles) {
            case 1: MODL_CD_CASE(1, true);
            case 2: MODL_CD_CASE(2, true);
            case 3: MODL_CD_CASE(3, true);
            case 4: MODL_CD_CASE(4, true);
            case 5: MODL_CD_CASE(5, true);
            case 6: MODL_CD_CASE(6, true);
            case 7: MODL_CD_CASE(7, true);
            case 8: MODL_CD_CASE(8, true);
            case 9: MODL_CD_CASE(9, true);
            default: MODL_CD_CASE(10, true);
        }
    }
    if (tiles <= 1) MODL_CD_CASE(1, false);
    if (tiles <= 2) MODL_CD_CASE(2, false);
    if (tiles <= 4) MODL_CD_CASE(4, false);
    if (tiles <= 8) MODL_CD_CASE(8, false);
    if (tiles <= 16) MODL_CD_CASE(16, false);
    MODL_CD_CASE(32, false);
#undef MODL_CD_CASE
}


template int cd_launch<float>(modl_ctx *, const float *, int64_t, const float *, const float *, float *, const int64_t *,
                              float *, int64_t, int64_t, float, float, float, int, int, int32_t *, cudaStream_t);
template int cd_launch<double>(modl_ctx *, const double *, int64_t, const double *, const double *, double *, const int64_t *,
                               double *, int64_t, int64_t, double, double, double, int, int, int32_t *, cudaStream_t);

}  // namespace modl
