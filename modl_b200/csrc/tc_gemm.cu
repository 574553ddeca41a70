// tc_gemm.cu -- launchers of the tcgen05 3xTF32 GEMM and its pack kernels (tc_gemm.cuh).
#include "gemm_simt.cuh"
#include "launch.h"
#include "tc_gemm.cuh"

namespace modl {

// split-K reduction of the two-destination form: columns [0, N1) of the N-wide product go to C, columns
// [n_split, n_split + N2) to C2; fixed summation order over the slices
__global__ void tc_splitk_reduce2_kernel(int M, int N, int splits, float alpha, const float *__restrict__ part, float beta,
                                         float *__restrict__ C, int64_t ldc, int N1, float *__restrict__ C2, int64_t ldc2,
                                         int n_split, int N2)
{
    const int64_t total = (int64_t)M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(e / N), n = (int)(e % N);
        float *c = nullptr;
        if (n < N1) c = C + (int64_t)m * ldc + n;
        else if (n >= n_split && n < n_split + N2) c = C2 + (int64_t)m * ldc2 + (n - n_split);
        if (c == nullptr) continue;
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + e];
        float r = alpha * s;
        if (beta != 0.f) r = fmaf(beta, *c, r);
        *c = r;
    }
}

size_t tc_packed_elems(int64_t rows, int64_t kd, int64_t rpb) { return tc_packed_floats(rows, kd, rpb); }

// Accumulator tile width for an N-column output produced by `mtiles` row tiles: the multiple of 16
// in [128, 160] whose tile count wastes the least of the last wave over the SMs (one CTA per SM).
// p = 10000, M = 256: bn = 144 -> 70 x 2 = 140 CTAs in ONE wave instead of 158 in two; with sm_avail = 132 (the
// dictionary update's 16-CTA cluster is resident on the others) bn = 160 -> 126 CTAs.
int tc_pick_bn(const modl_ctx *ctx, int64_t N, int64_t mtiles, int sm_avail)
{
    int best = TC_ROWS;
    double best_cost = 1e300;
    const int64_t sms = sm_avail > 0 && sm_avail < ctx->sm_count ? sm_avail : ctx->sm_count;
    for (int bn = 128; bn <= 160; bn += 16) {
        const int64_t ctas = ceil_div(N, bn) * mtiles;
        const int64_t waves = ceil_div(ctas, sms);
        const double cost = (double)waves * bn;            // time ~ waves x tile width
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}
int64_t tc_rows_padded(int64_t rows) { return tc_row_blocks(rows) * TC_ROWS; }

int tc_pack_rows(modl_ctx *ctx, const float *src, int64_t ld, int64_t rows, int64_t p, const int64_t *subset,
                 int64_t kd, float *packed, int64_t row0, int64_t rows_pad_end, float *plain, int64_t ldp,
                 float *norm2, cudaStream_t st)
{
    const int64_t n = rows_pad_end - row0;
    if (n <= 0 || kd <= 0) return MODL_OK;
    tc_pack_rows_kernel<<<grid_for(ctx, n, 16), 256, 0, st>>>(src, ld, (int)rows, (int)p, subset, (int)kd, packed, row0,
                                                               rows_pad_end, plain, ldp, norm2);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

int tc_pack_cols(modl_ctx *ctx, const float *src, int64_t ld, int64_t kd, int64_t rows, float *packed, int rpb,
                 cudaStream_t st)
{
    if (rows <= 0 || kd <= 0) return MODL_OK;
    const int64_t total = tc_row_blocks(rows, rpb) * rpb * tc_k_blocks(kd) * TC_CHUNKS;
    tc_pack_cols_kernel<<<grid_for(ctx, ceil_div(total, 256), 16), 256, 0, st>>>(src, ld, (int)kd, (int)rows, packed, rpb);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

int tc_gemm(modl_ctx *ctx, const float *Apacked, const float *Bpacked, int64_t M, int64_t N, int64_t Kd, float alpha,
            float beta, float *C, int64_t ldc, int bn, cudaStream_t st, WsSlot part_slot, float *C2, int64_t ldc2,
            int64_t n_split, int64_t N1, const float *Braw, int64_t ldb_raw)
{
    if (M <= 0 || N <= 0) return MODL_OK;
    MODL_REQUIRE(Kd >= 1, "tc_gemm needs a non-empty contraction");
    MODL_REQUIRE(bn >= 16 && bn <= TC_MAX_BN && bn % 16 == 0, "tc_gemm tile width");
    TcGemmParams P;
    P.A = Apacked; P.B = Bpacked; P.C = C; P.ldc = ldc; P.M = (int)M; P.N = (int)N;
    MODL_REQUIRE((Bpacked != nullptr) != (Braw != nullptr), "tc_gemm takes the B operand packed or raw");
    P.Braw = Braw; P.ldb_raw = ldb_raw; P.Kd = (int)Kd;
    MODL_REQUIRE(Braw == nullptr || bn <= 160, "tc_gemm: the raw-B form converts tiles of at most 160 columns");
    P.nkb = (int)tc_k_blocks(Kd);
    P.alpha = alpha; P.beta = beta;
    P.sbo = 128u;
    P.bn = bn;
    P.C2 = C2; P.ldc2 = ldc2;
    P.n_split = C2 ? (int)n_split : 0;
    P.N1 = C2 ? (int)N1 : (int)N;
    P.N2 = C2 ? (int)(N - n_split) : 0;
    MODL_REQUIRE(C2 == nullptr || (n_split % bn == 0 && N1 <= n_split && n_split < N), "tc_gemm second destination");
    P.nacc = 512 / bn < 4 ? 512 / bn : 4;
    P.tmem_cols = 32;
    while ((int)P.tmem_cols < P.nacc * bn) P.tmem_cols <<= 1;
    const size_t stage_bytes = TC_BLOCK_BYTES + (size_t)2 * bn * TC_BK * 4;
    P.stages = (int)((TC_SMEM_BUDGET - 1024) / stage_bytes);
    if (P.stages > TC_MAX_STAGES) P.stages = TC_MAX_STAGES;
    MODL_REQUIRE(P.stages >= 2, "tc_gemm tile does not fit shared memory");
    const size_t smem_bytes = (size_t)P.stages * stage_bytes + 1024;
    static bool configured[64] = {};        // per device: function attributes are per device
    if (ctx->device < 0 || ctx->device >= 64 || !configured[ctx->device]) {
        MODL_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BUDGET));
        if (ctx->device >= 0 && ctx->device < 64) configured[ctx->device] = true;
    }
    const int64_t tiles = tc_row_blocks(M) * tc_row_blocks(N, bn);
    // split the contraction until the grid covers the SMs (one CTA per SM: 193 KB of shared memory each)
    int64_t splits = (int64_t)ctx->sm_count / tiles;            // never more CTAs than SMs: a second wave costs a whole tile time
    if (splits > P.nkb) splits = P.nkb;
    if (splits > 32) splits = 32;
    if (splits < 1 || (C2 != nullptr && !ctx->opt_tc_split2)) splits = 1;
    P.kb_per_split = (int)ceil_div(P.nkb, splits);
    splits = ceil_div(P.nkb, P.kb_per_split);
    P.part = nullptr;
    if (splits > 1) MODL_TRY(ws<float>(ctx, part_slot, (size_t)(splits * M * N), &P.part));
    dim3 grid((unsigned)tc_row_blocks(N, bn), (unsigned)tc_row_blocks(M), (unsigned)splits);
    tc_gemm_kernel<<<grid, TC_THREADS, smem_bytes, st>>>(P);
    MODL_LAUNCH_CHECK(ctx);
    if (splits > 1) {
        const int64_t total = M * N;
        int blocks = (int)(ceil_div(total, 256) < 4 * ctx->sm_count ? ceil_div(total, 256) : 4 * ctx->sm_count);
        if (C2 == nullptr)
            gemm_splitk_reduce_kernel<float><<<blocks, 256, 0, st>>>((int)M, (int)N, (int)splits, alpha, P.part, beta, C, ldc);
        else
            tc_splitk_reduce2_kernel<<<blocks, 256, 0, st>>>((int)M, (int)N, (int)splits, alpha, P.part, beta, C, ldc, P.N1, C2, ldc2,
                                                             P.n_split, P.N2);
        MODL_LAUNCH_CHECK(ctx);
    }
    return MODL_OK;
}

}  // namespace modl
