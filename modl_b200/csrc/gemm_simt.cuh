// gemm_simt.cuh -- register-tiled FP32/FP64 GEMM on the CUDA cores, used for every dense
// contraction of the step in full precision (TF32 inputs would flip CD stop
// decisions) and as the f64 path.  Deterministic: split-K partials are reduced in a fixed
// order by a second pass, never with atomics, so two fits with the same seed are bitwise
// identical (reference test: test_dict_fact.py:90-112).
//
//   C[M x N] = alpha * op(A) . op(B) + beta * C
//   A_KMAJOR : A stored M x K row-major (lda)      A_MMAJOR : A stored K x M row-major
//   B_KMAJOR : B stored N x K row-major (ldb)      B_NMAJOR : B stored K x N row-major
//
// Call sites (reference lines are what each product replaces):
//   G  = r D_sub D_sub^T,  Dx = r X_sub D_sub^T      KMAJOR x KMAJOR   dict_fact.py:595,604
//   C_ = (1-w) C_ + w/b code^T code                   MMAJOR x NMAJOR   dict_fact.py:573
//   B_ = (1-w) B_ + w/b code^T X                      MMAJOR x NMAJOR   dict_fact.py:564
//   G_ -/+= D_sub D_sub^T                             KMAJOR x KMAJOR   dict_fact.py:668,713
#pragma once
#include "common.cuh"
#include "launch.h"

namespace modl {

constexpr int GEMM_BM = 64, GEMM_BN = 64, GEMM_BK = 16, GEMM_THREADS = 256;

template <typename T, int LA, int LB>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_simt_kernel(int M, int N, int K, int k_chunk, T alpha, const T *__restrict__ A, int64_t lda,
                 const T *__restrict__ B, int64_t ldb, T beta, T *__restrict__ C, int64_t ldc,
                 T *__restrict__ part)
{
    __shared__ T As[GEMM_BK][GEMM_BM + 4];
    __shared__ T Bs[GEMM_BK][GEMM_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
    const int kz0 = blockIdx.z * k_chunk;
    const int kz1 = min(K, kz0 + k_chunk);
    const int ty = tid >> 4, tx = tid & 15;

    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

    for (int k0 = kz0; k0 < kz1; k0 += GEMM_BK) {
        // ---- stage the A and B tiles (zero-filled outside the matrix / K range) ----
#pragma unroll
        for (int e = tid; e < GEMM_BM * GEMM_BK; e += GEMM_THREADS) {
            int mm, kk;
            if (LA == A_KMAJOR) { kk = e % GEMM_BK; mm = e / GEMM_BK; }
            else                { mm = e % GEMM_BM; kk = e / GEMM_BM; }
            const int gm = m0 + mm, gk = k0 + kk;
            T v = T(0);
            if (gm < M && gk < kz1)
                v = (LA == A_KMAJOR) ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
            As[kk][mm] = v;
        }
#pragma unroll
        for (int e = tid; e < GEMM_BN * GEMM_BK; e += GEMM_THREADS) {
            int nn, kk;
            if (LB == B_KMAJOR) { kk = e % GEMM_BK; nn = e / GEMM_BK; }
            else                { nn = e % GEMM_BN; kk = e / GEMM_BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            T v = T(0);
            if (gn < N && gk < kz1)
                v = (LB == B_KMAJOR) ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    if (part != nullptr) {
        T *dst = part + (int64_t)blockIdx.z * M * N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = m0 + ty * 4 + i;
            if (gm >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + tx * 4 + j;
                if (gn < N) dst[(int64_t)gm * N + gn] = acc[i][j];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = m0 + ty * 4 + i;
            if (gm >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + tx * 4 + j;
                if (gn >= N) continue;
                T *c = C + (int64_t)gm * ldc + gn;
                T r = alpha * acc[i][j];
                if (beta != T(0)) r = fma(beta, *c, r);
                *c = r;
            }
        }
    }
}

// C = alpha * sum_z part[z] + beta * C   (fixed summation order over z)
template <typename T>
__global__ void gemm_splitk_reduce_kernel(int M, int N, int splits, T alpha, const T *__restrict__ part,
                                          T beta, T *__restrict__ C, int64_t ldc)
{
    const int64_t total = (int64_t)M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        T s = T(0);
        for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + e];
        const int m = (int)(e / N), n = (int)(e % N);
        T *c = C + (int64_t)m * ldc + n;
        T r = alpha * s;
        if (beta != T(0)) r = fma(beta, *c, r);
        *c = r;
    }
}

template <typename T>
int gemm_simt(modl_ctx *ctx, int la, int lb, int64_t M, int64_t N, int64_t K, T alpha, const T *A,
              int64_t lda, const T *B, int64_t ldb, T beta, T *C, int64_t ldc, cudaStream_t st, WsSlot part_slot)
{
    if (M <= 0 || N <= 0) return MODL_OK;
    const int64_t tiles = ceil_div(M, GEMM_BM) * ceil_div(N, GEMM_BN);
    int64_t splits = 1;
    if (K > 0) {
        const int64_t want = ceil_div(2 * (int64_t)ctx->sm_count, tiles);
        const int64_t cap = K / (4 * GEMM_BK) > 0 ? K / (4 * GEMM_BK) : 1;
        splits = want < cap ? want : cap;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
    }
    int64_t k_chunk = round_up(ceil_div(K > 0 ? K : 1, splits), GEMM_BK);
    splits = K > 0 ? ceil_div(K, k_chunk) : 1;
    T *part = nullptr;
    if (splits > 1) MODL_TRY(ws<T>(ctx, part_slot, (size_t)(splits * M * N), &part));
    dim3 grid((unsigned)ceil_div(N, GEMM_BN), (unsigned)ceil_div(M, GEMM_BM), (unsigned)splits);
#define MODL_GEMM_LAUNCH(LA_, LB_)                                                              \
    gemm_simt_kernel<T, LA_, LB_><<<grid, GEMM_THREADS, 0, st>>>(                               \
        (int)M, (int)N, (int)K, (int)k_chunk, alpha, A, lda, B, ldb, beta, C, ldc, part)
    if (la == A_KMAJOR && lb == B_KMAJOR) MODL_GEMM_LAUNCH(A_KMAJOR, B_KMAJOR);
    else if (la == A_MMAJOR && lb == B_NMAJOR) MODL_GEMM_LAUNCH(A_MMAJOR, B_NMAJOR);
    else if (la == A_KMAJOR && lb == B_NMAJOR) MODL_GEMM_LAUNCH(A_KMAJOR, B_NMAJOR);
    else MODL_GEMM_LAUNCH(A_MMAJOR, B_KMAJOR);
#undef MODL_GEMM_LAUNCH
    MODL_LAUNCH_CHECK(ctx);
    if (splits > 1) {
        const int64_t total = M * N;
        int blocks = (int)(ceil_div(total, 256) < 4 * ctx->sm_count ? ceil_div(total, 256) : 4 * ctx->sm_count);
        gemm_splitk_reduce_kernel<T><<<blocks, 256, 0, st>>>((int)M, (int)N, (int)splits, alpha, part, beta, C, ldc);
        MODL_LAUNCH_CHECK(ctx);
    }
    return MODL_OK;
}

}  // namespace modl
