// ridge_kernels.cuh -- the l1_ratio == 0 branch of the code solve: (G + alpha I) code_i = Dx_i
// by Cholesky, replacing LAPACK posv as called by the reference
// [ref: modl/decomposition/dict_fact_fast.pyx:174-197 (shared Gram, one factorisation, b
// right-hand sides) and :82-94 (one factorisation per sample)].
//
//   chol_factor_kernel : one CTA per matrix.  Builds F = mirror(U) with A + alpha I = U^T U,
//                        i.e. F[i][j] = U[min(i,j)][max(i,j)], so that both triangular solves
//                        read contiguous rows.  Non-positive pivots raise *info (LAPACK info>0;
//                        the reference ignores it, we surface it through modl_ctx_check_info).
//   chol_solve_kernel  : one warp per right-hand side, x distributed over the lanes; forward
//                        substitution with rows of U^T... in axpy form, then backward.
#pragma once
#include "common.cuh"

namespace modl {

// A: k x k symmetric (row-major), stride a_stride between matrices; F: k x k per matrix.
template <typename T>
__global__ void __launch_bounds__(1024)
chol_factor_kernel(const T *__restrict__ A, int64_t a_stride, T alpha, T *__restrict__ F, int k, int *info)
{
    const T *Am = A + (int64_t)blockIdx.x * a_stride;
    T *Fm = F + (int64_t)blockIdx.x * (int64_t)k * k;
    const int tid = threadIdx.x, nt = blockDim.x;
    // work copy: upper triangle (j >= i) of A + alpha I
    for (int e = tid; e < k * k; e += nt) {
        const int i = e / k, j = e % k;
        T v = Am[e];
        if (i == j) v += alpha;
        Fm[e] = v;
    }
    __syncthreads();
    __shared__ T pivot;
    for (int j = 0; j < k; ++j) {
        if (tid == 0) {
            T d = Fm[(int64_t)j * k + j];
            if (!(d > T(0))) { atomicMax(info, j + 1); d = T(1); }
            d = t_sqrt(d);
            Fm[(int64_t)j * k + j] = d;
            pivot = d;
        }
        __syncthreads();
        const T d = pivot;
        // scale row j of U right of the diagonal
        for (int c = j + 1 + tid; c < k; c += nt) Fm[(int64_t)j * k + c] = Fm[(int64_t)j * k + c] / d;
        __syncthreads();
        // trailing update of the upper triangle: U[i][c] -= U[j][i] * U[j][c], j < i <= c
        const int m = k - j - 1;
        for (int e = tid; e < m * m; e += nt) {
            const int i = j + 1 + e / m, c = j + 1 + e % m;
            if (c >= i) Fm[(int64_t)i * k + c] = fma(-Fm[(int64_t)j * k + i], Fm[(int64_t)j * k + c], Fm[(int64_t)i * k + c]);
        }
        __syncthreads();
    }
    // mirror the factor into the lower triangle
    for (int e = tid; e < k * k; e += nt) {
        const int i = e / k, j = e % k;
        if (j < i) Fm[e] = Fm[(int64_t)j * k + i];
    }
}

// Solve U^T U x = rhs for every row of `rhs` (b x k, in place -> x), one warp per row.
// F is the mirrored factor (f_stride elements between per-sample factors, 0 = shared).
// The solution is also scattered to code[indices[ii]] and code_batch[ii].
template <typename T, int TILES>
__global__ void chol_solve_kernel(const T *__restrict__ F, int64_t f_stride, T *__restrict__ rhs,
                                  T *__restrict__ code, const int64_t *__restrict__ indices,
                                  T *__restrict__ code_batch, int b, int k)
{
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    for (int ii = blockIdx.x * warps + (threadIdx.x >> 5); ii < b; ii += gridDim.x * warps) {
        const T *Fm = F + (int64_t)ii * f_stride;
        T x[TILES];
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const int c = J * 32 + lane;
            x[J] = c < k ? rhs[(int64_t)ii * k + c] : T(0);
        }
        // forward: U^T z = rhs.  z_j = x_j / U[j][j]; x[c] -= U[j][c] z_j for c > j.
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            for (int l = 0; l < 32; ++l) {
                const int j = J * 32 + l;
                if (j >= k) break;
                const T *row = Fm + (int64_t)j * k;
                T r[TILES];
#pragma unroll
                for (int JJ = 0; JJ < TILES; ++JJ) {
                    const int c = JJ * 32 + lane;
                    r[JJ] = (JJ >= J && c < k) ? row[c] : T(0);
                }
                const T zj = __shfl_sync(kFullMask, x[J] / r[J], l);   // lane l holds U[j][j] in r[J]
                if (lane == l) x[J] = zj;
#pragma unroll
                for (int JJ = 0; JJ < TILES; ++JJ) {
                    const int c = JJ * 32 + lane;
                    if (JJ >= J && c > j) x[JJ] = fma(-zj, r[JJ], x[JJ]);
                }
            }
        }
        // backward: U x = z.  x_j = z_j / U[j][j]; z[i] -= U[i][j] x_j for i < j
        // (U[i][j] for i < j is F[j][i], the mirrored lower part of row j).
#pragma unroll
        for (int J = TILES - 1; J >= 0; --J) {
            for (int l = 31; l >= 0; --l) {
                const int j = J * 32 + l;
                if (j >= k) continue;
                const T *row = Fm + (int64_t)j * k;
                T r[TILES];
#pragma unroll
                for (int JJ = 0; JJ < TILES; ++JJ) {
                    const int c = JJ * 32 + lane;
                    r[JJ] = (JJ <= J && c < k) ? row[c] : T(0);
                }
                const T xj = __shfl_sync(kFullMask, x[J] / r[J], l);
                if (lane == l) x[J] = xj;
#pragma unroll
                for (int JJ = 0; JJ < TILES; ++JJ) {
                    const int c = JJ * 32 + lane;
                    if (JJ <= J && c < j) x[JJ] = fma(-xj, r[JJ], x[JJ]);
                }
            }
        }
        const int64_t rowi = indices ? indices[ii] : (int64_t)ii;
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const int c = J * 32 + lane;
            if (c < k) {
                rhs[(int64_t)ii * k + c] = x[J];      // the reference leaves the solution in Dx too
                code[rowi * k + c] = x[J];
                if (code_batch) code_batch[(int64_t)ii * k + c] = x[J];
            }
        }
    }
}

}  // namespace modl
