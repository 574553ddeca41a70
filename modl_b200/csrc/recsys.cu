// recsys.cu -- the missing-value path of RecsysDictFact:
// kernels + C-ABI entry points (include/modl_b200.h, "Recsys" block).
//
// The reference walks the rows of a CSR matrix one by one in Python
// [ref: modl/decomposition/recsys.py:147-213, 254-265].  Within one minibatch the dictionary is
// constant, so the per-row code solves are independent (one CTA per row here), and the B_ update
// -- sequential over rows in the reference -- is a per-COLUMN recurrence (only rows that observe
// column j touch B_[:, j] and feature_n_iter_[j]), hence parallel over the union of observed
// columns once the batch's entries are ordered by (column, position in the batch).
//
// Layout: the CSR arrays stay resident in HBM; the dictionary is kept twice, D (k x p, what the
// shared dictionary-update kernels work on) and Dt (p x k), so that gathering the atoms' values at
// an observed column is ONE contiguous k-vector instead of k sector-sized reads p elements apart.
// Everything here is HBM/gather bound (k <= 128, tens of flops per byte at most): CUDA cores,
// coalesced k-vectors, shared-memory staging of the gathered tile -- no tensor cores.
#include <type_traits>

#include "basic_kernels.cuh"
#include "common.cuh"
#include "launch.h"

namespace modl {

constexpr int RS_TILE_E = 16;     // observed entries staged per step of the Gram kernel
constexpr int RS_KC = 8;          // atoms per thread of the B_ recurrence

// ---------------------------------------------------------------------------------------
// One CTA per row of the batch:  G = D_sub D_sub^T + (alpha / reduction) I,  Dx = D_sub x_sub,
// D_sub = components_[:, observed columns], reduction = n_features / len(observed)
// [ref: recsys.py:169-180 and :254-265].  A thread owns a 4 x 4 block of G (MAXB blocks when
// k > 64); the gathered Dt rows of RS_TILE_E entries are staged in shared memory, zero padded.
// ---------------------------------------------------------------------------------------
template <typename T, int MAXB>
__global__ void __launch_bounds__(256)
recsys_gram_dx_kernel(const T *__restrict__ Dt, int64_t ldt, const int64_t *__restrict__ indptr,
                      const int32_t *__restrict__ indices, const T *__restrict__ data,
                      const int64_t *__restrict__ rows, int64_t row0, int k, int64_t p, double alpha,
                      T *__restrict__ G, T *__restrict__ Dx)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    T *tile = reinterpret_cast<T *>(rs_smem);
    const int kp = (k + 3) & ~3, nb = kp >> 2;
    T *xs = tile + RS_TILE_E * kp;
    const int tid = threadIdx.x;
    const int64_t r = rows ? rows[blockIdx.x] : row0 + (int64_t)blockIdx.x;
    const int64_t lo = indptr[r], hi = indptr[r + 1];

    T acc[MAXB][16];
#pragma unroll
    for (int m = 0; m < MAXB; ++m)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[m][i] = T(0);
    T dx = T(0);

    for (int64_t e0 = lo; e0 < hi; e0 += RS_TILE_E) {
        const int ne = (int)((hi - e0) < (int64_t)RS_TILE_E ? (hi - e0) : (int64_t)RS_TILE_E);
        for (int i = tid; i < RS_TILE_E * kp; i += 256) {
            const int e = i / kp, a = i - e * kp;
            T v = T(0);
            if (e < ne && a < k) v = Dt[(int64_t)indices[e0 + e] * ldt + a];
            tile[i] = v;
        }
        if (tid < RS_TILE_E) xs[tid] = tid < ne ? data[e0 + tid] : T(0);
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MAXB; ++m) {
            const int blk = tid + m * 256;
            if (blk < nb * nb) {
                const int bi = blk / nb, bj = blk - bi * nb;
#pragma unroll 4
                for (int e = 0; e < RS_TILE_E; ++e) {
                    const T *row = tile + e * kp;
                    T a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { a[i] = row[4 * bi + i]; b[i] = row[4 * bj + i]; }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[m][i * 4 + j] = fma(a[i], b[j], acc[m][i * 4 + j]);
                }
            }
        }
        if (tid < k) {
#pragma unroll 4
            for (int e = 0; e < RS_TILE_E; ++e) dx = fma(tile[e * kp + tid], xs[e], dx);
        }
        __syncthreads();
    }

    // alpha / reduction, reduction = n_features / len_subset  [ref: :174, :177]; an empty row never reaches
    // the solver in the reference (:167); it gets the identity here so that the factorisation stays defined
    const T ridge = hi > lo ? (T)(alpha / ((double)p / (double)(hi - lo))) : T(1);
    T *Gr = G + (int64_t)blockIdx.x * k * k;
#pragma unroll
    for (int m = 0; m < MAXB; ++m) {
        const int blk = tid + m * 256;
        if (blk < nb * nb) {
            const int bi = blk / nb, bj = blk - bi * nb;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int a = 4 * bi + i, b = 4 * bj + j;
                    if (a < k && b < k) Gr[(int64_t)a * k + b] = acc[m][i * 4 + j] + (a == b ? ridge : T(0));
                }
        }
    }
    if (tid < k) Dx[(int64_t)blockIdx.x * k + tid] = dx;
}

// ---------------------------------------------------------------------------------------
// B_ recurrence of one minibatch, one thread per (observed column, RS_KC atoms)
// [ref: recsys.py:168, :182-185]:  for every row of the batch that observes column j, in batch order,
//     feature_n_iter_[j] += 1;  w_B = min(1, w n_iter_ / feature_n_iter_[j])
//     B_[:, j] = (1 - w_B) B_[:, j] + code_[row] (x_rowj w_B)
// The reference evaluates w_B and both products in float64 and rounds B_ to its dtype after the scaling and
// again after the addition; so does this kernel.  Threads of a warp hold consecutive entries of the sorted
// column list, so their B_ accesses fall in neighbouring sectors of each atom's row.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
recsys_update_B_kernel(T *__restrict__ B, int64_t ldb, const T *__restrict__ code, const int64_t *__restrict__ subset,
                       const int64_t *__restrict__ col_ptr, const int64_t *__restrict__ entry_row,
                       const T *__restrict__ entry_val, const int64_t *__restrict__ feature_n_iter, int64_t s, int k,
                       double w_n_iter)
{
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= s) return;
    const int k0 = blockIdx.y * RS_KC;
    const int64_t col = subset[j];
    const int64_t lo = col_ptr[j], hi = col_ptr[j + 1];
    int64_t cnt = feature_n_iter[col];
    T b[RS_KC];
#pragma unroll
    for (int c = 0; c < RS_KC; ++c) b[c] = (k0 + c < k) ? B[(int64_t)(k0 + c) * ldb + col] : T(0);
    for (int64_t e = lo; e < hi; ++e) {
        cnt += 1;
        const double wB = fmin(1.0, w_n_iter / (double)cnt);
        const double keep = 1.0 - wB;
        const double xw = (double)entry_val[e] * wB;
        const T *cr = code + entry_row[e] * (int64_t)k;
#pragma unroll
        for (int c = 0; c < RS_KC; ++c) {
            if (k0 + c < k) {
                const T scaled = (T)__dmul_rn((double)b[c], keep);
                b[c] = (T)__dadd_rn((double)scaled, __dmul_rn((double)cr[k0 + c], xw));
            }
        }
    }
#pragma unroll
    for (int c = 0; c < RS_KC; ++c)
        if (k0 + c < k) B[(int64_t)(k0 + c) * ldb + col] = b[c];
}

// feature_n_iter_[subset[j]] += number of rows of the batch that observe it  [ref: recsys.py:168]
__global__ void recsys_count_kernel(int64_t *__restrict__ feature_n_iter, const int64_t *__restrict__ subset,
                                    const int64_t *__restrict__ col_ptr, int64_t s)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < s; j += (int64_t)gridDim.x * blockDim.x)
        feature_n_iter[subset[j]] += col_ptr[j + 1] - col_ptr[j];
}

// ---------------------------------------------------------------------------------------
// Dt[subset[j], :] = D[:, subset[j]]  (subset NULL: every column) through a 32 x 32 shared tile: reads run
// along the (sorted) columns, writes along the atoms.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
recsys_transpose_kernel(const T *__restrict__ D, int64_t ldd, T *__restrict__ Dt, int64_t ldt,
                        const int64_t *__restrict__ subset, int64_t s, int k)
{
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    const int a0 = blockIdx.y * 32;
    {
        const int64_t j = j0 + tx;
        const int64_t col = j < s ? (subset ? subset[j] : j) : -1;
        for (int a = ty; a < 32; a += 8)
            tile[a][tx] = (col >= 0 && a0 + a < k) ? D[(int64_t)(a0 + a) * ldd + col] : T(0);
    }
    __syncthreads();
    for (int jj = ty; jj < 32; jj += 8) {
        const int64_t j = j0 + jj;
        if (j < s && a0 + tx < k) {
            const int64_t col = subset ? subset[j] : j;
            Dt[col * ldt + a0 + tx] = tile[tx][jj];
        }
    }
}

// ---------------------------------------------------------------------------------------
// out[e] = <code[u], components[:, indices[e]]> for every stored entry e of row u
// [ref: _predict, recsys_fast.pyx:10-37]: float64 accumulation in atom order, multiply and add rounded
// separately (the reference is plain C without contraction), so float64 estimators reproduce it bit for bit.
// One warp per row, lanes over its entries.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
recsys_predict_kernel(const T *__restrict__ code, const T *__restrict__ Dt, int64_t ldt,
                      const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, int64_t n_rows, int k,
                      double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < n_rows; u += nwarps) {
        const T *cu = code + u * (int64_t)k;
        const int64_t lo = indptr[u], hi = indptr[u + 1];
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const T *q = Dt + (int64_t)indices[e] * ldt;
            double dot = 0.0;
            for (int a = 0; a < k; ++a) dot = __dadd_rn(dot, __dmul_rn((double)cu[a], (double)q[a]));
            out[e] = dot;
        }
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
template <typename T>
static int recsys_gram_dx(modl_ctx *ctx, const T *Dt, int64_t ldt, const int64_t *indptr, const int32_t *indices,
                          const T *data, const int64_t *rows, int64_t row0, int64_t b, int64_t k, int64_t p, double alpha,
                          T *G, T *Dx, cudaStream_t st)
{
    MODL_REQUIRE(ctx && Dt && indptr && indices && data && G && Dx && b >= 0 && p >= 1 && ldt >= k, "recsys_gram_dx arguments");
    MODL_REQUIRE(k >= 1 && k <= 128, "the recsys path supports 1 <= n_components <= 128");
    if (b == 0) return MODL_OK;
    const int kp = (int)round_up(k, 4);
    const size_t smem = sizeof(T) * (size_t)(RS_TILE_E * kp + RS_TILE_E);
    if (k <= 64)
        recsys_gram_dx_kernel<T, 1><<<(unsigned)b, 256, smem, st>>>(Dt, ldt, indptr, indices, data, rows, row0, (int)k, p,
                                                                     alpha, G, Dx);
    else
        recsys_gram_dx_kernel<T, 4><<<(unsigned)b, 256, smem, st>>>(Dt, ldt, indptr, indices, data, rows, row0, (int)k, p,
                                                                     alpha, G, Dx);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int recsys_update_B(modl_ctx *ctx, T *B, int64_t ldb, const T *code, const int64_t *subset, const int64_t *col_ptr,
                           const int64_t *entry_row, const T *entry_val, int64_t *feature_n_iter, int64_t s, int64_t k,
                           double w, int64_t n_iter, cudaStream_t st)
{
    MODL_REQUIRE(ctx && B && code && subset && col_ptr && entry_row && entry_val && feature_n_iter && s >= 0 && k >= 1,
                 "recsys_update_B arguments");
    if (s == 0) return MODL_OK;
    dim3 grid((unsigned)ceil_div(s, 128), (unsigned)ceil_div(k, RS_KC));
    // w * n_iter_ is formed once, in double, as the reference does before dividing by the counts [ref: :183]
    recsys_update_B_kernel<T><<<grid, 128, 0, st>>>(B, ldb, code, subset, col_ptr, entry_row, entry_val, feature_n_iter, s,
                                                    (int)k, w * (double)n_iter);
    MODL_LAUNCH_CHECK(ctx);
    recsys_count_kernel<<<grid_for(ctx, ceil_div(s, 256), 8), 256, 0, st>>>(feature_n_iter, subset, col_ptr, s);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int recsys_update_C(modl_ctx *ctx, T *C, const T *code, const int64_t *rows, int64_t b, int64_t k, double w,
                           cudaStream_t st)
{
    MODL_REQUIRE(ctx && C && code && b >= 1 && k >= 1, "recsys_update_C arguments");
    const T *cb = code;
    if (rows) {
        T *tmp = nullptr;
        MODL_TRY(ws<T>(ctx, WS_CODE_BATCH, (size_t)(b * k), &tmp));
        gather_rows_kernel<T><<<grid_for(ctx, b, 16), 128, 0, st>>>(code, k, rows, (int)b, (int)k, tmp, k);
        MODL_LAUNCH_CHECK(ctx);
        cb = tmp;
    }
    // C_ = (1 - w) C_ + (w / b) code^T code   [ref: recsys.py:156-157]
    return gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, k, b, (T)(w / (double)b), cb, k, cb, k, (T)(1.0 - w), C, k, st);
}

template <typename T>
static int recsys_sync_transposed(modl_ctx *ctx, const T *D, int64_t ldd, T *Dt, int64_t ldt, const int64_t *subset,
                                  int64_t s, int64_t k, cudaStream_t st)
{
    MODL_REQUIRE(ctx && D && Dt && s >= 0 && k >= 1 && ldt >= k, "recsys_sync_transposed arguments");
    if (s == 0) return MODL_OK;
    dim3 grid((unsigned)ceil_div(s, 32), (unsigned)ceil_div(k, 32));
    recsys_transpose_kernel<T><<<grid, 256, 0, st>>>(D, ldd, Dt, ldt, subset, s, (int)k);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int recsys_predict(modl_ctx *ctx, const T *code, const T *Dt, int64_t ldt, const int64_t *indptr,
                          const int32_t *indices, int64_t n_rows, int64_t k, double *out, cudaStream_t st)
{
    MODL_REQUIRE(ctx && code && Dt && indptr && indices && out && n_rows >= 0 && k >= 1 && ldt >= k, "recsys_predict arguments");
    if (n_rows == 0) return MODL_OK;
    recsys_predict_kernel<T><<<grid_for(ctx, ceil_div(n_rows, 8), 8), 256, 0, st>>>(code, Dt, ldt, indptr, indices, n_rows,
                                                                                    (int)k, out);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

}  // namespace modl

using namespace modl;

extern "C" {

#define MODL_DEFINE_RECSYS(SFX, T)                                                                                          \
    int modl_recsys_gram_dx_##SFX(modl_ctx *c, const T *Dt, int64_t ldt, const int64_t *indptr, const int32_t *indices,      \
                                  const T *data, const int64_t *rows, int64_t row0, int64_t b, int64_t k, int64_t p,        \
                                  double alpha, T *G, T *Dx, void *st)                                                      \
    { CtxGuard g_(c); return recsys_gram_dx<T>(c, Dt, ldt, indptr, indices, data, rows, row0, b, k, p, alpha, G, Dx, (cudaStream_t)st); }   \
    int modl_recsys_update_B_##SFX(modl_ctx *c, T *B, int64_t ldb, const T *code, const int64_t *subset,                    \
                                   const int64_t *col_ptr, const int64_t *entry_row, const T *entry_val,                    \
                                   int64_t *feature_n_iter, int64_t s, int64_t k, double w, int64_t n_iter, void *st)       \
    { CtxGuard g_(c); return recsys_update_B<T>(c, B, ldb, code, subset, col_ptr, entry_row, entry_val, feature_n_iter, s, k, w, n_iter,    \
                                (cudaStream_t)st); }                                                                        \
    int modl_recsys_update_C_##SFX(modl_ctx *c, T *C, const T *code, const int64_t *rows, int64_t b, int64_t k, double w,   \
                                   void *st)                                                                                \
    { CtxGuard g_(c); return recsys_update_C<T>(c, C, code, rows, b, k, w, (cudaStream_t)st); }                                             \
    int modl_recsys_sync_transposed_##SFX(modl_ctx *c, const T *D, int64_t ldd, T *Dt, int64_t ldt, const int64_t *subset,  \
                                          int64_t s, int64_t k, void *st)                                                   \
    { CtxGuard g_(c); return recsys_sync_transposed<T>(c, D, ldd, Dt, ldt, subset, s, k, (cudaStream_t)st); }                               \
    int modl_recsys_predict_##SFX(modl_ctx *c, const T *code, const T *Dt, int64_t ldt, const int64_t *indptr,              \
                                  const int32_t *indices, int64_t n_rows, int64_t k, double *out, void *st)                 \
    { CtxGuard g_(c); return recsys_predict<T>(c, code, Dt, ldt, indptr, indices, n_rows, k, out, (cudaStream_t)st); }

MODL_DEFINE_RECSYS(f32, float)
MODL_DEFINE_RECSYS(f64, double)

}  // extern "C"
