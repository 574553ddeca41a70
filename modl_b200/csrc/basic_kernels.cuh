// basic_kernels.cuh -- HBM-bound helper kernels of the step: column gathers/scatters of
// the feature subset, full-row squared norms, running averages, elastic-net row helpers.
#pragma once
#include "common.cuh"

namespace modl {

// ---------------------------------------------------------------------------------------
// dst[r, j] = src[r, subset[j]]  (j < s)  and, optionally, norm2[r] = |src[r, :p]|^2.
// One CTA per row: the row is streamed once, coalesced, for the norm; the gather then hits
// L1/L2.  [ref: dict_fact.py:589 components_[:, subset], :594 X[:, subset]; the squared row
// norm is what the CD stop test scales its tolerance with, dict_fact_fast.pyx:334-336]
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_cols_kernel(const T *__restrict__ src, int64_t ld, int rows, int p,
                   const int64_t *__restrict__ subset, int s, T *__restrict__ dst, int64_t ldd,
                   T *__restrict__ norm2)
{
    __shared__ T scratch[33];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const T *row = src + (int64_t)r * ld;
        if (norm2 != nullptr) {
            T acc = T(0);
            for (int j = threadIdx.x; j < p; j += blockDim.x) {
                const T v = row[j];
                acc = fma(v, v, acc);
            }
            acc = block_sum(acc, scratch);
            if (threadIdx.x == 0) norm2[r] = acc;
        }
        if (dst != nullptr) {
            T *out = dst + (int64_t)r * ldd;
            for (int j = threadIdx.x; j < s; j += blockDim.x) out[j] = row[subset[j]];
        }
    }
}

// dst[r, subset[j]] = src[r, j]   [ref: dict_fact.py:709 components_[:, subset] = panel]
template <typename T>
__global__ void __launch_bounds__(256)
scatter_cols_kernel(const T *__restrict__ src, int64_t lds, int rows,
                    const int64_t *__restrict__ subset, int s, T *__restrict__ dst, int64_t ldd)
{
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const T *in = src + (int64_t)r * lds;
        T *row = dst + (int64_t)r * ldd;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += gridDim.x * blockDim.x)
            row[subset[j]] = in[j];
    }
}

// out[r, j] = keep * B[r, subset[j]] + inc[r, j]: the B_[:, subset] panel as it will be once the
// all-reduced increments are folded in (same fma as xpby_kernel)   [ref: dict_fact.py:532, :559-566]
template <typename T>
__global__ void __launch_bounds__(256)
gather_axpby_kernel(const T *__restrict__ B, int64_t ldb, int rows, const int64_t *__restrict__ subset, int s, T keep,
                    const T *__restrict__ inc, int64_t ldi, T *__restrict__ out, int64_t ldo)
{
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const T *row = B + (int64_t)r * ldb;
        for (int j = threadIdx.x; j < s; j += blockDim.x) {
            const T x = inc[(int64_t)r * ldi + j];
            out[(int64_t)r * ldo + j] = (keep != T(0)) ? fma(keep, row[subset[j]], x) : x;
        }
    }
}

// The two folds the sharded step does after its all-reduce, in ONE launch (it sits on the critical path of every rank):
//   C_[r, :] = keep * C_[r, :] + incC[r, :]                      (same fma as xpby_kernel)
//   out[r, j] = keep * B[r, subset[j]] + incB[r, j]              (same fma as gather_axpby_kernel)
template <typename T>
__global__ void __launch_bounds__(256)
apply_sub_kernel(T *__restrict__ Cmat, const T *__restrict__ incC, int k, const T *__restrict__ B, int64_t ldb,
                 const int64_t *__restrict__ subset, int s, T keep, const T *__restrict__ incB, int64_t ldi,
                 T *__restrict__ out, int64_t ldo)
{
    for (int r = blockIdx.x; r < k; r += gridDim.x) {
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            const int64_t i = (int64_t)r * k + j;
            Cmat[i] = (keep != T(0)) ? fma(keep, Cmat[i], incC[i]) : incC[i];
        }
        const T *row = B + (int64_t)r * ldb;
        for (int j = threadIdx.x; j < s; j += blockDim.x) {
            const T x = incB[(int64_t)r * ldi + j];
            out[(int64_t)r * ldo + j] = (keep != T(0)) ? fma(keep, row[subset[j]], x) : x;
        }
    }
}

// dst[ii, :] = src[indices[ii], :]
template <typename T>
__global__ void gather_rows_kernel(const T *__restrict__ src, int64_t ld, const int64_t *__restrict__ indices,
                                   int rows, int cols, T *__restrict__ dst, int64_t ldd)
{
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int64_t sr = indices ? indices[r] : r;
        for (int j = threadIdx.x; j < cols; j += blockDim.x)
            dst[(int64_t)r * ldd + j] = src[sr * ld + j];
    }
}

template <typename T>
__global__ void fill_kernel(T *x, int64_t n, T v)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = v;
}

// y = y + a * x  (flat)   [ref: dict_fact.py:700 components_subset += w * step_size * gradient_subset]
template <typename T>
__global__ void axpy_kernel(T *y, const T *x, int64_t n, T a)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = fma(a, x[i], y[i]);
}

// y = b * y + x  (flat): folds the all-reduced statistics increments into C_ / B_
template <typename T>
__global__ void xpby_kernel(T *y, const T *x, int64_t n, T b)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (b != T(0)) ? fma(b, y[i], x[i]) : x[i];
}

// ---------------------------------------------------------------------------------------
// Running averages of the 'average' estimators.
//   G_average[row] = (1 - w) G_average[row] + w G        [ref: dict_fact_fast.pyx:217-228]
//   Dx_average[row] = (1 - w) Dx_average[row] + w Dx ; Dx = Dx_average[row]
//                                                          [ref: dict_fact.py:596-601]
// The reference evaluates (1 - w) in double and rounds the product to the array dtype.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void update_g_average_kernel(T *__restrict__ G_average, const T *__restrict__ G,
                                        const T *__restrict__ w_sample, const int64_t *__restrict__ indices,
                                        int b, int64_t kk)
{
    for (int ii = blockIdx.y; ii < b; ii += gridDim.y) {
        const T w = w_sample[ii];
        const double keep = 1.0 - (double)w;
        T *dst = G_average + (indices ? indices[ii] : (int64_t)ii) * kk;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < kk;
             e += (int64_t)gridDim.x * blockDim.x) {
            T g = (T)((double)dst[e] * keep);
            g = g + G[e] * w;          // two roundings, like the reference's `+=` of a product
            dst[e] = g;
        }
    }
}

template <typename T>
__global__ void update_dx_average_kernel(T *__restrict__ Dx_average, T *__restrict__ Dx,
                                         const T *__restrict__ w_sample, const int64_t *__restrict__ indices,
                                         int b, int k)
{
    for (int ii = blockIdx.x; ii < b; ii += gridDim.x) {
        const T w = w_sample[ii];
        const T keep = T(1) - w;      // numpy: 1 - w_sample[:, None] in the array dtype
        T *avg = Dx_average + (indices ? indices[ii] : (int64_t)ii) * k;
        T *dx = Dx + (int64_t)ii * k;
        for (int j = threadIdx.x; j < k; j += blockDim.x) {
            T a = avg[j] * keep;
            a = a + dx[j] * w;
            avg[j] = a;
            dx[j] = a;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Elastic-net helpers on rows (one CTA per row).
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T enet_term(T v, T l1_ratio)
{
    const T a = t_abs(v);
    return a * (l1_ratio + (T(1) - l1_ratio) * a);     // [ref: enet.pyx:146-147]
}

template <typename T>
__global__ void __launch_bounds__(256)
enet_norm_rows_kernel(const T *__restrict__ v, int rows, int n, int64_t ld, T l1_ratio,
                      T *__restrict__ out, T sign, T *__restrict__ accumulate_into)
{
    __shared__ T scratch[33];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        T acc = T(0);
        for (int j = threadIdx.x; j < n; j += blockDim.x) acc += enet_term(v[(int64_t)r * ld + j], l1_ratio);
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) {
            if (out) out[r] = acc;
            if (accumulate_into) accumulate_into[r] += sign * acc;
        }
    }
}

// Threshold of the elastic-net-ball projection of one vector held in memory `v` (n values,
// read through `load(j)`), computed cooperatively by the whole CTA.
// The reference finds the active set (rho, s) with a pivot partition (enet.pyx:87-110) and
// then solves a quadratic for the threshold l (:113-119).  The active set is unique, so we
// reach the same (rho, s) by the monotone fixed point
//        A_0 = all,   l_t = root(A_t),   A_{t+1} = { |v_j| > l_t }
// which only ever shrinks towards the true active set (each l_t is a lower bound of the true
// threshold because inactive entries contribute non-positive terms), and stops when the set
// is unchanged.  Sums are taken in double.  Returns l (>= 0) to every thread; `inside` is set
// when the vector already lies in the ball (projection = identity).
template <typename T, typename Load>
__device__ T enet_threshold_block(Load load, int n, T radius_over_l1, T gamma, bool *inside, double *dscratch)
{
    // pass 0: norm = sum |v| (1 + gamma/2 |v|), also gives the stats of A_0
    double s1 = 0, s2 = 0, cnt = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double a = fabs((double)load(j));
        s1 += a; s2 += a * a; cnt += 1;
    }
    s1 = block_sum(s1, dscratch);
    s2 = block_sum(s2, dscratch);
    cnt = block_sum(cnt, dscratch);
    const double g = (double)gamma, R = (double)radius_over_l1;
    const double norm = s1 + 0.5 * g * s2;
    if ((T)norm <= radius_over_l1) { *inside = true; return T(0); }
    *inside = false;
    double l = 0;
    for (int it = 0; it < 64; ++it) {
        // root of (R g^2 + g rho/2) l^2 + (2 R g + rho) l + (R - s) = 0, s = s1 + g/2 s2
        const double sa = s1 + 0.5 * g * s2;
        double lnew;
        if (g != 0) {
            const double qa = g * g * R + 0.5 * g * cnt;
            const double qd = 2 * R * g + cnt;
            const double qc = R - sa;
            lnew = (-qd + sqrt(qd * qd - 4 * qa * qc)) / (2 * qa);
        } else {
            lnew = (sa - R) / cnt;
        }
        double n1 = 0, n2 = 0, nc = 0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double a = fabs((double)load(j));
            if (a > lnew) { n1 += a; n2 += a * a; nc += 1; }
        }
        n1 = block_sum(n1, dscratch);
        n2 = block_sum(n2, dscratch);
        nc = block_sum(nc, dscratch);
        l = lnew;
        if (nc == cnt || nc == 0) break;   // active set unchanged (or degenerate): l is final
        s1 = n1; s2 = n2; cnt = nc;
    }
    return (T)l;
}

// Three block-wide sums at once, each reduced by exactly the tree of block_sum() (same shuffles, same order of
// the per-warp partials), so the results are bit-identical to three block_sum calls at a third of the barriers.
// `scratch` holds >= 3 * nw + 3 doubles (nw = warps per CTA, <= 12).
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    __syncthreads();   // protect scratch from a previous use
    if (lane == 0) { scratch[wid] = a; scratch[nw + wid] = b; scratch[2 * nw + wid] = c; }
    __syncthreads();
    if (wid == 0) {
        double ta = lane < nw ? scratch[lane] : 0.0, tb = lane < nw ? scratch[nw + lane] : 0.0,
               tc = lane < nw ? scratch[2 * nw + lane] : 0.0;
        ta = warp_sum(ta); tb = warp_sum(tb); tc = warp_sum(tc);
        if (lane == 0) { scratch[3 * nw] = ta; scratch[3 * nw + 1] = tb; scratch[3 * nw + 2] = tc; }
    }
    __syncthreads();
    a = scratch[3 * nw]; b = scratch[3 * nw + 1]; c = scratch[3 * nw + 2];
}

// enet_threshold_block with the vector held in REGISTERS: every thread loads its n / blockDim values once (NV
// per thread at most; requires n <= NV * blockDim.x and blockDim.x <= 384) and the fixed-point passes run on
// registers with one fused reduction each -- no memory traffic and three barriers per pass instead of nine.
// Same per-thread summation order and the same reduction trees as enet_threshold_block: bit-identical result.
template <typename T, int NV, typename Load>
__device__ T enet_threshold_block_regs(Load load, int n, T radius_over_l1, T gamma, bool *inside, double *dscratch)
{
    T av[NV];                                        // |v_j| of my entries; -1 marks "past the end"
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int j = threadIdx.x + i * blockDim.x;
        av[i] = j < n ? t_abs(load(j)) : T(-1);
    }
    double s1 = 0, s2 = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (av[i] >= T(0)) {
            const double a = (double)av[i];
            s1 += a; s2 += a * a; cnt += 1;
        }
    }
    block_sum3(s1, s2, cnt, dscratch);
    const double g = (double)gamma, R = (double)radius_over_l1;
    const double norm = s1 + 0.5 * g * s2;
    if ((T)norm <= radius_over_l1) { *inside = true; return T(0); }
    *inside = false;
    double l = 0;
    for (int it = 0; it < 64; ++it) {
        const double sa = s1 + 0.5 * g * s2;
        double lnew;
        if (g != 0) {
            const double qa = g * g * R + 0.5 * g * cnt;
            const double qd = 2 * R * g + cnt;
            const double qc = R - sa;
            lnew = (-qd + sqrt(qd * qd - 4 * qa * qc)) / (2 * qa);
        } else {
            lnew = (sa - R) / cnt;
        }
        double n1 = 0, n2 = 0, nc = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double a = (double)av[i];
            if (av[i] >= T(0) && a > lnew) { n1 += a; n2 += a * a; nc += 1; }
        }
        block_sum3(n1, n2, nc, dscratch);
        l = lnew;
        if (nc == cnt || nc == 0) break;
        s1 = n1; s2 = n2; cnt = nc;
    }
    return (T)l;
}

template <typename T>
__device__ __forceinline__ T enet_shrink(T v, T l, T gamma)
{
    // sign(v) * max(|v| - l, 0) / (1 + l gamma)        [ref: enet.pyx:120-121]
    const T a = t_abs(v) - l;
    const T m = a > T(0) ? a : T(0);
    return (v >= T(0) ? m : -m) / (T(1) + l * gamma);
}

// out[r] = projection of v[r] on the elastic-net ball of radius radius[r]  [ref: enet.pyx:38-122]
template <typename T>
__global__ void __launch_bounds__(256)
enet_projection_rows_kernel(const T *__restrict__ v, T *__restrict__ out, int rows, int n, int64_t ld,
                            const T *__restrict__ radius, T l1_ratio)
{
    __shared__ T scratch[33];
    __shared__ double dscratch[33];
    __shared__ bool inside;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const T *vr = v + (int64_t)r * ld;
        T *outr = out + (int64_t)r * ld;
        const T rad = radius[r];
        if (rad == T(0)) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) outr[j] = T(0);
            continue;
        }
        if (l1_ratio == T(0)) {
            T acc = T(0);
            for (int j = threadIdx.x; j < n; j += blockDim.x) acc = fma(vr[j], vr[j], acc);
            acc = block_sum(acc, scratch);
            const T nrm = (acc <= rad) ? T(1) : t_sqrt(acc / rad);
            for (int j = threadIdx.x; j < n; j += blockDim.x) outr[j] = vr[j] / nrm;
        } else {
            const T gamma = T(2) / l1_ratio - T(2);
            const T R = rad / l1_ratio;
            bool ins_local;
            const T l = enet_threshold_block<T>([&](int j) { return vr[j]; }, n, R, gamma, &ins_local, dscratch);
            if (threadIdx.x == 0) inside = ins_local;
            __syncthreads();
            const bool ins = inside;
            for (int j = threadIdx.x; j < n; j += blockDim.x)
                outr[j] = ins ? vr[j] : enet_shrink(vr[j], l, gamma);
            __syncthreads();
        }
    }
}

// X[r] *= S with S chosen so that enet_norm(X[r]) == radius   [ref: enet.pyx:150-168]
template <typename T>
__global__ void __launch_bounds__(256)
enet_scale_rows_kernel(T *__restrict__ X, int rows, int n, int64_t ld, T l1_ratio, T radius)
{
    __shared__ T scratch[33];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        T *xr = X + (int64_t)r * ld;
        T l1 = T(0), l2 = T(0);
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            l1 += t_abs(xr[j]);
            l2 = fma(xr[j], xr[j], l2);
        }
        l1 = block_sum(l1, scratch);
        l2 = block_sum(l2, scratch);
        l1 = l1 * l1_ratio;
        l2 = (T)((double)l2 * (1.0 - (double)l1_ratio));
        T S = T(0);
        if (l2 != T(0))
            S = (T)((-(double)l1 + sqrt((double)(l1 * l1) + (4.0 * (double)radius) * (double)l2)) / (2.0 * (double)l2));
        else if (l1 != T(0))
            S = radius / l1;
        for (int j = threadIdx.x; j < n; j += blockDim.x) xr[j] = xr[j] * S;
        __syncthreads();
    }
}

}  // namespace modl
