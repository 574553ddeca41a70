// cd_kernels.cuh -- batched elastic-net regression by cyclic coordinate descent over a Gram
// matrix: one warp per sample.
//
// Replaces enet_coordinate_descent_gram [ref: modl/decomposition/dict_fact_fast.pyx:270-427]
// and the per-sample loops that call it (:198-214 shared Gram, :95-112 per-sample Gram).
//
// Mapping to the hardware
//   * A sample's state lives in the registers of ONE warp: lane l holds coordinates
//     {32 J + l}, J = 0..TILES-1, of w (code), q (= Dx row) and H (= Q w).  k <= 32*TILES.
//   * Coordinates are visited in the reference's cyclic order (:354-355).  The owner lane
//     computes the soft-thresholded update; two warp shuffles broadcast (w_old, w_new) and all
//     lanes apply the rank-1 corrections  H -= w_old Q[c,:],  H += w_new Q[c,:]  as two
//     separate FMAs, like the reference's two axpy calls (:359-363, :374-377).
//   * Shared-Gram variant: the k x k Gram is kept in shared memory as its lower triangle of
//     32 x 32 tiles (k = 256 -> 36 tiles = 144 KB, fits where the full 256 KB matrix does
//     not).  A tile stores element (a, c) at a*32 + ((a + c) & 31): rows AND
//     columns of a tile are bank-conflict free, so row c of the symmetric matrix is read as
//     tile rows left of the diagonal and tile columns below it.
//   * Per-sample-Gram / large-k variant: rows stream from global memory (coalesced 128 B per
//     tile column), strided per sample.
//   * The stop test mirrors the reference arithmetic [ref: :388-427]: sums in the working
//     precision, the duality-gap combination in double exactly where the Cython float
//     specialisation promotes (R_norm2, const, A_norm2).
#pragma once
#include "common.cuh"

namespace modl {

constexpr int CD_TILE = 32;
constexpr int CD_TILE_ELEMS = CD_TILE * CD_TILE;

__host__ __device__ inline int cd_tri(int I) { return I * (I + 1) / 2; }
__host__ __device__ inline size_t cd_packed_elems(int tiles) { return (size_t)cd_tri(tiles) * CD_TILE_ELEMS; }

// Fill the packed lower-triangular tile image of G in shared memory (whole CTA).
template <typename T>
__device__ void cd_load_packed_gram(T *sG, const T *__restrict__ G, int k, int tiles)
{
    const int total = cd_tri(tiles) * CD_TILE_ELEMS;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int t = e / CD_TILE_ELEMS, within = e % CD_TILE_ELEMS;
        int I = 0;
        while (cd_tri(I + 1) <= t) ++I;
        const int J = t - cd_tri(I);
        const int a = within / CD_TILE, c = within % CD_TILE;
        const int gi = I * CD_TILE + a, gj = J * CD_TILE + c;
        T v = T(0);
        if (gi < k && gj < k) v = G[(int64_t)gi * k + gj];
        sG[t * CD_TILE_ELEMS + a * CD_TILE + ((a + c) & 31)] = v;
    }
}

// Same image written to GLOBAL memory once per Gram matrix (grid: one CTA per lower tile), so
// that every CTA of the solver can pull it into shared memory with one TMA bulk copy per tile
// instead of 36 K scalar loads each.
template <typename T>
__global__ void __launch_bounds__(256)
cd_pack_gram_kernel(const T *__restrict__ G, int k, int tiles, T *__restrict__ packed)
{
    const int t = blockIdx.x;
    if (t == cd_tri(tiles)) {
        // one extra CTA: H for the all-ones warm start of a first visit [ref: dict_fact.py:470, code_ = ones], i.e. the
        // rows of G added in coordinate order -- bit for bit what the solver's  H += w_c G[c,:]  loop produces for w = 1
        // (fma(1, g, h) rounds once, like the add), so that warps whose sample still carries the warm start skip k
        // dependent row loads
        T *h_ones = packed + (int64_t)cd_tri(tiles) * CD_TILE_ELEMS;
        for (int j = threadIdx.x; j < tiles * CD_TILE; j += blockDim.x) {
            T acc = T(0);
            if (j < k) {
                int c = 0;
                for (; c + 32 <= k; c += 32) {          // 32 loads in flight, added in coordinate order
                    T v[32];
#pragma unroll
                    for (int u = 0; u < 32; ++u) v[u] = G[(int64_t)(c + u) * k + j];
#pragma unroll
                    for (int u = 0; u < 32; ++u) acc = acc + v[u];
                }
                for (; c < k; ++c) acc = acc + G[(int64_t)c * k + j];
            }
            h_ones[j] = acc;
        }
        return;
    }
    int I = 0;
    while (cd_tri(I + 1) <= t) ++I;
    const int J = t - cd_tri(I);
    for (int within = threadIdx.x; within < CD_TILE_ELEMS; within += blockDim.x) {
        const int a = within / CD_TILE, c = within % CD_TILE;
        const int gi = I * CD_TILE + a, gj = J * CD_TILE + c;
        T v = T(0);
        if (gi < k && gj < k) v = G[(int64_t)gi * k + gj];
        packed[(int64_t)t * CD_TILE_ELEMS + a * CD_TILE + ((a + c) & 31)] = v;
    }
}

// TMA (cp.async.bulk) copy of the packed image global -> shared, completion on an mbarrier.
__device__ __forceinline__ void cd_bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned chunk,
                                             unsigned long long *mbar)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
        for (unsigned off = 0; off < bytes; off += chunk) {
            const unsigned n = (bytes - off) < chunk ? (bytes - off) : chunk;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst + off), "l"((const char *)gmem_src + off), "r"(n), "r"(bar) : "memory");
        }
    }
    // everybody waits for phase 0 of the barrier
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar) : "memory");
    }
}

// explicit shared-window loads (32-bit address computed once per warp; the per-tile offsets fold
// into the instruction immediates)
__device__ __forceinline__ float lds_real(unsigned addr, float) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_real(unsigned addr, double) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Row `c = 32*Jc + l` of the symmetric matrix from the packed image; lane gets columns 32*JJ+lane.
// sbase: shared address of the image; swz = (l + lane) & 31.
template <typename T, int TILES>
__device__ __forceinline__ void cd_row_packed(unsigned sbase, int Jc, int l, int lane, T (&r)[TILES])
{
    const unsigned swz = (unsigned)((l + lane) & 31);
    const unsigned rowoff = sbase + (unsigned)sizeof(T) * ((unsigned)l * CD_TILE + swz);      // tile row l
    const unsigned coloff = sbase + (unsigned)sizeof(T) * ((unsigned)lane * CD_TILE + swz);   // tile column l
#pragma unroll
    for (int JJ = 0; JJ < TILES; ++JJ) {
        if (JJ <= Jc) r[JJ] = lds_real(rowoff + (unsigned)sizeof(T) * (unsigned)((cd_tri(Jc) + JJ) * CD_TILE_ELEMS), T(0));
        else          r[JJ] = lds_real(coloff + (unsigned)sizeof(T) * (unsigned)((cd_tri(JJ) + Jc) * CD_TILE_ELEMS), T(0));
    }
}

template <typename T, int TILES>
__device__ __forceinline__ void cd_row_global(const T *__restrict__ Gs, int k, int c, int lane, T (&r)[TILES])
{
    const T *row = Gs + (int64_t)c * k;
#pragma unroll
    for (int JJ = 0; JJ < TILES; ++JJ) {
        const int col = JJ * CD_TILE + lane;
        r[JJ] = col < k ? row[col] : T(0);
    }
}

// h += c * r over the register tiles.  float: packed FFMA2 (fma.rn.f32x2, new on sm_100) halves
// the instruction count of the two rank-1 updates that dominate a coordinate step.
template <int TILES>
__device__ __forceinline__ void cd_axpy_tiles(float (&h)[TILES], const float (&r)[TILES], float c)
{
    const float2 cc = make_float2(c, c);
#pragma unroll
    for (int m = 0; m + 1 < TILES; m += 2) {
        const float2 hv = __ffma2_rn(cc, make_float2(r[m], r[m + 1]), make_float2(h[m], h[m + 1]));
        h[m] = hv.x;
        h[m + 1] = hv.y;
    }
    if (TILES & 1) h[TILES - 1] = fmaf(c, r[TILES - 1], h[TILES - 1]);
}
template <int TILES>
__device__ __forceinline__ void cd_axpy_tiles(double (&h)[TILES], const double (&r)[TILES], double c)
{
#pragma unroll
    for (int m = 0; m < TILES; ++m) h[m] = fma(c, r[m], h[m]);
}

// One coordinate-descent solve for the sample owned by this warp.
// POS (non-negative codes) is a template parameter: the sign tests sit on the dependent chain of every coordinate step
template <typename T, int TILES, bool PACKED, bool POS>
__device__ __forceinline__ int cd_solve_warp(unsigned sbase, const T *__restrict__ Gs, int k, int lane,
                                             T (&w)[TILES], const T (&q)[TILES], T ynorm2, T alpha,
                                             T beta, T tol, int max_iter, const T *__restrict__ h_ones = nullptr)
{
    constexpr bool positive = POS;
    T h[TILES], r[TILES], inv[TILES], diag[TILES];
    unsigned dead = 0;      // bit J: my coordinate 32 J + lane has a zero diagonal (or is padding)
    // 1 / (Q[c,c] + beta) for my own coordinates, once per sample.  The per-step division of
    // the reference (:372-373) becomes multiply + one Newton correction (correctly rounded in
    // all but pathological cases), which keeps the IEEE-division slow path -- triggered by the
    // many zero numerators of a sparse code -- off the dependent chain.
#pragma unroll
    for (int J = 0; J < TILES; ++J) {
        const int c = J * CD_TILE + lane;
        T dg;
        if (PACKED) dg = lds_real(sbase + (unsigned)sizeof(T) * (unsigned)((cd_tri(J) + J) * CD_TILE_ELEMS + lane * CD_TILE + ((2 * lane) & 31)), T(0));
        else        dg = c < k ? Gs[(int64_t)c * k + c] : T(0);
        if (dg == T(0) || c >= k) dead |= (1u << J);
        diag[J] = dg;
        inv[J] = T(1) / (dg + beta);
        h[J] = T(0);
    }
    // ---- H = Q w accumulated column by column (Q symmetric)  [ref: :340-347] ----
    bool ones = h_ones != nullptr;
#pragma unroll
    for (int J = 0; J < TILES; ++J) ones = ones && (w[J] == ((J * CD_TILE + lane) < k ? T(1) : T(0)));
    ones = __all_sync(kFullMask, ones);
    if (ones) {     // the warm start of a first visit: H precomputed once per Gram matrix (cd_pack_gram_kernel), same bits
#pragma unroll
        for (int J = 0; J < TILES; ++J) h[J] = h_ones[J * CD_TILE + lane];
    }
#pragma unroll
    for (int J = 0; J < TILES; ++J) {
        const int lmax = ones ? 0 : min(CD_TILE, k - J * CD_TILE);
        for (int l = 0; l < lmax; ++l) {
            const T wc = __shfl_sync(kFullMask, w[J], l);
            if (wc != T(0)) {
                if (PACKED) cd_row_packed<T, TILES>(sbase, J, l, lane, r);
                else        cd_row_global<T, TILES>(Gs, k, J * CD_TILE + l, lane, r);
                cd_axpy_tiles<TILES>(h, r, wc);
            }
        }
    }

    const T d_w_tol = tol;
    const T tolv = tol * ynorm2;                         // [ref: :336]
    int sweeps = 0;
    for (int it = 0; it < max_iter; ++it) {
        sweeps = it + 1;
        T w0[TILES];
#pragma unroll
        for (int J = 0; J < TILES; ++J) w0[J] = w[J];
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const int lmax = min(CD_TILE, k - J * CD_TILE);
            const bool my_live = !((dead >> J) & 1u) && lane < lmax;   // Q[c,c] == 0: coordinate skipped [ref: :357-358]
            // Active-set walk over the 32 coordinates of tile J.  A coordinate whose weight is zero
            // and whose soft-threshold candidate is zero is an exact no-op of the reference's step
            // (neither axpy runs, :359 and :374 test w[ii] != 0), and H is untouched by no-ops, so
            // every lane can test its own coordinate against the CURRENT H in parallel; the warp
            // then jumps to the first coordinate that does something, executes it exactly like the
            // sequential loop would, and re-tests the lanes after it.  Sparse codes (a few percent
            // of non-zeros) therefore cost a few dependent steps per tile instead of 32.
            unsigned todo = kFullMask;
            while (true) {
                const T tmp0 = q[J] - h[J];
                const T mag0 = t_abs(tmp0) - alpha;
                const bool moves = mag0 > T(0) && !(positive && tmp0 < T(0));
                const unsigned act = __ballot_sync(kFullMask, my_live && (w[J] != T(0) || moves)) & todo;
                // The new weight of MY coordinate as if it were the next one visited -- every lane, before the ballot
                // resolves and before any row is loaded: on the visited lane l the reference's first update
                // H -= w_old Q[c,:] leaves h[J] = fma(-w[J], Q[c,c], h[J]) (row element r[J] on lane l IS the diagonal, which
                // the lane keeps in a register), and everything the candidate needs is the lane's own state.  Same
                // operations on the same operands as computing it after the row arrives (bit-identical), but the
                // soft threshold and the Newton division now overlap the ballot and the shared-memory row load instead of
                // following them on the dependent chain.
                const T hl = fma(-w[J], diag[J], h[J]);
                const T tmp = q[J] - hl;
                const T mag = t_abs(tmp) - alpha;
                T ms = mag > T(0) ? mag : T(0);
                ms = (tmp < T(0)) ? (positive ? T(0) : -ms) : ms;
                const T den = diag[J] + beta;
                T cand = ms * inv[J];
                cand = fma(fma(-cand, den, ms), inv[J], cand);         // Newton step: ms / den
                if (act == 0u) break;
                const int l = __ffs(act) - 1;
                todo = (l == 31) ? 0u : (kFullMask << (l + 1));
                if (PACKED) cd_row_packed<T, TILES>(sbase, J, l, lane, r);
                else        cd_row_global<T, TILES>(Gs, k, J * CD_TILE + l, lane, r);
                const T w_old = __shfl_sync(kFullMask, w[J], l);
                const T w_new = __shfl_sync(kFullMask, cand, l);
                // H -= w_old Q[c,:] -- a zero coefficient leaves H bit-unchanged; then H += w_new Q[c,:]  [ref: :359-376]
                cd_axpy_tiles<TILES>(h, r, -w_old);
                w[J] = (lane == l) ? w_new : w[J];
                cd_axpy_tiles<TILES>(h, r, w_new);
            }
        }
        // max |w| and max |delta w| of this sweep over the visited coordinates [ref: :379-386]
        T w_max = T(0), d_w_max = T(0);
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const T aw = ((dead >> J) & 1u) ? T(0) : t_abs(w[J]);
            const T dw = t_abs(w[J] - w0[J]);
            w_max = aw > w_max ? aw : w_max;
            d_w_max = dw > d_w_max ? dw : d_w_max;
        }
        w_max = warp_max(w_max);
        d_w_max = warp_max(d_w_max);
        if (w_max == T(0) || d_w_max / w_max < d_w_tol || it == max_iter - 1) {
            // ---- duality gap  [ref: :388-427] ----
            T qw = T(0), wh = T(0), w2 = T(0), l1 = T(0);
            T dual = positive ? -INFINITY : T(0);
#pragma unroll
            for (int J = 0; J < TILES; ++J) {
                const bool real = (J * CD_TILE + lane) < k;
                qw = fma(w[J], q[J], qw);
                wh = fma(w[J], h[J], wh);
                w2 = fma(w[J], w[J], w2);
                l1 += t_abs(w[J]);
                const T xta = q[J] - h[J] - beta * w[J];
                if (real) {
                    const T cmp = positive ? xta : t_abs(xta);
                    if (cmp > dual) dual = cmp;
                }
            }
            qw = warp_sum(qw); wh = warp_sum(wh); w2 = warp_sum(w2); l1 = warp_sum(l1);
            dual = warp_max(dual);
            const double R_norm2 = (double)(ynorm2 + wh) - 2.0 * (double)qw;
            double cst;
            T gap;
            if (dual > alpha) {
                cst = (double)alpha / (double)dual;
                const double A_norm2 = R_norm2 * (cst * cst);
                gap = (T)(0.5 * (R_norm2 + A_norm2));
            } else {
                cst = 1.0;
                gap = (T)R_norm2;
            }
            gap = (T)((double)gap + ((((double)(alpha * l1) - cst * (double)ynorm2) + cst * (double)qw)
                                     + ((0.5 * (double)beta) * (1.0 + cst * cst)) * (double)w2));
            if (gap < tolv) break;
        }
    }
    return sweeps;
}

// grid: any; block: 32 * warps.  Dynamic smem (PACKED): cd_packed_elems(TILES) * sizeof(T).
// g_stride: elements between consecutive per-sample Gram matrices (0 = one shared Gram).
template <typename T, int TILES, bool PACKED>
__global__ void cd_regression_kernel(const T *__restrict__ G, int64_t g_stride, const T *__restrict__ Dx,
                                     const T *__restrict__ xnorm2, T *__restrict__ code,
                                     const int64_t *__restrict__ indices, T *__restrict__ code_batch,
                                     int b, int k, T alpha, T beta, T tol, int max_iter, int positive,
                                     int32_t *__restrict__ sweeps_out)
{
    extern __shared__ __align__(128) unsigned char cd_smem_raw[];
    T *sG = reinterpret_cast<T *>(cd_smem_raw);
    if (PACKED) {
        // G points at the tile-packed image (cd_pack_gram_kernel); one TMA bulk copy per tile
        __shared__ __align__(8) unsigned long long mbar;
        cd_bulk_load(sG, G, (unsigned)(cd_packed_elems(TILES) * sizeof(T)), (unsigned)(CD_TILE_ELEMS * sizeof(T)), &mbar);
    }
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    for (int ii = blockIdx.x * warps + (threadIdx.x >> 5); ii < b; ii += gridDim.x * warps) {
        const int64_t row = indices ? indices[ii] : (int64_t)ii;
        T *wrow = code + row * k;
        const T *qrow = Dx + (int64_t)ii * k;
        T w[TILES], q[TILES];
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const int c = J * CD_TILE + lane;
            w[J] = c < k ? wrow[c] : T(0);
            q[J] = c < k ? qrow[c] : T(0);
        }
        const T *Gs = G + (int64_t)ii * g_stride;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sG);
        const T *h_ones = PACKED ? G + cd_packed_elems(TILES) : (const T *)nullptr;
        const int sw = positive ? cd_solve_warp<T, TILES, PACKED, true>(sbase, Gs, k, lane, w, q, xnorm2[ii], alpha, beta, tol,
                                                                        max_iter, h_ones)
                                : cd_solve_warp<T, TILES, PACKED, false>(sbase, Gs, k, lane, w, q, xnorm2[ii], alpha, beta, tol,
                                                                         max_iter, h_ones);
#pragma unroll
        for (int J = 0; J < TILES; ++J) {
            const int c = J * CD_TILE + lane;
            if (c < k) {
                wrow[c] = w[J];
                if (code_batch) code_batch[(int64_t)ii * k + c] = w[J];
            }
        }
        if (sweeps_out && lane == 0) sweeps_out[ii] = sw;
    }
}

}  // namespace modl
