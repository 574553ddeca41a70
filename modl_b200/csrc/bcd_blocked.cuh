// bcd_blocked.cuh -- dictionary update with the projection scalars solved BLOCK-WISE in coefficient space, for the
// L2 ball without positivity (comp_l1_ratio == 0, comp_pos == False: the benchmark configuration) and panels that fit
// the shared memory of one thread-block cluster.  Same mathematics as the per-atom kernels (bcd_kernels.cuh,
// bcd_pilot.cuh) [ref: modl/decomposition/dict_fact.py:675-694, modl/utils/math/enet.pyx:62-70]; what changes is WHERE
// the k-step dependent chain runs.
//
// The per-atom kernels exchange one scalar over the cluster per atom (|v_t|^2, the squared norm of the candidate row):
// k dependent cluster round trips, ~1 400 cycles each.  But for the L2 ball the chain only couples the atoms through
// SCALARS.  Inside a block of M = 16 consecutive atoms (a_1 .. a_M in update order), with every earlier block applied,
//
//     g_t = d_t + (B_sub[a_t] - C[a_t,:] . D_sub) / C[a_t,a_t]      the candidate if no atom of the block had moved,
//     v_t = g_t - sum_{j<t} L_tj delta_j,    L_tj = C[a_t,a_j] / C[a_t,a_t],
//     n_t = alpha_t v_t,  alpha_t = min(1, sqrt(radius_t / |v_t|^2)),   delta_t = n_t - d_t,
//
// so every v_t, n_t, delta_t lives in the span of the 2M = 32 vectors {g_1..g_M, d_1..d_M}, and |v_t|^2 = w_t^T G w_t
// with G their 32 x 32 Gram matrix and w_t the coefficient vector of v_t.  Per block the cluster therefore does
//
//   S0  apply the previous block (n_t = sum_r coef[t][r] basis_r on the CTA's columns), repair the look-ahead product
//       with the previous block's deltas, form g_t / d_t for this block                       (all warps, local)
//   S1  partial Gram of the 32 basis vectors over the CTA's columns (4x4 register tiles)      (9 warps, local)
//   S2  ONE all-to-all of the 576 partial Gram entries (cp.async.bulk shared -> peer shared, mbarrier complete_tx)
//   S3  fixed-order sum of the 16 partials -> G, bit-identical in every CTA
//   S4  warp 0: the 16-step scalar recurrence in registers (lane r = basis vector r; w_t, G w_t by the same
//       recurrence, one warp reduction per atom) -> coef;  warps 1-11 meanwhile: look-ahead product
//       C[block b+1, :] . D_sub for the next block and its operand loads
//
// i.e. ONE cluster exchange per 16 atoms, and a dependent chain of ~16 x (one warp reduction + one rsqrt) per block
// that touches registers only.  Every CTA solves the same 32-dimensional problem redundantly from bit-identical inputs
// with the same instruction sequence, so all CTAs apply the same scalars (no broadcast, no atomics: deterministic).
// Arithmetic differs from the per-atom kernels in the way |v_t|^2 is assembled (quadratic form instead of a direct
// sum of squares; relative difference ~1e-7 in float) and in the new row being a linear combination evaluated once.
#pragma once
#include "bcd_pilot.cuh"

namespace modl {

constexpr int BB_M = 16;                 // atoms per block
constexpr int BB_NB = 2 * BB_M;          // basis vectors per block (= warp size)
constexpr int BB_THREADS = 384;
constexpr int BB_TILES = 36;             // 4x4 tiles of the upper triangle of the 32x32 Gram (8x8 tile grid)
constexpr int BB_GRAM = BB_TILES * 16;   // values each CTA contributes per block
constexpr int BB_LA = 8;                 // atoms per pass of the look-ahead product
constexpr int BB_MLD = BB_NB + 1;        // row pitch of the Gram matrix in shared memory
enum { BB_BAR_WORKERS = 1 };

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
// acc.xy += a.xy * b.xy  (float: one packed FFMA2 -- the fma pipe issues an FFMA2 at the rate of an FFMA)
__device__ __forceinline__ void pair_mul_fma(const Pair<float> &a, const Pair<float> &b, Pair<float> &acc)
{
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(acc.x, acc.y));
    acc.x = r.x; acc.y = r.y;
}
__device__ __forceinline__ void pair_mul_fma(const Pair<double> &a, const Pair<double> &b, Pair<double> &acc)
{
    acc.x = fma(a.x, b.x, acc.x); acc.y = fma(a.y, b.y, acc.y);
}

// shared-memory footprint in bytes -- must match the carve-up in the kernel
template <typename T>
__host__ __device__ inline size_t bcd_blocked_smem_bytes(int64_t k, int64_t ncp)
{
    const int64_t kp = round_up(k, 32);
    const int64_t elems = k * ncp                 // Ds
                          + kp * BB_M             // Cblk
                          + 2 * BB_M * ncp        // Rraw, Brow
                          + BB_NB * ncp           // basis
                          + BB_M * ncp            // dlt
                          + BB_M * BB_NB          // coef
                          + 2 * BB_M * BB_M       // Lblk, Cx
                          + BB_NB * BB_MLD + 32   // Mfull (padded to a multiple of 4)
                          + 2 * kp                // cnorm, rad
                          + 2 * BB_M              // caa, rcaa
                          + 2 * BB_GRAM           // stage (two blocks)
                          + BCD_MAX_CLUSTER * BB_GRAM;   // recv (also: scratch of the look-ahead product)
    return (size_t)elems * sizeof(T) + (size_t)kp * sizeof(int) + 64;
}

// largest scratch the look-ahead product needs (elements), must fit the recv area
__host__ __device__ inline int64_t bcd_blocked_red_elems(int64_t ncp)
{
    const int64_t np = ncp / 2;
    const int64_t igw = (BB_THREADS - 32) / np > 0 ? (BB_THREADS - 32) / np : 1;
    return igw * BB_LA * ncp;
}

template <typename T>
__global__ void __launch_bounds__(BB_THREADS, 1)
bcd_blocked_kernel(BcdParams<T> P)
{
    extern __shared__ __align__(16) unsigned char bb_smem_raw[];
    const int k = P.k, s = P.s, lds = P.lds;
    const int nblk = gridDim.x, g = blockIdx.x;
    const int c0 = min(s, g * P.cols_per_cta);
    const int c1 = min(s, c0 + P.cols_per_cta);
    const int nc = c1 - c0;
    const int ncp = (int)round_up(P.cols_per_cta, 32);
    const int kp = (int)round_up(k, 32);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nbk = (k + BB_M - 1) / BB_M;
    const int NP = ncp >> 1;                                 // column pairs of the slice

    // ---- shared memory carve-up (see bcd_blocked_smem_bytes) ----
    T *Ds = reinterpret_cast<T *>(bb_smem_raw);              // [k][ncp]   the CTA's slice of D_sub
    T *Cblk = Ds + (size_t)k * ncp;                          // [M][kp]    C[a_j, :] of the block in flight / ahead
    T *Rraw = Cblk + kp * BB_M;                              // [M][ncp]   look-ahead product C[a_j,:] . D_sub (previous block not applied)
    T *Brow = Rraw + BB_M * ncp;                             // [M][ncp]   B_sub rows of the block
    T *basis = Brow + BB_M * ncp;                            // [2M][ncp]  g_1..g_M, d_1..d_M
    T *dlt = basis + BB_NB * ncp;                            // [M][ncp]   deltas of the block just applied
    T *coef = dlt + BB_M * ncp;                              // [M][2M]    n_t = sum_r coef[t][r] basis_r
    T *Lblk = coef + BB_M * BB_NB;                           // [M][M]     L_tj (j < t), zero elsewhere
    T *Cx = Lblk + BB_M * BB_M;                              // [M][M]     C[a_j, a_i(previous block)] at [i*M + j]
    T *Mfull = Cx + BB_M * BB_M;                             // [2M][2M+1] Gram matrix of the basis
    T *cnorm = Mfull + BB_NB * BB_MLD + 32;                  // [kp]       comp_norm_ on entry
    T *rad = cnorm + kp;                                     // [kp]       radius used for every atom
    T *caa = rad + kp;                                       // [M]        C[a_t, a_t]
    T *rcaa = caa + BB_M;                                    // [M]        1 / C[a_t, a_t]
    T *stage = rcaa + BB_M;                                  // [2][GRAM]  this CTA's partial Gram, by block parity
    T *recv = stage + 2 * BB_GRAM;                           // [16][GRAM] partial Grams of every CTA
    T *red = recv;                                           //            scratch of the look-ahead product (S4 only)
    int *ord_s = reinterpret_cast<int *>(recv + BCD_MAX_CLUSTER * BB_GRAM);   // [kp] update order
    __shared__ __align__(8) unsigned long long xbar;
    __shared__ unsigned char tile_i[BB_TILES], tile_j[BB_TILES];
    const unsigned xbar_addr = (unsigned)__cvta_generic_to_shared(&xbar);
    constexpr unsigned kGramBytes = BB_GRAM * sizeof(T);

    // debug stamps of CTA 0: thread 0 (solver warp) [b][8] phase starts, thread 32 (first product warp) look-ahead
    // start / end per block at 8k + 16 + 2b, kernel-level stamps at 8k .. 8k + 7
    long long *stamp = (P.timing && g == 0 && tid == 0) ? P.timing : nullptr;
    long long *wstamp = (P.timing && g == 0 && tid == 32) ? P.timing + (int64_t)8 * k + 16 : nullptr;
#define BB_STAMP(b_, slot_) do { if (stamp) stamp[(int64_t)(b_) * 8 + (slot_)] = clock64(); } while (0)
    if (stamp) stamp[(int64_t)8 * k] = clock64();

    constexpr int VE = 16 / (int)sizeof(T);                  // elements per 16-byte copy
    const bool c_vec = (k % VE == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    const bool p_vec = (lds % VE == 0) && (c0 % VE == 0) && ((reinterpret_cast<uintptr_t>(P.Dp) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(P.Bp) & 15) == 0);

    // operands of block b: C[a_j, :] and B_sub[a_j, my columns]; every copy asynchronous and in flight at once
    auto issue_loads = [&](int b) {
        const int mb = min(BB_M, k - b * BB_M);
        if (c_vec) {
            const int nv = k / VE;
            for (int e = tid; e < BB_M * nv; e += BB_THREADS) {
                const int j = e / nv, iv = (e % nv) * VE;
                if (j < mb) cp_async_16(Cblk + j * kp + iv, P.C + (int64_t)ord_s[b * BB_M + j] * k + iv);
            }
        } else {
            for (int e = tid; e < BB_M * k; e += BB_THREADS) {
                const int j = e / k, i = e % k;
                if (j < mb) cp_async_elem(Cblk + j * kp + i, P.C + (int64_t)ord_s[b * BB_M + j] * k + i);
            }
        }
        const int nvb = ncp / VE;
        for (int e = tid; e < BB_M * nvb; e += BB_THREADS) {
            const int j = e / nvb, cv = (e % nvb) * VE;
            T *dst = Brow + j * ncp + cv;
            if (j < mb && p_vec && cv + VE <= nc) {
                cp_async_16(dst, P.Bp + (int64_t)ord_s[b * BB_M + j] * lds + c0 + cv);
            } else {
#pragma unroll
                for (int u = 0; u < VE; ++u) {
                    if (j < mb && cv + u < nc) cp_async_elem(dst + u, P.Bp + (int64_t)ord_s[b * BB_M + j] * lds + c0 + cv + u);
                    else dst[u] = T(0);
                }
            }
        }
        cp_async_commit();
    };

    // ---- prologue: D slice (asynchronous copies), norms, order, tables ----
    for (int i = tid; i < kp; i += BB_THREADS) {
        cnorm[i] = i < k ? P.comp_norm[i] : T(0);
        ord_s[i] = i < k ? P.order[i] : 0;
    }
    {
        const T *Dg = P.Dp + c0;
        const int nv = ncp / VE;
        for (int e = tid; e < k * nv; e += BB_THREADS) {
            const int i = e / nv, cv = (e % nv) * VE;
            T *dst = Ds + i * ncp + cv;
            if (p_vec && cv + VE <= nc) {
                cp_async_16(dst, Dg + (int64_t)i * lds + cv);
            } else {
#pragma unroll
                for (int u = 0; u < VE; ++u) {
                    if (cv + u < nc) cp_async_elem(dst + u, Dg + (int64_t)i * lds + cv + u);
                    else dst[u] = T(0);
                }
            }
        }
    }
    for (int e = tid; e < BB_M * kp; e += BB_THREADS) Cblk[e] = T(0);       // rows of a short last block stay zero
    for (int e = tid; e < BB_M * ncp; e += BB_THREADS) dlt[e] = T(0);
    for (int e = tid; e < BB_M * BB_NB; e += BB_THREADS) coef[e] = T(0);
    for (int e = tid; e < BB_M * BB_M; e += BB_THREADS) Cx[e] = T(0);
    if (tid < BB_TILES) {
        int I = 0, rem = tid;
        while (rem >= 8 - I) { rem -= 8 - I; ++I; }
        tile_i[tid] = (unsigned char)I;
        tile_j[tid] = (unsigned char)(I + rem);
    }
    if (tid == 0) {
        mbar_init(xbar_addr, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();                        // ord_s, zeroed Cblk
    issue_loads(0);
    // every peer is resident and its mbarrier initialised before anybody sends
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (g == 0 && tid == 0) bcd_signal_start(P);
    cp_async_wait_all();
    __syncthreads();
    if (stamp) stamp[(int64_t)8 * k + 1] = clock64();

    // look-ahead product of the block whose C rows are in Cblk, against the shared slice as it is now (warps 1..11):
    // Rraw[j] = C[a_j,:] . D_sub.  Thread = (column pair, group of rows), BB_LA atoms per pass, FFMA2.
    auto lookahead = [&]() {
        constexpr int NW = BB_THREADS - 32;
        const int wt = tid - 32;
        const int IGW = max(1, NW / NP);
        const int pr = wt % NP, ig = wt / NP;
        const int RB = (((k + IGW - 1) / IGW) + 3) & ~3;
        for (int pass = 0; pass < BB_M / BB_LA; ++pass) {
            if (ig < IGW) {
                const int r0 = ig * RB, r1 = min(k, r0 + RB);
                const Pair<T> *dcol = reinterpret_cast<const Pair<T> *>(Ds) + pr;
                const T *Cb = Cblk + pass * BB_LA * kp;
                Pair<T> acc[BB_LA];
#pragma unroll
                for (int j = 0; j < BB_LA; ++j) acc[j].x = acc[j].y = T(0);
                int i = r0;
                for (; i + 4 <= r1; i += 4) {
                    const Pair<T> d0 = dcol[(i + 0) * NP], d1 = dcol[(i + 1) * NP], d2 = dcol[(i + 2) * NP], d3 = dcol[(i + 3) * NP];
#pragma unroll
                    for (int j = 0; j < BB_LA; ++j) {
                        const Quad<T> cq = *reinterpret_cast<const Quad<T> *>(Cb + j * kp + i);
                        pair_fma(cq.x, d0, acc[j]); pair_fma(cq.y, d1, acc[j]); pair_fma(cq.z, d2, acc[j]); pair_fma(cq.w, d3, acc[j]);
                    }
                }
                for (; i < r1; ++i) {
                    const Pair<T> d = dcol[i * NP];
#pragma unroll
                    for (int j = 0; j < BB_LA; ++j) pair_fma(Cb[j * kp + i], d, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < BB_LA; ++j)
                    *reinterpret_cast<Pair<T> *>(red + ((size_t)ig * BB_LA + j) * ncp + 2 * pr) = acc[j];
            }
            named_sync(BB_BAR_WORKERS, NW);
            for (int e = wt; e < BB_LA * ncp; e += NW) {
                T sum = T(0);
                for (int gi = 0; gi < IGW; ++gi) sum += red[(size_t)gi * BB_LA * ncp + e];   // fixed order
                Rraw[pass * BB_LA * ncp + e] = sum;
            }
            named_sync(BB_BAR_WORKERS, NW);
        }
    };

    if (wid > 0) lookahead();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");      // matched by the wait before the first send
    if (stamp) stamp[(int64_t)8 * k + 2] = clock64();

    // new rows and deltas of block pb on my columns, from the coefficients the solver left.
    // item = (column pair, pair of atoms); FFMA2 over the column pair.
    auto apply_block = [&](int pb) {
        const int mbp = min(BB_M, k - pb * BB_M);
        for (int e = tid; e < (BB_M / 2) * NP; e += BB_THREADS) {
            const int q = e / NP, pr = e % NP;
            const Pair<T> *bcol = reinterpret_cast<const Pair<T> *>(basis) + pr;
            Pair<T> acc0 = {T(0), T(0)}, acc1 = {T(0), T(0)};
#pragma unroll
            for (int r = 0; r < BB_NB; r += 4) {
                const Pair<T> b0 = bcol[(r + 0) * NP], b1 = bcol[(r + 1) * NP], b2 = bcol[(r + 2) * NP], b3 = bcol[(r + 3) * NP];
                const Quad<T> f0 = *reinterpret_cast<const Quad<T> *>(coef + (2 * q) * BB_NB + r);
                const Quad<T> f1 = *reinterpret_cast<const Quad<T> *>(coef + (2 * q + 1) * BB_NB + r);
                pair_fma(f0.x, b0, acc0); pair_fma(f0.y, b1, acc0); pair_fma(f0.z, b2, acc0); pair_fma(f0.w, b3, acc0);
                pair_fma(f1.x, b0, acc1); pair_fma(f1.y, b1, acc1); pair_fma(f1.z, b2, acc1); pair_fma(f1.w, b3, acc1);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * q + u;
                const Pair<T> nv = u ? acc1 : acc0;
                Pair<T> dv = {T(0), T(0)};
                if (j < mbp) {
                    const Pair<T> dold = bcol[(BB_M + j) * NP];
                    dv.x = nv.x - dold.x; dv.y = nv.y - dold.y;
                    *reinterpret_cast<Pair<T> *>(Ds + ord_s[pb * BB_M + j] * ncp + 2 * pr) = nv;
                }
                *reinterpret_cast<Pair<T> *>(dlt + j * ncp + 2 * pr) = dv;
            }
        }
    };

    for (int b = 0; b < nbk; ++b) {
        const int mb = min(BB_M, k - b * BB_M);
        const int par = b & 1;
        BB_STAMP(b, 0);
        // ---- S0a: small tables of this block (its C rows are resident); apply the previous block ----
        if (tid < BB_M) {
            const T d = (tid < mb) ? Cblk[tid * kp + ord_s[b * BB_M + tid]] : T(1);
            caa[tid] = d;
            rcaa[tid] = T(1) / d;
        }
        if (b > 0 && tid < BB_M * BB_M) {
            const int i = tid / BB_M, j = tid % BB_M;
            Cx[tid] = Cblk[j * kp + ord_s[(b - 1) * BB_M + i]];                 // C[a_j, a_i(prev)]
        }
        if (b > 0) apply_block(b - 1);
        __syncthreads();
        BB_STAMP(b, 1);
        // ---- S0b: repair the look-ahead product with the previous block's deltas; basis of this block ----
        for (int e = tid; e < (BB_M / 2) * NP; e += BB_THREADS) {
            const int q = e / NP, pr = e % NP;
            Pair<T> dot0 = *reinterpret_cast<const Pair<T> *>(Rraw + (2 * q) * ncp + 2 * pr);
            Pair<T> dot1 = *reinterpret_cast<const Pair<T> *>(Rraw + (2 * q + 1) * ncp + 2 * pr);
            if (b > 0) {
#pragma unroll
                for (int i = 0; i < BB_M; ++i) {
                    const Pair<T> dv = *reinterpret_cast<const Pair<T> *>(dlt + i * ncp + 2 * pr);
                    const Pair<T> cq = *reinterpret_cast<const Pair<T> *>(Cx + i * BB_M + 2 * q);
                    pair_fma(cq.x, dv, dot0);
                    pair_fma(cq.y, dv, dot1);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * q + u;
                Pair<T> gv = {T(0), T(0)}, dold = {T(0), T(0)};
                if (j < mb) {
                    const T ca = caa[j], rc = rcaa[j];
                    const Pair<T> dot = u ? dot1 : dot0;
                    dold = *reinterpret_cast<const Pair<T> *>(Ds + ord_s[b * BB_M + j] * ncp + 2 * pr);
                    const Pair<T> bv = *reinterpret_cast<const Pair<T> *>(Brow + j * ncp + 2 * pr);
                    const T gx = (bv.x - dot.x) + ca * dold.x, gy = (bv.y - dot.y) + ca * dold.y;   // [ref: :679-683]
                    T qx = gx * rc, qy = gy * rc;
                    qx = fma(fma(-qx, ca, gx), rc, qx);                            // grad / caa, Newton-corrected
                    qy = fma(fma(-qy, ca, gy), rc, qy);
                    const bool upd = ca > T(1e-20);                                // else do not update [ref: :681-683]
                    gv.x = upd ? qx : dold.x; gv.y = upd ? qy : dold.y;
                }
                *reinterpret_cast<Pair<T> *>(basis + j * ncp + 2 * pr) = gv;
                *reinterpret_cast<Pair<T> *>(basis + (BB_M + j) * ncp + 2 * pr) = dold;
            }
        }
        if (tid < BB_M * BB_M) {
            const int t = tid / BB_M, j = tid % BB_M;
            T l = T(0);
            if (j < t && t < mb && caa[t] > T(1e-20)) {
                const T c1 = Cblk[t * kp + ord_s[b * BB_M + j]], ca = caa[t], rc = rcaa[t];   // C[a_t, a_j]
                l = c1 * rc;
                l = fma(fma(-l, ca, c1), rc, l);
            }
            Lblk[tid] = l;
        }
        __syncthreads();
        BB_STAMP(b, 2);
        if (b + 1 < nbk) issue_loads(b + 1);        // Cblk / Brow of block b are consumed
        // ---- S1: partial Gram of the 32 basis vectors over my columns (4x4 register tiles, 8 column parts) ----
        if (wid < BB_TILES / 4) {
            const int ti = 4 * wid + (lane >> 3), part = lane & 7;
            const int I = tile_i[ti], J = tile_j[ti];
            const Quad<T> *ri = reinterpret_cast<const Quad<T> *>(basis + (4 * I) * ncp);
            const Quad<T> *rj = reinterpret_cast<const Quad<T> *>(basis + (4 * J) * ncp);
            const int nq = ncp >> 2;                                            // float4 groups per row
            Pair<T> acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) acc[a][bb].x = acc[a][bb].y = T(0);
            for (int f = part; f < nq; f += 8) {
                Quad<T> x[4], y[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { x[a] = ri[a * nq + f]; y[a] = rj[a * nq + f]; }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) {
                        pair_mul_fma(Pair<T>{x[a].x, x[a].y}, Pair<T>{y[bb].x, y[bb].y}, acc[a][bb]);
                        pair_mul_fma(Pair<T>{x[a].z, x[a].w}, Pair<T>{y[bb].z, y[bb].w}, acc[a][bb]);
                    }
            }
            T out[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    T v = acc[a][bb].x + acc[a][bb].y;
                    v += __shfl_xor_sync(kFullMask, v, 1);
                    v += __shfl_xor_sync(kFullMask, v, 2);
                    v += __shfl_xor_sync(kFullMask, v, 4);
                    out[a][bb] = v;
                }
            if (part == 0) {
                T *dst = stage + par * BB_GRAM + ti * 16;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    Quad<T> o;
                    o.x = out[a][0]; o.y = out[a][1]; o.z = out[a][2]; o.w = out[a][3];
                    *reinterpret_cast<Quad<T> *>(dst + 4 * a) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // visible to the bulk copies below
            }
        }
        __syncthreads();
        BB_STAMP(b, 3);
        // every peer has consumed the previous exchange (and finished the scratch use of its receive area)
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        BB_STAMP(b, 4);
        // ---- S2: all-to-all of the partial Grams ----
        if (tid == 0) mbar_expect_tx(xbar_addr, (unsigned)nblk * kGramBytes);
        if (tid < nblk) {
            const unsigned src = (unsigned)__cvta_generic_to_shared(stage + par * BB_GRAM);
            const unsigned dst = mapa_u32((unsigned)__cvta_generic_to_shared(recv + (size_t)g * BB_GRAM), (unsigned)tid);
            const unsigned bar = mapa_u32(xbar_addr, (unsigned)tid);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst), "r"(src), "r"(kGramBytes), "r"(bar) : "memory");
        }
        mbar_wait(xbar_addr, (unsigned)par);
        BB_STAMP(b, 5);
        // ---- S3: G = sum of the partials, in CTA order ----
        for (int v = tid; v < BB_GRAM; v += BB_THREADS) {
            T part[BCD_MAX_CLUSTER];
#pragma unroll
            for (int q = 0; q < BCD_MAX_CLUSTER; ++q) part[q] = q < nblk ? recv[(size_t)q * BB_GRAM + v] : T(0);
            T a0 = (part[0] + part[1]) + (part[2] + part[3]), a1 = (part[4] + part[5]) + (part[6] + part[7]);
            T a2 = (part[8] + part[9]) + (part[10] + part[11]), a3 = (part[12] + part[13]) + (part[14] + part[15]);
            const T sum = (a0 + a1) + (a2 + a3);
            const int ti = v >> 4, a = (v >> 2) & 3, bb = v & 3;
            const int r = 4 * tile_i[ti] + a, c = 4 * tile_j[ti] + bb;
            Mfull[r * BB_MLD + c] = sum;
            if (tile_i[ti] != tile_j[ti]) Mfull[c * BB_MLD + r] = sum;
        }
        cp_async_wait_all();
        __syncthreads();
        BB_STAMP(b, 6);
        // ---- S4: warp 0 solves the block's scalars; the other warps run the look-ahead product of the next block ----
        if (wid == 0) {
            T Mrow[BB_NB];
#pragma unroll
            for (int c = 0; c < BB_NB; ++c) Mrow[c] = Mfull[lane * BB_MLD + c];
            T dl[BB_M], zz[BB_M];
#pragma unroll
            for (int t = 0; t < BB_M; ++t) {
                if (t < mb) {
                    T w = (lane == t) ? T(1) : T(0);
                    T y = Mrow[t];
#pragma unroll
                    for (int j = 0; j < t; ++j) {
                        const T l = Lblk[t * BB_M + j];
                        w = fma(-l, dl[j], w);
                        y = fma(-l, zz[j], y);
                    }
                    T n2 = w * y;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(kFullMask, n2, o);
                    n2 = n2 > T(0) ? n2 : T(0);
                    const int a = ord_s[b * BB_M + t];
                    const T radius = cnorm[a] + Mfull[(BB_M + t) * BB_MLD + BB_M + t];   // comp_norm_[k] += enet_norm(old row) [ref: :676-678]
                    if (lane == 0) rad[a] = radius;
                    T rn = T(1);
                    if (radius == T(0)) {
                        rn = T(0);                                                  // [ref: enet.pyx:56-58]
                    } else {
                        const T x = n2 / radius;
                        if (x > T(1)) rn = bcd_rsqrt(x);                            // v / sqrt(|v|^2 / radius) [ref: enet.pyx:62-70]
                    }
                    const T cf = rn * w;
                    coef[t * BB_NB + lane] = cf;
                    dl[t] = cf - ((lane == BB_M + t) ? T(1) : T(0));
                    zz[t] = fma(rn, y, -Mrow[BB_M + t]);
                } else {
                    coef[t * BB_NB + lane] = T(0);
                    dl[t] = zz[t] = T(0);
                }
            }
            BB_STAMP(b, 7);
        } else if (b + 1 < nbk) {
            if (wstamp) wstamp[2 * b] = clock64();
            lookahead();
            if (wstamp) wstamp[2 * b + 1] = clock64();
        }
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (stamp) stamp[(int64_t)8 * k + 3] = clock64();
    apply_block(nbk - 1);
    __syncthreads();
    if (stamp) stamp[(int64_t)8 * k + 4] = clock64();

    // ---- epilogue: norms of the new atoms, write-back (as in bcd_pilot_kernel) ----
    T *napart = P.part;                                        // [nblk][k]
    for (int i = wid; i < k; i += BB_THREADS / 32) {
        T acc = T(0);
        for (int c = lane; c < nc; c += 32) acc += enet_term(Ds[i * ncp + c], P.l1_ratio);
        acc = warp_sum(acc);
        if (lane == 0) napart[(int64_t)g * k + i] = acc;
    }
    if (p_vec) {
        const int nv = ncp / VE;
        for (int e = tid; e < k * nv; e += BB_THREADS) {
            const int i = e / nv, cv = (e % nv) * VE;
            if (cv + VE <= nc) {
                *reinterpret_cast<uint4 *>(P.Dp + (int64_t)i * lds + c0 + cv) = *reinterpret_cast<const uint4 *>(Ds + i * ncp + cv);
            } else {
                for (int u = 0; u < VE; ++u)
                    if (cv + u < nc) P.Dp[(int64_t)i * lds + c0 + cv + u] = Ds[i * ncp + cv + u];
            }
        }
    } else {
        for (int e = tid; e < k * ncp; e += BB_THREADS) {
            const int i = e / ncp, c = e % ncp;
            if (c < nc) P.Dp[(int64_t)i * lds + c0 + c] = Ds[e];
        }
    }
    if (stamp) stamp[(int64_t)8 * k + 5] = clock64();
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (stamp) stamp[(int64_t)8 * k + 6] = clock64();
    if (g == 0) {
        for (int i = tid; i < k; i += BB_THREADS) {
            T part[BCD_MAX_CLUSTER];
#pragma unroll
            for (int q = 0; q < BCD_MAX_CLUSTER; ++q) part[q] = q < nblk ? __ldcg(napart + (int64_t)q * k + i) : T(0);   // all in flight
            T na = T(0);
#pragma unroll
            for (int q = 0; q < BCD_MAX_CLUSTER; ++q) na += part[q];                     // fixed order
            P.comp_norm[i] = rad[i] - na;                                                // [ref: :690-692]
        }
    }
    if (stamp) stamp[(int64_t)8 * k + 7] = clock64();
#undef BB_STAMP
}

}  // namespace modl
