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
constexpr int BB_GRAM = BB_TILES * 16;   // entries of the (tile-packed) Gram
constexpr int BB_MLD = BB_NB + 1;        // row pitch of the Gram matrix in shared memory
constexpr int BB_RS_SLOTS = 52;          // (owned tile, sender) slots of the reduce-scatter: ceil(36 / n) * n <= 51 for n <= 16
constexpr int BB_RED_CAP = 9216;         // scratch of the look-ahead product (elements)
enum { BB_BAR_WORKERS = 1 };
constexpr int BB_SOLVER = BB_THREADS / 32 - 1;   // the solver warp: the HIGHEST warp id (the issue arbiter favours high ids)
constexpr int BB_NWORK = 256;            // warps 0..7: the look-ahead product (two per scheduler), behind the solver
constexpr int BB_NPH = BB_THREADS;       // every thread takes part in the per-block phases
// 1 / sqrt(x) for x >= 1: approximation + one Newton step, no denormal handling
__device__ __forceinline__ float bb_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return fmaf(0.5f * y, fmaf(-x * y, y, 1.f), y);
}
__device__ __forceinline__ double bb_rsqrt(double x) { return 1.0 / sqrt(x); }

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
// TMA bulk copies: global -> shared (completes bytes on an mbarrier), shared -> global (bulk group)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned bar)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, unsigned bytes)
{
    const unsigned src = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(src), "r"(bytes) : "memory");
}
// four values into a peer's shared memory, completing their bytes on the PEER's mbarrier (one-way, no fences)
__device__ __forceinline__ void st_async_quad(unsigned remote, unsigned remote_mbar, float a, float b, float c, float d)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                 ::"r"(remote), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)),
                   "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async_quad(unsigned remote, unsigned remote_mbar, double a, double b, double c, double d)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];\n"
                 ::"r"(remote), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(remote_mbar) : "memory");
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];\n"
                 ::"r"(remote + 16u), "l"(__double_as_longlong(c)), "l"(__double_as_longlong(d)), "r"(remote_mbar) : "memory");
}

// shared-memory footprint in bytes -- must match the carve-up in the kernel
template <typename T>
__host__ __device__ inline size_t bcd_blocked_smem_bytes(int64_t k, int64_t ncp)
{
    const int64_t kp = round_up(k, 32);
    const int64_t elems = k * ncp                 // Ds
                          + 2 * kp * BB_M         // Cblk (two blocks)
                          + 3 * BB_M * ncp        // Rraw (two blocks), Brow
                          + BB_NB * ncp           // basis
                          + BB_M * ncp            // dlt
                          + BB_M * BB_NB          // coef
                          + 2 * BB_M * BB_M       // LT, Cx
                          + BB_NB * BB_MLD + 32   // Mfull (padded to a multiple of 4)
                          + kp                    // cnorm
                          + 2 * BB_M              // caa, rcaa
                          + BB_GRAM               // Mflat
                          + 2 * BB_RS_SLOTS * 16  // rsrecv (by block parity)
                          + BB_RED_CAP;           // red
    return (size_t)elems * sizeof(T) + (size_t)kp * sizeof(int) + 64;
}

// the look-ahead product needs at least one row group of 16 x ncp partial sums in its scratch
__host__ __device__ inline bool bcd_blocked_fits(int64_t ncp) { return BB_M * ncp <= BB_RED_CAP; }

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
    const int NQ = ncp >> 2;                                 // groups of 4 columns of the slice
    const bool worker = tid < BB_NWORK;
    const int pt = tid;                                      // index among the phase threads (all of them)

    // ---- shared memory carve-up (see bcd_blocked_smem_bytes) ----
    T *Ds = reinterpret_cast<T *>(bb_smem_raw);              // [k][ncp]     the CTA's slice of D_sub
    T *Cblk = Ds + (size_t)k * ncp;                          // [2][M][kp]   C[a_j, :] by block parity
    T *Rraw = Cblk + 2 * kp * BB_M;                          // [2][M][ncp]  look-ahead product C[a_j,:] . D_sub, by block parity
    T *Brow = Rraw + 2 * BB_M * ncp;                         // [M][ncp]     B_sub rows of the block
    T *basis = Brow + BB_M * ncp;                            // [2M][ncp]    g_1..g_M, d_1..d_M
    T *dlt = basis + BB_NB * ncp;                            // [M][ncp]     deltas of the block just applied
    T *coef = dlt + BB_M * ncp;                              // [M][2M]      n_t = sum_r coef[t][r] basis_r
    T *LT = coef + BB_M * BB_NB;                             // [M][M]       L_tj at [j*M + t] (j < t), zero elsewhere
    T *Cx = LT + BB_M * BB_M;                                // [M][M]       C[a_j, a_i(previous block)] at [i*M + j]
    T *Mfull = Cx + BB_M * BB_M;                             // [2M][2M+1]   Gram matrix of the basis
    T *cnorm = Mfull + BB_NB * BB_MLD + 32;                  // [kp]         comp_norm_ on entry
    T *caa = cnorm + kp;                                     // [M]          C[a_t, a_t]
    T *rcaa = caa + BB_M;                                    // [M]          1 / C[a_t, a_t]
    T *Mflat = rcaa + BB_M;                                  // [36][16]     the summed Gram, tile-packed (all-gather target)
    T *rsrecv = Mflat + BB_GRAM;                             // [2][slot][sender][16] partial tiles this CTA sums, by block parity
    T *red = rsrecv + 2 * BB_RS_SLOTS * 16;                  // scratch of the look-ahead product
    int *ord_s = reinterpret_cast<int *>(red + BB_RED_CAP);  // [kp] update order
    __shared__ __align__(8) unsigned long long xbar[3];      // [0] reduce-scatter arrivals, [1] all-gather arrivals, [2] operand copies
    __shared__ unsigned char tile_i[BB_TILES], tile_j[BB_TILES];
    const unsigned bar1 = (unsigned)__cvta_generic_to_shared(&xbar[0]);
    const unsigned bar2 = bar1 + 8;
    const unsigned ldbar = bar1 + 16;
    unsigned ld_batches = 0;                                  // operand batches issued so far (phase of ldbar)
    const int n_owned = (BB_TILES - g + nblk - 1) / nblk;    // tiles ti with ti % nblk == g  (g < nblk <= 16 < 36)

    // debug stamps of CTA 0.  First phase thread: [b][0..6]; lane 0 of the solver warp: [b][7] = solver done;
    // thread 0 (a worker): look-ahead start / end per block at 8k + 16 + 4b; kernel-level stamps at 8k .. 8k + 7
    long long *stamp = (P.timing && g == 0 && tid == 0) ? P.timing : nullptr;
    long long *wstamp = (P.timing && g == 0 && tid == 32) ? P.timing + (int64_t)8 * k + 16 : nullptr;
    long long *sstamp = (P.timing && g == 0 && tid == 32 * BB_SOLVER) ? P.timing : nullptr;
#define BB_STAMP(b_, slot_) do { if (stamp) stamp[(int64_t)(b_) * 8 + (slot_)] = clock64(); } while (0)
    // a stamp right after a barrier records when the warp ARRIVED (the barrier blocks at the next dependent access), so
    // the release time is taken after a volatile shared load
#define BB_REL_STAMP(b_, slot_) do { if (stamp) { const int v_ = *reinterpret_cast<volatile int *>(ord_s); \
        stamp[(int64_t)(b_) * 8 + (slot_)] = clock64() + (v_ & 0); } } while (0)
    if (stamp) stamp[(int64_t)8 * k] = clock64();

    constexpr int VE = 16 / (int)sizeof(T);                  // elements per 16-byte copy
    const bool c_vec = (k % VE == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
    const bool p_vec = (lds % VE == 0) && (c0 % VE == 0) && ((reinterpret_cast<uintptr_t>(P.Dp) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(P.Bp) & 15) == 0);
    // whole rows move with TMA bulk copies (one instruction per row) when every row start and length is 16-byte aligned
    const bool bulk = c_vec && p_vec && nc > 0 && (nc % VE == 0);
    const unsigned row_bytes = (unsigned)nc * (unsigned)sizeof(T), crow_bytes = (unsigned)k * (unsigned)sizeof(T);

    // asynchronous operand copies, spread over `nth` threads (index `t`): C[a_j, :] of block bc, B_sub rows of block bb
    auto load_C = [&](int bc, int t, int nth) {
        const int mb = min(BB_M, k - bc * BB_M);
        T *dstC = Cblk + (bc & 1) * kp * BB_M;
        if (bulk) {
            if (t < mb) bulk_g2s(dstC + t * kp, P.C + (int64_t)ord_s[bc * BB_M + t] * k, crow_bytes, ldbar);
        } else if (c_vec) {
            const int nv = k / VE;
            for (int e = t; e < BB_M * nv; e += nth) {
                const int j = e / nv, iv = (e % nv) * VE;
                if (j < mb) cp_async_16(dstC + j * kp + iv, P.C + (int64_t)ord_s[bc * BB_M + j] * k + iv);
            }
        } else {
            for (int e = t; e < BB_M * k; e += nth) {
                const int j = e / k, i = e % k;
                if (j < mb) cp_async_elem(dstC + j * kp + i, P.C + (int64_t)ord_s[bc * BB_M + j] * k + i);
            }
        }
    };
    auto load_B = [&](int bb, int t, int nth) {
        const int mb = min(BB_M, k - bb * BB_M);
        const int nvb = ncp / VE;
        if (bulk) {         // (the padding columns were zeroed once in the prologue)
            const int j = t - 32;
            if (j >= 0 && j < mb) bulk_g2s(Brow + j * ncp, P.Bp + (int64_t)ord_s[bb * BB_M + j] * lds + c0, row_bytes, ldbar);
            return;
        }
        for (int e = t; e < BB_M * nvb; e += nth) {
            const int j = e / nvb, cv = (e % nvb) * VE;
            T *dst = Brow + j * ncp + cv;
            if (j < mb && p_vec && cv + VE <= nc) {
                cp_async_16(dst, P.Bp + (int64_t)ord_s[bb * BB_M + j] * lds + c0 + cv);
            } else {
#pragma unroll
                for (int u = 0; u < VE; ++u) {
                    if (j < mb && cv + u < nc) cp_async_elem(dst + u, P.Bp + (int64_t)ord_s[bb * BB_M + j] * lds + c0 + cv + u);
                    else dst[u] = T(0);
                }
            }
        }
    };

    // one batch of operand copies = one phase of ldbar (bulk mode) or one cp.async group (otherwise)
    auto arm_batch = [&](unsigned bytes) {
        if (bulk && tid == 0) mbar_expect_tx(ldbar, bytes);
    };
    auto wait_batch = [&]() {
        if (bulk) { mbar_wait(ldbar, ld_batches & 1u); }
        else cp_async_wait_all();
        ld_batches += 1;
    };

    // ---- prologue: D slice (asynchronous copies), norms, order, tables ----
    for (int i = tid; i < kp; i += BB_THREADS) {
        cnorm[i] = i < k ? P.comp_norm[i] : T(0);
        ord_s[i] = i < k ? P.order[i] : 0;
    }
    if (bulk) {
        for (int e = tid; e < k * (ncp - nc); e += BB_THREADS) {                // padding columns of the slice and of the B rows
            const int i = e / (ncp - nc), c = nc + e % (ncp - nc);
            Ds[i * ncp + c] = T(0);
            if (i < BB_M) Brow[i * ncp + c] = T(0);
        }
    } else {
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
    for (int e = tid; e < 2 * BB_M * kp; e += BB_THREADS) Cblk[e] = T(0);   // rows of a short last block stay zero
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
        mbar_init(bar1, 1);
        mbar_init(bar2, 1);
        mbar_init(ldbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the zero fills above precede the bulk copies below
    __syncthreads();                        // ord_s, zeroed Cblk
    {
        const int mb0 = min(BB_M, k), mb1 = nbk > 1 ? min(BB_M, k - BB_M) : 0;
        arm_batch((unsigned)k * row_bytes + (unsigned)(mb0 + mb1) * crow_bytes + (unsigned)mb0 * row_bytes);
        if (bulk)
            for (int i = tid; i < k; i += BB_THREADS) bulk_g2s(Ds + i * ncp, P.Dp + (int64_t)i * lds + c0, row_bytes, ldbar);
    }
    load_C(0, tid, BB_THREADS);
    if (nbk > 1) load_C(1, tid, BB_THREADS);
    load_B(0, tid, BB_THREADS);
    cp_async_commit();
    // every peer is resident and its mbarriers initialised before anybody sends
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (g == 0 && tid == 0) bcd_signal_start(P);
    wait_batch();
    __syncthreads();
    if (stamp) stamp[(int64_t)8 * k + 1] = clock64();

    // Look-ahead product of block bl (worker warps): Rraw[bl & 1][j] = C[a_j,:] . D_sub against the shared slice as
    // it is now.  A thread owns FOUR columns x all 16 atoms x a group of rows: per 4 rows 4 + 16 128-bit shared loads
    // feed 128 FFMA2 (the fma pipe issues one FFMA2 per ~2.5 cycles and scheduler, a broadcast 128-bit load costs ~2.5
    // cycles of the shared-memory port: the loop is fma-bound).  Row groups are combined through a scratch that holds
    // half of them: the first half stores, the second half adds in place, then the sums are added in group order.
    auto lookahead = [&](int bl) {
        constexpr int NW = BB_NWORK;
        const T *Cb = Cblk + (bl & 1) * kp * BB_M;
        T *Rb = Rraw + (bl & 1) * BB_M * ncp;
        if constexpr (sizeof(T) == 8) {
            // double: two columns per thread (the four-column tile would not fit the register file), one round
            const int NP = ncp >> 1;
            const int IGW = max(1, min(NW / NP, BB_RED_CAP / (BB_M * ncp)));
            const int pr = tid % NP, ig = tid / NP;
            const int RB = (((k + IGW - 1) / IGW) + 3) & ~3;
            if (ig < IGW) {
                const int r0 = min(k, ig * RB), r1 = min(k, r0 + RB);
                const Pair<T> *dcol = reinterpret_cast<const Pair<T> *>(Ds) + pr;
                Pair<T> acc[BB_M];
#pragma unroll
                for (int j = 0; j < BB_M; ++j) acc[j].x = acc[j].y = T(0);
                for (int i = r0; i < r1; ++i) {
                    const Pair<T> d = dcol[i * NP];
#pragma unroll
                    for (int j = 0; j < BB_M; ++j) pair_fma(Cb[j * kp + i], d, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < BB_M; ++j)
                    *reinterpret_cast<Pair<T> *>(red + ((size_t)ig * BB_M + j) * ncp + 2 * pr) = acc[j];
            }
            named_sync(BB_BAR_WORKERS, NW);
            for (int e = tid; e < BB_M * ncp; e += NW) {
                T sum = T(0);
                for (int gi = 0; gi < IGW; ++gi) sum += red[(size_t)gi * BB_M * ncp + e];   // fixed order
                Rb[e] = sum;
            }
            named_sync(BB_BAR_WORKERS, NW);
            return;
        }
        if (ncp == 96) {
            // The benchmark width (k x 1250 over 16 CTAs).  One WARP per group of rows; lane = (half of the atoms, six
            // columns): 8 atoms x 3 column pairs per thread, so that the eight warps carry exactly equal shares, two per
            // scheduler (the fma pipe issues one FFMA2 per ~2.5 cycles and scheduler: 6144 warp-FFMA2 per block).
            const int ig = wid, ah = lane >> 4, sx = lane & 15;
            const int RB = (((k + 7) / 8) + 3) & ~3;
            const int r0 = min(k, ig * RB), r1 = min(k, r0 + RB);
            const T *dcol = Ds + 6 * sx;
            const T *Ca = Cb + (8 * ah) * kp;
            Pair<T> acc[8][3];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[j][c].x = acc[j][c].y = T(0);
            int i = r0;
            for (; i + 4 <= r1; i += 4) {
                Pair<T> d[4][3];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < 3; ++c) d[u][c] = *reinterpret_cast<const Pair<T> *>(dcol + (i + u) * ncp + 2 * c);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const Quad<T> c4 = *reinterpret_cast<const Quad<T> *>(Ca + j * kp + i);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        pair_fma(c4.x, d[0][c], acc[j][c]); pair_fma(c4.y, d[1][c], acc[j][c]);
                        pair_fma(c4.z, d[2][c], acc[j][c]); pair_fma(c4.w, d[3][c], acc[j][c]);
                    }
                }
            }
            for (; i < r1; ++i) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const Pair<T> dv = *reinterpret_cast<const Pair<T> *>(dcol + i * ncp + 2 * c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) pair_fma(Ca[j * kp + i], dv, acc[j][c]);
                }
            }
            // row groups 0..3 store, 4..7 add in place, then the four sums are added in group order (fixed order)
            T *slot = red + ((size_t)(ig & 3) * BB_M + 8 * ah) * ncp + 6 * sx;
            if (ig < 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) *reinterpret_cast<Pair<T> *>(slot + j * ncp + 2 * c) = acc[j][c];
            }
            named_sync(BB_BAR_WORKERS, NW);
            if (ig >= 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        Pair<T> o = *reinterpret_cast<Pair<T> *>(slot + j * ncp + 2 * c);
                        o.x += acc[j][c].x; o.y += acc[j][c].y;
                        *reinterpret_cast<Pair<T> *>(slot + j * ncp + 2 * c) = o;
                    }
            }
            named_sync(BB_BAR_WORKERS, NW);
            for (int e = tid; e < BB_M * ncp; e += NW) {
                T sum = red[e];
#pragma unroll
                for (int gi = 1; gi < 4; ++gi) sum += red[(size_t)gi * BB_M * ncp + e];
                Rb[e] = sum;
            }
            named_sync(BB_BAR_WORKERS, NW);
            return;
        }
        const int half = max(1, BB_RED_CAP / (BB_M * ncp));
        const int IGW = max(1, min(NW / NQ, 2 * half));
        const int cq = tid % NQ, ig = tid / NQ;
        const int RB = (((k + IGW - 1) / IGW) + 3) & ~3;
        Pair<T> acc[BB_M][2];
        const bool on = ig < IGW;
        if (on) {
            const int r0 = min(k, ig * RB), r1 = min(k, r0 + RB);
            const Quad<T> *dcol = reinterpret_cast<const Quad<T> *>(Ds) + cq;
#pragma unroll
            for (int j = 0; j < BB_M; ++j) acc[j][0].x = acc[j][0].y = acc[j][1].x = acc[j][1].y = T(0);
            int i = r0;
            for (; i + 4 <= r1; i += 4) {
                const Quad<T> d0 = dcol[(i + 0) * NQ], d1 = dcol[(i + 1) * NQ], d2 = dcol[(i + 2) * NQ], d3 = dcol[(i + 3) * NQ];
#pragma unroll
                for (int h = 0; h < BB_M; h += 8) {
                    Quad<T> c4[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) c4[j] = *reinterpret_cast<const Quad<T> *>(Cb + (h + j) * kp + i);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        pair_fma(c4[j].x, Pair<T>{d0.x, d0.y}, acc[h + j][0]); pair_fma(c4[j].x, Pair<T>{d0.z, d0.w}, acc[h + j][1]);
                        pair_fma(c4[j].y, Pair<T>{d1.x, d1.y}, acc[h + j][0]); pair_fma(c4[j].y, Pair<T>{d1.z, d1.w}, acc[h + j][1]);
                        pair_fma(c4[j].z, Pair<T>{d2.x, d2.y}, acc[h + j][0]); pair_fma(c4[j].z, Pair<T>{d2.z, d2.w}, acc[h + j][1]);
                        pair_fma(c4[j].w, Pair<T>{d3.x, d3.y}, acc[h + j][0]); pair_fma(c4[j].w, Pair<T>{d3.z, d3.w}, acc[h + j][1]);
                    }
                }
            }
            for (; i < r1; ++i) {
                const Quad<T> d = dcol[i * NQ];
#pragma unroll
                for (int j = 0; j < BB_M; ++j) {
                    const T c = Cb[j * kp + i];
                    pair_fma(c, Pair<T>{d.x, d.y}, acc[j][0]); pair_fma(c, Pair<T>{d.z, d.w}, acc[j][1]);
                }
            }
            if (ig < half) {
#pragma unroll
                for (int j = 0; j < BB_M; ++j) {
                    Quad<T> o; o.x = acc[j][0].x; o.y = acc[j][0].y; o.z = acc[j][1].x; o.w = acc[j][1].y;
                    *reinterpret_cast<Quad<T> *>(red + ((size_t)ig * BB_M + j) * ncp + 4 * cq) = o;
                }
            }
        }
        named_sync(BB_BAR_WORKERS, NW);
        if (on && ig >= half) {
#pragma unroll
            for (int j = 0; j < BB_M; ++j) {
                Quad<T> *slot = reinterpret_cast<Quad<T> *>(red + ((size_t)(ig - half) * BB_M + j) * ncp + 4 * cq);
                Quad<T> o = *slot;
                o.x += acc[j][0].x; o.y += acc[j][0].y; o.z += acc[j][1].x; o.w += acc[j][1].y;
                *slot = o;
            }
        }
        named_sync(BB_BAR_WORKERS, NW);
        const int nsum = min(IGW, half);
        for (int e = tid; e < BB_M * ncp; e += NW) {
            T sum = T(0);
            for (int gi = 0; gi < nsum; ++gi) sum += red[(size_t)gi * BB_M * ncp + e];   // fixed order
            Rb[e] = sum;
        }
        named_sync(BB_BAR_WORKERS, NW);        // the scratch may be reused
    };

    if (worker) lookahead(0);
    __syncthreads();
    if (stamp) stamp[(int64_t)8 * k + 2] = clock64();

    // new rows and deltas of block pb on my columns, from the coefficients the solver left:  n_j = sum_r coef[j][r] basis_r.
    // item = (4 columns, 2 atoms): 192 threads; 6 128-bit loads feed 16 FFMA2.
    auto apply_block = [&](int pb) {
        const int mbp = min(BB_M, k - pb * BB_M);
        for (int e = pt; e < 8 * NQ; e += BB_NPH) {
            const int ap = e / NQ, cq = e % NQ;
            const Quad<T> *bcol = reinterpret_cast<const Quad<T> *>(basis) + cq;
            Pair<T> acc[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) acc[u][0].x = acc[u][0].y = acc[u][1].x = acc[u][1].y = T(0);
#pragma unroll
            for (int r = 0; r < BB_NB; r += 4) {
                const Quad<T> b0 = bcol[(r + 0) * NQ], b1 = bcol[(r + 1) * NQ], b2 = bcol[(r + 2) * NQ], b3 = bcol[(r + 3) * NQ];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const Quad<T> f = *reinterpret_cast<const Quad<T> *>(coef + (2 * ap + u) * BB_NB + r);
                    pair_fma(f.x, Pair<T>{b0.x, b0.y}, acc[u][0]); pair_fma(f.x, Pair<T>{b0.z, b0.w}, acc[u][1]);
                    pair_fma(f.y, Pair<T>{b1.x, b1.y}, acc[u][0]); pair_fma(f.y, Pair<T>{b1.z, b1.w}, acc[u][1]);
                    pair_fma(f.z, Pair<T>{b2.x, b2.y}, acc[u][0]); pair_fma(f.z, Pair<T>{b2.z, b2.w}, acc[u][1]);
                    pair_fma(f.w, Pair<T>{b3.x, b3.y}, acc[u][0]); pair_fma(f.w, Pair<T>{b3.z, b3.w}, acc[u][1]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ap + u;
                Quad<T> nv; nv.x = acc[u][0].x; nv.y = acc[u][0].y; nv.z = acc[u][1].x; nv.w = acc[u][1].y;
                Quad<T> dv; dv.x = dv.y = dv.z = dv.w = T(0);
                if (j < mbp) {
                    const Quad<T> dold = bcol[(BB_M + j) * NQ];
                    dv.x = nv.x - dold.x; dv.y = nv.y - dold.y; dv.z = nv.z - dold.z; dv.w = nv.w - dold.w;
                    *reinterpret_cast<Quad<T> *>(Ds + ord_s[pb * BB_M + j] * ncp + 4 * cq) = nv;
                }
                *reinterpret_cast<Quad<T> *>(dlt + j * ncp + 4 * cq) = dv;
            }
        }
    };

    for (int b = 0; b < nbk; ++b) {
        const int mb = min(BB_M, k - b * BB_M);
        const unsigned par = (unsigned)(b & 1);
        const T *Cb = Cblk + (b & 1) * kp * BB_M;
        BB_REL_STAMP(b, 0);
        // ---- S0a: small tables of this block, apply the previous block ----
        {
            if (pt >= BB_NPH - BB_M) {
                const int t = pt - (BB_NPH - BB_M);
                const T d = (t < mb) ? Cb[t * kp + ord_s[b * BB_M + t]] : T(1);
                caa[t] = d;
                rcaa[t] = T(1) / d;
            }
            if (b > 0) {
                for (int e = pt; e < BB_M * BB_M; e += BB_NPH) {
                    const int i = e / BB_M, j = e % BB_M;
                    Cx[e] = Cb[j * kp + ord_s[(b - 1) * BB_M + i]];              // C[a_j, a_i(prev)]
                }
                apply_block(b - 1);
            }
        }
        __syncthreads();                    // the shared slice now holds every block before b
        BB_REL_STAMP(b, 1);
        // ---- S0b: repair the look-ahead product with the previous block's deltas; basis of this block.
        //      item = (4 columns, 2 atoms) ----
        const T *Rb = Rraw + (b & 1) * BB_M * ncp;
        for (int e = pt; e < 8 * NQ; e += BB_NPH) {
            const int ap = e / NQ, cq = e % NQ;
            Pair<T> dot[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const Quad<T> rr = *reinterpret_cast<const Quad<T> *>(Rb + (2 * ap + u) * ncp + 4 * cq);
                dot[u][0].x = rr.x; dot[u][0].y = rr.y; dot[u][1].x = rr.z; dot[u][1].y = rr.w;
            }
            if (b > 0) {
#pragma unroll
                for (int i = 0; i < BB_M; ++i) {
                    const Quad<T> dv = *reinterpret_cast<const Quad<T> *>(dlt + i * ncp + 4 * cq);
                    const Pair<T> cx = *reinterpret_cast<const Pair<T> *>(Cx + i * BB_M + 2 * ap);
                    pair_fma(cx.x, Pair<T>{dv.x, dv.y}, dot[0][0]); pair_fma(cx.x, Pair<T>{dv.z, dv.w}, dot[0][1]);
                    pair_fma(cx.y, Pair<T>{dv.x, dv.y}, dot[1][0]); pair_fma(cx.y, Pair<T>{dv.z, dv.w}, dot[1][1]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = 2 * ap + u;
                Quad<T> gv, dold;
                gv.x = gv.y = gv.z = gv.w = dold.x = dold.y = dold.z = dold.w = T(0);
                if (j < mb) {
                    const T ca = caa[j], rc = rcaa[j];
                    const bool upd = ca > T(1e-20);                                // else do not update [ref: :681-683]
                    dold = *reinterpret_cast<const Quad<T> *>(Ds + ord_s[b * BB_M + j] * ncp + 4 * cq);
                    const Quad<T> bv = *reinterpret_cast<const Quad<T> *>(Brow + j * ncp + 4 * cq);
                    auto cand = [&](T bvv, T dotv, T dv) {
                        const T gr = (bvv - dotv) + ca * dv;                       // [ref: :679-683]
                        T qv = gr * rc;
                        qv = fma(fma(-qv, ca, gr), rc, qv);                        // grad / caa, Newton-corrected
                        return upd ? qv : dv;
                    };
                    gv.x = cand(bv.x, dot[u][0].x, dold.x); gv.y = cand(bv.y, dot[u][0].y, dold.y);
                    gv.z = cand(bv.z, dot[u][1].x, dold.z); gv.w = cand(bv.w, dot[u][1].y, dold.w);
                }
                *reinterpret_cast<Quad<T> *>(basis + j * ncp + 4 * cq) = gv;
                *reinterpret_cast<Quad<T> *>(basis + (BB_M + j) * ncp + 4 * cq) = dold;
            }
        }
        for (int e = pt; e < BB_M * BB_M; e += BB_NPH) {
            const int t = e / BB_M, j = e % BB_M;
            T l = T(0);
            if (j < t && t < mb && caa[t] > T(1e-20)) {
                const T c1 = Cb[t * kp + ord_s[b * BB_M + j]], ca = caa[t], rc = rcaa[t];   // C[a_t, a_j]
                l = c1 * rc;
                l = fma(fma(-l, ca, c1), rc, l);
            }
            LT[j * BB_M + t] = l;
        }
        if (pt == BB_NPH - 1) {     // arm this block's two exchanges (bytes this CTA will receive)
            mbar_expect_tx(bar1, (unsigned)(n_owned * nblk * 16 * (int)sizeof(T)));
            mbar_expect_tx(bar2, (unsigned)(BB_GRAM * (int)sizeof(T)));
        }
        __syncthreads();
        BB_REL_STAMP(b, 2);
        // operands ahead: this block's C rows and B rows are consumed
        const bool batch = b + 1 < nbk;
        if (batch) {
            arm_batch((b + 2 < nbk ? (unsigned)min(BB_M, k - (b + 2) * BB_M) * crow_bytes : 0u) +
                      (unsigned)min(BB_M, k - (b + 1) * BB_M) * row_bytes);
            if (b + 2 < nbk) load_C(b + 2, pt, BB_NPH);
            load_B(b + 1, pt, BB_NPH);
            cp_async_commit();
        }
        // ---- S1: partial Gram of the 32 basis vectors over my columns: 4x4 register tiles x 8 column parts (288 threads),
        //          parts combined with three shuffle levels (measured against a shared-memory combine + block barrier on the
        //          same box: 2 980 vs 3 670 cycles for the phase), then lane `part` of a tile's group pushes row `part` of
        //          the tile into the CTA that sums the tile (reduce-scatter over distributed shared memory) ----
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
            if (part < 4) {         // lane `part` of the tile's group sends row `part` of the tile
                T q0 = out[0][0], q1 = out[0][1], q2 = out[0][2], q3 = out[0][3];
#pragma unroll
                for (int a = 1; a < 4; ++a)
                    if (part == a) { q0 = out[a][0]; q1 = out[a][1]; q2 = out[a][2]; q3 = out[a][3]; }
                const int owner = ti % nblk, slot = ti / nblk;
                const unsigned local = (unsigned)__cvta_generic_to_shared(rsrecv + par * (BB_RS_SLOTS * 16) + ((slot * nblk + g) * 16 + 4 * part));
                st_async_quad(mapa_u32(local, (unsigned)owner), mapa_u32(bar1, (unsigned)owner), q0, q1, q2, q3);
            }
        }
        BB_STAMP(b, 3);
        // ---- S2: the owner of a tile sums its 16 partials (fixed order) and pushes the sums to every CTA (all-gather).
        //          item = (row of an owned tile, destination CTA): each sums its four values itself, no hand-over ----
        if (pt < n_owned * 4 * nblk) {
            const int qd = pt / nblk, pe = pt % nblk;
            const int slot = qd >> 2, a = qd & 3;
            mbar_wait(bar1, par);
            const T *rsb = rsrecv + par * (BB_RS_SLOTS * 16);
            Quad<T> sum = *reinterpret_cast<const Quad<T> *>(rsb + ((slot * nblk + 0) * 16 + 4 * a));
            for (int q = 1; q < nblk; ++q) {
                const Quad<T> v = *reinterpret_cast<const Quad<T> *>(rsb + ((slot * nblk + q) * 16 + 4 * a));
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            const int ti = slot * nblk + g;
            const unsigned local = (unsigned)__cvta_generic_to_shared(Mflat + ti * 16 + 4 * a);
            st_async_quad(mapa_u32(local, (unsigned)pe), mapa_u32(bar2, (unsigned)pe), sum.x, sum.y, sum.z, sum.w);
        }
        BB_STAMP(b, 4);
        mbar_wait(bar2, par);
        BB_STAMP(b, 5);
        // ---- S3: unpack the tiles into the symmetric matrix ----
        for (int v = pt; v < BB_GRAM; v += BB_NPH) {
            const T val = Mflat[v];
            const int ti = v >> 4, a = (v >> 2) & 3, bb = v & 3;
            const int r = 4 * tile_i[ti] + a, c = 4 * tile_j[ti] + bb;
            Mfull[r * BB_MLD + c] = val;
            if (tile_i[ti] != tile_j[ti]) Mfull[c * BB_MLD + r] = val;
        }
        if (batch) wait_batch();
        __syncthreads();
        BB_REL_STAMP(b, 6);
        // ---- S4: the solver warp runs the block's scalar recurrence; warps 0..7 run the look-ahead product of the next
        //          block meanwhile (the shared slice holds every block before b: what the product wants).
        //          Solver: lane r holds component r of the coefficient vectors; w_t and G w_t are kept up to date for every
        //          pending atom by a right-looking rank-1 update, so the dependent chain of an atom is one warp reduction
        //          (five shuffle levels), the clamp, one rsqrt + Newton step and two FMAs.  (A variant that moved the
        //          reduction off the chain with the three-term expansion of |v_t|^2 -- two reductions in flight, as in the
        //          pilot kernel -- was slower in place: 8 650 vs 7 370 cycles per block; a warp issues in order, so the
        //          second reduction's shuffles queue in front of the scalar chain instead of beside it.) ----
        if (worker && b + 1 < nbk) {
            if (wstamp) wstamp[4 * b] = clock64();
            lookahead(b + 1);
            if (wstamp) wstamp[4 * b + 1] = clock64();
        }
        if (wid == BB_SOLVER) {
            T Mrow[BB_NB];
#pragma unroll
            for (int c = 0; c < BB_NB; ++c) Mrow[c] = Mfull[lane * BB_MLD + c];
            // radius of atom t (lane t):  comp_norm_[k] += enet_norm(old row)  [ref: :676-678]; |d_t|^2 is a Gram diagonal
            const int a_l = ord_s[b * BB_M + (lane & (BB_M - 1))];
            const T rad_l = cnorm[a_l] + Mfull[(BB_M + (lane & (BB_M - 1))) * BB_MLD + BB_M + (lane & (BB_M - 1))];
            const T rinv_l = rad_l != T(0) ? T(1) / rad_l : T(0);
            T wacc[BB_M], yacc[BB_M];      // lane r's component of w_t and of G w_t, accumulated right-looking
            T cn_mine = T(0);              // lane t keeps the new comp_norm_ of atom t: ONE store per block after the chain (a
                                           // store per atom put a shared load -> address -> store sequence in front of every
                                           // atom's reduction: a warp issues in order)
#pragma unroll
            for (int t = 0; t < BB_M; ++t) { wacc[t] = (lane == t) ? T(1) : T(0); yacc[t] = Mrow[t]; }
#pragma unroll
            for (int t = 0; t < BB_M; ++t) {
                const T w = wacc[t], y = yacc[t];
                T n2 = w * y;                                       // |v_t|^2 = w^T G w
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(kFullMask, n2, o);
                n2 = n2 > T(0) ? n2 : T(0);
                const T radius = __shfl_sync(kFullMask, rad_l, t), rinv = __shfl_sync(kFullMask, rinv_l, t);
                const T x = n2 * rinv;
                const T rs = bb_rsqrt(x > T(1) ? x : T(1));
                T rn = x > T(1) ? rs : T(1);                        // v / sqrt(|v|^2 / radius) outside the ball [ref: enet.pyx:62-70]
                rn = radius == T(0) ? T(0) : rn;                    // [ref: enet.pyx:56-58]
                const T cf = rn * w;
                const T dlv = cf - ((lane == BB_M + t) ? T(1) : T(0));       // delta_t = n_t - d_t
                const T zv = fma(rn, y, -Mrow[BB_M + t]);                    // G delta_t
#pragma unroll
                for (int t2 = t + 1; t2 < BB_M; ++t2) {
                    const T l = LT[t * BB_M + t2];                           // L_{t2, t}
                    wacc[t2] = fma(-l, dlv, wacc[t2]);
                    yacc[t2] = fma(-l, zv, yacc[t2]);
                }
                coef[t * BB_NB + lane] = cf;
                // comp_norm_[k] -= enet_norm(new row) [ref: :690-692]: |n_t|^2 = rn^2 |v_t|^2 (the same value in every CTA)
                const T cn_t = radius - (rn * rn) * n2;
                cn_mine = (lane == t) ? cn_t : cn_mine;
            }
            if (g == 0 && lane < mb) P.comp_norm[ord_s[b * BB_M + lane]] = cn_mine;
            if (sstamp) sstamp[(int64_t)b * 8 + 7] = clock64();
        }
        __syncthreads();                    // coefficients, look-ahead product of block b+1, operands of blocks b+1 / b+2
    }
    if (stamp) stamp[(int64_t)8 * k + 3] = clock64();
    apply_block(nbk - 1);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the rows written above are read by the bulk stores
    __syncthreads();
    if (stamp) stamp[(int64_t)8 * k + 4] = clock64();

    // ---- epilogue: write the slice back ----
    if (bulk) {
        for (int i = tid; i < k; i += BB_THREADS) bulk_s2g(P.Dp + (int64_t)i * lds + c0, Ds + i * ncp, row_bytes);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
    } else if (p_vec) {
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
    // nobody leaves while a peer may still address its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (stamp) stamp[(int64_t)8 * k + 7] = stamp[(int64_t)8 * k + 6] = clock64();
#undef BB_STAMP
#undef BB_REL_STAMP
}

}  // namespace modl
