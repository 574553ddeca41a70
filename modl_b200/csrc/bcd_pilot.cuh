// bcd_pilot.cuh -- latency-optimised dictionary update for panels that fit the shared memory of
// ONE thread-block cluster (the common case: k * s * sizeof(T) <= ~2.5 MB, e.g. BASELINE
// config 2: 256 x 1250 floats).  Same mathematics and same fixed-order reductions as
// bcd_update_kernel (bcd_kernels.cuh) [ref: modl/decomposition/dict_fact.py:675-694]; what
// changes is how the k-step dependent chain is scheduled on the SM:
//
//   * warp specialisation.  Warp 0 of every CTA (the "pilot") owns the critical path of an atom
//     step -- candidate row on the CTA's columns, three partial sums, all-to-all exchange,
//     projection -- entirely with warp-level primitives: no __syncthreads on the chain.
//   * blocked look-ahead.  The row products C[a,:] . D_sub that every atom needs are produced by
//     the other 7 warps ("workers") for M = 8 atoms at a time as an 8 x k x cols register-tiled
//     product (D slice read from shared memory once per 8 atoms instead of once per atom, FFMA2),
//     one block AHEAD of the pilot and against the dictionary as it was one block ago; the pilot
//     repairs the staleness with rank-1 corrections  sum_j' C[a, a_j'] (d_new - d_old)_j'  over the
//     <= 8 + 8 atoms updated since (exact algebra, a handful of FMAs per column).
//   * the exchange is st.async + mbarrier over distributed shared memory (one-way latency).
//
// Roles synchronise once per block of 8 atoms with named barriers; new atom rows are staged and
// committed to the shared D slice at block boundaries so the workers always read a consistent
// snapshot.
//
//   * pipelined norm exchange (PIPE, L2 ball without positivity -- the benchmark configuration).  The only
//     thing atom t needs from the rest of the cluster is a scalar: |v_t|^2, the squared norm of its candidate
//     row, which fixes the projection scale alpha_t.  The candidate is LINEAR in the previous atom's scale,
//         v_t = u_t - g_t alpha_{t-1} v_{t-1},     g_t = C[a_t, a_{t-1}] / C[a_t, a_t],
//     with u_t free of alpha_{t-1} (it needs alpha_{t-2} and older only), so
//         |v_t|^2 = |u_t|^2 - 2 g_t alpha_{t-1} <u_t, v_{t-1}> + (g_t alpha_{t-1})^2 |v_{t-1}|^2 .
//     The cluster therefore exchanges the partial sums of (|u_t|^2, <u_t, v_{t-1}>) one atom EARLY: the all-to-all
//     of atom t+1 is sent before the one of atom t is waited for, two exchanges are in flight, and the per-atom
//     period drops from (local work + exchange latency) to max(local work, exchange latency).  Exact algebra; the
//     squared norm is assembled from three terms instead of summed directly (relative difference ~1e-7 in float).
#pragma once
#include "bcd_kernels.cuh"

namespace modl {

constexpr int BP_M = 8;                       // atoms per look-ahead block
// The per-atom chain is latency bound, and most of its length is the serial work of ONE warp over
// its 2-6 column groups.  So the chain is carried by BP_PW pilot warps, each owning every BP_PW-th
// 32-column group of the CTA's slice (one group per pilot at the benchmark shape); they combine
// their two partial sums through shared memory and a 96-thread named barrier, warp 0 sends, all wait.
constexpr int BP_PW = 3;                      // pilot warps
constexpr int BP_THREADS = 384;               // warps 0-2 = pilots, warps 3-11 = workers
constexpr int BP_WORKERS = BP_THREADS - 32 * BP_PW;
constexpr int BP_SYNCED = BP_THREADS;         // threads that meet at the pilot <-> worker barriers
enum { BP_BAR_PRODUCT = 1, BP_BAR_SNAPSHOT = 2, BP_BAR_WORKERS = 3, BP_BAR_PILOTS = 4 };

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }

// shared-memory footprint (elements of T, then bytes) -- must match the carve-up in the kernel
template <typename T>
__host__ __device__ inline size_t bcd_pilot_smem_bytes(int64_t k, int64_t ncp)
{
    const int64_t kp = round_up(k, 32);
    const int64_t np = ncp / 2;
    const int64_t igw = BP_WORKERS / np > 0 ? BP_WORKERS / np : 1;
    const int64_t elems = 2 * kp * BP_M          // Cblk
                          + 2 * BP_M * ncp       // Rbuf
                          + 2 * BP_M * ncp       // brows
                          + 4 * BP_M * ncp       // vnew, delta (two blocks each)
                          + 2 * kp               // cnorm, rad
                          + 4 * BCD_MAX_CLUSTER * BCD_NPART   // xch (BP_XRING)
                          + igw * BP_M * ncp     // red
                          + 2 * 224;             // per-block tables (two coefficient tables, diagonals, atom ids), x2
    return (size_t)elems * sizeof(T) + 40 * sizeof(double) + (size_t)k * ncp * sizeof(T);
}

// warp-level version of enet_threshold_block (basic_kernels.cuh): same monotone active-set fixed point.
template <typename T, typename Load>
__device__ T enet_threshold_warp(Load load, int n, T radius_over_l1, T gamma, bool *inside)
{
    const int lane = threadIdx.x & 31;
    double s1 = 0, s2 = 0, cnt = 0;
    for (int j = lane; j < n; j += 32) {
        const double a = fabs((double)load(j));
        s1 += a; s2 += a * a; cnt += 1;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2); cnt = warp_sum(cnt);
    const double g = (double)gamma, R = (double)radius_over_l1;
    const double norm = s1 + 0.5 * g * s2;
    if ((T)norm <= radius_over_l1) { *inside = true; return T(0); }
    *inside = false;
    double l = 0;
    for (int it = 0; it < 64; ++it) {
        const double sa = s1 + 0.5 * g * s2;
        double lnew;
        if (g != 0) {
            const double qa = g * g * R + 0.5 * g * cnt, qd = 2 * R * g + cnt, qc = R - sa;
            lnew = (-qd + sqrt(qd * qd - 4 * qa * qc)) / (2 * qa);
        } else {
            lnew = (sa - R) / cnt;
        }
        double n1 = 0, n2 = 0, nc = 0;
        for (int j = lane; j < n; j += 32) {
            const double a = fabs((double)load(j));
            if (a > lnew) { n1 += a; n2 += a * a; nc += 1; }
        }
        n1 = warp_sum(n1); n2 = warp_sum(n2); nc = warp_sum(nc);
        l = lnew;
        if (nc == cnt || nc == 0) break;
        s1 = n1; s2 = n2; cnt = nc;
    }
    return (T)l;
}

constexpr int BP_XRING = 4;                   // exchange slots / mbarriers in rotation (two exchanges in flight need four)

template <typename T, int NCL, bool ENET, bool PIPE = false>
__global__ void __launch_bounds__(BP_THREADS, 1)
bcd_pilot_kernel(BcdParams<T> P)
{
    static_assert(!(ENET && PIPE), "the pipelined exchange is for the L2 ball");
    extern __shared__ __align__(16) unsigned char bp_smem_raw[];
    const int k = P.k, s = P.s, lds = P.lds;
    const int nblk = gridDim.x, g = blockIdx.x;
    const int c0 = min(s, g * P.cols_per_cta);
    const int c1 = min(s, c0 + P.cols_per_cta);
    const int nc = c1 - c0;
    const int ncp = (int)round_up(P.cols_per_cta, 32);
    const int kp = (int)round_up(k, 32);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr bool enet = ENET;
    const int nbk = (k + BP_M - 1) / BP_M;
    const int NP = ncp >> 1;
    const int IGW = max(1, BP_WORKERS / NP);
    constexpr int TAB = 224;                                  // per-block table: cfm 64 | cfp 64 | rcv 8 | cav 8 | ordc 16 ints

    // ---- shared memory carve-up (see bcd_pilot_smem_bytes) ----
    T *Cblk = reinterpret_cast<T *>(bp_smem_raw);            // [2][kp][M] : C[A[j], i] at [i*M + j]
    T *Rbuf = Cblk + 2 * kp * BP_M;                           // [2][M][ncp] look-ahead products
    T *brows = Rbuf + 2 * BP_M * ncp;                         // [2][M][ncp] B_sub rows
    T *vnew = brows + 2 * BP_M * ncp;                         // [2][M][ncp] staged new rows, by block parity
    T *delta = vnew + 2 * BP_M * ncp;                         // [2][M][ncp] new - old, by block parity
    T *cnorm = delta + 2 * BP_M * ncp;                        // [kp] comp_norm_ on entry
    T *rad = cnorm + kp;                                      // [kp] radius used for every atom
    T *xch = rad + kp;                                        // [BP_XRING][16][4] exchange slots
    T *red = xch + BP_XRING * BCD_MAX_CLUSTER * BCD_NPART;    // [IGW][M][ncp]
    T *tabs = red + (size_t)IGW * BP_M * ncp;                 // [2][TAB]
    double *dscratch = reinterpret_cast<double *>(tabs + 2 * TAB);
    T *Ds = reinterpret_cast<T *>(dscratch + 40);             // [k][ncp]
    __shared__ __align__(8) unsigned long long xbar[BP_XRING];
    const unsigned xbar_addr = (unsigned)__cvta_generic_to_shared(xbar);
    const unsigned xch_addr = (unsigned)__cvta_generic_to_shared(xch);
    constexpr unsigned kSlotBytes = BCD_NPART * sizeof(T);

    long long *xstamp = (P.timing && g == 0 && tid == 0) ? P.timing + (int64_t)8 * k : nullptr;   // debug
    if (xstamp) xstamp[0] = clock64();
    const T *Dg = P.Dp + c0;
    constexpr int VE = 16 / (int)sizeof(T);                   // elements per 128-bit access
    const bool vec_ok = (lds % VE == 0) && (c0 % VE == 0) && ((reinterpret_cast<uintptr_t>(P.Dp) & 15) == 0);
    if (vec_ok) {
        const int nv = ncp / VE;
#pragma unroll 4
        for (int e = tid; e < k * nv; e += BP_THREADS) {
            const int i = e / nv, cv = (e % nv) * VE;
            alignas(16) T tmp[VE];
#pragma unroll
            for (int u = 0; u < VE; ++u) tmp[u] = T(0);
            if (cv < nc) *reinterpret_cast<uint4 *>(tmp) = *reinterpret_cast<const uint4 *>(Dg + (int64_t)i * lds + cv);
#pragma unroll
            for (int u = 0; u < VE; ++u) Ds[i * ncp + cv + u] = (cv + u < nc) ? tmp[u] : T(0);
        }
    } else {
        for (int e = tid; e < k * ncp; e += BP_THREADS) {
            const int i = e / ncp, c = e % ncp;
            Ds[e] = (c < nc) ? Dg[(int64_t)i * lds + c] : T(0);
        }
    }
    for (int i = tid; i < k; i += BP_THREADS) cnorm[i] = P.comp_norm[i];
    for (int e = tid; e < 4 * BP_M * ncp; e += BP_THREADS) vnew[e] = T(0);       // vnew and delta (contiguous)
    for (int e = tid; e < BP_XRING * BCD_MAX_CLUSTER * BCD_NPART; e += BP_THREADS) xch[e] = T(0);
    if (tid == 0) {
        for (int r = 0; r < BP_XRING; ++r) mbar_init(xbar_addr + 8 * r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // every peer is resident and its mbarriers initialised before anybody sends
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (g == 0 && tid == 0) bcd_signal_start(P);
    if (xstamp) xstamp[1] = clock64();

    if (wid >= BP_PW) {
        // =========================== workers: everything off the critical path ===========================
        const int wt = tid - 32 * BP_PW;
        const int pr = wt % NP, ig = wt / NP;
        const int RB = (((k + IGW - 1) / IGW) + 3) & ~3;
        for (int b = 0; b < nbk; ++b) {
            if (b > 0) named_sync(BP_BAR_SNAPSHOT, BP_SYNCED);     // pilot is done with block b-2 and its buffers
            const int nb = b & 1;
            const int mb = min(BP_M, k - b * BP_M);
            T *Cb = Cblk + nb * kp * BP_M;
            T *Rb = Rbuf + nb * BP_M * ncp;
            T *Bb = brows + nb * BP_M * ncp;
            T *tab = tabs + nb * TAB;
            // coefficients (transposed) and B_sub rows of the block's atoms: all copies in flight at once
            for (int e = wt; e < k * BP_M; e += BP_WORKERS) {
                const int j = e / k, i = e % k;
                if (j < mb) cp_async_elem(Cb + i * BP_M + j, P.C + (int64_t)P.order[b * BP_M + j] * k + i);
                else Cb[i * BP_M + j] = T(0);
            }
            for (int e = wt; e < BP_M * ncp; e += BP_WORKERS) {
                const int j = e / ncp, c = e % ncp;
                if (j < mb && c < nc) cp_async_elem(Bb + e, P.Bp + (int64_t)P.order[b * BP_M + j] * lds + c0 + c);
                else Bb[e] = T(0);
            }
            cp_async_commit();
            // commit the rows staged two blocks ago: the shared slice becomes the dictionary at the start of block b-1
            if (b >= 2) {
                const T *vs = vnew + nb * BP_M * ncp;
                for (int e = wt; e < BP_M * ncp; e += BP_WORKERS) {
                    const int jp = e / ncp, c = e % ncp;
                    Ds[P.order[(b - 2) * BP_M + jp] * ncp + c] = vs[e];
                }
            }
            cp_async_wait_all();
            named_sync(BP_BAR_WORKERS, BP_WORKERS);
            // small tables for the pilot: repair coefficients, diagonals, atom ids
            if (wt < BP_M * BP_M) {
                const int j = wt / BP_M, jp = wt % BP_M;
                tab[wt] = (jp < j && j < mb) ? Cb[P.order[b * BP_M + jp] * BP_M + j] : T(0);                   // cfm
                tab[64 + wt] = (b > 0 && j < mb) ? Cb[P.order[(b - 1) * BP_M + jp] * BP_M + j] : T(0);         // cfp
            } else if (wt < BP_M * BP_M + BP_M) {
                const int j = wt - BP_M * BP_M;
                const T d = (j < mb) ? Cb[P.order[b * BP_M + j] * BP_M + j] : T(1);
                tab[136 + j] = d;                                                                             // cav
                tab[128 + j] = T(1) / d;                                                                      // rcv
            } else if (wt < BP_M * BP_M + 2 * BP_M) {
                const int j = wt - BP_M * BP_M - BP_M;
                reinterpret_cast<int *>(tab + 144)[j] = (j < mb) ? P.order[b * BP_M + j] : 0;                 // ordc
            }
            if (ig < IGW) {
                const int r0 = ig * RB, r1 = min(k, r0 + RB);
                const Pair<T> *dcol = reinterpret_cast<const Pair<T> *>(Ds) + pr;
                const int rs = ncp >> 1;
                Pair<T> acc[BP_M];
#pragma unroll
                for (int j = 0; j < BP_M; ++j) acc[j].x = acc[j].y = T(0);
#pragma unroll 2
                for (int i = r0; i < r1; ++i) {
                    const Pair<T> d = dcol[i * rs];
                    const Quad<T> q0 = *reinterpret_cast<const Quad<T> *>(Cb + i * BP_M);
                    const Quad<T> q1 = *reinterpret_cast<const Quad<T> *>(Cb + i * BP_M + 4);
                    pair_fma(q0.x, d, acc[0]); pair_fma(q0.y, d, acc[1]); pair_fma(q0.z, d, acc[2]); pair_fma(q0.w, d, acc[3]);
                    pair_fma(q1.x, d, acc[4]); pair_fma(q1.y, d, acc[5]); pair_fma(q1.z, d, acc[6]); pair_fma(q1.w, d, acc[7]);
                }
#pragma unroll
                for (int j = 0; j < BP_M; ++j) {
                    red[((size_t)ig * BP_M + j) * ncp + 2 * pr] = acc[j].x;
                    red[((size_t)ig * BP_M + j) * ncp + 2 * pr + 1] = acc[j].y;
                }
            }
            named_sync(BP_BAR_WORKERS, BP_WORKERS);
            for (int e = wt; e < BP_M * ncp; e += BP_WORKERS) {
                T sum = T(0);
                for (int gi = 0; gi < IGW; ++gi) sum += red[(size_t)gi * BP_M * ncp + e];   // fixed order
                Rb[e] = sum;
            }
            __threadfence_block();
            named_arrive(BP_BAR_PRODUCT, BP_SYNCED);
        }
    } else {
        if constexpr (PIPE) {
        // =========================== pilots, pipelined norm exchange ===========================
        constexpr int NCLW = (NCL + BP_PW - 1) / BP_PW;    // column groups per pilot warp
        __shared__ double ppsum_raw[BP_XRING * BP_PW * 3];
        T *ppsum = reinterpret_cast<T *>(ppsum_raw);       // [ring][BP_PW][3] per-warp partial sums
        unsigned rslot[BP_XRING], rbar[BP_XRING];
#pragma unroll
        for (unsigned r = 0; r < BP_XRING; ++r) {
            const unsigned peer = (unsigned)(lane < nblk ? lane : 0);
            rslot[r] = mapa_u32(xch_addr + (r * BCD_MAX_CLUSTER + (unsigned)g) * kSlotBytes, peer);
            rbar[r] = mapa_u32(xbar_addr + 8 * r, peer);
        }
        // exchange number t uses ring entry t & 3, phase (t >> 2) & 1 of its mbarrier.  Entry t+1 may be written into a
        // peer while that peer has not consumed entry t yet; entry t+1 replaces entry t-3, which the peer read before it
        // sent its entry t-1 -- and that one has been received here.
        auto xsend = [&](int t, T p0, T p1, T p2) {
            const unsigned r = (unsigned)t & (BP_XRING - 1);
            if (tid == 0) mbar_expect_tx(xbar_addr + 8 * r, (unsigned)nblk * kSlotBytes);
            p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2);
            T *ps = ppsum + r * (3 * BP_PW);
            if (lane == 0) { ps[3 * wid] = p0; ps[3 * wid + 1] = p1; ps[3 * wid + 2] = p2; }
            named_sync(BP_BAR_PILOTS, 32 * BP_PW);
            if (wid == 0) {
                T t0 = ps[0], t1 = ps[1], t2 = ps[2];
#pragma unroll
                for (int w = 1; w < BP_PW; ++w) { t0 += ps[3 * w]; t1 += ps[3 * w + 1]; t2 += ps[3 * w + 2]; }   // fixed order
                if (lane < nblk) st_async_triplet(rslot[r], rbar[r], t0, t1, t2);
            }
        };
        auto xwait = [&](int t, T &o0, T &o1, T &o2) {
            const unsigned r = (unsigned)t & (BP_XRING - 1);
            mbar_wait(xbar_addr + 8 * r, ((unsigned)t >> 2) & 1u);
            const T *src = xch + (size_t)r * BCD_MAX_CLUSTER * BCD_NPART;
            T a0[4], a1[4], a2[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                a0[c4] = src[(4 * c4) * BCD_NPART];
                a1[c4] = src[(4 * c4) * BCD_NPART + 1];
                a2[c4] = src[(4 * c4) * BCD_NPART + 2];
#pragma unroll
                for (int u = 1; u < 4; ++u) {
                    a0[c4] += src[(4 * c4 + u) * BCD_NPART];
                    a1[c4] += src[(4 * c4 + u) * BCD_NPART + 1];
                    a2[c4] += src[(4 * c4 + u) * BCD_NPART + 2];
                }
            }
            o0 = (a0[0] + a0[1]) + (a0[2] + a0[3]);
            o1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
            o2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
        };
        // per-block tables of atom t
        auto tab_of = [&](int t) { return tabs + ((t / BP_M) & 1) * TAB; };
        // coupling of atom t with the atom updated just before it (c1) and two before it (c2): C[a_t, a_{t-1}], C[a_t, a_{t-2}]
        auto coef1 = [&](int t) -> T {
            const int j = t % BP_M;
            const T *tb = tab_of(t);
            return j >= 1 ? tb[j * BP_M + j - 1] : (t >= BP_M ? tb[64 + BP_M - 1] : T(0));
        };
        auto coef2 = [&](int t) -> T {
            const int j = t % BP_M;
            const T *tb = tab_of(t);
            return j >= 2 ? tb[j * BP_M + j - 2] : (t >= BP_M ? tb[64 + j * BP_M + BP_M - 2 + j] : T(0));
        };
        // pre_t = B_sub[a_t] - (look-ahead product + repairs for every atom up to t-3), the old row of atom t and its
        // enet_norm partial: everything of atom t that is known once atom t-3 is final.
        auto precompute = [&](int t, T (&pre)[NCLW], T (&dold)[NCLW], T &nb_l) {
            const int b = t / BP_M, jn = t % BP_M, cur = b & 1;
            if (jn == 0) named_sync(BP_BAR_PRODUCT, BP_SYNCED);      // look-ahead product, B rows and tables of block b are ready
            const T *Rb = Rbuf + cur * BP_M * ncp;
            const T *Bb = brows + cur * BP_M * ncp;
            const T *tb = tabs + cur * TAB;
            const T *cfm = tb, *cfp = tb + 64;
            const int a = reinterpret_cast<const int *>(tb + 144)[jn];
            const T *dcur = delta + cur * BP_M * ncp;
            const T *dprev = delta + (cur ^ 1) * BP_M * ncp;
            const int nprev = BP_M - (jn < 2 ? 2 - jn : 0);          // atoms of the previous block that are final (t-3 and older)
            nb_l = T(0);
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) {
                const int m = wid + mm * BP_PW;
                pre[mm] = dold[mm] = T(0);
                if (m >= NCL) continue;
                const int c = lane + 32 * m;
                T dot = Rb[jn * ncp + c], dot2 = T(0);
#pragma unroll
                for (int jp = 0; jp < BP_M; ++jp)
                    if (jp < nprev) dot = fma(cfp[jn * BP_M + jp], dprev[jp * ncp + c], dot);
#pragma unroll
                for (int jp = 0; jp < BP_M - 3; ++jp)
                    if (jp < jn - 2) dot2 = fma(cfm[jn * BP_M + jp], dcur[jp * ncp + c], dot2);
                dold[mm] = Ds[a * ncp + c];
                pre[mm] = Bb[jn * ncp + c] - (dot + dot2);
                nb_l += enet_term(dold[mm], P.l1_ratio);
            }
        };
        // u_t from its base row:  u = (base + C[a_t, a_{t-1}] d_old[t-1] + C[a_t, a_t] d_old[t]) / C[a_t, a_t]   [ref: :676-685]
        auto make_u = [&](int t, const T (&base)[NCLW], const T (&dold_prev)[NCLW], const T (&dold)[NCLW], T (&u)[NCLW], T &g_out) {
            const T *tb = tab_of(t);
            const int j = t % BP_M;
            const T caa = tb[136 + j], rcaa = tb[128 + j];
            const bool upd = caa > T(1e-20);                         // [ref: :681-683]
            const T c1 = coef1(t);
            T gq = c1 * rcaa;
            gq = fma(fma(-gq, caa, c1), rcaa, gq);                   // c1 / caa, Newton-corrected
            g_out = upd ? gq : T(0);
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) {
                const T grad = fma(c1, dold_prev[mm], base[mm]) + caa * dold[mm];
                T q = grad * rcaa;
                q = fma(fma(-q, caa, grad), rcaa, q);
                u[mm] = upd ? q : dold[mm];
            }
        };
        // projection scale of atom t from the exchanged sums  [ref: enet.pyx:62-70]
        auto scale_of = [&](int t, T nb, T sv2, T &rnrm, T &nrm) {
            const int a = reinterpret_cast<const int *>(tab_of(t) + 144)[t % BP_M];
            const T radius = cnorm[a] + nb;                          // comp_norm_[k] += subset_norm  [ref: :676-678]
            if (tid == 0) rad[a] = radius;
            rnrm = T(1); nrm = T(1);
            if (radius == T(0)) {
                rnrm = T(0);
            } else {
                const T x = sv2 / radius;
                if (x > T(1)) { rnrm = bcd_rsqrt(x); nrm = x * rnrm; }
            }
        };
        // final row of atom t: v / nrm, staged for the workers, and its change
        auto finalize = [&](int t, const T (&v)[NCLW], const T (&dold)[NCLW], T rnrm, T nrm, T (&dnew)[NCLW], T (&dlt)[NCLW]) {
            const int cur = (t / BP_M) & 1, j = t % BP_M;
            T *vcur = vnew + cur * BP_M * ncp;
            T *dcur = delta + cur * BP_M * ncp;
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) {
                const int m = wid + mm * BP_PW;
                dnew[mm] = dlt[mm] = T(0);
                if (m >= NCL) continue;
                const int c = lane + 32 * m;
                const T tt = v[mm] * rnrm;
                const T q = fma(fma(-tt, nrm, v[mm]), rnrm, tt);     // v / nrm [ref: enet.pyx:69-70]; 0 when radius == 0
                dnew[mm] = q;
                dlt[mm] = q - dold[mm];
                vcur[j * ncp + c] = q;
                dcur[j * ncp + c] = dlt[mm];
            }
        };

        T vp[NCLW], dop[NCLW], uc[NCLW], doc[NCLW], pn[NCLW], don[NCLW], nbn = T(0);
        T gc = T(0), sv2p = T(0), rnrm = T(1), nrm = T(1);
        {   // ---- prologue: atoms 0 and 1 enter the pipeline ----
            T pre0[NCLW], nb0, g0;
            T zero[NCLW];
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) zero[mm] = T(0);
            precompute(0, pre0, dop, nb0);
            if (nbk > 1) named_arrive(BP_BAR_SNAPSHOT, BP_SYNCED);   // workers may start on block 1
            make_u(0, pre0, zero, dop, vp, g0);                      // no predecessor: v_0 = u_0
            T s11 = T(0);
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) s11 = fma(vp[mm], vp[mm], s11);
            xsend(0, nb0, s11, T(0));
            if (k > 1) {
                T pre1[NCLW], nb1;
                precompute(1, pre1, doc, nb1);                       // nothing final yet: the bare product
                make_u(1, pre1, dop, doc, uc, gc);
                T s1 = T(0), s2 = T(0);
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) { s1 = fma(uc[mm], uc[mm], s1); s2 = fma(uc[mm], vp[mm], s2); }
                xsend(1, nb1, s1, s2);
            }
            if (k > 2) precompute(2, pn, don, nbn);
            T nb, u2, uv;
            xwait(0, nb, u2, uv);
            sv2p = u2;
            scale_of(0, nb, sv2p, rnrm, nrm);
        }
        for (int t = 1; t < k; ++t) {
            if (P.timing && g == 0 && tid == 0) P.timing[(int64_t)t * 8] = clock64();
            // A: atom t-1 is final
            T dnew[NCLW], dlt[NCLW];
            finalize(t - 1, vp, dop, rnrm, nrm, dnew, dlt);
            if (t % BP_M == 0 && t / BP_M + 1 < nbk) named_arrive(BP_BAR_SNAPSHOT, BP_SYNCED);   // block t/8 - 1 is complete: its buffers may be recycled
            // B: candidate row of atom t
            T vc[NCLW];
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) vc[mm] = fma(-gc, dnew[mm], uc[mm]);
            // C, D: u of atom t+1 and its sums leave one exchange early
            T un[NCLW], gn = T(0);
            if (t + 1 < k) {
                const T c2 = coef2(t + 1);
                T base[NCLW];
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) base[mm] = fma(-c2, dlt[mm], pn[mm]);
                make_u(t + 1, base, doc, don, un, gn);
                T s1 = T(0), s2 = T(0);
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) { s1 = fma(un[mm], un[mm], s1); s2 = fma(un[mm], vc[mm], s2); }
                xsend(t + 1, nbn, s1, s2);
            }
            // between the send and the wait: everything of atom t+2 that is already determined
            T pn2[NCLW], don2[NCLW], nbn2 = T(0);
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) pn2[mm] = don2[mm] = T(0);
            if (t + 2 < k) precompute(t + 2, pn2, don2, nbn2);
            // E, F: |v_t|^2 from the sums sent one atom ago, then the projection scale of atom t
            T nb, u2, uv;
            xwait(t, nb, u2, uv);
            const T h = gc * rnrm;                                   // g_t alpha_{t-1}
            T sv2 = fma(h, fma(h, sv2p, T(-2) * uv), u2);
            sv2 = sv2 > T(0) ? sv2 : T(0);
            scale_of(t, nb, sv2, rnrm, nrm);
            sv2p = sv2;
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) {
                vp[mm] = vc[mm]; dop[mm] = doc[mm]; uc[mm] = un[mm]; doc[mm] = don[mm]; pn[mm] = pn2[mm]; don[mm] = don2[mm];
            }
            gc = gn;
            nbn = nbn2;
        }
        {   // the last atom
            T dnew[NCLW], dlt[NCLW];
            finalize(k - 1, vp, dop, rnrm, nrm, dnew, dlt);
        }
        } else {
        // =========================== pilots: the dependent chain ===========================
        // Pilot warp w owns the 32-column groups m = w, w + BP_PW, ... (< NCL) of the CTA's slice, lane l
        // the column l + 32 m; padded columns carry zeros through every formula.
        constexpr int NCLW = (NCL + BP_PW - 1) / BP_PW;    // column groups per pilot warp
        __shared__ double psum_raw[2 * BP_PW * 2];
        T *psum = reinterpret_cast<T *>(psum_raw);         // [2 parities][BP_PW][2] per-warp partial sums
        unsigned xi = 0;                                   // exchange counter
        long long *tstamp = nullptr;
#define BP_STAMP(slot) do { if (tstamp && tid == 0) tstamp[(slot)] = clock64(); } while (0)
        // peer addresses of my slot and of the peer's mbarrier, per parity
        unsigned rslot[2], rbar[2];
#pragma unroll
        for (unsigned par = 0; par < 2; ++par) {
            const unsigned peer = (unsigned)(lane < nblk ? lane : 0);
            rslot[par] = mapa_u32(xch_addr + (par * BCD_MAX_CLUSTER + (unsigned)g) * kSlotBytes, peer);
            rbar[par] = mapa_u32(xbar_addr + 8 * par, peer);
        }
        // The exchange in two halves, so that independent work can sit between the send and the wait.  The
        // pilots combine their partial sums through shared memory (96-thread named barrier) and warp 0 sends
        // ONE slot per peer.  (Measured alternative: every pilot warp sending its own slot removes the barrier
        // but triples the DSMEM traffic and the slots to add: 1 460 vs 1 240 cycles per atom.)
        auto exchange_send = [&](T p0, T p1) {
            const unsigned par = xi & 1u;
            if (tid == 0) mbar_expect_tx(xbar_addr + 8 * par, (unsigned)nblk * kSlotBytes);   // armed early, off the chain
            p0 = warp_sum(p0); p1 = warp_sum(p1);
            T *ps = psum + par * (2 * BP_PW);
            if (lane == 0) { ps[2 * wid] = p0; ps[2 * wid + 1] = p1; }
            if (enet) __threadfence();                     // my slice of the candidate row is visible before the send
            named_sync(BP_BAR_PILOTS, 32 * BP_PW);
            BP_STAMP(2);
            if (wid == 0) {
                T t0 = ps[0], t1 = ps[1];
#pragma unroll
                for (int w = 1; w < BP_PW; ++w) { t0 += ps[2 * w]; t1 += ps[2 * w + 1]; }   // fixed order
                if (enet) __threadfence();
                if (lane < nblk) st_async_triplet(rslot[par], rbar[par], t0, t1, T(0));
            }
            BP_STAMP(3);
        };
        auto exchange_wait = [&](T &o0, T &o1) {
            const unsigned par = xi & 1u;
            mbar_wait(xbar_addr + 8 * par, (xi >> 1) & 1u);
            BP_STAMP(4);
            if (enet) __threadfence();
            // slots of absent peers stay zero.  Every lane reads all 16 slots (broadcast loads) and adds them in
            // the same fixed tree (four chains of four, then a pair of pairs): bit-identical in every CTA, and
            // shorter than a 4-round shuffle butterfly on the dependent chain
            const T *src = xch + (size_t)par * BCD_MAX_CLUSTER * BCD_NPART;
            T a0[4], a1[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                a0[c4] = src[(4 * c4) * BCD_NPART];
                a1[c4] = src[(4 * c4) * BCD_NPART + 1];
#pragma unroll
                for (int u = 1; u < 4; ++u) {
                    a0[c4] += src[(4 * c4 + u) * BCD_NPART];
                    a1[c4] += src[(4 * c4 + u) * BCD_NPART + 1];
                }
            }
            o0 = (a0[0] + a0[1]) + (a0[2] + a0[3]);
            o1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
            xi += 1;
        };

        for (int b = 0; b < nbk; ++b) {
            const int cur = b & 1;
            const int mb = min(BP_M, k - b * BP_M);
            const T *Rb = Rbuf + cur * BP_M * ncp;
            const T *Bb = brows + cur * BP_M * ncp;
            const T *tab = tabs + cur * TAB;
            const T *cfm = tab, *cfp = tab + 64, *rcv = tab + 128, *cav = tab + 136;
            const int *ordc = reinterpret_cast<const int *>(tab + 144);
            T *vcur = vnew + cur * BP_M * ncp;
            T *dcur = delta + cur * BP_M * ncp;
            const T *dprev = delta + (cur ^ 1) * BP_M * ncp;
            if (xstamp) xstamp[16 + 2 * b] = clock64();
            named_sync(BP_BAR_PRODUCT, BP_SYNCED);               // look-ahead product, B rows and tables of block b are ready
            if (xstamp) xstamp[16 + 2 * b + 1] = clock64();
            if (b + 1 < nbk) named_arrive(BP_BAR_SNAPSHOT, BP_SYNCED);   // workers may recycle the buffers of block b-1

            // Everything atom j needs EXCEPT the contribution of the atom updated just before it:
            //   base = B_sub[a_j] - (look-ahead product + repairs for the previous block and for atoms < j-1 of this one)
            // It does not depend on the projection in flight, so it is computed for atom j+1 between the send and
            // the wait of atom j's exchange: most of the candidate's work leaves the dependent chain.
            auto precompute = [&](int jn, T (&base)[NCLW], T (&dold)[NCLW], T &nb_l) {
                const int a = ordc[jn];
                nb_l = T(0);
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) {
                    const int m = wid + mm * BP_PW;
                    base[mm] = dold[mm] = T(0);
                    if (m >= NCL) continue;
                    const int c = lane + 32 * m;
                    T dot = Rb[jn * ncp + c], dot2 = T(0);
#pragma unroll
                    for (int jp = 0; jp < BP_M; ++jp) dot = fma(cfp[jn * BP_M + jp], dprev[jp * ncp + c], dot);
#pragma unroll
                    for (int jp = 0; jp < BP_M - 2; ++jp)
                        if (jp < jn - 1) dot2 = fma(cfm[jn * BP_M + jp], dcur[jp * ncp + c], dot2);
                    dold[mm] = Ds[a * ncp + c];
                    base[mm] = Bb[jn * ncp + c] - (dot + dot2);
                    nb_l += enet_term(dold[mm], P.l1_ratio);
                }
            };

            T base[NCLW], dold[NCLW], dlast[NCLW], nb_l;
#pragma unroll
            for (int mm = 0; mm < NCLW; ++mm) dlast[mm] = T(0);
            precompute(0, base, dold, nb_l);
            for (int j = 0; j < mb; ++j) {
                tstamp = (P.timing && g == 0) ? P.timing + (int64_t)(b * BP_M + j) * 8 : nullptr;
                BP_STAMP(0);
                const int a = ordc[j];
                const T caa = cav[j];
                const bool upd = caa > T(1e-20);                    // [ref: :681-683]
                const T rcaa = rcv[j];
                const unsigned par = xi & 1u;
                const T clast = j > 0 ? cfm[j * BP_M + j - 1] : T(0);   // C[a_j, a_{j-1}]
                const T cn_a = cnorm[a];                             // read before the wait: off the dependent chain
                // ---- candidate row on my columns [ref: :676-685] ----
                T sv2_l = T(0);
                T v[NCLW];
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) {
                    const int m = wid + mm * BP_PW;
                    v[mm] = T(0);
                    if (m >= NCL) continue;
                    const int c = lane + 32 * m;
                    const T grad = fma(-clast, dlast[mm], base[mm]) + caa * dold[mm];
                    T q = grad * rcaa;
                    q = fma(fma(-q, caa, grad), rcaa, q);           // grad / caa, Newton-corrected
                    q = upd ? q : dold[mm];
                    if (P.positive && q < T(0)) q = T(0);           // [ref: :684-685]
                    v[mm] = q;
                    sv2_l = fma(q, q, sv2_l);
                    if (enet && c < nc) P.vrow[(int64_t)par * s + c0 + c] = q;
                }
                BP_STAMP(1);
                exchange_send(nb_l, sv2_l);
                // between the send and the wait: the next atom's base row (uses atoms < j of this block only)
                T nbase[NCLW], ndold[NCLW], nnb_l = T(0);
                if (j + 1 < mb) precompute(j + 1, nbase, ndold, nnb_l);
                T nb, sv2;
                exchange_wait(nb, sv2);
                BP_STAMP(5);
                const T radius = cn_a + nb;                          // comp_norm_[k] += subset_norm  [ref: :676-678]
                if (tid == 0) rad[a] = radius;
                // projection of the candidate on the ball of "radius" [ref: enet.pyx:38-122].  The L2 case is a
                // pure rescale v / nrm (Newton-corrected reciprocal multiply); radius == 0 maps to a zero scale.
                T lthr = T(0), gamma = T(0), nrm = T(1), rnrm = T(1);
                bool shrink = false;
                if (!enet) {
                    // nrm = sqrt(|v|^2 / radius) when the candidate leaves the ball [ref: enet.pyx:62-70]: one division,
                    // then 1 / nrm from the hardware reciprocal square root + one Newton step, nrm = x / nrm
                    if (radius == T(0)) {
                        rnrm = T(0);
                    } else {
                        const T x = sv2 / radius;
                        if (x > T(1)) { rnrm = bcd_rsqrt(x); nrm = x * rnrm; }
                    }
                } else if (radius == T(0)) {
                    rnrm = T(0);
                } else {
                    gamma = T(2) / P.l1_ratio - T(2);
                    const T *vr = P.vrow + (int64_t)par * s;
                    bool ins;
                    const T l = enet_threshold_warp<T>([&](int jj) { return __ldcg(vr + jj); }, s, radius / P.l1_ratio, gamma, &ins);
                    shrink = !ins;
                    lthr = l;
                }
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) {
                    const int m = wid + mm * BP_PW;
                    if (m >= NCL) continue;
                    const int c = lane + 32 * m;
                    T q = v[mm];
                    if (enet && shrink) {
                        q = enet_shrink(q, lthr, gamma);
                    } else {
                        const T t = q * rnrm;
                        q = fma(fma(-t, nrm, q), rnrm, t);          // v / nrm [ref: enet.pyx:69-70]; 0 when radius == 0
                    }
                    vcur[j * ncp + c] = q;
                    dlast[mm] = q - dold[mm];
                    dcur[j * ncp + c] = dlast[mm];
                }
#pragma unroll
                for (int mm = 0; mm < NCLW; ++mm) { base[mm] = nbase[mm]; dold[mm] = ndold[mm]; }
                nb_l = nnb_l;
                BP_STAMP(6);
                BP_STAMP(7);
            }
            tstamp = nullptr;
        }
#undef BP_STAMP
            }   // !PIPE
    }
    // ---- epilogue: commit the last two staged blocks, norms of the new atoms, write-back ----
    if (xstamp) xstamp[2] = clock64();
    __syncthreads();
    if (xstamp) xstamp[3] = clock64();
    for (int bb = max(0, nbk - 2); bb < nbk; ++bb) {
        const int mb = min(BP_M, k - bb * BP_M);
        const T *vs = vnew + (bb & 1) * BP_M * ncp;
        for (int e = tid; e < mb * ncp; e += BP_THREADS) {
            const int jp = e / ncp, c = e % ncp;
            Ds[P.order[bb * BP_M + jp] * ncp + c] = vs[e];
        }
    }
    __syncthreads();
    // comp_norm_[a] = radius_a - enet_norm(new atom) [ref: :690-692]: per-CTA partial norms -> global -> CTA 0
    T *napart = P.part;                                        // [nblk][k]
    for (int i = wid; i < k; i += BP_THREADS / 32) {
        T acc = T(0);
        for (int c = lane; c < nc; c += 32) acc += enet_term(Ds[i * ncp + c], P.l1_ratio);
        acc = warp_sum(acc);
        if (lane == 0) napart[(int64_t)g * k + i] = acc;
    }
    if (vec_ok) {
        const int nv = ncp / VE;
#pragma unroll 4
        for (int e = tid; e < k * nv; e += BP_THREADS) {
            const int i = e / nv, cv = (e % nv) * VE;
            if (cv + VE <= nc) {
                *reinterpret_cast<uint4 *>(P.Dp + (int64_t)i * lds + c0 + cv) = *reinterpret_cast<const uint4 *>(Ds + i * ncp + cv);
            } else {
                for (int u = 0; u < VE; ++u)
                    if (cv + u < nc) P.Dp[(int64_t)i * lds + c0 + cv + u] = Ds[i * ncp + cv + u];
            }
        }
    } else {
        for (int e = tid; e < k * ncp; e += BP_THREADS) {
            const int i = e / ncp, c = e % ncp;
            if (c < nc) P.Dp[(int64_t)i * lds + c0 + c] = Ds[e];
        }
    }
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (g == 0) {
        for (int i = tid; i < k; i += BP_THREADS) {
            T part[BCD_MAX_CLUSTER];
#pragma unroll
            for (int q = 0; q < BCD_MAX_CLUSTER; ++q) part[q] = q < nblk ? __ldcg(napart + (int64_t)q * k + i) : T(0);   // all in flight
            T na = T(0);
#pragma unroll
            for (int q = 0; q < BCD_MAX_CLUSTER; ++q) na += part[q];                     // fixed order
            P.comp_norm[i] = rad[i] - na;
        }
    }
    if (xstamp) xstamp[4] = clock64();
}

}  // namespace modl
