// bcd_kernels.cuh -- the sequential-over-atoms block-coordinate dictionary update on the
// gathered subset panels, replacing the Python loop of DictFact._update_dict
// [ref: modl/decomposition/dict_fact.py:675-694] together with enet_norm / enet_projection
// [ref: modl/utils/math/enet.pyx:38-148] that it calls per atom.
//
// Formulation.  The reference keeps a full gradient panel up to date with two rank-1 BLAS
// `ger` calls per atom (2 x k x s flops each).  Only row `a` of that panel is consumed when
// atom `a` is visited, and the panel invariant is  grad = B_sub - C . D_sub(current), so we
// evaluate that one row on demand:
//       g = (B_sub[a,:] - C[a,:] . D_sub) + C[a,a] D_sub[a,:]          (== grad[a] after ger +1)
// which halves the flops and needs no gradient panel.
//
// Parallelisation.  The s subset columns are split over the CTAs of ONE launch; every CTA
// keeps its column slice of D_sub in shared memory (when it fits) and walks the atoms in the
// same order.  The only cross-CTA dependency per atom is the projection radius test, which
// needs sums over all s columns: each CTA publishes three partial sums (and, for the
// elastic-net ball, its slice of the candidate row), all CTAs meet at ONE barrier per atom,
// then every CTA reduces the partials in the same fixed order (bitwise identical decision
// everywhere, no atomics) and finishes its own columns.
//
// Latency engineering (the kernel is a chain of k dependent atom steps):
//   * single thread-block cluster (<= 16 CTAs): partial sums are PUSHED into every peer's
//     shared memory (DSMEM stores) and the hardware cluster barrier is split into
//     arrive.release ... wait.acquire;
//   * between arrive and wait every CTA already computes the next atom's row product
//     C[a',:] . D_sub over all rows except the one being updated, so that after the barrier
//     only a rank-1 fix-up remains on the critical path;
//   * the C rows (two atoms ahead) and B_sub rows (one atom ahead) are prefetched with
//     cp.async into shared-memory rings; comp_norm_ is staged in shared memory once;
//   * enet_norm of the new atom is reduced together with the next atom's partial sums.
// Larger panels fall back to a cooperative launch over all SMs with the same code and a
// global-memory exchange + sense barrier.
#pragma once
#include <cooperative_groups.h>

#include "basic_kernels.cuh"
#include "common.cuh"

namespace modl {

namespace cg = cooperative_groups;

constexpr int BCD_THREADS = 256;
constexpr int BCD_NPART = 4;     // partial sums exchanged per atom and CTA (3 used)
constexpr int BCD_MAX_CLUSTER = 16;

// elements of the cross-row-block reduction scratch for a slice of `ncp` (padded) columns
__host__ __device__ inline int64_t bcd_red_elems(int64_t ncp)
{
    const int64_t np = ncp / 2 > 0 ? ncp / 2 : 1;
    const int64_t ig = BCD_THREADS / np > 0 ? BCD_THREADS / np : 1;
    const int64_t a = ig * ncp;
    return round_up(a > BCD_THREADS ? a : BCD_THREADS, 32);
}

template <typename T>
struct BcdParams {
    T *Dp;               // k x s panel of components_[:, subset]      (in/out), ld = lds
    const T *Bp;         // k x s panel of B_[:, subset]                          ld = lds
    const T *C;          // k x k
    T *comp_norm;        // k, in/out
    const int32_t *order;// k
    int k, s, lds;
    T l1_ratio;
    int positive;
    int cols_per_cta;    // columns owned by each CTA (last one may own fewer / none)
    int chunk;           // column-lane width (multiple of 32, <= BCD_THREADS)
    int d_in_smem;       // keep the CTA's D slice in shared memory
    int use_cluster;     // 1: DSMEM exchange + hardware cluster barrier, 0: global memory + sense barrier
    unsigned *bar;       // global barrier counter (zero on entry)
    unsigned *flags;     // [2][nblk] per-CTA epoch flags (zero on entry); used instead of `bar` when flag_barrier != 0
    int flag_barrier;
    T *part;             // [2][nblk][BCD_NPART]       (global exchange, use_cluster == 0)
    T *vrow;             // [2][s]   candidate rows (elastic-net ball only)
    long long *timing;   // debug: [k][8] clock64 stamps of CTA 0 (NULL = off)
    unsigned *start_flag;    // optional: *start_flag = start_serial as soon as every CTA of the launch is resident
    unsigned start_serial;   //   (the second stream of modl_partial_fit_* waits for it, so that its kernels cannot
                             //   take the SMs this launch needs to be placed)
};

// "the whole launch is resident": called by one thread after the first cluster / grid-wide barrier
template <typename T>
__device__ __forceinline__ void bcd_signal_start(const BcdParams<T> &P)
{
    if (P.start_flag != nullptr) {
        *reinterpret_cast<volatile unsigned *>(P.start_flag) = P.start_serial;
        __threadfence_system();
    }
}

// 1 / sqrt(x): the hardware approximation plus one Newton step (float), exact division (double)
__device__ __forceinline__ float bcd_rsqrt(float x)
{
    const float y = rsqrtf(x);
    return fmaf(0.5f * y, fmaf(-x * y, y, 1.f), y);
}
__device__ __forceinline__ double bcd_rsqrt(double x) { return 1.0 / sqrt(x); }

template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gmem_src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(sa), "l"(gmem_src), "n"(sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- cluster exchange: st.async pushes 16 (float) / 32 (double) bytes into a peer's slot and
// completes that many bytes on the PEER's mbarrier: one-way latency, no fences, no barrier ----
__device__ __forceinline__ unsigned mapa_u32(unsigned local_smem_addr, unsigned cta_rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_async_triplet(unsigned remote_slot, unsigned remote_mbar, float a, float b, float c)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                 ::"r"(remote_slot), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(0u),
                   "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async_triplet(unsigned remote_slot, unsigned remote_mbar, double a, double b, double c)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];\n"
                 ::"r"(remote_slot), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(remote_mbar) : "memory");
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];\n"
                 ::"r"(remote_slot + 16u), "l"(__double_as_longlong(c)), "l"(0ll), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol error becomes a trap (launch failure), never a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
#ifdef MODL_UNBOUNDED_WAIT
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    return;
#endif
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && spin > (1u << 22)) __trap();
    }
}

__device__ __forceinline__ void bcd_arrive(bool use_cluster, unsigned *bar, unsigned &epoch)
{
    if (use_cluster) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
        if (threadIdx.x == 0) {
            epoch += 1;
            __threadfence();
            atomicAdd(bar, 1u);
        }
    }
}

__device__ __forceinline__ void bcd_wait(bool use_cluster, unsigned *bar, unsigned nblk, unsigned epoch)
{
    if (use_cluster) {
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        if (threadIdx.x == 0) {
            const unsigned target = epoch * nblk;
            while (true) {
                unsigned cur;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(cur) : "l"(bar) : "memory");
                if (cur >= target) break;
                __nanosleep(20);
            }
            __threadfence();
        }
        __syncthreads();
    }
}

// Grid barrier without a shared counter ("bcd_flag_barrier"): every CTA owns one flag per atom parity; publishing is a
// release store of the atom's serial number into the CTA's own flag (no atomics, no contention on one address), waiting
// is warp 0 polling the nblk flags with ONE coalesced load per round.  The data the flag guards (this CTA's partial
// sums, its slice of the candidate row) is written before the __syncthreads + fence that precede the store.
__device__ __forceinline__ void bcd_flag_arrive(unsigned *flags, int par, int nblk, int g, unsigned serial)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(flags + par * nblk + g), "r"(serial) : "memory");
    }
}
__device__ __forceinline__ void bcd_flag_wait(const unsigned *flags, int par, int nblk, unsigned serial)
{
    if (threadIdx.x < 32) {
        const unsigned *f = flags + par * nblk;
        while (true) {
            bool ok = true;
            for (int q = threadIdx.x; q < nblk; q += 32) {
                unsigned cur;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(cur) : "l"(f + q) : "memory");
                ok = ok && (cur >= serial);
            }
            if (__all_sync(kFullMask, ok)) break;
            __nanosleep(20);
        }
        __threadfence();
    }
    __syncthreads();
}

// dotn[c] = sum_i crow[i] * D[i][c] over ALL rows, for the CTA's columns (all threads
// cooperate).  Shared-memory path: a thread owns a PAIR of adjacent columns and a contiguous
// block of rows, so the C row is read as 128-bit and the D slice as 64-bit shared loads
// (~1.6 instructions per FMA); partial sums of the row blocks are combined in a fixed order.
template <typename T> struct alignas(2 * sizeof(T)) Pair { T x, y; };
__device__ __forceinline__ void pair_fma(float c, const Pair<float> &d, Pair<float> &acc)
{
    const float2 r = __ffma2_rn(make_float2(c, c), make_float2(d.x, d.y), make_float2(acc.x, acc.y));
    acc.x = r.x; acc.y = r.y;
}
__device__ __forceinline__ void pair_fma(double c, const Pair<double> &d, Pair<double> &acc)
{
    acc.x = fma(c, d.x, acc.x); acc.y = fma(c, d.y, acc.y);
}
template <typename T> struct alignas(16) Quad { T x, y, z, w; };

template <typename T>
__device__ __forceinline__ void bcd_row_product(const T *crow, const T *Ds, const T *Dg, int64_t lds, bool d_in_smem,
                                                int k, int nc, int ncp, int CW, T *red, T *dotn)
{
    const int tid = threadIdx.x;
    if (d_in_smem) {
        const int NP = ncp >> 1;                       // column pairs
        const int IG = BCD_THREADS / NP;               // row blocks (>= 1 because ncp <= 2 * BCD_THREADS)
        const int RB = (((k + IG - 1) / IG) + 3) & ~3; // rows per block, multiple of 4
        const int pr = tid % NP, ig = tid / NP;
        if (ig < IG) {
            const int r0 = ig * RB, r1 = min(k, r0 + RB);
            const Pair<T> *dcol = reinterpret_cast<const Pair<T> *>(Ds) + pr;      // Ds[i][2*pr .. 2*pr+1]
            const int rs = ncp >> 1;                                              // row stride in pairs
            Pair<T> accA = {T(0), T(0)}, accB = {T(0), T(0)};                      // two independent FMA chains
            int i = r0;
#pragma unroll 2
            for (; i + 4 <= r1; i += 4) {
                const Quad<T> cq = *reinterpret_cast<const Quad<T> *>(crow + i);
                const Pair<T> d0 = dcol[(i + 0) * rs], d1 = dcol[(i + 1) * rs], d2 = dcol[(i + 2) * rs], d3 = dcol[(i + 3) * rs];
                pair_fma(cq.x, d0, accA);
                pair_fma(cq.y, d1, accB);
                pair_fma(cq.z, d2, accA);
                pair_fma(cq.w, d3, accB);
            }
            for (; i < r1; ++i) pair_fma(crow[i], dcol[i * rs], accA);
            red[ig * ncp + 2 * pr] = accA.x + accB.x;
            red[ig * ncp + 2 * pr + 1] = accA.y + accB.y;
        }
        __syncthreads();
        for (int c = tid; c < nc; c += BCD_THREADS) {
            T dot = T(0);
            for (int gi = 0; gi < IG; ++gi) dot += red[gi * ncp + c];
            dotn[c] = dot;
        }
        __syncthreads();
    } else {
        const int IG = BCD_THREADS / CW;
        const int cl = tid % CW, ig = tid / CW;
        for (int cb = 0; cb < nc; cb += CW) {
            const int c = cb + cl;
            T a0 = T(0);
            if (ig < IG && c < nc) {
                const T *col = Dg + c;
#pragma unroll 4
                for (int i = ig; i < k; i += IG) a0 = fma(crow[i], col[(int64_t)i * lds], a0);
            }
            if (ig < IG) red[ig * CW + cl] = a0;
            __syncthreads();
            if (ig == 0 && c < nc) {
                T dot = T(0);
                for (int gi = 0; gi < IG; ++gi) dot += red[gi * CW + cl];
                dotn[c] = dot;
            }
            __syncthreads();
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(BCD_THREADS, 1)
bcd_update_kernel(BcdParams<T> P)
{
    extern __shared__ __align__(16) unsigned char bcd_smem_raw[];
    const int k = P.k, s = P.s, lds = P.lds;
    const int nblk = gridDim.x, g = blockIdx.x;
    const int c0 = min(s, g * P.cols_per_cta);
    const int c1 = min(s, c0 + P.cols_per_cta);
    const int nc = c1 - c0;                       // columns owned (may be 0)
    const int ncp = (int)round_up(P.cols_per_cta, 32);
    const int kp = (int)round_up(k, 32);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int CW = P.chunk;
    const bool enet = P.l1_ratio != T(0);
    const bool use_cluster = P.use_cluster != 0;
    const bool d_in_smem = P.d_in_smem != 0;

    // ---- shared memory carve-up (all offsets multiples of 32 elements) ----
    T *Ca = reinterpret_cast<T *>(bcd_smem_raw);              // 3 * kp : ring of C rows
    T *cnorm = Ca + 3 * kp;                                   // kp     : comp_norm_ staged
    T *red = cnorm + kp;                                      // max(BCD_THREADS, row blocks * ncp)
    T *vown = red + bcd_red_elems(ncp);                       // ncp : candidate / new atom slice
    T *dotn = vown + ncp;                                     // ncp : C[a,:] . D for the atom being visited
    T *brow = dotn + ncp;                                     // 2 * ncp : ring of B_sub rows
    T *xch = brow + 2 * ncp;                                  // 2 * BCD_MAX_CLUSTER * BCD_NPART exchange slots
    T *wred = xch + 2 * BCD_MAX_CLUSTER * BCD_NPART;          // 16 warps * 4
    double *dscratch = reinterpret_cast<double *>(wred + 64); // 40 doubles
    T *Ds = reinterpret_cast<T *>(dscratch + 40);             // k * ncp (optional)
    __shared__ bool sh_inside;
    __shared__ T sh_l;
    __shared__ __align__(8) unsigned long long xbar[2];      // one mbarrier per atom parity (cluster exchange)
    const unsigned xbar_addr = (unsigned)__cvta_generic_to_shared(xbar);
    const unsigned xch_addr = (unsigned)__cvta_generic_to_shared(xch);
    constexpr unsigned kSlotBytes = BCD_NPART * sizeof(T);
    if (use_cluster && tid == 0) {
        mbar_init(xbar_addr, 1);
        mbar_init(xbar_addr + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }

    const T *Dg = P.Dp + c0;                                  // my columns of the global panel
    if (d_in_smem) {
        for (int e = tid; e < k * ncp; e += BCD_THREADS) {
            const int i = e / ncp, c = e % ncp;
            Ds[e] = (c < nc) ? Dg[(int64_t)i * lds + c] : T(0);
        }
    }
    for (int i = tid; i < k; i += BCD_THREADS) cnorm[i] = P.comp_norm[i];
    // C rows of the first two atoms, B row of the first atom
    {
        const int a0 = P.order[0], a1 = P.order[k > 1 ? 1 : 0];
        for (int i = tid; i < k; i += BCD_THREADS) {
            Ca[i] = P.C[(int64_t)a0 * k + i];
            Ca[kp + i] = P.C[(int64_t)a1 * k + i];
        }
        for (int c = tid; c < nc; c += BCD_THREADS) brow[c] = P.Bp[(int64_t)a0 * lds + c0 + c];
    }
    __syncthreads();
    bcd_row_product<T>(Ca, Ds, Dg, lds, d_in_smem, k, nc, ncp, CW, red, dotn);

    cg::cluster_group cluster = cg::this_cluster();
    if (use_cluster) {   // every peer is resident before anybody stores into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    if (g == 0 && tid == 0) bcd_signal_start(P);   // cluster and cooperative launches are co-scheduled: all CTAs are resident
    unsigned epoch = 0;
    T na_carry = T(0);           // enet_norm partial of the previous atom held by this thread
    T radius_prev = T(0);
    int a_prev = -1;

    for (int oi = 0; oi <= k; ++oi) {
        const bool live = oi < k;                  // oi == k: only flushes the last atom's norm
        const int a = live ? P.order[oi] : 0;
        const int an = (oi + 1 < k) ? P.order[oi + 1] : -1;
        const int par = oi & 1;
        const T *Ccur = Ca + (oi % 3) * kp;
        const T *Cnext = Ca + ((oi + 1) % 3) * kp;
        T *Cpre = Ca + ((oi + 2) % 3) * kp;
        const T *bcur = brow + par * ncp;
        T *bnext = brow + (par ^ 1) * ncp;
        const T caa = live ? Ccur[a] : T(1);

#define BCD_STAMP(slot) do { if (P.timing && g == 0 && tid == 0 && live) P.timing[(int64_t)oi * 8 + (slot)] = clock64(); } while (0)
        BCD_STAMP(0);
        // ---------------- S1: candidate row on my columns + partial sums ----------------
        T nb_local = T(0), sv2_local = T(0);
        if (live) {
            for (int c = tid; c < nc; c += BCD_THREADS) {
                const T dold = d_in_smem ? Ds[a * ncp + c] : Dg[(int64_t)a * lds + c];
                const T grad = (bcur[c] - dotn[c]) + caa * dold;
                T v = (caa > T(1e-20)) ? grad / caa : dold;          // [ref: :681-683]
                if (P.positive && v < T(0)) v = T(0);                // [ref: :684-685]
                vown[c] = v;
                nb_local += enet_term(dold, P.l1_ratio);             // [ref: :676-678]
                sv2_local = fma(v, v, sv2_local);
                if (enet) P.vrow[(int64_t)par * s + c0 + c] = v;
            }
        }
        {
            T r0 = warp_sum(nb_local), r1 = warp_sum(sv2_local), r2 = warp_sum(na_carry);
            if (lane == 0) { wred[wid * 4 + 0] = r0; wred[wid * 4 + 1] = r1; wred[wid * 4 + 2] = r2; }
        }
        __syncthreads();
        if (wid == 0) {
            // fixed-order sum over the warps (every lane the same), then publish to every CTA (slot [par][g])
            constexpr int NW = BCD_THREADS / 32;
            T t0 = T(0), t1 = T(0), t2 = T(0);
#pragma unroll
            for (int w = 0; w < NW; ++w) { t0 += wred[w * 4]; t1 += wred[w * 4 + 1]; t2 += wred[w * 4 + 2]; }
            if (use_cluster) {
                if (enet) __threadfence();      // my slice of the candidate row (global) before the signal
                if (lane == 0) mbar_expect_tx(xbar_addr + 8 * par, (unsigned)nblk * kSlotBytes);   // what I will receive
                if (lane < nblk)
                    st_async_triplet(mapa_u32(xch_addr + (unsigned)(par * BCD_MAX_CLUSTER + g) * kSlotBytes, (unsigned)lane),
                                     mapa_u32(xbar_addr + 8 * par, (unsigned)lane), t0, t1, t2);
            } else if (lane == 0) {
                T *dst = P.part + ((int64_t)par * nblk + g) * BCD_NPART;
                dst[0] = t0; dst[1] = t1; dst[2] = t2;
            }
        }
        BCD_STAMP(1);
        if (!use_cluster) {
            if (P.flag_barrier) bcd_flag_arrive(P.flags, par, nblk, g, (unsigned)oi + 1u);
            else bcd_arrive(false, P.bar, epoch);
        }
        BCD_STAMP(2);

        // ---------------- S2 (overlaps the barrier): prefetch + next atom's row product ----------------
        if (oi + 2 < k) {
            const int a2 = P.order[oi + 2];
            for (int i = tid; i < k; i += BCD_THREADS) cp_async_elem(Cpre + i, P.C + (int64_t)a2 * k + i);
        }
        if (an >= 0)
            for (int c = tid; c < nc; c += BCD_THREADS) cp_async_elem(bnext + c, P.Bp + (int64_t)an * lds + c0 + c);
        cp_async_commit();
        if (an >= 0)
            bcd_row_product<T>(Cnext, Ds, Dg, lds, d_in_smem, k, nc, ncp, CW, red, dotn);

        BCD_STAMP(3);
        if (use_cluster) {
            mbar_wait(xbar_addr + 8 * par, (unsigned)((oi >> 1) & 1));
            if (enet) __threadfence();
        } else {
            if (P.flag_barrier) bcd_flag_wait(P.flags, par, nblk, (unsigned)oi + 1u);
            else bcd_wait(false, P.bar, (unsigned)nblk, epoch);
        }
        BCD_STAMP(4);

        // ---------------- S3: global sums, projection, write-back, fix-up ----------------
        T nb = T(0), sv2 = T(0), na_prev = T(0);
        if (use_cluster) {
            // every warp reduces the <= 16 CTA partials with the same shuffle tree (bit-identical everywhere)
            const T *src = xch + (par * BCD_MAX_CLUSTER + lane) * BCD_NPART;
            const bool on = lane < nblk;
            nb = warp_sum(on ? src[0] : T(0)); sv2 = warp_sum(on ? src[1] : T(0)); na_prev = warp_sum(on ? src[2] : T(0));
        } else {
            // lanes stride over the CTA partials, then the same shuffle tree in every warp of every CTA
            const T *src = P.part + (int64_t)par * nblk * BCD_NPART;
            T p0 = T(0), p1 = T(0), p2 = T(0);
            for (int q = lane; q < nblk; q += 32) {
                p0 += __ldcg(src + q * BCD_NPART); p1 += __ldcg(src + q * BCD_NPART + 1); p2 += __ldcg(src + q * BCD_NPART + 2);
            }
            nb = warp_sum(p0); sv2 = warp_sum(p1); na_prev = warp_sum(p2);
        }
        if (a_prev >= 0 && tid == 0)
            cnorm[a_prev] = radius_prev - na_prev;                    // comp_norm_[k] -= subset_norm  [ref: :690-692]
        if (!live) break;
        const T radius = cnorm[a] + nb;                               // comp_norm_[k] += subset_norm  [ref: :676-678]

        T lthr = T(0), gamma = T(0), nrm = T(1);
        int mode = 0;                                                 // 0: scale by 1/nrm, 1: zero, 2: shrink by lthr
        if (radius == T(0)) {                                         // [ref: enet.pyx:57-59]
            mode = 1;
        } else if (!enet) {                                           // [ref: enet.pyx:62-70]
            nrm = (sv2 <= radius) ? T(1) : t_sqrt(sv2 / radius);
        } else {                                                      // [ref: enet.pyx:72-121]
            gamma = T(2) / P.l1_ratio - T(2);
            const T R = radius / P.l1_ratio;
            const T *vr = P.vrow + (int64_t)par * s;
            bool ins;
            // the whole candidate row, redundantly in every CTA (no extra grid barrier): held in registers when it
            // fits (96 floats / 48 doubles per thread), else re-read from L2 on every pass of the fixed point
            constexpr int NV_BIG = sizeof(T) == 4 ? 96 : 48;
            auto ld = [&](int j) { return __ldcg(vr + j); };
            T l;
            if (s <= 16 * BCD_THREADS) l = enet_threshold_block_regs<T, 16>(ld, s, R, gamma, &ins, dscratch);
            else if (s <= NV_BIG * BCD_THREADS) l = enet_threshold_block_regs<T, NV_BIG>(ld, s, R, gamma, &ins, dscratch);
            else l = enet_threshold_block<T>(ld, s, R, gamma, &ins, dscratch);
            if (tid == 0) { sh_inside = ins; sh_l = l; }
            __syncthreads();
            if (!sh_inside) { mode = 2; lthr = sh_l; }
        }
        BCD_STAMP(5);
        cp_async_wait_all();                                          // next C / B rows have landed (this thread's part)
        BCD_STAMP(6);
        const T can = (an >= 0) ? Cnext[a] : T(0);
        na_carry = T(0);
        for (int c = tid; c < nc; c += BCD_THREADS) {
            T v = vown[c];
            if (mode == 1) v = T(0);
            else if (mode == 2) v = enet_shrink(v, lthr, gamma);
            else v = v / nrm;
            na_carry += enet_term(v, P.l1_ratio);
            T dold;
            if (d_in_smem) {
                dold = Ds[a * ncp + c];
                Ds[a * ncp + c] = v;                                  // global panel is refreshed once, at the end
            } else {
                dold = Dg[(int64_t)a * lds + c];
                P.Dp[(int64_t)a * lds + c0 + c] = v;
            }
            // the next row product was taken over the OLD row a: rank-1 correction with the change
            dotn[c] = fma(can, v - dold, dotn[c]);
        }
        radius_prev = radius;
        a_prev = a;
        __syncthreads();
        BCD_STAMP(7);
    }
#undef BCD_STAMP
    // ---- write-back ----
    __syncthreads();
    if (use_cluster) {   // nobody leaves while a peer could still address its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    }
    if (d_in_smem) {
        for (int e = tid; e < k * ncp; e += BCD_THREADS) {
            const int i = e / ncp, c = e % ncp;
            if (c < nc) P.Dp[(int64_t)i * lds + c0 + c] = Ds[e];
        }
    }
    if (g == 0)
        for (int i = tid; i < k; i += BCD_THREADS) P.comp_norm[i] = cnorm[i];
}

}  // namespace modl
