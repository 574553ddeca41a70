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
// needs sums over all s columns: each CTA publishes partial sums (and, for the elastic-net
// ball, its slice of the candidate row), all CTAs meet at ONE barrier per atom, then every
// CTA reduces the partials in the same fixed order (bitwise identical decision everywhere,
// no atomics) and finishes its own columns.  The barrier is a thread-block-cluster hardware
// barrier when the launch is a single cluster (<= 16 CTAs; ~0.2 us), else a global-memory
// sense barrier under a cooperative launch.
#pragma once
#include "basic_kernels.cuh"
#include "common.cuh"

namespace modl {

constexpr int BCD_THREADS = 512;
constexpr int BCD_NPART = 4;     // partial sums exchanged per atom and CTA

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
    int use_cluster;     // 1: hardware cluster barrier, 0: global sense barrier
    unsigned *bar;       // global barrier counter (zero on entry)
    T *part;             // [2][nblk][BCD_NPART]
    T *vrow;             // [2][s]   candidate rows (elastic-net ball only)
    T *na_part;          // [nblk][k] per-CTA partial of enet_norm(new atom)
    T *radius_log;       // [k] radius used for atom a (written by CTA 0)
};

__device__ __forceinline__ void bcd_barrier(bool use_cluster, unsigned *bar, unsigned nblk, unsigned &epoch)
{
    if (use_cluster) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    } else {
        __syncthreads();
        if (threadIdx.x == 0) {
            epoch += 1;
            __threadfence();
            atomicAdd(bar, 1u);
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
    const int tid = threadIdx.x;
    const int CW = P.chunk, IG = BCD_THREADS / CW;
    const int cl = tid % CW, ig = tid / CW;       // column lane / row group (ig >= IG: idle)
    const bool enet = P.l1_ratio != T(0);

    // ---- shared memory carve-up ----
    T *Ca = reinterpret_cast<T *>(bcd_smem_raw);              // k      : row a of C
    T *red = Ca + round_up(k, 32);                        // IG*CW  : cross-group partial dots
    T *vown = red + BCD_THREADS;                              // ncp    : candidate / new atom slice
    T *scratch = vown + ncp;                                  // 40
    double *dscratch = reinterpret_cast<double *>(scratch + 40);   // 40 doubles (8-byte aligned below)
    T *Ds = reinterpret_cast<T *>(dscratch + 40);             // k * ncp (optional)
    __shared__ bool sh_inside;
    __shared__ T sh_l;

    if (P.d_in_smem) {
        for (int e = tid; e < k * ncp; e += BCD_THREADS) {
            const int i = e / ncp, c = e % ncp;
            Ds[e] = (c < nc) ? P.Dp[(int64_t)i * lds + c0 + c] : T(0);
        }
    }
    __syncthreads();

    unsigned epoch = 0;
    for (int oi = 0; oi < k; ++oi) {
        const int a = P.order[oi];
        const int par = oi & 1;
        for (int i = tid; i < k; i += BCD_THREADS) Ca[i] = P.C[(int64_t)a * k + i];
        __syncthreads();
        const T caa = Ca[a];
        T nb_local = T(0), sv2_local = T(0);

        // ---------------- phase A: candidate row on my columns ----------------
        for (int cb = 0; cb < nc; cb += CW) {
            const int c = cb + cl;                      // column within my slice
            T acc = T(0);
            if (ig < IG && c < nc) {
                if (P.d_in_smem) {
                    for (int i = ig; i < k; i += IG) acc = fma(Ca[i], Ds[i * ncp + c], acc);
                } else {
                    const T *col = P.Dp + c0 + c;
                    for (int i = ig; i < k; i += IG) acc = fma(Ca[i], col[(int64_t)i * lds], acc);
                }
            }
            if (ig < IG) red[ig * CW + cl] = acc;
            __syncthreads();
            if (ig == 0 && c < nc) {
                T dot = T(0);
                for (int gi = 0; gi < IG; ++gi) dot += red[gi * CW + cl];
                const T dold = P.d_in_smem ? Ds[a * ncp + c] : P.Dp[(int64_t)a * lds + c0 + c];
                const T grad = (P.Bp[(int64_t)a * lds + c0 + c] - dot) + caa * dold;
                T v = (caa > T(1e-20)) ? grad / caa : dold;         // [ref: :681-683]
                if (P.positive && v < T(0)) v = T(0);               // [ref: :684-685]
                vown[c] = v;
                nb_local += enet_term(dold, P.l1_ratio);            // [ref: :676-678]
                sv2_local = fma(v, v, sv2_local);
                if (enet) P.vrow[(int64_t)par * s + c0 + c] = v;
            }
            __syncthreads();
        }
        nb_local = block_sum(nb_local, scratch);
        sv2_local = block_sum(sv2_local, scratch);
        if (tid == 0) {
            T *dst = P.part + ((int64_t)par * nblk + g) * BCD_NPART;
            dst[0] = nb_local;
            dst[1] = sv2_local;
        }

        bcd_barrier(P.use_cluster != 0, P.bar, (unsigned)nblk, epoch);

        // ---------------- phase B: global sums, projection, write-back ----------------
        T nb = T(0), sv2 = T(0);
        {
            const T *src = P.part + (int64_t)par * nblk * BCD_NPART;
            for (int q = 0; q < nblk; ++q) {          // same order in every CTA
                nb += __ldcg(src + q * BCD_NPART + 0);
                sv2 += __ldcg(src + q * BCD_NPART + 1);
            }
        }
        const T radius = P.comp_norm[a] + nb;         // comp_norm_[k] += subset_norm
        if (g == 0 && tid == 0) P.radius_log[a] = radius;

        T na_local = T(0);
        if (radius == T(0)) {                                         // [ref: enet.pyx:57-59]
            for (int c = tid; c < nc; c += BCD_THREADS) vown[c] = T(0);
        } else if (!enet) {                                           // [ref: enet.pyx:62-70]
            const T nrm = (sv2 <= radius) ? T(1) : t_sqrt(sv2 / radius);
            for (int c = tid; c < nc; c += BCD_THREADS) vown[c] = vown[c] / nrm;
        } else {                                                      // [ref: enet.pyx:72-121]
            const T gamma = T(2) / P.l1_ratio - T(2);
            const T R = radius / P.l1_ratio;
            const T *vr = P.vrow + (int64_t)par * s;
            bool ins;
            const T l = enet_threshold_block<T>([&](int j) { return __ldcg(vr + j); }, s, R, gamma, &ins, dscratch);
            if (tid == 0) { sh_inside = ins; sh_l = l; }
            __syncthreads();
            if (!sh_inside) {
                const T lt = sh_l;
                for (int c = tid; c < nc; c += BCD_THREADS) vown[c] = enet_shrink(vown[c], lt, gamma);
            }
        }
        __syncthreads();
        for (int c = tid; c < nc; c += BCD_THREADS) {
            const T v = vown[c];
            na_local += enet_term(v, P.l1_ratio);                     // [ref: :690-692]
            if (P.d_in_smem) Ds[a * ncp + c] = v;
            P.Dp[(int64_t)a * lds + c0 + c] = v;                      // write-through
        }
        na_local = block_sum(na_local, scratch);
        if (tid == 0) P.na_part[(int64_t)g * k + a] = na_local;
        __syncthreads();
    }

    // ---- comp_norm_[a] = radius_a - enet_norm(new atom): final fixed-order reduction ----
    bcd_barrier(P.use_cluster != 0, P.bar, (unsigned)nblk, epoch);
    if (g == 0) {
        for (int a = tid; a < k; a += BCD_THREADS) {
            T na = T(0);
            for (int q = 0; q < nblk; ++q) na += __ldcg(P.na_part + (int64_t)q * k + a);
            P.comp_norm[a] = __ldcg(P.radius_log + a) - na;
        }
    }
}

}  // namespace modl
