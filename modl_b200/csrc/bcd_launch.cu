// bcd_launch.cu -- launch geometry of the dictionary-update kernels (bcd_kernels.cuh, bcd_pilot.cuh).
#include "bcd_kernels.cuh"
#include "bcd_pilot.cuh"
#include "bcd_blocked.cuh"
#include "launch.h"

#include <mutex>
#include <vector>

namespace modl {

// ---------------------------------------------------------------------------------------
// dictionary update
// ---------------------------------------------------------------------------------------
template <typename T, bool ENET, bool PIPE>
static const void *pilot_kernel_for_e(int ncl)
{
    switch (ncl) {
        case 1: return (const void *)bcd_pilot_kernel<T, 1, ENET, PIPE>;
        case 2: return (const void *)bcd_pilot_kernel<T, 2, ENET, PIPE>;
        case 3: return (const void *)bcd_pilot_kernel<T, 3, ENET, PIPE>;
        case 4: return (const void *)bcd_pilot_kernel<T, 4, ENET, PIPE>;
        case 5: return (const void *)bcd_pilot_kernel<T, 5, ENET, PIPE>;
        default: return (const void *)bcd_pilot_kernel<T, 6, ENET, PIPE>;
    }
}
// pipe: two norm exchanges in flight (L2 ball, no positivity clamp: the candidate row is linear in the previous scale)
template <typename T>
static const void *pilot_kernel_for(int ncl, bool enet, bool pipe)
{
    if (enet) return pilot_kernel_for_e<T, true, false>(ncl);
    return pipe ? pilot_kernel_for_e<T, false, true>(ncl) : pilot_kernel_for_e<T, false, false>(ncl);
}

template <typename T>
int bcd_update(modl_ctx *ctx, T *Dp, const T *Bp, int64_t lds, const T *C, T *comp_norm,
                      const int32_t *d_order, int64_t k, int64_t s, T l1_ratio, int positive, cudaStream_t st,
                      unsigned *start_flag, unsigned start_serial)
{
    if (s <= 0 || k <= 0) return MODL_OK;
    auto kern = bcd_update_kernel<T>;
    const size_t budget = (size_t)ctx->max_smem_optin - 1024;
    auto base_smem = [&](int64_t ncp) {
        return (size_t)(4 * round_up(k, 32) + bcd_red_elems(ncp) + 4 * ncp + 2 * BCD_MAX_CLUSTER * BCD_NPART + 64) * sizeof(T) +
               40 * sizeof(double);
    };
    BcdParams<T> P;
    P.Dp = Dp; P.Bp = Bp; P.C = C; P.comp_norm = comp_norm; P.order = d_order;
    P.k = (int)k; P.s = (int)s; P.lds = (int)lds; P.l1_ratio = l1_ratio; P.positive = positive;
    P.start_flag = start_flag; P.start_serial = start_serial;

    const bool pipe = ctx->opt_bcd_pipeline != 0 && l1_ratio == T(0) && !positive;
    int nblk = 0, use_cluster = 0, d_in_smem = 0, use_pilot = 0;
    int64_t cols = 0;
    size_t smem = 0;
    // 1) a single thread-block cluster with the panel resident in shared memory
    if (ctx->cluster_ok && ctx->opt_bcd_cluster >= 2) {
        for (int cs = 16; cs >= 2 && !nblk; cs >>= 1) {
            if (cs > ctx->opt_bcd_cluster) continue;
            const int64_t c = round_up(ceil_div(s, cs), 4), ncp = round_up(c, 32);
            // variant 2 = block-wise coefficient-space solve (L2 ball, no positivity), 1 = per-atom look-ahead pilot,
            // 0 = plain per-atom kernel
            const bool blocked_ok = ctx->opt_bcd_blocked != 0 && l1_ratio == T(0) && !positive;
            for (int pilot = blocked_ok ? 2 : (ctx->opt_bcd_pilot ? 1 : 0); pilot >= 0 && !nblk; --pilot) {
                size_t need;
                if (pilot == 2) {
                    if (!bcd_blocked_fits(ncp)) continue;
                    need = bcd_blocked_smem_bytes<T>(k, ncp);
                } else if (pilot == 1 && !ctx->opt_bcd_pilot) {
                    continue;
                } else if (pilot) {
                    if (ncp > 192) continue;
                    need = bcd_pilot_smem_bytes<T>(k, ncp);
                } else {
                    need = base_smem(ncp) + (size_t)k * ncp * sizeof(T);
                    if (ncp > 2 * BCD_THREADS) continue;
                }
                if (need > budget) continue;
                const void *fn = pilot == 2 ? (const void *)bcd_blocked_kernel<T>
                                 : pilot ? pilot_kernel_for<T>((int)(ncp / 32), l1_ratio != T(0), pipe) : (const void *)kern;
                // (kernel, cluster size, shared memory) combinations already validated on this device: skip the
                // attribute and occupancy queries (tens of microseconds of host time per step)
                struct Seen { const void *fn; int cs; size_t need; int device; };
                static std::vector<Seen> seen;
                static std::mutex seen_mu;
                bool known = false;
                {
                    std::lock_guard<std::mutex> lk(seen_mu);
                    for (const Seen &e : seen) known = known || (e.fn == fn && e.cs == cs && e.need >= need && e.device == ctx->device);
                }
                if (known) {
                    nblk = cs; use_cluster = 1; d_in_smem = 1; cols = c; smem = need; use_pilot = pilot;
                    continue;
                }
                if (cs > 8 && cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                // always the whole opt-in budget: the attribute is per FUNCTION, and a later, smaller request (another
                // cluster size of the same kernel) would otherwise lower it under a cached larger one
                if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) {
                    cudaGetLastError();
                    continue;
                }
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs); cfg.blockDim = dim3(pilot == 2 ? BB_THREADS : pilot ? BP_THREADS : BCD_THREADS);
                cfg.dynamicSmemBytes = need; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int nclusters = 0;
                if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1) {
                    cudaGetLastError();
                    continue;
                }
                nblk = cs; use_cluster = 1; d_in_smem = 1; cols = c; smem = need; use_pilot = pilot;
                {
                    std::lock_guard<std::mutex> lk(seen_mu);
                    seen.push_back(Seen{fn, cs, need, ctx->device});
                }
            }
        }
    }
    // 2) cooperative launch over the SMs with a global barrier
    if (!nblk) {
        // one CTA per >= 32 columns by default; "bcd_coop_min_cols" trades row-product parallelism (tiny: k x ncp FMAs
        // per atom and CTA) for fewer participants at the per-atom grid barrier
        int64_t want = ceil_div(s, ctx->opt_bcd_coop_min_cols > 32 ? ctx->opt_bcd_coop_min_cols : 32);
        if (want > ctx->sm_count) want = ctx->sm_count;
        if (want < 1) want = 1;
        cols = ceil_div(s, want);
        const int64_t ncp = round_up(cols, 32);
        size_t need = base_smem(ncp) + (size_t)k * ncp * sizeof(T);
        d_in_smem = need <= budget && ncp <= 2 * BCD_THREADS;
        if (!d_in_smem) need = base_smem(ncp);
        MODL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));   // per function: never lower it
        int per_sm = 0;
        MODL_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BCD_THREADS, need));
        MODL_REQUIRE(per_sm >= 1, "dictionary-update kernel does not fit on an SM");
        nblk = (int)ceil_div(s, cols);
        MODL_REQUIRE(nblk <= per_sm * ctx->sm_count, "cooperative grid too large");
        smem = need;
    }
    const int64_t nchunks = ceil_div(cols, 128);
    int64_t cw = round_up(ceil_div(cols, nchunks), 32);
    if (cw > 128) cw = 128;
    P.cols_per_cta = (int)cols; P.chunk = (int)cw; P.d_in_smem = d_in_smem; P.use_cluster = use_cluster;

    // exchange workspace: [barrier 256 B][flags 2*nblk u32, padded][part 2*nblk*4][vrow 2*s]
    const size_t n_t = (size_t)(2 * nblk * BCD_NPART) + (size_t)nblk * k + 2 * (size_t)s;
    const size_t sync_bytes = 256 + (size_t)round_up(2 * nblk * (int64_t)sizeof(unsigned), 256);
    unsigned char *base = nullptr;
    MODL_TRY(ws<unsigned char>(ctx, WS_BCD_SYNC, sync_bytes + n_t * sizeof(T), &base));
    P.bar = reinterpret_cast<unsigned *>(base);
    P.flags = reinterpret_cast<unsigned *>(base + 256);
    P.flag_barrier = ctx->opt_bcd_flag_barrier;
    T *tb = reinterpret_cast<T *>(base + sync_bytes);
    P.part = tb; tb += 2 * nblk * BCD_NPART + (size_t)nblk * k;
    P.vrow = tb;
    MODL_CUDA_TRY(cudaMemsetAsync(P.bar, 0, sync_bytes, st));
    P.timing = nullptr;
    if (ctx->opt_bcd_timing) {
        long long *tbuf = nullptr;
        MODL_TRY(ws<long long>(ctx, WS_MISC, (size_t)(8 * k + 16 + 2 * ceil_div(k, 8) + 8), &tbuf));
        P.timing = tbuf;
        ctx->bcd_timing_k = (int)k;
    }

    ctx->bcd_grid_wide = use_cluster ? 0 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(use_pilot == 2 ? BB_THREADS : use_pilot ? BP_THREADS : BCD_THREADS);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (use_cluster) {
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = nblk; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    } else {
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
    }
    cfg.attrs = at; cfg.numAttrs = 1;
    if (use_pilot == 2) {
        MODL_CUDA_TRY(cudaLaunchKernelEx(&cfg, bcd_blocked_kernel<T>, P));
    } else if (use_pilot) {
        void *args[] = {&P};
        MODL_CUDA_TRY(cudaLaunchKernelExC(&cfg, pilot_kernel_for<T>((int)(round_up(cols, 32) / 32), l1_ratio != T(0), pipe), args));
    } else {
        MODL_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
    }
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}


template int bcd_update<float>(modl_ctx *, float *, const float *, int64_t, const float *, float *, const int32_t *, int64_t,
                               int64_t, float, int, cudaStream_t, unsigned *, unsigned);
template int bcd_update<double>(modl_ctx *, double *, const double *, int64_t, const double *, double *, const int32_t *,
                                int64_t, int64_t, double, int, cudaStream_t, unsigned *, unsigned);

}  // namespace modl
