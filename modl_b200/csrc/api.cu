// api.cu -- C-ABI entry points (include/modl_b200.h): context, launch geometry, and the
// fused minibatch step that mirrors DictFact._single_batch_fit
// [ref: modl/decomposition/dict_fact.py:495-526].
#include <cstdarg>
#include <cstdlib>
#include <mutex>
#include <new>
#include <type_traits>
#include <vector>

#include "basic_kernels.cuh"
#include "common.cuh"
#include "launch.h"

namespace modl {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace modl

using namespace modl;

int modl_ctx::reserve(WsSlot slot, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 16;
    if (slot_bytes[slot] < bytes) {
        if (capturing) return MODL_EGROW;     // cudaFree / cudaMalloc are not capturable: the caller re-runs the step eagerly
        // grow-only; freeing is stream-ordered-safe because cudaFree synchronises the device
        if (slot_ptr[slot]) MODL_CUDA_TRY(cudaFree(slot_ptr[slot]));
        slot_ptr[slot] = nullptr;
        slot_bytes[slot] = 0;
        size_t cap = bytes + bytes / 4 + 256;
        MODL_CUDA_TRY(cudaMalloc(&slot_ptr[slot], cap));
        slot_bytes[slot] = cap;
    }
    *out = slot_ptr[slot];
    return MODL_OK;
}

extern "C" {

int modl_version(void) { return 100; }

void modl_struct_sizes(int64_t *h_out3)
{
    if (!h_out3) return;
    h_out3[0] = (int64_t)sizeof(modl_step_params);
    h_out3[1] = (int64_t)sizeof(modl_fit_params);
    h_out3[2] = (int64_t)sizeof(modl_fit_batches);
}
const char *modl_last_error(void) { return modl::g_err; }

static int ctx_init(modl_ctx *c, int device)
{
    CtxGuard g_(c);                      // makes `device` current; the caller's device is restored on return
    MODL_CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    MODL_CUDA_TRY(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    MODL_CUDA_TRY(cudaDeviceGetAttribute(&c->cluster_ok, cudaDevAttrClusterLaunch, device));
    if (const char *e = getenv("MODL_BCD_CLUSTER")) c->opt_bcd_cluster = atoi(e);
    if (const char *e = getenv("MODL_BCD_BLOCKED")) c->opt_bcd_blocked = atoi(e);
    if (const char *e = getenv("MODL_TC_RAW_B")) c->opt_tc_raw_b = atoi(e);
    if (const char *e = getenv("MODL_TC_SPLIT2")) c->opt_tc_split2 = atoi(e);
    if (const char *e = getenv("MODL_BCD_PIPELINE")) c->opt_bcd_pipeline = atoi(e);
    if (const char *e = getenv("MODL_CD_WARPS")) c->opt_cd_warps = atoi(e);
    if (const char *e = getenv("MODL_FORCE_GLOBAL_GRAM")) c->opt_force_global_gram = atoi(e);
    if (const char *e = getenv("MODL_TC_GEMM")) c->opt_tc_gemm = atoi(e);
    int *info = nullptr;
    MODL_TRY(ws<int>(c, WS_INFO, 4, &info));
    MODL_CUDA_TRY(cudaMemset(info, 0, 4 * sizeof(int)));
    return MODL_OK;
}

int modl_ctx_create(int device, modl_ctx **out)
{
    MODL_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int count = 0;
    MODL_CUDA_TRY(cudaGetDeviceCount(&count));
    MODL_REQUIRE(device >= 0 && device < count, "no such CUDA device");
    modl_ctx *c = new (std::nothrow) modl_ctx();
    if (!c) return MODL_ENOMEM;
    c->device = device;
    const int status = ctx_init(c, device);
    if (status != MODL_OK) {
        modl_ctx_destroy(c);
        return status;
    }
    *out = c;
    return MODL_OK;
}

void modl_ctx_destroy(modl_ctx *ctx)
{
    if (!ctx) return;
    {
        CtxGuard g_(ctx);
        for (int i = 0; i < WS_COUNT; ++i)
            if (ctx->slot_ptr[i]) cudaFree(ctx->slot_ptr[i]);
        for (int i = 0; i < 2; ++i)
            if (ctx->ev_fork[i]) cudaEventDestroy(ctx->ev_fork[i]);
    }
    delete ctx;
}

int modl_ctx_sm_count(const modl_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
int64_t modl_ctx_launch_count(const modl_ctx *ctx) { return ctx ? ctx->launches : 0; }

int modl_ctx_set_option(modl_ctx *ctx, const char *name, int value)
{
    MODL_REQUIRE(ctx && name, "null");
    CtxGuard g_(ctx);
    if (!strcmp(name, "bcd_cluster")) ctx->opt_bcd_cluster = value;
    else if (!strcmp(name, "cd_warps")) ctx->opt_cd_warps = value;
    else if (!strcmp(name, "force_global_gram")) ctx->opt_force_global_gram = value;
    else if (!strcmp(name, "bcd_timing")) ctx->opt_bcd_timing = value;
    else if (!strcmp(name, "tc_gemm")) ctx->opt_tc_gemm = value;
    else if (!strcmp(name, "bcd_pilot")) ctx->opt_bcd_pilot = value;
    else if (!strcmp(name, "bcd_pipeline")) ctx->opt_bcd_pipeline = value;
    else if (!strcmp(name, "bcd_blocked")) ctx->opt_bcd_blocked = value;
    else if (!strcmp(name, "tc_raw_b")) ctx->opt_tc_raw_b = value;
    else if (!strcmp(name, "tc_split2")) ctx->opt_tc_split2 = value;
    else if (!strcmp(name, "bcd_coop_min_cols")) ctx->opt_bcd_coop_min_cols = value;
    else if (!strcmp(name, "bcd_flag_barrier")) ctx->opt_bcd_flag_barrier = value;
    else if (!strcmp(name, "drop_workspace")) {
        // debug / memory pressure: give the grow-only scratch slots back (they are re-reserved by the next call that needs
        // them; a graph-replayed step that meets an empty slot falls back to plain launches for that call)
        MODL_CUDA_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < WS_COUNT; ++i) {
            if (i == WS_INFO || !ctx->slot_ptr[i]) continue;
            MODL_CUDA_TRY(cudaFree(ctx->slot_ptr[i]));
            ctx->slot_ptr[i] = nullptr;
            ctx->slot_bytes[i] = 0;
        }
        ctx->code_packed = nullptr;
        ctx->panel_b_ready = 0;
    }
    else { set_error("unknown option %s", name); return MODL_EINVAL; }
    return MODL_OK;
}

int modl_ctx_profile(modl_ctx *ctx, int enable)
{
    MODL_REQUIRE(ctx, "null ctx");
    CtxGuard g_(ctx);
    if (enable && !ctx->prof_ev[0]) {
        for (int i = 0; i < 33; ++i) MODL_CUDA_TRY(cudaEventCreate(&ctx->prof_ev[i]));
    }
    ctx->prof_on = enable ? 1 : 0;
    ctx->prof_n = 0;
    if (enable) {
        for (int i = 0; i < MODL_PROF_PHASES; ++i) ctx->prof_ms[i] = 0;
        ctx->prof_steps = 0;
    }
    return MODL_OK;
}

int modl_ctx_profile_read(modl_ctx *ctx, double *h_ms, int64_t *h_steps)
{
    MODL_REQUIRE(ctx && h_ms && h_steps, "null");
    for (int i = 0; i < MODL_PROF_PHASES; ++i) h_ms[i] = ctx->prof_ms[i];
    *h_steps = ctx->prof_steps;
    return MODL_OK;
}

// debug: mean clock cycles between the 8 stamps of the last dictionary-update launch (7 gaps)
int modl_debug_bcd_timing(modl_ctx *ctx, double *h_gaps7)
{
    MODL_REQUIRE(ctx && h_gaps7 && ctx->bcd_timing_k > 0 && ctx->slot_ptr[WS_MISC], "no timing recorded");
    CtxGuard g_(ctx);
    const int k = ctx->bcd_timing_k;
    std::vector<long long> t((size_t)8 * k);
    MODL_CUDA_TRY(cudaDeviceSynchronize());
    MODL_CUDA_TRY(cudaMemcpy(t.data(), ctx->slot_ptr[WS_MISC], sizeof(long long) * t.size(), cudaMemcpyDeviceToHost));
    for (int j = 0; j < 7; ++j) {
        double acc = 0;
        for (int i = 1; i < k; ++i) acc += (double)(t[(size_t)i * 8 + j + 1] - t[(size_t)i * 8 + j]);
        h_gaps7[j] = acc / (k - 1);
    }
    if (getenv("MODL_BCD_TIMING_VERBOSE")) {
        std::vector<long long> x(16 + 2 * ((size_t)k / 8 + 1));
        cudaMemcpy(x.data(), (long long *)ctx->slot_ptr[WS_MISC] + 8 * k, sizeof(long long) * x.size(), cudaMemcpyDeviceToHost);
        fprintf(stderr, "bcd pilot: prologue %lld, loop %lld, final sync %lld, write-back %lld cycles; first atom at +%lld\n",
                x[1] - x[0], x[2] - x[1], x[3] - x[2], x[4] - x[3], t[0] - x[0]);
        double wsync = 0, blk = 0;
        int nb = k / 8;
        for (int b = 0; b < nb; ++b) wsync += (double)(x[16 + 2 * b + 1] - x[16 + 2 * b]);
        for (int b = 1; b < nb; ++b) blk += (double)(x[16 + 2 * b] - x[16 + 2 * (b - 1)]);
        fprintf(stderr, "bcd pilot: mean wait for workers %.0f, mean block period %.0f cycles; atom 8->9 gap %lld, atom 7 end -> block1 start %lld\n",
                wsync / nb, blk / (nb - 1), t[9 * 8] - t[8 * 8 + 7], x[16 + 2] - t[7 * 8 + 7]);
    }
    return MODL_OK;
}

// debug: raw clock64 stamps of the last dictionary-update launch
int modl_debug_bcd_stamps(modl_ctx *ctx, long long *h_out, int n)
{
    MODL_REQUIRE(ctx && h_out && n > 0 && ctx->slot_ptr[WS_MISC], "no timing recorded");
    CtxGuard g_(ctx);
    MODL_CUDA_TRY(cudaDeviceSynchronize());
    MODL_CUDA_TRY(cudaMemcpy(h_out, ctx->slot_ptr[WS_MISC], sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost));
    return MODL_OK;
}

// Synchronises `stream` and reports (then clears) the sticky numerical status:
// MODL_ENOTSPD if a Cholesky pivot was non-positive since the last check.
int modl_ctx_check_info(modl_ctx *ctx, void *stream)
{
    MODL_REQUIRE(ctx, "null ctx");
    CtxGuard g_(ctx);
    int h[4] = {0, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    MODL_CUDA_TRY(cudaMemcpyAsync(h, ctx->slot_ptr[WS_INFO], sizeof(h), cudaMemcpyDeviceToHost, st));
    MODL_CUDA_TRY(cudaStreamSynchronize(st));
    if (h[0] != 0) {
        MODL_CUDA_TRY(cudaMemsetAsync(ctx->slot_ptr[WS_INFO], 0, sizeof(h), st));
        set_error("Cholesky: leading minor of order %d is not positive definite", h[0]);
        return MODL_ENOTSPD;
    }
    return MODL_OK;
}

}  // extern "C"

namespace modl {

// ---------------------------------------------------------------------------------------
// gathers
// ---------------------------------------------------------------------------------------
template <typename T>
static int gather_cols(modl_ctx *ctx, const T *src, int64_t ld, int64_t rows, int64_t p,
                       const int64_t *subset, int64_t s, T *dst, int64_t ldd, T *norm2, cudaStream_t st)
{
    if (rows <= 0) return MODL_OK;
    gather_cols_kernel<T><<<grid_for(ctx, rows, 16), 256, 0, st>>>(src, ld, (int)rows, (int)p, subset, (int)s,
                                                                    dst, ldd, norm2);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int scatter_cols(modl_ctx *ctx, const T *src, int64_t lds, int64_t rows, const int64_t *subset,
                        int64_t s, T *dst, int64_t ldd, cudaStream_t st)
{
    if (rows <= 0 || s <= 0) return MODL_OK;
    dim3 grid((unsigned)ceil_div(s, 256), (unsigned)(rows < 65535 ? rows : 65535));
    scatter_cols_kernel<T><<<grid, 256, 0, st>>>(src, lds, (int)rows, subset, (int)s, dst, ldd);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

// regression front: CD or ridge.  xnorm2 is required for CD.
template <typename T>
static int regression(modl_ctx *ctx, const T *G, int64_t g_stride, T *Dx, const T *xnorm2, T *code,
                      const int64_t *indices, T *code_batch, int64_t b, int64_t k, T l1_ratio, T alpha,
                      int positive, T tol, int max_iter, int32_t *sweeps, cudaStream_t st)
{
    if (l1_ratio == T(0)) {
        if (sweeps) MODL_CUDA_TRY(cudaMemsetAsync(sweeps, 0, sizeof(int32_t) * (size_t)b, st));
        return ridge_solve<T>(ctx, G, g_stride, Dx, code, indices, code_batch, b, k, alpha, st);
    }
    return cd_launch<T>(ctx, G, g_stride, Dx, xnorm2, code, indices, code_batch, b, k, alpha * l1_ratio,
                        alpha * (T(1) - l1_ratio), tol, max_iter, positive, sweeps, st);
}

// The small per-step inputs straight out of pinned (mapped) host memory: feature subset (int64) and atom order
// (int64 -> int32).  One launch, no copy-engine traffic.
__global__ void __launch_bounds__(256)
fetch_step_inputs_kernel(const int64_t *__restrict__ h_subset, int64_t *__restrict__ d_subset, int s,
                         const int64_t *__restrict__ h_order, int32_t *__restrict__ d_order, int k)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
    for (int i = t; i < s; i += n) d_subset[i] = h_subset[i];
    for (int i = t; i < k; i += n) d_order[i] = (int32_t)h_order[i];
}

// An event the step records in the middle of a call (other streams wait on it).  While the call is being captured into a
// graph the record becomes an event-record NODE: the event fires when the graph's launch reaches it, and a
// cudaStreamWaitEvent issued after cudaGraphLaunch waits for exactly that.
static int record_step_event(modl_ctx *ctx, void *ev, cudaStream_t st)
{
    if (ctx->capturing == 1) MODL_CUDA_TRY(cudaEventRecordWithFlags(static_cast<cudaEvent_t>(ev), st, cudaEventRecordExternal));
    else MODL_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(ev), st));
    return MODL_OK;
}

// upload the atom order as int32 (two buffers: the step in flight and the one being prefetched)
static int upload_order(modl_ctx *ctx, const int64_t *h_order, int64_t k, int32_t **d_order, cudaStream_t st, int slot = 0,
                        bool ready = false)
{
    MODL_TRY(ws<int32_t>(ctx, slot ? WS_ORDER2 : WS_ORDER, (size_t)k, d_order));
    if (ready) return MODL_OK;          // uploaded by this step's MODL_PHASE_PREFETCH
    if (ctx->capturing) return MODL_EGROW;   // a copy out of a temporary cannot live in a graph: run the step eagerly
    std::vector<int32_t> tmp((size_t)k);
    for (int64_t i = 0; i < k; ++i) {
        MODL_REQUIRE(h_order[i] >= 0 && h_order[i] < k, "order is not a permutation of range(k)");
        tmp[(size_t)i] = (int32_t)h_order[i];
    }
    // pageable source: the runtime stages the bytes before returning, so `tmp` may die here
    MODL_CUDA_TRY(cudaMemcpyAsync(*d_order, tmp.data(), sizeof(int32_t) * (size_t)k, cudaMemcpyHostToDevice, st));
    return MODL_OK;
}

// Dictionary update given an already-gathered (or to-be-gathered) D panel.
//   Dpanel: k x s (ld = s) workspace holding components_[:, subset] if panel_ready, else filled here.
template <typename T>
static int update_dict_impl(modl_ctx *ctx, T *components, int64_t ldd, const T *B, int64_t ldb, const T *C,
                            T *comp_norm, T *G_full, const int64_t *subset, int64_t s, const int64_t *h_order,
                            int64_t k, int64_t p, T comp_l1_ratio, int comp_pos, int mode, double w,
                            double step_size, T *Dpanel, bool panel_ready, cudaStream_t st, bool bpanel_ready = false,
                            int slot = 0, bool order_ready = false, unsigned *start_flag = nullptr, unsigned start_serial = 0)
{
    MODL_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (variational) or 1 (sgd)");
    if (k <= 0) return MODL_OK;
    const int64_t lds = panel_ld(s);
    T *Bp = nullptr;
    prof_mark(ctx, st, MODL_PROF_DICT_PREP);
    MODL_TRY(ws<T>(ctx, slot ? WS_PANEL_B2 : WS_PANEL_B, (size_t)(k * lds), &Bp));
    if (!panel_ready) MODL_TRY(gather_cols<T>(ctx, components, ldd, k, p, subset, s, Dpanel, lds, nullptr, st));
    if (!bpanel_ready)
        MODL_TRY(gather_cols<T>(ctx, B, ldb, k, p, subset, s, Bp, lds, nullptr, st));   // gradient_[:, subset] = B_[:, subset]
    const bool small_subset = (double)s < (double)p / 2.;
    if (G_full && small_subset && s > 0)        // G_ -= D_sub D_sub^T   [ref: :667-668]
        MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, k, k, s, T(-1), Dpanel, lds, Dpanel, lds, T(1), G_full, k, st));

    int32_t *d_order = nullptr;
    MODL_TRY(upload_order(ctx, h_order, k, &d_order, st, slot, order_ready));
    prof_mark(ctx, st, MODL_PROF_DICT_BCD);
    if (mode == 0) {
        MODL_TRY(bcd_update<T>(ctx, Dpanel, Bp, lds, C, comp_norm, d_order, k, s, comp_l1_ratio, comp_pos, st, start_flag,
                               start_serial));
    } else if (s > 0) {
        // 'sgd' branch [ref: :695-708]: no sequential dependency between atoms
        enet_norm_rows_kernel<T><<<grid_for(ctx, k, 16), 256, 0, st>>>(Dpanel, (int)k, (int)s, lds, comp_l1_ratio,
                                                                        nullptr, T(1), comp_norm);
        MODL_LAUNCH_CHECK(ctx);
        // grad = B_sub - C . D_sub
        MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_NMAJOR, k, s, k, T(-1), C, k, Dpanel, lds, T(1), Bp, lds, st));
        axpy_kernel<T><<<grid_for(ctx, ceil_div(k * lds, 256), 8), 256, 0, st>>>(Dpanel, Bp, k * lds,
                                                                                 (T)(w * step_size));
        MODL_LAUNCH_CHECK(ctx);
        enet_projection_rows_kernel<T><<<grid_for(ctx, k, 16), 256, 0, st>>>(Dpanel, Dpanel, (int)k, (int)s, lds,
                                                                              comp_norm, comp_l1_ratio);
        MODL_LAUNCH_CHECK(ctx);
        enet_norm_rows_kernel<T><<<grid_for(ctx, k, 16), 256, 0, st>>>(Dpanel, (int)k, (int)s, lds, comp_l1_ratio,
                                                                        nullptr, T(-1), comp_norm);
        MODL_LAUNCH_CHECK(ctx);
    }
    prof_mark(ctx, st, MODL_PROF_DICT_POST);
    MODL_TRY(scatter_cols<T>(ctx, Dpanel, lds, k, subset, s, components, ldd, st));   // [ref: :709]
    if (G_full) {                                                                       // [ref: :711-715]
        if (small_subset) {
            if (s > 0)
                MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, k, k, s, T(1), Dpanel, lds, Dpanel, lds, T(1), G_full, k, st));
        } else {
            MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, k, k, p, T(1), components, ldd, components, ldd, T(0), G_full, k, st));
        }
    }
    return MODL_OK;
}

// ---------------------------------------------------------------------------------------
// typed implementations behind the C entry points
// ---------------------------------------------------------------------------------------
template <typename T>
static int enet_norm_impl(modl_ctx *ctx, const T *v, int64_t rows, int64_t n, int64_t ld, T l1_ratio, T *out, void *stream)
{
    MODL_REQUIRE(ctx && v && out && rows >= 0 && n >= 0, "enet_norm arguments");
    if (rows == 0) return MODL_OK;
    enet_norm_rows_kernel<T><<<grid_for(ctx, rows, 16), 256, 0, (cudaStream_t)stream>>>(v, (int)rows, (int)n, ld, l1_ratio,
                                                                                         out, T(0), nullptr);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int enet_projection_impl(modl_ctx *ctx, const T *v, T *out, int64_t rows, int64_t n, int64_t ld, const T *radius,
                                T l1_ratio, void *stream)
{
    MODL_REQUIRE(ctx && v && out && radius && rows >= 0 && n >= 0, "enet_projection arguments");
    if (rows == 0) return MODL_OK;
    enet_projection_rows_kernel<T><<<grid_for(ctx, rows, 16), 256, 0, (cudaStream_t)stream>>>(v, out, (int)rows, (int)n, ld,
                                                                                               radius, l1_ratio);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

template <typename T>
static int enet_scale_impl(modl_ctx *ctx, T *X, int64_t rows, int64_t n, int64_t ld, T l1_ratio, T radius, void *stream)
{
    MODL_REQUIRE(ctx && X && rows >= 0 && n >= 0, "enet_scale arguments");
    if (rows == 0) return MODL_OK;
    enet_scale_rows_kernel<T><<<grid_for(ctx, rows, 16), 256, 0, (cudaStream_t)stream>>>(X, (int)rows, (int)n, ld, l1_ratio, radius);
    MODL_LAUNCH_CHECK(ctx);
    return MODL_OK;
}

// [G ; Dx] = scale [D' ; X'] . D'^T on the tensor cores (float only), D' = D[:, subset] (or D when
// subset == NULL), contraction length kd.  The pack kernels fuse the column gather, the hi/lo split
// and (for X) the full-row squared norm; plain_D (optional) receives the plain k x kd panel the
// dictionary update works on.
static int gram_dx_tc(modl_ctx *ctx, const float *D, int64_t ldd, const float *X, int64_t ldx, const int64_t *subset,
                      int64_t kd, int64_t k, int64_t b, int64_t p, float scale, float *G, float *Dx, float *xnorm2,
                      float *plain_D, int64_t lds, cudaStream_t st, float *plain_X = nullptr, int x_mode = 0)
{
    // x_mode 0: everything; 1: only the X rows of the packed panel (+ plain X panel, row norms) -- the part of
    // the step that does not depend on the dictionary, which MODL_PHASE_PREFETCH runs ahead of time on another
    // stream; 2: the X rows are already there (the rest of the work).
    const bool want_dx = Dx != nullptr && b > 0;
    const int64_t rows = k + (want_dx ? b : 0);
    float *packed = nullptr;
    MODL_TRY(ws<float>(ctx, WS_TC_A, tc_packed_elems(rows, kd), &packed));
    prof_mark(ctx, st, MODL_PROF_GATHER);
    if (x_mode != 1)
        MODL_TRY(tc_pack_rows(ctx, D, ldd, k, p, subset, kd, packed, 0, want_dx ? k : tc_rows_padded(k), plain_D, lds, nullptr, st));
    if (x_mode != 2) {
        if (want_dx)
            MODL_TRY(tc_pack_rows(ctx, X, ldx, b, p, subset, kd, packed, k, tc_rows_padded(k + b), plain_X, lds, xnorm2, st));
        else if (xnorm2 && b > 0)
            MODL_TRY(gather_cols<float>(ctx, X, ldx, b, p, nullptr, 0, (float *)nullptr, 0, xnorm2, st));
    }
    if (x_mode == 1) return MODL_OK;
    prof_mark(ctx, st, MODL_PROF_GRAM);
    if (G != nullptr && (!want_dx || Dx == G + k * k)) {
        MODL_TRY(tc_gemm(ctx, packed, packed, rows, k, kd, scale, 0.f, G, k, 128, st)); // G and Dx are one (k + b) x k matrix
    } else {
        float *gdx = nullptr;
        MODL_TRY(ws<float>(ctx, WS_GDX, (size_t)(rows * k), &gdx));
        MODL_TRY(tc_gemm(ctx, packed, packed, rows, k, kd, scale, 0.f, gdx, k, 128, st));
        if (G) MODL_CUDA_TRY(cudaMemcpyAsync(G, gdx, sizeof(float) * (size_t)(k * k), cudaMemcpyDeviceToDevice, st));
        if (want_dx)
            MODL_CUDA_TRY(cudaMemcpyAsync(Dx, gdx + k * k, sizeof(float) * (size_t)(b * k), cudaMemcpyDeviceToDevice, st));
    }
    return MODL_OK;
}

template <typename T>
static inline bool use_tc(const modl_ctx *ctx) { return std::is_same<T, float>::value && ctx->opt_tc_gemm != 0; }

// G / Dx / xnorm2 products; panel (optional out): where D_sub (k x s) was left
template <typename T>
static int gram_dx_impl(modl_ctx *ctx, const T *D, int64_t ldd, const T *X, int64_t ldx, const int64_t *subset, int64_t s,
                        int64_t k, int64_t b, int64_t p, T scale, T *G, T *Dx, T *xnorm2, T **panel_out, cudaStream_t st,
                        int x_mode = 0)
{
    MODL_REQUIRE(ctx && D && k >= 1 && p >= 1 && b >= 0, "gram_dx arguments");
    MODL_REQUIRE(x_mode == 0 || subset != nullptr, "the split X / D gather needs a feature subset");
    MODL_REQUIRE(X != nullptr || (Dx == nullptr && xnorm2 == nullptr), "X required for Dx / xnorm2");
    if (subset == nullptr) {
        if constexpr (std::is_same<T, float>::value) {
            if (use_tc<T>(ctx) && (G || (Dx && b > 0))) {
                if (panel_out) *panel_out = nullptr;
                return gram_dx_tc(ctx, D, ldd, X, ldx, nullptr, p, k, b, p, scale, G, Dx, xnorm2, nullptr, 0, st);
            }
        }
        prof_mark(ctx, st, MODL_PROF_GATHER);
        if (xnorm2 && b > 0) MODL_TRY(gather_cols<T>(ctx, X, ldx, b, p, nullptr, 0, (T *)nullptr, 0, xnorm2, st));
        prof_mark(ctx, st, MODL_PROF_GRAM);
        if (G) MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, k, k, p, scale, D, ldd, D, ldd, T(0), G, k, st));
        if (Dx && b > 0) MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, b, k, p, scale, X, ldx, D, ldd, T(0), Dx, k, st));
        if (panel_out) *panel_out = nullptr;
        return MODL_OK;
    }
    MODL_REQUIRE(s >= 0 && s <= p, "subset length");
    const int64_t lds = panel_ld(s);
    T *panel = nullptr;
    // D_sub and X_sub panels live in separate slots: the prefetch of step t+1 rewrites X_sub (whose row pitch changes
    // with the subset length) while the dictionary update of step t still works on D_sub
    T *Xsub = nullptr;
    MODL_TRY(ws<T>(ctx, WS_PANEL_DX, (size_t)(k * lds), &panel));
    MODL_TRY(ws<T>(ctx, WS_PANEL_X, (size_t)((b > 0 ? b : 1) * lds), &Xsub));
    T *Dsub = panel;
    if constexpr (std::is_same<T, float>::value) {
        if (use_tc<T>(ctx) && s > 0 && (G || (Dx && b > 0))) {
            if (panel_out) *panel_out = Dsub;
            return gram_dx_tc(ctx, D, ldd, X, ldx, subset, s, k, b, p, scale, G, Dx, xnorm2, Dsub, lds, st, Xsub, x_mode);
        }
    }
    prof_mark(ctx, st, MODL_PROF_GATHER);
    if (x_mode != 1) MODL_TRY(gather_cols<T>(ctx, D, ldd, k, p, subset, s, Dsub, lds, nullptr, st));
    if (x_mode != 2 && b > 0 && (Dx || xnorm2))
        MODL_TRY(gather_cols<T>(ctx, X, ldx, b, p, subset, s, Dx ? Xsub : (T *)nullptr, lds, xnorm2, st));
    if (x_mode == 1) return MODL_OK;
    prof_mark(ctx, st, MODL_PROF_GRAM);
    if (G) MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, k, k, s, scale, Dsub, lds, Dsub, lds, T(0), G, k, st));
    if (Dx && b > 0) MODL_TRY(gemm_simt<T>(ctx, A_KMAJOR, B_KMAJOR, b, k, s, scale, Xsub, lds, Dsub, lds, T(0), Dx, k, st));
    if (panel_out) *panel_out = Dsub;
    return MODL_OK;
}

template <typename T>
static int regression_entry(modl_ctx *ctx, const T *G, int64_t g_stride, T *Dx, const T *X, int64_t ldx, int64_t p,
                            const T *xnorm2, T *code, const int64_t *indices, int64_t b, int64_t k, T l1_ratio, T alpha,
                            int positive, T tol, int max_iter, int32_t *sweeps, void *stream)
{
    MODL_REQUIRE(ctx && G && Dx && code && b >= 0 && k >= 1, "regression arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const T *norms = xnorm2;
    if (l1_ratio != T(0) && norms == nullptr && b > 0) {
        MODL_REQUIRE(X != nullptr && p >= 1, "CD needs the data rows X (or xnorm2) for its stop test");
        T *tmp = nullptr;
        MODL_TRY(ws<T>(ctx, WS_XNORM, (size_t)b, &tmp));
        MODL_TRY(gather_cols<T>(ctx, X, ldx, b, p, nullptr, 0, (T *)nullptr, 0, tmp, st));
        norms = tmp;
    }
    return regression<T>(ctx, G, g_stride, Dx, norms, code, indices, nullptr, b, k, l1_ratio, alpha, positive, tol,
                         max_iter, sweeps, st);
}

template <typename T>
static int update_stats_impl(modl_ctx *ctx, const T *code, const int64_t *indices, const T *X, int64_t ldx, T *C, T *B,
                             int64_t ldb, double w, int64_t b, int64_t k, int64_t p, int overwrite, cudaStream_t st,
                             int64_t global_batch = 0, bool increments_only = false)
{
    MODL_REQUIRE(ctx && code && b >= 1 && k >= 1, "update_stats arguments");
    const T *cb = code;
    if (indices) {
        T *tmp = nullptr;
        MODL_TRY(ws<T>(ctx, WS_CODE_BATCH, (size_t)(b * k), &tmp));
        gather_rows_kernel<T><<<grid_for(ctx, b, 16), 128, 0, st>>>(code, k, indices, (int)b, (int)k, tmp, k);
        MODL_LAUNCH_CHECK(ctx);
        cb = tmp;
    }
    const double bt = (double)(global_batch > 0 ? global_batch : b);
    const T a = overwrite ? (T)(1.0 / bt) : (T)(w / bt);
    const T be = (overwrite || increments_only) ? T(0) : (T)(1.0 - w);
    MODL_REQUIRE(B == nullptr || (X != nullptr && p >= 1), "X required for the B_ update");
    if constexpr (std::is_same<T, float>::value) {
        if (use_tc<T>(ctx)) {
            // contraction over the batch: code^T and X^T as packed split panels (transposing packs)
            float *codeP = nullptr, *XP = nullptr;
            MODL_TRY(ws<float>(ctx, WS_TC_CODE, tc_packed_elems(k, b), &codeP));
            MODL_TRY(tc_pack_cols(ctx, cb, k, b, k, codeP, 128, st));
            if (C && B) {
                // ONE launch for both statistics: [B_ | C_] = code^T [X | code].  The code rows are appended to the
                // packed X^T operand at the next tile boundary; tiles beyond it are written to C_.
                const int bn = tc_pick_bn(ctx, p + k, ceil_div(k, 128));
                const int64_t n_split = round_up(p, bn), nkb = ceil_div(b, 32);
                MODL_TRY(ws<float>(ctx, WS_TC_X, tc_packed_elems(n_split + k, b, bn), &XP));
                MODL_TRY(tc_pack_cols(ctx, X, ldx, b, p, XP, bn, st));
                MODL_TRY(tc_pack_cols(ctx, cb, k, b, k, XP + (size_t)(n_split / bn) * nkb * (2 * bn * 32), bn, st));
                return tc_gemm(ctx, codeP, XP, k, n_split + k, b, a, be, B, ldb, bn, st, WS_GEMM_PART, C, k, n_split, p);
            }
            if (C) MODL_TRY(tc_gemm(ctx, codeP, codeP, k, k, b, a, be, C, k, 128, st));
            if (B) {
                const int bn = tc_pick_bn(ctx, p, ceil_div(k, 128));       // e.g. p = 10000, k = 256: 144 -> 140 CTAs, one wave
                MODL_TRY(ws<float>(ctx, WS_TC_X, tc_packed_elems(p, b, bn), &XP));
                MODL_TRY(tc_pack_cols(ctx, X, ldx, b, p, XP, bn, st));
                MODL_TRY(tc_gemm(ctx, codeP, XP, k, p, b, a, be, B, ldb, bn, st));
            }
            return MODL_OK;
        }
    }
    if (C) MODL_TRY(gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, k, b, a, cb, k, cb, k, be, C, k, st));
    if (B) MODL_TRY(gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, p, b, a, cb, k, X, ldx, be, B, ldb, st));
    return MODL_OK;
}

// Packed X[:, subset]^T (the B operand of the subset statistics product) -- depends on the batch and the subset
// only, so MODL_PHASE_PREFETCH builds it ahead of time.  The packed code^T is appended at the next tile boundary.
static inline int64_t xs_split(int64_t s) { return s > 0 ? round_up(s, 128) : 0; }
static int pack_xsub(modl_ctx *ctx, const float *Xsub, int64_t lds, int64_t s, int64_t b, int64_t k, float **XsP, cudaStream_t st,
                     bool do_pack, int slot)
{
    MODL_TRY(ws<float>(ctx, slot ? WS_TC_XS2 : WS_TC_XS, tc_packed_elems(xs_split(s) + k, b), XsP));
    if (do_pack && s > 0) MODL_TRY(tc_pack_cols(ctx, Xsub, lds, b, s, *XsP, 128, st));
    return MODL_OK;
}

// What the dictionary update waits for, and nothing else:
//   raw (sharded) form: inc_sub = a [code^T code | code^T X[:, subset]]   (k x k, then k x lds), summed over ranks by the caller;
//   fused (one GPU) form, inc_sub == NULL:  C_ = keep C_ + a code^T code  and  Bp = keep Bp + a code^T X[:, subset], where Bp
//   already holds B_[:, subset] -- ONE tensor-core launch, [Bp | C_] = code^T [X_sub | code] with the decay in the epilogue.
// Xsub is the plain b x lds panel of X[:, subset]; xs_packed says its packed transpose is already in WS_TC_XS.
template <typename T>
static int stats_sub_impl(modl_ctx *ctx, const T *cb, const T *Xsub, int64_t lds, int64_t s, T *inc_sub, T *Cmat, T *Bp, T a,
                          T keep, int64_t b, int64_t k, cudaStream_t st, bool xs_packed, int slot = 0)
{
    T *outC = inc_sub ? inc_sub : Cmat, *outB = inc_sub ? inc_sub + k * k : Bp;
    const T be = inc_sub ? T(0) : keep;
    if constexpr (std::is_same<T, float>::value) {
        if (use_tc<T>(ctx)) {
            float *XsP = nullptr;
            MODL_TRY(pack_xsub(ctx, Xsub, lds, s, b, k, &XsP, st, !xs_packed, slot));
            const int64_t nkb = ceil_div(b, 32);
            float *codeP = XsP + (size_t)(xs_split(s) / 128) * nkb * (2 * 128 * 32);
            MODL_TRY(tc_pack_cols(ctx, cb, k, b, k, codeP, 128, st));
            ctx->code_packed = codeP;
            if (s > 0)
                return tc_gemm(ctx, codeP, XsP, k, xs_split(s) + k, b, a, be, outB, lds, 128, st, WS_GEMM_PART, outC, k, xs_split(s), s);
            return tc_gemm(ctx, codeP, codeP, k, k, b, a, be, outC, k, 128, st);
        }
    }
    MODL_TRY(gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, k, b, a, cb, k, cb, k, be, outC, k, st));
    if (s > 0) MODL_TRY(gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, s, b, a, cb, k, Xsub, lds, be, outB, lds, st));
    return MODL_OK;
}

// The full-width statistic on whatever stream the caller chose: out = be * out + a * code^T X  (k x p).
// sm_avail: SMs the product may count on (the dictionary update's cluster holds the others while it runs).
template <typename T>
static int stats_b_impl(modl_ctx *ctx, const T *cb, const T *X, int64_t ldx, T *out, int64_t ldo, T a, T be, int64_t b,
                        int64_t k, int64_t p, cudaStream_t st, int sm_avail)
{
    if constexpr (std::is_same<T, float>::value) {
        if (use_tc<T>(ctx)) {
            const float *codeP = ctx->code_packed;
            float *XP = nullptr;
            if (!codeP) {
                float *cp = nullptr;
                MODL_TRY(ws<float>(ctx, WS_TC_CODE, tc_packed_elems(k, b), &cp));
                MODL_TRY(tc_pack_cols(ctx, cb, k, b, k, cp, 128, st));
                codeP = cp;
            }
            const int bn = tc_pick_bn(ctx, p, ceil_div(k, 128), sm_avail);
            if (ctx->opt_tc_raw_b) {
                // X straight from its rows: the GEMM's own warps split it into the TF32 hi / lo operand in shared memory
                (void)XP;
                return tc_gemm(ctx, codeP, nullptr, k, p, b, a, be, out, ldo, bn, st, WS_GEMM_PART2, nullptr, 0, 0, 0, X, ldx);
            }
            MODL_TRY(ws<float>(ctx, WS_TC_X, tc_packed_elems(p, b, bn), &XP));
            MODL_TRY(tc_pack_cols(ctx, X, ldx, b, p, XP, bn, st));
            return tc_gemm(ctx, codeP, XP, k, p, b, a, be, out, ldo, bn, st, WS_GEMM_PART2);
        }
    }
    return gemm_simt<T>(ctx, A_MMAJOR, B_NMAJOR, k, p, b, a, cb, k, X, ldx, be, out, ldo, st, WS_GEMM_PART2);
}

template <typename T>
static int update_dict_entry(modl_ctx *ctx, T *components, int64_t ldd, const T *B, int64_t ldb, const T *C, T *comp_norm,
                             T *G_full, const int64_t *subset, int64_t s, const int64_t *h_order, int64_t k, int64_t p,
                             T comp_l1_ratio, int comp_pos, int mode, double w, double step_size, void *stream)
{
    MODL_REQUIRE(ctx && components && B && C && comp_norm && subset && h_order, "update_dict arguments");
    MODL_REQUIRE(s >= 0 && s <= p, "subset length");
    T *panel = nullptr;
    MODL_TRY(ws<T>(ctx, WS_PANEL_DX, (size_t)(k * panel_ld(s)), &panel));
    return update_dict_impl<T>(ctx, components, ldd, B, ldb, C, comp_norm, G_full, subset, s, h_order, k, p, comp_l1_ratio,
                               comp_pos, mode, w, step_size, panel, false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------
// the fused step  [ref: dict_fact.py:507-533, 577-648]
// ---------------------------------------------------------------------------------------
template <typename T>
int batch_fit_impl(modl_ctx *ctx, const modl_step_params *q, void *stream)
{
    MODL_REQUIRE(ctx && q, "null arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t k = q->n_components, b = q->batch_size, p = q->n_features, s = q->subset_len;
    MODL_REQUIRE(k >= 1 && b >= 1 && p >= 1 && q->n_samples >= 1, "shapes");
    MODL_REQUIRE(q->X && q->components && q->code && q->C && q->B && q->comp_norm, "state pointers");
    MODL_REQUIRE(q->h_subset && q->h_order && s >= 0 && s <= p, "subset / order");
    MODL_REQUIRE(q->Dx_agg >= 0 && q->Dx_agg <= 2 && q->G_agg >= 0 && q->G_agg <= 2, "aggregation mode");
    MODL_REQUIRE(q->G_agg != MODL_AGG_FULL || q->G_full, "G_agg='full' needs G_full");
    MODL_REQUIRE(q->G_agg != MODL_AGG_AVERAGE || (q->G_average && q->w_sample), "G_agg='average' needs G_average, w_sample");
    MODL_REQUIRE(q->Dx_agg != MODL_AGG_AVERAGE || (q->Dx_average && q->w_sample), "Dx_agg='average' needs Dx_average, w_sample");
    MODL_REQUIRE(q->slot == 0 || q->slot == 1, "slot");
    ctx->prof_n = 0;
    prof_mark(ctx, st, MODL_PROF_GATHER);
    const T *X = static_cast<const T *>(q->X);
    T *D = static_cast<T *>(q->components);
    T *code = static_cast<T *>(q->code);
    const T r = (T)q->reduction;
    const int flag_bits = MODL_PHASE_REUSE_SUBSET | MODL_PHASE_INPUTS_READY | MODL_PHASE_FUSED_APPLY | MODL_PHASE_GATHER_B;
    const bool gather_b = (q->phases & MODL_PHASE_GATHER_B) != 0;
    const bool inputs_ready = (q->phases & MODL_PHASE_INPUTS_READY) != 0;     // this step's MODL_PHASE_PREFETCH has run
    const bool reuse_subset = (q->phases & MODL_PHASE_REUSE_SUBSET) != 0 || inputs_ready;
    const bool fused_apply = (q->phases & MODL_PHASE_FUSED_APPLY) != 0;
    const int phases = (q->phases & ~flag_bits) ? (q->phases & ~flag_bits)
                                                : (MODL_PHASE_CODE | MODL_PHASE_STATS | MODL_PHASE_APPLY | MODL_PHASE_DICT);
    const int slot = q->slot;
    T *inc = static_cast<T *>(q->stats_inc);
    T *inc_sub = static_cast<T *>(q->inc_sub);
    const double bt = (double)(q->global_batch > 0 ? q->global_batch : b);
    const T av = q->optimizer_sgd ? (T)(1.0 / bt) : (T)(q->w / bt);
    const T keep = q->optimizer_sgd ? T(0) : (T)(1.0 - q->w);
    const int64_t lds = panel_ld(s);
    // which parts of the code phase work on the feature subset [ref: :589-604]
    const bool need_sub = q->Dx_agg != MODL_AGG_FULL || q->G_agg != MODL_AGG_FULL;
    const bool dx_sub = q->Dx_agg != MODL_AGG_FULL;
    const bool g_sub = q->G_agg != MODL_AGG_FULL;
    const bool x_ahead = need_sub && dx_sub && s > 0;     // the X side of the gather can run ahead of the dictionary

    if (phases == MODL_PHASE_APPLY_B) {
        // B_ = (1-w) B_ + all-reduced increments; touches no workspace, so it may run on a side stream
        MODL_REQUIRE(inc != nullptr, "APPLY_B without stats_inc");
        xpby_kernel<T><<<grid_for(ctx, ceil_div(k * p, 256), 8), 256, 0, st>>>(static_cast<T *>(q->B), inc + k * k, k * p, keep);
        MODL_LAUNCH_CHECK(ctx);
        ctx->prof_n = 0;
        return MODL_OK;
    }
    if (phases == MODL_PHASE_STATS_B) {
        // full-width B statistic from the batch code left in the workspace by this step's CODE phase
        MODL_REQUIRE(ctx->slot_ptr[WS_CODE_BATCH] != nullptr, "STATS_B before any CODE phase");
        const T *cbp = static_cast<const T *>(ctx->slot_ptr[WS_CODE_BATCH]);
        ctx->prof_n = 0;
        if (inc) return stats_b_impl<T>(ctx, cbp, X, q->ldx, inc + k * k, p, av, T(0), b, k, p, st, q->sm_avail);
        return stats_b_impl<T>(ctx, cbp, X, q->ldx, static_cast<T *>(q->B), p, av, keep, b, k, p, st, q->sm_avail);
    }
    MODL_REQUIRE(!(phases & (MODL_PHASE_APPLY_B | MODL_PHASE_STATS_B)), "APPLY_B / STATS_B run in calls of their own");
    MODL_REQUIRE(!(phases & MODL_PHASE_STATS_SUB) || ((phases & MODL_PHASE_CODE) && (inc_sub || fused_apply) && !(phases & MODL_PHASE_STATS)),
                 "STATS_SUB needs CODE in the same call, inc_sub (or FUSED_APPLY), and excludes STATS");
    MODL_REQUIRE(!(phases & MODL_PHASE_APPLY_SUB) || inc_sub, "APPLY_SUB without inc_sub");
    MODL_REQUIRE(!(phases & MODL_PHASE_STATS) || (phases & MODL_PHASE_CODE), "STATS phase needs CODE in the same call");
    MODL_REQUIRE(inc != nullptr || q->phases == 0 || !(phases & MODL_PHASE_APPLY) || (phases & MODL_PHASE_STATS),
                 "APPLY without stats_inc");
    MODL_REQUIRE(!(phases & MODL_PHASE_PREFETCH) || phases == MODL_PHASE_PREFETCH, "PREFETCH runs in a call of its own");

    // subset -> device
    int64_t *d_subset = nullptr;
    MODL_TRY(ws<int64_t>(ctx, slot ? WS_SUBSET2 : WS_SUBSET, (size_t)(s > 0 ? s : 1), &d_subset));
    const bool fetch_mapped = q->h_inputs_mapped != 0 && phases == MODL_PHASE_PREFETCH;
    if (s > 0 && !reuse_subset && !fetch_mapped)
        MODL_CUDA_TRY(cudaMemcpyAsync(d_subset, q->h_subset, sizeof(int64_t) * (size_t)s, cudaMemcpyHostToDevice, st));

    T *xnorm2 = nullptr, *Gw = nullptr, *Dxw = nullptr, *cb = nullptr, *panel = nullptr;
    MODL_TRY(ws<T>(ctx, WS_XNORM, (size_t)b, &xnorm2));
    MODL_TRY(ws<T>(ctx, WS_G, (size_t)(k * k + b * k), &Gw));   // [G ; Dx]: one (k + b) x k matrix
    Dxw = Gw + k * k;
    MODL_TRY(ws<T>(ctx, WS_CODE_BATCH, (size_t)(b * k), &cb));

    if (phases == MODL_PHASE_PREFETCH) {
        // Everything of the step that depends on the batch, the subset and the atom order but NOT on the
        // dictionary or on this step's codes: index uploads, the X side of the subset gather (packed rows,
        // row norms), the packed X[:, subset]^T operand of the subset statistics and, on one GPU, the
        // B_[:, subset] panel the fused statistics product accumulates into.  Runs on another stream while
        // the previous step's dictionary update holds its 16 SMs.
        int32_t *d_order = nullptr;
        if (fetch_mapped) {
            for (int64_t i = 0; i < k; ++i)
                MODL_REQUIRE(q->h_order[i] >= 0 && q->h_order[i] < k, "order is not a permutation of range(k)");
            MODL_TRY(ws<int32_t>(ctx, slot ? WS_ORDER2 : WS_ORDER, (size_t)k, &d_order));
            const int64_t n = s > k ? s : k;
            fetch_step_inputs_kernel<<<(unsigned)(n < 2048 ? 1 : (n + 2047) / 2048), 256, 0, st>>>(q->h_subset, d_subset, (int)s,
                                                                                                    q->h_order, d_order, (int)k);
            MODL_LAUNCH_CHECK(ctx);
        } else {
            MODL_TRY(upload_order(ctx, q->h_order, k, &d_order, st, slot));
        }
        if (x_ahead) {
            MODL_TRY(gram_dx_impl<T>(ctx, D, p, X, q->ldx, d_subset, s, k, b, p, r, g_sub ? Gw : (T *)nullptr, Dxw, xnorm2, &panel, st, 1));
            if constexpr (std::is_same<T, float>::value) {
                if (use_tc<T>(ctx)) {
                    float *XsP = nullptr;
                    float *Xs = nullptr;
                    MODL_TRY(ws<float>(ctx, WS_PANEL_X, (size_t)(b * lds), &Xs));
                    MODL_TRY(pack_xsub(ctx, Xs, lds, s, b, k, &XsP, st, true, slot));
                }
            }
        }
        if (fused_apply && s > 0) {
            T *Bp = nullptr;
            MODL_TRY(ws<T>(ctx, slot ? WS_PANEL_B2 : WS_PANEL_B, (size_t)(k * lds), &Bp));
            MODL_TRY(gather_cols<T>(ctx, static_cast<const T *>(q->B), p, k, p, d_subset, s, Bp, lds, nullptr, st));
        }
        ctx->prof_n = 0;
        return MODL_OK;
    }

    T *panel_keep = nullptr;
    bool b_early = false;
    if (phases & MODL_PHASE_CODE) {
    if (gather_b && fused_apply && (phases & MODL_PHASE_STATS_SUB) && ctx->fork_stream && s > 0) {
        // B_[:, subset] on the forked stream, beside the Gram / code solve (B_ is final: the previous step has ended)
        for (int i = 0; i < 2; ++i)
            if (!ctx->ev_fork[i]) MODL_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork[i], cudaEventDisableTiming));
        T *Bp = nullptr;
        MODL_TRY(ws<T>(ctx, slot ? WS_PANEL_B2 : WS_PANEL_B, (size_t)(k * lds), &Bp));
        MODL_CUDA_TRY(cudaEventRecord(ctx->ev_fork[0], st));
        MODL_CUDA_TRY(cudaStreamWaitEvent(ctx->fork_stream, ctx->ev_fork[0], 0));
        MODL_TRY(gather_cols<T>(ctx, static_cast<const T *>(q->B), p, k, p, d_subset, s, Bp, lds, nullptr, ctx->fork_stream));
        MODL_CUDA_TRY(cudaEventRecord(ctx->ev_fork[1], ctx->fork_stream));
        b_early = true;
    }
    // ---- _compute_code [ref: :577-648] ----
    const int x_mode = (inputs_ready && x_ahead) ? 2 : 0;
    if (need_sub) {
        MODL_TRY(gram_dx_impl<T>(ctx, D, p, X, q->ldx, d_subset, s, k, b, p, r, g_sub ? Gw : (T *)nullptr,
                                 dx_sub ? Dxw : (T *)nullptr, xnorm2, &panel, st, x_mode));
    }
    if (!dx_sub) {   // Dx = X . D^T over all features [ref: :592]
        MODL_TRY(gram_dx_impl<T>(ctx, D, p, X, q->ldx, nullptr, 0, k, b, p, T(1), (T *)nullptr, Dxw,
                                 need_sub ? (T *)nullptr : xnorm2, nullptr, st));
    }
    prof_mark(ctx, st, MODL_PROF_AVERAGE);
    if (q->Dx_agg == MODL_AGG_AVERAGE) {
        update_dx_average_kernel<T><<<grid_for(ctx, b, 16), 128, 0, st>>>(static_cast<T *>(q->Dx_average), Dxw,
                                                                          static_cast<const T *>(q->w_sample), q->indices, (int)b, (int)k);
        MODL_LAUNCH_CHECK(ctx);
    }
    const T *Guse = Gw;
    int64_t g_stride = 0;
    T *g_rows = nullptr;
    if (q->G_agg == MODL_AGG_FULL) {
        Guse = static_cast<const T *>(q->G_full);
    } else if (q->G_agg == MODL_AGG_AVERAGE) {
        T *Gav = static_cast<T *>(q->G_average);
        dim3 grid((unsigned)grid_for(ctx, ceil_div(k * k, 256), 2), (unsigned)(b < 65535 ? b : 65535));
        update_g_average_kernel<T><<<grid, 256, 0, st>>>(Gav, Gw, static_cast<const T *>(q->w_sample), q->indices, (int)b, k * k);
        MODL_LAUNCH_CHECK(ctx);
        if (q->indices) {   // the solver wants the batch's matrices contiguous: gather the rows
            MODL_TRY(ws<T>(ctx, WS_GROWS, (size_t)(b * k * k), &g_rows));
            gather_rows_kernel<T><<<grid_for(ctx, b, 16), 256, 0, st>>>(Gav, k * k, q->indices, (int)b, (int)(k * k), g_rows, k * k);
            MODL_LAUNCH_CHECK(ctx);
            Guse = g_rows;
        } else {
            Guse = Gav;
        }
        g_stride = k * k;
    }
    prof_mark(ctx, st, MODL_PROF_CODE);
    MODL_TRY(regression<T>(ctx, Guse, g_stride, Dxw, xnorm2, code, q->indices, cb, b, k, (T)q->code_l1_ratio,
                           (T)q->code_alpha, q->code_pos, (T)q->tol, q->max_iter, q->sweeps, st));

    panel_keep = panel;
    ctx->code_packed = nullptr;
    if (phases & MODL_PHASE_STATS_SUB) {
        prof_mark(ctx, st, MODL_PROF_STATS);
        T *Xsub = nullptr;
        MODL_TRY(ws<T>(ctx, WS_PANEL_X, (size_t)(b * lds), &Xsub));
        const bool x_there = need_sub && dx_sub && s > 0;   // X[:, subset] was gathered by the code phase (or its prefetch)
        if (!x_there)                                       // Dx_agg == 'full'
            MODL_TRY(gather_cols<T>(ctx, X, q->ldx, b, p, d_subset, s, Xsub, lds, nullptr, st));
        T *Bp = nullptr;
        if (fused_apply) {
            MODL_TRY(ws<T>(ctx, slot ? WS_PANEL_B2 : WS_PANEL_B, (size_t)(k * lds), &Bp));
            if (b_early)
                MODL_CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_fork[1], 0));
            else if ((!inputs_ready || gather_b) && s > 0)
                MODL_TRY(gather_cols<T>(ctx, static_cast<const T *>(q->B), p, k, p, d_subset, s, Bp, lds, nullptr, st));
        }
        MODL_TRY(stats_sub_impl<T>(ctx, cb, Xsub, lds, s, fused_apply ? (T *)nullptr : inc_sub, static_cast<T *>(q->C), Bp, av,
                                   keep, b, k, st, inputs_ready && x_there, slot));
        if (fused_apply) {
            ctx->panel_b_ready = 1;
            if (q->ev_after_apply_sub) MODL_TRY(record_step_event(ctx, q->ev_after_apply_sub, st));
        }
    }
    // ---- _update_C / _update_B [ref: :559-575] ----
    if (phases & MODL_PHASE_STATS) {
        prof_mark(ctx, st, MODL_PROF_STATS);
        if (inc)
        {
            MODL_TRY(update_stats_impl<T>(ctx, cb, nullptr, X, q->ldx, inc, inc + k * k, p, q->w, b, k, p, q->optimizer_sgd,
                                          st, q->global_batch, true));
            if (inc_sub) {   // compact copy of what the dictionary update needs first: [C inc | B inc[:, subset]]
                MODL_CUDA_TRY(cudaMemcpyAsync(inc_sub, inc, sizeof(T) * (size_t)(k * k), cudaMemcpyDeviceToDevice, st));
                MODL_TRY(gather_cols<T>(ctx, inc + k * k, p, k, p, d_subset, s, inc_sub + k * k, lds, nullptr, st));
            }
        }
        else
            MODL_TRY(update_stats_impl<T>(ctx, cb, nullptr, X, q->ldx, static_cast<T *>(q->C), static_cast<T *>(q->B), p,
                                          q->w, b, k, p, q->optimizer_sgd, st, q->global_batch, false));
    }
    }   // MODL_PHASE_CODE
    if ((phases & MODL_PHASE_APPLY) && inc) {
        prof_mark(ctx, st, MODL_PROF_STATS);
        xpby_kernel<T><<<grid_for(ctx, ceil_div(k * k, 256), 4), 256, 0, st>>>(static_cast<T *>(q->C), inc, k * k, keep);
        MODL_LAUNCH_CHECK(ctx);
        xpby_kernel<T><<<grid_for(ctx, ceil_div(k * p, 256), 8), 256, 0, st>>>(static_cast<T *>(q->B), inc + k * k, k * p, keep);
        MODL_LAUNCH_CHECK(ctx);
    }
    if (phases & MODL_PHASE_APPLY_SUB) {
        prof_mark(ctx, st, MODL_PROF_STATS);
        T *Bp = nullptr;
        MODL_TRY(ws<T>(ctx, slot ? WS_PANEL_B2 : WS_PANEL_B, (size_t)(k * lds), &Bp));
        // C_ and the B_[:, subset] panel in one launch (s == 0: the second loop is empty)
        apply_sub_kernel<T><<<grid_for(ctx, k, 2), 256, 0, st>>>(static_cast<T *>(q->C), inc_sub, (int)k, static_cast<const T *>(q->B), p,
                                                                 d_subset, (int)s, keep, inc_sub + k * k, lds, Bp, lds);
        MODL_LAUNCH_CHECK(ctx);
        ctx->panel_b_ready = 1;
        if (q->ev_after_apply_sub) MODL_TRY(record_step_event(ctx, q->ev_after_apply_sub, st));
    }
    if (phases & MODL_PHASE_DICT) {
    // ---- _update_dict [ref: :650-715] ----
    T *Dpanel = panel_keep;
    bool ready = panel_keep != nullptr;
    if (!ready) MODL_TRY(ws<T>(ctx, WS_PANEL_DX, (size_t)(k * lds), &Dpanel));
    MODL_TRY(update_dict_impl<T>(ctx, D, p, static_cast<const T *>(q->B), p, static_cast<const T *>(q->C),
                                 static_cast<T *>(q->comp_norm), q->G_agg == MODL_AGG_FULL ? static_cast<T *>(q->G_full) : (T *)nullptr,
                                 d_subset, s, q->h_order, k, p, (T)q->comp_l1_ratio, q->comp_pos, q->optimizer_sgd ? 1 : 0, q->w,
                                 q->step_size, Dpanel, ready, st, ctx->panel_b_ready != 0, slot, inputs_ready,
                                 static_cast<unsigned *>(q->start_flag), q->start_serial));
    ctx->panel_b_ready = 0;
    }   // MODL_PHASE_DICT
    if (ctx->prof_on && ctx->prof_n > 0) {
        cudaEventRecord(ctx->prof_ev[ctx->prof_n], st);
        MODL_CUDA_TRY(cudaEventSynchronize(ctx->prof_ev[ctx->prof_n]));
        for (int i = 0; i < ctx->prof_n; ++i) {
            float ms = 0;
            MODL_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]));
            ctx->prof_ms[ctx->prof_phase[i]] += ms;
        }
        ctx->prof_steps += 1;
    }
    ctx->prof_n = 0;
    return MODL_OK;
}

template int batch_fit_impl<float>(modl_ctx *, const modl_step_params *, void *);
template int batch_fit_impl<double>(modl_ctx *, const modl_step_params *, void *);

}  // namespace modl

// ---------------------------------------------------------------------------------------
// extern "C" shims
// ---------------------------------------------------------------------------------------
extern "C" {

#define MODL_DEFINE_TYPED(SFX, T)                                                                                           \
    int modl_enet_norm_##SFX(modl_ctx *c, const T *v, int64_t rows, int64_t n, int64_t ld, T l1, T *out, void *st)          \
    { CtxGuard g_(c); return enet_norm_impl<T>(c, v, rows, n, ld, l1, out, st); }                                                           \
    int modl_enet_projection_##SFX(modl_ctx *c, const T *v, T *out, int64_t rows, int64_t n, int64_t ld, const T *radius,   \
                                   T l1, void *st)                                                                          \
    { CtxGuard g_(c); return enet_projection_impl<T>(c, v, out, rows, n, ld, radius, l1, st); }                                             \
    int modl_enet_scale_##SFX(modl_ctx *c, T *X, int64_t rows, int64_t n, int64_t ld, T l1, T radius, void *st)             \
    { CtxGuard g_(c); return enet_scale_impl<T>(c, X, rows, n, ld, l1, radius, st); }                                                       \
    int modl_gram_dx_##SFX(modl_ctx *c, const T *D, int64_t ldd, const T *X, int64_t ldx, const int64_t *subset, int64_t s, \
                           int64_t k, int64_t b, int64_t p, T scale, T *G, T *Dx, T *xnorm2, void *st)                      \
    { CtxGuard g_(c); return gram_dx_impl<T>(c, D, ldd, X, ldx, subset, s, k, b, p, scale, G, Dx, xnorm2, nullptr, (cudaStream_t)st); }     \
    int modl_enet_regression_single_gram_##SFX(modl_ctx *c, const T *G, T *Dx, const T *X, int64_t ldx, int64_t p,          \
                                               const T *xnorm2, T *code, const int64_t *indices, int64_t b, int64_t k,      \
                                               T l1, T alpha, int positive, T tol, int max_iter, int32_t *sweeps, void *st) \
    { CtxGuard g_(c); return regression_entry<T>(c, G, 0, Dx, X, ldx, p, xnorm2, code, indices, b, k, l1, alpha, positive, tol, max_iter,   \
                                 sweeps, st); }                                                                             \
    int modl_enet_regression_multi_gram_##SFX(modl_ctx *c, const T *G, T *Dx, const T *X, int64_t ldx, int64_t p,           \
                                              const T *xnorm2, T *code, const int64_t *indices, int64_t b, int64_t k,       \
                                              T l1, T alpha, int positive, T tol, int max_iter, int32_t *sweeps, void *st)  \
    { CtxGuard g_(c); return regression_entry<T>(c, G, k * k, Dx, X, ldx, p, xnorm2, code, indices, b, k, l1, alpha, positive, tol,         \
                                 max_iter, sweeps, st); }                                                                   \
    int modl_update_G_average_##SFX(modl_ctx *c, T *Gav, const T *G, const T *w, const int64_t *indices, int64_t b,         \
                                    int64_t k, void *st)                                                                    \
    {                                                                                                                       \
        MODL_REQUIRE(c && Gav && G && w && b >= 0 && k >= 1, "update_G_average arguments");                                 \
        CtxGuard g_(c);                                                                                                     \
        if (b == 0) return MODL_OK;                                                                                         \
        dim3 grid((unsigned)grid_for(c, ceil_div(k * k, 256), 2), (unsigned)(b < 65535 ? b : 65535));                       \
        update_g_average_kernel<T><<<grid, 256, 0, (cudaStream_t)st>>>(Gav, G, w, indices, (int)b, k * k);                  \
        MODL_LAUNCH_CHECK(c);                                                                                               \
        return MODL_OK;                                                                                                     \
    }                                                                                                                       \
    int modl_update_Dx_average_##SFX(modl_ctx *c, T *Dav, T *Dx, const T *w, const int64_t *indices, int64_t b, int64_t k,  \
                                     void *st)                                                                              \
    {                                                                                                                       \
        MODL_REQUIRE(c && Dav && Dx && w && b >= 0 && k >= 1, "update_Dx_average arguments");                               \
        CtxGuard g_(c);                                                                                                     \
        if (b == 0) return MODL_OK;                                                                                         \
        update_dx_average_kernel<T><<<grid_for(c, b, 16), 128, 0, (cudaStream_t)st>>>(Dav, Dx, w, indices, (int)b, (int)k); \
        MODL_LAUNCH_CHECK(c);                                                                                               \
        return MODL_OK;                                                                                                     \
    }                                                                                                                       \
    int modl_update_stats_##SFX(modl_ctx *c, const T *code, const int64_t *indices, const T *X, int64_t ldx, T *C, T *B,    \
                                int64_t ldb, double w, int64_t b, int64_t k, int64_t p, int overwrite, void *st)            \
    { CtxGuard g_(c); return update_stats_impl<T>(c, code, indices, X, ldx, C, B, ldb, w, b, k, p, overwrite, (cudaStream_t)st); }          \
    int modl_update_dict_##SFX(modl_ctx *c, T *comp, int64_t ldd, const T *B, int64_t ldb, const T *C, T *comp_norm,        \
                               T *G_full, const int64_t *subset, int64_t s, const int64_t *h_order, int64_t k, int64_t p,   \
                               T l1, int pos, int mode, double w, double step, void *st)                                    \
    { CtxGuard g_(c); return update_dict_entry<T>(c, comp, ldd, B, ldb, C, comp_norm, G_full, subset, s, h_order, k, p, l1, pos, mode, w,   \
                                  step, st); }                                                                              \
    int modl_batch_fit_##SFX(modl_ctx *c, const modl_step_params *prm, void *st)                                            \
    { CtxGuard g_(c); return batch_fit_impl<T>(c, prm, st); }

MODL_DEFINE_TYPED(f32, float)
MODL_DEFINE_TYPED(f64, double)

}  // extern "C"
