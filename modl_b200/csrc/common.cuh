// common.cuh -- shared plumbing of the sm_100a kernels: context, workspace arena, error
// reporting, warp/block reduction helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/modl_b200.h"

namespace modl {

void set_error(const char *fmt, ...);

#define MODL_CUDA_TRY(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::modl::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                              cudaGetErrorString(_e));                                   \
            return MODL_ECUDA;                                                           \
        }                                                                                \
    } while (0)

#define MODL_TRY(expr)                  \
    do {                                \
        int _s = (expr);                \
        if (_s != MODL_OK) return _s;   \
    } while (0)

#define MODL_REQUIRE(cond, msg)                                                 \
    do {                                                                        \
        if (!(cond)) {                                                          \
            ::modl::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return MODL_EINVAL;                                                 \
        }                                                                       \
    } while (0)

// internal status (never crosses the C ABI): a workspace slot must grow while the call is being stream-captured
#define MODL_EGROW 100

// check the launch that was just issued
#define MODL_LAUNCH_CHECK(ctx)                \
    do {                                      \
        (ctx)->launches += 1;                 \
        MODL_CUDA_TRY(cudaGetLastError());    \
    } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Named, grow-only device scratch slots.  A slot keeps its allocation across calls, so the
// steady-state minibatch loop performs no cudaMalloc.  All work is stream-ordered on the
// caller's stream; slots are reused call after call on that same stream.
enum WsSlot {
    WS_SUBSET = 0,   // int64[s]
    WS_ORDER,        // int32[k]
    WS_PANEL_DX,     // D_sub  k x s_pad
    WS_PANEL_B,      // B_sub / gradient panel  k x s_pad
    WS_XNORM,        // real[b]
    WS_G,            // k x k
    WS_DX,           // b x k
    WS_CODE_BATCH,   // b x k
    WS_GEMM_PART,    // split-K partials
    WS_BCD_SYNC,     // barrier counter + partial sums + v rows
    WS_CHOL,         // k x k factor (x b for per-sample Grams)
    WS_GROWS,        // gathered per-sample Gram matrices
    WS_GPACK,        // tile-packed lower triangle of the shared Gram
    WS_TC_A,         // packed split panel [D_sub ; X_sub]
    WS_TC_CODE,      // packed split panel code^T
    WS_TC_X,         // packed split panel X^T
    WS_GDX,          // [G ; Dx] of the tensor-core path, (k + b) x k
    WS_TC_XS,        // packed split panel X[:, subset]^T
    WS_GEMM_PART2,   // split-K partials of the side-stream statistics product
    WS_MISC,         // small scalars
    WS_INFO,         // int status flags
    WS_SUBSET2,      // second copies of the per-step inputs: the step in flight keeps its own while the next
    WS_ORDER2,       //   step's are prefetched on another stream (modl_step_params.slot)
    WS_PANEL_B2,
    WS_INDICES,      // int64[b] sample rows of a batch, two slots
    WS_INDICES2,
    WS_WSAMPLE,      // real[b] per-sample weights of the 'average' modes, two slots
    WS_WSAMPLE2,
    WS_PANEL_X,      // X_sub  b x s_pad
    WS_TC_XS2,       // second packed X[:, subset]^T | code^T panel (slot 1): the full-width product of step t reads the code^T
                     //   part of its panel while the prefetch of step t+1 packs the next one
    WS_COUNT
};

}  // namespace modl

struct modl_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    int cluster_ok = 0;
    int64_t launches = 0;
    // tunables (modl_ctx_set_option)
    int opt_bcd_cluster = 16;     // largest thread-block cluster tried by the dictionary update (0 = never)
    int opt_cd_warps = 0;         // warps per CTA of the CD kernel (0 = auto)
    int opt_force_global_gram = 0;// debug: never keep the Gram in shared memory
    int opt_bcd_pilot = 1;        // use the warp-specialised look-ahead dictionary kernel when the panel fits a cluster
    int opt_tc_gemm = 1;          // float contractions on the tensor cores (tcgen05 3xTF32); 0 = CUDA-core FFMA GEMM
    int opt_bcd_flag_barrier = 0; // grid-wide dictionary update: per-CTA epoch flags instead of one atomic counter (measured r02_a: slower, stays off)
    int opt_bcd_coop_min_cols = 128; // grid-wide dictionary update: fewest columns per CTA (fewer CTAs at the per-atom grid barrier;
                                    // measured r02_a: image shape 8.3 -> 5.5 us/atom at 128, worse again at 256)
    int opt_bcd_timing = 0;       // debug: record clock64 stamps inside the dictionary update
    int bcd_timing_k = 0;
    int opt_tc_split2 = 1;        // split the contraction of the two-destination GEMM ([B_[:, subset] | C_]: 26 tiles on 148 SMs otherwise)
    int opt_tc_raw_b = 1;         // full-width B_ product: the GEMM converts raw X rows itself (no packed X^T panel through HBM)
    int opt_bcd_blocked = 1;      // L2 ball without positivity, panel fits one cluster: block-wise coefficient-space solve (bcd_blocked.cuh)
    int opt_bcd_pipeline = 1;     // pilot kernel, L2 ball without positivity: keep two norm exchanges in flight (bcd_pilot.cuh)
    const float *code_packed = nullptr;   // packed code^T (A operand, 128-row blocks) of the current step, or NULL
    int panel_b_ready = 0;        // the B_[:, subset] panel of the current step is complete in WS_PANEL_B[slot]
    int bcd_grid_wide = 0;        // the last dictionary update ran on the cooperative grid (panel larger than one cluster): the
                                  // minibatch loop then keeps plain launches (fit_loop.cu)
    int capturing = 0;            // the launches of this call are being captured into a CUDA graph (fit_loop.cu): a workspace
                                  // slot that would have to grow returns MODL_EGROW instead (the step is then run eagerly).
                                  // 1: the mid-call event is one other streams wait for (an event-record node); 2: it is
                                  // the fork point of a second captured stream (a plain record)
    // set by the minibatch loop around a fused-graph step: a second stream on which the call may run work that is
    // independent of the dictionary (the B_[:, subset] gather) beside its critical path; ev_fork orders the two
    cudaStream_t fork_stream = nullptr;
    cudaEvent_t ev_fork[2] = {};
    // optional per-phase device timing of the fused step (modl_ctx_profile)
    int prof_on = 0;
    int prof_n = 0;                       // marks recorded in the current step
    int prof_phase[32] = {};
    cudaEvent_t prof_ev[33] = {};
    double prof_ms[MODL_PROF_PHASES] = {};
    int64_t prof_steps = 0;
    void *slot_ptr[modl::WS_COUNT] = {};
    size_t slot_bytes[modl::WS_COUNT] = {};
    // one call at a time per context: the workspace slots and the per-step flags above are shared state
    std::recursive_mutex mu;

    // returns a device pointer with at least `bytes` capacity (grow-only)
    int reserve(modl::WsSlot slot, size_t bytes, void **out);
};

namespace modl {

// Every extern "C" entry that takes a context opens with one of these: it serialises the calls on that
// context (workspace slots are shared state) and makes the context's device current for the duration of
// the call, restoring the caller's device on exit -- a process may drive several GPUs.
struct CtxGuard {
    modl_ctx *c;
    int prev = -1;
    explicit CtxGuard(const modl_ctx *ctx) : c(const_cast<modl_ctx *>(ctx)) {
        if (!c) return;
        c->mu.lock();
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != c->device && cudaSetDevice(c->device) == cudaSuccess) prev = cur;
    }
    ~CtxGuard() {
        if (!c) return;
        if (prev >= 0) cudaSetDevice(prev);
        c->mu.unlock();
    }
    CtxGuard(const CtxGuard &) = delete;
    CtxGuard &operator=(const CtxGuard &) = delete;
};

// "phase `ph` starts now" (no-op unless profiling is on)
inline void prof_mark(modl_ctx *ctx, cudaStream_t st, int ph) {
    if (!ctx->prof_on || ctx->prof_n >= 32) return;
    ctx->prof_phase[ctx->prof_n] = ph;
    cudaEventRecord(ctx->prof_ev[ctx->prof_n], st);
    ctx->prof_n += 1;
}

template <typename T>
inline int ws(modl_ctx *ctx, WsSlot slot, size_t count, T **out) {
    void *p = nullptr;
    int st = ctx->reserve(slot, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return st;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr unsigned kFullMask = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(kFullMask, v, o);
        v = u > v ? u : v;
    }
    return v;
}

// Block-wide sum; every thread gets the result.  `scratch` holds >= 33 elements of T.
// Deterministic (fixed tree).  Contains __syncthreads().
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();   // protect scratch from a previous use
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        T t = lane < nw ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

template <typename T> __device__ __forceinline__ T t_abs(T x);
template <> __device__ __forceinline__ float t_abs<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double t_abs<double>(double x) { return fabs(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }

#endif  // __CUDACC__

}  // namespace modl
