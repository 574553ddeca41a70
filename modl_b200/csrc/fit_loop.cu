// fit_loop.cu -- DictFact.partial_fit as ONE C call per call of the estimator method
// [ref: modl/decomposition/dict_fact.py:313-337 (partial_fit) and :495-526 (_single_batch_fit)].
//
// The reference walks the batches of a partial_fit call in Python; here the walk, the integer bookkeeping of
// every step (feature subset, n_iter_, sample_n_iter_, batch weight) and the stream choreography live behind the
// C ABI, so that the host cost of a step is its kernel launches and nothing else.
//
// Schedule of step t on two streams (main = the caller's stream, side = owned, lower priority):
//
//   side : PREFETCH(t)    index uploads, X rows of the packed [D_sub ; X_sub] panel + row norms, packed
//                         X[:, subset]^T, B_[:, subset] panel            -- nothing here depends on the dictionary
//   main : wait(PREFETCH) D rows -> [G ; Dx] (tcgen05) -> code solve -> [panel | C_] += code^T [X_sub | code]
//          ----- ev_code ----- dictionary update (16-CTA cluster, sequential over atoms) -> scatter
//   side : wait(ev_code), wait(dictionary kernel RESIDENT)  B_ = (1-w) B_ + w/b code^T X  over all p columns
//
// so the critical path of a step is  D gather -> Gram -> codes -> subset statistics -> dictionary update, and the
// full-width product (60 % of the step's flops, 40 MB of its HBM traffic) plus the next step's input preparation
// run on the 132 SMs the dictionary update leaves idle.  "Resident" matters: a 16-CTA cluster cannot be placed once
// one-CTA-per-SM GEMM tiles occupy every SM (round-1 finding), so the side stream waits on a flag that the
// dictionary kernel writes from its first instructions (cuStreamWaitValue32), and the full-width product sizes its
// grid for sm_count - 16.
//
// Sample-sharded data parallelism: with a communicator attached (modl_fit_set_comm) every rank runs
// this loop on its rows; the k x (k + s) increments the dictionary update needs are summed with ncclAllReduce on
// the main stream, the k x p increment of B_ on the side stream with a second communicator, behind the
// dictionary update.  NCCL is resolved at run time from the library the process already has (torch's).
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <new>
#include <vector>

#include "launch.h"

namespace modl {

// ---------------------------------------------------------------------------------------
// NCCL, resolved at run time (no link-time dependency: a single-GPU user never needs it)
// ---------------------------------------------------------------------------------------
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy torch.distributed already loaded
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    });
    return &api;
}

#define MODL_NCCL_TRY(expr)                                                                           \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            ::modl::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl_api()->GetErrorString(_r)); \
            return MODL_ECUDA;                                                                        \
        }                                                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------
// stream memory operations (driver entry points resolved through the runtime)
// ---------------------------------------------------------------------------------------
struct StreamMemOps {
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    bool ok = false;
};

static StreamMemOps *stream_mem_ops()
{
    static StreamMemOps ops;
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDriverEntryPointQueryResult qr;
        void *f = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            ops.wait32 = reinterpret_cast<decltype(ops.wait32)>(f);
        f = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            ops.write32 = reinterpret_cast<decltype(ops.write32)>(f);
        cudaGetLastError();
        ops.ok = ops.wait32 && ops.write32;
    });
    return &ops;
}

}  // namespace modl

using namespace modl;

// ---------------------------------------------------------------------------------------
// the loop object
// ---------------------------------------------------------------------------------------
struct modl_fit {
    static constexpr int NSTAGE = 3;     // device slots of host batches in flight
    static constexpr int NRING = 8;      // pinned copies of the small per-step inputs in flight

    modl_ctx *ctx = nullptr;
    cudaStream_t side = nullptr, copy = nullptr, d2h = nullptr;
    cudaEvent_t ev_pre[2] = {}, ev_code[2] = {}, ev_sub[2] = {}, ev_side[2] = {}, ev_d2h[2] = {};
    bool side_pending[2] = {false, false}, d2h_pending[2] = {false, false};
    int64_t step = 0;                    // steps run so far (slot = step & 1)
    unsigned serial = 0;                 // value the dictionary kernel of the current step writes to *d_flag
    unsigned *d_flag = nullptr;

    // host batches: device staging slots (and pinned bounce buffers for pageable sources)
    void *stage[NSTAGE] = {};
    void *bounce[NSTAGE] = {};
    size_t stage_bytes = 0, bounce_bytes = 0;
    cudaEvent_t ev_copied[NSTAGE] = {}, ev_free[NSTAGE] = {};
    bool stage_busy[NSTAGE] = {}, stage_copied[NSTAGE] = {};
    int64_t host_batches = 0;            // host batches staged so far (slot = host_batches % NSTAGE)

    // pinned ring: subset | order | indices | w_sample
    unsigned char *ring[NRING] = {};
    size_t ring_bytes = 0;
    cudaEvent_t ev_ring[NRING] = {};
    bool ring_busy[NRING] = {};

    // sharding
    int world = 1, rank = 0;
    ncclComm_t comm_main = nullptr, comm_side = nullptr;
    void *inc = nullptr, *inc_sub = nullptr;
    size_t inc_bytes = 0, inc_sub_bytes = 0;

    int overlap = 1;                     // 0: one stream, one fused call per step (the round-1 schedule)
    // CUDA graphs (one GPU, two-stream schedule): the launches of each of the step's three calls are stream-captured on
    // `cap` (nothing ever runs there), folded into last step's executable graph with cudaGraphExecUpdate (new pointers,
    // subset length, weights: same topology) and launched on the stream the call belongs to.  Inside a graph a kernel
    // follows its predecessor in ~1 us instead of ~3 us, and -- the reason this exists -- stays there while pinned host
    // rows saturate PCIe (stream launches: 5-10 us each then, scripts/ubench/graph_update_ubench.cu).
    //   graph = 1 (default): the three calls of the two-stream schedule as three graphs; the second stream and the code
    //   read-back wait for the dictionary kernel's residency flag instead of an event in the middle of the critical path (an
    //   event-record node cuts a graph into separately submitted pieces).
    //   graph = 2 (opt-in, measured 1-2 % slower: profiles/r02_w4_*): ONE graph per step on the caller's stream holds the
    //   critical path AND, forked inside the graph after the subset statistics, the full-width product (behind a one-thread
    //   kernel that polls the residency flag), joined at the end; the B_[:, subset] panel is gathered by the graph itself on
    //   the forked stream (B_ is final when the previous graph has ended), so the second stream only prepares the X side of
    //   the next step and the cycle step t -> step t+1 crosses no stream.
    int graph = 1;
    static constexpr int GRAPH_WARM_STEPS = 4;      // eager steps first: they size the workspace of both slots
    cudaStream_t cap = nullptr, cap2 = nullptr;
    cudaEvent_t ev_gfork[2] = {}, ev_gjoin[2] = {}; // fork / join of the second captured stream (never waited on outside)
    bool prev_fused = false;                        // the previous step ran as one fused graph (or its eager twin)
    cudaGraphExec_t gexec[2][4] = {};               // [slot][0 prefetch, 1 critical path (sharded: up to the all-reduce), 2 full-width
                                                    //  product, 3 sharded: fold-in + dictionary update]
    std::vector<cudaGraphExec_t> gexec_retired;     // replaced executables: destroyed once nothing can be in flight
    int64_t graph_launches = 0, graph_updates_failed = 0, graph_fallbacks = 0;
    int gate = 1;                        // side stream waits for the dictionary kernel to be resident
    cudaEvent_t ev_call = nullptr;       // "the caller's stream has reached this partial_fit call"
    cudaEvent_t code_ev[2] = {};         // the event that marked ev_code of the last step on each slot (a trace event when tracing)

    // optional timeline of the first steps (modl_fit_trace): CUDA events with timing on the streams of the loop
    static constexpr int TRACE_POINTS = 10;
    int trace_cap = 0, trace_n = 0;
    cudaEvent_t trace_base = nullptr;
    std::vector<cudaEvent_t> trace_ev;   // [trace_cap][TRACE_POINTS]
    std::vector<unsigned char> trace_set;
};

namespace modl {

static int grow(void **ptr, size_t *have, size_t need)
{
    if (*have >= need) return MODL_OK;
    if (*ptr) MODL_CUDA_TRY(cudaFree(*ptr));
    *ptr = nullptr; *have = 0;
    MODL_CUDA_TRY(cudaMalloc(ptr, need + need / 8 + 256));
    *have = need + need / 8 + 256;
    return MODL_OK;
}

static int fit_init(modl_fit *f)
{
    int lo = 0, hi = 0;
    MODL_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));      // lo = least urgent
    MODL_CUDA_TRY(cudaStreamCreateWithPriority(&f->side, cudaStreamNonBlocking, lo));
    MODL_CUDA_TRY(cudaStreamCreateWithPriority(&f->copy, cudaStreamNonBlocking, lo));
    MODL_CUDA_TRY(cudaStreamCreateWithPriority(&f->d2h, cudaStreamNonBlocking, lo));
    for (int i = 0; i < 2; ++i) {
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_pre[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_code[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_sub[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_side[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_d2h[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < modl_fit::NSTAGE; ++i) {
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_copied[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_free[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < modl_fit::NRING; ++i) MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_ring[i], cudaEventDisableTiming));
    MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_call, cudaEventDisableTiming));
    MODL_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&f->d_flag), 256));
    MODL_CUDA_TRY(cudaMemset(f->d_flag, 0, 256));
    if (const char *e = getenv("MODL_FIT_OVERLAP")) f->overlap = atoi(e);
    if (const char *e = getenv("MODL_FIT_GATE")) f->gate = atoi(e);
    if (const char *e = getenv("MODL_FIT_GRAPH")) f->graph = atoi(e);
    MODL_CUDA_TRY(cudaStreamCreateWithFlags(&f->cap, cudaStreamNonBlocking));
    MODL_CUDA_TRY(cudaStreamCreateWithFlags(&f->cap2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_gfork[i], cudaEventDisableTiming));
        MODL_CUDA_TRY(cudaEventCreateWithFlags(&f->ev_gjoin[i], cudaEventDisableTiming));
    }
    return MODL_OK;
}

static void fit_free(modl_fit *f)
{
    cudaDeviceSynchronize();
    if (f->comm_main) nccl_api()->CommDestroy(f->comm_main);
    if (f->comm_side) nccl_api()->CommDestroy(f->comm_side);
    for (int i = 0; i < 2; ++i) {
        if (f->ev_pre[i]) cudaEventDestroy(f->ev_pre[i]);
        if (f->ev_code[i]) cudaEventDestroy(f->ev_code[i]);
        if (f->ev_sub[i]) cudaEventDestroy(f->ev_sub[i]);
        if (f->ev_side[i]) cudaEventDestroy(f->ev_side[i]);
        if (f->ev_d2h[i]) cudaEventDestroy(f->ev_d2h[i]);
    }
    for (int i = 0; i < modl_fit::NSTAGE; ++i) {
        if (f->ev_copied[i]) cudaEventDestroy(f->ev_copied[i]);
        if (f->ev_free[i]) cudaEventDestroy(f->ev_free[i]);
        if (f->stage[i]) cudaFree(f->stage[i]);
        if (f->bounce[i]) cudaFreeHost(f->bounce[i]);
    }
    for (int i = 0; i < modl_fit::NRING; ++i) {
        if (f->ev_ring[i]) cudaEventDestroy(f->ev_ring[i]);
        if (f->ring[i]) cudaFreeHost(f->ring[i]);
    }
    if (f->ev_call) cudaEventDestroy(f->ev_call);
    if (f->trace_base) cudaEventDestroy(f->trace_base);
    for (cudaEvent_t e : f->trace_ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 4; ++j) if (f->gexec[i][j]) cudaGraphExecDestroy(f->gexec[i][j]);
    for (cudaGraphExec_t e : f->gexec_retired) cudaGraphExecDestroy(e);
    if (f->cap) cudaStreamDestroy(f->cap);
    if (f->cap2) cudaStreamDestroy(f->cap2);
    for (int i = 0; i < 2; ++i) {
        if (f->ev_gfork[i]) cudaEventDestroy(f->ev_gfork[i]);
        if (f->ev_gjoin[i]) cudaEventDestroy(f->ev_gjoin[i]);
    }
    if (f->d_flag) cudaFree(f->d_flag);
    if (f->inc) cudaFree(f->inc);
    if (f->inc_sub) cudaFree(f->inc_sub);
    if (f->side) cudaStreamDestroy(f->side);
    if (f->copy) cudaStreamDestroy(f->copy);
    if (f->d2h) cudaStreamDestroy(f->d2h);
}

// pinned slot of the small per-step inputs; blocks (almost never) until the uploads out of it have been issued and done
static int ring_slot(modl_fit *f, int64_t step, size_t need, unsigned char **out)
{
    const int r = (int)(step % modl_fit::NRING);
    if (f->ring_busy[r]) {
        MODL_CUDA_TRY(cudaEventSynchronize(f->ev_ring[r]));
        f->ring_busy[r] = false;
    }
    if (f->ring_bytes < need) {
        for (int i = 0; i < modl_fit::NRING; ++i) {
            if (f->ring_busy[i]) { MODL_CUDA_TRY(cudaEventSynchronize(f->ev_ring[i])); f->ring_busy[i] = false; }
            if (f->ring[i]) MODL_CUDA_TRY(cudaFreeHost(f->ring[i]));
            f->ring[i] = nullptr;
        }
        f->ring_bytes = 0;
        const size_t cap = need + need / 4 + 4096;
        for (int i = 0; i < modl_fit::NRING; ++i) MODL_CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&f->ring[i]), cap));
        f->ring_bytes = cap;
    }
    *out = f->ring[r];
    return MODL_OK;
}

// timeline point `pt` of the step being enqueued (no-op unless tracing); returns the event so that it can double as a
// dependency
static cudaEvent_t trace_mark(modl_fit *f, int pt, cudaStream_t st)
{
    if (f->trace_n >= f->trace_cap) return nullptr;
    const size_t i = (size_t)f->trace_n * modl_fit::TRACE_POINTS + (size_t)pt;
    cudaEventRecord(f->trace_ev[i], st);
    f->trace_set[i] = 1;
    return f->trace_ev[i];
}

// issue the host -> device copy of one batch into the next staging slot (copy stream)
template <typename T>
static int stage_batch(modl_fit *f, const T *h_rows, int64_t ldx, int64_t rows, int64_t p, int pinned, int *slot_out)
{
    const int j = (int)(f->host_batches % modl_fit::NSTAGE);
    const size_t bytes = (size_t)rows * (size_t)p * sizeof(T);
    if (f->stage_busy[j]) {                       // kernels of the batch that used this slot no longer read it
        MODL_CUDA_TRY(cudaStreamWaitEvent(f->copy, f->ev_free[j], 0));
        f->stage_busy[j] = false;
    }
    const void *src = h_rows;
    size_t src_pitch = (size_t)ldx * sizeof(T);
    if (!pinned) {
        if (f->stage_copied[j]) MODL_CUDA_TRY(cudaEventSynchronize(f->ev_copied[j]));   // the DMA out of this bounce buffer is over
        unsigned char *dst = static_cast<unsigned char *>(f->bounce[j]);
        for (int64_t r = 0; r < rows; ++r)
            memcpy(dst + (size_t)r * p * sizeof(T), reinterpret_cast<const unsigned char *>(h_rows) + (size_t)r * src_pitch,
                   (size_t)p * sizeof(T));
        src = f->bounce[j];
        src_pitch = (size_t)p * sizeof(T);
    }
    if (f->trace_n < f->trace_cap && f->trace_cap > 0) {
        // (attributed to the step that will consume the slot only when it is the next one: good enough for a timeline)
        trace_mark(f, 0, f->copy);
    }
    if (src_pitch == (size_t)p * sizeof(T))
        MODL_CUDA_TRY(cudaMemcpyAsync(f->stage[j], src, bytes, cudaMemcpyHostToDevice, f->copy));
    else
        MODL_CUDA_TRY(cudaMemcpy2DAsync(f->stage[j], (size_t)p * sizeof(T), src, src_pitch, (size_t)p * sizeof(T), (size_t)rows,
                                        cudaMemcpyHostToDevice, f->copy));
    MODL_CUDA_TRY(cudaEventRecord(f->ev_copied[j], f->copy));
    if (f->trace_n < f->trace_cap) trace_mark(f, 1, f->copy);
    f->stage_copied[j] = true;
    f->host_batches += 1;
    *slot_out = j;
    return MODL_OK;
}

// One thread polls the residency flag of the dictionary kernel (first node of the full-width product's branch inside the
// fused graph: a one-CTA-per-SM GEMM that starts first leaves no room for the 16-CTA cluster).  Bounded: a dictionary
// kernel that never started must not hang the device.
__global__ void wait_flag_kernel(unsigned *flag, unsigned serial)
{
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<const volatile unsigned *>(flag) - serial) < 0) {
        if (clock64() - t0 > 40000000LL) {          // ~20 ms: give up (counted: modl_fit_graph_stats)
            atomicAdd(flag + 1, 1u);
            break;
        }
        __nanosleep(100);
    }
}

// The fused step on streams (a, b2): critical path on a; after the subset statistics b2 forks, waits for the dictionary
// kernel to be resident and runs the full-width product; a waits for b2 at the end.  Captured (a = cap, b2 = cap2) this is
// the step's graph; on real streams it is the same step with plain launches.
template <typename T>
static int enqueue_fused(modl_fit *f, const modl_step_params *q0, int slot, cudaStream_t a, cudaStream_t b2)
{
    modl_ctx *ctx = f->ctx;
    modl_step_params q = *q0;
    q.phases = MODL_PHASE_CODE | MODL_PHASE_STATS_SUB | MODL_PHASE_DICT | MODL_PHASE_INPUTS_READY | MODL_PHASE_FUSED_APPLY |
               MODL_PHASE_GATHER_B;
    q.ev_after_apply_sub = f->ev_gfork[slot];
    ctx->fork_stream = b2;                          // the B_[:, subset] gather runs there, beside the Gram and the code solve
    const int status = batch_fit_impl<T>(ctx, &q, a);
    ctx->fork_stream = nullptr;
    MODL_TRY(status);
    MODL_CUDA_TRY(cudaStreamWaitEvent(b2, f->ev_gfork[slot], 0));
    if (q.subset_len > 0) {                         // a dictionary kernel was launched: it raises the flag
        wait_flag_kernel<<<1, 1, 0, b2>>>(f->d_flag, f->serial);
        MODL_LAUNCH_CHECK(ctx);
    }
    q.phases = MODL_PHASE_STATS_B;
    q.inc_sub = nullptr;
    q.ev_after_apply_sub = nullptr;
    q.sm_avail = ctx->sm_count - 16;
    MODL_TRY(batch_fit_impl<T>(ctx, &q, b2));
    MODL_CUDA_TRY(cudaEventRecord(f->ev_gjoin[slot], b2));
    MODL_CUDA_TRY(cudaStreamWaitEvent(a, f->ev_gjoin[slot], 0));
    return MODL_OK;
}

// One call of the step (`which`: 0 prefetch, 1 critical path, 2 full-width product) on `st`: launched directly, or --
// `use_graph` -- captured, folded into the slot's executable graph and launched as one.  A call that cannot be captured
// (a workspace slot has to grow, an option took a path with a pageable upload) is run eagerly instead.
template <typename T>
static int run_call(modl_fit *f, bool use_graph, int which, int slot, const modl_step_params *q, cudaStream_t st, bool fused = false)
{
    modl_ctx *ctx = f->ctx;
    // fused: the whole step (two captured streams); its eager twin runs the branch on the loop's second stream
    auto eager = [&]() { return fused ? enqueue_fused<T>(f, q, slot, st, f->side) : batch_fit_impl<T>(ctx, q, st); };
    if (!use_graph) return eager();
    const float *code_packed = ctx->code_packed;
    const int panel_b_ready = ctx->panel_b_ready;
    const int64_t launches = ctx->launches;
    if (cudaStreamBeginCapture(f->cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        f->graph = 0;
        return eager();
    }
    ctx->capturing = fused ? 2 : 1;
    const int status = fused ? enqueue_fused<T>(f, q, slot, f->cap, f->cap2) : batch_fit_impl<T>(ctx, q, f->cap);
    ctx->capturing = 0;
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(f->cap, &g);
    if (status != MODL_OK || ce != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        if (status != MODL_OK && status != MODL_EGROW) return status;
        if (status == MODL_OK) f->graph = 0;        // something on this path is not capturable: stop trying
        ctx->code_packed = code_packed; ctx->panel_b_ready = panel_b_ready; ctx->launches = launches;
        f->graph_fallbacks += 1;
        return eager();
    }
    cudaGraphExec_t &ex = f->gexec[slot][which];
    bool current = false;
    if (ex) {
        cudaGraphExecUpdateResultInfo info;
        current = cudaGraphExecUpdate(ex, g, &info) == cudaSuccess;
        if (!current) {                             // another topology (empty subset, another kernel variant): rebuild
            cudaGetLastError();
            f->gexec_retired.push_back(ex);         // a launch of it may still be running
            ex = nullptr;
            f->graph_updates_failed += 1;
        }
    }
    if (!current && cudaGraphInstantiate(&ex, g, 0) != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(g);
        ex = nullptr;
        f->graph = 0;
        ctx->code_packed = code_packed; ctx->panel_b_ready = panel_b_ready; ctx->launches = launches;
        f->graph_fallbacks += 1;
        return eager();
    }
    cudaGraphDestroy(g);
    MODL_CUDA_TRY(cudaGraphLaunch(ex, st));
    f->graph_launches += 1;
    return MODL_OK;
}

template <typename T>
static int partial_fit_impl(modl_fit *f, const modl_fit_params *e, const modl_fit_batches *io, void *stream)
{
    modl_ctx *ctx = f->ctx;
    cudaStream_t main_st = (cudaStream_t)stream;
    const int64_t k = e->n_components, p = e->n_features, bs = e->batch_size, n = io->n_rows;
    MODL_REQUIRE(k >= 1 && p >= 1 && bs >= 1 && n >= 0 && e->n_samples >= 1, "shapes");
    MODL_REQUIRE(io->X && io->ldx >= p, "X");
    MODL_REQUIRE(io->sampler && io->h_orders && io->h_n_iter, "sampler / orders / n_iter");
    MODL_REQUIRE(io->x_location >= 0 && io->x_location <= 2, "x_location");
    MODL_REQUIRE(f->world == 1 || (f->comm_main && f->comm_side), "sharded loop without communicators");
    const bool need_w = e->Dx_agg == MODL_AGG_AVERAGE || e->G_agg == MODL_AGG_AVERAGE;
    MODL_REQUIRE(!need_w || io->h_w_sample, "the 'average' modes need h_w_sample");
    const int64_t nb = ceil_div(n, bs);
    if (nb == 0) return MODL_OK;
    const bool host_x = io->x_location != 0;
    const bool sharded = f->world > 1;
    // the two-stream schedule covers the variational optimizer; 'sgd' and profiled runs take the single fused call
    const bool overlap = f->overlap && !e->optimizer_sgd && !ctx->prof_on;
    StreamMemOps *smo = stream_mem_ops();
    const bool gate = overlap && f->gate && smo->ok;
    if (f->gexec_retired.size() > 64) {             // rare: topology changes pile up only with alternating empty subsets
        MODL_CUDA_TRY(cudaDeviceSynchronize());
        for (cudaGraphExec_t ex : f->gexec_retired) cudaGraphExecDestroy(ex);
        f->gexec_retired.clear();
    }
    const T *X = static_cast<const T *>(io->X);

    if (host_x) {
        const size_t need = (size_t)bs * (size_t)p * sizeof(T);
        if (f->stage_bytes < need) {
            MODL_CUDA_TRY(cudaDeviceSynchronize());
            for (int i = 0; i < modl_fit::NSTAGE; ++i) {
                if (f->stage[i]) MODL_CUDA_TRY(cudaFree(f->stage[i]));
                f->stage[i] = nullptr;
                f->stage_busy[i] = f->stage_copied[i] = false;
            }
            for (int i = 0; i < modl_fit::NSTAGE; ++i) MODL_CUDA_TRY(cudaMalloc(&f->stage[i], need));
            f->stage_bytes = need;
        }
        if (io->x_location == 2 && f->bounce_bytes < need) {
            for (int i = 0; i < modl_fit::NSTAGE; ++i) {
                if (f->stage_copied[i]) MODL_CUDA_TRY(cudaEventSynchronize(f->ev_copied[i]));
                if (f->bounce[i]) MODL_CUDA_TRY(cudaFreeHost(f->bounce[i]));
                f->bounce[i] = nullptr;
            }
            for (int i = 0; i < modl_fit::NSTAGE; ++i) MODL_CUDA_TRY(cudaMallocHost(&f->bounce[i], need));
            f->bounce_bytes = need;
        }
    }
    if (sharded) {
        MODL_TRY(grow(&f->inc, &f->inc_bytes, sizeof(T) * (size_t)(k * (k + p))));
        MODL_TRY(grow(&f->inc_sub, &f->inc_sub_bytes, sizeof(T) * (size_t)(k * k + k * panel_ld(p))));
    }

    MODL_REQUIRE(io->fence >= 0 && io->fence <= 2, "fence");
    const bool fence = io->fence == 1 || (io->fence == 0 && !host_x) || f->step == 0;
    if (overlap && fence) {
        // whatever the caller enqueued on its stream before this call (a new X, a state array set by hand) is visible to
        // the streams of the loop.  The previous call's dictionary update is on that stream too, so this also keeps the next
        // block's input preparation from starting under it: callers whose rows are final say so (io->fence)
        MODL_CUDA_TRY(cudaEventRecord(f->ev_call, main_st));
        MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_call, 0));
    }
    const size_t ring_need = sizeof(int64_t) * (size_t)(p + k + bs) + sizeof(T) * (size_t)bs + 64;
    std::vector<int> staged((size_t)nb, -1);
    int64_t issued = 0;                      // host batches of this call whose copy has been issued
    const int ahead = modl_fit::NSTAGE - 1;  // copies kept in flight beyond the batch being enqueued

    for (int64_t i = 0; i < nb; ++i) {
        const int64_t r0 = i * bs, b = (r0 + bs <= n ? bs : n - r0);
        const int slot = (int)(f->step & 1), prev = slot ^ 1;
        // (not while the dictionary update runs on the cooperative grid -- panels larger than one cluster: with the side
        //  stream's graphs beside that kernel the fMRI-shaped step ran 2x slower, profiles/r02_nr_*; plain launches there)
        const bool use_graph = overlap && f->graph && f->step >= modl_fit::GRAPH_WARM_STEPS && !ctx->bcd_grid_wide;
        // ---- host rows on their way ----
        if (host_x) {
            for (; issued < nb && issued <= i + ahead - 1; ++issued) {
                const int64_t q0 = issued * bs, qb = (q0 + bs <= n ? bs : n - q0);
                MODL_TRY(stage_batch<T>(f, X + q0 * io->ldx, io->ldx, qb, p, io->x_location == 1, &staged[(size_t)issued]));
            }
        }
        const int sj = host_x ? staged[(size_t)i] : -1;
        const T *Xb = host_x ? static_cast<const T *>(f->stage[sj]) : X + r0 * io->ldx;
        const int64_t ldxb = host_x ? p : io->ldx;

        // ---- the integer / scalar statements of the step, in the reference's order [ref: :507-515, :672] ----
        unsigned char *rp = nullptr;
        MODL_TRY(ring_slot(f, f->step, ring_need, &rp));
        int64_t *h_subset = reinterpret_cast<int64_t *>(rp);
        int64_t *h_order = h_subset + p;
        int64_t *h_idx = h_order + k;
        T *h_w = reinterpret_cast<T *>(h_idx + bs);
        const int64_t s = modl_sampler_yield_subset(io->sampler, e->reduction, h_subset);            // [ref: :507]
        MODL_REQUIRE(s >= 0 && s <= p, "sampler returned an invalid subset");
        const int64_t b_global = b * f->world;          // equal shards: every rank passes the same row count (checked by the caller)
        // contiguous ascending rows (the common case) address the per-sample state through a pointer offset
        int64_t base = io->h_sample_indices ? io->h_sample_indices[r0] : r0;
        bool contiguous = true;
        if (io->h_sample_indices)
            for (int64_t j = 1; j < b && contiguous; ++j) contiguous = io->h_sample_indices[r0 + j] == base + j;
        for (int64_t j = 0; j < b; ++j) {
            const int64_t row = io->h_sample_indices ? io->h_sample_indices[r0 + j] : r0 + j;
            MODL_REQUIRE(row >= 0 && row < e->n_samples, "sample index out of range");
            h_idx[j] = row;
        }
        if (io->update_counters) {
            *io->h_n_iter += b_global;                                                                // [ref: :509]
            if (io->h_sample_n_iter)
                for (int64_t j = 0; j < b; ++j) io->h_sample_n_iter[h_idx[j]] += 1;                    // [ref: :510]
        }
        const double w = modl_batch_weight(*io->h_n_iter, b_global, e->learning_rate, 0.);             // [ref: :515]
        memcpy(h_order, io->h_orders + i * k, sizeof(int64_t) * (size_t)k);                            // [ref: :672]
        if (need_w) memcpy(h_w, static_cast<const T *>(io->h_w_sample) + r0, sizeof(T) * (size_t)b);
        if (io->h_last_subset) memcpy(io->h_last_subset, h_subset, sizeof(int64_t) * (size_t)s);
        if (io->h_last_subset_len) *io->h_last_subset_len = s;

        // ---- the parameter block of the step ----
        modl_step_params q;
        memset(&q, 0, sizeof(q));
        q.n_features = p; q.n_components = k; q.batch_size = b;
        q.X = Xb; q.ldx = ldxb;
        q.h_subset = h_subset; q.subset_len = s; q.h_order = h_order; q.w = w;
        q.components = e->components; q.C = e->C; q.B = e->B; q.comp_norm = e->comp_norm; q.G_full = e->G_full;
        q.reduction = e->reduction; q.code_alpha = e->code_alpha; q.code_l1_ratio = e->code_l1_ratio;
        q.comp_l1_ratio = e->comp_l1_ratio; q.tol = e->tol; q.step_size = e->step_size;
        q.max_iter = e->max_iter; q.code_pos = e->code_pos; q.comp_pos = e->comp_pos; q.Dx_agg = e->Dx_agg; q.G_agg = e->G_agg;
        q.optimizer_sgd = e->optimizer_sgd;
        q.sweeps = io->sweeps;
        q.global_batch = sharded ? b_global : 0;
        q.slot = slot;
        q.h_inputs_mapped = 1;          // subset and order live in the pinned ring slot of this step
        const int64_t off = contiguous ? base : 0;
        q.n_samples = e->n_samples - off;
        q.code = static_cast<T *>(e->code) + off * k;
        q.Dx_average = e->Dx_average ? static_cast<T *>(e->Dx_average) + off * k : nullptr;
        q.G_average = e->G_average ? static_cast<T *>(e->G_average) + off * k * k : nullptr;
        cudaStream_t up_st = overlap ? f->side : main_st;       // stream of the small uploads
        if (overlap && f->side_pending[prev]) {
            // the buffers PREFETCH rewrites were last read before ev_code of the previous step; B_ (gathered below on one
            // GPU) is final once the previous step's full-width product has run -- same stream, in order
            if (f->code_ev[prev]) MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->code_ev[prev], 0));
        }
        if (!contiguous) {
            int64_t *d_idx = nullptr;
            MODL_TRY(ws<int64_t>(ctx, slot ? WS_INDICES2 : WS_INDICES, (size_t)bs, &d_idx));
            MODL_CUDA_TRY(cudaMemcpyAsync(d_idx, h_idx, sizeof(int64_t) * (size_t)b, cudaMemcpyHostToDevice, up_st));
            q.indices = d_idx;
        }
        if (need_w) {
            T *d_w = nullptr;
            MODL_TRY(ws<T>(ctx, slot ? WS_WSAMPLE2 : WS_WSAMPLE, (size_t)bs, &d_w));
            MODL_CUDA_TRY(cudaMemcpyAsync(d_w, h_w, sizeof(T) * (size_t)b, cudaMemcpyHostToDevice, up_st));
            q.w_sample = d_w;
        }

        if (!overlap) {
            // ---- one stream: the fused call (sharded: two calls around the all-reduce) ----
            if (host_x) MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_copied[sj], 0));
            if (!sharded) {
                q.phases = 0;
                MODL_TRY(batch_fit_impl<T>(ctx, &q, main_st));
            } else {
                q.stats_inc = f->inc;
                q.phases = MODL_PHASE_CODE | MODL_PHASE_STATS;
                MODL_TRY(batch_fit_impl<T>(ctx, &q, main_st));
                MODL_NCCL_TRY(nccl_api()->AllReduce(f->inc, f->inc, (size_t)(k * (k + p)), sizeof(T) == 4 ? ncclFloat : ncclDouble,
                                                    ncclSum, f->comm_main, main_st));
                q.phases = MODL_PHASE_APPLY | MODL_PHASE_DICT | MODL_PHASE_REUSE_SUBSET;
                MODL_TRY(batch_fit_impl<T>(ctx, &q, main_st));
            }
            MODL_CUDA_TRY(cudaEventRecord(f->ev_ring[f->step % modl_fit::NRING], main_st));
            f->ring_busy[f->step % modl_fit::NRING] = true;
            if (io->h_code_out) {
                MODL_CUDA_TRY(cudaMemcpyAsync(static_cast<T *>(io->h_code_out) + r0 * k, ctx->slot_ptr[WS_CODE_BATCH],
                                              sizeof(T) * (size_t)(b * k), cudaMemcpyDeviceToHost, main_st));
            }
            if (host_x) {
                MODL_CUDA_TRY(cudaEventRecord(f->ev_free[sj], main_st));
                f->stage_busy[sj] = true;
            }
            f->step += 1;
            continue;
        }

        if (use_graph && gate && !sharded && f->graph >= 2) {
            // ---- one graph per step (see modl_fit::graph) ----
            // side: the X side of this step's inputs.  Its buffers were last read by the previous step's subset statistics:
            // the dictionary kernel that follows them has raised the flag (an eager predecessor: its code event as well)
            if (f->side_pending[prev] && f->code_ev[prev]) MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->code_ev[prev], 0));
            if (f->serial > 0) smo->wait32((CUstream)f->side, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WAIT_VALUE_GEQ);
            if (host_x) MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_copied[sj], 0));
            trace_mark(f, 2, f->side);
            q.phases = MODL_PHASE_PREFETCH;                 // without FUSED_APPLY: the graph gathers B_[:, subset] itself
            MODL_TRY(run_call<T>(f, true, 0, slot, &q, f->side));
            MODL_CUDA_TRY(cudaEventRecord(f->ev_pre[slot], f->side));
            trace_mark(f, 3, f->side);
            MODL_CUDA_TRY(cudaEventRecord(f->ev_ring[f->step % modl_fit::NRING], f->side));
            f->ring_busy[f->step % modl_fit::NRING] = true;

            // main: the step
            MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_pre[slot], 0));
            for (int j = 0; j < 2; ++j)
                if (f->side_pending[j]) {                   // an eager predecessor's full-width product (second stream): B_ is final
                    MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_side[j], 0));
                    f->side_pending[j] = false;
                }
            if (f->d2h_pending[prev]) {                     // the previous batch code has been read back before CODE overwrites it
                MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_d2h[prev], 0));
                f->d2h_pending[prev] = false;
            }
            f->serial += 1;
            q.start_flag = f->d_flag;
            q.start_serial = f->serial;
            trace_mark(f, 4, main_st);
            f->code_ev[slot] = nullptr;
            MODL_TRY(run_call<T>(f, true, 1, slot, &q, main_st, true));
            if (s == 0)                                     // no dictionary kernel: the flag still reaches the serial
                smo->write32((CUstream)main_st, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WRITE_VALUE_DEFAULT);
            trace_mark(f, 6, main_st);
            if (io->h_code_out) {
                smo->wait32((CUstream)f->d2h, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WAIT_VALUE_GEQ);
                trace_mark(f, 5, f->d2h);                   // "codes done", as the read-back stream sees it
                MODL_CUDA_TRY(cudaMemcpyAsync(static_cast<T *>(io->h_code_out) + r0 * k, ctx->slot_ptr[WS_CODE_BATCH],
                                              sizeof(T) * (size_t)(b * k), cudaMemcpyDeviceToHost, f->d2h));
                MODL_CUDA_TRY(cudaEventRecord(f->ev_d2h[slot], f->d2h));
                trace_mark(f, 9, f->d2h);
                f->d2h_pending[slot] = true;
            }
            if (f->trace_n < f->trace_cap) f->trace_n += 1;
            if (host_x) {                                   // the graph's full-width product was the last reader of the rows
                MODL_CUDA_TRY(cudaEventRecord(f->ev_free[sj], main_st));
                f->stage_busy[sj] = true;
            }
            f->prev_fused = true;
            f->step += 1;
            continue;
        }
        if (f->prev_fused) {
            // back to plain launches (graphs switched off, or a sharded / profiled call): the second stream must see the
            // end of the last fused graph (B_ final, its buffers read) before it prepares this step
            MODL_CUDA_TRY(cudaEventRecord(f->ev_call, main_st));
            MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_call, 0));
            f->prev_fused = false;
        }

        // ---- side: everything of the step that does not depend on the dictionary ----
        if (host_x) MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_copied[sj], 0));
        trace_mark(f, 2, f->side);
        q.phases = MODL_PHASE_PREFETCH | (sharded ? 0 : MODL_PHASE_FUSED_APPLY);
        MODL_TRY(run_call<T>(f, use_graph, 0, slot, &q, f->side));
        MODL_CUDA_TRY(cudaEventRecord(f->ev_pre[slot], f->side));
        trace_mark(f, 3, f->side);
        MODL_CUDA_TRY(cudaEventRecord(f->ev_ring[f->step % modl_fit::NRING], f->side));
        f->ring_busy[f->step % modl_fit::NRING] = true;

        // ---- main: the critical path ----
        MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_pre[slot], 0));
        if (f->d2h_pending[prev]) {              // the previous batch code has been read back before CODE overwrites it
            MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_d2h[prev], 0));
            f->d2h_pending[prev] = false;
        }
        f->serial += 1;
        q.start_flag = gate ? f->d_flag : nullptr;
        q.start_serial = f->serial;
        trace_mark(f, 4, main_st);
        // ev_code: the codes are solved, the subset statistics folded in; the dictionary update follows on this stream
        // Graph-replayed step with the residency flag: no event in the middle of the critical path (an event-record node
        // cuts the graph into separately submitted pieces, each paying the launch latency PCIe traffic inflates).  The
        // dictionary kernel follows the subset statistics in stream order, so "its cluster is resident" (the flag the
        // second stream already waits for) implies "codes solved, subset statistics folded in".
        const bool by_flag = use_graph && gate && !sharded;
        cudaEvent_t evc = by_flag ? nullptr : f->ev_code[slot];
        if (!by_flag && f->trace_n < f->trace_cap) {
            evc = f->trace_ev[(size_t)f->trace_n * modl_fit::TRACE_POINTS + 5];
            f->trace_set[(size_t)f->trace_n * modl_fit::TRACE_POINTS + 5] = 1;
        }
        f->code_ev[slot] = evc;                             // NULL: the second stream has waited for the flag, in order
        if (!sharded) {
            q.phases = MODL_PHASE_CODE | MODL_PHASE_STATS_SUB | MODL_PHASE_DICT | MODL_PHASE_INPUTS_READY | MODL_PHASE_FUSED_APPLY;
            q.ev_after_apply_sub = evc;                     // recorded between the subset statistics and the dictionary update
            MODL_TRY(run_call<T>(f, use_graph, 1, slot, &q, main_st));
        } else {
            q.stats_inc = f->inc;
            q.inc_sub = f->inc_sub;
            q.phases = MODL_PHASE_CODE | MODL_PHASE_STATS_SUB | MODL_PHASE_INPUTS_READY;
            MODL_TRY(run_call<T>(f, use_graph, 1, slot, &q, main_st));
            MODL_CUDA_TRY(cudaEventRecord(evc, main_st));
            MODL_NCCL_TRY(nccl_api()->AllReduce(f->inc_sub, f->inc_sub, (size_t)(k * k + k * panel_ld(s)),
                                                sizeof(T) == 4 ? ncclFloat : ncclDouble, ncclSum, f->comm_main, main_st));
            q.phases = MODL_PHASE_APPLY_SUB | MODL_PHASE_DICT | MODL_PHASE_INPUTS_READY;
            q.ev_after_apply_sub = f->ev_sub[slot];         // B_[:, subset] has been read
            MODL_TRY(run_call<T>(f, use_graph, 3, slot, &q, main_st));
        }
        if (gate && !(by_flag && s > 0))    // whatever the dictionary phase launched (or did not: empty subset), the flag reaches
            smo->write32((CUstream)main_st, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WRITE_VALUE_DEFAULT);
        trace_mark(f, 6, main_st);

        // ---- read-back of the batch code (optional), on its own stream ----
        if (io->h_code_out) {
            if (by_flag) smo->wait32((CUstream)f->d2h, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WAIT_VALUE_GEQ);
            else MODL_CUDA_TRY(cudaStreamWaitEvent(f->d2h, evc, 0));
            MODL_CUDA_TRY(cudaMemcpyAsync(static_cast<T *>(io->h_code_out) + r0 * k, ctx->slot_ptr[WS_CODE_BATCH],
                                          sizeof(T) * (size_t)(b * k), cudaMemcpyDeviceToHost, f->d2h));
            MODL_CUDA_TRY(cudaEventRecord(f->ev_d2h[slot], f->d2h));
            trace_mark(f, 9, f->d2h);
            f->d2h_pending[slot] = true;
        }

        // ---- side: the full-width statistic, behind the dictionary update ----
        if (!by_flag) MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, evc, 0));
        if (gate) smo->wait32((CUstream)f->side, (CUdeviceptr)(uintptr_t)f->d_flag, f->serial, CU_STREAM_WAIT_VALUE_GEQ);
        if (by_flag) trace_mark(f, 5, f->side);             // "codes done", as the second stream sees it
        trace_mark(f, 7, f->side);
        q.phases = MODL_PHASE_STATS_B;
        q.inc_sub = nullptr;
        q.ev_after_apply_sub = nullptr;
        q.sm_avail = ctx->sm_count - 16;
        MODL_TRY(run_call<T>(f, use_graph, 2, slot, &q, f->side));
        if (sharded) {
            MODL_NCCL_TRY(nccl_api()->AllReduce(static_cast<T *>(f->inc) + k * k, static_cast<T *>(f->inc) + k * k, (size_t)(k * p),
                                                sizeof(T) == 4 ? ncclFloat : ncclDouble, ncclSum, f->comm_side, f->side));
            MODL_CUDA_TRY(cudaStreamWaitEvent(f->side, f->ev_sub[slot], 0));
            q.phases = MODL_PHASE_APPLY_B;
            MODL_TRY(batch_fit_impl<T>(ctx, &q, f->side));
        }
        MODL_CUDA_TRY(cudaEventRecord(f->ev_side[slot], f->side));
        trace_mark(f, 8, f->side);
        if (f->trace_n < f->trace_cap) f->trace_n += 1;
        f->side_pending[slot] = true;
        if (host_x) {
            MODL_CUDA_TRY(cudaEventRecord(f->ev_free[sj], f->side));
            f->stage_busy[sj] = true;
        }
        f->step += 1;
    }

    // the caller's stream sees the whole call: B_ is final, the batch code has been read back
    if (overlap) {
        const int last = (int)((f->step - 1) & 1);
        if (f->side_pending[last]) MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_side[last], 0));   // (a fused step joins inside its graph)
        if (f->d2h_pending[last]) MODL_CUDA_TRY(cudaStreamWaitEvent(main_st, f->ev_d2h[last], 0));
    }
    if (io->wait_host) {
        // the caller may reuse X (and read h_code_out) on return
        if (host_x)
            for (int i = 0; i < modl_fit::NSTAGE; ++i)
                if (f->stage_copied[i]) MODL_CUDA_TRY(cudaEventSynchronize(f->ev_copied[i]));
        if (io->h_code_out) {
            if (overlap) { for (int i = 0; i < 2; ++i) if (f->d2h_pending[i]) MODL_CUDA_TRY(cudaEventSynchronize(f->ev_d2h[i])); }
            else MODL_CUDA_TRY(cudaStreamSynchronize(main_st));
        }
    }
    return MODL_OK;
}

}  // namespace modl

extern "C" {

int modl_fit_create(modl_ctx *ctx, modl_fit **out)
{
    MODL_REQUIRE(ctx && out, "null");
    *out = nullptr;
    CtxGuard g_(ctx);
    modl_fit *f = new (std::nothrow) modl_fit();
    if (!f) return MODL_ENOMEM;
    f->ctx = ctx;
    const int status = fit_init(f);
    if (status != MODL_OK) {
        fit_free(f);
        delete f;
        return status;
    }
    *out = f;
    return MODL_OK;
}

void modl_fit_destroy(modl_fit *f)
{
    if (!f) return;
    {
        CtxGuard g_(f->ctx);
        fit_free(f);
    }
    delete f;
}

int modl_fit_set_option(modl_fit *f, const char *name, int value)
{
    MODL_REQUIRE(f && name, "null");
    CtxGuard g_(f->ctx);
    if (!strcmp(name, "overlap")) f->overlap = value;
    else if (!strcmp(name, "gate")) f->gate = value;
    else if (!strcmp(name, "graph")) f->graph = value;
    else { set_error("unknown option %s", name); return MODL_EINVAL; }
    return MODL_OK;
}

int modl_fit_graph_stats(modl_fit *f, int64_t *h_out4)
{
    MODL_REQUIRE(f && h_out4, "null");
    CtxGuard g_(f->ctx);
    h_out4[0] = f->graph_launches; h_out4[1] = f->graph_updates_failed; h_out4[2] = f->graph_fallbacks;
    unsigned timeouts = 0;
    MODL_CUDA_TRY(cudaMemcpy(&timeouts, f->d_flag + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));     // synchronises
    h_out4[3] = (int64_t)timeouts;
    return MODL_OK;
}

/* Debug: timeline of the next `steps` steps of the two-stream schedule (0 = off).  modl_fit_trace_read synchronises the
 * device and writes, per traced step, the milliseconds since the trace was armed at which each of these points was
 * reached on its stream (NaN = not recorded):  0 H2D start, 1 H2D end (copy stream); 2 PREFETCH start, 3 PREFETCH end
 * (second stream); 4 critical path start, 5 codes + subset statistics done, 6 dictionary update done (caller's stream);
 * 7 full-width product start, 8 full-width product end (second stream); 9 code read-back done. */
int modl_fit_trace(modl_fit *f, int steps)
{
    MODL_REQUIRE(f && steps >= 0 && steps <= 4096, "trace steps");
    CtxGuard g_(f->ctx);
    MODL_CUDA_TRY(cudaDeviceSynchronize());
    for (cudaEvent_t e : f->trace_ev) if (e) cudaEventDestroy(e);
    f->trace_ev.clear(); f->trace_set.clear();
    f->trace_cap = steps; f->trace_n = 0;
    if (steps == 0) return MODL_OK;
    if (!f->trace_base) MODL_CUDA_TRY(cudaEventCreate(&f->trace_base));
    f->trace_ev.assign((size_t)steps * modl_fit::TRACE_POINTS, nullptr);
    f->trace_set.assign((size_t)steps * modl_fit::TRACE_POINTS, 0);
    for (cudaEvent_t &e : f->trace_ev) MODL_CUDA_TRY(cudaEventCreate(&e));
    MODL_CUDA_TRY(cudaEventRecord(f->trace_base, nullptr));
    MODL_CUDA_TRY(cudaEventSynchronize(f->trace_base));
    return MODL_OK;
}

int modl_fit_trace_read(modl_fit *f, double *h_ms, int capacity_steps, int *h_steps)
{
    MODL_REQUIRE(f && h_ms && h_steps && capacity_steps >= 0, "trace buffers");
    CtxGuard g_(f->ctx);
    MODL_CUDA_TRY(cudaDeviceSynchronize());
    const int n = f->trace_n < capacity_steps ? f->trace_n : capacity_steps;
    for (int i = 0; i < n; ++i)
        for (int pt = 0; pt < modl_fit::TRACE_POINTS; ++pt) {
            const size_t e = (size_t)i * modl_fit::TRACE_POINTS + (size_t)pt;
            float ms = 0.f;
            double v = std::nan("");
            if (f->trace_set[e] && cudaEventElapsedTime(&ms, f->trace_base, f->trace_ev[e]) == cudaSuccess) v = ms;
            h_ms[e] = v;
        }
    cudaGetLastError();
    *h_steps = n;
    return MODL_OK;
}

int modl_nccl_unique_id(void *h_out, int64_t capacity)
{
    MODL_REQUIRE(h_out && capacity >= (int64_t)sizeof(ncclUniqueId), "modl_nccl_unique_id needs a 128-byte buffer");
    NcclApi *api = nccl_api();
    if (!api->ok) { set_error("NCCL (libnccl.so.2) could not be loaded"); return MODL_ECUDA; }
    MODL_NCCL_TRY(api->GetUniqueId(static_cast<ncclUniqueId *>(h_out)));
    return MODL_OK;
}

int modl_fit_set_comm(modl_fit *f, int world, int rank, const void *h_id_main, const void *h_id_side)
{
    MODL_REQUIRE(f && world >= 1 && rank >= 0 && rank < world, "communicator shape");
    CtxGuard g_(f->ctx);
    if (world == 1) { f->world = 1; f->rank = 0; return MODL_OK; }
    MODL_REQUIRE(h_id_main && h_id_side, "two NCCL unique ids (main and side stream communicators)");
    NcclApi *api = nccl_api();
    if (!api->ok) { set_error("NCCL (libnccl.so.2) could not be loaded"); return MODL_ECUDA; }
    ncclUniqueId a, b;
    memcpy(&a, h_id_main, sizeof(a));
    memcpy(&b, h_id_side, sizeof(b));
    MODL_NCCL_TRY(api->CommInitRank(&f->comm_main, world, a, rank));
    MODL_NCCL_TRY(api->CommInitRank(&f->comm_side, world, b, rank));
    f->world = world;
    f->rank = rank;
    return MODL_OK;
}

int modl_fit_synchronize(modl_fit *f)
{
    MODL_REQUIRE(f, "null");
    CtxGuard g_(f->ctx);
    MODL_CUDA_TRY(cudaStreamSynchronize(f->copy));
    MODL_CUDA_TRY(cudaStreamSynchronize(f->side));
    MODL_CUDA_TRY(cudaStreamSynchronize(f->d2h));
    return MODL_OK;
}

int modl_partial_fit_f32(modl_fit *f, const modl_fit_params *est, const modl_fit_batches *io, void *stream)
{
    MODL_REQUIRE(f && est && io, "null");
    CtxGuard g_(f->ctx);
    return partial_fit_impl<float>(f, est, io, stream);
}

int modl_partial_fit_f64(modl_fit *f, const modl_fit_params *est, const modl_fit_batches *io, void *stream)
{
    MODL_REQUIRE(f && est && io, "null");
    CtxGuard g_(f->ctx);
    return partial_fit_impl<double>(f, est, io, stream);
}

}  // extern "C"
