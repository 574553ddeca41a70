// Micro-benchmark behind the graph-replayed minibatch step: kernel-to-kernel latency of (A) plain stream launches,
// (B) a static CUDA graph, (C) a graph re-captured and cudaGraphExecUpdate'd before every launch (new kernel arguments and
// grid sizes each time -- what a step with a fresh subset / batch pointer needs), each alone and under a saturating pinned
// host->device copy loop on another stream.  Also prints the host cost of capture + update + launch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o graph_update_ubench graph_update_ubench.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void spin_kernel(long long cycles, int *sink, int tag)
{
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 0 && blockIdx.x == 0 && sink) sink[0] = tag;
}

static const int KPG = 10;      // kernels per graph (one "step")
static const int NG = 40;       // steps

static void enqueue_step(cudaStream_t st, int step, int *sink)
{
    for (int i = 0; i < KPG; ++i)
        spin_kernel<<<8 + (step + i) % 5, 64, 0, st>>>(10000, sink, step * KPG + i);    // ~5 us each
}

int main()
{
    cudaStream_t st, bg;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&bg, cudaStreamNonBlocking));
    int *sink; CK(cudaMalloc(&sink, 256));
    const size_t nbytes = 20u << 20;
    void *h, *d; CK(cudaMallocHost(&h, nbytes)); CK(cudaMalloc(&d, nbytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    cudaGraph_t g; cudaGraphExec_t ex_static, ex_dyn;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed)); enqueue_step(st, 0, sink); CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ex_static, g, 0)); CK(cudaGraphInstantiate(&ex_dyn, g, 0)); CK(cudaGraphDestroy(g));

    for (int rep = 0; rep < 2; ++rep)
    for (int load = 0; load < 2; ++load) {
        for (int mode = 0; mode < 3; ++mode) {
            CK(cudaDeviceSynchronize());
            if (load) for (int i = 0; i < 30; ++i) CK(cudaMemcpyAsync(d, h, nbytes, cudaMemcpyHostToDevice, bg));
            CK(cudaEventRecord(e0, st));
            const auto t0 = std::chrono::steady_clock::now();
            int failed = 0;
            for (int s = 0; s < NG; ++s) {
                if (mode == 0) enqueue_step(st, s, sink);
                else if (mode == 1) CK(cudaGraphLaunch(ex_static, st));
                else {
                    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed)); enqueue_step(st, s, sink); CK(cudaStreamEndCapture(st, &g));
                    cudaGraphExecUpdateResultInfo info;
                    if (cudaGraphExecUpdate(ex_dyn, g, &info) != cudaSuccess) {
                        cudaGetLastError(); ++failed;
                        CK(cudaGraphExecDestroy(ex_dyn)); CK(cudaGraphInstantiate(&ex_dyn, g, 0));
                    }
                    CK(cudaGraphDestroy(g));
                    CK(cudaGraphLaunch(ex_dyn, st));
                }
            }
            const auto t1 = std::chrono::steady_clock::now();
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaDeviceSynchronize());
            const char *names[3] = {"A stream launches", "B static graph", "C re-captured + ExecUpdate graph"};
            printf("%-34s %-10s device %.2f us per kernel   host %.1f us per step of %d kernels   (update failures %d)\n", names[mode],
                   load ? "under H2D" : "alone", ms * 1e3 / (NG * KPG), std::chrono::duration<double, std::micro>(t1 - t0).count() / NG, KPG, failed);
        }
    }
    return 0;
}
