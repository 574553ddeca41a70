// Micro-benchmarks behind the design of bcd_blocked.cuh (one CTA, 384 threads, one SM):
//   cost of warp-uniform (broadcast) shared loads, FFMA2 issue rate, SHFL latency with and without shared-memory traffic.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_ubench smem_ubench.cu && ./smem_ubench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// mode 0: LDS.128 broadcast (all lanes same address); 1: LDS.128 distinct (512 B per warp); 2: LDS.32 broadcast;
// 3: LDS.64 distinct (256 B per warp)
template <int MODE>
__global__ void lds_kernel(float *out, long long *cyc, int iters, int nwarps)
{
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (wid < nwarps) {
        t0 = clk();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int base = ((it * 16 + u) * 64 + wid * 128) & 4095;
                if (MODE == 0) { float4 v = *reinterpret_cast<const float4 *>(sm + base); acc += v.x + v.y + v.z + v.w; }
                if (MODE == 1) { float4 v = *reinterpret_cast<const float4 *>(sm + ((base + 4 * lane) & 8188)); acc += v.x + v.y + v.z + v.w; }
                if (MODE == 2) { acc += sm[base]; }
                if (MODE == 3) { float2 v = *reinterpret_cast<const float2 *>(sm + ((base + 2 * lane) & 8190)); acc += v.x + v.y; }
            }
        }
        t1 = clk();
    }
    out[threadIdx.x] = acc;
    if (lane == 0 && wid < nwarps) cyc[wid] = t1 - t0;
}

// FFMA2 with 8 independent accumulators per thread
__global__ void ffma2_kernel(float *out, long long *cyc, int iters, int nwarps)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float2 a[8];
    for (int j = 0; j < 8; ++j) a[j] = make_float2(lane * 0.001f + j, 1.f);
    float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    long long t0 = 0, t1 = 0;
    if (wid < nwarps) {
        t0 = clk();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ffma2_rn(a[j], m, c);
        }
        t1 = clk();
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += a[j].x + a[j].y;
    out[threadIdx.x] = s;
    if (lane == 0 && wid < nwarps) cyc[wid] = t1 - t0;
}

// warp 0: dependent chain of butterfly reductions (5 SHFL + 5 FADD each); warps 1..: background traffic
// bg 0: none; 1: LDS.128 broadcast stream; 2: FFMA2 stream; 3: LDS.128 distinct stream
__global__ void shfl_kernel(float *out, long long *cyc, int iters, int bg, int bgwarps, int solver_wid)
{
    __shared__ __align__(16) float sm[8192];
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-6f;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc = lane * 0.01f;
    if (wid == solver_wid) {
        const long long t0 = clk();
        for (int it = 0; it < iters; ++it) {
            float v = acc * 1.0001f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc = v * 0.03f + lane;
        }
        const long long t1 = clk();
        if (lane == 0) { cyc[0] = t1 - t0; stop = 1; }
    } else if (bg != 0 && (wid < bgwarps + (solver_wid == 0 ? 1 : 0)) && wid != solver_wid) {
        float2 a0 = make_float2(acc, 1.f), a1 = a0, a2 = a0, a3 = a0;
        const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
        int it = 0;
        while (!stop) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int base = ((it * 8 + u) * 64 + wid * 128) & 4095;
                if (bg == 1) { float4 v = *reinterpret_cast<const float4 *>(sm + base); acc += v.x + v.y + v.z + v.w; }
                if (bg == 3) { float4 v = *reinterpret_cast<const float4 *>(sm + ((base + 4 * lane) & 8188)); acc += v.x + v.y + v.z + v.w; }
                if (bg == 2) { a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c); }
            }
            ++it;
        }
        acc += a0.x + a1.y + a2.x + a3.y;
    }
    out[threadIdx.x] = acc;
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 256);
    long long h[16];
    const int iters = 2000;
    const char *names[4] = {"LDS.128 broadcast", "LDS.128 distinct (512 B/warp)", "LDS.32 broadcast", "LDS.64 distinct (256 B/warp)"};
    for (int mode = 0; mode < 4; ++mode)
        for (int nw : {1, 4, 9, 12}) {
            if (mode == 0) lds_kernel<0><<<1, 384>>>(out, cyc, iters, nw);
            if (mode == 1) lds_kernel<1><<<1, 384>>>(out, cyc, iters, nw);
            if (mode == 2) lds_kernel<2><<<1, 384>>>(out, cyc, iters, nw);
            if (mode == 3) lds_kernel<3><<<1, 384>>>(out, cyc, iters, nw);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
            printf("%-32s %2d warps: %.2f cycles per load per warp, %.2f SM cycles per warp-load overall\n", names[mode], nw,
                   (double)mx / (iters * 16), (double)mx / (iters * 16) / nw);
        }
    for (int nw : {1, 4, 8, 12}) {
        ffma2_kernel<<<1, 384>>>(out, cyc, iters, nw);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("FFMA2 x8 independent            %2d warps: %.2f cycles per FFMA2 per warp\n", nw, (double)mx / (iters * 8));
    }
    const char *bgn[4] = {"alone", "8 warps LDS.128 broadcast", "8 warps FFMA2", "8 warps LDS.128 distinct"};
    for (int sw : {0, 11})
        for (int bg = 0; bg < 4; ++bg) {
            shfl_kernel<<<1, 384>>>(out, cyc, 2000, bg, 8, sw);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            printf("butterfly chain (5 SHFL+FADD) on warp %2d, background %-28s: %.1f cycles per reduction\n", sw, bgn[bg], (double)h[0] / 2000);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
