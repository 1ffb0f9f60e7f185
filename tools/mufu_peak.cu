// Measured throughput of the special-function (XU / MUFU) pipe of the GPU this runs on: the denominator of the roofline
// bench.py quotes for the MPJPE kernel (21 square roots per pair) and of the sweeps' ex2 floor.  Chip-wide Gop/s for
// sqrt.approx (MUFU.SQRT), rsqrt.approx (MUFU.RSQ) and ex2.approx (MUFU.EX2), 8 independent chains per thread, 32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_peak tools/mufu_peak.cu && /tmp/mufu_peak > profiles/r02_mufu_peak.json
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float mufu(float x)
{
    float r;
    if (OP == 0) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <int OP>
__global__ void __launch_bounds__(256) chain(float *out, int iters, float seed)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + 0.125f * i + 1e-3f * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = mufu<OP>(v[i]);           // sqrt / rsqrt iterate towards 1, ex2 towards its fixed range
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
double run(int sms)
{
    float *out;
    const int blocks = sms * 4;                       // 4 x 256 threads = 32 warps per SM
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    const int iters = 4000;
    chain<OP><<<blocks, 256>>>(out, 50, OP == 2 ? -1.0f : 1.5f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        chain<OP><<<blocks, 256>>>(out, iters, OP == 2 ? -1.0f : 1.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * 256 * iters * 32.0;
        const double gops = ops / (ms * 1e-3) / 1e9;
        if (gops > best) best = gops;
    }
    cudaFree(out);
    return best;
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double sq = run<0>(sms), rs = run<1>(sms), ex = run<2>(sms);
    const double per_clk = sq * 1e9 / (sms * (khz * 1e3));
    printf("{\"sms\": %d, \"clock_rate_mhz\": %.0f, \"mufu_sqrt_gops\": %.1f, \"mufu_rsq_gops\": %.1f, \"mufu_ex2_gops\": %.1f, "
           "\"sqrt_lanes_per_clk_per_sm_at_max_clock\": %.2f, \"nominal_gops_at_max_clock\": %.1f, "
           "\"how\": \"8 independent MUFU chains per thread, 32 warps per SM, best of 5 launches, CUDA events\"}\n",
           sms, khz / 1e3, sq, rs, ex, per_clk, sms * 16.0 * khz * 1e3 / 1e9);
    return 0;
}
