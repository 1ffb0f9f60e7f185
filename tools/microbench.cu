// Pipe-throughput microbenchmarks for the MPJPE kernel's instruction mix (B200).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c){ f2 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ float ffma(float a,float b,float c){ float d; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(d):"f"(a),"f"(b),"f"(c)); return d; }
__device__ __forceinline__ float rsq(float a){ float d; asm volatile("rsqrt.approx.ftz.f32 %0,%1;" : "=f"(d):"f"(a)); return d; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b){ unsigned d; asm volatile("add.u32 %0,%1,%2;" : "=r"(d):"r"(a),"r"(b)); return d; }
__device__ __forceinline__ float fmnmx(float a,float b){ float d; asm volatile("max.f32 %0,%1,%2;" : "=f"(d):"f"(a),"f"(b)); return d; }

// MODE: counts per loop iteration of (ffma2, ffma, mufu, alu)
template <int NF2, int NF1, int NMU, int NAL>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    f2 a2[8]; float a1[8]; float m[8]; unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a2[i] = (f2)(threadIdx.x + i) | ((f2)(i + 1) << 40); a1[i] = seed + i; m[i] = seed * (i + 2); u[i] = i; }
    const f2 c2 = 0x3f8000003f800000ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < NF2; ++i) a2[(r + i) & 7] = fma2(a2[(r + i) & 7], c2, c2);
#pragma unroll
            for (int i = 0; i < NF1; ++i) a1[(r + i) & 7] = ffma(a1[(r + i) & 7], 1.0001f, 0.5f);
#pragma unroll
            for (int i = 0; i < NMU; ++i) m[(r + i) & 7] = rsq(m[(r + i) & 7]);
#pragma unroll
            for (int i = 0; i < NAL; ++i) u[(r + i) & 7] = iadd(u[(r + i) & 7], 0x800000u);
        }
    }
    float acc = 0; 
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += (float)(a2[i] & 0xffff) + a1[i] + m[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NF2, int NF1, int NMU, int NAL>
void run(const char *name, int warps_per_sm)
{
    float *out; cudaMalloc(&out, 148 * 16 * 256 * 4);
    const int iters = 2000;
    const int blocks = 148 * warps_per_sm / 8;
    k<NF2, NF1, NMU, NAL><<<blocks, 256>>>(out, 10, 1.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NF2, NF1, NMU, NAL><<<blocks, 256>>>(out, iters, 1.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions per SMSP per cycle assuming 1965 MHz
    double cyc = ms * 1e-3 * 1.965e9;
    double per_smsp_iters = (double)iters * 8 * (warps_per_sm / 4.0);
    printf("%-28s warps/SM %2d: %.3f ms  cycles per (loop body of %d f2 + %d f1 + %d mufu + %d alu) per SMSP-warp-slot: %.2f\n",
           name, warps_per_sm, ms, NF2, NF1, NMU, NAL, cyc / per_smsp_iters);
    cudaFree(out);
}

int main()
{
    for (int w : {8, 16, 32}) {
        if (w == 8) { run<8,0,0,0>("ffma2 x8", 8); run<0,8,0,0>("ffma x8", 8); run<0,0,8,0>("mufu x8", 8); run<0,0,0,8>("iadd x8", 8);
                      run<8,0,2,0>("ffma2 x8 + mufu x2", 8); run<4,8,0,0>("ffma2 x4 + ffma x8", 8); run<8,0,2,4>("ffma2 x8 + mufu x2 + alu x4", 8);
                      run<0,16,2,0>("ffma x16 + mufu x2", 8); run<8,8,0,0>("ffma2 x8 + ffma x8", 8); }
        if (w == 16) { run<8,0,0,0>("ffma2 x8", 16); run<0,8,0,0>("ffma x8", 16); run<0,0,8,0>("mufu x8", 16); run<0,0,0,8>("iadd x8", 16);
                      run<8,0,2,0>("ffma2 x8 + mufu x2", 16); run<4,8,0,0>("ffma2 x4 + ffma x8", 16); run<8,0,2,4>("ffma2 x8 + mufu x2 + alu x4", 16);
                      run<0,16,2,0>("ffma x16 + mufu x2", 16); run<8,8,0,0>("ffma2 x8 + ffma x8", 16); }
        if (w == 32) { run<8,0,0,0>("ffma2 x8", 32); run<0,8,0,0>("ffma x8", 32); run<0,0,8,0>("mufu x8", 32);
                      run<8,0,2,0>("ffma2 x8 + mufu x2", 32); run<8,0,2,4>("ffma2 x8 + mufu x2 + alu x4", 32); run<0,16,2,0>("ffma x16 + mufu x2", 32); }
    }
    return 0;
}
