// simhand_b200: device self-tests of the exact fp32 arithmetic the MPJPE weight path relies on.
// Each test compares the branch-free fast form (smh_common.cuh) with the IEEE intrinsic over its whole
// domain (exhaustive) or a large pseudo-random sample, and returns counters:
//   out[0] = values tested, out[1] = mismatches, out[2] = first mismatching input bits, out[3] = max ulp distance
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

__device__ __forceinline__ void report(uint64_t *out, uint64_t tested, uint64_t bad, uint64_t first_bad, uint64_t max_ulp)
{
    // block-level aggregation is not needed: counters are cheap and mismatches are expected to be zero
    if (tested) atomicAdd((unsigned long long *)&out[0], (unsigned long long)tested);
    if (bad) {
        atomicAdd((unsigned long long *)&out[1], (unsigned long long)bad);
        atomicMin((unsigned long long *)&out[2], (unsigned long long)first_bad);
        atomicMax((unsigned long long *)&out[3], (unsigned long long)max_ulp);
    }
}

__device__ __forceinline__ uint64_t ulp_dist(float a, float b)
{
    long long ia = (long long)(int)__float_as_uint(a), ib = (long long)(int)__float_as_uint(b);
    if (ia < 0) ia = -(ia & 0x7fffffffll);
    if (ib < 0) ib = -(ib & 0x7fffffffll);
    long long dlt = ia - ib;
    return (uint64_t)(dlt < 0 ? -dlt : dlt);
}

// 0: sqrt_rn_fast == __fsqrt_rn for x = 0 and every float in [2^-101, FLT_MAX]
__global__ void selftest_sqrt(uint64_t *out)
{
    const uint64_t lo = 0x0d000000ull, hi = 0x7f7fffffull;
    uint64_t tested = 0, bad = 0, first = ~0ull, mx = 0;
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi + 1; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = (b == hi + 1) ? 0.f : __uint_as_float((uint32_t)b);
        const float a = sqrt_rn_fast(x), r = __fsqrt_rn(x);
        ++tested;
        if (__float_as_uint(a) != __float_as_uint(r)) {
            ++bad;
            if (b < first) first = b;
            uint64_t u = ulp_dist(a, r);
            if (u > mx) mx = u;
        }
    }
    report(out, tested, bad, first, mx);
}

// 1: packed sqrt2_rn_fast == __fsqrt_rn on both lanes (every 7th float of the domain, lanes offset)
__global__ void selftest_sqrt2(uint64_t *out)
{
    const uint64_t lo = 0x0d000000ull, hi = 0x7f7fffffull;
    uint64_t tested = 0, bad = 0, first = ~0ull, mx = 0;
    for (uint64_t b = lo + 7ull * ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x); b + 3 <= hi; b += 7ull * gridDim.x * blockDim.x) {
        const float x0 = __uint_as_float((uint32_t)b), x1 = __uint_as_float((uint32_t)(b + 3));
        float a0, a1;
        unpack2(sqrt2_rn_fast(pack2(x0, x1)), a0, a1);
        const float r0 = __fsqrt_rn(x0), r1 = __fsqrt_rn(x1);
        tested += 2;
        if (__float_as_uint(a0) != __float_as_uint(r0) || __float_as_uint(a1) != __float_as_uint(r1)) {
            ++bad;
            if (b < first) first = b;
            uint64_t u = max(ulp_dist(a0, r0), ulp_dist(a1, r1));
            if (u > mx) mx = u;
        }
    }
    report(out, tested, bad, first, mx);
}

// 2: div_fast(x, 21) == __fdiv_rn(x, 21) for x = 0 and every float in [2^-60, 2^70]
__global__ void selftest_div21(uint64_t *out)
{
    const uint64_t lo = 0x21800000ull /* 2^-60 */, hi = 0x62800000ull /* 2^70 */;
    const DivConst d21 = make_div(21.0f);
    uint64_t tested = 0, bad = 0, first = ~0ull, mx = 0;
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi + 1; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = (b == hi + 1) ? 0.f : __uint_as_float((uint32_t)b);
        const float a = div_fast(x, d21), r = __fdiv_rn(x, 21.0f);
        ++tested;
        if (__float_as_uint(a) != __float_as_uint(r)) {
            ++bad;
            if (b < first) first = b;
            uint64_t u = ulp_dist(a, r);
            if (u > mx) mx = u;
        }
    }
    report(out, tested, bad, first, mx);
}

__device__ __forceinline__ uint32_t mix32(uint64_t &s)
{
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t x = (uint32_t)(s >> 33) ^ (uint32_t)(s >> 13);
    return x * 2654435761u;
}

// 3: weight division (Dmax - D) / Dmax: div_fast vs __fdiv_rn on pseudo-random (D, Dmax), 2^34 samples
__global__ void selftest_divw(uint64_t *out)
{
    uint64_t seed = 0x9e3779b97f4a7c15ull * (1 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
    uint64_t tested = 0, bad = 0, first = ~0ull, mx = 0;
    for (int outer = 0; outer < 64; ++outer) {
        // Dmax: random mantissa, exponent in 2^-10 .. 2^20
        const uint32_t eb = 117u + (mix32(seed) % 31u);
        const float c = __uint_as_float((eb << 23) | (mix32(seed) & 0x7fffffu));
        const DivConst dc = make_div(c);
        for (int inner = 0; inner < 1024; ++inner) {
            // D uniform in [0, Dmax] with a random low mantissa, as MPJPE values are
            const float u = (float)(mix32(seed) >> 8) * (1.0f / 16777216.0f);
            float dval = u * c;
            dval = __uint_as_float(__float_as_uint(dval) ^ (mix32(seed) & 0xffu));
            if (!(dval <= c)) dval = c;
            const float num = __fsub_rn(c, dval);
            const float a = div_fast(num, dc), r = __fdiv_rn(num, c);
            ++tested;
            if (__float_as_uint(a) != __float_as_uint(r)) {
                ++bad;
                uint64_t key = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(dval);
                if (key < first) first = key;
                uint64_t ud = ulp_dist(a, r);
                if (ud > mx) mx = ud;
            }
        }
    }
    report(out, tested, bad, first, mx);
}

// 4: sqrt_fma_pipe / sqrt2_fma_pipe (the MUFU-free square root of the 16-bit tile image) against the correctly rounded
// double sqrt for x = 0 and every float in [2^-101, FLT_MAX]: out[1] = values off by more than 7.5e-7 relative (8.0e-7 for
// the accumulating form sqrt2_fma_pipe_acc; or a
// non-zero result for x = 0, or the two forms disagreeing), out[3] = largest relative error in units of 1e-9
__global__ void selftest_sqrt_fma_pipe(uint64_t *out)
{
    const uint64_t lo = 0x0d000000ull, hi = 0x7f7fffffull;
    uint64_t tested = 0, bad = 0, first = ~0ull, mx = 0;
    for (uint64_t b = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= hi + 1; b += (uint64_t)gridDim.x * blockDim.x) {
        const float x = (b == hi + 1) ? 0.f : __uint_as_float((uint32_t)b);
        const float a = sqrt_fma_pipe(x);
        float p0, p1, q0, q1;
        unpack2(sqrt2_fma_pipe(pack2(x, x)), p0, p1);
        unpack2(sqrt2_fma_pipe_acc(pack2(x, x), pack2(0.f, 0.f)), q0, q1);        // the accumulating form, same bound
        const double r = sqrt((double)x);
        const double rel = r > 0.0 ? fabs((double)a - r) / r : (a == 0.f ? 0.0 : 1.0);
        const uint64_t u = (uint64_t)(rel * 1e9);
        ++tested;
        if (u > mx) mx = u;
        const double relq = r > 0.0 ? fabs((double)q0 - r) / r : (q0 == 0.f ? 0.0 : 1.0);
        if (rel > 7.5e-7 || relq > 8.0e-7 || q0 != q1 || __float_as_uint(p0) != __float_as_uint(a) ||
            __float_as_uint(p1) != __float_as_uint(a)) {
            ++bad;
            if (b < first) first = b;
        }
    }
    if (tested) atomicAdd((unsigned long long *)&out[0], (unsigned long long)tested);
    if (bad) {
        atomicAdd((unsigned long long *)&out[1], (unsigned long long)bad);
        atomicMin((unsigned long long *)&out[2], (unsigned long long)first);
    }
    atomicMax((unsigned long long *)&out[3], (unsigned long long)mx);
}

int launch_selftest(int which, uint64_t *out, int64_t out_words, cudaStream_t stream)
{
    (void)out_words;
    static const uint64_t init[8] = {0, 0, ~0ull, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(out, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return set_error((int)e, "selftest init: %s", cudaGetErrorString(e));
    switch (which) {
        case 0: selftest_sqrt<<<148 * 16, 256, 0, stream>>>(out); break;
        case 1: selftest_sqrt2<<<148 * 16, 256, 0, stream>>>(out); break;
        case 2: selftest_div21<<<148 * 16, 256, 0, stream>>>(out); break;
        case 3: selftest_divw<<<148 * 16, 256, 0, stream>>>(out); break;
        case 4: selftest_sqrt_fma_pipe<<<148 * 16, 256, 0, stream>>>(out); break;
        default: return set_error(SMH_E_MODE, "unknown selftest %d", which);
    }
    return check_launch("selftest");
}

}  // namespace smh

// ----------------------------------------------------------------------------------------------
// tcgen05 probe: one CTA computes S = A B^T (kind::tf32, both operands K-major SWIZZLE_128B from the staged
// tf32 z blocks) and then dZ = bf16(S) Z_B (kind::f16: A operand = packed bf16 in TMEM, B operand = the staged
// bf16 block read MN-major), with every descriptor field supplied by the caller.  tests/ uses it to pin the
// descriptor encodings the sweep kernels hard-code (smh_sweep_tc.cu) against a host matmul.
// params: [0] idesc1 [1] a_lbo [2] a_sbo [3] b_lbo [4] b_sbo [5] kstep_bytes [6] box_stride_a [7] box_stride_b
//         [8] idesc2 [9] b2_lbo [10] b2_sbo [11] b2_kstep_bytes [12] a2 TMEM columns per K step
// ----------------------------------------------------------------------------------------------
namespace smh {

struct ProbeParams {
    uint32_t v[16];
};

__global__ void __launch_bounds__(128, 1)
tc_probe_kernel(const float *__restrict__ zt, const uint16_t *__restrict__ zb, int blk_a, int blk_b, ProbeParams p,
                float *__restrict__ s_out, uint32_t *__restrict__ g_out, float *__restrict__ dz_out,
                uint32_t *__restrict__ fail)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = sm, *sB = sm + 65536, *sBb = sm + 65536 + 32768;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 65536 + 32768 + 16384);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tc_alloc(tmem_slot, 256);
        tc_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], 65536 + 32768 + 16384);
        for (int kb = 0; kb < 4; ++kb) {
            bulk_g2s(sA + kb * 16384, zt + (int64_t)blk_a * kBlockFloats + kb * 2048, 8192, &bars[0]);
            bulk_g2s(sA + kb * 16384 + 8192, zt + (int64_t)(blk_a + 1) * kBlockFloats + kb * 2048, 8192, &bars[0]);
        }
        bulk_g2s(sB, zt + (int64_t)blk_b * kBlockFloats, 32768, &bars[0]);
        bulk_g2s(sBb, zb + (int64_t)blk_b * kBlockFloats, 16384, &bars[0]);
        mbar_wait(&bars[0], 0, fail, 101);
        tc_fence_after();
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        for (int kb = 0; kb < 4; ++kb)
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = umma_desc_sw128(sA_u + kb * p.v[6] + ks * p.v[5], p.v[1], p.v[2]);
                const uint64_t bd = umma_desc_sw128(sB_u + kb * p.v[7] + ks * p.v[5], p.v[3], p.v[4]);
                tc_mma_ss_tf32(tmem_base, ad, bd, p.v[0], (kb | ks) ? 1u : 0u);
            }
        tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0, fail, 102);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t v[32], pk[16];
        tc_ld32(lane_addr + chunk * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) s_out[r * 64 + chunk * 32 + c] = __uint_as_float(v[c]);
#pragma unroll
        for (int c = 0; c < 16; ++c) pk[c] = pack_bf16x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
        tc_st16(lane_addr + chunk * 16, pk);
    }
    tc_wait_st();
    {   // read G back (checks that tcgen05.st landed where the MMA will look)
        uint32_t v[32];
        tc_ld32(lane_addr, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) g_out[r * 32 + c] = v[c];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t sBb_u = smem_u32(sBb);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t bd = umma_desc_sw128(sBb_u + ks * p.v[11], p.v[9], p.v[10]);
            tc_mma_ts_f16(tmem_base + 128, tmem_base + ks * p.v[12], bd, p.v[8], ks ? 1u : 0u);
        }
        tc_commit(&bars[2]);
    }
    mbar_wait(&bars[2], 0, fail, 103);
    tc_fence_after();
    for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t v[32];
        tc_ld32(lane_addr + 128 + chunk * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) dz_out[r * 128 + chunk * 32 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tmem_base, 256);
}

}  // namespace smh

extern "C" int smh_tc_probe(const float *zt_dev, const void *zb_dev, int blk_a, int blk_b, const uint32_t *params16_host,
                            float *s_out_dev, uint32_t *g_out_dev, float *dz_out_dev, uint32_t *fail_dev, void *stream)
{
    using namespace smh;
    if (!zt_dev || !zb_dev || !params16_host || !s_out_dev || !g_out_dev || !dz_out_dev || !fail_dev)
        return set_error(SMH_E_ARG, "null pointer");
    ProbeParams p;
    for (int i = 0; i < 16; ++i) p.v[i] = params16_host[i];
    const int smem = 1024 + 65536 + 32768 + 16384 + 64;
    cudaError_t e = cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error((int)e, "probe smem attr: %s", cudaGetErrorString(e));
    tc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(zt_dev, (const uint16_t *)zb_dev, blk_a, blk_b, p, s_out_dev,
                                                          g_out_dev, dz_out_dev, fail_dev);
    return check_launch("tc_probe_kernel");
}

extern "C" void smh_tc_default_params(uint32_t *params16_host)
{
    using namespace smh;
    for (int i = 0; i < 16; ++i) params16_host[i] = 0;
    params16_host[0] = umma_idesc_tf32(kTile, kTaskN, 0, 0);
    params16_host[1] = 16;      // a_lbo (unused for swizzled K-major; encoded as 1)
    params16_host[2] = 1024;    // a_sbo: 8 rows x 128 B
    params16_host[3] = 16;
    params16_host[4] = 1024;
    params16_host[5] = 32;      // 8 tf32 = 32 B per K step inside the 128 B swizzle row
    params16_host[6] = 16384;   // A: next 32-column box (128 rows x 128 B)
    params16_host[7] = 8192;    // B: next 32-column box (64 rows x 128 B)
    params16_host[8] = umma_idesc_bf16(kTile, kD, 0, 1);
    params16_host[9] = 8192;    // MN-major bf16 B: stride between 64-element (128 B) MN atoms
    params16_host[10] = 1024;   // stride between 8-row K groups
    params16_host[11] = 2048;   // K step of 16 sample rows
    params16_host[12] = 8;      // A (TMEM) columns per K step: 16 packed bf16
}
