// simhand_b200 K0: all-pairs MPJPE tiles (src/models/utils.py:251-255).
//
// D_ij = mean_k || a_ik - a_jk ||_2 over 21 joints, bit-exact with torch-CPU:
//   n_k = sqrt_rn(fma(dy, dy, dx*dx));  s = ((((n16+n17)+n18)+n19)+n20);  s += (n_k + n_{k+8}), k = 0..7;
//   D = s / 21                                                       (oracle/smh_oracle.c)
// D is bitwise symmetric with a zero diagonal, so only the upper-triangular 128x128 tiles assigned to this
// rank are evaluated and stored (one CTA per tile, layout smh_common.cuh: dist_index); the sweeps read a
// stored tile directly for (I, J) and transposed for (J, I).  The global max (utils.py:255) is folded in
// with an integer atomicMax (D >= 0).  The global min is the diagonal, +0 (utils.py:256).
//
// CUDA-core kernel: per ordered pair 21 MUFU.RSQ and ~190 FMA-pipe lane-operations; the packed
// FADD2/FMUL2/FFMA2 forms halve the issue slots.  This is the kernel the roofline in bench.py is quoted on.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

template <bool FAST>
__device__ __forceinline__ float mpjpe_one(const f2 (&ax)[10], const f2 (&ay)[10], float ax20, float ay20,
                                           const float *__restrict__ col, const DivConst &div21)
{
    f2 nn[10];
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        const float4 b = *reinterpret_cast<const float4 *>(col + 4 * p);   // (bx_2p, bx_2p+1, by_2p, by_2p+1)
        f2 dx = sub2(ax[p], pack2(b.x, b.y));
        f2 dy = sub2(ay[p], pack2(b.z, b.w));
        f2 x = fma2(dy, dy, mul2(dx, dx));
        if (FAST) {
            nn[p] = sqrt2_rn_fast(x);
        } else {
            float x0, x1;
            unpack2(x, x0, x1);
            nn[p] = pack2(__fsqrt_rn(x0), __fsqrt_rn(x1));
        }
    }
    const float2 b20 = *reinterpret_cast<const float2 *>(col + 40);
    const float dx20 = __fsub_rn(ax20, b20.x), dy20 = __fsub_rn(ay20, b20.y);
    const float x20 = __fmaf_rn(dy20, dy20, __fmul_rn(dx20, dx20));
    const float n20 = FAST ? sqrt_rn_fast(x20) : __fsqrt_rn(x20);

    float a, b;
    unpack2(nn[8], a, b);                 // (n16, n17)
    float s = __fadd_rn(a, b);
    unpack2(nn[9], a, b);                 // (n18, n19)
    s = __fadd_rn(s, a);
    s = __fadd_rn(s, b);
    s = __fadd_rn(s, n20);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        f2 t = add2(nn[p], nn[p + 4]);    // (n_2p + n_2p+8, n_2p+1 + n_2p+9)
        unpack2(t, a, b);
        s = __fadd_rn(s, a);
        s = __fadd_rn(s, b);
    }
    return FAST ? div_fast(s, div21) : __fdiv_rn(s, 21.0f);
}

template <bool FAST>
__device__ __forceinline__ void mpjpe_tile_body(const float *__restrict__ jp, float *__restrict__ tile_out, int I,
                                                int J, int m, float *cs, float &vmax)
{
    const int t = threadIdx.x;
    const int r = t & 127;
    const int h = t >> 7;
    // stage the 128 column samples (contiguous 128 x 44 floats)
    {
        const float4 *src = reinterpret_cast<const float4 *>(jp + (int64_t)J * kTile * kJP);
        float4 *dst = reinterpret_cast<float4 *>(cs);
        for (int i = t; i < kTile * kJP / 4; i += 256) dst[i] = src[i];
    }
    // this thread's row sample in registers
    f2 ax[10], ay[10];
    float ax20, ay20;
    {
        const float4 *rowp = reinterpret_cast<const float4 *>(jp + ((int64_t)I * kTile + r) * kJP);
#pragma unroll
        for (int p = 0; p < 10; ++p) {
            float4 v = rowp[p];
            ax[p] = pack2(v.x, v.y);
            ay[p] = pack2(v.z, v.w);
        }
        float4 v = rowp[10];
        ax20 = v.x;
        ay20 = v.y;
    }
    __syncthreads();
    const DivConst div21 = make_div(21.0f);
    const bool row_ok = (I * kTile + r) < m;
    const int col_limit = m - J * kTile;          // columns >= col_limit are padding
#pragma unroll 1
    for (int cq = 0; cq < 16; ++cq) {
        const int c0 = h * 64 + cq * 4;
        float dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            dv[u] = mpjpe_one<FAST>(ax, ay, ax20, ay20, cs + (c0 + u) * kJP, div21);
            if (row_ok && (c0 + u) < col_limit) vmax = fmaxf(vmax, dv[u]);
        }
        *reinterpret_cast<float4 *>(tile_out + dist_index(r, c0)) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
}

__global__ void __launch_bounds__(256, 2)
mpjpe_kernel(const int2 *__restrict__ tiles, const float *__restrict__ jp, float *__restrict__ dist, int m,
             Stats *__restrict__ stats)
{
    __shared__ __align__(16) float cs[kTile * kJP];
    __shared__ float wmax[8];
    const int2 ij = tiles[blockIdx.x];
    float *tile_out = dist + (int64_t)blockIdx.x * kTileFloats;
    float vmax = 0.f;
    const uint32_t flags = stats->flags;
    if (flags & (SMH_FLAG_SLOW_DOMAIN | SMH_FLAG_NONFINITE))
        mpjpe_tile_body<false>(jp, tile_out, ij.x, ij.y, m, cs, vmax);
    else
        mpjpe_tile_body<true>(jp, tile_out, ij.x, ij.y, m, cs, vmax);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = wmax[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = fmaxf(v, wmax[w]);
        atomicMax(&stats->dmax_bits, __float_as_uint(v));
    }
}

int launch_mpjpe(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                 cudaStream_t stream)
{
    (void)dims;
    if (lay.n_stored_tiles == 0) return 0;
    mpjpe_kernel<<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.jp, ws.dist, lay.m, (Stats *)ws.stats);
    return check_launch("mpjpe_kernel");
}

}  // namespace smh
