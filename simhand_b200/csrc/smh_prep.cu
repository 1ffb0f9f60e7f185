// simhand_b200 K3a: input staging.
//
// One warp per sample row.  Replaces the torch.cat calls of src/models/utils.py:237-239 and :407 and the
// positive-pair MPJPE of :229-231 (the weight normalisation of :233-235 happens in smh_finalize once the
// global max/min are known).  Writes
//   zt  : z rounded to tf32 (round-to-nearest) in the pre-swizzled 64-row block layout the sweeps stage
//         with one linear bulk copy (smh_common.cuh: zt_index)
//   zb  : bf16 copy of z in the pre-swizzled block layout the backward sweep reads MN-major (zb_index)
//   zh  : fp16 copy of z, same layout (forward logit operand of the fp16 engine)
//   jp  : joints packed as 10 x (x_k, x_k+1, y_k, y_k+1) + (x_20, y_20, 0, 0) per sample
//   posd: D_{k,k+N}, with the exact operation order of the all-pairs kernel, so posd[k] is bitwise
//         D[k, k+N]
// and folds the fast-domain check of the exact sqrt into stats.flags.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

__device__ __forceinline__ const float *sample_ptr(const float *base, int k, int n_local, int64_t rank_stride,
                                                   int64_t row_stride)
{
    return base + (int64_t)(k / n_local) * rank_stride + (int64_t)(k % n_local) * row_stride;
}

__global__ void __launch_bounds__(256) prep_kernel(smh_inputs_t in, int n, int d, int mp, bool round_tf32, int diff, int images,
                                                   float *__restrict__ zt, uint16_t *__restrict__ zb,
                                                   uint16_t *__restrict__ zh,
                                                   float *__restrict__ jp, float *__restrict__ posd,
                                                   Stats *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int m = 2 * n;
    const bool has_joints = in.j1_dev != nullptr;         // optional on the materialised-weights path
    uint32_t bound_bits = 0u;                             // running max of D(row, sample 0) over this warp's rows
    // positive-pair extrema of this warp's rows (lane 0): combined per block, two atomics per block instead of two per row
    // (16384 atomics on two addresses serialised in the L2 were most of this kernel's time)
    uint32_t pmax_w = 0u, pmin_inv_w = 0u;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < mp; row += gridDim.x * warps_per_block) {
        float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
        float jx = 0.f, jy = 0.f;
        const bool live = row < m;
        const int v = row >= n ? 1 : 0;
        const int k = row - v * n;
        if (live) {
            const float *zp = sample_ptr(v ? in.z2_dev : in.z1_dev, k, in.n_local, in.z_rank_stride, in.z_row_stride);
            const int c = 4 * lane;
            if (c + 3 < d && ((reinterpret_cast<uintptr_t>(zp + c) & 15) == 0)) {
                zv = *reinterpret_cast<const float4 *>(zp + c);
            } else {
                if (c + 0 < d) zv.x = zp[c + 0];
                if (c + 1 < d) zv.y = zp[c + 1];
                if (c + 2 < d) zv.z = zp[c + 2];
                if (c + 3 < d) zv.w = zp[c + 3];
            }
            if (round_tf32) {
                zv.x = to_tf32(zv.x);
                zv.y = to_tf32(zv.y);
                zv.z = to_tf32(zv.z);
                zv.w = to_tf32(zv.w);
            }
            if (has_joints && lane < kJ) {
                const float *jb = sample_ptr(v ? in.j2_dev : in.j1_dev, k, in.n_local, in.j_rank_stride,
                                             in.j_sample_stride) +
                                  (int64_t)lane * in.j_joint_stride;
                jx = jb[0];
                jy = jb[in.j_coord_stride];
            }
        }
        // only the images the selected engine stages are written (bit 0: fp32/tf32, 1: bf16, 2: fp16)
        if (images & 1) *reinterpret_cast<float4 *>(zt + zt_index(row, 4 * lane)) = zv;
        // bf16 copy (value operand of dz += G z): 4 consecutive columns = 8 bytes inside one 16-byte chunk
        if (images & 2)
            *reinterpret_cast<uint2 *>(zb + zb_index(row, 4 * lane)) = make_uint2(pack_bf16x2(zv.x, zv.y), pack_bf16x2(zv.z, zv.w));
        // fp16 copy, same layout: 11-bit significand = the precision of tf32 for |z| <= 1 (forward logit operand)
        if (images & 4)
            *reinterpret_cast<uint2 *>(zh + zb_index(row, 4 * lane)) = make_uint2(pack_f16x2(zv.x, zv.y), pack_f16x2(zv.z, zv.w));

        // packed joints
        float *jrow = jp + (int64_t)row * kJP;
        if (lane < 20) {
            int p = lane >> 1, u = lane & 1;
            jrow[4 * p + u] = jx;
            jrow[4 * p + 2 + u] = jy;
        } else if (lane == 20) {
            jrow[40] = jx;
            jrow[41] = jy;
        } else if (lane == 21) {
            jrow[42] = 0.f;
            jrow[43] = 0.f;
        }

        // domain of the branch-free exact sqrt: every coordinate is 0 or 2^-24 <= |c| <= 2^60, finite
        uint32_t bad = 0;
        if (live && lane < kJ) {
            float ax = fabsf(jx), ay = fabsf(jy);
            bool fin = (ax <= 3.0e38f) && (ay <= 3.0e38f);   // false for NaN/Inf
            bool okx = (ax == 0.f) || (ax >= 5.9604645e-8f && ax <= 1.1529215e18f);
            bool oky = (ay == 0.f) || (ay >= 5.9604645e-8f && ay <= 1.1529215e18f);
            bad = (fin ? 0u : SMH_FLAG_NONFINITE) | ((okx && oky) ? 0u : SMH_FLAG_SLOW_DOMAIN);
        }
        bad |= __shfl_xor_sync(0xffffffffu, bad, 16);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 8);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 4);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 2);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
        if (bad && lane == 0) atomicOr(&stats->flags, bad);

        // D(row, sample 0): by the triangle inequality D_ij <= D_i0 + D_0j, so 2 max_i D_i0 bounds the whole matrix
        // (scale of the 16-bit tile image, SMH_DIMS_Q16_TILES; any summation order will do)
        if (has_joints && live) {
            float nk = 0.f;
            if (lane < kJ) {
                const float *rb = sample_ptr(in.j1_dev, 0, in.n_local, in.j_rank_stride, in.j_sample_stride) +
                                  (int64_t)lane * in.j_joint_stride;
                const float ex = jx - rb[0], ey = jy - rb[in.j_coord_stride];
                nk = sqrtf(fmaf(ey, ey, ex * ex));
            }
            const float d0 = warp_sum(nk) * (1.0f / 21.0f) * 1.0001f;        // + margin for the rounding of this sum
            if (d0 >= 0.f && d0 <= 3.0e38f) bound_bits = max(bound_bits, __float_as_uint(d0));
        }

        // positive pair (k, k + N): utils.py:229-231, IEEE sqrt/div (any domain), ATen summation order
        if (has_joints && row < n) {
            float dx = 0.f, dy = 0.f;
            if (lane < kJ) {
                const float *pb = sample_ptr(in.j2_dev, k, in.n_local, in.j_rank_stride, in.j_sample_stride) +
                                  (int64_t)lane * in.j_joint_stride;
                dx = __fsub_rn(jx, pb[0]);
                dy = __fsub_rn(jy, pb[in.j_coord_stride]);
            }
            float dk;
            if (diff == SMH_DIFF_MPJPE) {
                const float nk = lane < kJ ? __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))) : 0.f;
                float s = __shfl_sync(0xffffffffu, nk, 16);
#pragma unroll
                for (int q = 17; q <= 20; ++q) s = __fadd_rn(s, __shfl_sync(0xffffffffu, nk, q));
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    s = __fadd_rn(s, __fadd_rn(__shfl_sync(0xffffffffu, nk, q), __shfl_sync(0xffffffffu, nk, q + 8)));
                dk = __fdiv_rn(s, 21.0f);
            } else if (diff == SMH_DIFF_EUCLID) {
                dk = __fsqrt_rn(warp_sum(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))));       // utils.py:265-274: || p1 - p2 ||_2
            } else {
                // w_abs / w_o_abs (utils.py:219-227): per-coordinate mean over the joints, then the 2-norm
                if (diff == SMH_DIFF_W_ABS) {
                    dx = fabsf(dx);
                    dy = fabsf(dy);
                }
                const float mx = __fdiv_rn(warp_sum(dx), 21.0f), my = __fdiv_rn(warp_sum(dy), 21.0f);
                dk = __fsqrt_rn(__fmaf_rn(my, my, __fmul_rn(mx, mx)));
            }
            if (lane == 0) {
                posd[k] = dk;
                uint32_t b = __float_as_uint(dk);
                if (b <= 0x7f800000u) {     // non-negative, not NaN
                    pmax_w = max(pmax_w, b);
                    pmin_inv_w = max(pmin_inv_w, 0x7fffffffu - b);
                } else {
                    atomicOr(&stats->flags, SMH_FLAG_NONFINITE);
                }
            }
        }
    }
    // one atomic per block and quantity (distance bound, positive-pair max, inverted positive-pair min; 0 = nothing seen)
    __shared__ uint32_t wb[3][8];
    if (lane == 0) {
        wb[0][threadIdx.x >> 5] = bound_bits;
        wb[1][threadIdx.x >> 5] = pmax_w;
        wb[2][threadIdx.x >> 5] = pmin_inv_w;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t v = wb[threadIdx.x][0];
        for (int w = 1; w < warps_per_block; ++w) v = max(v, wb[threadIdx.x][w]);
        uint32_t *dst = threadIdx.x == 0 ? &stats->dbound_bits : (threadIdx.x == 1 ? &stats->pmax_bits : &stats->pmin_inv);
        if (v != 0u) atomicMax(dst, v);
    }
}

int launch_prep(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                int engine, cudaStream_t stream)
{
    const bool round_tf32 = engine == SMH_ENGINE_TC_TF32;
    // fp32 engine: the fp32 image only; tensor-core engines: bf16 (backward, and the bf16 forward) + their forward image
    const int images = engine == SMH_ENGINE_FP32 ? 1 : (2 | (engine == SMH_ENGINE_TC_TF32 ? 1 : 0) |
                                                         (engine == SMH_ENGINE_TC_FP16 ? 4 : 0));
    const int mp = lay.tiles_per_side * kTile;
    const int blocks = (mp + 7) / 8;
    prep_kernel<<<blocks, 256, 0, stream>>>(in, dims.n, dims.d, mp, round_tf32, dims.diff_type, images, ws.zt, ws.zb, ws.zh, ws.jp, ws.posd, (Stats *)ws.stats);
    return check_launch("prep_kernel");
}

}  // namespace smh
