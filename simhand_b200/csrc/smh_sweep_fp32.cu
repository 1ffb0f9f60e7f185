// simhand_b200 K1/K2, exact-fp32 engine (SMH_ENGINE_FP32): the forward and backward sweeps with the dense
// contractions on the CUDA cores (FFMA, fp32 accumulate).  Same task plan, same MPJPE tile reads and the
// same epilogue arithmetic as the tcgen05 engine (smh_sweep_tc.cu); used when bit-faithful fp32 logits are
// wanted and as the on-device cross-check of the tensor-core engine.
//
//   forward  (src/models/utils.py:411-417): S = z z^T, E = exp(S * W / tau) off the diagonal, neg_i += sum_j E_ij
//   backward (autograd of :411-426): G_ij = W_ij E_ij (1/neg_i + 1/neg_j), dzacc_i += sum_j G_ij z_j
//
// One CTA (256 threads) walks strips of tasks that share a 128-row block; a task is 128 rows x 64 columns.
// Thread (ty, tx) = (t / 16, t % 16) owns rows ty + 16 p and columns tx + 16 q, which makes every shared-memory
// access of the two contractions a broadcast or a unit-stride read with a 129-float pitch.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

constexpr int kPitch = kD + 1;                                            // 129 floats
constexpr int kFp32Smem = (kTile * kPitch + 2 * kTaskN * kPitch) * 4;     // As + Zs + Gs = 132096 B

template <bool BWD>
__global__ void __launch_bounds__(256, 1)
sweep_fp32_kernel(const int4 *__restrict__ tasks, const int2 *__restrict__ strips, const int *__restrict__ cta_ptr,
                  const float *__restrict__ zt, const float *__restrict__ dist, const float *__restrict__ rn,
                  const __grid_constant__ Peers peers, const Stats *__restrict__ stats, int m, int n, int n_local, float k2, int wmode, float lambda_neg)
{
    extern __shared__ __align__(16) float smem[];
    float *As = smem;                         // [128][129]  row block of z
    float *Zs = As + kTile * kPitch;          // [64][129]   column block of z
    float *Gs = Zs + kTaskN * kPitch;         // [64][129]   G^T: Gs[j][i]

    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int s_begin = cta_ptr[blockIdx.x], s_end = cta_ptr[blockIdx.x + 1];      // this CTA's strips
    const float dmax = __uint_as_float(stats->dmax_bits);
    const DivConst divw = make_div(dmax);     // Dmax - Dmin with Dmin = +0 (diagonal)
    const float mu = wmode == 3 ? (float)(stats->dsum / ((double)m * (double)m)) : 0.f;     // non_linear: mean D

    for (int s = s_begin; s < s_end; ++s) {
        const int2 strip = strips[s];
        const int I = tasks[strip.x].x;
        __syncthreads();
        for (int idx = t; idx < kTile * kD; idx += 256) {
            int r = idx >> 7, c = idx & 127;
            As[r * kPitch + c] = zt[zt_index((int64_t)I * kTile + r, c)];
        }
        float rowsum[8];
        float dz[8][8];
        float rni[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            rowsum[p] = 0.f;
            int gi = I * kTile + ty + 16 * p;
            rni[p] = (BWD && gi < m) ? rn[gi] : 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) dz[p][q] = 0.f;
        }

        for (int ti = strip.x; ti < strip.y; ++ti) {
            const int4 task = tasks[ti];
            const int cj = task.y;
            const float *tile = dist + (int64_t)task.z * kTileFloats;
            const bool transposed = task.w & kTaskTransposed;
            const bool diagonal = task.w & kTaskDiagonal;
            __syncthreads();                              // previous task done with Zs / Gs
            for (int idx = t; idx < kTaskN * kD; idx += 256) {
                int r = idx >> 7, c = idx & 127;
                Zs[r * kPitch + c] = zt[zt_index((int64_t)cj * kTaskN + r, c)];
            }
            __syncthreads();
            // S micro-tile
            float acc[8][4];
#pragma unroll
            for (int p = 0; p < 8; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
#pragma unroll 4
            for (int k = 0; k < kD; ++k) {
                float a[8], b[4];
#pragma unroll
                for (int p = 0; p < 8; ++p) a[p] = As[(ty + 16 * p) * kPitch + k];
#pragma unroll
                for (int q = 0; q < 4; ++q) b[q] = Zs[(tx + 16 * q) * kPitch + k];
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
            }
            // epilogue
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int r = ty + 16 * p;
                const int gi = I * kTile + r;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int jl = tx + 16 * q;
                    const int gj = cj * kTaskN + jl;
                    const int cc = (cj & 1) * 64 + jl;            // column inside the 128-wide stored tile
                    // wmode 1: unit weights, the tile is never read; 2: the tile holds the materialised W
                    const float dv = wmode == 1 ? 0.f : (transposed ? tile[dist_index(cc, r)] : tile[dist_index(r, cc)]);
                    float w = wmode == 1 ? 1.0f : (wmode == 2 ? dv : div_fast(__fsub_rn(dmax, dv), divw));
                    if (wmode == 3) w = __fdiv_rn(1.0f, 1.0f + expf(lambda_neg * (dv - mu)));     // utils.py:346
                    float e = ex2_approx(acc[p][q] * w * k2);
                    const bool valid = (gi < m) && (gj < m) && !(diagonal && gi == gj);
                    e = valid ? e : 0.f;
                    if (!BWD) {
                        rowsum[p] += e;
                    } else {
                        const float rnj = (gj < m) ? rn[gj] : 0.f;
                        // materialised W is not symmetric in general: row term from the direct visit of the tile,
                        // column term from the transposed visit of its mirror
                        const float rs = wmode == 2 ? (transposed ? rnj : rni[p]) : rni[p] + rnj;
                        Gs[jl * kPitch + r] = valid ? w * e * rs : 0.f;
                    }
                }
            }
            if (BWD) {
                __syncthreads();
#pragma unroll 2
                for (int j = 0; j < kTaskN; ++j) {
                    float g[8], zv[8];
#pragma unroll
                    for (int p = 0; p < 8; ++p) g[p] = Gs[j * kPitch + ty + 16 * p];
#pragma unroll
                    for (int q = 0; q < 8; ++q) zv[q] = Zs[j * kPitch + tx + 16 * q];
#pragma unroll
                    for (int p = 0; p < 8; ++p)
#pragma unroll
                        for (int q = 0; q < 8; ++q) dz[p][q] = fmaf(g[p], zv[q], dz[p][q]);
                }
            }
        }
        // strip flush
        if (!BWD) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                float v = rowsum[p];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                const int gi = I * kTile + ty + 16 * p;
                if (tx == 0 && gi < m)
                    for (int pp = 0; pp < peers.world; ++pp) atomicAdd(peers.neg(pp) + gi, v);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int gi = I * kTile + ty + 16 * p;
                if (gi < m) {
                    float *orow = dz_row_ptr(peers, gi, n, n_local);
#pragma unroll
                    for (int q = 0; q < 8; ++q) atomicAdd(orow + tx + 16 * q, dz[p][q]);
                }
            }
        }
    }
}

int launch_sweep_fp32(bool backward, int wmode, const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan,
                      const WsView &ws, const Peers &peers, float temperature, cudaStream_t stream)
{
    if (lay.n_strips == 0) return 0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)sms;
    const int grid = kNumCtas;                      // the plan is cut for exactly this many CTAs
    const float k2 = 1.4426950408889634f / temperature;
    const int n_local = dims.n / dims.world;
    cudaError_t e;
    if (backward) {
        e = cudaFuncSetAttribute(sweep_fp32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFp32Smem);
        if (e != cudaSuccess) return set_error((int)e, "fp32 sweep smem attr: %s", cudaGetErrorString(e));
        sweep_fp32_kernel<true><<<grid, 256, kFp32Smem, stream>>>(plan.tasks, plan.strips, plan.cta_ptr, ws.zt,
                                                                 ws.dist, ws.rn, peers,
                                                                 (const Stats *)ws.stats, lay.m, dims.n, n_local, k2, wmode, dims.lambda_neg);
    } else {
        e = cudaFuncSetAttribute(sweep_fp32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFp32Smem);
        if (e != cudaSuccess) return set_error((int)e, "fp32 sweep smem attr: %s", cudaGetErrorString(e));
        sweep_fp32_kernel<false><<<grid, 256, kFp32Smem, stream>>>(plan.tasks, plan.strips_fwd, plan.cta_ptr_fwd, ws.zt,
                                                                  ws.dist, ws.rn, peers,
                                                                  (const Stats *)ws.stats, lay.m, dims.n, n_local, k2, wmode, dims.lambda_neg);
    }
    return check_launch("sweep_fp32_kernel");
}

}  // namespace smh
