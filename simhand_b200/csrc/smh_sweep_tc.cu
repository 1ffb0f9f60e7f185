// simhand_b200 K1/K2, tensor-core engine (SMH_ENGINE_TC_TF32): the fused forward / backward sweeps.
//
//   forward  (src/models/utils.py:411-417): S = z z^T on tcgen05 (kind::tf32, fp32 accumulate in TMEM); the
//            epilogue turns each S tile into E = exp(S * W / tau) with W built on the fly from the stored MPJPE
//            tile (W = (Dmax - D) / Dmax, correctly rounded), masks the diagonal and accumulates the row sums.
//            The 2N x 2N logit matrix never leaves the SM.
//   backward (autograd of :411-426): the same S tile and weights; the epilogue writes
//            G = W E (1/neg_i + 1/neg_j) back into the TMEM columns S came from (packed bf16) and a second
//            tcgen05.mma (kind::f16: A = G from TMEM, B = the staged bf16 z block read MN-major; tf32 operands
//            cannot be read MN-major from a SWIZZLE_128B image) accumulates dzacc_I += G z_J in fp32 in TMEM
//            across the whole strip; the softmax is never materialised.
//
// Persistent CTAs (one per SM) walk strips of 128x64 tasks that share a 128-row block.  Warp roles:
//   warp 0      producer: cp.async.bulk of the z blocks (pre-swizzled SWIZZLE_128B images, smh_prep.cu) and of
//               the MPJPE tile pieces, completing on mbarriers
//   warp 1      one elected thread issues tcgen05.mma and tcgen05.commit
//   warps 2..5  epilogue, one row per thread (TMEM lane == row): tcgen05.ld -> weights -> ex2 -> row sums
//               (forward) or tcgen05.st of G (backward); strip flush of dzacc with red.global.add.v4.f32
// Every pipeline wait is bounded (smh_common.cuh: mbar_wait) so a protocol bug cannot hang the device.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

constexpr int kTcThreads = 192;
constexpr int kStages = 2;
constexpr int kABytes = kTile * kD * 4;                 // 65536
constexpr int kBBytes = kTaskN * kD * 4;                // 32768  tf32 block (logit operand)
constexpr int kBbBytes = kTaskN * kD * 2;               // 16384  bf16 block (value operand, backward only)
constexpr int kPieceBytes = 1024;                       // 64 rows x 16 B of a stored tile
constexpr int kPiecePitch = 1040;                       // +16 B: conflict-free transposed reads
constexpr int kDBytes = 32 * kPiecePitch;               // 33280
constexpr int kNumBars = 24;
constexpr int kTcSmemFwd = 1024 /*align slack*/ + kABytes + kStages * kBBytes + kStages * kDBytes + kNumBars * 8 + 16;
constexpr int kTcSmemBwd = kTcSmemFwd + kStages * kBbBytes;
static_assert(kTcSmemBwd <= 232448, "backward sweep exceeds the 227 KB shared-memory limit");

struct TcBars {
    uint64_t full_b[kStages], empty_b[kStages], full_d[kStages], empty_d[kStages];
    uint64_t a_full, a_empty;
    uint64_t sg_full[2], sg_empty[2], g_ready[2];
    uint64_t dz_full, dz_empty;
};
static_assert(sizeof(TcBars) <= kNumBars * 8, "barrier block too small");

// one 32-column chunk of the epilogue.  v[] holds S on entry and (backward) G bits on exit.
template <bool BWD, bool TRANSPOSED, bool MASKED>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&v)[32], const unsigned char *dstage, int r, int chunk,
                                               int gi, int gj0, int m, bool diagonal, float dmax,
                                               const DivConst &divw, float k2, float rni,
                                               const float *__restrict__ rn, float &rowsum)
{
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int jl = chunk * 32 + q * 4;                       // first of 4 columns inside the task
        float dv[4];
        if (!TRANSPOSED) {
            const float4 t4 = *reinterpret_cast<const float4 *>(dstage + ((r >> 6) * 16 + (jl >> 2)) * kPiecePitch +
                                                                (r & 63) * 16);
            dv[0] = t4.x; dv[1] = t4.y; dv[2] = t4.z; dv[3] = t4.w;
        } else {
            const unsigned char *base = dstage + (r >> 2) * kPiecePitch + (r & 3) * 4;
#pragma unroll
            for (int u = 0; u < 4; ++u) dv[u] = *reinterpret_cast<const float *>(base + (jl + u) * 16);
        }
        float4 rnj = make_float4(0.f, 0.f, 0.f, 0.f);
        if (BWD) rnj = __ldg(reinterpret_cast<const float4 *>(rn + gj0 + jl));
        const float rnjv[4] = {rnj.x, rnj.y, rnj.z, rnj.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = q * 4 + u;
            const float s = __uint_as_float(v[c]);
            const float w = div_fast(__fsub_rn(dmax, dv[u]), divw);
            float e = ex2_approx(s * (w * k2));
            if (MASKED) {
                const int gj = gj0 + jl + u;
                const bool valid = (gi < m) && (gj < m) && !(diagonal && gi == gj);
                e = valid ? e : 0.f;
            }
            if (!BWD) {
                rowsum += e;
            } else {
                v[c] = __float_as_uint(w * e * (rni + rnjv[u]));
            }
        }
    }
}

template <bool BWD>
__global__ void __launch_bounds__(kTcThreads, 1)
sweep_tc_kernel(const int4 *__restrict__ tasks, const int2 *__restrict__ strips, int n_strips,
                const float *__restrict__ zt, const uint16_t *__restrict__ zb, const float *__restrict__ dist,
                const float *__restrict__ rn,
                float *__restrict__ neg, float *__restrict__ dzacc, Stats *__restrict__ stats, int m, int n,
                int n_local, float k2)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = sm;
    unsigned char *sB = sA + kABytes;
    unsigned char *sD = sB + kStages * kBBytes;
    unsigned char *sBb = sD + kStages * kDBytes;                 // backward only
    TcBars *bars = reinterpret_cast<TcBars *>(sBb + (BWD ? kStages * kBbBytes : 0));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(bars) + kNumBars * 8);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t *fail = &stats->fail_site;
    constexpr uint32_t kTmemCols = BWD ? 256u : 128u;
    constexpr uint32_t kDzCol = 128u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&bars->full_b[i], 1);
            mbar_init(&bars->empty_b[i], 1);
            mbar_init(&bars->full_d[i], 1);
            mbar_init(&bars->empty_d[i], 4);
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->sg_full[i], 1);
            mbar_init(&bars->sg_empty[i], BWD ? 1 : 4);
            mbar_init(&bars->g_ready[i], 4);
        }
        mbar_init(&bars->dz_full, 1);
        mbar_init(&bars->dz_empty, 4);
        mbar_fence_init();
    }
    if (warp == 2) {
        tc_alloc(tmem_slot, kTmemCols);
        tc_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        int stage = 0;
        uint32_t ph = 0, a_ph = 0;
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x) {
            const int2 strip = strips[s];
            const int I = tasks[strip.x].x;
            if (lane == 0) {
                mbar_wait(&bars->a_empty, a_ph ^ 1u, fail, 1);
                mbar_arrive_expect_tx(&bars->a_full, kABytes);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) {
                    bulk_g2s(sA + kb * 16384, zt + (int64_t)(2 * I) * kBlockFloats + kb * 2048, 8192, &bars->a_full);
                    bulk_g2s(sA + kb * 16384 + 8192, zt + (int64_t)(2 * I + 1) * kBlockFloats + kb * 2048, 8192,
                             &bars->a_full);
                }
            }
            a_ph ^= 1u;
            for (int ti = strip.x; ti < strip.y; ++ti) {
                const int4 task = tasks[ti];
                if (lane == 0) {
                    mbar_wait(&bars->empty_b[stage], ph ^ 1u, fail, 2);
                    mbar_arrive_expect_tx(&bars->full_b[stage], kBBytes + (BWD ? kBbBytes : 0));
                    bulk_g2s(sB + stage * kBBytes, zt + (int64_t)task.y * kBlockFloats, kBBytes, &bars->full_b[stage]);
                    if (BWD)
                        bulk_g2s(sBb + stage * kBbBytes, zb + (int64_t)task.y * kBlockFloats, kBbBytes,
                                 &bars->full_b[stage]);
                    mbar_wait(&bars->empty_d[stage], ph ^ 1u, fail, 3);
                    mbar_arrive_expect_tx(&bars->full_d[stage], 32 * kPieceBytes);
                }
                __syncwarp();
                {
                    // piece `lane`: direct -> (rh = lane / 16, c4 = half * 16 + lane % 16); transposed -> (rh = half, c4 = lane)
                    const int half = task.y & 1;
                    const int src_piece = (task.w & kTaskTransposed) ? (half * 32 + lane)
                                                                    : ((lane >> 4) * 32 + half * 16 + (lane & 15));
                    bulk_g2s(sD + stage * kDBytes + lane * kPiecePitch,
                             dist + (int64_t)task.z * kTileFloats + src_piece * 256, kPieceBytes, &bars->full_d[stage]);
                }
                stage ^= 1;
                if (stage == 0) ph ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc_tf32(kTile, kTaskN, 0, 0);
            constexpr uint32_t idesc2 = umma_idesc_bf16(kTile, kD, 0, 1);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB), sBb_u = smem_u32(sBb);
            int stage = 0, sb = 0;
            uint32_t ph = 0, sph = 0, a_ph = 0, dz_ph = 0;
            bool pending = false, p_first = false;
            int p_stage = 0, p_sb = 0;
            uint32_t p_sph = 0;
            auto mma2 = [&]() {
                mbar_wait(&bars->g_ready[p_sb], p_sph, fail, 4);
                if (p_first) {
                    mbar_wait(&bars->dz_empty, dz_ph ^ 1u, fail, 5);
                    dz_ph ^= 1u;
                }
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kTaskN / 16; ++ks) {
                    // MN-major bf16 B: 64-element (128 B) atoms along d at LBO = 8 KiB, 8-row K groups at SBO = 1 KiB,
                    // 16 sample rows (2 KiB) per K step; A = 16 packed-bf16 K values = 8 TMEM columns per step
                    const uint64_t bdesc = umma_desc_sw128(sBb_u + p_stage * kBbBytes + ks * 2048, 8192, 1024);
                    tc_mma_ts_f16(tmem_base + kDzCol, tmem_base + p_sb * kTaskN + ks * 8, bdesc, idesc2,
                                  (p_first && ks == 0) ? 0u : 1u);
                }
                tc_commit(&bars->sg_empty[p_sb]);
                tc_commit(&bars->empty_b[p_stage]);
                pending = false;
            };
            for (int s = blockIdx.x; s < n_strips; s += gridDim.x) {
                const int2 strip = strips[s];
                mbar_wait(&bars->a_full, a_ph, fail, 6);
                a_ph ^= 1u;
                for (int ti = strip.x; ti < strip.y; ++ti) {
                    mbar_wait(&bars->full_b[stage], ph, fail, 7);
                    mbar_wait(&bars->sg_empty[sb], sph ^ 1u, fail, 8);
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t adesc = umma_desc_sw128(sA_u + kb * 16384 + ks * 32, 16, 1024);
                            const uint64_t bdesc = umma_desc_sw128(sB_u + stage * kBBytes + kb * 8192 + ks * 32, 16, 1024);
                            tc_mma_ss_tf32(tmem_base + sb * kTaskN, adesc, bdesc, idesc1, (kb | ks) ? 1u : 0u);
                        }
                    }
                    tc_commit(&bars->sg_full[sb]);
                    if (!BWD) tc_commit(&bars->empty_b[stage]);
                    if (BWD) {
                        if (pending) mma2();
                        pending = true;
                        p_stage = stage;
                        p_sb = sb;
                        p_sph = sph;
                        p_first = (ti == strip.x);
                    }
                    stage ^= 1;
                    if (stage == 0) ph ^= 1u;
                    sb ^= 1;
                    if (sb == 0) sph ^= 1u;
                }
                tc_commit(&bars->a_empty);            // every MMA1 of the strip has read sA
                if (BWD) {
                    if (pending) mma2();
                    tc_commit(&bars->dz_full);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int w4 = warp & 3;                       // TMEM lane quadrant this warp may touch
        const int r = w4 * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
        const float dmax = __uint_as_float(stats->dmax_bits);
        const DivConst divw = make_div(dmax);          // Dmax - Dmin, Dmin = +0
        int stage = 0, sb = 0;
        uint32_t ph = 0, sph = 0, dz_ph = 0;
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x) {
            const int2 strip = strips[s];
            const int I = tasks[strip.x].x;
            const int gi = I * kTile + r;
            const bool row_ok = gi < m;
            const float rni = (BWD && row_ok) ? rn[gi] : 0.f;
            float rowsum = 0.f;
            for (int ti = strip.x; ti < strip.y; ++ti) {
                const int4 task = tasks[ti];
                const int gj0 = task.y * kTaskN;
                const bool transposed = task.w & kTaskTransposed;
                const bool masked = task.w & (kTaskDiagonal | kTaskRagged);
                const bool diagonal = task.w & kTaskDiagonal;
                mbar_wait(&bars->sg_full[sb], sph, fail, 9);
                mbar_wait(&bars->full_d[stage], ph, fail, 10);
                tc_fence_after();
                const unsigned char *dstage = sD + stage * kDBytes;
#pragma unroll 1
                for (int chunk = 0; chunk < 2; ++chunk) {
                    uint32_t v[32];
                    const uint32_t taddr = lane_addr + sb * kTaskN + chunk * 32;
                    tc_ld32(taddr, v);
                    tc_wait_ld();
                    if (masked) {
                        if (transposed)
                            epilogue_chunk<BWD, true, true>(v, dstage, r, chunk, gi, gj0, m, diagonal, dmax, divw, k2, rni, rn, rowsum);
                        else
                            epilogue_chunk<BWD, false, true>(v, dstage, r, chunk, gi, gj0, m, diagonal, dmax, divw, k2, rni, rn, rowsum);
                    } else {
                        if (transposed)
                            epilogue_chunk<BWD, true, false>(v, dstage, r, chunk, gi, gj0, m, diagonal, dmax, divw, k2, rni, rn, rowsum);
                        else
                            epilogue_chunk<BWD, false, false>(v, dstage, r, chunk, gi, gj0, m, diagonal, dmax, divw, k2, rni, rn, rowsum);
                    }
                    if (BWD) {
                        // G as packed bf16x2 into the first 32 columns of the buffer S came from
                        uint32_t pk[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            pk[c] = pack_bf16x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
                        tc_st16(lane_addr + sb * kTaskN + chunk * 16, pk);
                    }
                }
                if (BWD) tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(BWD ? &bars->g_ready[sb] : &bars->sg_empty[sb]);
                    mbar_arrive(&bars->empty_d[stage]);
                }
                stage ^= 1;
                if (stage == 0) ph ^= 1u;
                sb ^= 1;
                if (sb == 0) sph ^= 1u;
            }
            if (!BWD) {
                if (row_ok) atomicAdd(neg + gi, rowsum);
            } else {
                mbar_wait(&bars->dz_full, dz_ph, fail, 11);
                dz_ph ^= 1u;
                tc_fence_after();
                float *orow = dzacc + (row_ok ? dz_out_row(gi, n, n_local) : 0) * kD;
#pragma unroll 1
                for (int chunk = 0; chunk < 4; ++chunk) {
                    uint32_t v[32];
                    tc_ld32(lane_addr + kDzCol + chunk * 32, v);
                    tc_wait_ld();
                    if (row_ok) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            red_add_v4(orow + chunk * 32 + q * 4, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->dz_empty);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, kTmemCols);
}

int launch_sweep_tc(bool backward, const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan,
                    const WsView &ws, float temperature, cudaStream_t stream)
{
    if (lay.n_strips == 0) return 0;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = lay.n_strips < sms ? lay.n_strips : sms;
    const float k2 = 1.4426950408889634f / temperature;
    const int n_local = dims.n / dims.world;
    cudaError_t e;
    if (backward) {
        e = cudaFuncSetAttribute(sweep_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBwd);
        if (e != cudaSuccess) return set_error((int)e, "tc sweep smem attr: %s", cudaGetErrorString(e));
        sweep_tc_kernel<true><<<grid, kTcThreads, kTcSmemBwd, stream>>>(plan.tasks, plan.strips, lay.n_strips, ws.zt,
                                                                       ws.zb, ws.dist, ws.rn, ws.neg, ws.dzacc,
                                                                    (Stats *)ws.stats, lay.m, dims.n, n_local, k2);
    } else {
        e = cudaFuncSetAttribute(sweep_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemFwd);
        if (e != cudaSuccess) return set_error((int)e, "tc sweep smem attr: %s", cudaGetErrorString(e));
        sweep_tc_kernel<false><<<grid, kTcThreads, kTcSmemFwd, stream>>>(plan.tasks, plan.strips, lay.n_strips, ws.zt,
                                                                        ws.zb, ws.dist, ws.rn, ws.neg, ws.dzacc,
                                                                     (Stats *)ws.stats, lay.m, dims.n, n_local, k2);
    }
    return check_launch("sweep_tc_kernel");
}

}  // namespace smh
