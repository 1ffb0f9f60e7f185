// simhand_b200: the sharded step with the exchange fused into the kernels (smh_exchange_t.fused).
//
// One step of a rank is six launches and no barrier or copy kernel:
//   shard_prep_kernel      own 2 * n_local rows -> operand images + packed joints stored into EVERY rank's workspace
//                          (the all-gather of the inputs, done on the converted rows), positive-pair distance and
//                          <z1_k, z2_k> of the own samples to every rank, max / min / flags / distance bound by remote
//                          atomicMax, local accumulators zeroed, stage 1 signalled
//   mpjpe_kernel           head: waits for stage 1 of every rank; tail: last CTA pushes Dmax, signals stage 2
//   sweep_tc_kernel<fwd>   head: waits for stage 2; tail: the CTAs ship the partial row sums, signal stage 3
//   rn_fused_kernel        head: waits for stage 3; row sums = rank-ordered sum of the partials, reciprocals
//   sweep_tc_kernel<bwd>   tail: the CTAs ship the partial gradient rows to their owners, signal stage 4
//   finalize_fused_kernel  head: waits for stage 4; own-row gradients (rank-ordered sum of the partials) and the loss of the
//                          global batch, evaluated identically on every rank from the delivered row sums / positives
// Signals are monotonic step counters in the ranks' signal blocks (CUDA-graph replay safe).  Per-step scalars that peers
// write into (max / min / flags / bound, positives) are double-buffered by step parity: a rank can be at most one
// shard_prep ahead of the slowest rank (its next mpjpe head waits for everybody's stage 1), everything else a peer
// writes is consumed before the writer can pass the stage-4 wait of the same step.  Every wait is bounded; a timeout
// poisons the whole group (sticky) and every later loss is NaN.
#include <math_constants.h>

#include <algorithm>

#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

struct ShardPrepArgs {
    smh_inputs_t in;            // this rank's tensors (n_local samples per view)
    int n, d, n_local;
    int images;                 // bit 0: fp32 / tf32 image, 1: bf16, 2: fp16
    int round_tf32;
    int diff;
    long long off_zt, off_zb, off_zh, off_jp;
    long long zero_off, zero_bytes;   // local accumulators: [off_neg, off_posd)
    int push_images;            // 0: smh_shard_push_z ships the z images on a parallel branch
};

__global__ void __launch_bounds__(256) shard_prep_kernel(const __grid_constant__ ShardPrepArgs a, const __grid_constant__ Peers pe)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    uint32_t *sig = pe.my_sig();
    const uint32_t epoch = sig[kSigEpoch] + 1u;           // written by this rank's previous step (stream order)
    Stats *gs = pe.gstats(pe.rank, epoch);
    Stats *st = pe.stats(pe.rank);
    const smh_inputs_t &in = a.in;
    PhaseClock clk(pe, 0);

    if (blockIdx.x == 0) {
        // rank 0 publishes the joints of global sample 0 first: every rank bounds its rows against them.  Each coordinate
        // travels as one 8-byte word {bits, epoch}: the store is atomic, the tag is its own "valid" flag (no fence, no
        // separate flag: one NVLink latency)
        if (pe.rank == 0 && threadIdx.x < 42) {
            const float v = in.j1_dev[(int64_t)(threadIdx.x >> 1) * in.j_joint_stride + (threadIdx.x & 1) * in.j_coord_stride];
            const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
            for (int p = 0; p < pe.world; ++p)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long *>(pe.sig[p] + kSigPivot) + threadIdx.x), "l"(w) : "memory");
        }
        // the scalars of the NEXT step (same slot as the previous one, which this rank has finished reading; peers
        // write it only after they have seen this rank's stage 4 of the current step)
        if (threadIdx.x < 12) reinterpret_cast<uint32_t *>(pe.gstats(pe.rank, epoch + 1u))[threadIdx.x] = 0u;
        if (threadIdx.x == 0) {
            st->dmax_bits = 0u;                            // rank-local maximum of the MPJPE kernel
            st->flags = 0u;
            st->fail_site = 0u;
            st->dsum = 0.0;
            st->reserved = 0u;                             // work counter of the persistent MPJPE kernel
        }
    }

    // zero the local accumulators (row sums, reciprocals, row terms, gradient accumulator)
    {
        float4 *z4 = reinterpret_cast<float4 *>(pe.ws[pe.rank] + a.zero_off);
        const long long n4 = a.zero_bytes / 16;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
            z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    uint32_t bad_acc = 0u, pmax_bits = 0u, pmin_inv = 0u;
    const int rows = 2 * a.n_local;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < rows; w += gridDim.x * wpb) {
        const int v = w >= a.n_local ? 1 : 0;
        const int kl = w - v * a.n_local;
        const int kg = pe.rank * a.n_local + kl;
        const int64_t grow = (int64_t)v * a.n + kg;
        const float *zp = (v ? in.z2_dev : in.z1_dev) + (int64_t)kl * in.z_row_stride;
        float4 zraw = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            const int c = 4 * lane;
            if (c + 3 < a.d && ((reinterpret_cast<uintptr_t>(zp + c) & 15) == 0)) {
                zraw = *reinterpret_cast<const float4 *>(zp + c);
            } else {
                if (c + 0 < a.d) zraw.x = zp[c + 0];
                if (c + 1 < a.d) zraw.y = zp[c + 1];
                if (c + 2 < a.d) zraw.z = zp[c + 2];
                if (c + 3 < a.d) zraw.w = zp[c + 3];
            }
        }
        float jx = 0.f, jy = 0.f;
        if (lane < kJ) {
            const float *jb = (v ? in.j2_dev : in.j1_dev) + (int64_t)kl * in.j_sample_stride + (int64_t)lane * in.j_joint_stride;
            jx = jb[0];
            jy = jb[in.j_coord_stride];
        }
        float4 zv = zraw;
        if (a.round_tf32) {
            zv.x = to_tf32(zv.x);
            zv.y = to_tf32(zv.y);
            zv.z = to_tf32(zv.z);
            zv.w = to_tf32(zv.w);
        }
        const uint2 b16 = make_uint2(pack_bf16x2(zv.x, zv.y), pack_bf16x2(zv.z, zv.w));
        const uint2 h16 = make_uint2(pack_f16x2(zv.x, zv.y), pack_f16x2(zv.z, zv.w));
        const int images = a.push_images ? a.images : 0;
        // packed joints: lane L < 10 holds (x_2L, x_2L+1, y_2L, y_2L+1), lane 10 holds (x_20, y_20, 0, 0)
        float4 jq;
        {
            const int s0 = lane < 10 ? 2 * lane : 20, s1 = lane < 10 ? 2 * lane + 1 : 20;
            const float xa = __shfl_sync(0xffffffffu, jx, s0), xb = __shfl_sync(0xffffffffu, jx, s1);
            const float ya = __shfl_sync(0xffffffffu, jy, s0), yb = __shfl_sync(0xffffffffu, jy, s1);
            jq = lane < 10 ? make_float4(xa, xb, ya, yb) : make_float4(xa, ya, 0.f, 0.f);
        }
        const int64_t i_zt = zt_index(grow, 4 * lane), i_zb = zb_index(grow, 4 * lane);
        for (int p = 0; p < pe.world; ++p) {
            unsigned char *w8 = pe.ws[p];
            if (images & 1) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(w8 + a.off_zt) + i_zt) = zv;
            if (images & 2) *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(w8 + a.off_zb) + i_zb) = b16;
            if (images & 4) *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(w8 + a.off_zh) + i_zb) = h16;
            if (lane < 11) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(w8 + a.off_jp) + grow * kJP + 4 * lane) = jq;
        }
        // domain of the branch-free exact sqrt (as smh_prep.cu)
        if (lane < kJ) {
            const float ax = fabsf(jx), ay = fabsf(jy);
            const bool fin = (ax <= 3.0e38f) && (ay <= 3.0e38f);
            const bool okx = (ax == 0.f) || (ax >= 5.9604645e-8f && ax <= 1.1529215e18f);
            const bool oky = (ay == 0.f) || (ay >= 5.9604645e-8f && ay <= 1.1529215e18f);
            bad_acc |= (fin ? 0u : SMH_FLAG_NONFINITE) | ((okx && oky) ? 0u : SMH_FLAG_SLOW_DOMAIN);
        }
        // positive pair (k, k + N): both views of a sample live on the same rank.  utils.py:229-231 and the positive logit
        // <z1_k, z2_k> in fp32 from the unrounded z (utils.py:420-423); the view-1 warp does it.
        if (v == 0) {
            const float *z2p = in.z2_dev + (int64_t)kl * in.z_row_stride;
            float dot = 0.f;
            {
                const int c = 4 * lane;
                if (c + 0 < a.d) dot = fmaf(zraw.x, z2p[c + 0], dot);
                if (c + 1 < a.d) dot = fmaf(zraw.y, z2p[c + 1], dot);
                if (c + 2 < a.d) dot = fmaf(zraw.z, z2p[c + 2], dot);
                if (c + 3 < a.d) dot = fmaf(zraw.w, z2p[c + 3], dot);
            }
            dot = warp_sum(dot);
            float dx = 0.f, dy = 0.f;
            if (lane < kJ) {
                const float *pb = in.j2_dev + (int64_t)kl * in.j_sample_stride + (int64_t)lane * in.j_joint_stride;
                dx = __fsub_rn(jx, pb[0]);
                dy = __fsub_rn(jy, pb[in.j_coord_stride]);
            }
            float dk;
            if (a.diff == SMH_DIFF_MPJPE) {
                const float nk = lane < kJ ? __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))) : 0.f;
                float s = __shfl_sync(0xffffffffu, nk, 16);
#pragma unroll
                for (int q = 17; q <= 20; ++q) s = __fadd_rn(s, __shfl_sync(0xffffffffu, nk, q));
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    s = __fadd_rn(s, __fadd_rn(__shfl_sync(0xffffffffu, nk, q), __shfl_sync(0xffffffffu, nk, q + 8)));
                dk = __fdiv_rn(s, 21.0f);
            } else if (a.diff == SMH_DIFF_EUCLID) {
                dk = __fsqrt_rn(warp_sum(__fmaf_rn(dy, dy, __fmul_rn(dx, dx))));       // utils.py:265-274: || p1 - p2 ||_2
            } else {
                if (a.diff == SMH_DIFF_W_ABS) {
                    dx = fabsf(dx);
                    dy = fabsf(dy);
                }
                const float mx = __fdiv_rn(warp_sum(dx), 21.0f), my = __fdiv_rn(warp_sum(dy), 21.0f);
                dk = __fsqrt_rn(__fmaf_rn(my, my, __fmul_rn(mx, mx)));
            }
            if (lane == 0) {
                for (int p = 0; p < pe.world; ++p) {
                    float *pi = pe.posinfo(p, epoch);
                    pi[kg] = dk;
                    pi[a.n + kg] = dot;
                }
                const uint32_t b = __float_as_uint(dk);
                if (b <= 0x7f800000u) {
                    pmax_bits = max(pmax_bits, b);
                    pmin_inv = max(pmin_inv, 0x7fffffffu - b);
                } else {
                    bad_acc |= SMH_FLAG_NONFINITE;
                }
            }
        }
    }
    // fold this thread's scalars into the rank's own slot
    bad_acc |= __shfl_xor_sync(0xffffffffu, bad_acc, 16);
    bad_acc |= __shfl_xor_sync(0xffffffffu, bad_acc, 8);
    bad_acc |= __shfl_xor_sync(0xffffffffu, bad_acc, 4);
    bad_acc |= __shfl_xor_sync(0xffffffffu, bad_acc, 2);
    bad_acc |= __shfl_xor_sync(0xffffffffu, bad_acc, 1);
    if (lane == 0) {
        if (bad_acc) atomicOr_system(&gs->flags, bad_acc);
        if (pmax_bits | pmin_inv) {
            atomicMax_system(&gs->pmax_bits, pmax_bits);
            atomicMax_system(&gs->pmin_inv, pmin_inv);
        }
    }

    // D(row, global sample 0): by the triangle inequality 2 max_i D_i0 bounds every D_ij (scale of the 16-bit image)
    __shared__ float pivot[42];
    __shared__ uint32_t wb[8];
    clk.lap();                                   // [0] zeroing + images + positives issued
    if (threadIdx.x < 42) {
        // poll this coordinate's tagged word (bounded like every cross-rank wait)
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(sig + kSigPivot) + threadIdx.x;
        unsigned long long w;
        const unsigned long long t0 = global_ns();
        const unsigned long long limit = (unsigned long long)(pe.timeout_ms ? pe.timeout_ms : 30000u) * 1000000ull;
        for (uint32_t spin = 1;; ++spin) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
            if ((uint32_t)(w >> 32) == epoch) break;
            if ((spin & 63u) == 0u) {
                if (global_ns() - t0 > limit) {
                    poison_group(pe, 120u);
                    break;
                }
                __nanosleep(64);
            }
        }
        pivot[threadIdx.x] = __uint_as_float((uint32_t)w);
    }
    __syncthreads();
    clk.lap();                                   // [1] pivot arrived
    uint32_t bound_bits = 0u;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < rows; w += gridDim.x * wpb) {
        const int v = w >= a.n_local ? 1 : 0;
        const int kl = w - v * a.n_local;
        float nk = 0.f;
        if (lane < kJ) {
            const float *jb = (v ? in.j2_dev : in.j1_dev) + (int64_t)kl * in.j_sample_stride + (int64_t)lane * in.j_joint_stride;
            const float ex = jb[0] - pivot[2 * lane], ey = jb[in.j_coord_stride] - pivot[2 * lane + 1];
            nk = sqrtf(fmaf(ey, ey, ex * ex));
        }
        const float d0 = warp_sum(nk) * (1.0f / 21.0f) * 1.0001f;
        if (d0 >= 0.f && d0 <= 3.0e38f) bound_bits = max(bound_bits, __float_as_uint(d0));
    }
    if (lane == 0) wb[threadIdx.x >> 5] = bound_bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = wb[0];
        for (int w = 1; w < wpb; ++w) b = max(b, wb[w]);
        if (b != 0u) atomicMax_system(&gs->dbound_bits, b);
        // one fence per block, after the block barrier: it orders every thread's stores into the peers (observed
        // through the barrier) before the ticket -- a fence in each of the 65 k threads cost 20-30 us
        clk.lap();                               // [2] bound rows done
        block_release_fence();
        clk.lap();                               // [3] fence
        const unsigned ticket = atomicAdd(&st->counter, 1u);
        if (ticket == gridDim.x - 1) {
            st->counter = 0u;
            __threadfence();
            // all-reduce of the rank's scalars: max is idempotent, so whatever the peers already added may ride along
            const uint32_t f = atomicOr_system(&gs->flags, 0u), pm = atomicMax_system(&gs->pmax_bits, 0u),
                           pn = atomicMax_system(&gs->pmin_inv, 0u), db = atomicMax_system(&gs->dbound_bits, 0u);
            for (int p = 0; p < pe.world; ++p) {
                if (p == pe.rank) continue;
                Stats *o = pe.gstats(p, epoch);
                if (f) atomicOr_system(&o->flags, f);
                atomicMax_system(&o->pmax_bits, pm);
                atomicMax_system(&o->pmin_inv, pn);
                atomicMax_system(&o->dbound_bits, db);
            }
            sig[kSigEpoch] = epoch;
            stage_signal(pe, 1, epoch);
        }
        clk.lap();                               // [4] block 0 done
    }
}

// The z-image part of the all-gather on its own: own rows -> operand images of the engine -> every rank's workspace; the
// rank's last block signals stage 5.  Launched on a second stream next to the MPJPE kernel (which needs only the joints):
// the transfer, 3/4 of the bytes a rank ships per step, leaves the critical path.
__global__ void __launch_bounds__(256)
shard_push_z_kernel(const __grid_constant__ ShardPrepArgs a, const __grid_constant__ Peers pe)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    uint32_t *sig = pe.my_sig();
    const uint32_t epoch = sig[kSigEpoch];                 // shard_prep of this step has completed (stream order)
    const smh_inputs_t &in = a.in;
    const int rows = 2 * a.n_local;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < rows; w += gridDim.x * wpb) {
        const int v = w >= a.n_local ? 1 : 0;
        const int kl = w - v * a.n_local;
        const int64_t grow = (int64_t)v * a.n + (int64_t)pe.rank * a.n_local + kl;
        const float *zp = (v ? in.z2_dev : in.z1_dev) + (int64_t)kl * in.z_row_stride;
        float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = 4 * lane;
        if (c + 3 < a.d && ((reinterpret_cast<uintptr_t>(zp + c) & 15) == 0)) {
            zv = *reinterpret_cast<const float4 *>(zp + c);
        } else {
            if (c + 0 < a.d) zv.x = zp[c + 0];
            if (c + 1 < a.d) zv.y = zp[c + 1];
            if (c + 2 < a.d) zv.z = zp[c + 2];
            if (c + 3 < a.d) zv.w = zp[c + 3];
        }
        if (a.round_tf32) {
            zv.x = to_tf32(zv.x);
            zv.y = to_tf32(zv.y);
            zv.z = to_tf32(zv.z);
            zv.w = to_tf32(zv.w);
        }
        const uint2 b16 = make_uint2(pack_bf16x2(zv.x, zv.y), pack_bf16x2(zv.z, zv.w));
        const uint2 h16 = make_uint2(pack_f16x2(zv.x, zv.y), pack_f16x2(zv.z, zv.w));
        const int64_t i_zt = zt_index(grow, 4 * lane), i_zb = zb_index(grow, 4 * lane);
        for (int p = 0; p < pe.world; ++p) {
            unsigned char *w8 = pe.ws[p];
            if (a.images & 1) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(w8 + a.off_zt) + i_zt) = zv;
            if (a.images & 2) *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(w8 + a.off_zb) + i_zb) = b16;
            if (a.images & 4) *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(w8 + a.off_zh) + i_zb) = h16;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        block_release_fence();
        const unsigned ticket = atomicAdd(sig + kSigTicketZ, 1u);
        if (ticket == gridDim.x - 1) {
            sig[kSigTicketZ] = 0u;
            stage_signal(pe, kStageZ, epoch);
        }
    }
}

static ShardPrepArgs make_prep_args(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine);

int launch_shard_push_z(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine,
                        const Peers &peers, cudaStream_t stream)
{
    const ShardPrepArgs a = make_prep_args(dims, lay, in, engine);
    // a modest grid: the kernel shares the SMs with the MPJPE kernel and is bound by NVLink, not by its own parallelism
    int blocks = (2 * a.n_local + 7) / 8;
    if (blocks > kNumCtas * 2) blocks = kNumCtas * 2;
    shard_push_z_kernel<<<blocks, 256, 0, stream>>>(a, peers);
    return check_launch("shard_push_z_kernel");
}

static ShardPrepArgs make_prep_args(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine)
{
    ShardPrepArgs a;
    a.in = in;
    a.n = dims.n;
    a.d = dims.d;
    a.n_local = dims.n / dims.world;
    a.round_tf32 = engine == SMH_ENGINE_TC_TF32;
    a.images = engine == SMH_ENGINE_FP32 ? 1 : (2 | (engine == SMH_ENGINE_TC_TF32 ? 1 : 0) |
                                                 (engine == SMH_ENGINE_TC_FP16 ? 4 : 0));
    a.diff = dims.diff_type;
    a.off_zt = lay.off_zt;
    a.off_zb = lay.off_zb;
    a.off_zh = lay.off_zh;
    a.off_jp = lay.off_jp;
    a.zero_off = lay.off_neg;
    a.zero_bytes = lay.off_posd - lay.off_neg;
    a.push_images = 1;
    return a;
}

int launch_shard_prep(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine,
                      const Peers &peers, cudaStream_t stream)
{
    const bool no_images = (engine & SMH_SHARD_PREP_NO_IMAGES) != 0;
    ShardPrepArgs a = make_prep_args(dims, lay, in, engine & ~SMH_SHARD_PREP_NO_IMAGES);
    a.push_images = no_images ? 0 : 1;
    int blocks = (2 * a.n_local + 7) / 8;
    if (blocks > kNumCtas * 8) blocks = kNumCtas * 8;
    shard_prep_kernel<<<blocks, 256, 0, stream>>>(a, peers);
    return check_launch("shard_prep_kernel");
}

// neg_i = rank-ordered sum of the partial row sums every rank delivered; 1 / neg_i for the backward sweep.
// signal4: loss-only step (no backward sweep follows): this launch closes stage 4.
__global__ void __launch_bounds__(256)
rn_fused_kernel(const __grid_constant__ Peers pe, float *__restrict__ neg, const float *__restrict__ negparts, float *__restrict__ rn, int m,
                int mp, int signal4)
{
    uint32_t *sig = pe.my_sig();
    const uint32_t epoch = sig[kSigEpoch];
    stage_wait(pe, 3, epoch);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mp) {
        float v = 0.f;
        for (int p = 0; p < pe.world; ++p) v += __ldcg(negparts + (int64_t)p * mp + i);
        neg[i] = v;
        rn[i] = (i < m) ? __frcp_rn(v) : 0.f;
    }
    if (signal4) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned ticket = atomicAdd(sig + kSigTicket, 1u);
            if (ticket == gridDim.x - 1) {
                sig[kSigTicket] = 0u;
                stage_signal(pe, 4, epoch);
            }
        }
    }
}

int launch_rn_fused(const smh_layout_t &lay, const WsView &ws, const Peers &peers, bool signal4, cudaStream_t stream)
{
    const int mp = lay.tiles_per_side * kTile;
    rn_fused_kernel<<<(mp + 255) / 256, 256, 0, stream>>>(peers, ws.neg, ws.negparts, ws.rn, lay.m, mp, signal4 ? 1 : 0);
    return check_launch("rn_fused_kernel");
}

// mean of the joint distance over all M^2 ordered pairs (non_linear weights): rank-ordered sum of the delivered parts
__device__ __forceinline__ double fused_dsum(const Peers &pe)
{
    const double *parts = reinterpret_cast<const double *>(pe.lossparts(pe.rank) + 16);
    double t = 0.0;
    for (int p = 0; p < pe.world; ++p) t += __ldcg(parts + p);
    return t;
}

// Gradients of the own rows and the loss of the global batch.  `in` describes this rank's tensors.
__global__ void __launch_bounds__(256)
finalize_fused_kernel(smh_inputs_t in, int n, int d, const __grid_constant__ Peers pe, const float *__restrict__ neg, float *__restrict__ rowloss,
                      const float *__restrict__ dzparts, int pos_mode, float lambda_pos, float inv_tau, float grad_scale,
                      float *__restrict__ loss_out, float *__restrict__ dz1, float *__restrict__ dz2,
                      int64_t dz_row_stride)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    uint32_t *sig = pe.my_sig();
    const uint32_t epoch = sig[kSigEpoch];
    PhaseClock clk(pe, 5);
    stage_wait(pe, 4, epoch);
    clk.lap();                                   // [0] stage wait
    Stats *st = pe.stats(pe.rank);
    const Stats *gs = pe.gstats(pe.rank, epoch);
    const float *posd = pe.posinfo(pe.rank, epoch);
    const float *dots = posd + n;
    const int m = 2 * n;
    const float pmax = __uint_as_float(__ldcg(&gs->pmax_bits));
    const float pmin = __uint_as_float(0x7fffffffu - __ldcg(&gs->pmin_inv));
    const float pden = __fsub_rn(pmax, pmin);
    const int n_local = in.n_local;
    const int k_lo = pe.rank * n_local;
    const float gsf = grad_scale * inv_tau / (float)m;
    const int64_t part_stride = (int64_t)2 * n_local * kD;
    __shared__ float pos_mean_s;
    __shared__ float red[256];
    if (pos_mode == 3) {
        float acc = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) acc += __ldcg(posd + i);
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int s2 = 128; s2 > 0; s2 >>= 1) {
            if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
            __syncthreads();
        }
        if (threadIdx.x == 0) pos_mean_s = red[0] / (float)n;
        __syncthreads();
    }
    const float pos_mean = pos_mode == 3 ? pos_mean_s : 0.f;

    auto pos_weight = [&](int k) {
        const float pd = __ldcg(posd + k);
        float wp = pos_mode == 1 ? 1.0f : __fdiv_rn(__fsub_rn(pmax, pd), pden);
        if (pos_mode == 3) wp = __fdiv_rn(1.0f, 1.0f + expf(lambda_pos * (pd - pos_mean)));
        return wp;
    };
    // loss terms of ALL rows, one thread per row (every rank evaluates the global loss from the delivered row sums and
    // positives: no further exchange)                                                               utils.py:420-426
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < m; row += gridDim.x * blockDim.x) {
        const int k = row >= n ? row - n : row;
        rowloss[row] = logf(__ldcg(neg + row)) - __ldcg(dots + k) * pos_weight(k) * inv_tau;
    }
    clk.lap();                                   // [1] loss rows
    // gradients of the OWN rows, one warp per row: rank-ordered sum of the delivered partials minus the positive term
    if (dz1 != nullptr) {
        for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < 2 * n_local; w += gridDim.x * wpb) {
            const int v = w >= n_local ? 1 : 0;
            const int kl = w - v * n_local;
            const float two_wp = 2.f * pos_weight(k_lo + kl);
            const float *zp = (v ? in.z1_dev : in.z2_dev) + (int64_t)kl * in.z_row_stride;        // the partner's row
            const float *src = dzparts + (int64_t)w * kD;
            float *dst = (v ? dz2 : dz1) + (int64_t)kl * dz_row_stride;
            if (d == kD && ((reinterpret_cast<uintptr_t>(zp) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
                // one 16-byte load per partial and lane, four of them in flight before the first add; rank order
                const float4 zq = *(reinterpret_cast<const float4 *>(zp) + lane);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int p0 = 0; p0 < pe.world; p0 += 4) {
                    float4 part[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (p0 + q < pe.world)
                            part[q] = __ldcg(reinterpret_cast<const float4 *>(src + (int64_t)(p0 + q) * part_stride) + lane);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (p0 + q < pe.world) {
                            acc.x += part[q].x;
                            acc.y += part[q].y;
                            acc.z += part[q].z;
                            acc.w += part[q].w;
                        }
                }
                *(reinterpret_cast<float4 *>(dst) + lane) =
                    make_float4(gsf * (acc.x - two_wp * zq.x), gsf * (acc.y - two_wp * zq.y), gsf * (acc.z - two_wp * zq.z),
                                gsf * (acc.w - two_wp * zq.w));
            } else {
                for (int c = lane; c < d; c += 32) {
                    float acc = __ldcg(src + c);
                    for (int p = 1; p < pe.world; ++p) acc += __ldcg(src + (int64_t)p * part_stride + c);   // rank order
                    dst[c] = gsf * (acc - two_wp * zp[c]);
                }
            }
        }
    }

    clk.lap();                                   // [2] own-row gradients
    // last block: the row terms in a fixed order (the same on every rank)
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd(&st->counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed order (thread t owns rows t, t + 256, ...; four independent partial sums keep 8 loads in flight)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = threadIdx.x;
    for (; i + 768 < m; i += 1024) {
        const float v0 = __ldcg(rowloss + i), v1 = __ldcg(rowloss + i + 256), v2 = __ldcg(rowloss + i + 512),
                    v3 = __ldcg(rowloss + i + 768);
        a0 += v0;
        a1 += v1;
        a2 += v2;
        a3 += v3;
    }
    for (; i < m; i += 256) a0 += __ldcg(rowloss + i);
    const float acc = (a0 + a1) + (a2 + a3);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st->counter = 0u;
        float loss = red[0] / (float)m;
        if ((__ldcg(&gs->flags) & SMH_FLAG_NONFINITE) || st->fail_site != 0u || ld_acquire_sys(sig + kSigPoison) != 0u)
            loss = CUDART_NAN_F;
        st->loss = loss;
        if (loss_out) *loss_out = loss;
    }
}

int launch_finalize_fused(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                          int pos_mode, float temperature, float grad_scale, float *loss, float *dz1, float *dz2,
                          int64_t dz_row_stride, const Peers &peers, cudaStream_t stream)
{
    const int n_local = dims.n / dims.world;
    int blocks = std::max((lay.m + 255) / 256, (2 * n_local + 7) / 8);           // a thread per loss row, a warp per own row
    if (blocks > 4 * kNumCtas) blocks = 4 * kNumCtas;
    finalize_fused_kernel<<<blocks, 256, 0, stream>>>(in, dims.n, dims.d, peers, ws.neg, ws.rowloss, ws.dzparts, pos_mode,
                                                      dims.lambda_pos, 1.0f / temperature, grad_scale, loss, dz1, dz2,
                                                      dz_row_stride);
    return check_launch("finalize_fused_kernel");
}

}  // namespace smh
