// simhand_b200: the two kernels of the peer exchange that are not fused into a compute kernel.
//   push_inputs_kernel  push-based all-gather: every rank packs its [z1|z2|joints1|joints2] straight from the caller's
//                       tensors into slot `rank` of every peer's gathered-input buffer (coalesced stores on peer
//                       pointers, NVLink)
//   barrier_kernel      device-side barrier between the phases of a step.  Monotonic counters (word 0 = barriers this
//                       rank has entered, word 8 + p = last barrier peer p announced), so the same kernel node can be
//                       replayed from a CUDA graph without host-side epochs.  The wait is bounded (smh_exchange_t.timeout_ms,
//                       30 s by default); a timeout poisons the group and the loss becomes NaN.
// This is the UNFUSED form of the exchange (14 launches per step); the default is the fused one in smh_shard.cu.
//   exchange_neg_kernel / exchange_dz_kernel   all-gather of the partial row sums and reduce-scatter payload of the
//                       partial gradient rows: plain 16-byte stores into slot `rank` of the peers' partial buffers; the
//                       consumers add the partials in rank order (deterministic).  (Adding straight into the peers'
//                       accumulators from the sweep epilogues was measured 4-5x slower at 8 ranks: ~1 M small remote
//                       reductions per rank and step.)
// The Dmax all-reduce is fused into the MPJPE kernel (its last CTA pushes the rank's maximum to every peer).
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

struct PushArgs {
    float *dst[kMaxPeers];      // peer p's gathered-input buffer, already offset to this rank's chunk
    int world;
    int aligned16;              // every dst base is 16-byte aligned
};

// One warp per (view, local sample): gathers the caller's (possibly strided) z row and [21, 2] joint view and writes
// them in the packed chunk layout [z1 | z2 | joints1 | joints2] straight into every peer's buffer.
__global__ void __launch_bounds__(256) push_inputs_kernel(smh_inputs_t in, PushArgs a, int n_local, int d)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int64_t off_z2 = (int64_t)n_local * d, off_j1 = 2 * off_z2, off_j2 = off_j1 + (int64_t)n_local * 42;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < 2 * n_local; w += gridDim.x * wpb) {
        const int v = w >= n_local ? 1 : 0;
        const int k = w - v * n_local;
        const float *zp = (v ? in.z2_dev : in.z1_dev) + (int64_t)k * in.z_row_stride;
        const float *jb = (v ? in.j2_dev : in.j1_dev) + (int64_t)k * in.j_sample_stride;
        const int64_t zo = (v ? off_z2 : 0) + (int64_t)k * d;
        const int64_t jo = (v ? off_j2 : off_j1) + (int64_t)k * 42;
        if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(zp) | (uintptr_t)(zo * 4)) & 15) == 0 && a.aligned16) {
            for (int c = lane * 4; c < d; c += 128) {
                const float4 val = *reinterpret_cast<const float4 *>(zp + c);
#pragma unroll 4
                for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4 *>(a.dst[p] + zo + c) = val;
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                const float val = zp[c];
                for (int p = 0; p < a.world; ++p) a.dst[p][zo + c] = val;
            }
        }
        for (int c = lane; c < 42; c += 32) {
            const float val = jb[(int64_t)(c >> 1) * in.j_joint_stride + (c & 1) * in.j_coord_stride];
            for (int p = 0; p < a.world; ++p) a.dst[p][jo + c] = val;
        }
    }
}

int launch_push_inputs(const smh_exchange_t &exch, const smh_inputs_t &in, int n_local, int d, cudaStream_t stream)
{
    PushArgs a;
    a.world = exch.world;
    const int64_t chunk = 2ll * n_local * (d + 42);
    for (int p = 0; p < exch.world; ++p) {
        if (!exch.xin_peer[p]) return set_error(SMH_E_ARG, "exchange xin_peer[%d] is null", p);
        a.dst[p] = (float *)exch.xin_peer[p] + (int64_t)exch.rank * chunk;
    }
    a.aligned16 = 1;
    for (int p = 0; p < exch.world; ++p)
        if (reinterpret_cast<uintptr_t>(a.dst[p]) & 15) a.aligned16 = 0;
    int blocks = (2 * n_local + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    push_inputs_kernel<<<blocks, 256, 0, stream>>>(in, a, n_local, d);
    return check_launch("push_inputs_kernel");
}

// all-gather of the partial row sums: this rank's neg[0, Mp) -> negparts[rank][.] on every rank (16-byte stores)
__global__ void __launch_bounds__(256) exchange_neg_kernel(Peers pe, int mp)
{
    const float4 *src = reinterpret_cast<const float4 *>(pe.neg(pe.rank));
    const int n4 = mp / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const float4 v = src[i];
        for (int p = 0; p < pe.world; ++p)
            reinterpret_cast<float4 *>(pe.negparts(p) + (int64_t)pe.rank * mp)[i] = v;
    }
}

int launch_exchange_neg(const smh_layout_t &lay, const Peers &peers, cudaStream_t stream)
{
    const int mp = lay.tiles_per_side * kTile;
    int blocks = (mp / 4 + 255) / 256;
    exchange_neg_kernel<<<blocks, 256, 0, stream>>>(peers, mp);
    return check_launch("exchange_neg_kernel");
}

// payload of the reduce-scatter: rows [p * 2 n_local, (p + 1) * 2 n_local) of this rank's full (rank-major) partial
// gradient go to dzparts[rank][.] on rank p; the owner adds the `world` blocks in rank order in smh_finalize
__global__ void __launch_bounds__(256) exchange_dz_kernel(Peers pe, int n_local)
{
    const int64_t block4 = (int64_t)2 * n_local * kD / 4;                 // float4 per destination rank
    const float4 *src = reinterpret_cast<const float4 *>(pe.dzacc(pe.rank));
    const int64_t total = block4 * pe.world;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i / block4);
        const int64_t j = i - (int64_t)p * block4;
        reinterpret_cast<float4 *>(pe.dzparts(p))[(int64_t)pe.rank * block4 + j] = src[i];
    }
}

int launch_exchange_dz(const smh_dims_t &dims, const smh_layout_t &lay, const Peers &peers, cudaStream_t stream)
{
    (void)lay;
    const int n_local = dims.n / dims.world;
    exchange_dz_kernel<<<148 * 4, 256, 0, stream>>>(peers, n_local);
    return check_launch("exchange_dz_kernel");
}

struct BarrierArgs {
    uint32_t *sig[kMaxPeers];
    int world, rank;
    unsigned timeout_ms;
};

__global__ void __launch_bounds__(32) barrier_kernel(BarrierArgs a)
{
    __shared__ uint32_t cnt_s;
    uint32_t *mine = a.sig[a.rank];
    if (threadIdx.x == 0) {
        cnt_s = mine[0] + 1u;
        mine[0] = cnt_s;
    }
    __syncwarp();
    const uint32_t cnt = cnt_s;
    __threadfence_system();                        // everything this rank wrote to peers is visible before the signal
    const int p = threadIdx.x;
    if (p < a.world && p != a.rank) {
        volatile uint32_t *theirs = a.sig[p] + 8 + a.rank;
        *theirs = cnt;                             // announce: this rank has entered barrier `cnt`
        volatile uint32_t *from_p = mine + 8 + p;
        // bounded: a missing peer must not hang the device -- but a timeout is a FAILURE, never a fall-through: the
        // group is poisoned (sticky word on every rank) and smh_finalize turns the loss into NaN
        const unsigned long long t0 = global_ns();
        const unsigned long long limit = (unsigned long long)(a.timeout_ms ? a.timeout_ms : 30000u) * 1000000ull;
        uint32_t spin = 0;
        while ((int32_t)(*from_p - cnt) < 0) {
            if ((++spin & 255u) == 0u && global_ns() - t0 > limit) {
                for (int q = 0; q < a.world; ++q) atomicCAS_system(a.sig[q] + kSigPoison, 0u, 130u);
                break;
            }
        }
    }
    __threadfence_system();
}

int launch_barrier(const smh_exchange_t &exch, cudaStream_t stream)
{
    BarrierArgs a;
    a.world = exch.world;
    a.rank = exch.rank;
    a.timeout_ms = exch.timeout_ms;
    for (int p = 0; p < exch.world; ++p) {
        if (!exch.signal_peer[p]) return set_error(SMH_E_ARG, "exchange signal_peer[%d] is null", p);
        a.sig[p] = (uint32_t *)exch.signal_peer[p];
    }
    barrier_kernel<<<1, 32, 0, stream>>>(a);
    return check_launch("barrier_kernel");
}

}  // namespace smh
