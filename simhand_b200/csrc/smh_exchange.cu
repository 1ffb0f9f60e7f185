// simhand_b200: the two kernels of the peer exchange that are not fused into a compute kernel.
//   push_inputs_kernel  push-based all-gather: every rank writes its packed [z1|z2|joints1|joints2] into slot `rank` of
//                       every peer's gathered-input buffer (plain coalesced 16-byte stores on peer pointers, NVLink)
//   barrier_kernel      device-side barrier between the phases of a step.  Monotonic counters (word 0 = barriers this
//                       rank has entered, word 8 + p = last barrier peer p announced), so the same kernel node can be
//                       replayed from a CUDA graph without host-side epochs.
// The fused parts live in the compute kernels: Dmax push (smh_mpjpe.cu), row-sum all-reduce and gradient reduce-scatter
// in the sweep epilogues (smh_sweep_tc.cu / smh_sweep_fp32.cu).
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

struct PushArgs {
    float *dst[kMaxPeers];
    int world;
};

__global__ void __launch_bounds__(256) push_inputs_kernel(const float4 *__restrict__ src, PushArgs a, int64_t n4)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        for (int p = 0; p < a.world; ++p) reinterpret_cast<float4 *>(a.dst[p])[i] = v;
    }
}

int launch_push_inputs(const smh_exchange_t &exch, const float *local, int64_t floats, cudaStream_t stream)
{
    PushArgs a;
    a.world = exch.world;
    for (int p = 0; p < exch.world; ++p) {
        if (!exch.xin_peer[p]) return set_error(SMH_E_ARG, "exchange xin_peer[%d] is null", p);
        a.dst[p] = (float *)exch.xin_peer[p] + (int64_t)exch.rank * floats;
    }
    const int64_t n4 = floats / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    push_inputs_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(local), a, n4);
    return check_launch("push_inputs_kernel");
}

struct BarrierArgs {
    uint32_t *sig[kMaxPeers];
    int world, rank;
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
        const long long t0 = clock64();
        while ((int32_t)(*from_p - cnt) < 0) {
            if (clock64() - t0 > 4000000000ll) break;      // bounded: a missing peer must not hang the device
        }
    }
    __threadfence_system();
}

int launch_barrier(const smh_exchange_t &exch, cudaStream_t stream)
{
    BarrierArgs a;
    a.world = exch.world;
    a.rank = exch.rank;
    for (int p = 0; p < exch.world; ++p) {
        if (!exch.signal_peer[p]) return set_error(SMH_E_ARG, "exchange signal_peer[%d] is null", p);
        a.sig[p] = (uint32_t *)exch.signal_peer[p];
    }
    barrier_kernel<<<1, 32, 0, stream>>>(a);
    return check_launch("barrier_kernel");
}

}  // namespace smh
