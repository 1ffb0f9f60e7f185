// simhand_b200 K1/K2, tensor-core engines (SMH_ENGINE_TC_FP16 / _TF32 / _BF16): the fused forward / backward sweeps.
//
//   forward  (src/models/utils.py:411-417): S = z z^T on tcgen05 (kind::f16 from the fp16 or bf16 image of z, or kind::tf32
//            from its tf32 image; fp32 accumulate in TMEM); the epilogue turns each S tile into E = exp(S * W / tau) with
//            W * log2(e) / tau built on the fly from the staged distance tile (one FFMA per weight), masks the diagonal and
//            accumulates the row sums.  The 2N x 2N logit matrix never leaves the SM.
//   backward (autograd of :411-426): the same S tile and weights; the epilogue writes
//            G = W E (1/neg_i + 1/neg_j) back into the TMEM columns S came from (packed bf16) and a second
//            tcgen05.mma (kind::f16: A = G from TMEM, B = the staged bf16 z block read MN-major; tf32 operands
//            cannot be read MN-major from a SWIZZLE_128B image) accumulates dzacc_I += G z_J in fp32 in TMEM
//            across the whole strip; the softmax is never materialised.  The backward recomputes S from the bf16
//            image in every engine (its effect on the gradient is ~1e-5 max|g|, below the bf16 rounding of G), which
//            lets one staged bf16 block serve both contractions (K-major for S, MN-major for dz) and frees the
//            shared memory for a deeper prefetch of the distance tiles.
//
// Persistent CTAs (one per SM, 640 threads) walk strips of 128x64 tasks that share a 128-row block.  Warp roles:
//   warp 0      tile producer: cp.async.bulk of the distance-tile halves (fp32, or the 16-bit image: 8 stages), on mbarriers
//   warp 18     operand producer: cp.async.bulk of the z blocks (pre-swizzled SWIZZLE_128B images, smh_prep.cu)
//   warp 1      issues the logit tcgen05.mma and tcgen05.commit
//   warp 19     (backward) issues the value tcgen05.mma: dz += G z
//   warps 2..17 epilogue, two groups of 8 warps taking alternate tasks; one row per thread (TMEM lane == row); two
//               warps of a group share a TMEM lane quadrant and split the 64 columns of a task: tcgen05.ld -> weights ->
//               ex2 -> row sums (forward) or tcgen05.st of G (backward); strip flush of dzacc with red.global.add.v4.f32
// The single-thread roles run their control flow warp-uniformly and guard only the issuing instruction with elect.sync
// (smh_common.cuh: elect_one).  Every pipeline wait is bounded (mbar_wait) so a protocol bug cannot hang the device.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

// Development aid (python -m simhand_b200.build --trace): per-role cycle counters of every sweep CTA, written to the
// rowloss region of the workspace (unused until finalize) and read by tools/trace_sweeps.py.  Compiled out otherwise.
#ifdef SMH_TRACE
#define TR_DECL(n) long long tr_[n] = {}; long long tr_t_ = clock64(); const long long tr_t0_ = tr_t_
#define TR_LAP(i) do { const long long now_ = clock64(); tr_[i] += now_ - tr_t_; tr_t_ = now_; } while (0)
#define TR_COUNT(i) (++tr_[i])
#define TR_STORE(slot, n) do { for (int i_ = 0; i_ < (n); ++i_) trace[(blockIdx.x * 4 + (slot)) * 8 + i_] = tr_[i_]; \
                               trace[(blockIdx.x * 4 + (slot)) * 8 + 7] = clock64() - tr_t0_; } while (0)
#else
#define TR_DECL(n)
#define TR_LAP(i)
#define TR_COUNT(i)
#define TR_STORE(slot, n)
#endif

constexpr int kEpiGroups = 2;                            // epilogue groups take alternate tasks
constexpr int kGroupWarps = 8;                           // 4 TMEM lane quadrants x 2 column halves
constexpr int kEpiWarps = kEpiGroups * kGroupWarps;      // 16
constexpr int kTcThreads = 64 + 32 * kEpiWarps + 64;     // 640: tile producer, logit MMA, 16 epilogue, operand producer, value MMA
constexpr int kSBufs = 4;                                // S / G buffers in TMEM (64 columns each)
constexpr int kMaxBStages = 4;                           // z blocks come from L2
constexpr int kMaxDStages = 8;                           // MPJPE tile pieces come from HBM: deeper prefetch
constexpr int kNumBars = 58;
constexpr int kCtlBytes = kNumBars * 8 + kMaxDStages * 16 + 16;   // barriers, staged task records, TMEM base

// Q16: the distance tiles are the 16-bit image (SMH_DIMS_Q16_TILES): half the bytes per task, twice the stages in flight
template <bool SBF16, bool Q16>
struct TcCfg {
    static constexpr int kABytes = SBF16 ? kTile * kD * 2 : kTile * kD * 4;        // row block of z
    static constexpr int kBBytes = SBF16 ? kTaskN * kD * 2 : kTaskN * kD * 4;      // column block of z
    static constexpr int kBStages = SBF16 ? 4 : 2;
    static constexpr int kDBytes = Q16 ? 16384 : 32768;                            // the half of a stored tile a task needs
    static constexpr int kDStages = Q16 ? (SBF16 ? 8 : 4) : (SBF16 ? 4 : 3);
    static constexpr int kSmem = 1024 /*align slack*/ + kABytes + kBStages * kBBytes + kDStages * kDBytes + kCtlBytes;
    static_assert(kSmem <= 232448, "sweep exceeds the 227 KB shared-memory limit");
};

struct TcBars {
    uint64_t full_b[kMaxBStages], empty_b[kMaxBStages], empty_d[kMaxDStages];
    // One "tile landed" barrier per (stage, consuming epilogue group).  A parity wait can only tell adjacent phases
    // apart, so a barrier must be waited on by a single party that sees every one of its phases; with an odd number
    // of stages the fills of a stage alternate between the two groups.
    uint64_t full_d[kMaxDStages][kEpiGroups];
    uint64_t a_full, a_empty;
    uint64_t sg_full[kSBufs], sg_empty[kSBufs], g_ready[kSBufs];
    uint64_t dz_full[2], dz_empty[2];        // two gradient accumulators in TMEM: strip k uses k % 2
};
static_assert(sizeof(TcBars) <= kNumBars * 8, "barrier block too small");

// Shared-memory reads of the staged MPJPE tile by 32-bit shared address (the generic-pointer form costs an address
// translation per load).
__device__ __forceinline__ void lds_f2x2(uint32_t addr, f2 &a, f2 &b)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
// one 16-bit fixed-point distance as a float: LDS.U16 zero-extends, I2FP.F32.U32 converts on the integer pipe (one
// instruction, where "or 2^23's exponent, then subtract 2^23" costs an integer instruction plus half a packed FADD2 = two
// dispatch cycles: the sweeps are bound by the dispatch port)
__device__ __forceinline__ float lds_q16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return __uint2float_rn(v);
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// The 32 columns [chunk * 32, chunk * 32 + 32) of one task for one row; v[] holds S on entry.
//   forward : rowsum += E,  E = 2^(S * wk),  wk = W * k2 = k2 - D * (k2 / Dmax)      (one FFMA per weight)
//   backward: pk[] = bf16x2 of G' = wk * E * (1/neg_i + 1/neg_j) = k2 * G; the strip flush multiplies by 1 / k2
// The tensor-core engines do not need the correctly rounded W of the fp32 engine (the logits carry 2^-11 operand
// rounding): the fused form is within 1 ulp of k2 in absolute terms.  Packed f32x2 arithmetic throughout.
// negc2 / k2c2: slope and offset of wk in the staged value.  Unit weights: slope 0 and the tile is not read.
// Materialised weights (the tile holds W itself): slope k2, offset 0, and, W not being symmetric in general, the
// backward visits a tile once for the row term (cs2 = 0: G = W_ij E_ij / neg_i) and once transposed for the column
// term (rni = 0, cs2 = 1: G = W_ji E_ji / neg_j).  Q16: the staged tile is the 16-bit image (SMH_DIMS_Q16_TILES).
// SIG: non_linear weights (utils.py:346), W = 1 / (1 + exp(lambda (D - mean D))).  A template parameter, not a run-time
// flag: as a flag ptxas predicated the two extra MUFU per element into the linear path (backward sweep XU pipe 21 % ->
// 70 % busy), and an out-of-line helper cost more in register traffic than it saved.
// (Tried and not faster: computing all 32 weights before the wait on the S buffer -- the two-phase form overlaps the MMA
// wait with the tile reads but gains nothing and spills in the backward sweep.)
template <bool BWD, bool TRANSPOSED, bool MASKED, bool Q16, bool SIG>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], uint32_t dstage_s, int r,
                                               int chunk, int gi, int gj0, int m, bool diagonal, f2 negc2, f2 k2c2,
                                               float rni, f2 cs2, bool no_tile, float k2,
                                               const float *__restrict__ rn, f2 (&rowsum)[2])
{
    uint32_t ta[8];                                   // transposed reads: one address per (row & 7) XOR pattern
    uint32_t dbase = 0, rx = 0;
    if (TRANSPOSED) {
        // thread = stored column r: 16-byte slot (c, stored row ^ (c & 7)) with c = r / 4 (fp32) or r / 8 (16-bit image)
        const uint32_t c = (uint32_t)r >> (Q16 ? 3 : 2), x = c & 7u;
        const uint32_t sub = Q16 ? ((uint32_t)r & 7u) * 2u : ((uint32_t)r & 3u) * 4u;
        const uint32_t base = dstage_s + c * 1024u + sub + (uint32_t)chunk * 512u;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) ta[k] = base + ((k ^ x) << 4);
    } else {
        // thread = row r; per 64-row half the staged piece holds 16 (fp32: 4 columns each) or 8 (16-bit: 8 columns each)
        // column groups of 1 KiB; a 32-column chunk is 8 resp. 4 of them
        dbase = dstage_s + ((uint32_t)r >> 6) * (Q16 ? 8192u : 16384u) + (uint32_t)chunk * (Q16 ? 4096u : 8192u);
        rx = ((uint32_t)r & 63u) << 4;
    }
    const f2 rni2 = pack2(rni, rni);
    const f2 magic2 = pack2(8388608.0f, 8388608.0f);
    f2 dnext01 = 0, dnext23 = 0;                      // 16-bit image: one 16-byte read serves two q steps
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int jl = chunk * 32 + q * 4;                       // first of 4 columns inside the task
        f2 d01 = 0, d23 = 0;                                     // unit weights: the tile is never read (0.0f x2)
        if (no_tile) {
        } else if (!TRANSPOSED) {
            if (!Q16) {
                // column group c4 = chunk * 8 + q of the staged half: its low 3 bits (q) drive the XOR
                lds_f2x2(dbase + (uint32_t)q * 1024u + (rx ^ ((uint32_t)q << 4)), d01, d23);
            } else if ((q & 1) == 0) {
                // column group c8 = chunk * 4 + q / 2: eight 16-bit values; 0x4B000000 | q is the float 2^23 + q
                const uint32_t g = (uint32_t)(q >> 1), c8l = (uint32_t)chunk * 4u + g;
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                             : "r"(dbase + g * 1024u + (rx ^ ((c8l & 7u) << 4))));
                d01 = sub2(pack2(__uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7610)),
                                 __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7632))), magic2);
                d23 = sub2(pack2(__uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7610)),
                                 __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7632))), magic2);
                dnext01 = sub2(pack2(__uint_as_float(__byte_perm(w2, 0x4B000000u, 0x7610)),
                                     __uint_as_float(__byte_perm(w2, 0x4B000000u, 0x7632))), magic2);
                dnext23 = sub2(pack2(__uint_as_float(__byte_perm(w3, 0x4B000000u, 0x7610)),
                                     __uint_as_float(__byte_perm(w3, 0x4B000000u, 0x7632))), magic2);
            } else {
                d01 = dnext01;
                d23 = dnext23;
            }
        } else {
            const uint32_t off = (uint32_t)(q >> 1) * 128u;      // stored rows jl .. jl + 3: bits 3.. of the row index
            const int k0 = (q & 1) * 4;
            if (!Q16) {
                d01 = pack2(lds_f32(ta[k0] + off), lds_f32(ta[k0 + 1] + off));
                d23 = pack2(lds_f32(ta[k0 + 2] + off), lds_f32(ta[k0 + 3] + off));
            } else {
                d01 = pack2(lds_q16(ta[k0] + off), lds_q16(ta[k0 + 1] + off));
                d23 = pack2(lds_q16(ta[k0 + 2] + off), lds_q16(ta[k0 + 3] + off));
            }
        }
        f2 wk01 = fma2(d01, negc2, k2c2), wk23 = fma2(d23, negc2, k2c2);
        if (SIG) {
            // the fma above produced lambda log2(e) (D - mean D)
            float t[4];
            unpack2(wk01, t[0], t[1]);
            unpack2(wk23, t[2], t[3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = k2 * rcp_approx(1.0f + ex2_approx(t[u]));
            wk01 = pack2(t[0], t[1]);
            wk23 = pack2(t[2], t[3]);
        }
        f2 rs01 = 0, rs23 = 0;
        if (BWD) {
            const float4 rnj = __ldg(reinterpret_cast<const float4 *>(rn + gj0 + jl));
            rs01 = fma2(pack2(rnj.x, rnj.y), cs2, rni2);        // cs = 1: 1/neg_i + 1/neg_j (exact, as an add)
            rs23 = fma2(pack2(rnj.z, rnj.w), cs2, rni2);
        }
        const f2 a01 = mul2(pack2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), wk01);
        const f2 a23 = mul2(pack2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), wk23);
        float a[4], e[4];
        unpack2(a01, a[0], a[1]);
        unpack2(a23, a[2], a[3]);
#pragma unroll
        for (int u = 0; u < 4; ++u) e[u] = ex2_approx(a[u]);
        bool valid[4];
        if (MASKED) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int gj = gj0 + jl + u;
                valid[u] = (gi < m) && (gj < m) && !(diagonal && gi == gj);
                e[u] = valid[u] ? e[u] : 0.f;
            }
        }
        if (!BWD) {
            rowsum[0] = add2(rowsum[0], pack2(e[0], e[1]));
            rowsum[1] = add2(rowsum[1], pack2(e[2], e[3]));
        } else {
            const f2 g01 = mul2(mul2(wk01, pack2(e[0], e[1])), rs01);
            const f2 g23 = mul2(mul2(wk23, pack2(e[2], e[3])), rs23);
            float g[4];
            unpack2(g01, g[0], g[1]);
            unpack2(g23, g[2], g[3]);
            if (MASKED) {
#pragma unroll
                for (int u = 0; u < 4; ++u) g[u] = valid[u] ? g[u] : 0.f;     // also drops NaN from padded distances
            }
            pk[2 * q] = pack_bf16x2(g[0], g[1]);
            pk[2 * q + 1] = pack_bf16x2(g[2], g[3]);
        }
    }
}

// BWD: backward sweep (always reads the bf16 image).  SBF16: logits from a 16-bit image `zb` (bf16, or fp16 in the
// forward sweep of the fp16 engine: the caller passes that image and the matching instruction descriptor idesc1),
// else from the tf32 image `zt`.
// FUSED: the exchange of the sharded step rides in this kernel's head and tail (smh_shard.cu); a separate instantiation, so
// the single-GPU kernel carries none of it.
// PLAIN: linear weights from the joints (wmode 0), the default; the other weightings (unit, materialised, non_linear) share
// the !PLAIN instantiation and tell each other apart at run time.
template <bool BWD, bool SBF16, bool Q16, bool FUSED, bool PLAIN>
__global__ void __launch_bounds__(kTcThreads, 1)
sweep_tc_kernel(const int4 *__restrict__ tasks, const int2 *__restrict__ strips, const int *__restrict__ cta_ptr,
                const float *__restrict__ zt, const uint16_t *__restrict__ zb, const void *__restrict__ dist,
                float *__restrict__ rn, const __grid_constant__ Peers peers, const __grid_constant__ Peers xp, Stats *__restrict__ stats, int m, int n, int n_local,
                float k2, float inv_k2, int wmode, float lambda_neg, uint32_t idesc1, long long *trace)
{
    static_assert(!BWD || SBF16, "the backward sweep stages only the bf16 image");
    using Cfg = TcCfg<SBF16, Q16>;
    constexpr int kABytes = Cfg::kABytes, kBBytes = Cfg::kBBytes, kDStages = Cfg::kDStages, kBStages = Cfg::kBStages;
    constexpr int kDBytes = Cfg::kDBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sA = sm;
    unsigned char *sB = sA + kABytes;
    unsigned char *sD = sB + kBStages * kBBytes;
    unsigned char *ctl = sD + kDStages * kDBytes;
    TcBars *bars = reinterpret_cast<TcBars *>(ctl);
    int4 *task_slot = reinterpret_cast<int4 *>(ctl + kNumBars * 8);          // task record travelling with a D stage
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ctl + kNumBars * 8 + kMaxDStages * 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int s_begin = cta_ptr[blockIdx.x], s_end = cta_ptr[blockIdx.x + 1];      // this CTA's strips
    uint32_t *fail = &stats->fail_site;
    constexpr uint32_t kTmemCols = BWD ? 512u : 256u;
    constexpr uint32_t kDzCol = kSBufs * kTaskN;                 // 256: gradient accumulator behind the S buffers

    if (threadIdx.x == 0) {
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(&bars->full_b[i], 1);
            mbar_init(&bars->empty_b[i], 1);
        }
        for (int i = 0; i < kDStages; ++i) {
            mbar_init(&bars->full_d[i][0], 1);
            mbar_init(&bars->full_d[i][1], 1);
            mbar_init(&bars->empty_d[i], kGroupWarps);
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int i = 0; i < kSBufs; ++i) {
            mbar_init(&bars->sg_full[i], 1);
            mbar_init(&bars->sg_empty[i], BWD ? 1 : kGroupWarps);
            mbar_init(&bars->g_ready[i], kGroupWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->dz_full[i], 1);
            mbar_init(&bars->dz_empty[i], kEpiWarps);
        }
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
    // Fused exchange (xp: the real ranks; `peers` stays the local view the accumulators are addressed through): the forward
    // sweep starts once every rank has delivered Dmax (stage 2); the backward sweep follows rn_fused_kernel, which has
    // already waited for the row sums.  gs: the step's combined scalars.
    uint32_t epoch = 0u;
    const Stats *gs = stats;
    PhaseClock clk(xp, BWD ? 4 : 2, FUSED);
    if (FUSED) {
        epoch = xp.my_sig()[kSigEpoch];
        if (!BWD) {
            stage_wait(xp, 2, epoch);            // Dmax of every rank
            stage_wait(xp, kStageZ, epoch);      // z images of every rank (shipped on a parallel branch under the MPJPE kernel)
        } else {
            // the row sums: neg_i = rank-ordered sum of the partials every rank delivered (stage 3), 1 / neg_i.  Every CTA
            // reduces a slice, a grid barrier publishes the vector (this used to be a launch of its own)
            stage_wait(xp, 3, epoch);
            const int mp = ((m + kTile - 1) / kTile) * kTile;
            float *neg_tot = xp.neg(xp.rank);
            const float *parts = xp.negparts(xp.rank);
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mp; i += gridDim.x * blockDim.x) {
                float v = 0.f;
                for (int p = 0; p < xp.world; ++p) v += __ldcg(parts + (int64_t)p * mp + i);
                neg_tot[i] = v;
                rn[i] = (i < m) ? __frcp_rn(v) : 0.f;
            }
            __syncthreads();
            if (threadIdx.x == 0) grid_barrier(xp, epoch * 8u + 2u);       // tags grow: fwd tail 1, bwd head 2, bwd tail 3
            __syncthreads();
        }
        gs = xp.gstats(xp.rank, epoch);
        clk.lap();                               // [0] stage wait (+ row-sum reduction)
    }

    if (warp == 0) {
        // ------------------------------------------------------------------ tile producer (MPJPE pieces, HBM)
        int dst = 0;
        uint32_t dph = 0, seq = 0;
        TR_DECL(7);
        for (int s = s_begin; s < s_end; ++s) {
            const int2 strip = strips[s];
            int4 task = tasks[strip.x];
            for (int ti = strip.x; ti < strip.y; ++ti, ++seq) {
                const int4 next = (ti + 1 < strip.y) ? tasks[ti + 1] : task;      // prefetch the next record
                uint64_t *full = &bars->full_d[dst][seq & 1u];                      // the group that takes task `seq`
                TR_LAP(0);
                mbar_wait(&bars->empty_d[dst], dph ^ 1u, fail, 3);                  // all lanes: warp-uniform control flow
                TR_LAP(1);                                                          // [1] waiting for a free tile stage
                if (elect_one()) {
                    task_slot[dst] = task;
                    mbar_arrive_expect_tx(full, kDBytes);
                    // fp32 tile: 64 KiB, 1 KiB per (row half, 4-column group); 16-bit image: 32 KiB, 1 KiB per (row half,
                    // 8-column group).  Either way a task needs one half of it:
                    const unsigned char *tile = reinterpret_cast<const unsigned char *>(dist) +
                                                (int64_t)task.z * (2 * kDBytes);
                    const int half = task.y & 1;
                    if (task.w & kTaskTransposed) {
                        // stored rows half*64 .. +63, all 128 stored columns: one contiguous slab
                        bulk_g2s(sD + dst * kDBytes, tile + half * kDBytes, kDBytes, full);
                    } else {
                        // stored columns half*64 .. +63: one slab per 64-row half
                        bulk_g2s(sD + dst * kDBytes, tile + half * (kDBytes / 2), kDBytes / 2, full);
                        bulk_g2s(sD + dst * kDBytes + kDBytes / 2, tile + kDBytes + half * (kDBytes / 2), kDBytes / 2, full);
                    }
                }
                task = next;
                if (++dst == kDStages) { dst = 0; dph ^= 1u; }
            }
        }
        if (lane == 0) TR_STORE(0, 7);
    } else if (warp == 2 + kEpiWarps) {
        // ------------------------------------------------------------------ operand producer (z blocks, L2)
        {
            int bst = 0;
            uint32_t bph = 0, a_ph = 0;
            TR_DECL(7);
            for (int s = s_begin; s < s_end; ++s) {
                const int2 strip = strips[s];
                int4 task = tasks[strip.x];
                const int I = task.x;
                TR_LAP(0);
                mbar_wait(&bars->a_empty, a_ph ^ 1u, fail, 1);
                TR_LAP(1);                                                          // [1] waiting for the A buffer
                if (elect_one()) {
                mbar_arrive_expect_tx(&bars->a_full, kABytes);
                if (SBF16) {
                    // two 64-row bf16 blocks -> two 64-column boxes of [128 rows][128 B]
#pragma unroll
                    for (int db = 0; db < 2; ++db) {
                        bulk_g2s(sA + db * 16384, zb + (int64_t)(2 * I) * kBlockFloats + db * 4096, 8192, &bars->a_full);
                        bulk_g2s(sA + db * 16384 + 8192, zb + (int64_t)(2 * I + 1) * kBlockFloats + db * 4096, 8192,
                                 &bars->a_full);
                    }
                } else {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        bulk_g2s(sA + kb * 16384, zt + (int64_t)(2 * I) * kBlockFloats + kb * 2048, 8192, &bars->a_full);
                        bulk_g2s(sA + kb * 16384 + 8192, zt + (int64_t)(2 * I + 1) * kBlockFloats + kb * 2048, 8192,
                                 &bars->a_full);
                    }
                }
                }
                a_ph ^= 1u;
                for (int ti = strip.x; ti < strip.y; ++ti) {
                    const int4 next = (ti + 1 < strip.y) ? tasks[ti + 1] : task;
                    TR_LAP(0);
                    mbar_wait(&bars->empty_b[bst], bph ^ 1u, fail, 2);
                    TR_LAP(2);                                                      // [2] waiting for a free z-block stage
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&bars->full_b[bst], kBBytes);
                        if (SBF16)
                            bulk_g2s(sB + bst * kBBytes, zb + (int64_t)task.y * kBlockFloats, kBBytes, &bars->full_b[bst]);
                        else
                            bulk_g2s(sB + bst * kBBytes, zt + (int64_t)task.y * kBlockFloats, kBBytes, &bars->full_b[bst]);
                    }
                    task = next;
                    if (++bst == kBStages) { bst = 0; bph ^= 1u; }
                }
            }
            if (lane == 0) TR_STORE(1, 7);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one elected lane issues; the
        // control flow and every operand stay warp-uniform, see elect_one)
        {
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            // descriptors = base + (byte offset / 16) in the start-address field (umma_desc_lo): one add per MMA operand
            const uint32_t a_lo = umma_desc_lo(sA_u, 16), b_lo0 = umma_desc_lo(sB_u, 16), d_hi = umma_desc_hi_sw128(1024);
            uint32_t a_ph = 0;
            uint32_t seq = 0;                       // logit MMAs issued so far by this CTA
            TR_DECL(7);
            // logit contraction of task `seq` into S buffer seq % kSBufs (operands and buffer are known to be ready)
            auto mma1 = [&](bool last_of_strip) {
                const uint32_t sb = seq % kSBufs, bst = seq % kBStages;
                const uint32_t b_lo = b_lo0 + bst * (uint32_t)(kBBytes >> 4);
                tc_fence_after();
                if (elect_one()) {
                if (SBF16) {
                    // K-major 16-bit: 64 columns (128 B) per box, 16 columns (32 B) per K step
#pragma unroll
                    for (int db = 0; db < 2; ++db) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t adesc = umma_desc_pack(a_lo + (uint32_t)(db * (16384 >> 4) + ks * 2), d_hi);
                            const uint64_t bdesc = umma_desc_pack(b_lo + (uint32_t)(db * (8192 >> 4) + ks * 2), d_hi);
                            tc_mma_ss_f16(tmem_base + sb * kTaskN, adesc, bdesc, idesc1, (db | ks) ? 1u : 0u);
                        }
                    }
                } else {
                    // K-major tf32: 32 columns (128 B) per box, 8 columns (32 B) per K step
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t adesc = umma_desc_pack(a_lo + (uint32_t)(kb * (16384 >> 4) + ks * 2), d_hi);
                            const uint64_t bdesc = umma_desc_pack(b_lo + (uint32_t)(kb * (8192 >> 4) + ks * 2), d_hi);
                            tc_mma_ss_tf32(tmem_base + sb * kTaskN, adesc, bdesc, idesc1, (kb | ks) ? 1u : 0u);
                        }
                    }
                }
                tc_commit(&bars->sg_full[sb]);
                if (!BWD) tc_commit(&bars->empty_b[bst]);
                if (last_of_strip) tc_commit(&bars->a_empty);           // every logit MMA of the strip has read sA
                }
                __syncwarp();
            };
            {
                for (int s = s_begin; s < s_end; ++s) {
                    const int2 strip = strips[s];
                    TR_LAP(0);
                    mbar_wait(&bars->a_full, a_ph, fail, 6);
                    TR_LAP(1);                                                      // [1] waiting for A
                    a_ph ^= 1u;
#pragma unroll 1
                    for (int ti = strip.x; ti < strip.y; ++ti, ++seq) {
                        mbar_wait(&bars->full_b[seq % kBStages], (seq / kBStages) & 1u, fail, 7);
                        TR_LAP(2);                                                  // [2] waiting for the z block
                        mbar_wait(&bars->sg_empty[seq % kSBufs], ((seq / kSBufs) & 1u) ^ 1u, fail, 8);
                        TR_LAP(3);                                                  // [3] waiting for a free S buffer
                        mma1(ti + 1 == strip.y);
                        TR_LAP(4);                                                  // [4] issuing the logit MMAs
                    }
                }
            }
            if (lane == 0) TR_STORE(2, 7);
        }
    } else if (BWD && warp == 3 + kEpiWarps) {
        // ------------------------------------------------------------------ value-MMA issuer (backward only).  A second
        // issuing warp: the logit MMAs (warp 1) wait for operands and S buffers, the value MMAs for G from the epilogue;
        // with one thread doing both, each kind queued behind the other's waits (first a fixed interleave, then a polling
        // loop that spent half its time polling).  tcgen05.commit tracks the MMAs of the committing thread, and every
        // dependency between the two streams goes through an mbarrier, so they need no ordering between them.
        constexpr uint32_t idesc2 = umma_idesc_bf16(kTile, kD, 0, 1);
        const uint32_t sB_u = smem_u32(sB);
        const uint32_t vb_lo0 = umma_desc_lo(sB_u, 8192), vd_hi = umma_desc_hi_sw128(1024);
        uint32_t q = 0;
        for (int s = s_begin; s < s_end; ++s) {
            const int2 strip = strips[s];
            // gradient accumulator of this strip: the epilogue drains the other one while this strip's MMAs run
            const uint32_t sk = (uint32_t)(s - s_begin), acc = sk & 1u;
            const uint32_t dz_tmem = tmem_base + kDzCol + acc * (uint32_t)kD;
#pragma unroll 1
            for (int ti = strip.x; ti < strip.y; ++ti, ++q) {
                const uint32_t sbq = q % kSBufs, bstq = q % kBStages;
                const bool first = ti == strip.x, last = ti + 1 == strip.y;
                mbar_wait(&bars->g_ready[sbq], (q / kSBufs) & 1u, fail, 4);
                if (first) mbar_wait(&bars->dz_empty[acc], ((sk >> 1) & 1u) ^ 1u, fail, 5);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kTaskN / 16; ++ks) {
                        // MN-major bf16 B: 64-element (128 B) atoms along d at LBO = 8 KiB, 8-row K groups at SBO = 1 KiB,
                        // 16 sample rows (2 KiB) per K step.  A = packed bf16 G: columns [0,16) hold task columns 0..31,
                        // columns [32,48) hold task columns 32..63 (each epilogue half overwrites its own S columns).
                        const uint64_t bdesc = umma_desc_pack(vb_lo0 + bstq * (uint32_t)(kBBytes >> 4) + (uint32_t)(ks * (2048 >> 4)), vd_hi);
                        const uint32_t a_col = (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u;
                        tc_mma_ts_f16(dz_tmem, tmem_base + sbq * kTaskN + a_col, bdesc, idesc2,
                                      (first && ks == 0) ? 0u : 1u);
                    }
                    tc_commit(&bars->sg_empty[sbq]);
                    tc_commit(&bars->empty_b[bstq]);
                    if (last) tc_commit(&bars->dz_full[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 2 && warp < 2 + kEpiWarps) {
        // ------------------------------------------------------------------ epilogue (warps 2..17)
        const int e = warp - 2;
        const int group = e >> 3;                      // takes the tasks with (sequence number % 2) == group
        const int half = (e >> 2) & 1;                 // which 32 of the task's 64 columns
        const int w4 = warp & 3;                       // TMEM lane quadrant this warp may touch
        const int r = w4 * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
        const float dmax = __uint_as_float(gs->dmax_bits);
        // wk = W * k2 = k2 - D * (k2 / Dmax)  (Dmin = +0, the diagonal); unit weights: wk = k2
        // Compile-time false for the default weighting: as run-time flags these three put a branch and two register-pair
        // clears into every 4-element step of the epilogue (CS2R 0.53, BSSY + BSYNC 0.4 instructions per element in the ncu
        // counts) -- on a kernel bound by the dispatch port: forward 146 -> 128 us, backward 206 -> 188 us at 2N = 16384.
        static_assert(PLAIN || !Q16, "the 16-bit image exists for the linear weights from the joints only");
        const bool no_tile = !PLAIN && wmode == 1, dense = !PLAIN && wmode == 2, sigmoid = !PLAIN && wmode == 3;
        float negc = no_tile ? 0.f : (dense ? k2 : -__fdiv_rn(k2, dmax));
        float addc = dense ? 0.f : k2;
        if (sigmoid) {
            // exponent of the sigmoid in base 2: lambda log2(e) (D - mean D), mean over all M^2 ordered pairs
            double dsum = stats->dsum;
            if (FUSED) {                     // rank-ordered sum of the parts every rank delivered with stage 2
                const double *parts = reinterpret_cast<const double *>(xp.lossparts(xp.rank) + 16);
                dsum = 0.0;
                for (int p = 0; p < xp.world; ++p) dsum += __ldcg(parts + p);
            }
            const float mu = (float)(dsum / ((double)m * (double)m));
            negc = lambda_neg * 1.4426950408889634f;
            addc = -mu * negc;
        }
        if (Q16 && !no_tile) {
            // the staged value is q = D * qscale: fold 1 / qscale into the slope (wk = k2 - q k2 / (Dmax qscale))
            const float qs = q16_scale(__uint_as_float(gs->dbound_bits));
            negc = qs > 0.f ? __fdiv_rn(negc, qs) : negc;
        }
        const f2 negc2 = pack2(negc, negc), k2c2 = pack2(addc, addc);
        const uint32_t sD_s = smem_u32(sD), task_slot_s = smem_u32(task_slot);
        uint32_t seq = 0;
        f2 rowsum[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
        // Backward: the flush of a strip's gradient accumulator is DEFERRED by one strip.  Waiting for the strip's last value
        // MMA right after its last task drained the whole pipeline at every strip boundary (a strip cost ~4 task times:
        // profiles/r02_strip_cost.txt); with two accumulators the epilogue goes straight on to the next strip and drains
        // the finished accumulator after that strip's tasks, when its MMAs have long completed.
        int pend_row = -1;                            // first row of the strip whose accumulator is still to be drained
        uint32_t pend_k = 0;                          // its index among this CTA's strips
        auto flush_dz = [&](int row_first, uint32_t sk) {
            const uint32_t acc = sk & 1u;
            mbar_wait(&bars->dz_full[acc], (sk >> 1) & 1u, fail, 11);
            tc_fence_after();
            const int gi2 = row_first + r;
            const bool ok2 = gi2 < m;
            // fused reduce-scatter: the gradient rows are added straight into the owning rank's accumulator
            float *orow = dz_row_ptr(peers, ok2 ? gi2 : 0, n, n_local);
            const int chunk = group * 2 + half;       // every epilogue warp drains 32 of the 128 columns of its lane quadrant
            uint32_t dv[32];
            tc_ld32(lane_addr + kDzCol + acc * (uint32_t)kD + chunk * 32, dv);
            tc_wait_ld();
            if (ok2) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    red_add_v4(orow + chunk * 32 + q * 4, inv_k2 * __uint_as_float(dv[4 * q]),
                               inv_k2 * __uint_as_float(dv[4 * q + 1]), inv_k2 * __uint_as_float(dv[4 * q + 2]),
                               inv_k2 * __uint_as_float(dv[4 * q + 3]));        // the accumulator holds k2 * dz
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->dz_empty[acc]);
        };
        TR_DECL(7);
        for (int s = s_begin; s < s_end; ++s) {
            const int2 strip = strips[s];
            int row_block = -1;
            float rni = 0.f;
            // this group's tasks only: every other one, starting at the first whose sequence number has the group's parity
            const uint32_t seq_end = seq + (uint32_t)(strip.y - strip.x);
            for (seq += (seq ^ (uint32_t)group) & 1u; seq < seq_end; seq += 2) {
                const uint32_t dst = seq % kDStages, sb = seq % kSBufs;
                // fills of a stage seen by this group: every fill (even stage count) or every other one (odd)
                const uint32_t my_fill = (kDStages & 1) ? (seq / kDStages) >> 1 : seq / kDStages;
                TR_LAP(0);
                mbar_wait(&bars->full_d[dst][group], my_fill & 1u, fail, 10);
                TR_LAP(1);                                                          // [1] waiting for the tile
                int4 task;                            // the record travelling with the stage: one LDS.128 by shared address
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(task.x), "=r"(task.y), "=r"(task.z), "=r"(task.w)
                             : "r"(task_slot_s + dst * 16u));
                const int gi = task.x * kTile + r;
                const bool row_ok = gi < m;
                const int gj0 = task.y * kTaskN;
                const bool transposed = task.w & kTaskTransposed;
                const bool masked = task.w & (kTaskDiagonal | kTaskRagged);
                const bool diagonal = task.w & kTaskDiagonal;
                if (row_block < 0) {
                    row_block = task.x;
                    if (BWD) rni = row_ok ? rn[gi] : 0.f;
                }
                const float rni_t = (dense && transposed) ? 0.f : rni;
                const float cs = (dense && !transposed) ? 0.f : 1.f;
                const f2 cs2 = pack2(cs, cs);
                TR_LAP(0);
                mbar_wait(&bars->sg_full[sb], (seq / kSBufs) & 1u, fail, 9);
                TR_LAP(2);                                                          // [2] waiting for S
                tc_fence_after();
                const uint32_t dstage = sD_s + dst * kDBytes;
                uint32_t v[32], pk[16];
                tc_ld32(lane_addr + sb * kTaskN + half * 32, v);
                tc_wait_ld();
                TR_LAP(3);                                                          // [3] tcgen05.ld + wait
#define SMH_EPI(T, M, S)                                                                                              \
    epilogue_chunk<BWD, T, M, Q16, S>(v, pk, dstage, r, half, gi, gj0, m, diagonal, negc2, k2c2, rni_t, cs2, no_tile, k2, \
                                      rn, rowsum)
                if (sigmoid) {                       // non_linear weights: the masked form serves every task
                    if (transposed) SMH_EPI(true, true, true);
                    else SMH_EPI(false, true, true);
                } else if (masked) {
                    if (transposed) SMH_EPI(true, true, false);
                    else SMH_EPI(false, true, false);
                } else {
                    if (transposed) SMH_EPI(true, false, false);
                    else SMH_EPI(false, false, false);
                }
#undef SMH_EPI
                if (BWD) {
                    // G' as packed bf16x2 over the first half of this warp's own S columns
                    tc_st16(lane_addr + sb * kTaskN + half * 32, pk);
                    tc_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(BWD ? &bars->g_ready[sb] : &bars->sg_empty[sb]);
                    mbar_arrive(&bars->empty_d[dst]);
                }
                TR_LAP(4);                                                          // [4] weights, exp, sums / G, arrive
                TR_COUNT(6);
            }
            seq = seq_end;
            TR_LAP(0);
            // strip flush: both groups hold partial results for the strip's row block
            const int gi = (row_block < 0 ? 0 : row_block) * kTile + r;
            const bool row_ok = row_block >= 0 && gi < m;
            if (!BWD) {
                // fused all-reduce(SUM): the row sums go into every rank's `neg` (peer pointers over NVLink)
                if (row_ok) {
                    float r0, r1, r2, r3;
                    unpack2(rowsum[0], r0, r1);
                    unpack2(rowsum[1], r2, r3);
                    const float part = (r0 + r1) + (r2 + r3);
                    for (int p = 0; p < peers.world; ++p) atomicAdd(peers.neg(p) + gi, part);
                }
                rowsum[0] = rowsum[1] = pack2(0.f, 0.f);
            } else {
                if (pend_row >= 0) flush_dz(pend_row, pend_k);       // the PREVIOUS strip's accumulator
                pend_row = tasks[strip.x].x * kTile;
                pend_k = (uint32_t)(s - s_begin);
            }
            TR_LAP(5);                                                              // [5] strip flush
        }
        if (BWD && pend_row >= 0) flush_dz(pend_row, pend_k);         // the CTA's last strip
        if (e == 0 && lane == 0) TR_STORE(3, 7);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, kTmemCols);
    if (FUSED) {
        // Tail of the fused exchange: once every CTA of the rank has flushed, the CTAs ship the rank's partial results --
        // forward: all-gather of the partial row sums into slot `rank` of every rank's negparts; backward: reduce-scatter
        // payload, rows [p * 2 n_local, (p + 1) * 2 n_local) of the gradient accumulator into slot `rank` of rank p's
        // dzparts -- with plain 16-byte stores over NVLink; the rank's last CTA signals the stage.
        uint32_t *sig = xp.my_sig();
        clk.lap();                               // [1] this CTA's share of the sweep
        if (threadIdx.x == 0) grid_barrier(xp, epoch * 8u + (BWD ? 3u : 1u));
        __syncthreads();
        clk.lap();                               // [2] every CTA of the rank has flushed
        if (!BWD) {
            const int mp4 = ((m + kTile - 1) / kTile) * kTile / 4;
            const float4 *src = reinterpret_cast<const float4 *>(xp.neg(xp.rank));
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mp4; i += gridDim.x * blockDim.x) {
                const float4 v = __ldcg(src + i);
                for (int p = 0; p < xp.world; ++p) reinterpret_cast<float4 *>(xp.negparts(p))[(int64_t)xp.rank * mp4 + i] = v;
            }
        } else {
            const int64_t block4 = (int64_t)2 * n_local * kD / 4;
            const int64_t total = block4 * xp.world;
            const float4 *src = reinterpret_cast<const float4 *>(xp.dzacc(xp.rank));
            for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
                const int p = (int)(i / block4);
                reinterpret_cast<float4 *>(xp.dzparts(p))[(int64_t)xp.rank * block4 + (i - (int64_t)p * block4)] = __ldcg(src + i);
            }
        }
        __syncthreads();
        clk.lap();                               // [3] payload stores issued
        if (threadIdx.x == 0) {
            block_release_fence();               // after the CTA barrier: orders every thread's peer stores before the ticket
            clk.lap();                           // [4] fence
            const uint32_t ticket = atomicAdd(sig + kSigTicket, 1u);
            if (ticket == gridDim.x - 1) {
                sig[kSigTicket] = 0u;
                stage_signal(xp, BWD ? 4 : 3, epoch);
            }
        }
    }
}

template <bool BWD, bool SBF16, bool Q16, bool FUSED, bool PLAIN>
static int launch_one(int wmode, const uint16_t *half_image, uint32_t idesc1, const smh_dims_t &dims,
                      const smh_layout_t &lay, const PlanView &plan, const WsView &ws, const Peers &peers,
                      const Peers &xp, float temperature, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    (void)sms;
    const int grid = kNumCtas;                      // the plan is cut for exactly this many CTAs
    const float k2 = 1.4426950408889634f / temperature;
    const float inv_k2 = (float)(0.6931471805599453 * (double)temperature);
    const int n_local = dims.n / dims.world;
    constexpr int smem = TcCfg<SBF16, Q16>::kSmem;
    cudaError_t e = cudaFuncSetAttribute(sweep_tc_kernel<BWD, SBF16, Q16, FUSED, PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error((int)e, "tc sweep smem attr: %s", cudaGetErrorString(e));
    sweep_tc_kernel<BWD, SBF16, Q16, FUSED, PLAIN><<<grid, kTcThreads, smem, stream>>>(plan.tasks, BWD ? plan.strips : plan.strips_fwd,
                                                                   BWD ? plan.cta_ptr : plan.cta_ptr_fwd, ws.zt, half_image,
                                                                   ws.dist, ws.rn, peers, xp, (Stats *)ws.stats, lay.m,
                                                                   dims.n, n_local, k2, inv_k2, wmode, dims.lambda_neg, idesc1,
                                                                        reinterpret_cast<long long *>(ws.rowloss));
    return check_launch("sweep_tc_kernel");
}

int launch_sweep_tc(bool backward, int logit_format, int wmode, const smh_dims_t &dims, const smh_layout_t &lay,
                    const PlanView &plan, const WsView &ws, const Peers &peers, const Peers &xp, float temperature,
                    cudaStream_t stream)
{
    if (lay.n_strips == 0 && !xp.fused) return 0;          // fused exchange: a rank without tasks still signals its stage
    const uint32_t id_bf16 = umma_idesc_bf16(kTile, kTaskN, 0, 0), id_f16 = umma_idesc_f16(kTile, kTaskN, 0, 0),
                   id_tf32 = umma_idesc_tf32(kTile, kTaskN, 0, 0);
    const bool q16 = dims.flags & SMH_DIMS_Q16_TILES;
    if (q16 && wmode != 0) return set_error(SMH_E_MODE, "SMH_DIMS_Q16_TILES: linear weights from the joints only");
#define SMH_SWEEP_F(B, S, F, IMG, ID)                                                                                  \
    (q16 ? launch_one<B, S, true, F, true>(wmode, IMG, ID, dims, lay, plan, ws, peers, xp, temperature, stream)                \
         : (wmode == 0 ? launch_one<B, S, false, F, true>(wmode, IMG, ID, dims, lay, plan, ws, peers, xp, temperature, stream) \
                       : launch_one<B, S, false, F, false>(wmode, IMG, ID, dims, lay, plan, ws, peers, xp, temperature, stream)))
#define SMH_SWEEP(B, S, IMG, ID) (xp.fused ? SMH_SWEEP_F(B, S, true, IMG, ID) : SMH_SWEEP_F(B, S, false, IMG, ID))
    if (backward) return SMH_SWEEP(true, true, ws.zb, id_bf16);
    if (logit_format == 1) return SMH_SWEEP(false, true, ws.zb, id_bf16);
    if (logit_format == 2) return SMH_SWEEP(false, true, ws.zh, id_f16);
    return SMH_SWEEP(false, false, ws.zb, id_tf32);
#undef SMH_SWEEP
#undef SMH_SWEEP_F
}

}  // namespace smh
