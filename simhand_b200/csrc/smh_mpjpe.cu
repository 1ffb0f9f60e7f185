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
// CUDA-core kernel: per pair 21 MUFU and ~190 FMA-pipe lane-operations; the packed FADD2/FMUL2/FFMA2 forms halve the
// instruction count (not the dispatch cycles: see joint_pair).  This is the kernel the roofline in bench.py is quoted on.
#include <cstdlib>

#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

// Fused exchange (smh_shard.cu): head = wait until every rank has delivered its rows (stage 1) and take the step's
// combined scalars; tail = the rank's last CTA pushes Dmax (and the non-finite flag) to every rank and, unless a
// distance-sum pass follows (non_linear weights), signals stage 2.
struct DistCtx {
    uint32_t epoch;
    const Stats *gs;          // scalars the kernel reads (distance bound, domain flags)
};
template <bool FUSED>
__device__ __forceinline__ DistCtx dist_head(Stats *stats, const Peers &peers)
{
    DistCtx c;
    c.epoch = 0u;
    c.gs = stats;
    if (FUSED) {
        c.epoch = peers.my_sig()[kSigEpoch];
        PhaseClock clk(peers, 1);
        stage_wait(peers, 1, c.epoch);
        clk.lap();                               // [0] stage wait
        c.gs = peers.gstats(peers.rank, c.epoch);
    }
    return c;
}
// thread 0 of every CTA, after the CTA's maximum went into stats->dmax_bits / stats->flags (rank-local)
__device__ __forceinline__ void dist_tail(Stats *stats, const Peers &peers, const DistCtx &c, bool signal2)
{
    if (peers.world <= 1) return;
    __threadfence();
    const unsigned ticket = atomicAdd(&stats->ticket2, 1u);
    if (ticket != gridDim.x - 1) return;
    const uint32_t mine = atomicMax(&stats->dmax_bits, 0u);
    if (peers.fused) {
        const uint32_t fl = atomicOr(&stats->flags, 0u);
        for (int p = 0; p < peers.world; ++p) {
            Stats *o = peers.gstats(p, c.epoch);
            atomicMax_system(&o->dmax_bits, mine);
            if (fl) atomicOr_system(&o->flags, fl);
        }
        stats->ticket2 = 0u;
        if (signal2) stage_signal(peers, 2, c.epoch);
        return;
    }
    for (int p = 0; p < peers.world; ++p)
        if (p != peers.rank) atomicMax(&peers.stats(p)->dmax_bits, mine);
    stats->ticket2 = 0u;
    __threadfence_system();
}

// MODE 0: IEEE intrinsics (any input).  MODE 1: branch-free exact forms with the zero guard.  MODE 2: without the
// guard (a coincident joint gives NaN; the caller repairs that pair with MODE 1).
// distances of joints (2p, 2p+1) of one pair of samples
// APPROX (16-bit tile image only): one MUFU.SQRT per joint and no correction step -- the value is about to be rounded to
// 16 bits, so the ~2^-22 relative error of the approximation is invisible, and sqrt(+0) = +0 needs no guard.  The exact
// forms stay for the fp32 tiles (weights API, fp32 engine, exact_weights=True).
//
// What bounds the 16-bit-image kernel (ncu: profiles/r02_ncu_mpjpe_pipes.txt).  Per pair of samples and warp the XU pipe
// needs 8 cycles per MUFU: 168 for 21 joints.  The rest of the pair (differences, squares, sums, 11 broadcast LDS) is ~100
// instructions, half of them packed FADD2 / FMUL2 / FFMA2 -- and a packed instruction holds the SMSP's dispatch port for TWO
// cycles (its pipe time, sm__pipe_fma_cycles_active, is exactly 2 x the packed count + the scalar count), so the pair costs
// ~147 dispatch cycles.  The kernel runs at 190 cycles per pair and warp: XU 89 % busy, dispatch 75 %.
// Bit p of kFmaSqrtMask (p = 0..9: joints 2p, 2p+1; bit 10: joint 20) sends that square root through sqrt2_fma_pipe (integer
// seed + 6 packed FMA-pipe operations, +-7.2e-7 relative) instead of MUFU.SQRT: 8 XU cycles less, 9 dispatch cycles more per
// joint.  The two limits cross between one and two joints; measured (profiles/r02_mpjpe_fma_sqrt.txt, 8256 tiles):
// 0 joints 691 us, 2 joints 679 us, 3: 697, 4: 702, 5: 726, 6: 748, 8: 801.  With 4 joints moved ncu shows XU 71 %, FMA 67 %,
// issue 62 % + the packed instructions once more = 90 % of the dispatch cycles: neither pipe binds any more, the port does.
// Also measured, not faster: scalar FFMA for the moved joints (730 us), 2 CTAs per SM at 126 registers (689 / 680 us with 0 / 4
// joints moved), two rows per thread sharing every LDS (mpjpe_tile_body_2r: 703 us -- the shared-memory instructions are
// not what holds the MUFU queue back).
#ifndef SMH_MPJPE_FMA_SQRT_MASK
#define SMH_MPJPE_FMA_SQRT_MASK 0x040
#endif
constexpr unsigned kFmaSqrtMask = SMH_MPJPE_FMA_SQRT_MASK;
#ifndef SMH_MPJPE_Q16_CTAS
#define SMH_MPJPE_Q16_CTAS 3                    // resident CTAs per SM of the 16-bit-image form (80 registers)
#endif
#ifndef SMH_MPJPE_Q16_ROWS
#define SMH_MPJPE_Q16_ROWS 1                    // rows per thread of the 16-bit-image form (2: mpjpe_tile_body_2r, needs 2 CTAs)
#endif
#ifndef SMH_MPJPE_EXACT_CTAS
#define SMH_MPJPE_EXACT_CTAS 2                  // resident CTAs per SM of the exact form (124 registers)
#endif

// squared distances of joints (2p, 2p+1) of one pair of samples
__device__ __forceinline__ f2 joint_x(const f2 ax, const f2 ay, const float *__restrict__ col, int p)
{
    const float4 b = *reinterpret_cast<const float4 *>(col + 4 * p);   // (bx_2p, bx_2p+1, by_2p, by_2p+1)
    f2 dx = sub2(ax, pack2(b.x, b.y));
    f2 dy = sub2(ay, pack2(b.z, b.w));
    return fma2(dy, dy, mul2(dx, dx));
}
template <int MODE, bool APPROX = false>
__device__ __forceinline__ f2 joint_pair(const f2 ax, const f2 ay, const float *__restrict__ col, int p)
{
    f2 x = joint_x(ax, ay, col, p);
    if (APPROX) {
        if ((kFmaSqrtMask >> p) & 1u) return sqrt2_fma_pipe(x);
        float x0, x1;
        unpack2(x, x0, x1);
        return pack2(sqrt_approx(x0), sqrt_approx(x1));
    }
    if (MODE == 2) return sqrt2_rn_fast_nz(x);
    if (MODE == 1) return sqrt2_rn_fast(x);
    float x0, x1;
    unpack2(x, x0, x1);
    return pack2(__fsqrt_rn(x0), __fsqrt_rn(x1));
}

// Joints are evaluated in the order the ATen summation consumes them (16..20 first, then k and k+8 together), so only
// two pair results are live at a time.
// SUM (16-bit tile image): return the 21-term sum s instead of D = s / 21 (the image scales s directly, and max_ij D_ij =
// (max_ij s_ij) / 21: one division per tile instead of one per pair), with the approximate square roots of joint_pair.
template <int MODE, bool SUM>
__device__ __forceinline__ float mpjpe_one(const f2 (&ax)[10], const f2 (&ay)[10], float ax20, float ay20,
                                           const float *__restrict__ col, const DivConst &div21)
{
    constexpr bool FAST = MODE != 0;
    float a, b;
    if (SUM) {
        // the image is rounded to 16 bits: the ATen summation order buys nothing here, so the 21 terms are added as packed
        // pairs (9 FADD2 + 2 FADD instead of 4 FADD2 + 14 FADD: the kernel is as much issue- as pipe-bound)
        f2 acc = joint_pair<MODE, true>(ax[0], ay[0], col, 0);
#pragma unroll
        for (int p = 1; p < 10; ++p) {
            if ((kFmaSqrtMask >> p) & 1u)
                acc = sqrt2_fma_pipe_acc(joint_x(ax[p], ay[p], col, p), acc);      // the sum rides on the chain's last FMA
            else
                acc = add2(acc, joint_pair<MODE, true>(ax[p], ay[p], col, p));
        }
        const float2 b20 = *reinterpret_cast<const float2 *>(col + 40);
        const float dx20 = __fsub_rn(ax20, b20.x), dy20 = __fsub_rn(ay20, b20.y);
        const float x20 = __fmaf_rn(dy20, dy20, __fmul_rn(dx20, dx20));
        unpack2(acc, a, b);
        return __fadd_rn(__fadd_rn(a, b), ((kFmaSqrtMask >> 10) & 1u) ? sqrt_fma_pipe(x20) : sqrt_approx(x20));
    }
    unpack2(joint_pair<MODE, SUM>(ax[8], ay[8], col, 8), a, b);          // (n16, n17)
    float s = __fadd_rn(a, b);
    unpack2(joint_pair<MODE, SUM>(ax[9], ay[9], col, 9), a, b);          // (n18, n19)
    s = __fadd_rn(s, a);
    s = __fadd_rn(s, b);
    {
        const float2 b20 = *reinterpret_cast<const float2 *>(col + 40);
        const float dx20 = __fsub_rn(ax20, b20.x), dy20 = __fsub_rn(ay20, b20.y);
        const float x20 = __fmaf_rn(dy20, dy20, __fmul_rn(dx20, dx20));
        s = __fadd_rn(s, SUM ? sqrt_approx(x20)
                             : (MODE == 2 ? sqrt_rn_fast_nz(x20) : (MODE == 1 ? sqrt_rn_fast(x20) : __fsqrt_rn(x20))));
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        f2 t = add2(joint_pair<MODE, SUM>(ax[p], ay[p], col, p), joint_pair<MODE, SUM>(ax[p + 4], ay[p + 4], col, p + 4));
        unpack2(t, a, b);                 // (n_2p + n_2p+8, n_2p+1 + n_2p+9)
        s = __fadd_rn(s, a);
        s = __fadd_rn(s, b);
    }
    if (SUM) return s;
    return FAST ? div_fast(s, div21) : __fdiv_rn(s, 21.0f);
}

// MODE as in joint_pair.  vmax_bits: running maximum of the integer image of D (D >= 0; NaN is larger than any finite).
// NCOLS: columns per thread (64: one CTA per stored tile; 32: one CTA per 64-column half).  col0: first column (inside
// the tile) of this thread's run; cs holds the staged columns starting at tile column cs0.
// Q16: the tile is stored as 16-bit fixed point q = round(D * qscale) = round(s * qscale / 21) (SMH_DIMS_Q16_TILES) and
// vmax_bits tracks the exact fp32 sum s (the caller divides the tile maximum by 21).  The rounding to integer rides on
// the FMA pipe: fma(s, qscale / 21, 2^23) has q in its low mantissa bits.
// RAGGED: the tile reaches past row / column m (padding rows are zeros: their distances must not enter the maximum); a full
// tile skips the per-column bounds test (a compare and a select per pair of a kernel bound by its dispatch port).
template <int MODE, int UN, int NCOLS, bool Q16, bool RAGGED = true>
__device__ __forceinline__ void mpjpe_tile_body(const float *__restrict__ jp, void *__restrict__ tile_out, int I,
                                                int J, int m, const float *cs, int cs0, int col0,
                                                uint32_t &vmax_bits, float qscale)
{
    const int r = threadIdx.x & 127;
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
    const DivConst div21 = make_div(21.0f);
    const bool row_ok = (I * kTile + r) < m;
    const int col_limit = m - J * kTile;          // columns >= col_limit are padding
    uint2 qprev = make_uint2(0u, 0u);             // 16-bit image: two 4-column steps share one 16-byte store
#pragma unroll 1
    for (int cq = 0; cq < NCOLS / UN; ++cq) {
        const int c0 = col0 + cq * UN;
        float dv[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            dv[u] = mpjpe_one<MODE, Q16>(ax, ay, ax20, ay20, cs + (c0 - cs0 + u) * kJP, div21);
            if (!RAGGED || (row_ok && (c0 + u) < col_limit)) vmax_bits = max(vmax_bits, __float_as_uint(dv[u]));
        }
#pragma unroll
        for (int u = 0; u < UN; u += 4) {
            if (Q16) {
                uint32_t q[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) q[t] = __float_as_uint(__fmaf_rn(dv[u + t], qscale, 8388608.0f));
                // low 16 bits of each: (q0 | q1 << 16, q2 | q3 << 16)
                const uint2 pk = make_uint2(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410));
                static_assert(!Q16 || UN == 4, "the 16-bit store pairs consecutive 4-column steps");
                if (cq & 1)
                    *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(tile_out) + distq_index(r, c0 - 4)) =
                        make_uint4(qprev.x, qprev.y, pk.x, pk.y);
                qprev = pk;
            } else {
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(tile_out) + dist_index(r, c0 + u)) =
                    make_float4(dv[u], dv[u + 1], dv[u + 2], dv[u + 3]);
            }
        }
    }
}

// 16-bit image, two rows per thread (r and r + 64) x NCOLS columns: every broadcast LDS of a column's joints serves two pairs,
// which halves the shared-memory instructions queued on the MIO path next to the MUFU.SQRT (ncu: mio_throttle is the top
// stall of the one-row form).  84 registers of row data: 2 CTAs per SM.
template <int NCOLS>
__device__ __forceinline__ void mpjpe_tile_body_2r(const float *__restrict__ jp, void *__restrict__ tile_out, int I, int J,
                                                   int m, const float *cs, int cs0, int col0, uint32_t &vmax_bits,
                                                   float qscale)
{
    const int r = threadIdx.x & 63;
    f2 ax[2][10], ay[2][10];
    float ax20[2], ay20[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 *rowp = reinterpret_cast<const float4 *>(jp + ((int64_t)I * kTile + r + 64 * h) * kJP);
#pragma unroll
        for (int p = 0; p < 10; ++p) {
            float4 v = rowp[p];
            ax[h][p] = pack2(v.x, v.y);
            ay[h][p] = pack2(v.z, v.w);
        }
        float4 v = rowp[10];
        ax20[h] = v.x;
        ay20[h] = v.y;
    }
    const DivConst div21 = make_div(21.0f);
    const bool row_ok[2] = {(I * kTile + r) < m, (I * kTile + r + 64) < m};
    const int col_limit = m - J * kTile;
    uint2 qprev[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
#pragma unroll 1
    for (int cq = 0; cq < NCOLS / 4; ++cq) {
        const int c0 = col0 + cq * 4;
        float dv[2][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float *col = cs + (c0 - cs0 + u) * kJP;
#pragma unroll
            for (int h = 0; h < 2; ++h) {          // same column pointer: the loads are shared between the two rows
                dv[h][u] = mpjpe_one<2, true>(ax[h], ay[h], ax20[h], ay20[h], col, div21);
                if (row_ok[h] && (c0 + u) < col_limit) vmax_bits = max(vmax_bits, __float_as_uint(dv[h][u]));
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t q[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) q[t] = __float_as_uint(__fmaf_rn(dv[h][t], qscale, 8388608.0f));
            const uint2 pk = make_uint2(__byte_perm(q[0], q[1], 0x5410), __byte_perm(q[2], q[3], 0x5410));
            if (cq & 1)
                *reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(tile_out) + distq_index(r + 64 * h, c0 - 4)) =
                    make_uint4(qprev[h].x, qprev[h].y, pk.x, pk.y);
            qprev[h] = pk;
        }
    }
}

// One work item = one 128 / PARTS-column part of a stored tile: thread = row x (64 / PARTS) columns.  Returns the block
// maximum of the integer image of the item's values (of D, or of the 21-term sum for the 16-bit image); NaN sorts on top.
template <int PARTS, bool Q16>
__device__ __forceinline__ uint32_t mpjpe_item(const int2 *__restrict__ tiles, const float *__restrict__ jp,
                                               void *__restrict__ dist, int m, int tile_id, int part, bool slow, float qscale,
                                               float *cs, uint32_t *wmax)
{
    constexpr int kCols = kTile / PARTS;              // columns staged per item
    constexpr int kPerThread = kCols / 2;
    const int cs0 = part * kCols;
    const int col0 = cs0 + (threadIdx.x >> 7) * kPerThread;
    const int2 ij = tiles[tile_id];
    void *tile_out = reinterpret_cast<unsigned char *>(dist) + (int64_t)tile_id * kTileFloats * (Q16 ? 2 : 4);
    {
        const float4 *src = reinterpret_cast<const float4 *>(jp + ((int64_t)ij.y * kTile + cs0) * kJP);
        float4 *dst = reinterpret_cast<float4 *>(cs);
        for (int i = threadIdx.x; i < kCols * kJP / 4; i += 256) dst[i] = src[i];
    }
    __syncthreads();
    auto block_max = [&](uint32_t v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = v;
        __syncthreads();
        v = wmax[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = max(v, wmax[w]);
        return v;
    };
    uint32_t vmax_bits = 0u;
    if (Q16 && SMH_MPJPE_Q16_ROWS == 2) {
        constexpr int kPer2 = kCols / 4;              // thread = rows (r, r + 64) x a quarter of the item's columns
        mpjpe_tile_body_2r<kPer2>(jp, tile_out, ij.x, ij.y, m, cs, cs0, cs0 + (threadIdx.x >> 6) * kPer2, vmax_bits, qscale);
        return block_max(vmax_bits);
    }
    if (Q16) {
        // one body: the approximate square roots take any input (sqrt(+0) = +0, non-finite values end up as NaN / inf in
        // the maximum and flag the step), so neither the IEEE path nor the guarded redo exists for the 16-bit image
        if ((ij.y + 1) * kTile > m || (ij.x + 1) * kTile > m)
            mpjpe_tile_body<2, 4, kPerThread, Q16, true>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);
        else
            mpjpe_tile_body<2, 4, kPerThread, Q16, false>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);
        return block_max(vmax_bits);
    }
    if (slow)
        mpjpe_tile_body<0, 4, kPerThread, Q16>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);
    else if (ij.x == ij.y || (ij.y + 1) * kTile > m || (ij.x + 1) * kTile > m)
        mpjpe_tile_body<1, 4, kPerThread, Q16>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);   // zero distances
    else            // off the diagonal and inside m in both directions (measured: dropping the bounds test here, as the 16-bit
                    // form does, is 1 % slower -- 928 vs 917 us -- the exact form schedules better with it)
        mpjpe_tile_body<2, 4, kPerThread, Q16, true>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);
    uint32_t bmax = block_max(vmax_bits);
    if (!slow && bmax > 0x7f800000u) {
        // a coincident joint in an off-diagonal tile: the unguarded form produced NaN somewhere; redo it guarded
        vmax_bits = 0u;
        mpjpe_tile_body<1, 4, kPerThread, Q16>(jp, tile_out, ij.x, ij.y, m, cs, cs0, col0, vmax_bits, qscale);
        bmax = block_max(vmax_bits);
    }
    return bmax;
}

// One CTA per work item.  The first n_full tiles (whole waves of the resident CTAs: 3 per SM for the 16-bit image, 2 for the
// exact form) are full-tile items; the tiles of an under-filled last wave are cut in halves, two CTAs each, so that the SMs
// are not left with a single CTA for the length of a full tile (a rank's 1032 tiles at 8 GPUs: 888 full + 144 x 2 halves).
// Measured alternatives, all slower: half / quarter items throughout, and a persistent grid pulling items from an atomic
// counter (profiles/r02_mpjpe_modes.txt) -- the per-item operand loads are not overlapped with the arithmetic.
template <bool Q16, bool FUSED>
__global__ void __launch_bounds__(256, Q16 ? SMH_MPJPE_Q16_CTAS : SMH_MPJPE_EXACT_CTAS)
mpjpe_kernel(const int2 *__restrict__ tiles, const float *__restrict__ jp, void *__restrict__ dist, int m, int n_full,
             Stats *__restrict__ stats, const __grid_constant__ Peers peers, int signal2)
{
    __shared__ __align__(16) float cs[kTile * kJP];
    __shared__ uint32_t wmax[8];
    const DistCtx dc = dist_head<FUSED>(stats, peers);
    const float qscale = Q16 ? q16_scale(__uint_as_float(dc.gs->dbound_bits)) * (1.0f / 21.0f) : 0.f;    // applied to the sum
    const uint32_t flags = dc.gs->flags;
    const bool slow = flags & (SMH_FLAG_SLOW_DOMAIN | SMH_FLAG_NONFINITE);
    uint32_t bmax;
    if ((int)blockIdx.x < n_full) {
        bmax = mpjpe_item<1, Q16>(tiles, jp, dist, m, (int)blockIdx.x, 0, slow, qscale, cs, wmax);
    } else {
        const int j = (int)blockIdx.x - n_full;
        bmax = mpjpe_item<2, Q16>(tiles, jp, dist, m, n_full + (j >> 1), j & 1, slow, qscale, cs, wmax);
    }
    if (threadIdx.x == 0) {
        if (bmax > 0x7f800000u)
            atomicOr(&stats->flags, SMH_FLAG_NONFINITE);      // non-finite inputs (IEEE path): the loss is NaN
        else
            atomicMax(&stats->dmax_bits, Q16 ? __float_as_uint(__fdiv_rn(__uint_as_float(bmax), 21.0f)) : bmax);
        // fused all-reduce(MAX): the last CTA of this rank pushes the rank's maximum into every peer's stats
        dist_tail(stats, peers, dc, signal2 != 0);
    }
}

// ----------------------------------------------------------------------------------------------
// diff_type w_abs / w_o_abs (utils.py:241-249): D_ij = || ((dx_k + dy_k) / 2)_k ||_2 over the 21 joints, with |.| on
// dx, dy for w_abs.  One sqrt per pair: a light kernel (same tiles, same layout, same max reduction).  Symmetric with
// a zero diagonal like the MPJPE, so the upper triangle is enough.  IEEE sqrt: no input-domain restriction.
// ----------------------------------------------------------------------------------------------
// MODE 0: w_o_abs, 1: w_abs, 2: euclid (SMH_DIFF_EUCLID, the *_with_pca weightings: D_ij = || a - b ||_2 over the 42
// packed coordinates, utils.py:282-293)
template <int MODE>
__global__ void __launch_bounds__(256, 2)
altdist_kernel(const int2 *__restrict__ tiles, const float *__restrict__ jp, float *__restrict__ dist, int m,
               Stats *__restrict__ stats, const __grid_constant__ Peers peers, int signal2)
{
    __shared__ __align__(16) float cs[kTile * kJP];
    __shared__ uint32_t wmax[8];
    const DistCtx dc = peers.fused ? dist_head<true>(stats, peers) : dist_head<false>(stats, peers);
    const int2 ij = tiles[blockIdx.x];
    float *tile_out = dist + (int64_t)blockIdx.x * kTileFloats;
    {
        const float4 *src = reinterpret_cast<const float4 *>(jp + (int64_t)ij.y * kTile * kJP);
        float4 *dst = reinterpret_cast<float4 *>(cs);
        for (int i = threadIdx.x; i < kTile * kJP / 4; i += 256) dst[i] = src[i];
    }
    __syncthreads();
    const int r = threadIdx.x & 127, col0 = (threadIdx.x >> 7) * 64;
    float a[kJP];
    {
        const float4 *rowp = reinterpret_cast<const float4 *>(jp + ((int64_t)ij.x * kTile + r) * kJP);
#pragma unroll
        for (int p = 0; p < 11; ++p) {
            const float4 v = rowp[p];
            a[4 * p] = v.x; a[4 * p + 1] = v.y; a[4 * p + 2] = v.z; a[4 * p + 3] = v.w;
        }
    }
    const bool row_ok = (ij.x * kTile + r) < m;
    const int col_limit = m - ij.y * kTile;
    uint32_t vmax_bits = 0u;
#pragma unroll 1
    for (int c0 = col0; c0 < col0 + 64; c0 += 4) {
        float dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float *b = cs + (c0 + u) * kJP;
            float acc = 0.f;
#pragma unroll
            for (int p = 0; p < 10; ++p) {          // packed (x_2p, x_2p+1, y_2p, y_2p+1)
                float dx0 = a[4 * p] - b[4 * p], dx1 = a[4 * p + 1] - b[4 * p + 1];
                float dy0 = a[4 * p + 2] - b[4 * p + 2], dy1 = a[4 * p + 3] - b[4 * p + 3];
                if (MODE == 2) {
                    acc = fmaf(dx0, dx0, acc);
                    acc = fmaf(dy0, dy0, acc);
                    acc = fmaf(dx1, dx1, acc);
                    acc = fmaf(dy1, dy1, acc);
                    continue;
                }
                if (MODE == 1) { dx0 = fabsf(dx0); dx1 = fabsf(dx1); dy0 = fabsf(dy0); dy1 = fabsf(dy1); }
                const float t0 = 0.5f * (dx0 + dy0), t1 = 0.5f * (dx1 + dy1);
                acc = fmaf(t0, t0, acc);
                acc = fmaf(t1, t1, acc);
            }
            float dx = a[40] - b[40], dy = a[41] - b[41];
            if (MODE == 2) {
                acc = fmaf(dx, dx, acc);
                acc = fmaf(dy, dy, acc);
            } else {
                if (MODE == 1) { dx = fabsf(dx); dy = fabsf(dy); }
                const float t = 0.5f * (dx + dy);
                acc = fmaf(t, t, acc);
            }
            dv[u] = __fsqrt_rn(acc);
            if (row_ok && (c0 + u) < col_limit) vmax_bits = max(vmax_bits, __float_as_uint(dv[u]));
        }
        *reinterpret_cast<float4 *>(tile_out + dist_index(r, c0)) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax_bits = max(vmax_bits, __shfl_xor_sync(0xffffffffu, vmax_bits, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = vmax_bits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t v = wmax[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) v = max(v, wmax[w]);
        if (v > 0x7f800000u)
            atomicOr(&stats->flags, SMH_FLAG_NONFINITE);
        else
            atomicMax(&stats->dmax_bits, v);
        dist_tail(stats, peers, dc, signal2 != 0);
    }
}

// non_linear weights (utils.py:343-346) need mean_ij D_ij over all M^2 ordered pairs: one pass over the stored tiles
// (an off-diagonal tile stands for both (I, J) and (J, I)), summed in double.
// Fused exchange: the rank's last CTA stores the rank's partial sum into slot `rank` of every rank's part array (the
// consumers add the parts in rank order) and signals stage 2.
__global__ void __launch_bounds__(256)
tile_sum_kernel(const int2 *__restrict__ tiles, const float *__restrict__ dist, int m, Stats *__restrict__ stats,
                const __grid_constant__ Peers peers)
{
    __shared__ double part[8];
    const int2 ij = tiles[blockIdx.x];
    const float *tile = dist + (int64_t)blockIdx.x * kTileFloats;
    float acc = 0.f;
    for (int idx = threadIdx.x; idx < kTileFloats / 4; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31;
        const float4 v = *reinterpret_cast<const float4 *>(tile + dist_index(r, 4 * c4));
        const int gi = ij.x * kTile + r, gj = ij.y * kTile + 4 * c4;
        if (gi < m) {
            if (gj < m) acc += v.x;
            if (gj + 1 < m) acc += v.y;
            if (gj + 2 < m) acc += v.z;
            if (gj + 3 < m) acc += v.w;
        }
    }
    double d = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w];
        atomicAdd(&stats->dsum, ij.x == ij.y ? t : 2.0 * t);
        if (peers.fused) {
            __threadfence();
            const unsigned ticket = atomicAdd(&stats->ticket2, 1u);
            if (ticket == gridDim.x - 1) {
                const double mine = atomicAdd(&stats->dsum, 0.0);
                for (int p = 0; p < peers.world; ++p)
                    reinterpret_cast<double *>(peers.lossparts(p) + 16)[peers.rank] = mine;
                stats->ticket2 = 0u;
                stage_signal(peers, 2, peers.my_sig()[kSigEpoch]);
            }
        }
    }
}

// a rank with nothing to do in a stage: wait for the previous stage, signal this one
__global__ void stage_relay_kernel(const __grid_constant__ Peers pe, int wait_stage, int stage)
{
    const uint32_t epoch = pe.my_sig()[kSigEpoch];
    if (wait_stage > 0) stage_wait(pe, wait_stage, epoch);
    if (threadIdx.x == 0) stage_signal(pe, stage, epoch);
}

int launch_mpjpe(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                 const Peers &peers, cudaStream_t stream)
{
    Stats *st = (Stats *)ws.stats;
    if (lay.n_stored_tiles == 0) {
        // a rank without tiles (fewer row-block pairs than ranks) still takes part in the exchange protocol
        if (peers.fused) {
            stage_relay_kernel<<<1, 32, 0, stream>>>(peers, 1, 2);
            return check_launch("stage_relay_kernel");
        }
        return 0;
    }
    const bool nonlinear = dims.weight_type == SMH_WEIGHT_NONLINEAR;
    const int signal2 = nonlinear ? 0 : 1;           // non_linear: the distance-sum pass closes stage 2
    if (peers.world > 1 && !peers.fused && (dims.diff_type != SMH_DIFF_MPJPE || nonlinear))
        return set_error(SMH_E_DIM, "w_abs / w_o_abs / non_linear on several ranks need the fused exchange");
    if (dims.diff_type != SMH_DIFF_MPJPE) {
        if (dims.diff_type == SMH_DIFF_W_ABS)
            altdist_kernel<1><<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.jp, ws.dist, lay.m, st, peers, signal2);
        else if (dims.diff_type == SMH_DIFF_EUCLID)
            altdist_kernel<2><<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.jp, ws.dist, lay.m, st, peers, signal2);
        else
            altdist_kernel<0><<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.jp, ws.dist, lay.m, st, peers, signal2);
    } else {
        // full-tile items for whole waves of the resident CTAs; an under-filled last wave (less than half) in half tiles
        const bool q16 = dims.flags & SMH_DIMS_Q16_TILES;
        const int resident = kNumCtas * (q16 ? SMH_MPJPE_Q16_CTAS : SMH_MPJPE_EXACT_CTAS);
        const int rem = lay.n_stored_tiles % resident;
        int n_full = (2 * rem <= resident) ? lay.n_stored_tiles - rem : lay.n_stored_tiles;
        if (const char *mode = getenv("SMH_MPJPE_HALVES")) n_full = atoi(mode) ? n_full : lay.n_stored_tiles;   // experiments
        const int grid = n_full + 2 * (lay.n_stored_tiles - n_full);
#define SMH_MPJPE(Q, F) mpjpe_kernel<Q, F><<<grid, 256, 0, stream>>>(plan.tiles, ws.jp, ws.dist, lay.m, n_full, st, peers, signal2)
        if (q16) { if (peers.fused) SMH_MPJPE(true, true); else SMH_MPJPE(true, false); }
        else     { if (peers.fused) SMH_MPJPE(false, true); else SMH_MPJPE(false, false); }
#undef SMH_MPJPE
    }
    int rc = check_launch(dims.diff_type != SMH_DIFF_MPJPE ? "altdist_kernel" : "mpjpe_kernel");
    if (rc || !nonlinear) return rc;
    tile_sum_kernel<<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.dist, lay.m, st, peers);
    return check_launch("tile_sum_kernel");
}

}  // namespace smh
