// simhand_b200: the small kernels around the sweeps.
//   rn_kernel            1 / neg_i                                        (feeds the backward sweep)
//   finalize_kernel      positive-pair weights (src/models/utils.py:233-235), positive logits (:420-423),
//                        loss = mean_i [log neg_i - S_ip Wp / tau] (:425-426), and the final gradient
//                        dz_i = dzacc_i / (M tau) - 2 Wp z_p(i) / (M tau).  The positive
//                        term is evaluated in fp32 from the caller's unrounded z in every mode.
//   weights_dense_kernel materialised pos_w [N] / neg_w [M, M] with the reference's shapes (:235, :259)
//   l2norm_*             K3: F.normalize forward/backward (simhand_w_model.py:56-58, 91-93)
#include <math_constants.h>

#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

// n_parts > 1 (peer exchange): neg_i = sum over ranks, in rank order, of the partial row sums delivered into negparts
__global__ void rn_kernel(float *__restrict__ neg, const float *__restrict__ negparts, int n_parts,
                          float *__restrict__ rn, int m, int mp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mp) return;
    float v = neg[i];
    if (n_parts > 1) {
        v = 0.f;
        for (int p = 0; p < n_parts; ++p) v += negparts[(int64_t)p * mp + i];
        neg[i] = v;
    }
    rn[i] = (i < m) ? __frcp_rn(v) : 0.f;
}

int launch_rn(const smh_layout_t &lay, const WsView &ws, int n_parts, cudaStream_t stream)
{
    const int mp = lay.tiles_per_side * kTile;
    rn_kernel<<<(mp + 255) / 256, 256, 0, stream>>>(ws.neg, ws.negparts, n_parts, ws.rn, lay.m, mp);
    return check_launch("rn_kernel");
}

__device__ __forceinline__ const float *sample_ptr_f(const float *base, int k, int n_local, int64_t rank_stride,
                                                     int64_t row_stride)
{
    return base + (int64_t)(k / n_local) * rank_stride + (int64_t)(k % n_local) * row_stride;
}

// phase 0: every row (single rank, NCCL transport).  Sharded finalize of the peer exchange: phase 1 evaluates the loss
// terms of this rank's own rows and stores their sum into slot `rank` of every peer's partial-loss array; phase 2 (after
// the barrier) writes the gradients of the own rows and combines the partial losses in rank order.
__global__ void __launch_bounds__(256)
finalize_kernel(smh_inputs_t in, int n, int d, int rank, const float *__restrict__ neg,
                const float *__restrict__ posd, float *__restrict__ rowloss, Stats *__restrict__ stats,
                const float *__restrict__ dzacc_src, int64_t src_row_offset, int n_parts, int64_t part_stride,
                int pos_mode, float lambda_pos, float inv_tau, float grad_scale,
                float *__restrict__ loss_out, float *__restrict__ dz1, float *__restrict__ dz2,
                int64_t dz_row_stride, int phase, const __grid_constant__ Peers peers)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int m = 2 * n;
    const float pmax = __uint_as_float(stats->pmax_bits);
    const float pmin = __uint_as_float(0x7fffffffu - stats->pmin_inv);
    const float pden = __fsub_rn(pmax, pmin);
    const int n_local = in.n_local;
    const int k_lo = rank * n_local, k_hi = k_lo + n_local;
    const float gs = grad_scale * inv_tau / (float)m;
    // the wide path: full-width rows whose every operand row starts on a 16-byte boundary
    const bool vec4 = d == kD && (in.z_row_stride & 3) == 0 && (in.z_rank_stride & 3) == 0 && (dz_row_stride & 3) == 0 &&
                      (part_stride & 3) == 0 &&
                      ((reinterpret_cast<uintptr_t>(in.z1_dev) | reinterpret_cast<uintptr_t>(in.z2_dev) |
                        reinterpret_cast<uintptr_t>(dz1) | reinterpret_cast<uintptr_t>(dz2) |
                        reinterpret_cast<uintptr_t>(dzacc_src)) & 15) == 0;
    __shared__ float pos_mean_s;
    if (pos_mode == 3) {
        // non_linear positives (utils.py:323-325): mean_k D_{k,k+N}, summed in the same fixed order by every block
        __shared__ float red[256];
        float a = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) a += posd[i];
        red[threadIdx.x] = a;
        __syncthreads();
        for (int s2 = 128; s2 > 0; s2 >>= 1) {
            if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
            __syncthreads();
        }
        if (threadIdx.x == 0) pos_mean_s = red[0] / (float)n;
        __syncthreads();
    }
    const float pos_mean = pos_mode == 3 ? pos_mean_s : 0.f;
    // rows this launch walks: all M, or the 2 * n_local rows of this rank (idx -> view v, sample k)
    const int rows = phase == 0 ? m : 2 * n_local;
    auto row_of = [&](int idx) { return phase == 0 ? idx : (idx / n_local) * n + k_lo + idx % n_local; };

    if (phase == 2 && blockIdx.x == 0 && threadIdx.x == 0) {
        float total = 0.f;
        for (int p = 0; p < peers.world; ++p) total += __ldcg(peers.lossparts(peers.rank) + p);       // rank order
        float loss = total / (float)m;
        if ((stats->flags & SMH_FLAG_NONFINITE) || stats->fail_site != 0u) loss = CUDART_NAN_F;
        if (peers.sig[peers.rank] != nullptr && ld_acquire_sys(peers.sig[peers.rank] + kSigPoison) != 0u)
            loss = CUDART_NAN_F;          // a barrier of the group timed out (now or earlier): nothing of this step is trusted
        stats->loss = loss;
        if (loss_out) *loss_out = loss;
    }

    for (int idx = blockIdx.x * warps_per_block + (threadIdx.x >> 5); idx < rows; idx += gridDim.x * warps_per_block) {
        const int row = row_of(idx);
        const int v = row >= n ? 1 : 0;
        const int k = row - v * n;
        const float *zi = sample_ptr_f(v ? in.z2_dev : in.z1_dev, k, n_local, in.z_rank_stride, in.z_row_stride);
        const float *zp = sample_ptr_f(v ? in.z1_dev : in.z2_dev, k, n_local, in.z_rank_stride, in.z_row_stride);
        // utils.py:235; 1: unit weights; 2: posd holds the caller's materialised Wp (smh_import_weights)
        float wp = pos_mode == 1 ? 1.0f : (pos_mode == 2 ? posd[k] : __fdiv_rn(__fsub_rn(pmax, posd[k]), pden));
        if (pos_mode == 3) wp = __fdiv_rn(1.0f, 1.0f + expf(lambda_pos * (posd[k] - pos_mean)));
        const bool grad_row = phase != 1 && dz1 != nullptr && k >= k_lo && k < k_hi;
        const float *src = dzacc_src + (dz_out_row(row, n, n_local) - src_row_offset) * kD;
        float *dst = (v ? dz2 : dz1) + (int64_t)(k - k_lo) * dz_row_stride;
        const float two_wp = 2.f * wp;
        if (vec4) {
            // d == 128, 16-byte aligned rows: one float4 per lane and operand, every load of the row issued before its first use
            const float4 a = *reinterpret_cast<const float4 *>(zi + 4 * lane);
            const float4 b = *reinterpret_cast<const float4 *>(zp + 4 * lane);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grad_row) {
                g = *reinterpret_cast<const float4 *>(src + 4 * lane);
                for (int p = 1; p < n_parts; ++p) {                                              // rank order
                    const float4 t = *reinterpret_cast<const float4 *>(src + (int64_t)p * part_stride + 4 * lane);
                    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
                }
            }
            if (phase != 2) {
                // the same per-lane order as the scalar path below (columns lane, lane + 32, ... there; 4 lane .. 4 lane + 3
                // here): a different but equally fixed summation order
                float dot = fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
                dot = warp_sum(dot);
                if (lane == 0) rowloss[row] = logf(neg[row]) - dot * wp * inv_tau;      // utils.py:420-426
            }
            if (grad_row)
                *reinterpret_cast<float4 *>(dst + 4 * lane) =
                    make_float4(gs * (g.x - two_wp * b.x), gs * (g.y - two_wp * b.y), gs * (g.z - two_wp * b.z),
                                gs * (g.w - two_wp * b.w));
            continue;
        }
        if (phase != 2) {
            float dot = 0.f;
            for (int c = lane; c < d; c += 32) dot = fmaf(zi[c], zp[c], dot);
            dot = warp_sum(dot);
            if (lane == 0) rowloss[row] = logf(neg[row]) - dot * wp * inv_tau;      // utils.py:420-426
        }
        if (grad_row) {
            for (int c = lane; c < d; c += 32) {
                float acc = src[c];
                for (int p = 1; p < n_parts; ++p) acc += src[(int64_t)p * part_stride + c];     // rank order
                dst[c] = gs * (acc - two_wp * zp[c]);
            }
        }
    }
    if (phase == 2) return;

    // last block reduces the per-row terms in a fixed order (deterministic loss)
    __shared__ bool is_last;
    __shared__ float part[256];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned ticket = atomicAdd(&stats->counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float acc = 0.f;
    int i = threadIdx.x;
    if (phase == 0) {
        // all rows, contiguous: four independent loads in flight per thread (64 dependent-latency loads per thread before)
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (; i + 768 < rows; i += 1024) {
            const float v0 = __ldcg(rowloss + i), v1 = __ldcg(rowloss + i + 256), v2 = __ldcg(rowloss + i + 512),
                        v3 = __ldcg(rowloss + i + 768);
            acc += v0; a1 += v1; a2 += v2; a3 += v3;
        }
        acc = (acc + a1) + (a2 + a3);
    }
    for (; i < rows; i += 256) acc += __ldcg(rowloss + row_of(i));
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
        if (threadIdx.x < s2) part[threadIdx.x] += part[threadIdx.x + s2];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats->counter = 0u;
        if (phase == 1) {
            // all-gather of the partial sums: one 4-byte store per peer
            for (int p = 0; p < peers.world; ++p) peers.lossparts(p)[peers.rank] = part[0];
            __threadfence_system();
            return;
        }
        float loss = part[0] / (float)m;
        if ((stats->flags & SMH_FLAG_NONFINITE) || stats->fail_site != 0u) loss = CUDART_NAN_F;
        stats->loss = loss;
        if (loss_out) *loss_out = loss;
    }
}

int launch_finalize(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                    const float *dzacc_src, bool local_block, int n_parts, int pos_mode, float temperature,
                    float grad_scale,
                    float *loss, float *dz1, float *dz2, int64_t dz_row_stride, int phase, const Peers &peers,
                    cudaStream_t stream)
{
    const int n_local = dims.n / dims.world;
    const int rows = phase == 0 ? lay.m : 2 * n_local;
    // 4 blocks per SM at most: every block ends with a ticket atomic on one address (last-block reduction of the loss),
    // and ~1200 of them serialised cost more than the few extra rows per warp
    // (4, 8 or 16 blocks per SM: no difference once the rows are read with float4 loads)
    const int blocks = (rows + 7) / 8 < 4 * kNumCtas ? (rows + 7) / 8 : 4 * kNumCtas;
    // a rank-local block (reduce-scattered buffer or peer-exchange accumulator) starts at this rank's first row
    const int64_t src_off = local_block ? (int64_t)dims.rank * 2 * n_local : 0;
    smh_inputs_t inp = in;
    inp.n_local = in.n_local;
    finalize_kernel<<<blocks, 256, 0, stream>>>(inp, dims.n, dims.d, dims.rank, ws.neg, ws.posd, ws.rowloss,
                                                (Stats *)ws.stats, dzacc_src, src_off, n_parts,
                                                (int64_t)2 * n_local * kD, pos_mode, dims.lambda_pos, 1.0f / temperature, grad_scale, loss, dz1,
                                                dz2, dz_row_stride, phase, peers);
    return check_launch("finalize_kernel");
}

// ----------------------------------------------------------------------------------------------
// materialised weights in: the caller's row-major neg_w [M, M] -> the stored-tile layout of the sweeps
// (one CTA per 128 x 128 tile; reads coalesced along rows, 16-byte swizzled stores; padding = 0)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
import_weights_kernel(const int2 *__restrict__ tiles, const float *__restrict__ neg_w, int64_t row_stride,
                      float *__restrict__ dist, int m)
{
    const int2 ij = tiles[blockIdx.x];
    float *tile = dist + (int64_t)blockIdx.x * kTileFloats;
    for (int idx = threadIdx.x; idx < kTileFloats / 4; idx += 256) {
        const int r = idx >> 5, c4 = idx & 31;
        const int gi = ij.x * kTile + r, gj = ij.y * kTile + 4 * c4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gi < m) {
            const float *src = neg_w + (int64_t)gi * row_stride + gj;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (gj + u < m) v[u] = src[u];
        }
        *reinterpret_cast<float4 *>(tile + dist_index(r, 4 * c4)) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void import_pos_kernel(const float *__restrict__ pos_w, float *__restrict__ posd, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) posd[i] = pos_w[i];
}

int launch_import_weights(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                          const float *neg_w, int64_t neg_row_stride, const float *pos_w, cudaStream_t stream)
{
    if (neg_w && lay.n_stored_tiles > 0)
        import_weights_kernel<<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, neg_w, neg_row_stride, ws.dist, lay.m);
    if (pos_w) import_pos_kernel<<<(dims.n + 255) / 256, 256, 0, stream>>>(pos_w, ws.posd, dims.n);
    return check_launch("import_weights_kernel");
}

// ----------------------------------------------------------------------------------------------
// materialised weights out
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
weights_dense_kernel(const int2 *__restrict__ tiles, const float *__restrict__ dist, const Stats *__restrict__ stats,
                     float *__restrict__ neg_w, int m, bool nonlinear, float lambda_neg)
{
    const float mu = nonlinear ? (float)(stats->dsum / ((double)m * (double)m)) : 0.f;
    const int2 ij = tiles[blockIdx.x];
    const float *tile = dist + (int64_t)blockIdx.x * kTileFloats;
    const float dmax = __uint_as_float(stats->dmax_bits);
    const bool slow = stats->flags & (SMH_FLAG_SLOW_DOMAIN | SMH_FLAG_NONFINITE);
    const DivConst divw = make_div(dmax);
    for (int idx = threadIdx.x; idx < kTileFloats; idx += 256) {
        // direct orientation: consecutive threads walk a row of the output
        int r = idx >> 7, c = idx & 127;
        int gi = ij.x * kTile + r, gj = ij.y * kTile + c;
        if (gi < m && gj < m) {
            float dv = tile[dist_index(r, c)];
            float num = __fsub_rn(dmax, dv);
            float w = slow ? __fdiv_rn(num, dmax) : div_fast(num, divw);
            if (nonlinear) w = __fdiv_rn(1.0f, 1.0f + expf(lambda_neg * (dv - mu)));        // utils.py:346
            neg_w[(int64_t)gi * m + gj] = w;
        }
        if (ij.x != ij.y) {
            // mirrored orientation: element (c', r') of the stored tile lands at row J*128 + c', col I*128 + r'
            int cc = idx >> 7, rr = idx & 127;
            int gi2 = ij.y * kTile + cc, gj2 = ij.x * kTile + rr;
            if (gi2 < m && gj2 < m) {
                float dv = tile[dist_index(rr, cc)];
                float num = __fsub_rn(dmax, dv);
                float w = slow ? __fdiv_rn(num, dmax) : div_fast(num, divw);
                if (nonlinear) w = __fdiv_rn(1.0f, 1.0f + expf(lambda_neg * (dv - mu)));
                neg_w[(int64_t)gi2 * m + gj2] = w;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
pos_weights_kernel(const float *__restrict__ posd, const Stats *__restrict__ stats, float *__restrict__ pos_w, int n,
                   bool nonlinear, float lambda_pos)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (nonlinear) {
        __shared__ float red[256];
        float a = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) a += posd[i];
        red[threadIdx.x] = a;
        __syncthreads();
        for (int s2 = 128; s2 > 0; s2 >>= 1) {
            if (threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
            __syncthreads();
        }
        const float mean = red[0] / (float)n;
        if (k < n) pos_w[k] = __fdiv_rn(1.0f, 1.0f + expf(lambda_pos * (posd[k] - mean)));        // utils.py:325
        return;
    }
    if (k >= n) return;
    const float pmax = __uint_as_float(stats->pmax_bits);
    const float pmin = __uint_as_float(0x7fffffffu - stats->pmin_inv);
    pos_w[k] = __fdiv_rn(__fsub_rn(pmax, posd[k]), __fsub_rn(pmax, pmin));
}

int launch_weights_dense(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                         float *pos_w, float *neg_w, cudaStream_t stream)
{
    if (pos_w) {
        pos_weights_kernel<<<(dims.n + 255) / 256, 256, 0, stream>>>(ws.posd, (const Stats *)ws.stats, pos_w, dims.n,
                                                                      dims.weight_type == SMH_WEIGHT_NONLINEAR, dims.lambda_pos);
        int rc = check_launch("pos_weights_kernel");
        if (rc) return rc;
    }
    if (neg_w && lay.n_stored_tiles > 0) {
        weights_dense_kernel<<<lay.n_stored_tiles, 256, 0, stream>>>(plan.tiles, ws.dist, (const Stats *)ws.stats,
                                                                    neg_w, lay.m,
                                                                    dims.weight_type == SMH_WEIGHT_NONLINEAR,
                                                                    dims.lambda_neg);
        return check_launch("weights_dense_kernel");
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------
// K3: row-wise L2 normalisation
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, float *__restrict__ norm, int64_t rows, int d,
                  float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    for (int64_t row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        const float *xr = x + row * d;
        float *yr = y + row * d;
        float ss = 0.f;
        if (vec) {
            for (int c = lane * 4; c < d; c += 128) {
                float4 v = *reinterpret_cast<const float4 *>(xr + c);
                ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
            }
        } else {
            for (int c = lane; c < d; c += 32) ss = fmaf(xr[c], xr[c], ss);
        }
        ss = warp_sum(ss);
        const float nrm = sqrtf(ss);
        const float inv = 1.0f / fmaxf(nrm, eps);
        if (vec) {
            for (int c = lane * 4; c < d; c += 128) {
                float4 v = *reinterpret_cast<const float4 *>(xr + c);
                *reinterpret_cast<float4 *>(yr + c) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
            }
        } else {
            for (int c = lane; c < d; c += 32) yr[c] = xr[c] * inv;
        }
        if (norm && lane == 0) norm[row] = nrm;
    }
}

// dx = (dy - y (y . dy)) / max(||x||, eps)   (for ||x|| >= eps; below eps F.normalize is a plain scale)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float *__restrict__ y, const float *__restrict__ norm, const float *__restrict__ dy,
                  float *__restrict__ dx, int64_t rows, int d, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    for (int64_t row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        const float *yr = y + row * d, *gr = dy + row * d;
        float *xr = dx + row * d;
        float dot = 0.f;
        for (int c = lane; c < d; c += 32) dot = fmaf(yr[c], gr[c], dot);
        dot = warp_sum(dot);
        const float nrm = norm[row];
        const float inv = 1.0f / fmaxf(nrm, eps);
        if (nrm < eps) dot = 0.f;
        for (int c = lane; c < d; c += 32) xr[c] = (gr[c] - yr[c] * dot) * inv;
    }
}

// autograd backward of the fused loss: both saved gradients times the upstream scalar, one launch
__global__ void __launch_bounds__(256)
scale_pair_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b, const float *__restrict__ scale,
                  float4 *__restrict__ oa, float4 *__restrict__ ob, int64_t n4, int64_t tail_from,
                  const float *__restrict__ at, const float *__restrict__ bt, float *__restrict__ oat,
                  float *__restrict__ obt, int64_t count)
{
    const float s = __ldg(scale);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 va = a[i], vb = b[i];
        oa[i] = make_float4(s * va.x, s * va.y, s * va.z, s * va.w);
        ob[i] = make_float4(s * vb.x, s * vb.y, s * vb.z, s * vb.w);
    }
    if (blockIdx.x == 0)
        for (int64_t i = tail_from + threadIdx.x; i < count; i += blockDim.x) {
            oat[i] = s * at[i];
            obt[i] = s * bt[i];
        }
}

int launch_scale_pair(const float *a, const float *b, const float *scale, float *oa, float *ob, int64_t count,
                      cudaStream_t stream)
{
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(oa) |
                       reinterpret_cast<uintptr_t>(ob)) & 15) == 0;
    const int64_t n4 = vec ? count / 4 : 0;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    scale_pair_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b),
                                                       scale, reinterpret_cast<float4 *>(oa), reinterpret_cast<float4 *>(ob),
                                                       n4, n4 * 4, a, b, oa, ob, count);
    return check_launch("scale_pair_kernel");
}

int launch_l2norm_fwd(const float *x, float *y, float *norm, int64_t rows, int d, float eps, cudaStream_t stream)
{
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    l2norm_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, y, norm, rows, d, eps);
    return check_launch("l2norm_fwd_kernel");
}

int launch_l2norm_bwd(const float *y, const float *norm, const float *dy, float *dx, int64_t rows, int d, float eps,
                      cudaStream_t stream)
{
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    l2norm_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(y, norm, dy, dx, rows, d, eps);
    return check_launch("l2norm_bwd_kernel");
}

}  // namespace smh
