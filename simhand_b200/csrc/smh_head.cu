// simhand_b200 K5: the projection head fused with the first normalisation.
//
// Reference (src/models/unsupervised/simclr_model.py:22-39, called at simhand_w_model.py:45-58):
//     Linear(in, hidden, bias=True) -> BatchNorm1d(hidden) [training: batch statistics] -> ReLU -> Linear(hidden, out, bias=False)
//     -> F.normalize(dim=1)
// with in = 2048, hidden = 512, out = 128 for the ResNet-50 recipe, on [2B, in] encodings under 16-bit autocast.
//
// Forward, two tcgen05 kernels:
//   head_gemm1_kernel   H = X W1^T + b1 (bf16 or fp16 operands staged by TMA, fp32 accumulate in TMEM, 256 x 256 tile per CTA,
//                       3-stage mbarrier pipeline).  Epilogue: bias, rounding to the 16-bit activation type (the autocast
//                       semantics of the reference: BatchNorm sees the rounded Linear output), H stored, and the per-column
//                       sum / sum of squares of the rounded values reduced across the tile's rows with a shuffle butterfly
//                       -> one atomicAdd per column and CTA.  The BatchNorm statistics cost no extra pass over H.
//   head_gemm2_kernel   per 128-row tile: BatchNorm scale/shift from the column sums (every CTA derives them; CTA 0 also
//                       updates running_mean / running_var and saves mean / rstd), then the K loop stages H by TMA, applies
//                       BN + ReLU IN PLACE in shared memory (the activation never goes back to HBM), feeds it as the A operand
//                       of P = relu(bn(H)) W2^T, and the epilogue L2-normalises the 128-wide rows straight out of TMEM
//                       (thread = row) -> Y fp32 and the row norms.
// Backward: head_bwd1_kernel fuses normalise-backward, dA = dP W2 (tcgen05), the ReLU mask (recomputed from H), the store of
// A = relu(bn(H)) for dW2 and the two column reductions BatchNorm's backward needs; head_bwd2_kernel is the BatchNorm backward
// itself.  The three plain GEMMs that remain (dW2 = dP^T A, dW1 = dH^T X, dX = dH W1) are library GEMMs on the host side.
#include <cuda.h>
#include <math_constants.h>

#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

// ----------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched through the runtime: the library does not link libcuda)
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D row-major [rows][cols] of 16-bit elements, box = [box_rows][64 columns] (128 bytes: one SWIZZLE_128B row)
static int make_map_16(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int64_t row_stride_elems,
                       int box_rows, bool fp16)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return set_error(SMH_E_ARCH, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)row_stride_elems * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                    gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(SMH_E_ARG, "cuTensorMapEncodeTiled failed (%d): base must be 16-byte aligned, row stride a multiple of 16 bytes", (int)r);
    return 0;
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ float bf16_round(float v) { return __uint_as_float(pack_bf16x2(v, 0.f) << 16); }
__device__ __forceinline__ float f16_round(float v)
{
    uint16_t h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    float r;
    asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
    return r;
}
__device__ __forceinline__ float unpack16(uint32_t w, int hi, bool fp16)
{
    const uint16_t h = hi ? (uint16_t)(w >> 16) : (uint16_t)(w & 0xffffu);
    if (!fp16) return __uint_as_float((uint32_t)h << 16);
    float r;
    asm("cvt.f32.f16 %0, %1;" : "=f"(r) : "h"(h));
    return r;
}
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, bool fp16)
{
    return fp16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}

// Column sums over the 32 rows a warp holds (lane = row, v[i] = column i): after the butterfly lane l holds column l.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float keep = up ? v[i + s] : v[i];
            const float send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// ----------------------------------------------------------------------------------------------
// G1: H = X W1^T + b1, column sums of H and H^2
// ----------------------------------------------------------------------------------------------
constexpr int kG1Threads = 320;                  // TMA producer, MMA issuer, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int kG1Stages = 3;
constexpr int kG1StageBytes = 2 * 256 * 128;     // A box [256 rows][64] + B box [256 rows][64], 16-bit
constexpr int kG1Smem = 1024 + kG1Stages * kG1StageBytes + 256;

__global__ void __launch_bounds__(kG1Threads, 1)
head_gemm1_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                  const float *__restrict__ bias, uint16_t *__restrict__ h_out, float *__restrict__ colsum, int rows, int in_dim,
                  int hidden, int fp16)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + kG1Stages * kG1StageBytes);        // full[3], empty[3], acc_full
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    __shared__ uint32_t fail_s;
    __shared__ float part[4][2][256];             // per-warp column partials (sum, sum of squares)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 256, col0 = blockIdx.y * 256;
    const int kblocks = in_dim / 64;
    if (threadIdx.x == 0) {
        fail_s = 0u;
        for (int i = 0; i < kG1Stages; ++i) {
            mbar_init(&bars[i], 1);
            mbar_init(&bars[3 + i], 1);
        }
        mbar_init(&bars[6], 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        tc_alloc(tmem_slot, 512);
        tc_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % kG1Stages;
            mbar_wait(&bars[3 + s], ((kb / kG1Stages) & 1) ^ 1, &fail_s, 1);
            if (elect_one()) {
                unsigned char *dst = sm + s * kG1StageBytes;
                mbar_arrive_expect_tx(&bars[s], kG1StageBytes);
                tma_load_2d(dst, &map_x, kb * 64, row0, &bars[s]);
                tma_load_2d(dst + 256 * 128, &map_w1, kb * 64, col0, &bars[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = fp16 ? umma_idesc_f16(128, 256, 0, 0) : umma_idesc_bf16(128, 256, 0, 0);
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % kG1Stages;
            mbar_wait(&bars[s], (kb / kG1Stages) & 1, &fail_s, 2);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_u = smem_u32(sm + s * kG1StageBytes), b_u = a_u + 256 * 128;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t bdesc = umma_desc_sw128(b_u + ks * 32, 16, 1024);
#pragma unroll
                    for (int mh = 0; mh < 2; ++mh) {
                        const uint64_t adesc = umma_desc_sw128(a_u + mh * 16384 + ks * 32, 16, 1024);
                        tc_mma_ss_f16(tmem_base + mh * 256, adesc, bdesc, idesc, (kb | ks) ? 1u : 0u);
                    }
                }
                tc_commit(&bars[3 + s]);
                if (kb + 1 == kblocks) tc_commit(&bars[6]);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue: 8 warps, thread = row of a 128-row
        // half; the two warps of a TMEM lane quadrant split the 256 columns of the tile
        const int w4 = warp & 3;                       // TMEM lane quadrant this warp may touch
        const int qi = (warp - 2) & 3;                 // slot of this warp's partial sums
        const int chalf = (warp - 2) >> 2;             // which 128 of the 256 columns
        mbar_wait(&bars[6], 0, &fail_s, 3);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
        for (int c = lane; c < 128; c += 32) {
            part[qi][0][chalf * 128 + c] = 0.f;
            part[qi][1][chalf * 128 + c] = 0.f;
        }
        __syncwarp();
        for (int mh = 0; mh < 2; ++mh) {
            const int grow = row0 + mh * 128 + w4 * 32 + lane;
            const bool row_ok = grow < rows;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                const int ch = chalf * 4 + c4;
                uint32_t v[32];
                tc_ld32(lane_addr + mh * 256 + ch * 32, v);
                tc_wait_ld();
                float hv[32];
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float b0 = __ldg(bias + col0 + ch * 32 + i), b1 = __ldg(bias + col0 + ch * 32 + i + 1);
                    const float x0 = __uint_as_float(v[i]) + b0, x1 = __uint_as_float(v[i + 1]) + b1;
                    pk[i >> 1] = pack16x2(x0, x1, fp16 != 0);
                    // statistics of the ROUNDED activation (what BatchNorm sees under autocast)
                    hv[i] = row_ok ? unpack16(pk[i >> 1], 0, fp16 != 0) : 0.f;
                    hv[i + 1] = row_ok ? unpack16(pk[i >> 1], 1, fp16 != 0) : 0.f;
                }
                if (row_ok) {
                    uint4 *dst = reinterpret_cast<uint4 *>(h_out + (int64_t)grow * hidden + col0 + ch * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                float sq[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) sq[i] = hv[i] * hv[i];
                const float s1 = warp_transpose_sum(hv, lane);
                const float s2 = warp_transpose_sum(sq, lane);
                part[qi][0][ch * 32 + lane] += s1;
                part[qi][1][ch * 32 + lane] += s2;
            }
        }
        tc_fence_before();
        // combine the four row quadrants, one atomic per column and statistic
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int t = threadIdx.x - 64;
        for (int c = t; c < 2 * 256; c += 256) {
            const int st = c >> 8, col = c & 255;
            const float tot = part[0][st][col] + part[1][st][col] + part[2][st][col] + part[3][st][col];
            atomicAdd(colsum + (int64_t)st * hidden + col0 + col, tot);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, 512);
    if (threadIdx.x == 0 && fail_s != 0u) colsum[0] = CUDART_NAN_F;          // a timed-out wait poisons the statistics
}

// ----------------------------------------------------------------------------------------------
// G2: Y = normalize(relu(bn(H)) W2^T)
// ----------------------------------------------------------------------------------------------
constexpr int kG2Threads = 192;
constexpr int kG2Stages = 2;
constexpr int kG2StageBytes = 2 * 128 * 256;     // H chunk [128 rows][128 cols] + W2 chunk [128 out][128 cols], 16-bit
constexpr int kG2Smem = 1024 + kG2Stages * kG2StageBytes + 256;

struct BnParams {
    const float *colsum;         // [2][hidden] from G1
    const float *gamma, *beta;   // [hidden]
    float *running_mean, *running_var;      // [hidden] or null
    float *save_mean, *save_rstd;           // [hidden]
    float eps, momentum;
    int training;
};

__global__ void __launch_bounds__(kG2Threads, 1)
head_gemm2_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w2, BnParams bn,
                  float *__restrict__ y_out, float *__restrict__ norm_out, int rows, int hidden, int out_dim, float norm_eps,
                  int fp16)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + kG2Stages * kG2StageBytes);   // full[2], empty[2], xformed[2], acc_full
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    __shared__ uint32_t fail_s;
    __shared__ float scale_s[1024], shift_s[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;
    const int chunks = hidden / 128;
    if (threadIdx.x == 0) {
        fail_s = 0u;
        for (int i = 0; i < kG2Stages; ++i) {
            mbar_init(&bars[i], 1);
            mbar_init(&bars[2 + i], 1);
            mbar_init(&bars[4 + i], 128);
        }
        mbar_init(&bars[6], 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        tc_alloc(tmem_slot, 128);
        tc_relinquish();
    }
    // BatchNorm1d, training mode: biased variance for the normalisation, unbiased for the running estimate
    for (int c = threadIdx.x; c < hidden; c += kG2Threads) {
        float mean = bn.colsum[c] / (float)rows;
        float var = fmaxf(bn.colsum[hidden + c] / (float)rows - mean * mean, 0.f);
        if (!bn.training) {                        // eval: the running estimates (nn.BatchNorm1d.eval())
            mean = bn.running_mean[c];
            var = bn.running_var[c];
        }
        const float rstd = rsqrtf(var + bn.eps);
        const float sc = bn.gamma[c] * rstd;
        scale_s[c] = sc;
        shift_s[c] = bn.beta[c] - mean * sc;
        if (blockIdx.x == 0) {
            bn.save_mean[c] = mean;
            bn.save_rstd[c] = rstd;
            if (bn.running_mean && bn.training) {
                bn.running_mean[c] = (1.f - bn.momentum) * bn.running_mean[c] + bn.momentum * mean;
                bn.running_var[c] = (1.f - bn.momentum) * bn.running_var[c] +
                                    bn.momentum * var * ((float)rows / (float)max(rows - 1, 1));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kG2Stages;
            mbar_wait(&bars[2 + s], ((kc / kG2Stages) & 1) ^ 1, &fail_s, 1);
            if (elect_one()) {
                unsigned char *dst = sm + s * kG2StageBytes;
                mbar_arrive_expect_tx(&bars[s], kG2StageBytes);
                tma_load_2d(dst, &map_h, kc * 128, row0, &bars[s]);
                tma_load_2d(dst + 16384, &map_h, kc * 128 + 64, row0, &bars[s]);
                tma_load_2d(dst + 32768, &map_w2, kc * 128, 0, &bars[s]);
                tma_load_2d(dst + 49152, &map_w2, kc * 128 + 64, 0, &bars[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = fp16 ? umma_idesc_f16(128, 128, 0, 0) : umma_idesc_bf16(128, 128, 0, 0);
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kG2Stages;
            mbar_wait(&bars[4 + s], (kc / kG2Stages) & 1, &fail_s, 2);          // BN + ReLU applied in place
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_u = smem_u32(sm + s * kG2StageBytes), b_u = a_u + 32768;
#pragma unroll
                for (int bx = 0; bx < 2; ++bx) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adesc = umma_desc_sw128(a_u + bx * 16384 + ks * 32, 16, 1024);
                        const uint64_t bdesc = umma_desc_sw128(b_u + bx * 16384 + ks * 32, 16, 1024);
                        tc_mma_ss_f16(tmem_base, adesc, bdesc, idesc, (kc | bx | ks) ? 1u : 0u);
                    }
                }
                tc_commit(&bars[2 + s]);
                if (kc + 1 == chunks) tc_commit(&bars[6]);
            }
            __syncwarp();
        }
    } else {
        const int w4 = warp & 3;
        const int r = w4 * 32 + lane;                 // row inside the tile == TMEM lane
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kG2Stages;
            mbar_wait(&bars[s], (kc / kG2Stages) & 1, &fail_s, 3);
            unsigned char *hs = sm + s * kG2StageBytes;
            // thread = row: 2 boxes x 8 chunks of 16 bytes; physical chunk = logical ^ (row & 7)
#pragma unroll
            for (int bx = 0; bx < 2; ++bx) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint4 *p = reinterpret_cast<uint4 *>(hs + bx * 16384 + r * 128 + ((j ^ (r & 7)) << 4));
                    uint4 q = *p;
                    uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    const int cbase = kc * 128 + bx * 64 + j * 8;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float a0 = fmaxf(fmaf(unpack16(w[u], 0, fp16 != 0), scale_s[cbase + 2 * u], shift_s[cbase + 2 * u]), 0.f);
                        const float a1 = fmaxf(fmaf(unpack16(w[u], 1, fp16 != 0), scale_s[cbase + 2 * u + 1], shift_s[cbase + 2 * u + 1]), 0.f);
                        w[u] = pack16x2(a0, a1, fp16 != 0);
                    }
                    *p = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(&bars[4 + s]);
        }
        // epilogue: thread = row, 128 fp32 columns straight out of TMEM
        mbar_wait(&bars[6], 0, &fail_s, 4);
        tc_fence_after();
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
        const int grow = row0 + r;
        float ss = 0.f;
        for (int ch = 0; ch < out_dim / 32; ++ch) {
            uint32_t v[32];
            tc_ld32(lane_addr + ch * 32, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ss = fmaf(__uint_as_float(v[i]), __uint_as_float(v[i]), ss);
        }
        const float nrm = sqrtf(ss);
        const float inv = 1.0f / fmaxf(nrm, norm_eps);
        for (int ch = 0; ch < out_dim / 32; ++ch) {
            uint32_t v[32];
            tc_ld32(lane_addr + ch * 32, v);
            tc_wait_ld();
            if (grow < rows) {
                float4 *dst = reinterpret_cast<float4 *>(y_out + (int64_t)grow * out_dim + ch * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    dst[q] = make_float4(__uint_as_float(v[4 * q]) * inv, __uint_as_float(v[4 * q + 1]) * inv,
                                         __uint_as_float(v[4 * q + 2]) * inv, __uint_as_float(v[4 * q + 3]) * inv);
            }
        }
        if (grow < rows) norm_out[grow] = (fail_s != 0u) ? CUDART_NAN_F : nrm;
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, 128);
}

// ----------------------------------------------------------------------------------------------
// backward 1: dP = normalize_bwd(dY), dA = dP W2 (tcgen05), dHn = dA * 1[bn(H) > 0]; stores dP, A = relu(bn(H)), dHn and the
// column sums of dHn and dHn * Hhat
// ----------------------------------------------------------------------------------------------
constexpr int kB1Threads = 192;
constexpr int kB1Stages = 2;
constexpr int kB1StageBytes = 2 * 128 * 256;     // W2T chunk [128 hidden][128 out] + H chunk [128 rows][128 hidden]
constexpr int kB1Smem = 1024 + 32768 + kB1Stages * kB1StageBytes + 256;

__global__ void __launch_bounds__(kB1Threads, 1)
head_bwd1_kernel(const __grid_constant__ CUtensorMap map_w2t, const __grid_constant__ CUtensorMap map_h,
                 const float *__restrict__ dy, const float *__restrict__ y, const float *__restrict__ norm,
                 const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ save_mean,
                 const float *__restrict__ save_rstd, uint16_t *__restrict__ dp_out, uint16_t *__restrict__ a_out,
                 uint16_t *__restrict__ dhn_out, float *__restrict__ colsum, int rows, int hidden, int out_dim, float norm_eps,
                 int fp16)
{
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sP = sm;                                                   // dP tile [128 rows][128 out], 2 boxes
    unsigned char *sS = sm + 32768;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sS + kB1Stages * kB1StageBytes);   // full[2], empty[2], p_ready, acc_full[2], acc_free[2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);
    __shared__ uint32_t fail_s;
    __shared__ float part[4][2][128];
    // per-column BatchNorm constants: hn = h * sc + sh (sc = gamma rstd, sh = beta - mean sc); hhat = h * rs + nm
    __shared__ float sc_s[1024], sh_s[1024], rs_s[1024], nm_s[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;
    const int chunks = hidden / 128;
    for (int c = threadIdx.x; c < hidden; c += kB1Threads) {
        const float mean = save_mean[c], rstd = save_rstd[c];
        const float sc = gamma[c] * rstd;
        sc_s[c] = sc;
        sh_s[c] = beta[c] - mean * sc;
        rs_s[c] = rstd;
        nm_s[c] = -mean * rstd;
    }
    if (threadIdx.x == 0) {
        fail_s = 0u;
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[i], 1);
            mbar_init(&bars[2 + i], 128);       // stage free: the epilogue has read the H chunk (the MMA's read of W2T precedes it)
            mbar_init(&bars[5 + i], 1);         // accumulator chunk complete
            mbar_init(&bars[7 + i], 128);       // accumulator buffer drained
        }
        mbar_init(&bars[4], 128);
        mbar_fence_init();
    }
    if (warp == 2) {
        tc_alloc(tmem_slot, 256);
        tc_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kB1Stages;
            mbar_wait(&bars[2 + s], ((kc / kB1Stages) & 1) ^ 1, &fail_s, 1);
            if (elect_one()) {
                unsigned char *dst = sS + s * kB1StageBytes;
                mbar_arrive_expect_tx(&bars[s], kB1StageBytes);
                tma_load_2d(dst, &map_w2t, 0, kc * 128, &bars[s]);                 // W2T rows = hidden, cols = out
                tma_load_2d(dst + 16384, &map_w2t, 64, kc * 128, &bars[s]);
                tma_load_2d(dst + 32768, &map_h, kc * 128, row0, &bars[s]);
                tma_load_2d(dst + 49152, &map_h, kc * 128 + 64, row0, &bars[s]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t idesc = fp16 ? umma_idesc_f16(128, 128, 0, 0) : umma_idesc_bf16(128, 128, 0, 0);
        mbar_wait(&bars[4], 0, &fail_s, 2);                                        // dP tile written
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kB1Stages, ab = kc & 1;
            mbar_wait(&bars[s], (kc / kB1Stages) & 1, &fail_s, 3);
            mbar_wait(&bars[7 + ab], ((kc >> 1) & 1) ^ 1, &fail_s, 4);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_u = smem_u32(sP), b_u = smem_u32(sS + s * kB1StageBytes);
#pragma unroll
                for (int bx = 0; bx < 2; ++bx) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t adesc = umma_desc_sw128(a_u + bx * 16384 + ks * 32, 16, 1024);
                        const uint64_t bdesc = umma_desc_sw128(b_u + bx * 16384 + ks * 32, 16, 1024);
                        tc_mma_ss_f16(tmem_base + ab * 128, adesc, bdesc, idesc, (bx | ks) ? 1u : 0u);
                    }
                }
                tc_commit(&bars[5 + ab]);
            }
            __syncwarp();
        }
    } else {
        const int w4 = warp & 3, e = warp - 2;
        const int r = w4 * 32 + lane;
        const int grow = row0 + r;
        const bool row_ok = grow < rows;
        // normalise backward, thread = row: dP = (dY - Y (Y . dY)) / max(||P||, eps)   (a plain scale below eps)
        {
            const float4 *dyr = reinterpret_cast<const float4 *>(dy + (int64_t)(row_ok ? grow : 0) * out_dim);
            const float4 *yr = reinterpret_cast<const float4 *>(y + (int64_t)(row_ok ? grow : 0) * out_dim);
            const float nrm = row_ok ? norm[grow] : 1.f;
            float dot = 0.f;
            for (int q = 0; q < out_dim / 4; ++q) {
                const float4 a = yr[q], b = dyr[q];
                dot = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, dot))));
            }
            if (nrm < norm_eps) dot = 0.f;
            const float inv = row_ok ? 1.0f / fmaxf(nrm, norm_eps) : 0.f;
            for (int j = 0; j < out_dim / 8; ++j) {
                const float4 a0 = yr[2 * j], a1 = yr[2 * j + 1], b0 = dyr[2 * j], b1 = dyr[2 * j + 1];
                uint4 q;
                q.x = pack16x2((b0.x - a0.x * dot) * inv, (b0.y - a0.y * dot) * inv, fp16 != 0);
                q.y = pack16x2((b0.z - a0.z * dot) * inv, (b0.w - a0.w * dot) * inv, fp16 != 0);
                q.z = pack16x2((b1.x - a1.x * dot) * inv, (b1.y - a1.y * dot) * inv, fp16 != 0);
                q.w = pack16x2((b1.z - a1.z * dot) * inv, (b1.w - a1.w * dot) * inv, fp16 != 0);
                const int bx = j >> 3, jj = j & 7;
                *reinterpret_cast<uint4 *>(sP + bx * 16384 + r * 128 + ((jj ^ (r & 7)) << 4)) = q;
                if (row_ok) *reinterpret_cast<uint4 *>(dp_out + (int64_t)grow * out_dim + j * 8) = q;
            }
            fence_proxy_async();
            mbar_arrive(&bars[4]);
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)(w4 * 32) << 16);
        for (int kc = 0; kc < chunks; ++kc) {
            const int s = kc % kB1Stages, ab = kc & 1;
            mbar_wait(&bars[5 + ab], (kc >> 1) & 1, &fail_s, 5);
            tc_fence_after();
            const unsigned char *hs = sS + s * kB1StageBytes + 32768;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {                                   // 32 hidden columns at a time
                uint32_t v[32];
                tc_ld32(lane_addr + ab * 128 + ch * 32, v);
                tc_wait_ld();
                float g[32], gh[32];
                uint32_t pa[16], pg[16];
                const int bx = ch >> 1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int jj = (ch & 1) * 4 + j;
                    const uint4 q = *reinterpret_cast<const uint4 *>(hs + bx * 16384 + r * 128 + ((jj ^ (r & 7)) << 4));
                    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int i = j * 8 + u * 2 + hf;
                            const int c = kc * 128 + ch * 32 + i;
                            const float hval = unpack16(w[u], hf, fp16 != 0);
                            const float hhat = fmaf(hval, rs_s[c], nm_s[c]);
                            const float hn = fmaf(hval, sc_s[c], sh_s[c]);
                            const float da = __uint_as_float(v[i]);
                            const float dh = (hn > 0.f && row_ok) ? da : 0.f;
                            g[i] = dh;
                            gh[i] = dh * hhat;
                            const float act = row_ok ? fmaxf(hn, 0.f) : 0.f;
                            if (hf == 0) {
                                pa[i >> 1] = __float_as_uint(act);
                                pg[i >> 1] = __float_as_uint(dh);
                            } else {
                                pa[i >> 1] = pack16x2(__uint_as_float(pa[i >> 1]), act, fp16 != 0);
                                pg[i >> 1] = pack16x2(__uint_as_float(pg[i >> 1]), dh, fp16 != 0);
                            }
                        }
                    }
                }
                if (row_ok) {
                    uint4 *da_dst = reinterpret_cast<uint4 *>(a_out + (int64_t)grow * hidden + kc * 128 + ch * 32);
                    uint4 *dg_dst = reinterpret_cast<uint4 *>(dhn_out + (int64_t)grow * hidden + kc * 128 + ch * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        da_dst[q] = make_uint4(pa[4 * q], pa[4 * q + 1], pa[4 * q + 2], pa[4 * q + 3]);
                        dg_dst[q] = make_uint4(pg[4 * q], pg[4 * q + 1], pg[4 * q + 2], pg[4 * q + 3]);
                    }
                }
                const float s1 = warp_transpose_sum(g, lane);
                const float s2 = warp_transpose_sum(gh, lane);
                part[e][0][ch * 32 + lane] = s1;
                part[e][1][ch * 32 + lane] = s2;
            }
            tc_fence_before();
            mbar_arrive(&bars[7 + ab]);                 // accumulator buffer drained
            mbar_arrive(&bars[2 + s]);                  // stage (H chunk) consumed
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;
            for (int c = t; c < 2 * 128; c += 128) {
                const int st = c >> 7, col = c & 127;
                const float tot = part[0][st][col] + part[1][st][col] + part[2][st][col] + part[3][st][col];
                atomicAdd(colsum + (int64_t)st * hidden + kc * 128 + col, tot);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, 256);
    if (threadIdx.x == 0 && fail_s != 0u) colsum[0] = CUDART_NAN_F;
}

// backward 2: BatchNorm backward, dH = gamma rstd (dHn - mean(dHn) - Hhat mean(dHn Hhat)) = a_c dHn + b_c H + d_c per column
// (the three coefficients per column sit in shared memory; a pure 48 MB stream); dgamma, dbeta from the column sums
__global__ void __launch_bounds__(256)
head_bwd2_kernel(const uint16_t *__restrict__ dhn, const uint16_t *__restrict__ h, const float *__restrict__ colsum,
                 const float *__restrict__ gamma, const float *__restrict__ save_mean, const float *__restrict__ save_rstd,
                 uint16_t *__restrict__ dh_out, float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows, int hidden,
                 int fp16)
{
    __shared__ float ca[1024], cb[1024], cd[1024];
    const float inv_r = 1.0f / (float)rows;
    for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
        const float rstd = save_rstd[c], mean = save_mean[c], g = gamma[c];
        const float m1 = colsum[c] * inv_r, m2 = colsum[hidden + c] * inv_r;
        // Hhat = (H - mean) rstd
        ca[c] = g * rstd;
        cb[c] = -g * rstd * rstd * m2;
        cd[c] = g * rstd * (mean * rstd * m2 - m1);
    }
    __syncthreads();
    const int groups = hidden / 8;                       // 16-byte groups per row
    const int64_t total8 = rows * groups;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % groups) * 8;
        const uint4 qg = __ldcs(reinterpret_cast<const uint4 *>(dhn) + i), qh = __ldg(reinterpret_cast<const uint4 *>(h) + i);
        const uint32_t wg[4] = {qg.x, qg.y, qg.z, qg.w}, wh[4] = {qh.x, qh.y, qh.z, qh.w};
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + 2 * u;
            const float r0 = fmaf(ca[c], unpack16(wg[u], 0, fp16 != 0), fmaf(cb[c], unpack16(wh[u], 0, fp16 != 0), cd[c]));
            const float r1 = fmaf(ca[c + 1], unpack16(wg[u], 1, fp16 != 0), fmaf(cb[c + 1], unpack16(wh[u], 1, fp16 != 0), cd[c + 1]));
            o[u] = pack16x2(r0, r1, fp16 != 0);
        }
        reinterpret_cast<uint4 *>(dh_out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < hidden; c += blockDim.x) {
            dbeta[c] = colsum[c];
            dgamma[c] = colsum[hidden + c];
        }
}

static int check_head_shapes(int64_t rows, int in_dim, int hidden, int out_dim)
{
    if (rows <= 0) return set_error(SMH_E_ARG, "head: rows must be positive");
    if (in_dim <= 0 || in_dim % 64) return set_error(SMH_E_DIM, "head: in_dim must be a positive multiple of 64 (got %d)", in_dim);
    if (hidden <= 0 || hidden % 256 || hidden > 1024)
        return set_error(SMH_E_DIM, "head: hidden must be a multiple of 256 and <= 1024 (got %d)", hidden);
    if (out_dim != 128) return set_error(SMH_E_DIM, "head: out_dim must be 128 (got %d)", out_dim);
    return 0;
}

int launch_head_fwd(const smh_head_t &hd, cudaStream_t stream)
{
    int rc = check_head_shapes(hd.rows, hd.in_dim, hd.hidden, hd.out_dim);
    if (rc) return rc;
    const bool fp16 = hd.fp16 != 0;
    CUtensorMap mx, mw1, mh, mw2;
    if ((rc = make_map_16(&mx, hd.x_dev, hd.rows, hd.in_dim, hd.x_row_stride, 256, fp16))) return rc;
    if ((rc = make_map_16(&mw1, hd.w1_dev, hd.hidden, hd.in_dim, hd.in_dim, 256, fp16))) return rc;
    if ((rc = make_map_16(&mh, hd.h_dev, hd.rows, hd.hidden, hd.hidden, 128, fp16))) return rc;
    if ((rc = make_map_16(&mw2, hd.w2_dev, hd.out_dim, hd.hidden, hd.hidden, 128, fp16))) return rc;
    cudaError_t e = cudaMemsetAsync(hd.colsum_dev, 0, sizeof(float) * 2 * hd.hidden, stream);
    if (e != cudaSuccess) return set_error((int)e, "head memset: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(head_gemm1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kG1Smem);
    if (e != cudaSuccess) return set_error((int)e, "head smem attr: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(head_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2Smem);
    if (e != cudaSuccess) return set_error((int)e, "head smem attr: %s", cudaGetErrorString(e));
    dim3 g1((unsigned)((hd.rows + 255) / 256), (unsigned)(hd.hidden / 256));
    head_gemm1_kernel<<<g1, kG1Threads, kG1Smem, stream>>>(mx, mw1, hd.b1_dev, (uint16_t *)hd.h_dev, hd.colsum_dev, (int)hd.rows,
                                                          hd.in_dim, hd.hidden, fp16 ? 1 : 0);
    if ((rc = check_launch("head_gemm1_kernel"))) return rc;
    BnParams bn;
    bn.colsum = hd.colsum_dev;
    bn.gamma = hd.gamma_dev;
    bn.beta = hd.beta_dev;
    bn.running_mean = hd.running_mean_dev;
    bn.running_var = hd.running_var_dev;
    bn.save_mean = hd.save_mean_dev;
    bn.save_rstd = hd.save_rstd_dev;
    bn.eps = hd.bn_eps;
    bn.momentum = hd.bn_momentum;
    bn.training = hd.training;
    if (!hd.training && !hd.running_mean_dev) return set_error(SMH_E_ARG, "head: eval mode needs the running estimates");
    head_gemm2_kernel<<<(unsigned)((hd.rows + 127) / 128), kG2Threads, kG2Smem, stream>>>(mh, mw2, bn, hd.y_dev, hd.norm_dev,
                                                                                         (int)hd.rows, hd.hidden, hd.out_dim,
                                                                                         hd.norm_eps, fp16 ? 1 : 0);
    return check_launch("head_gemm2_kernel");
}

int launch_head_bwd(const smh_head_t &hd, const smh_head_bwd_t &bw, cudaStream_t stream)
{
    int rc = check_head_shapes(hd.rows, hd.in_dim, hd.hidden, hd.out_dim);
    if (rc) return rc;
    const bool fp16 = hd.fp16 != 0;
    CUtensorMap mw2t, mh;
    if ((rc = make_map_16(&mw2t, bw.w2t_dev, hd.hidden, hd.out_dim, hd.out_dim, 128, fp16))) return rc;
    if ((rc = make_map_16(&mh, hd.h_dev, hd.rows, hd.hidden, hd.hidden, 128, fp16))) return rc;
    cudaError_t e = cudaMemsetAsync(bw.colsum_dev, 0, sizeof(float) * 2 * hd.hidden, stream);
    if (e != cudaSuccess) return set_error((int)e, "head memset: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(head_bwd1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kB1Smem);
    if (e != cudaSuccess) return set_error((int)e, "head smem attr: %s", cudaGetErrorString(e));
    head_bwd1_kernel<<<(unsigned)((hd.rows + 127) / 128), kB1Threads, kB1Smem, stream>>>(
        mw2t, mh, bw.dy_dev, hd.y_dev, hd.norm_dev, hd.gamma_dev, hd.beta_dev, hd.save_mean_dev, hd.save_rstd_dev,
        (uint16_t *)bw.dp_dev, (uint16_t *)bw.a_dev, (uint16_t *)bw.dhn_dev, bw.colsum_dev, (int)hd.rows, hd.hidden, hd.out_dim,
        hd.norm_eps, fp16 ? 1 : 0);
    if ((rc = check_launch("head_bwd1_kernel"))) return rc;
    int64_t blocks = (hd.rows * hd.hidden / 8 + 255) / 256;
    if (blocks > kNumCtas * 8) blocks = kNumCtas * 8;
    head_bwd2_kernel<<<(int)blocks, 256, 0, stream>>>((const uint16_t *)bw.dhn_dev, (const uint16_t *)hd.h_dev, bw.colsum_dev,
                                                      hd.gamma_dev, hd.save_mean_dev, hd.save_rstd_dev, (uint16_t *)bw.dh_dev,
                                                      bw.dgamma_dev, bw.dbeta_dev, hd.rows, hd.hidden, fp16 ? 1 : 0);
    return check_launch("head_bwd2_kernel");
}

}  // namespace smh
