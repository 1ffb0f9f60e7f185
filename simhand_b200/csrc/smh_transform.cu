// simhand_b200 K4: fused projection-space transform.
//
// HandCLR_W / PeCLR_W.get_transformed_projections (src/models/unsupervised/simhand_w_model.py:55-94): every projection
// row of width d is d/2 2-D points (x_k, y_k):
//   y = x / max(||x||, eps)                                                         F.normalize        (:56-58)
//   y_x += tx (max_k y_x - min_k y_x),  y_y += ty (max_k y_y - min_k y_y)           translate_encodings (utils.py:661-684)
//   (c_x, c_y) = mean_k y;  a = cos(angle pi/180), b = sin(angle pi/180)            rotate_encoding     (utils.py:636-658)
//   r_x = a y_x + b y_y + (1 - a) c_x - b c_y;  r_y = -b y_x + a y_y + (1 - a) c_y + b c_x   (get_rotation_2D_matrix :606-633)
//   p = r / max(||r||, eps)                                                         F.normalize        (:91-93)
// The extents and the centre are detached in the reference, so the backward is normalise-bwd o rotation^T o normalise-bwd.
// The reference runs ~15 kernels and builds the rotation matrices on the HOST (utils.py:625, :652: a device->host->device
// round trip per step); here one warp handles one row in registers, forward and backward are one launch each.
#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// lane l holds columns 4l .. 4l+3 = points 2l (v[0], v[1]) and 2l+1 (v[2], v[3]); d is even and <= 128
__device__ __forceinline__ void load_row(const float *__restrict__ p, int lane, int d, float (&v)[4])
{
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (4 * lane + u < d) ? p[4 * lane + u] : 0.f;
}

__global__ void __launch_bounds__(256)
transform_fwd_kernel(const float *__restrict__ x, int64_t x_stride, const float *__restrict__ tx,
                     const float *__restrict__ ty, const float *__restrict__ angle, float *__restrict__ out,
                     int64_t out_stride, float *__restrict__ save, int64_t rows, int d, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    const float npts = (float)(d / 2);
    for (int64_t row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        float v[4];
        load_row(x + row * x_stride, lane, d, v);
        const bool p0 = 4 * lane < d, p1 = 4 * lane + 2 < d;               // which of this lane's two points exist
        const float n1 = sqrtf(warp_sum(fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])))));
        const float den1 = fmaxf(n1, eps);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __fdiv_rn(v[u], den1);
        if (tx != nullptr) {
            const float big = 3.0e38f;
            const float mxx = warp_max(fmaxf(p0 ? v[0] : -big, p1 ? v[2] : -big));
            const float mnx = warp_min(fminf(p0 ? v[0] : big, p1 ? v[2] : big));
            const float mxy = warp_max(fmaxf(p0 ? v[1] : -big, p1 ? v[3] : -big));
            const float mny = warp_min(fminf(p0 ? v[1] : big, p1 ? v[3] : big));
            const float sx = tx[row] * (mxx - mnx), sy = ty[row] * (mxy - mny);
            // only points that exist move: padding lanes (4 lane + u >= d) must stay 0 for the second norm
            if (p0) { v[0] += sx; v[1] += sy; }
            if (p1) { v[2] += sx; v[3] += sy; }
        }
        float a = 1.f, b = 0.f;
        if (angle != nullptr) {
            const float cx = warp_sum((p0 ? v[0] : 0.f) + (p1 ? v[2] : 0.f)) / npts;
            const float cy = warp_sum((p0 ? v[1] : 0.f) + (p1 ? v[3] : 0.f)) / npts;
            const float rad = __fdiv_rn(__fmul_rn(angle[row], 3.14159274101257324f), 180.0f);    // angle * np.pi / 180
            a = cosf(rad);
            b = sinf(rad);
            const float ox = (1.f - a) * cx - b * cy, oy = (1.f - a) * cy + b * cx;
            const float x0 = v[0], y0 = v[1], x1 = v[2], y1 = v[3];
            v[0] = fmaf(a, x0, fmaf(b, y0, ox));
            v[1] = fmaf(-b, x0, fmaf(a, y0, oy));
            v[2] = fmaf(a, x1, fmaf(b, y1, ox));
            v[3] = fmaf(-b, x1, fmaf(a, y1, oy));
            if (!p0) v[0] = v[1] = 0.f;
            if (!p1) v[2] = v[3] = 0.f;
        }
        const float n2 = sqrtf(warp_sum(fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], v[3] * v[3])))));
        const float den2 = fmaxf(n2, eps);
        float *o = out + row * out_stride;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (4 * lane + u < d) o[4 * lane + u] = __fdiv_rn(v[u], den2);
        if (lane == 0) *reinterpret_cast<float4 *>(save + 4 * row) = make_float4(n1, n2, a, b);
    }
}

// dx from dp: p = out of the forward, save = (||x||, ||r||, cos, sin) per row
__global__ void __launch_bounds__(256)
transform_bwd_kernel(const float *__restrict__ x, int64_t x_stride, const float *__restrict__ p, int64_t p_stride,
                     const float *__restrict__ save, const float *__restrict__ dp, int64_t dp_stride,
                     float *__restrict__ dx, int64_t dx_stride, int64_t rows, int d, float eps)
{
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    for (int64_t row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
        float xv[4], pv[4], g[4];
        load_row(x + row * x_stride, lane, d, xv);
        load_row(p + row * p_stride, lane, d, pv);
        load_row(dp + row * dp_stride, lane, d, g);
        const float4 sv = *reinterpret_cast<const float4 *>(save + 4 * row);
        const float n1 = sv.x, n2 = sv.y, a = sv.z, b = sv.w;
        // second normalise: dr = (dp - p (p . dp)) / max(||r||, eps)  (a plain scale below eps, as F.normalize)
        float dot = warp_sum(fmaf(pv[0], g[0], fmaf(pv[1], g[1], fmaf(pv[2], g[2], pv[3] * g[3]))));
        if (n2 < eps) dot = 0.f;
        const float inv2 = 1.0f / fmaxf(n2, eps);
        float dr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) dr[u] = (g[u] - pv[u] * dot) * inv2;
        // rotation transposed (the centre and the translation extents are detached): dy = R^T dr
        float dy[4];
        dy[0] = a * dr[0] - b * dr[1];
        dy[1] = b * dr[0] + a * dr[1];
        dy[2] = a * dr[2] - b * dr[3];
        dy[3] = b * dr[2] + a * dr[3];
        // first normalise
        const float inv1 = 1.0f / fmaxf(n1, eps);
        float yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) yv[u] = xv[u] * inv1;
        float dot1 = warp_sum(fmaf(yv[0], dy[0], fmaf(yv[1], dy[1], fmaf(yv[2], dy[2], yv[3] * dy[3]))));
        if (n1 < eps) dot1 = 0.f;
        float *o = dx + row * dx_stride;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (4 * lane + u < d) o[4 * lane + u] = (dy[u] - yv[u] * dot1) * inv1;
    }
}

int launch_transform_fwd(const float *x, int64_t x_stride, const float *tx, const float *ty, const float *angle,
                         float *out, int64_t out_stride, float *save, int64_t rows, int d, float eps,
                         cudaStream_t stream)
{
    int64_t blocks = (rows + 7) / 8;
    if (blocks > kNumCtas * 8) blocks = kNumCtas * 8;
    transform_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, x_stride, tx, ty, angle, out, out_stride, save, rows, d, eps);
    return check_launch("transform_fwd_kernel");
}

int launch_transform_bwd(const float *x, int64_t x_stride, const float *p, int64_t p_stride, const float *save,
                         const float *dp, int64_t dp_stride, float *dx, int64_t dx_stride, int64_t rows, int d,
                         float eps, cudaStream_t stream)
{
    int64_t blocks = (rows + 7) / 8;
    if (blocks > kNumCtas * 8) blocks = kNumCtas * 8;
    transform_bwd_kernel<<<(int)blocks, 256, 0, stream>>>(x, x_stride, p, p_stride, save, dp, dp_stride, dx, dx_stride,
                                                          rows, d, eps);
    return check_launch("transform_bwd_kernel");
}

}  // namespace smh
