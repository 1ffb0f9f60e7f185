// simhand_b200: declarations shared by the translation units of libsimhand_b200.so (host side).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/simhand_b200.h"
#include "smh_common.cuh"

namespace smh {

struct WsView {                 // device pointers carved out of the caller's workspace blob
    void *stats;                // Stats
    float *zt;
    uint16_t *zb;
    uint16_t *zh;
    float *jp;
    float *posd;
    float *neg;
    float *rn;
    float *rowloss;
    float *dzacc;
    float *negparts;
    float *dzparts;
    float *dist;
};

struct PlanView {               // device pointers into the caller's plan blob
    const int2 *tiles;          // (I, J) of each stored tile of this rank
    const int4 *tasks;          // (row block, 64-col tile, stored tile, flags)
    const int2 *strips;         // backward sweep: (first task, one-past-last task)
    const int *cta_ptr;         // strips of sweep CTA c: [cta_ptr[c], cta_ptr[c + 1])
    const int2 *strips_fwd;     // the forward sweep's own cuts of the same task list
    const int *cta_ptr_fwd;
};

int set_error(int code, const char *fmt, ...);
int check_launch(const char *what);

// kernel launchers (one per translation unit)
int launch_prep(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                int engine, cudaStream_t stream);
int launch_mpjpe(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                 const Peers &peers, cudaStream_t stream);
int launch_sweep_fp32(bool backward, int wmode /* 0 fused, 1 unit, 2 materialised */, const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan,
                      const WsView &ws, const Peers &peers, float temperature, cudaStream_t stream);
int launch_sweep_tc(bool backward, int logit_format /* 0 tf32, 1 bf16, 2 fp16 */, int wmode, const smh_dims_t &dims, const smh_layout_t &lay,
                    const PlanView &plan, const WsView &ws, const Peers &peers, const Peers &xp /* fused exchange, or world 1 */,
                    float temperature, cudaStream_t stream);
int launch_shard_prep(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine,
                      const Peers &peers, cudaStream_t stream);
int launch_shard_push_z(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, int engine,
                        const Peers &peers, cudaStream_t stream);
int launch_rn_fused(const smh_layout_t &lay, const WsView &ws, const Peers &peers, bool signal4, cudaStream_t stream);
int launch_finalize_fused(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                          int pos_mode, float temperature, float grad_scale, float *loss, float *dz1, float *dz2,
                          int64_t dz_row_stride, const Peers &peers, cudaStream_t stream);
int launch_push_inputs(const smh_exchange_t &exch, const smh_inputs_t &in, int n_local, int d, cudaStream_t stream);
int launch_barrier(const smh_exchange_t &exch, cudaStream_t stream);
int launch_rn(const smh_layout_t &lay, const WsView &ws, int n_parts, cudaStream_t stream);
int launch_exchange_neg(const smh_layout_t &lay, const Peers &peers, cudaStream_t stream);
int launch_exchange_dz(const smh_dims_t &dims, const smh_layout_t &lay, const Peers &peers, cudaStream_t stream);
int launch_import_weights(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                          const float *neg_w, int64_t neg_row_stride, const float *pos_w, cudaStream_t stream);
int launch_finalize(const smh_dims_t &dims, const smh_layout_t &lay, const smh_inputs_t &in, const WsView &ws,
                    const float *dzacc_src, bool local_block, int n_parts, int pos_mode /* 0 fused, 1 unit, 2 given */, float temperature, float grad_scale, float *loss,
                    float *dz1,
                    float *dz2, int64_t dz_row_stride, int phase /* 0 whole, 1 local loss part, 2 local grad + loss */,
                    const Peers &peers, cudaStream_t stream);
int launch_weights_dense(const smh_dims_t &dims, const smh_layout_t &lay, const PlanView &plan, const WsView &ws,
                         float *pos_w, float *neg_w, cudaStream_t stream);
int launch_scale_pair(const float *a, const float *b, const float *scale, float *oa, float *ob, int64_t count,
                      cudaStream_t stream);
int launch_l2norm_fwd(const float *x, float *y, float *norm, int64_t rows, int d, float eps, cudaStream_t stream);
int launch_l2norm_bwd(const float *y, const float *norm, const float *dy, float *dx, int64_t rows, int d,
                      float eps, cudaStream_t stream);
int launch_transform_fwd(const float *x, int64_t x_stride, const float *tx, const float *ty, const float *angle,
                         float *out, int64_t out_stride, float *save, int64_t rows, int d, float eps,
                         cudaStream_t stream);
int launch_transform_bwd(const float *x, int64_t x_stride, const float *p, int64_t p_stride, const float *save,
                         const float *dp, int64_t dp_stride, float *dx, int64_t dx_stride, int64_t rows, int d,
                         float eps, cudaStream_t stream);
int launch_head_fwd(const smh_head_t &hd, cudaStream_t stream);
int launch_head_bwd(const smh_head_t &hd, const smh_head_bwd_t &bw, cudaStream_t stream);
int launch_selftest(int which, uint64_t *out, int64_t out_words, cudaStream_t stream);

}  // namespace smh
