/*
 * simhand_b200 -- C ABI of the B200-native similarity-weighted NT-Xent hot path.
 *
 * This is the drop-in boundary for the loss operator of ut-vision/SiMHand.  The reference has no
 * native code: its boundary is two Python functions,
 *     get_weights_linear(joints1, joints2, diff_type)                    src/models/utils.py:218-261
 *     vanila_weights_contrastive_loss(z1, z2, pos_w, neg_w, temperature)  src/models/utils.py:391-427
 * called from contrastive_step() (src/models/unsupervised/simhand_w_model.py:122-136,
 * peclr_w_model.py:110-123, simclr_w_model.py:71-96).  The host mirror of those two functions lives in
 * simhand_b200/ops.py and binds the entry points below with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer named *_dev is a CUDA device pointer owned by the caller (PyTorch's caching
 *     allocator); the library never allocates, frees or retains device memory;
 *   - every call is asynchronous on the given stream and never synchronises the device;
 *   - return value: 0 = ok, < 0 = argument error (SMH_E_*), > 0 = cudaError_t of a failed launch;
 *     smh_last_error() returns a thread-local description of the last failure;
 *   - re-entrant: no global mutable state (safe under nn.DataParallel's one-thread-per-device use,
 *     src/experiments/main.py:155);
 *   - built for sm_100a only.  There is no CPU path and no other backend.
 *
 * One training step of the path is the sequence
 *     smh_prep -> smh_mpjpe -> [all-reduce MAX of stats, world > 1] -> smh_forward ->
 *     [all-reduce SUM of neg] -> smh_backward -> [reduce-scatter of dz] -> smh_finalize
 * over one workspace blob whose layout smh_layout() reports ("Data layout in HBM").
 */
#ifndef SIMHAND_B200_H
#define SIMHAND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMH_VERSION 100          /* 0.1.0 */
#define SMH_NUM_JOINTS 21        /* src/data_loader: joints are [B, 21, 3], the path uses [:, :, :2] */
#define SMH_MAX_DIM 128          /* projection width d <= 128 (reference: output_dim 128) */

/* error codes (negative) */
#define SMH_E_ARG (-1)           /* null pointer / non-positive size */
#define SMH_E_DIM (-2)           /* unsupported d, J or world/rank combination */
#define SMH_E_ALIGN (-3)         /* pointer or stride not aligned as documented */
#define SMH_E_SIZE (-4)          /* caller buffer too small */
#define SMH_E_ARCH (-5)          /* device is not sm_100 */
#define SMH_E_MODE (-6)          /* unknown engine / mode */

/* engine of the forward/backward sweeps (the dense contraction S = z z^T and dz = G z) */
#define SMH_ENGINE_TC_TF32 0     /* tcgen05: tf32 logits in the forward sweep, bf16 operands in the backward sweep */
#define SMH_ENGINE_FP32 1        /* CUDA-core FFMA, fp32 accumulate (exact-fp32 mode) */
#define SMH_ENGINE_TC_BF16 2     /* tcgen05: bf16 operands in both sweeps (bf16 mode) */
#define SMH_ENGINE_TC_FP16 3     /* tcgen05: fp16 logit operands in the forward sweep (11-bit significand, the precision of
                                  * tf32, at half the staging traffic; |z| <= 1 after L2 normalisation so the fp16 range is
                                  * no issue), bf16 operands in the backward sweep */
#define SMH_BACKWARD_RN_ONLY 0x200 /* OR into smh_backward's engine: only reduce the row sums (loss without gradient) */
#define SMH_UNIT_NEG_WEIGHTS 0x400 /* OR into smh_forward/backward's engine and smh_finalize's flags: W_ij == 1 (the reference's
                                    * vanila_pos_weights_contrastive_loss / vanila_contrastive_loss; smh_mpjpe can be skipped) */
#define SMH_UNIT_POS_WEIGHTS 0x800 /* OR into smh_finalize's flags: Wp_k == 1 (vanila_neg_weights_contrastive_loss, vanila_contrastive_loss) */
#define SMH_PREP_NO_ZERO 0x100   /* OR into smh_prep's engine: the accumulators were already zeroed by smh_prep_zero */
#define SMH_DENSE_WEIGHTS 0x2000 /* OR into smh_forward/backward's engine and smh_finalize's flags: the weights are the
                                  * materialised tensors handed to smh_import_weights (dims.flags has SMH_DIMS_DENSE_WEIGHTS) */

/* Sharded finalize (peer exchange; OR into smh_finalize's flags): every rank evaluates the loss terms of its OWN rows only.
 * LOSS_PART (after smh_backward, before the barrier that follows smh_exchange_dz): row terms of the local rows, their sum
 * stored into slot `rank` of every peer's partial-loss array.  GRAD (after that barrier): gradients of the local rows and
 * loss = rank-ordered sum of the partials / M.  Nothing reads the gathered inputs of other ranks after smh_prep, so the step
 * needs no closing barrier. */
#define SMH_FINALIZE_LOSS_PART 0x4000
#define SMH_FINALIZE_GRAD 0x8000
/* Fused exchange (exch->fused): smh_finalize takes the rank's LOCAL inputs (as smh_shard_prep), waits for the gradient
 * partials of every rank, writes the local gradients and the loss of the GLOBAL batch, which every rank evaluates in the
 * same fixed order from the delivered row sums and positive-pair terms (no further exchange). */

/* smh_dims_t.flags */
#define SMH_DIMS_DENSE_WEIGHTS 1  /* materialised-weights path (the reference's two-call API with real tensors,
                                   * utils.py:391): every (I, J) tile is stored, nothing is assumed symmetric; world == 1 */
#define SMH_DIMS_Q16_TILES 4      /* tensor-core engines, linear / mpjpe weighting: the distance tiles are stored as 16-bit
                                   * fixed point q = round(D * 65000 / Dbound), Dbound = 2 max_i D(i, 0) >= Dmax (triangle
                                   * inequality), which halves the bytes the sweeps stage per tile.  |W error| <= 1.6e-5, far
                                   * below the 2^-11 operand rounding of the logits.  Dmax itself stays exact.  The fp32
                                   * engine, smh_weights_dense and the other weightings need the fp32 tiles (flag clear). */
#define SMH_DIMS_DENSE_BACKWARD 2 /* with SMH_DIMS_DENSE_WEIGHTS: the task list of the backward sweep, which visits every
                                   * tile twice (W_ij for the row term, W_ji read transposed for the column term).  Same
                                   * layout as the forward list: one workspace, two plans. */

/* problem description shared by all calls */
typedef struct smh_dims {
    int32_t n;                   /* global per-view batch N; the loss sees M = 2N samples */
    int32_t d;                   /* projection width, 1..SMH_MAX_DIM */
    int32_t world;               /* ranks sharing the batch (1 = single GPU) */
    int32_t rank;                /* this rank, 0..world-1 */
    int32_t strip_len;           /* sweep tasks per strip (<= 0: library default) */
    int32_t flags;               /* SMH_DIMS_*; 0 = fused path (weights from the joints) */
    int32_t diff_type;           /* SMH_DIFF_*: which joint distance feeds the weights (utils.py:219-231, :241-253) */
    int32_t weight_type;         /* SMH_WEIGHT_*: linear (utils.py:218) or non_linear (utils.py:304) */
    float lambda_pos;            /* non_linear: Wp = 1 / (1 + exp(lambda_pos (D - mean D)))   (utils.py:323-325) */
    float lambda_neg;            /* non_linear: Wn = 1 / (1 + exp(lambda_neg (D - mean D)))   (utils.py:343-346) */
} smh_dims_t;

/* smh_dims_t.diff_type: distance between two samples' 21 x 2 joints a, b */
#define SMH_DIFF_MPJPE 0    /* mean_k ||a_k - b_k||                                   (bit-exact with torch-CPU) */
#define SMH_DIFF_W_ABS 1    /* negatives: || ((|dx_k| + |dy_k|) / 2)_k ||; positives: || mean_k (|dx_k|, |dy_k|) || */
#define SMH_DIFF_W_O_ABS 2  /* negatives: || ((dx_k + dy_k) / 2)_k ||;     positives: || mean_k (dx_k, dy_k) ||
                             * (the reference reduces over different axes for the two, utils.py:219-227 vs :241-249) */
#define SMH_DIFF_EUCLID 3   /* || a - b ||_2 over all 42 coordinates of the joint block: the *_with_pca weightings
                             * (utils.py:264-301, :349-388), whose inputs are [N, K <= 42] PCA coordinates (apply_pca,
                             * utils.py:192-215) handed over zero-padded as [N, 21, 2]; all three reference diff_types
                             * reduce to this distance there */
/* smh_dims_t.weight_type */
#define SMH_WEIGHT_LINEAR 0     /* W = (max D - D) / (max D - min D) */
#define SMH_WEIGHT_NONLINEAR 1  /* W = 1 / (1 + exp(lambda (D - mean D))); single rank */

/* byte offsets into the workspace blob (all 256-byte aligned) and table sizes */
typedef struct smh_layout {
    int64_t ws_bytes;            /* total workspace size */
    int64_t plan_bytes;          /* size of the task plan (host build, device copy by the caller) */
    int64_t off_stats;           /* smh_stats_t */
    int64_t off_zt;              /* [Tp*128][128] fp32, tf32-rounded z, pre-swizzled 64-row blocks */
    int64_t off_zb;              /* [Tp*128][128] bf16 copy of z, pre-swizzled 64-row blocks (backward value operand) */
    int64_t off_zh;              /* [Tp*128][128] fp16 copy of z, same block layout as zb (forward logit operand, fp16 engine) */
    int64_t off_jp;              /* [Tp*128][44] fp32 packed joints */
    int64_t off_posd;            /* [N] fp32 positive-pair MPJPE */
    int64_t off_neg;             /* [Tp*128] fp32 off-diagonal row sums (partial until all-reduced) */
    int64_t off_rn;              /* [Tp*128] fp32 1/neg */
    int64_t off_rowloss;         /* [Tp*128] fp32 per-row loss terms */
    int64_t off_dzacc;           /* [M][128] fp32 unscaled gradient accumulator (rank-major rows) */
    int64_t off_negparts;        /* [world][Tp*128] fp32 row-sum partials received from the ranks (peer exchange) */
    int64_t off_dzparts;         /* [world][2*n_local][128] fp32 gradient partials received from the ranks (peer exchange) */
    int64_t off_dist;            /* stored MPJPE tiles of this rank, 64 KiB each */
    int64_t off_posinfo;         /* (world > 1) [2 parities][2][N] fp32: positive-pair distance and <z1_k, z2_k> of every
                                  * sample, delivered by the owning ranks (fused exchange) */
    int32_t m;                   /* 2N */
    int32_t tiles_per_side;      /* Tp = ceil(M / 128) */
    int32_t n_stored_tiles;      /* upper-triangular 128x128 tiles assigned to this rank */
    int32_t n_tasks;             /* 128x64 sweep tasks of this rank */
    int32_t n_strips;            /* groups of consecutive tasks sharing a row block (backward sweep's cuts) */
    int32_t strip_len;           /* max tasks per strip */
    int32_t n_strips_fwd;        /* the same for the forward sweep's cuts of the task list */
    int32_t reserved;
} smh_layout_t;

/* device-resident scalars.  The first three words are order-preserving integer images of the
 * floats, so a multi-rank run combines them with one all-reduce(MAX) on int32[3]. */
typedef struct smh_stats {
    uint32_t dmax_bits;          /* float bits of max_ij D_ij (D >= 0)            (utils.py:255) */
    uint32_t pmax_bits;          /* float bits of max_k D_{k,k+N}                 (utils.py:233) */
    uint32_t pmin_inv;           /* 0x7fffffff - float bits of min_k D_{k,k+N}    (utils.py:234) */
    uint32_t flags;              /* SMH_FLAG_* */
    float loss;                  /* final loss (also written to the caller's pointer) */
    uint32_t counter;            /* internal: last-block-done ticket */
    uint32_t fail_site;          /* != 0: a bounded pipeline wait timed out (result invalid) */
    uint32_t ticket2;            /* internal: last-block-done ticket of the MPJPE kernel (peer exchange) */
    double dsum;                 /* non_linear weights: sum_ij D_ij over all ordered pairs (mean = dsum / M^2, utils.py:345) */
    uint32_t dbound_bits;        /* float bits of max_i D(i, 0): 2x this bounds every D_ij (SMH_DIMS_Q16_TILES scale) */
    uint32_t reserved;
} smh_stats_t;

#define SMH_FLAG_SLOW_DOMAIN 1u  /* joints outside the fast exact-sqrt domain: IEEE slow path used */
#define SMH_FLAG_NONFINITE 2u    /* a joint coordinate is NaN/Inf: loss is NaN, as in the reference */

/* inputs as the reference hands them over: two projections and two joint views.
 * joints are fp32 views [N, 21, 2] with arbitrary element strides (the callers pass the
 * non-contiguous joints[:, :, :2] slice of a [N, 21, 3] tensor, simhand_w_model.py:103-104).
 * For world > 1 the buffers are the all-gathered ones: sample k of view v lives at
 *   base_v + (k / n_local) * rank_stride + (k % n_local) * row_stride            (elements). */
typedef struct smh_inputs {
    const float *z1_dev, *z2_dev;        /* [N, d] fp32, rows L2-normalised by the caller */
    int64_t z_row_stride;                /* elements between consecutive samples (>= d) */
    const float *j1_dev, *j2_dev;        /* [N, 21, 2] fp32 views */
    int64_t j_sample_stride, j_joint_stride, j_coord_stride;   /* elements */
    int32_t n_local;                     /* samples per rank per view (N when world == 1) */
    int64_t z_rank_stride, j_rank_stride;/* elements between rank chunks (ignored when world == 1) */
} smh_inputs_t;

/* Peer exchange (world > 1, optional): the ranks' workspaces and input staging buffers are symmetric allocations
 * mapped into every process (torch.distributed._symmetric_memory or CUDA IPC); with it the library does the
 * collectives itself over NVLink: the MPJPE kernel's last CTA pushes Dmax to every peer (atomicMax);
 * smh_exchange_neg / smh_exchange_dz store this rank's partial row sums / gradient rows into slot `rank` of every
 * peer's (resp. the owning peer's) partial buffers with plain 16-byte stores, and the consumers (1/neg kernel,
 * finalize) add the `world` partials in rank order, so the reduction is deterministic; smh_barrier separates the
 * phases.  Without it (exch == NULL) the caller runs all-reduce / reduce-scatter between the calls. */
#define SMH_MAX_PEERS 8
#define SMH_SIGNAL_WORDS 256             /* uint32 words of a rank's signal block */
#define SMH_SIG_EPOCH 16                 /* word: steps this rank has started (fused exchange) */
#define SMH_SIG_POISON 17                /* word: != 0 once any cross-rank wait of the group timed out (sticky; every later
                                          * loss is NaN until the exchange is rebuilt); value = site of the first failure */
typedef struct smh_exchange {
    int32_t world, rank;
    void *ws_peer[SMH_MAX_PEERS];        /* workspace blob of every rank (ws_peer[rank] == the local ws_dev) */
    void *xin_peer[SMH_MAX_PEERS];       /* gathered-input buffer of every rank: world x chunk floats (unused when fused) */
    void *signal_peer[SMH_MAX_PEERS];    /* SMH_SIGNAL_WORDS x uint32 of every rank (zero-initialised once) */
    int32_t fused;                       /* 0: phases separated by smh_barrier launches (smh_push_inputs / smh_prep /
                                          * smh_exchange_*).  1: fused exchange -- smh_shard_prep ships every rank's operand
                                          * images, and the cross-rank signals and payloads ride in the heads and tails of
                                          * smh_mpjpe / smh_forward / smh_backward / smh_finalize (6 launches per step, no
                                          * barrier or exchange kernels).  The workspaces must be zero-filled once. */
    uint32_t timeout_ms;                 /* bound of every cross-rank wait; 0 = 30000.  A timeout poisons the group
                                          * (SMH_SIG_POISON on every rank) and the loss is NaN: never a silent wrong result */
} smh_exchange_t;

int smh_version(void);
const char *smh_last_error(void);

/* sizes and offsets for a problem. */
int smh_layout(const smh_dims_t *dims, smh_layout_t *out);

/* builds the task plan of this rank into caller-owned host memory (layout.plan_bytes); the caller
 * copies it to the device once per dims and passes the device copy to the sweeps. */
int smh_plan_build(const smh_dims_t *dims, void *plan_host, int64_t plan_bytes);

/* K3a: packs joints, stages z into sweep tiles (rounded to tf32 for SMH_ENGINE_TC_TF32, unrounded for
 * SMH_ENGINE_FP32), positive-pair MPJPE (utils.py:229-231), zeroes the accumulators.  Replaces the
 * torch.cat calls of utils.py:237-239 and :407. */
int smh_prep(const smh_dims_t *dims, const smh_inputs_t *in, void *ws_dev, int engine, void *stream);
/* zeroes the accumulators only (peer exchange: must precede the barrier after which peers may add into them) */
int smh_prep_zero(const smh_dims_t *dims, void *ws_dev, void *stream);

/* Materialised-weights path (dims.flags & SMH_DIMS_DENSE_WEIGHTS; replaces smh_mpjpe): copies neg_w [M, M] (fp32, row
 * stride neg_row_stride elements; any values, not assumed symmetric) into the tile layout the sweeps stage, and pos_w
 * [N] into ws.posd.  Either pointer may be NULL (that operand then comes from the unit-weight flags).  smh_prep may be
 * called with NULL joints on this path. */
int smh_import_weights(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, const float *neg_w_dev,
                       int64_t neg_row_stride, const float *pos_w_dev, void *stream);

/* K0: all-pairs MPJPE tiles of this rank (upper triangle) + running max (utils.py:251-255). */
int smh_mpjpe(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, const smh_exchange_t *exch, void *stream);

/* K1: weighted logits, exp and off-diagonal row sums (utils.py:411-417) -> ws.neg (partial sums of
 * this rank's tasks for all M rows). */
int smh_forward(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, float temperature,
                int engine, const smh_exchange_t *exch, void *stream);

/* K2: dz accumulation (autograd of utils.py:411-426) -> ws.dzacc (partial sums of
 * this rank's tasks for all M rows, rows in rank-major order).  Needs ws.neg complete. */
int smh_backward(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, float temperature,
                 int engine, const smh_exchange_t *exch, void *stream);

/* Fused exchange, first launch of a step (replaces smh_push_inputs + smh_prep_zero + smh_barrier + smh_prep): every rank
 * converts its OWN 2 * n_local rows (local_in: this rank's z1/z2/joints, rank strides ignored) into the operand images of
 * the selected engine and the packed joints and stores them into every rank's workspace over NVLink; the positive-pair
 * distance and <z1_k, z2_k> of its samples go to every rank's posinfo; the max / min / domain flags / distance bound are
 * combined with remote atomicMax; the local accumulators are zeroed; stage 1 is signalled.  The distance bound is taken
 * against global sample 0 (rank 0 publishes its joints first), so the 16-bit image has the scale of the single-GPU
 * step. */
int smh_shard_prep(const smh_dims_t *dims, const smh_inputs_t *local_in, void *ws_dev, const smh_exchange_t *exch,
                   int engine, void *stream);
/* OR into smh_shard_prep's engine: the operand images of z are NOT shipped by smh_shard_prep but by smh_shard_push_z, which
 * the caller launches on a second stream right after smh_shard_prep (a parallel branch of the step's CUDA graph): its
 * NVLink transfer (3/4 of the bytes a rank ships) then runs under the MPJPE kernel, which needs only the joints.  The
 * forward sweep waits for it (stage 5); the caller joins the streams before smh_forward. */
#define SMH_SHARD_PREP_NO_IMAGES 0x10000
int smh_shard_push_z(const smh_dims_t *dims, const smh_inputs_t *local_in, void *ws_dev, const smh_exchange_t *exch,
                     int engine, void *stream);

/* peer exchange: pack this rank's local inputs (n_local samples per view, described like smh_inputs_t with
 * rank strides ignored) as [z1|z2|joints1|joints2] into slot `rank` of every peer's gathered-input buffer
 * (the all-gather of SURVEY.md 8e, push-based over NVLink); chunk = 2 * n_local * (d + 42) floats */
int smh_push_inputs(const smh_exchange_t *exch, const smh_inputs_t *local_in, int32_t n_local, int32_t d, void *stream);
/* peer exchange: all-gather of the partial row sums (after smh_forward) / reduce-scatter payload of the partial
 * gradient rows (after smh_backward) into the peers' partial buffers */
int smh_exchange_neg(const smh_dims_t *dims, void *ws_dev, const smh_exchange_t *exch, void *stream);
int smh_exchange_dz(const smh_dims_t *dims, void *ws_dev, const smh_exchange_t *exch, void *stream);
/* peer exchange: device-side barrier over all ranks (monotonic counters in signal_peer; CUDA-graph safe) */
int smh_barrier(const smh_exchange_t *exch, void *stream);

/* loss (utils.py:420-426) over all M rows and, if dz1_dev != NULL, the gradients of the local
 * samples: dz = dzacc_src / (M tau) - 2 Wp z_partner / (M tau), scaled by grad_scale.
 * dzacc_src_dev: NULL = all M rows in ws.dzacc (world == 1); otherwise this rank's own [2 * n_local][128] block
 * (the reduce-scattered buffer).  With exch != NULL the gradient rows are the sum of the `world` partial blocks in
 * ws.dzparts (dzacc_src_dev is ignored). */
int smh_finalize(const smh_dims_t *dims, const smh_inputs_t *in, void *ws_dev,
                 const float *dzacc_src_dev, float temperature, float grad_scale,
                 float *loss_dev, float *dz1_dev, float *dz2_dev, int64_t dz_row_stride,
                 int flags, const smh_exchange_t *exch, void *stream);

/* autograd backward of the fused loss (the gradients were produced with the forward): out1 = *scale_dev * dz1,
 * out2 = *scale_dev * dz2 over `count` contiguous floats each, one launch (the upstream gradient of the 0-dim loss
 * stays on the device). */
int smh_scale_grads(const float *dz1_dev, const float *dz2_dev, const float *scale_dev, float *out1_dev, float *out2_dev,
                    int64_t count, void *stream);

/* materialised weights with the reference's return shapes: pos_w [N], neg_w [M, M] row-major
 * (utils.py:235, :259).  world == 1 only.  Needs smh_prep + smh_mpjpe. */
int smh_weights_dense(const smh_dims_t *dims, const void *plan_dev, void *ws_dev,
                      float *pos_w_dev, float *neg_w_dev, void *stream);

/* K3: row-wise L2 normalisation y = x / max(||x||, eps) and its backward
 * (F.normalize in simhand_w_model.py:56-58,91-93). */
int smh_l2norm_fwd(const float *x_dev, float *y_dev, float *norm_dev, int64_t rows, int32_t d,
                   float eps, void *stream);
int smh_l2norm_bwd(const float *y_dev, const float *norm_dev, const float *dy_dev, float *dx_dev,
                   int64_t rows, int32_t d, float eps, void *stream);

/* K4: fused projection-space transform of HandCLR_W / PeCLR_W.get_transformed_projections
 * (src/models/unsupervised/simhand_w_model.py:55-94, peclr_w_model.py:52-91): every row of x [rows, d] (d even, <= 128)
 * is d/2 2-D points; out = normalize(rotate(translate(normalize(x)))), i.e. F.normalize, translate_encodings
 * (utils.py:661-684: shift by tx/ty times the detached per-row extent), rotate_encoding (utils.py:636-658: rotation by
 * `angle_deg` about the detached per-row centroid, matrix of get_rotation_2D_matrix :606-633 built on the device),
 * F.normalize.  tx_dev/ty_dev (both or neither) and angle_deg_dev may be NULL (that stage is skipped, as when "crop" /
 * "rotate" is not in config.augmentation).  save_dev [rows][4] receives (||x||, ||rotated||, cos, sin) for the backward.
 * Strides are in elements. */
int smh_transform_fwd(const float *x_dev, int64_t x_row_stride, const float *tx_dev, const float *ty_dev,
                      const float *angle_deg_dev, float *out_dev, int64_t out_row_stride, float *save_dev,
                      int64_t rows, int32_t d, float eps, void *stream);
/* backward of smh_transform_fwd: dx = J^T dout (normalise-bwd o rotation^T o normalise-bwd; extents/centroid detached) */
int smh_transform_bwd(const float *x_dev, int64_t x_row_stride, const float *out_dev, int64_t out_row_stride,
                      const float *save_dev, const float *dout_dev, int64_t dout_row_stride, float *dx_dev,
                      int64_t dx_row_stride, int64_t rows, int32_t d, float eps, void *stream);

/* K5: the projection head fused with the first normalisation (src/models/unsupervised/simclr_model.py:22-39
 * Linear(in, hidden, bias) -> BatchNorm1d(hidden) in training mode -> ReLU -> Linear(hidden, out, no bias), followed by the
 * F.normalize of simhand_w_model.py:56-58), 16-bit operands (the reference trains under 16-bit autocast), fp32 accumulation.
 * smh_head_forward = two tcgen05 kernels: GEMM 1 with the BatchNorm column statistics reduced in its epilogue; GEMM 2 with
 * BN + ReLU applied in shared memory to its A operand and the row-wise L2 normalisation in its epilogue.
 * smh_head_backward = normalise-backward + dA = dP W2 (tcgen05) + ReLU mask + BatchNorm-backward reductions in one kernel and
 * the BatchNorm backward in a second: it produces dP, A = relu(bn(H)) and dH; the three plain GEMMs left
 * (dW2 = dP^T A, dW1 = dH^T X, dX = dH W1) are library GEMMs on the caller's side.
 * Shapes: in_dim % 64 == 0, hidden % 256 == 0 (<= 1024), out_dim == 128; 16-bit pointers 16-byte aligned. */
typedef struct smh_head {
    int64_t rows;                /* 2B */
    int32_t in_dim, hidden, out_dim;
    int32_t fp16;                /* 0: bf16 activations / weights, 1: fp16 */
    int32_t training;            /* 1: batch statistics (and running estimates updated); 0: eval, running estimates used */
    int32_t reserved;
    const void *x_dev;           /* [rows, in_dim] 16-bit encodings */
    int64_t x_row_stride;        /* elements */
    const void *w1_dev;          /* [hidden, in_dim] 16-bit */
    const float *b1_dev;         /* [hidden] fp32 */
    const float *gamma_dev, *beta_dev;            /* [hidden] fp32 BatchNorm affine */
    float *running_mean_dev, *running_var_dev;    /* [hidden] fp32, updated in place; NULL = not tracked */
    float bn_eps, bn_momentum;
    const void *w2_dev;          /* [out_dim, hidden] 16-bit */
    void *h_dev;                 /* out: [rows, hidden] 16-bit, Linear-1 output (kept for the backward) */
    float *colsum_dev;           /* scratch: [2, hidden] fp32 */
    float *save_mean_dev, *save_rstd_dev;         /* out: [hidden] fp32 batch statistics */
    float *y_dev;                /* out: [rows, out_dim] fp32 L2-normalised projections */
    float *norm_dev;             /* out: [rows] fp32 row norms before the normalisation */
    float norm_eps;              /* F.normalize eps (1e-12) */
} smh_head_t;

typedef struct smh_head_bwd {
    const float *dy_dev;         /* [rows, out_dim] fp32 gradient w.r.t. the normalised projections */
    const void *w2t_dev;         /* [hidden, out_dim] 16-bit: W2 transposed */
    void *dp_dev;                /* out: [rows, out_dim] 16-bit gradient w.r.t. Linear-2 output */
    void *a_dev;                 /* out: [rows, hidden] 16-bit relu(bn(H)) (for dW2 = dP^T A) */
    void *dhn_dev;               /* scratch: [rows, hidden] 16-bit */
    void *dh_dev;                /* out: [rows, hidden] 16-bit gradient w.r.t. Linear-1 output */
    float *colsum_dev;           /* scratch: [2, hidden] fp32 */
    float *dgamma_dev, *dbeta_dev;                /* out: [hidden] fp32 */
} smh_head_bwd_t;

int smh_head_forward(const smh_head_t *head, void *stream);
int smh_head_backward(const smh_head_t *head, const smh_head_bwd_t *bwd, void *stream);

/* device self-tests used by tests/.  out_dev receives out[0] = values tested, out[1] = failures, out[2] = first failing input
 * bits, out[3] = worst case.  which: 0 exact sqrt, 1 packed exact sqrt, 2 exact x / 21, 3 exact weight division (out[3] in
 * ulp, all against the IEEE intrinsics over the whole domain); 4 the MUFU-free square root of the 16-bit tile image against
 * the double-precision sqrt (bound 7.5e-7 relative, 8.0e-7 for the accumulating form; out[3] in units of 1e-9). */
int smh_selftest(int which, uint64_t *out_dev, int64_t out_words, void *stream);

/* diagnostic: one tcgen05 tile S = A B^T (tf32) and dZ = bf16(S) Z_B (bf16 value operand, MN-major) on staged
 * z blocks with caller-supplied descriptor fields (see smh_selftest.cu); tests/ pins the encodings hard-coded
 * in the sweeps with it. */
int smh_tc_probe(const float *zt_dev, const void *zb_dev, int blk_a, int blk_b, const uint32_t *params16_host,
                 float *s_out_dev, uint32_t *g_out_dev, float *dz_out_dev, uint32_t *fail_dev, void *stream);
void smh_tc_default_params(uint32_t *params16_host);

#ifdef __cplusplus
}
#endif
#endif /* SIMHAND_B200_H */
