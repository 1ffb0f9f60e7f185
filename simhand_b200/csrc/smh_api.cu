// simhand_b200: extern "C" entry points, workspace layout and the host-side task plan.
//
// The plan decides which 128x128 MPJPE tiles a rank stores (upper triangle, balanced over ranks)
// and which 128x64 sweep tasks it runs in the forward/backward sweeps; see DESIGN.md "Task plan".
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "smh_common.cuh"
#include "smh_internal.h"

namespace smh {

static thread_local char g_err[512] = "";
// cost of opening a strip in sweep-task times (enumerate_plan): backward (gradient-accumulator flush, row-block reload,
// pipeline restart) and forward (128 atomics)
constexpr double kStripCostBwd = 4.0, kStripCostFwd = 0.5;
constexpr double kCostTransposed = 0.0, kCostMasked = 0.0;      // extra cost of a transposed / masked task (see task_cost)

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int validate_dims(const smh_dims_t *dims)
{
    if (!dims) return set_error(SMH_E_ARG, "dims is null");
    if (dims->n <= 0) return set_error(SMH_E_ARG, "n must be positive (got %d)", dims->n);
    if (dims->d <= 0 || dims->d > SMH_MAX_DIM)
        return set_error(SMH_E_DIM, "d must be in 1..%d (got %d)", SMH_MAX_DIM, dims->d);
    if (dims->world <= 0 || dims->rank < 0 || dims->rank >= dims->world)
        return set_error(SMH_E_DIM, "bad world/rank %d/%d", dims->world, dims->rank);
    if (dims->n % dims->world != 0)
        return set_error(SMH_E_DIM, "n (%d) must be a multiple of world (%d)", dims->n, dims->world);
    if ((int64_t)dims->n * 2 > (1 << 22)) return set_error(SMH_E_DIM, "2N too large (%d)", dims->n * 2);
    if (dims->flags & ~(SMH_DIMS_DENSE_WEIGHTS | SMH_DIMS_DENSE_BACKWARD | SMH_DIMS_Q16_TILES))
        return set_error(SMH_E_ARG, "unknown dims.flags 0x%x", dims->flags);
    if ((dims->flags & SMH_DIMS_DENSE_BACKWARD) && !(dims->flags & SMH_DIMS_DENSE_WEIGHTS))
        return set_error(SMH_E_ARG, "SMH_DIMS_DENSE_BACKWARD needs SMH_DIMS_DENSE_WEIGHTS");
    if ((dims->flags & SMH_DIMS_DENSE_WEIGHTS) && dims->world != 1)
        return set_error(SMH_E_DIM, "the materialised-weights path is single-rank (world == 1)");
    if ((dims->flags & SMH_DIMS_Q16_TILES) &&
        ((dims->flags & SMH_DIMS_DENSE_WEIGHTS) || dims->diff_type != SMH_DIFF_MPJPE || dims->weight_type != SMH_WEIGHT_LINEAR))
        return set_error(SMH_E_MODE, "SMH_DIMS_Q16_TILES: only with linear / mpjpe weights built from the joints");
    if (dims->diff_type < SMH_DIFF_MPJPE || dims->diff_type > SMH_DIFF_EUCLID)
        return set_error(SMH_E_MODE, "unknown diff_type %d", dims->diff_type);
    if (dims->weight_type != SMH_WEIGHT_LINEAR && dims->weight_type != SMH_WEIGHT_NONLINEAR)
        return set_error(SMH_E_MODE, "unknown weight_type %d", dims->weight_type);
    if (dims->weight_type == SMH_WEIGHT_NONLINEAR && !(dims->lambda_pos == dims->lambda_pos && dims->lambda_neg == dims->lambda_neg))
        return set_error(SMH_E_ARG, "non_linear weights: lambda is NaN");
    return 0;
}

// owner of row block I: rows I and Tp-1-I are paired so every rank stores ~the same number of tiles
static inline int row_owner(int I, int tp, int world) { return std::min(I, tp - 1 - I) % world; }

struct HostPlan {
    std::vector<int2> tiles;
    std::vector<int4> tasks;
    // the CTA ranges are cut separately for the two sweeps (a strip costs the backward sweep ~4 task times, the forward
    // sweep next to nothing): [0] backward, [1] forward
    std::vector<int2> strips[2];
    std::vector<int> cta_ptr[2];       // kNumCtas + 1 offsets into strips
};

// Contiguous ranges of the ordered task list for the kNumCtas sweep CTAs, cut for equal COST: a task costs 1, every strip a
// CTA opens (a run of tasks with the same row block; a cut opens one too) costs `strip_cost` task times.  With ~28 tasks per
// CTA (8 ranks) a strip more or less is 10 % of a CTA's time: equal task counts left the slowest CTA of the backward sweep
// 15-20 % behind the first (profiles/r02_strip_cost*.txt).
// cost of one task in the units of a plain direct task: a transposed task reads the staged tile with four 4-byte (or
// 2-byte) shared loads where a direct one takes one 16-byte load; a masked one (diagonal / ragged) runs the checked epilogue
static double g_cost_transposed = 0.0, g_cost_masked = 0.0;
static inline double task_cost(const int4 &t)
{
    return 1.0 + ((t.w & kTaskTransposed) ? g_cost_transposed : 0.0) + ((t.w & (kTaskDiagonal | kTaskRagged)) ? g_cost_masked : 0.0);
}

static void cut_ranges(const std::vector<int4> &tasks, double strip_cost, int strip_len, std::vector<int2> *strips_out,
                       std::vector<int> *cta_ptr_out)
{
    const int n_ctas = (int)std::min<size_t>(kNumCtas, std::max<size_t>(tasks.size(), 1));
    std::vector<int2> strips;
    std::vector<int> cta_ptr(kNumCtas + 1, 0);
    auto opens_strip = [&](size_t j, size_t first) {          // does task j open a new strip in a range starting at `first`?
        return j == first || tasks[j].x != tasks[j - 1].x;
    };
    double remaining = 0.0;
    for (size_t j = 0; j < tasks.size(); ++j) remaining += task_cost(tasks[j]) + (opens_strip(j, 0) ? strip_cost : 0.0);
    size_t lo = 0;
    for (int c = 0; c < kNumCtas; ++c) {
        cta_ptr[c] = (int)strips.size();
        if (c >= n_ctas || lo >= tasks.size()) continue;
        size_t hi = lo;
        if (c == n_ctas - 1) {
            hi = tasks.size();
        } else {
            // every later CTA pays for the strip its cut opens
            const double target = (remaining + strip_cost * (n_ctas - 1 - c)) / (double)(n_ctas - c);
            double acc = 0.0;
            const size_t must_leave = (size_t)(n_ctas - 1 - c);          // at least one task for every later CTA
            while (hi < tasks.size() - must_leave) {
                const double inc = task_cost(tasks[hi]) + (opens_strip(hi, lo) ? strip_cost : 0.0);
                if (hi > lo && acc + 0.5 * inc > target) break;
                acc += inc;
                ++hi;
            }
        }
        for (size_t j = lo; j < hi; ++j) remaining -= task_cost(tasks[j]) + (opens_strip(j, 0) ? strip_cost : 0.0);
        size_t i = lo;
        while (i < hi) {
            size_t j = i;
            while (j < hi && tasks[j].x == tasks[i].x && (int)(j - i) < strip_len) ++j;
            strips.push_back(make_int2((int)i, (int)j));
            i = j;
        }
        lo = hi;
    }
    cta_ptr[kNumCtas] = (int)strips.size();
    strips_out->swap(strips);
    cta_ptr_out->swap(cta_ptr);
}

static void enumerate_plan(const smh_dims_t &dims, int strip_len, HostPlan *out, int *n_stored, int *n_tasks,
                           int *n_strips, int *n_strips_fwd)
{
    const int m = 2 * dims.n;
    const int tp = (m + kTile - 1) / kTile;
    std::vector<int2> tiles;
    std::vector<int4> tasks;
    const bool dense = dims.flags & SMH_DIMS_DENSE_WEIGHTS;
    if (dense) {
        // materialised weights: every (I, J) tile is stored (tile id = I * tp + J) and visited directly by row block I;
        // the backward list adds the transposed visit of the mirror tile (J, I) for the column term of the gradient
        for (int I = 0; I < tp; ++I)
            for (int J = 0; J < tp; ++J) tiles.push_back(make_int2(I, J));
        for (int I = 0; I < tp; ++I) {
            for (int J = 0; J < tp; ++J) {
                for (int half = 0; half < 2; ++half) {
                    const int cj = 2 * J + half;
                    if (cj * kTaskN >= m) continue;
                    int flags = (I == J ? kTaskDiagonal : 0);
                    if (I * kTile + kTile > m || cj * kTaskN + kTaskN > m) flags |= kTaskRagged;
                    tasks.push_back(make_int4(I, cj, I * tp + J, flags));
                    if (dims.flags & SMH_DIMS_DENSE_BACKWARD)
                        tasks.push_back(make_int4(I, cj, J * tp + I, flags | kTaskTransposed));
                }
            }
        }
    }
    for (int I = 0; I < tp && !dense; ++I) {
        if (row_owner(I, tp, dims.world) != dims.rank) continue;
        for (int J = I; J < tp; ++J) {
            int lt = (int)tiles.size();
            tiles.push_back(make_int2(I, J));
            for (int half = 0; half < 2; ++half) {
                int cj = 2 * J + half;
                if (cj * kTaskN >= m) continue;
                int flags = (I == J ? kTaskDiagonal : 0);
                if (I * kTile + kTile > m || cj * kTaskN + kTaskN > m) flags |= kTaskRagged;
                tasks.push_back(make_int4(I, cj, lt, flags));
            }
            if (J > I) {
                for (int half = 0; half < 2; ++half) {
                    int cj = 2 * I + half;
                    int flags = kTaskTransposed;
                    if (J * kTile + kTile > m) flags |= kTaskRagged;
                    tasks.push_back(make_int4(J, cj, lt, flags));
                }
            }
        }
    }
    // Execution order of the tasks: by super-tile pair {A, B} (16 x 16 row blocks of 128 samples), (A, B) before
    // (B, A), then row block, then column.  The CTAs take equal contiguous ranges of this list, so the ~5 CTAs that
    // work inside one super-tile pair at the same time read every stored MPJPE tile twice (direct and transposed)
    // within a ~32 MB window: the second read hits the 126 MB L2 instead of HBM.  Inside a range, a strip is a
    // maximal run of tasks with the same row block (the accumulators are flushed at its end).  When all stored tiles
    // of the rank fit in L2 anyway (sharded runs), plain (row block, column) order gives the longest strips.
    {
        // super-block edge: a rank's share of one super-tile pair (2 sbk^2 / world tiles) stays ~32 MB
        const long long tile_bytes = (long long)kTileFloats * ((dims.flags & SMH_DIMS_Q16_TILES) ? 2 : 4);
        int sbk = 16;
        while ((long long)2 * (sbk + 1) * (sbk + 1) * tile_bytes <= (32ll << 20) * dims.world) ++sbk;
        if ((long long)tiles.size() * tile_bytes <= (96ll << 20)) sbk = 4096;
        if (dense) sbk = 4096;                 // every tile is read once per visit type: longest strips
        auto key = [&](const int4 &t) {
            long long a = t.x / sbk, b = (t.y / 2) / sbk;
            long long lo = a < b ? a : b, hi = a < b ? b : a;
            return (((lo * 4096 + hi) * 2 + (a > b ? 1 : 0)) * 4096 + t.x) * 8192 + t.y;
        };
        std::sort(tasks.begin(), tasks.end(), [&](const int4 &x, const int4 &y) { return key(x) < key(y); });
    }
    // SMH_STRIP_COST / SMH_STRIP_COST_FWD override the weights for experiments
    double cost_bwd = kStripCostBwd, cost_fwd = kStripCostFwd;
    if (const char *e = getenv("SMH_STRIP_COST")) cost_bwd = atof(e);
    if (const char *e = getenv("SMH_STRIP_COST_FWD")) cost_fwd = atof(e);
    g_cost_transposed = kCostTransposed;
    g_cost_masked = kCostMasked;
    if (const char *e = getenv("SMH_COST_TRANSPOSED")) g_cost_transposed = atof(e);
    if (const char *e = getenv("SMH_COST_MASKED")) g_cost_masked = atof(e);
    std::vector<int2> strips[2];
    std::vector<int> cta_ptr[2];
    cut_ranges(tasks, cost_bwd, strip_len, &strips[0], &cta_ptr[0]);
    cut_ranges(tasks, cost_fwd, strip_len, &strips[1], &cta_ptr[1]);
    *n_stored = (int)tiles.size();
    *n_tasks = (int)tasks.size();
    *n_strips = (int)strips[0].size();
    *n_strips_fwd = (int)strips[1].size();
    if (out) {
        out->tiles.swap(tiles);
        out->tasks.swap(tasks);
        for (int k = 0; k < 2; ++k) {
            out->strips[k].swap(strips[k]);
            out->cta_ptr[k].swap(cta_ptr[k]);
        }
    }
}

// per-thread memo of the last layout: the entry points are called several times per step with the same
// dims and enumerating the plan costs ~1 ms at 2N = 16384 (thread-local, so the library stays re-entrant)
struct LayoutMemo {
    bool valid = false;
    smh_dims_t dims;
    smh_layout_t lay;
};
static thread_local LayoutMemo g_memo;

static int compute_layout(const smh_dims_t &dims, smh_layout_t *lay, HostPlan *plan)
{
    if (!plan && g_memo.valid && memcmp(&g_memo.dims, &dims, sizeof(dims)) == 0) {
        *lay = g_memo.lay;
        return 0;
    }
    memset(lay, 0, sizeof(*lay));
    const int m = 2 * dims.n;
    const int tp = (m + kTile - 1) / kTile;
    const int64_t mp = (int64_t)tp * kTile;
    lay->m = m;
    lay->tiles_per_side = tp;
    lay->strip_len = dims.strip_len > 0 ? dims.strip_len : 4096;
    enumerate_plan(dims, lay->strip_len, plan, &lay->n_stored_tiles, &lay->n_tasks, &lay->n_strips, &lay->n_strips_fwd);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        int64_t o = off;
        off = align_up(off + bytes, 1024);
        return o;
    };
    lay->off_stats = take(sizeof(Stats));
    lay->off_neg = take(mp * 4);
    lay->off_rn = take(mp * 4);
    lay->off_rowloss = take(mp * 4);
    lay->off_dzacc = take(mp * kD * 4);
    lay->off_posd = take((int64_t)dims.n * 4);
    lay->off_zt = take(mp * kD * 4);
    lay->off_zb = take(mp * kD * 2);
    lay->off_zh = take(mp * kD * 2);
    lay->off_jp = take(mp * kJP * 4);
    // partial buffers of the peer exchange (unused, and not allocated, on a single rank)
    // + 64 B tail: one partial loss per rank (sharded finalize)
    // + 128 B tail: one partial loss per rank (sharded finalize) | one partial distance sum per rank (non_linear)
    lay->off_negparts = take(dims.world > 1 ? (int64_t)dims.world * mp * 4 + 128 : 0);
    lay->off_dzparts = take(dims.world > 1 ? (int64_t)m * kD * 4 : 0);
    lay->off_posinfo = take(dims.world > 1 ? (int64_t)4 * dims.n * 4 : 0);
    lay->off_dist = take((int64_t)lay->n_stored_tiles * kTileFloats * ((dims.flags & SMH_DIMS_Q16_TILES) ? 2 : 4));
    lay->ws_bytes = off;
    lay->plan_bytes = align_up((int64_t)sizeof(PlanHeader), 16) + align_up((int64_t)lay->n_stored_tiles * 8, 16) +
                      (int64_t)lay->n_tasks * 16 + align_up((int64_t)lay->n_strips * 8, 16) +
                      align_up((int64_t)lay->n_strips_fwd * 8, 16) + 2 * align_up((int64_t)(kNumCtas + 1) * 4, 16);
    g_memo.dims = dims;
    g_memo.lay = *lay;
    g_memo.valid = true;
    return 0;
}

static WsView carve(void *ws, const smh_layout_t &lay)
{
    char *b = (char *)ws;
    WsView v;
    v.stats = b + lay.off_stats;
    v.zt = (float *)(b + lay.off_zt);
    v.zb = (uint16_t *)(b + lay.off_zb);
    v.zh = (uint16_t *)(b + lay.off_zh);
    v.jp = (float *)(b + lay.off_jp);
    v.posd = (float *)(b + lay.off_posd);
    v.neg = (float *)(b + lay.off_neg);
    v.rn = (float *)(b + lay.off_rn);
    v.rowloss = (float *)(b + lay.off_rowloss);
    v.dzacc = (float *)(b + lay.off_dzacc);
    v.negparts = (float *)(b + lay.off_negparts);
    v.dzparts = (float *)(b + lay.off_dzparts);
    v.dist = (float *)(b + lay.off_dist);
    return v;
}

static PlanView carve_plan(const void *plan, const smh_layout_t &lay)
{
    const char *b = (const char *)plan;
    int64_t o = align_up((int64_t)sizeof(PlanHeader), 16);
    PlanView v;
    v.tiles = (const int2 *)(b + o);
    o += align_up((int64_t)lay.n_stored_tiles * 8, 16);
    v.tasks = (const int4 *)(b + o);
    o += (int64_t)lay.n_tasks * 16;
    v.strips = (const int2 *)(b + o);
    o += align_up((int64_t)lay.n_strips * 8, 16);
    v.cta_ptr = (const int *)(b + o);
    o += align_up((int64_t)(kNumCtas + 1) * 4, 16);
    v.strips_fwd = (const int2 *)(b + o);
    o += align_up((int64_t)lay.n_strips_fwd * 8, 16);
    v.cta_ptr_fwd = (const int *)(b + o);
    return v;
}

static int make_peers(const smh_dims_t &dims, const smh_layout_t &lay, void *ws_dev, const smh_exchange_t *exch,
                      Peers *out)
{
    memset(out, 0, sizeof(*out));
    out->off_stats = lay.off_stats;
    out->off_neg = lay.off_neg;
    out->off_dzacc = lay.off_dzacc;
    out->off_negparts = lay.off_negparts;
    out->off_dzparts = lay.off_dzparts;
    out->off_lossparts = lay.off_negparts + (int64_t)dims.world * lay.tiles_per_side * kTile * 4;
    out->off_posinfo = lay.off_posinfo;
    out->n = dims.n;
    if (!exch) {
        out->world = 1;
        out->rank = 0;
        out->ws[0] = (unsigned char *)ws_dev;
        return 0;
    }
    if (exch->world != dims.world || exch->rank != dims.rank || exch->world > SMH_MAX_PEERS)
        return set_error(SMH_E_DIM, "exchange world/rank %d/%d does not match dims %d/%d (max %d peers)", exch->world,
                         exch->rank, dims.world, dims.rank, SMH_MAX_PEERS);
    if (exch->ws_peer[exch->rank] != ws_dev)
        return set_error(SMH_E_ARG, "exchange ws_peer[rank] must be the local workspace");
    out->world = exch->world;
    out->rank = exch->rank;
    for (int p = 0; p < exch->world; ++p) {
        if (!exch->ws_peer[p]) return set_error(SMH_E_ARG, "exchange ws_peer[%d] is null", p);
        out->ws[p] = (unsigned char *)exch->ws_peer[p];
        if (!exch->signal_peer[p]) return set_error(SMH_E_ARG, "exchange signal_peer[%d] is null", p);
        out->sig[p] = (uint32_t *)exch->signal_peer[p];
    }
    out->fused = exch->fused ? 1 : 0;
    out->timeout_ms = exch->timeout_ms;
    return 0;
}

static int check_ptr(const void *p, const char *name, int align)
{
    if (!p) return set_error(SMH_E_ARG, "%s is null", name);
    if (((uintptr_t)p) % align) return set_error(SMH_E_ALIGN, "%s must be %d-byte aligned", name, align);
    return 0;
}

static int check_inputs(const smh_dims_t &dims, const smh_inputs_t *in)
{
    if (!in) return set_error(SMH_E_ARG, "inputs is null");
    int rc;
    if ((rc = check_ptr(in->z1_dev, "z1", 4))) return rc;
    if ((rc = check_ptr(in->z2_dev, "z2", 4))) return rc;
    if (!(dims.flags & SMH_DIMS_DENSE_WEIGHTS) || in->j1_dev || in->j2_dev) {      // joints are optional with materialised weights
        if ((rc = check_ptr(in->j1_dev, "joints1", 4))) return rc;
        if ((rc = check_ptr(in->j2_dev, "joints2", 4))) return rc;
    }
    if (in->z_row_stride < dims.d) return set_error(SMH_E_ARG, "z_row_stride < d");
    if (in->n_local <= 0 || dims.n % in->n_local != 0)
        return set_error(SMH_E_DIM, "n_local (%d) must divide n (%d)", in->n_local, dims.n);
    return 0;
}

// 0: linear weights from the stored distance tiles, 1: unit weights, 2: the tiles hold the materialised weights,
// 3: non_linear (sigmoid) weights from the stored distance tiles
static int weight_mode(const smh_dims_t &dims, int engine, bool backward, int *out)
{
    const bool dense_dims = dims.flags & SMH_DIMS_DENSE_WEIGHTS;
    if (engine & SMH_DENSE_WEIGHTS) {
        if (!dense_dims) return set_error(SMH_E_MODE, "SMH_DENSE_WEIGHTS needs dims.flags & SMH_DIMS_DENSE_WEIGHTS");
        if (backward != ((dims.flags & SMH_DIMS_DENSE_BACKWARD) != 0))
            return set_error(SMH_E_MODE, "materialised weights: smh_forward takes the plan without, smh_backward the "
                                         "plan with SMH_DIMS_DENSE_BACKWARD");
        *out = 2;
        return 0;
    }
    if (engine & SMH_UNIT_NEG_WEIGHTS) {
        *out = 1;
        return 0;
    }
    if (dense_dims) return set_error(SMH_E_MODE, "dense dims need SMH_DENSE_WEIGHTS or SMH_UNIT_NEG_WEIGHTS in engine");
    *out = dims.weight_type == SMH_WEIGHT_NONLINEAR ? 3 : 0;
    return 0;
}

}  // namespace smh

using namespace smh;

extern "C" {

int smh_version(void) { return SMH_VERSION; }

const char *smh_last_error(void) { return g_err; }

int smh_layout(const smh_dims_t *dims, smh_layout_t *out)
{
    int rc = validate_dims(dims);
    if (rc) return rc;
    if (!out) return set_error(SMH_E_ARG, "layout out is null");
    return compute_layout(*dims, out, nullptr);
}

int smh_plan_build(const smh_dims_t *dims, void *plan_host, int64_t plan_bytes)
{
    int rc = validate_dims(dims);
    if (rc) return rc;
    if (!plan_host) return set_error(SMH_E_ARG, "plan_host is null");
    smh_layout_t lay;
    HostPlan hp;
    compute_layout(*dims, &lay, &hp);
    if (plan_bytes < lay.plan_bytes)
        return set_error(SMH_E_SIZE, "plan buffer too small: %lld < %lld", (long long)plan_bytes,
                         (long long)lay.plan_bytes);
    char *b = (char *)plan_host;
    memset(b, 0, (size_t)lay.plan_bytes);
    PlanHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = kPlanMagic;
    h.m = (uint32_t)lay.m;
    h.world = (uint32_t)dims->world;
    h.rank = (uint32_t)dims->rank;
    h.tiles_per_side = (uint32_t)lay.tiles_per_side;
    h.n_stored = (uint32_t)lay.n_stored_tiles;
    h.n_tasks = (uint32_t)lay.n_tasks;
    h.n_strips = (uint32_t)lay.n_strips;
    h.strip_len = (uint32_t)lay.strip_len;
    int64_t o = align_up((int64_t)sizeof(PlanHeader), 16);
    h.off_tiles = (uint32_t)o;
    if (!hp.tiles.empty()) memcpy(b + o, hp.tiles.data(), hp.tiles.size() * 8);
    o += align_up((int64_t)lay.n_stored_tiles * 8, 16);
    h.off_tasks = (uint32_t)o;
    if (!hp.tasks.empty()) memcpy(b + o, hp.tasks.data(), hp.tasks.size() * 16);
    o += (int64_t)lay.n_tasks * 16;
    h.off_strips = (uint32_t)o;
    if (!hp.strips[0].empty()) memcpy(b + o, hp.strips[0].data(), hp.strips[0].size() * 8);
    o += align_up((int64_t)lay.n_strips * 8, 16);
    h.off_cta = (uint32_t)o;
    memcpy(b + o, hp.cta_ptr[0].data(), hp.cta_ptr[0].size() * 4);
    o += align_up((int64_t)(kNumCtas + 1) * 4, 16);
    h.off_strips_fwd = (uint32_t)o;
    h.n_strips_fwd = (uint32_t)lay.n_strips_fwd;
    if (!hp.strips[1].empty()) memcpy(b + o, hp.strips[1].data(), hp.strips[1].size() * 8);
    o += align_up((int64_t)lay.n_strips_fwd * 8, 16);
    h.off_cta_fwd = (uint32_t)o;
    memcpy(b + o, hp.cta_ptr[1].data(), hp.cta_ptr[1].size() * 4);
    memcpy(b, &h, sizeof(h));
    return 0;
}

#define SMH_COMMON_PROLOGUE(need_plan)                                                      \
    int rc = validate_dims(dims);                                                           \
    if (rc) return rc;                                                                      \
    if ((rc = check_ptr(ws_dev, "workspace", 256))) return rc;                             \
    if (need_plan && (rc = check_ptr(plan_dev, "plan", 16))) return rc;                     \
    smh_layout_t lay;                                                                       \
    compute_layout(*dims, &lay, nullptr);                                                   \
    WsView ws = carve(ws_dev, lay);                                                         \
    cudaStream_t st = (cudaStream_t)stream;

int smh_prep(const smh_dims_t *dims, const smh_inputs_t *in, void *ws_dev, int engine, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    if ((rc = check_inputs(*dims, in))) return rc;
    const bool zero = !(engine & SMH_PREP_NO_ZERO);
    engine &= ~SMH_PREP_NO_ZERO;
    if (engine != SMH_ENGINE_TC_TF32 && engine != SMH_ENGINE_FP32 && engine != SMH_ENGINE_TC_BF16 &&
        engine != SMH_ENGINE_TC_FP16)
        return set_error(SMH_E_MODE, "unknown engine %d", engine);
    if (zero) {
        cudaError_t e = cudaMemsetAsync(ws.stats, 0, (size_t)(lay.off_posd - lay.off_stats), st);
        if (e != cudaSuccess) return set_error((int)e, "prep memset: %s", cudaGetErrorString(e));
    }
    return launch_prep(*dims, lay, *in, ws, engine, st);      // dims carries diff_type
}

int smh_prep_zero(const smh_dims_t *dims, void *ws_dev, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    cudaError_t e = cudaMemsetAsync(ws.stats, 0, (size_t)(lay.off_posd - lay.off_stats), st);
    if (e != cudaSuccess) return set_error((int)e, "prep memset: %s", cudaGetErrorString(e));
    return 0;
}

int smh_import_weights(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, const float *neg_w_dev,
                       int64_t neg_row_stride, const float *pos_w_dev, void *stream)
{
    SMH_COMMON_PROLOGUE(true)
    if (!neg_w_dev && !pos_w_dev) return set_error(SMH_E_ARG, "import_weights: nothing to import");
    if (neg_w_dev) {
        if (!(dims->flags & SMH_DIMS_DENSE_WEIGHTS))
            return set_error(SMH_E_MODE, "import_weights: neg_w needs dims.flags & SMH_DIMS_DENSE_WEIGHTS");
        if (neg_row_stride < 2 * (int64_t)dims->n) return set_error(SMH_E_ARG, "neg_row_stride < 2N");
        if ((rc = check_ptr(neg_w_dev, "neg_w", 4))) return rc;
    }
    return launch_import_weights(*dims, lay, carve_plan(plan_dev, lay), ws, neg_w_dev, neg_row_stride, pos_w_dev, st);
}

int smh_mpjpe(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, const smh_exchange_t *exch, void *stream)
{
    SMH_COMMON_PROLOGUE(true)
    if (dims->flags & SMH_DIMS_DENSE_WEIGHTS)
        return set_error(SMH_E_MODE, "smh_mpjpe: the materialised-weights path takes smh_import_weights instead");
    Peers peers;
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    return launch_mpjpe(*dims, lay, carve_plan(plan_dev, lay), ws, peers, st);
}

int smh_forward(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, float temperature, int engine,
                const smh_exchange_t *exch, void *stream)
{
    SMH_COMMON_PROLOGUE(true)
    if (!(temperature > 0.f)) return set_error(SMH_E_ARG, "temperature must be positive");
    PlanView pv = carve_plan(plan_dev, lay);
    // the sweep accumulates locally (`peers` = local view); the partial sums travel with smh_exchange_neg, or, with the
    // fused exchange (`xp` = the real ranks), in the tail of the sweep itself
    Peers peers, xp;
    if ((rc = make_peers(*dims, lay, ws_dev, nullptr, &peers))) return rc;
    if ((rc = make_peers(*dims, lay, ws_dev, (exch && exch->fused) ? exch : nullptr, &xp))) return rc;
    int wmode;
    if ((rc = weight_mode(*dims, engine, false, &wmode))) return rc;
    engine &= 0xff;
    if (engine == SMH_ENGINE_TC_TF32 || engine == SMH_ENGINE_TC_BF16 || engine == SMH_ENGINE_TC_FP16)
        return launch_sweep_tc(false, engine == SMH_ENGINE_TC_TF32 ? 0 : (engine == SMH_ENGINE_TC_BF16 ? 1 : 2), wmode,
                               *dims, lay, pv, ws, peers, xp, temperature, st);
    if (xp.fused) return set_error(SMH_E_MODE, "the fused exchange runs the tensor-core engines only");
    if (engine == SMH_ENGINE_FP32 && (dims->flags & SMH_DIMS_Q16_TILES))
        return set_error(SMH_E_MODE, "the fp32 engine reads fp32 distance tiles (clear SMH_DIMS_Q16_TILES)");
    if (engine == SMH_ENGINE_FP32) return launch_sweep_fp32(false, wmode, *dims, lay, pv, ws, peers, temperature, st);
    return set_error(SMH_E_MODE, "unknown engine %d", engine);
}

int smh_backward(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, float temperature, int engine,
                 const smh_exchange_t *exch, void *stream)
{
    SMH_COMMON_PROLOGUE(true)
    if (!(temperature > 0.f)) return set_error(SMH_E_ARG, "temperature must be positive");
    PlanView pv = carve_plan(plan_dev, lay);
    Peers peers, xp;
    if ((rc = make_peers(*dims, lay, ws_dev, nullptr, &peers))) return rc;
    if ((rc = make_peers(*dims, lay, ws_dev, (exch && exch->fused) ? exch : nullptr, &xp))) return rc;
    // peer exchange: the row sums are the rank-ordered sum of the partials every rank delivered
    if (xp.fused) {
        // loss-only step: the row sums are reduced by a small kernel that also closes stage 4; otherwise the backward
        // sweep's CTAs reduce them in its head (one launch less)
        if ((engine & SMH_BACKWARD_RN_ONLY) && (rc = launch_rn_fused(lay, ws, xp, true, st))) return rc;
    } else if ((rc = launch_rn(lay, ws, exch ? dims->world : 1, st))) {
        return rc;
    }
    if (engine & SMH_BACKWARD_RN_ONLY) return 0;
    int wmode;
    if ((rc = weight_mode(*dims, engine, true, &wmode))) return rc;
    engine &= 0xff;
    if (engine == SMH_ENGINE_TC_TF32 || engine == SMH_ENGINE_TC_BF16 || engine == SMH_ENGINE_TC_FP16)
        return launch_sweep_tc(true, 1, wmode, *dims, lay, pv, ws, peers, xp, temperature, st);
    if (xp.fused) return set_error(SMH_E_MODE, "the fused exchange runs the tensor-core engines only");
    if (engine == SMH_ENGINE_FP32 && (dims->flags & SMH_DIMS_Q16_TILES))
        return set_error(SMH_E_MODE, "the fp32 engine reads fp32 distance tiles (clear SMH_DIMS_Q16_TILES)");
    if (engine == SMH_ENGINE_FP32) return launch_sweep_fp32(true, wmode, *dims, lay, pv, ws, peers, temperature, st);
    return set_error(SMH_E_MODE, "unknown engine %d", engine);
}

int smh_finalize(const smh_dims_t *dims, const smh_inputs_t *in, void *ws_dev, const float *dzacc_src_dev,
                 float temperature, float grad_scale, float *loss_dev, float *dz1_dev, float *dz2_dev,
                 int64_t dz_row_stride, int flags, const smh_exchange_t *exch, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    if ((rc = check_inputs(*dims, in))) return rc;
    if (!(temperature > 0.f)) return set_error(SMH_E_ARG, "temperature must be positive");
    if ((dz1_dev == nullptr) != (dz2_dev == nullptr)) return set_error(SMH_E_ARG, "dz1/dz2 must both be set or null");
    if (dz1_dev && dz_row_stride < dims->d) return set_error(SMH_E_ARG, "dz_row_stride < d");
    // NULL: all rows in the local accumulator (rank-major); otherwise this rank's own [2 n_local][128] block
    bool local_block = dzacc_src_dev != nullptr;
    int n_parts = 1;
    if (!dzacc_src_dev) dzacc_src_dev = ws.dzacc;
    Peers peers;
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    if (exch && exch->fused) {
        if (in->n_local * dims->world != dims->n) return set_error(SMH_E_DIM, "fused finalize takes the rank's local inputs");
        const int pm = (flags & SMH_UNIT_POS_WEIGHTS) ? 1 : (dims->weight_type == SMH_WEIGHT_NONLINEAR ? 3 : 0);
        return launch_finalize_fused(*dims, lay, *in, ws, pm, temperature, grad_scale, loss_dev, dz1_dev, dz2_dev,
                                     dz_row_stride, peers, st);
    }
    int phase = 0;
    if (flags & (SMH_FINALIZE_LOSS_PART | SMH_FINALIZE_GRAD)) {
        if (!exch) return set_error(SMH_E_ARG, "sharded finalize phases need the peer exchange");
        if ((flags & SMH_FINALIZE_LOSS_PART) && (flags & SMH_FINALIZE_GRAD))
            return set_error(SMH_E_ARG, "finalize: one phase per call");
        phase = (flags & SMH_FINALIZE_LOSS_PART) ? 1 : 2;
    }
    if (exch) {
        if (exch->world != dims->world) return set_error(SMH_E_DIM, "exchange world does not match dims");
        dzacc_src_dev = ws.dzparts;           // [world][2 n_local][128], summed in rank order
        local_block = true;
        n_parts = dims->world;
    }
    const int pos_mode = (flags & SMH_UNIT_POS_WEIGHTS) ? 1
                         : ((flags & SMH_DENSE_WEIGHTS) ? 2 : (dims->weight_type == SMH_WEIGHT_NONLINEAR ? 3 : 0));
    return launch_finalize(*dims, lay, *in, ws, dzacc_src_dev, local_block, n_parts, pos_mode,
                           temperature, grad_scale, loss_dev, dz1_dev, dz2_dev, dz_row_stride, phase, peers, st);
}

int smh_weights_dense(const smh_dims_t *dims, const void *plan_dev, void *ws_dev, float *pos_w_dev,
                      float *neg_w_dev, void *stream)
{
    SMH_COMMON_PROLOGUE(true)
    if (dims->world != 1) return set_error(SMH_E_DIM, "smh_weights_dense needs world == 1");
    if (dims->flags & SMH_DIMS_Q16_TILES) return set_error(SMH_E_MODE, "smh_weights_dense needs the fp32 tiles");
    if (!pos_w_dev && !neg_w_dev) return set_error(SMH_E_ARG, "no output requested");
    return launch_weights_dense(*dims, lay, carve_plan(plan_dev, lay), ws, pos_w_dev, neg_w_dev, st);
}

int smh_shard_prep(const smh_dims_t *dims, const smh_inputs_t *local_in, void *ws_dev, const smh_exchange_t *exch,
                   int engine, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    (void)ws;
    if (!exch || !exch->fused) return set_error(SMH_E_ARG, "shard_prep needs an exchange with fused = 1");
    if (!local_in || !local_in->z1_dev || !local_in->z2_dev || !local_in->j1_dev || !local_in->j2_dev)
        return set_error(SMH_E_ARG, "shard_prep: null input pointer");
    if (local_in->z_row_stride < dims->d) return set_error(SMH_E_ARG, "z_row_stride < d");
    if (dims->world < 2) return set_error(SMH_E_DIM, "shard_prep: world must be >= 2");
    const int eng = engine & ~SMH_SHARD_PREP_NO_IMAGES;
    if (eng != SMH_ENGINE_TC_TF32 && eng != SMH_ENGINE_TC_BF16 && eng != SMH_ENGINE_TC_FP16)
        return set_error(SMH_E_MODE, "the fused exchange runs the tensor-core engines only (got %d)", eng);
    Peers peers;
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    return launch_shard_prep(*dims, lay, *local_in, engine, peers, st);
}

int smh_shard_push_z(const smh_dims_t *dims, const smh_inputs_t *local_in, void *ws_dev, const smh_exchange_t *exch,
                     int engine, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    (void)ws;
    if (!exch || !exch->fused) return set_error(SMH_E_ARG, "shard_push_z needs an exchange with fused = 1");
    if (!local_in || !local_in->z1_dev || !local_in->z2_dev) return set_error(SMH_E_ARG, "shard_push_z: null input pointer");
    if (local_in->z_row_stride < dims->d) return set_error(SMH_E_ARG, "z_row_stride < d");
    if (engine != SMH_ENGINE_TC_TF32 && engine != SMH_ENGINE_TC_BF16 && engine != SMH_ENGINE_TC_FP16)
        return set_error(SMH_E_MODE, "the fused exchange runs the tensor-core engines only (got %d)", engine);
    Peers peers;
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    return launch_shard_push_z(*dims, lay, *local_in, engine, peers, st);
}

int smh_push_inputs(const smh_exchange_t *exch, const smh_inputs_t *local_in, int32_t n_local, int32_t d, void *stream)
{
    if (!exch || !local_in) return set_error(SMH_E_ARG, "push_inputs: null exchange/inputs");
    if (exch->world < 1 || exch->world > SMH_MAX_PEERS) return set_error(SMH_E_DIM, "bad exchange world %d", exch->world);
    if (n_local <= 0 || d <= 0 || d > SMH_MAX_DIM) return set_error(SMH_E_DIM, "push_inputs: bad n_local/d %d/%d", n_local, d);
    if (!local_in->z1_dev || !local_in->z2_dev || !local_in->j1_dev || !local_in->j2_dev)
        return set_error(SMH_E_ARG, "push_inputs: null input pointer");
    return launch_push_inputs(*exch, *local_in, n_local, d, (cudaStream_t)stream);
}

int smh_exchange_neg(const smh_dims_t *dims, void *ws_dev, const smh_exchange_t *exch, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    (void)ws;
    Peers peers;
    if (!exch) return set_error(SMH_E_ARG, "exchange_neg: null exchange");
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    return launch_exchange_neg(lay, peers, st);
}

int smh_exchange_dz(const smh_dims_t *dims, void *ws_dev, const smh_exchange_t *exch, void *stream)
{
    const void *plan_dev = nullptr;
    SMH_COMMON_PROLOGUE(false)
    (void)plan_dev;
    (void)ws;
    Peers peers;
    if (!exch) return set_error(SMH_E_ARG, "exchange_dz: null exchange");
    if ((rc = make_peers(*dims, lay, ws_dev, exch, &peers))) return rc;
    return launch_exchange_dz(*dims, lay, peers, st);
}

int smh_barrier(const smh_exchange_t *exch, void *stream)
{
    if (!exch) return set_error(SMH_E_ARG, "barrier: null exchange");
    if (exch->world < 1 || exch->world > SMH_MAX_PEERS) return set_error(SMH_E_DIM, "bad exchange world %d", exch->world);
    return launch_barrier(*exch, (cudaStream_t)stream);
}

int smh_scale_grads(const float *dz1_dev, const float *dz2_dev, const float *scale_dev, float *out1_dev, float *out2_dev,
                    int64_t count, void *stream)
{
    if (!dz1_dev || !dz2_dev || !scale_dev || !out1_dev || !out2_dev) return set_error(SMH_E_ARG, "scale_grads: null pointer");
    if (count <= 0) return set_error(SMH_E_ARG, "scale_grads: count must be positive");
    return launch_scale_pair(dz1_dev, dz2_dev, scale_dev, out1_dev, out2_dev, count, (cudaStream_t)stream);
}

int smh_l2norm_fwd(const float *x_dev, float *y_dev, float *norm_dev, int64_t rows, int32_t d, float eps,
                   void *stream)
{
    if (!x_dev || !y_dev) return set_error(SMH_E_ARG, "x/y is null");
    if (rows <= 0 || d <= 0) return set_error(SMH_E_ARG, "rows and d must be positive");
    return launch_l2norm_fwd(x_dev, y_dev, norm_dev, rows, d, eps, (cudaStream_t)stream);
}

int smh_l2norm_bwd(const float *y_dev, const float *norm_dev, const float *dy_dev, float *dx_dev, int64_t rows,
                   int32_t d, float eps, void *stream)
{
    if (!y_dev || !norm_dev || !dy_dev || !dx_dev) return set_error(SMH_E_ARG, "null pointer");
    if (rows <= 0 || d <= 0) return set_error(SMH_E_ARG, "rows and d must be positive");
    return launch_l2norm_bwd(y_dev, norm_dev, dy_dev, dx_dev, rows, d, eps, (cudaStream_t)stream);
}

int smh_transform_fwd(const float *x_dev, int64_t x_row_stride, const float *tx_dev, const float *ty_dev,
                      const float *angle_deg_dev, float *out_dev, int64_t out_row_stride, float *save_dev,
                      int64_t rows, int32_t d, float eps, void *stream)
{
    if (!x_dev || !out_dev || !save_dev) return set_error(SMH_E_ARG, "transform_fwd: x/out/save is null");
    if ((tx_dev == nullptr) != (ty_dev == nullptr)) return set_error(SMH_E_ARG, "transform_fwd: tx and ty go together");
    if (rows <= 0) return set_error(SMH_E_ARG, "rows must be positive");
    if (d <= 0 || d > SMH_MAX_DIM || (d & 1)) return set_error(SMH_E_DIM, "d must be even and in 2..%d (got %d)", SMH_MAX_DIM, d);
    if (x_row_stride < d || out_row_stride < d) return set_error(SMH_E_ARG, "row stride < d");
    if (((uintptr_t)save_dev) % 16) return set_error(SMH_E_ALIGN, "save must be 16-byte aligned");
    return launch_transform_fwd(x_dev, x_row_stride, tx_dev, ty_dev, angle_deg_dev, out_dev, out_row_stride, save_dev,
                                rows, d, eps, (cudaStream_t)stream);
}

int smh_transform_bwd(const float *x_dev, int64_t x_row_stride, const float *out_dev, int64_t out_row_stride,
                      const float *save_dev, const float *dout_dev, int64_t dout_row_stride, float *dx_dev,
                      int64_t dx_row_stride, int64_t rows, int32_t d, float eps, void *stream)
{
    if (!x_dev || !out_dev || !save_dev || !dout_dev || !dx_dev) return set_error(SMH_E_ARG, "transform_bwd: null pointer");
    if (rows <= 0) return set_error(SMH_E_ARG, "rows must be positive");
    if (d <= 0 || d > SMH_MAX_DIM || (d & 1)) return set_error(SMH_E_DIM, "d must be even and in 2..%d (got %d)", SMH_MAX_DIM, d);
    if (x_row_stride < d || out_row_stride < d || dout_row_stride < d || dx_row_stride < d)
        return set_error(SMH_E_ARG, "row stride < d");
    if (((uintptr_t)save_dev) % 16) return set_error(SMH_E_ALIGN, "save must be 16-byte aligned");
    return launch_transform_bwd(x_dev, x_row_stride, out_dev, out_row_stride, save_dev, dout_dev, dout_row_stride, dx_dev,
                                dx_row_stride, rows, d, eps, (cudaStream_t)stream);
}

static int check_head(const smh_head_t *h)
{
    if (!h) return set_error(SMH_E_ARG, "head is null");
    if (!h->x_dev || !h->w1_dev || !h->b1_dev || !h->gamma_dev || !h->beta_dev || !h->w2_dev || !h->h_dev || !h->colsum_dev ||
        !h->save_mean_dev || !h->save_rstd_dev || !h->y_dev || !h->norm_dev)
        return set_error(SMH_E_ARG, "head: null pointer");
    if ((h->running_mean_dev == nullptr) != (h->running_var_dev == nullptr))
        return set_error(SMH_E_ARG, "head: running_mean and running_var go together");
    if (h->x_row_stride < h->in_dim) return set_error(SMH_E_ARG, "head: x_row_stride < in_dim");
    return 0;
}

int smh_head_forward(const smh_head_t *head, void *stream)
{
    int rc = check_head(head);
    if (rc) return rc;
    return launch_head_fwd(*head, (cudaStream_t)stream);
}

int smh_head_backward(const smh_head_t *head, const smh_head_bwd_t *bwd, void *stream)
{
    int rc = check_head(head);
    if (rc) return rc;
    if (!bwd || !bwd->dy_dev || !bwd->w2t_dev || !bwd->dp_dev || !bwd->a_dev || !bwd->dhn_dev || !bwd->dh_dev || !bwd->colsum_dev ||
        !bwd->dgamma_dev || !bwd->dbeta_dev)
        return set_error(SMH_E_ARG, "head backward: null pointer");
    return launch_head_bwd(*head, *bwd, (cudaStream_t)stream);
}

int smh_selftest(int which, uint64_t *out_dev, int64_t out_words, void *stream)
{
    if (!out_dev || out_words < 8) return set_error(SMH_E_ARG, "out buffer needs >= 8 words");
    return launch_selftest(which, out_dev, out_words, (cudaStream_t)stream);
}

}  // extern "C"
