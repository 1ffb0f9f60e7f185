/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the weighted NT-Xent hot path.
 *
 * Plain-C restatement of the reference algorithm.  Nothing under simhand_b200/ may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Follows (paths relative to the reference checkout):
 *   src/models/utils.py:229-235   positive-pair MPJPE and its linear weight
 *   src/models/utils.py:237-259   all-pairs MPJPE matrix, global max/min, linear weight
 *   src/models/utils.py:407-426   weighted NT-Xent loss (mean reduction, tau = 0.5 default)
 *   autograd of :407-426          closed form derived in SURVEY.md section 7.2
 *
 * Bit recipe of the fp32 MPJPE as torch-CPU evaluates utils.py:252-253 (pinned against the
 * reference executed in the build container, see oracle/gen_golden.py and
 * tests/test_oracle.py):
 *   n_k = sqrtf(fmaf(dy, dy, dx*dx))                         k = 0..20
 *   s   = ((((n16 + n17) + n18) + n19) + n20)
 *   s  += (n_k + n_{k+8})                                    k = 0..7, in order
 *   D   = s / 21.0f
 *   W   = (Dmax - D) / (Dmax - Dmin)                         each op rounded to fp32
 * The contrastive part is evaluated in double from the fp32 weights, so it is the
 * "true" value the fp32 reference approximates (its own fp32 noise is ~5e-8 relative).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; fmaf() must stay a fused op).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SMH_J 21

/* one MPJPE entry; a and b point at 42 floats (x0,y0,x1,y1,...) */
static inline float mpjpe_pair(const float *a, const float *b)
{
    float n[SMH_J];
    for (int k = 0; k < SMH_J; ++k) {
        float dx = a[2 * k] - b[2 * k];
        float dy = a[2 * k + 1] - b[2 * k + 1];
        float xx = dx * dx;
        n[k] = sqrtf(fmaf(dy, dy, xx));
    }
    float s = n[16];
    s = s + n[17];
    s = s + n[18];
    s = s + n[19];
    s = s + n[20];
    for (int k = 0; k < 8; ++k) {
        float t = n[k] + n[k + 8];
        s = s + t;
    }
    return s / 21.0f;
}

/* D[r - r0][c] for r in [r0, r1), c in [0, M); J is [M][42] contiguous. */
int smh_oracle_mpjpe_rows(const float *J, int M, int r0, int r1, float *D)
{
    if (!J || !D || M <= 0 || r0 < 0 || r1 > M || r0 > r1) return -1;
#pragma omp parallel for schedule(static)
    for (int r = r0; r < r1; ++r) {
        const float *a = J + (size_t)r * 42;
        float *out = D + (size_t)(r - r0) * M;
        for (int c = 0; c < M; ++c) out[c] = mpjpe_pair(a, J + (size_t)c * 42);
    }
    return 0;
}

/* Global stats of utils.py:255-256 without storing D: returns max and min over all M*M entries. */
int smh_oracle_mpjpe_minmax(const float *J, int M, float *dmax, float *dmin)
{
    if (!J || M <= 0) return -1;
    float gmax = -INFINITY, gmin = INFINITY;
    int has_nan = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(max : gmax) reduction(min : gmin) reduction(| : has_nan)
    for (int r = 0; r < M; ++r) {
        const float *a = J + (size_t)r * 42;
        for (int c = 0; c < M; ++c) {
            float v = mpjpe_pair(a, J + (size_t)c * 42);
            if (v != v) has_nan = 1;
            if (v > gmax) gmax = v;
            if (v < gmin) gmin = v;
        }
    }
    if (has_nan) gmax = gmin = NAN;
    *dmax = gmax;
    *dmin = gmin;
    return 0;
}

/* Positive-pair weights, utils.py:229-235.  J is the concatenated [2N][42]. */
int smh_oracle_pos_weights(const float *J, int N, float *pos_w, float *pmax_out, float *pmin_out)
{
    if (!J || !pos_w || N <= 0) return -1;
    float pmax = -INFINITY, pmin = INFINITY;
    for (int k = 0; k < N; ++k) {
        float v = mpjpe_pair(J + (size_t)k * 42, J + (size_t)(k + N) * 42);
        pos_w[k] = v;
        if (v > pmax) pmax = v;
        if (v < pmin) pmin = v;
    }
    float den = pmax - pmin;
    for (int k = 0; k < N; ++k) pos_w[k] = (pmax - pos_w[k]) / den;
    if (pmax_out) *pmax_out = pmax;
    if (pmin_out) *pmin_out = pmin;
    return 0;
}

/* Dense negative weights for rows [r0, r1): W[r - r0][c], utils.py:259. */
int smh_oracle_neg_weights_rows(const float *J, int M, int r0, int r1, float dmax, float dmin, float *W)
{
    int rc = smh_oracle_mpjpe_rows(J, M, r0, r1, W);
    if (rc) return rc;
    float den = dmax - dmin;
    size_t n = (size_t)(r1 - r0) * M;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) W[i] = (dmax - W[i]) / den;
    return 0;
}

/*
 * One full step in the reference's semantics: loss (utils.py:407-426) and its gradient
 * w.r.t. z = [z1; z2], double arithmetic on top of the fp32 weights.
 *   z    [M][d] fp32 (rows already L2-normalised by the caller, M = 2N)
 *   J    [M][42] fp32
 * outputs: loss (1), dz [M][d] double, neg [M] double (off-diagonal row sums), pos_w [N] fp32,
 *          stats[4] = {Dmax, Dmin, Pmax, Pmin}.
 */
int smh_oracle_step(const float *z, const float *J, int N, int d, double tau,
                    double *loss_out, double *dz, double *neg_out, float *pos_w, float *stats)
{
    if (!z || !J || N <= 0 || d <= 0 || !loss_out) return -1;
    const int M = 2 * N;
    float dmax, dmin, pmax, pmin;
    float *pw = pos_w ? pos_w : (float *)malloc(sizeof(float) * (size_t)N);
    double *neg = neg_out ? neg_out : (double *)malloc(sizeof(double) * (size_t)M);
    if (!pw || !neg) return -2;
    smh_oracle_mpjpe_minmax(J, M, &dmax, &dmin);
    smh_oracle_pos_weights(J, N, pw, &pmax, &pmin);
    if (stats) { stats[0] = dmax; stats[1] = dmin; stats[2] = pmax; stats[3] = pmin; }
    const float den = dmax - dmin;

    /* pass A: neg_i = sum_{j != i} exp(S_ij * W_ij / tau) */
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < M; ++i) {
        const float *zi = z + (size_t)i * d, *ji = J + (size_t)i * 42;
        double acc = 0.0;
        for (int j = 0; j < M; ++j) {
            if (j == i) continue;
            const float *zj = z + (size_t)j * d;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += (double)zi[k] * (double)zj[k];
            float w = (dmax - mpjpe_pair(ji, J + (size_t)j * 42)) / den;
            acc += exp(s * (double)w / tau);
        }
        neg[i] = acc;
    }
    /* loss = mean_i [ log neg_i - S_{i,p(i)} * Wp / tau ] */
    double loss = 0.0;
    for (int i = 0; i < M; ++i) {
        int p = i < N ? i + N : i - N;
        const float *zi = z + (size_t)i * d, *zp = z + (size_t)p * d;
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += (double)zi[k] * (double)zp[k];
        loss += log(neg[i]) - s * (double)pw[i % N] / tau;
    }
    *loss_out = loss / M;

    if (dz) {
        /* dz_i = 1/(M tau) sum_{j != i} W_ij E_ij (1/neg_i + 1/neg_j) z_j - 2/(M tau) Wp z_p(i) */
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < M; ++i) {
            const float *zi = z + (size_t)i * d, *ji = J + (size_t)i * 42;
            double *g = dz + (size_t)i * d;
            for (int k = 0; k < d; ++k) g[k] = 0.0;
            for (int j = 0; j < M; ++j) {
                if (j == i) continue;
                const float *zj = z + (size_t)j * d;
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += (double)zi[k] * (double)zj[k];
                float w = (dmax - mpjpe_pair(ji, J + (size_t)j * 42)) / den;
                double a = (double)w * exp(s * (double)w / tau) * (1.0 / neg[i] + 1.0 / neg[j]);
                for (int k = 0; k < d; ++k) g[k] += a * (double)zj[k];
            }
            int p = i < N ? i + N : i - N;
            const float *zp = z + (size_t)p * d;
            double wp = (double)pw[i % N];
            for (int k = 0; k < d; ++k) g[k] = (g[k] - 2.0 * wp * (double)zp[k]) / ((double)M * tau);
        }
    }
    if (!pos_w) free(pw);
    if (!neg_out) free(neg);
    return 0;
}

/* Bounded sample used by bench.py's CPU legs: rows [r0, r1) of the three sweeps of one step
 * (max pre-pass, forward row sums, backward rows); returns a checksum so the work is not elided. */
double smh_oracle_step_rows(const float *z, const float *J, int N, int d, double tau, int r0, int r1)
{
    const int M = 2 * N;
    double chk = 0.0;
    float dmax = 0.f;
    /* sweep 1: max over the sampled rows */
#pragma omp parallel for schedule(dynamic, 8) reduction(max : dmax)
    for (int i = r0; i < r1; ++i)
        for (int j = 0; j < M; ++j) {
            float v = mpjpe_pair(J + (size_t)i * 42, J + (size_t)j * 42);
            if (v > dmax) dmax = v;
        }
    if (!(dmax > 0.f)) dmax = 1.f;
    /* sweeps 2+3: forward sums and backward rows (neg_j approximated by neg_i: same op count) */
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : chk)
    for (int i = r0; i < r1; ++i) {
        const float *zi = z + (size_t)i * d, *ji = J + (size_t)i * 42;
        double acc = 0.0;
        for (int j = 0; j < M; ++j) {
            if (j == i) continue;
            const float *zj = z + (size_t)j * d;
            float s = 0.f;
            for (int k = 0; k < d; ++k) s += zi[k] * zj[k];
            float w = (dmax - mpjpe_pair(ji, J + (size_t)j * 42)) / dmax;
            acc += exp((double)(s * w) / tau);
        }
        double g[1024];
        int dd = d < 1024 ? d : 1024;
        for (int k = 0; k < dd; ++k) g[k] = 0.0;
        for (int j = 0; j < M; ++j) {
            if (j == i) continue;
            const float *zj = z + (size_t)j * d;
            float s = 0.f;
            for (int k = 0; k < d; ++k) s += zi[k] * zj[k];
            float w = (dmax - mpjpe_pair(ji, J + (size_t)j * 42)) / dmax;
            double a = (double)w * exp((double)(s * w) / tau) * (2.0 / acc);
            for (int k = 0; k < dd; ++k) g[k] += a * (double)zj[k];
        }
        chk += log(acc) + g[0] + g[dd - 1];
    }
    return chk;
}

int smh_oracle_version(void) { return 1; }
