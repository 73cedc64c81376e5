/*
 * glabc_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar CPU restatement of the reference's sampler inner loops (caofff/GL-ABC-MCMC,
 * `glabcmcmc` 1.0.1), written from the reference's Python source and pinned against golden
 * vectors produced by running that Python in the build container
 * (tests/golden/make_golden.py -> tests/golden/ npz files; tests/test_oracle_golden.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (gl-abc-mcmc_b200/) never does.
 *
 * Everything is float32 in the reference's operation order (compile with -ffp-contract=off),
 * float64 only where the reference is (resampling compare, GLMALA gradient statistics).
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * It shares only the POD parameter structs of include/glabc.h with the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/glabc.h"

#include <pthread.h>
#include <unistd.h>

#define ORACLE_EXPORT __attribute__((visibility("default")))

/* ---- chains are independent: a plain pthread fan-out over contiguous chain ranges ---------- */
static int g_threads = 0; /* 0 = all online cores */

typedef void (*range_fn)(void* ctx, int64_t c0, int64_t c1);
typedef struct { range_fn fn; void* ctx; int64_t c0, c1; } range_job;

static void* range_thunk(void* p)
{
    range_job* j = (range_job*)p;
    j->fn(j->ctx, j->c0, j->c1);
    return NULL;
}

static int effective_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_chains(range_fn fn, void* ctx, int64_t C)
{
    int nt = effective_threads();
    if (nt > 256) nt = 256;
    if ((int64_t)nt > C) nt = (int)(C > 0 ? C : 1);
    if (nt <= 1) { fn(ctx, 0, C); return; }
    pthread_t th[256];
    range_job jobs[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].c0 = C * t / nt; jobs[t].c1 = C * (t + 1) / nt;
        if (pthread_create(&th[t], NULL, range_thunk, &jobs[t]) != 0) { th[t] = 0; fn(ctx, jobs[t].c0, jobs[t].c1); }
    }
    for (int t = 0; t < nt; ++t) if (th[t]) pthread_join(th[t], NULL);
}

/* -------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11) — the counter-based generator the native-RNG mode of the product uses.  The reference
 * has no counterpart (it uses torch's global MT19937, GlobalMCMC.py:39); this is the shared
 * native-mode RNG spec of DESIGN.md so the CPU baseline and the kernels draw the same streams.
 * ------------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORACLE_EXPORT void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* native-mode stream layout (DESIGN.md "RNG streams"): counter = (chain_lo, chain_hi, block, slot).
 * slot 1 = the step block: one Philox call carries the 4 normals, U_b (16 bits) and U_a (24 bits) of a
 * step; slot 2+g = extra normal blocks; slot 0x80000000 = the float64 resampling uniform.           */
enum { SLOT_STEP = 1, SLOT_NORMAL = 2 /* + group index */ };

static inline void philox_block(uint64_t seed, uint64_t chain, uint32_t block, uint32_t slot, uint32_t out[4])
{
    const uint32_t ctr[4] = {(uint32_t)chain, (uint32_t)(chain >> 32), block, slot};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    oracle_philox4x32_10(ctr, key, out);
}

/* Box-Muller pair from two words: radius uniform from w0[31:8] (24 bits), angle from w1[31:12] (20) */
static inline void box_muller(uint32_t w0, uint32_t w1, float* n0, float* n1)
{
    const float u1 = fmaf((float)(w0 & 0xFFFFFF00u), 0x1p-32f, 0x1p-25f);
    const float r = sqrtf(-2.0f * logf(u1));
    const float a = (float)(w1 & 0xFFFFF000u) * (6.28318530717958647692f * 0x1p-32f);
    *n0 = r * cosf(a);
    *n1 = r * sinf(a);
}

/* the draws of step i: n normals (first four from the step block), U_b on a 16-bit grid, U_a on
 * torch.rand's 24-bit grid (B-16) */
static void native_step_draws(uint64_t seed, uint64_t chain, uint32_t step, int n, float* normals, float* u_b, float* u_a)
{
    uint32_t w[4];
    float z[4];
    philox_block(seed, chain, step, SLOT_STEP, w);
    box_muller(w[0], w[1], &z[0], &z[1]);
    box_muller(w[2], w[3], &z[2], &z[3]);
    for (int j = 0; j < 4 && j < n; ++j) normals[j] = z[j];
    *u_b = (float)(((w[0] & 0xFFu) << 8) | (w[2] & 0xFFu)) * 0x1p-16f;
    *u_a = (float)((w[1] & 0xFFFu) | ((w[3] << 12) & 0xFFF000u)) * 0x1p-24f;
    for (int g = 1; g * 4 < n; ++g) {
        uint32_t v[4];
        philox_block(seed, chain, step, SLOT_NORMAL + (uint32_t)(g - 1), v);
        box_muller(v[0], v[1], &z[0], &z[1]);
        box_muller(v[2], v[3], &z[2], &z[3]);
        for (int j = 0; j < 4 && g * 4 + j < n; ++j) normals[g * 4 + j] = z[j];
    }
}

/* n normals from the extra blocks slot0, slot0+1, ... of (chain, step) */
static void native_normals(uint64_t seed, uint64_t chain, uint32_t step, uint32_t slot0, int n, float* out)
{
    for (int g = 0; g * 4 < n; ++g) {
        uint32_t w[4];
        float z[4];
        philox_block(seed, chain, step, slot0 + (uint32_t)g, w);
        box_muller(w[0], w[1], &z[0], &z[1]);
        box_muller(w[2], w[3], &z[2], &z[3]);
        for (int j = 0; j < 4 && g * 4 + j < n; ++j) out[g * 4 + j] = z[j];
    }
}

/* -------------------------------------------------------------------------------------------
 * Distributions — distribution.py
 * ------------------------------------------------------------------------------------------- */
/* -0.5*d*log(2*pi) is a float64 scalar that torch casts to float32 before the subtraction
 * (verified in the container: (c - x) == float32(c) - x for 1e6 random x).                      */
static inline float half_log_2pi_f32(int d) { return (float)(-0.5 * (double)d * log(2.0 * M_PI)); }

/* torch.sum over <16 contiguous float32: ATen SumKernel.cpp row_sum with ilp_factor 4 — four
 * interleaved partial sums, the tail folded into partial 0, then partials 1..3 folded in.
 * Reproduces the tree SURVEY.md B-3 measured for 6 elements: ((((v0+v4)+v5)+v1)+v2)+v3.          */
static float torch_sum_f32(const float* v, int n)
{
    if (n >= 16) { /* vectorised path (16 lanes, AVX-512): lanes, then tail, then lanes folded */
        float lane[16];
        const int nv = n / 16;
        for (int l = 0; l < 16; ++l) lane[l] = 0.0f;
        for (int r = 0; r < nv; ++r)
            for (int l = 0; l < 16; ++l) lane[l] += v[r * 16 + l];
        float acc = 0.0f;
        for (int i = nv * 16; i < n; ++i) acc += v[i];
        for (int l = 0; l < 16; ++l) acc += lane[l];
        return acc;
    }
    float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int rows = n / 4;
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < 4; ++k) p[k] += v[r * 4 + k];
    for (int i = rows * 4; i < n; ++i) p[0] += v[i];
    for (int k = 1; k < 4; ++k) p[0] += p[k];
    return p[0];
}

/* DiagGaussian.log_prob, distribution.py:176-181 */
static float diag_gauss_log_prob(const float* z, const float* loc, const float* log_scale,
                                 const float* scale, int d)
{
    float t[GLABC_MAX_DIM];
    for (int i = 0; i < d; ++i) {
        const float r = (z[i] - loc[i]) / scale[i];
        t[i] = log_scale[i] + 0.5f * (r * r);
    }
    return half_log_2pi_f32(d) - torch_sum_f32(t, d);
}

/* DiagGaussian.forward, distribution.py:166-174: z = loc + exp(log_scale)*eps, log_p from eps */
static float diag_gauss_forward(const float* eps, const float* loc, const float* log_scale,
                                const float* scale, int d, float* z)
{
    float t[GLABC_MAX_DIM];
    for (int i = 0; i < d; ++i) {
        z[i] = loc[i] + scale[i] * eps[i];
        t[i] = log_scale[i] + 0.5f * (eps[i] * eps[i]);
    }
    return half_log_2pi_f32(d) - torch_sum_f32(t, d);
}

/* -------------------------------------------------------------------------------------------
 * ABC model plugin — examples/Mixture.py
 * ------------------------------------------------------------------------------------------- */
/* generate_samples, Mixture.py:13-26: |theta| + likelihood.sample() */
static void model_simulate(const glabc_model_t* m, const float* theta, const float* eps, float* y)
{
    for (int i = 0; i < m->y_dim; ++i) {
        const float noise = m->noise_loc[i] + m->noise_scale[i] * eps[i];
        const float mean = (m->family == GLABC_MODEL_ABS_NORMAL) ? fabsf(theta[i]) : theta[i];
        y[i] = mean + noise;
    }
}

/* prior_log_prob, Mixture.py:28-31 */
static float model_prior(const glabc_model_t* m, const float* theta)
{
    return diag_gauss_log_prob(theta, m->prior_loc, m->prior_log_scale, m->prior_scale, m->theta_dim);
}

/* discrepancy, Mixture.py:33-36 */
static float model_discrepancy(const glabc_model_t* m, const float* y)
{
    float t[GLABC_MAX_DIM];
    for (int i = 0; i < m->y_dim; ++i) {
        const float dy = y[i] - m->y_obs[i];
        t[i] = dy * dy;
    }
    return sqrtf(torch_sum_f32(t, m->y_dim));
}

/* calculate_log_kernel_dis, Mixture.py:47-53: DiagGaussian(1, 0, log eps).log_prob(dis) */
static float model_log_kernel_dis(const glabc_model_t* m, float dis)
{
    const float r = (dis - 0.0f) / m->eps_scale;
    return half_log_2pi_f32(1) - (m->eps_log_scale + 0.5f * (r * r));
}

/* calculate_log_kernel, Mixture.py:38-45 */
static float model_log_kernel(const glabc_model_t* m, const float* y)
{
    return model_log_kernel_dis(m, model_discrepancy(m, y));
}

/* -------------------------------------------------------------------------------------------
 * shared bookkeeping
 * ------------------------------------------------------------------------------------------- */
static inline size_t trace_index(const glabc_run_t* r, int64_t chain, int64_t row, int d)
{
    const int64_t c = r->trace_chain_off + chain;
    if (r->trace_layout == GLABC_TRACE_CHAIN_MAJOR) return (size_t)((c * r->trace_rows + row) * d);
    return (size_t)((row * r->trace_chains + c) * d);
}

static void stats_update(float* st, int d, const float* theta_new, const float* theta_prev)
{
    st[GLABC_STAT_STEPS] += 1.0f;
    int tri = 0;
    for (int i = 0; i < d; ++i) {
        st[GLABC_STAT_SUM + i] += theta_new[i];
        st[GLABC_STAT_SUM + d + i] += theta_new[i] * theta_new[i];
        for (int j = i; j < d; ++j, ++tri)
            st[GLABC_STAT_SUM + 2 * d + tri] += (theta_new[i] - theta_prev[i]) * (theta_new[j] - theta_prev[j]);
    }
}

static int check_common(const glabc_model_t* m, const glabc_run_t* r)
{
    if (!m || !r) return GLABC_ERR_INVALID;
    if (m->theta_dim < 1 || m->theta_dim > GLABC_MAX_DIM || m->y_dim != m->theta_dim) return GLABC_ERR_UNSUPPORTED;
    if (m->family != GLABC_MODEL_ABS_NORMAL && m->family != GLABC_MODEL_ID_NORMAL) return GLABC_ERR_UNSUPPORTED;
    if (r->n_chains < 0 || r->n_steps < 0 || !r->theta || !r->y) return GLABC_ERR_INVALID;
    if (r->trace_layout != GLABC_TRACE_NONE && !r->trace) return GLABC_ERR_INVALID;
    if (r->rng_mode == GLABC_RNG_REPLAY && !r->tape32) return GLABC_ERR_INVALID;
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * GlobalMCMC — GlobalMCMC.py:37-68
 * Draw order per step: U_b, N[1,d] (proposal), N[1,y_dim] (simulator), U_a.
 * ------------------------------------------------------------------------------------------- */
typedef struct { const glabc_model_t* m; const glabc_dist_t* lp; const glabc_dist_t* gp; const glabc_run_t* r; } sampler_job;

static void run_global_range(void* vctx, int64_t c_begin, int64_t c_end)
{
    const sampler_job* job = (const sampler_job*)vctx;
    const glabc_model_t* m = job->m;
    const glabc_dist_t* lp = job->lp;
    const glabc_dist_t* gp = job->gp;
    const glabc_run_t* r = job->r;
    const int d = m->theta_dim, yd = m->y_dim;
    const int slots = GLABC_TAPE_GLOBAL_SLOTS(d, yd);
    const int ns = GLABC_NSTATS(d);
    const float gf = r->global_frequency;
    const int64_t C = r->n_chains;

    for (int64_t c = c_begin; c < c_end; ++c) {
        float theta[GLABC_MAX_DIM], y[GLABC_MAX_DIM], st[GLABC_NSTATS(GLABC_MAX_DIM)];
        memcpy(theta, r->theta + c * d, sizeof(float) * d);
        memcpy(y, r->y + c * yd, sizeof(float) * yd);
        memset(st, 0, sizeof(st));
        if (r->write_row0 && r->trace_layout != GLABC_TRACE_NONE)
            memcpy(r->trace + trace_index(r, c, r->step_base, d), theta, sizeof(float) * d);

        for (int64_t s = 0; s < r->n_steps; ++s) {
            const int64_t i = r->step_base + 1 + s; /* the reference's loop variable, GlobalMCMC.py:37 */
            float u_b, u_a, eps_p[GLABC_MAX_DIM], eps_s[GLABC_MAX_DIM];
            if (r->rng_mode == GLABC_RNG_REPLAY) {
                const float* t = r->tape32 + (size_t)s * slots * C + c;
                u_b = t[0];
                for (int k = 0; k < d; ++k) eps_p[k] = t[(size_t)(1 + k) * C];
                for (int k = 0; k < yd; ++k) eps_s[k] = t[(size_t)(1 + d + k) * C];
                u_a = t[(size_t)(1 + d + yd) * C];
            } else {
                float z[2 * GLABC_MAX_DIM + 4];
                native_step_draws(r->seed, (uint64_t)(r->chain_id_base + c), (uint32_t)i, d + yd, z, &u_b, &u_a);
                memcpy(eps_p, z, sizeof(float) * d);
                memcpy(eps_s, z + d, sizeof(float) * yd);
            }

            const int is_global = u_b < gf; /* GlobalMCMC.py:39, float32 compare (B-15) */
            float theta_p[GLABC_MAX_DIM], y_p[GLABC_MAX_DIM], log_acc, prior_p, kern_p;
            if (is_global) {
                /* GlobalMCMC.py:40-46 */
                const float lq_p = diag_gauss_forward(eps_p, gp->a, gp->b, gp->c, d, theta_p);
                model_simulate(m, theta_p, eps_s, y_p);
                prior_p = model_prior(m, theta_p);
                kern_p = model_log_kernel(m, y_p);
                log_acc = prior_p + kern_p;
                log_acc = log_acc + diag_gauss_log_prob(theta, gp->a, gp->b, gp->c, d);
                log_acc = log_acc - lq_p;
                log_acc = log_acc - model_prior(m, theta);
                log_acc = log_acc - model_log_kernel(m, y);
            } else {
                /* GlobalMCMC.py:56-61: Local_Proposal.sample(1) + Theta_old */
                float z[GLABC_MAX_DIM];
                (void)diag_gauss_forward(eps_p, lp->a, lp->b, lp->c, d, z);
                for (int k = 0; k < d; ++k) theta_p[k] = z[k] + theta[k];
                model_simulate(m, theta_p, eps_s, y_p);
                prior_p = model_prior(m, theta_p);
                kern_p = model_log_kernel(m, y_p);
                log_acc = prior_p + kern_p;
                log_acc = log_acc - model_prior(m, theta);
                log_acc = log_acc - model_log_kernel(m, y);
            }
            const float log_w = logf(u_a);     /* GlobalMCMC.py:47,62 */
            const int accept = log_w < log_acc; /* strict; NaN rejects (B-17) */

            float prev[GLABC_MAX_DIM];
            memcpy(prev, theta, sizeof(float) * d);
            if (accept) {
                memcpy(theta, theta_p, sizeof(float) * d);
                memcpy(y, y_p, sizeof(float) * yd);
            }
            stats_update(st, d, theta, prev);
            st[GLABC_STAT_GLOBAL_STEPS] += (float)is_global;
            st[is_global ? GLABC_STAT_ACC_GLOBAL : GLABC_STAT_ACC_LOCAL] += (float)accept;
            if (r->trace_layout != GLABC_TRACE_NONE)
                memcpy(r->trace + trace_index(r, c, i, d), theta, sizeof(float) * d);
            if (r->debug) {
                float* g = r->debug + (size_t)s * GLABC_DEBUG_SLOTS * C + c;
                g[0] = (float)(is_global | (accept << 1));
                g[(size_t)1 * C] = prior_p;
                g[(size_t)2 * C] = kern_p;
                g[(size_t)3 * C] = log_acc;
            }
        }
        memcpy(r->theta + c * d, theta, sizeof(float) * d);
        memcpy(r->y + c * yd, y, sizeof(float) * yd);
        if (r->stats)
            for (int k = 0; k < ns; ++k) r->stats[c * ns + k] += st[k];
    }
}

ORACLE_EXPORT int oracle_run_global(const glabc_model_t* m, const glabc_dist_t* lp, const glabc_dist_t* gp,
                                    const glabc_run_t* r)
{
    int rc = check_common(m, r);
    if (rc) return rc;
    if (!lp || !gp || lp->kind != GLABC_DIST_DIAG_GAUSSIAN || gp->kind != GLABC_DIST_DIAG_GAUSSIAN)
        return GLABC_ERR_UNSUPPORTED;
    sampler_job job = {m, lp, gp, r};
    parallel_chains(run_global_range, &job, r->n_chains);
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * GLMCMC — GLMCMC.py:58-104; weight_sampling GLMCMC.py:7-22
 * Global draw order: U_b, N[K,d] (proposal), N[K,y_dim] (simulator), U64 (numpy float64).
 * Local  draw order: U_b, N[1,d], N[1,y_dim], U_a.
 * The prior-sentinel redraw loop (GLMCMC.py:92-93) cannot fire for the fused priors: it needs
 * prior_log_prob == 7*log(1e-10) exactly, and a DiagGaussian prior only hits that value by
 * coincidence of rounding; it is omitted here and in the kernels (DESIGN.md "omitted branches").
 * ------------------------------------------------------------------------------------------- */
static int weight_sampling(const float* w, int n, double ran)
{
    double s = 0.0; /* GLMCMC.py:18-22: python float (double) running sum of float32 weights */
    for (int j = 0; j < n; ++j) {
        s += (double)w[j];
        if (ran < s) return j;
    }
    return -1; /* None */
}

static void run_isir_range(void* vctx, int64_t c_begin, int64_t c_end)
{
    const sampler_job* job = (const sampler_job*)vctx;
    const glabc_model_t* m = job->m;
    const glabc_dist_t* lp = job->lp;
    const glabc_dist_t* ip = job->gp;
    const glabc_run_t* r = job->r;
    const int K = r->n_candidates;
    const int d = m->theta_dim, yd = m->y_dim;
    const int slots = GLABC_TAPE_ISIR_SLOTS(d, yd, K);
    const int ns = GLABC_NSTATS(d);
    const float gf = r->global_frequency;
    const int64_t C = r->n_chains;

    for (int64_t c = c_begin; c < c_end; ++c) {
        float theta[GLABC_MAX_DIM], y[GLABC_MAX_DIM], st[GLABC_NSTATS(GLABC_MAX_DIM)];
        memcpy(theta, r->theta + c * d, sizeof(float) * d);
        memcpy(y, r->y + c * yd, sizeof(float) * yd);
        memset(st, 0, sizeof(st));
        float lw_old = r->aux[c * GLABC_AUX_SLOTS + GLABC_AUX_LOGW];
        int local = r->aux[c * GLABC_AUX_SLOTS + GLABC_AUX_LOCAL] != 0.0f;
        if (r->write_row0 && r->trace_layout != GLABC_TRACE_NONE)
            memcpy(r->trace + trace_index(r, c, r->step_base, d), theta, sizeof(float) * d);

        for (int64_t s = 0; s < r->n_steps; ++s) {
            const int64_t i = r->step_base + 1 + s;
            const uint64_t gid = (uint64_t)(r->chain_id_base + c);
            float u_b, u_a = 0.0f;
            float eps_p[GLABC_MAX_K * GLABC_MAX_DIM], eps_s[GLABC_MAX_K * GLABC_MAX_DIM];
            double u64 = 0.0;
            if (r->rng_mode == GLABC_RNG_REPLAY) {
                const float* t = r->tape32 + (size_t)s * slots * C + c;
                u_b = t[0];
                for (int k = 0; k < K * d; ++k) eps_p[k] = t[(size_t)(1 + k) * C];
                for (int k = 0; k < K * yd; ++k) eps_s[k] = t[(size_t)(1 + K * d + k) * C];
                u_a = t[(size_t)(1 + K * d + K * yd) * C];
                u64 = r->tape64[(size_t)s * C + c];
            } else {
                /* native streams: the step block gives U_b, U_a and the normals of a local move
                 * (as GlobalMCMC); candidate j of a global move uses normal blocks SLOT_NORMAL + 8 +
                 * j*G .. (G = blocks per candidate); the float64 resampling uniform is built from 53
                 * bits of block (i, slot 2^31).                                                     */
                const int G = (d + yd + 3) / 4;
                float zl[2 * GLABC_MAX_DIM + 4];
                native_step_draws(r->seed, gid, (uint32_t)i, d + yd, zl, &u_b, &u_a);
                for (int j = 0; j < K; ++j) {
                    float z[2 * GLABC_MAX_DIM + 4];
                    native_normals(r->seed, gid, (uint32_t)i, SLOT_NORMAL + 8u + (uint32_t)(j * G), d + yd, z);
                    memcpy(eps_p + j * d, z, sizeof(float) * d);
                    memcpy(eps_s + j * yd, z + d, sizeof(float) * yd);
                }
                if (!(u_b < gf)) { /* local move: its proposal / simulator normals come from the step block */
                    memcpy(eps_p, zl, sizeof(float) * d);
                    memcpy(eps_s, zl + d, sizeof(float) * yd);
                }
                /* 53-bit resampling uniform: the step block's 24-bit U_a field on top of the spare
                 * U_a (24) / U_b (top 5) fields of candidate 0's first normal block */
                uint32_t w0[4], c0[4];
                philox_block(r->seed, gid, (uint32_t)i, SLOT_STEP, w0);
                philox_block(r->seed, gid, (uint32_t)i, SLOT_NORMAL + 8u, c0);
                const uint64_t ua_step = (w0[1] & 0xFFFu) | ((w0[3] << 12) & 0xFFF000u);
                const uint64_t ua_c0 = (c0[1] & 0xFFFu) | ((c0[3] << 12) & 0xFFF000u);
                u64 = (double)((ua_step << 29) | (ua_c0 << 5) | ((c0[0] & 0xFFu) >> 3)) * 0x1p-53;
            }

            const int is_global = u_b < gf; /* GLMCMC.py:59 */
            int changed = 0, ind = -1;
            float dbg[GLABC_DEBUG_SLOTS];
            memset(dbg, 0, sizeof(dbg));
            float prev[GLABC_MAX_DIM];
            memcpy(prev, theta, sizeof(float) * d);
            if (is_global) {
                if (local) { /* GLMCMC.py:60-64 */
                    lw_old = (model_prior(m, theta) + model_log_kernel(m, y)) -
                             diag_gauss_log_prob(theta, ip->a, ip->b, ip->c, d);
                }
                local = 0;
                float th[(GLABC_MAX_K + 1) * GLABC_MAX_DIM], x[(GLABC_MAX_K + 1) * GLABC_MAX_DIM];
                float lw[GLABC_MAX_K + 1], w[GLABC_MAX_K + 1] = {0};
                memcpy(th, theta, sizeof(float) * d);
                memcpy(x, y, sizeof(float) * yd);
                lw[0] = lw_old;
                for (int j = 0; j < K; ++j) { /* GLMCMC.py:66-74 */
                    float* tj = th + (j + 1) * d;
                    float* xj = x + (j + 1) * yd;
                    const float lq = diag_gauss_forward(eps_p + j * d, ip->a, ip->b, ip->c, d, tj);
                    model_simulate(m, tj, eps_s + j * yd, xj);
                    lw[j + 1] = (model_prior(m, tj) + model_log_kernel(m, xj)) - lq;
                }
                for (int j = 0; j <= K; ++j) { /* GLMCMC.py:78-81: no max shift (B-1) */
                    w[j] = expf(lw[j]);
                    if (isnan(w[j])) w[j] = 0.0f;
                }
                const float S = torch_sum_f32(w, K + 1); /* GLMCMC.py:82 */
                for (int j = 0; j <= K; ++j) w[j] = w[j] / S;
                ind = weight_sampling(w, K + 1, u64); /* GLMCMC.py:83 */
                if (ind > 0) {                        /* GLMCMC.py:84-88 */
                    memcpy(theta, th + ind * d, sizeof(float) * d);
                    memcpy(y, x + ind * yd, sizeof(float) * yd);
                    lw_old = lw[ind];
                    changed = 1;
                }
                dbg[1] = lw[0];
                dbg[2] = S;
                dbg[3] = w[0];
                for (int j = 0; j < K; ++j) dbg[4 + j] = lw[j + 1];
            } else { /* GLMCMC.py:90-104 */
                float z[GLABC_MAX_DIM], theta_p[GLABC_MAX_DIM], y_p[GLABC_MAX_DIM];
                (void)diag_gauss_forward(eps_p, lp->a, lp->b, lp->c, d, z);
                for (int k = 0; k < d; ++k) theta_p[k] = z[k] + theta[k];
                model_simulate(m, theta_p, eps_s, y_p);
                const float prior_p = model_prior(m, theta_p);
                const float kern_p = model_log_kernel(m, y_p);
                float log_acc = prior_p + kern_p;
                log_acc = log_acc - model_prior(m, theta);
                log_acc = log_acc - model_log_kernel(m, y);
                if (logf(u_a) < log_acc) {
                    local = 1;
                    memcpy(theta, theta_p, sizeof(float) * d);
                    memcpy(y, y_p, sizeof(float) * yd);
                    changed = 1;
                }
                dbg[1] = prior_p;
                dbg[2] = kern_p;
                dbg[3] = log_acc;
            }
            dbg[0] = (float)(is_global | (changed << 1) | ((ind + 1) << 8));

            stats_update(st, d, theta, prev);
            st[GLABC_STAT_GLOBAL_STEPS] += (float)is_global;
            st[is_global ? GLABC_STAT_ACC_GLOBAL : GLABC_STAT_ACC_LOCAL] += (float)changed;
            if (r->trace_layout != GLABC_TRACE_NONE)
                memcpy(r->trace + trace_index(r, c, i, d), theta, sizeof(float) * d);
            if (r->debug) {
                float* g = r->debug + (size_t)s * GLABC_DEBUG_SLOTS * C + c;
                for (int k = 0; k < GLABC_DEBUG_SLOTS; ++k) g[(size_t)k * C] = dbg[k];
            }
        }
        memcpy(r->theta + c * d, theta, sizeof(float) * d);
        memcpy(r->y + c * yd, y, sizeof(float) * yd);
        r->aux[c * GLABC_AUX_SLOTS + GLABC_AUX_LOGW] = lw_old;
        r->aux[c * GLABC_AUX_SLOTS + GLABC_AUX_LOCAL] = local ? 1.0f : 0.0f;
        if (r->stats)
            for (int k = 0; k < ns; ++k) r->stats[c * ns + k] += st[k];
    }
}

ORACLE_EXPORT int oracle_run_isir(const glabc_model_t* m, const glabc_dist_t* lp, const glabc_dist_t* ip,
                                  const glabc_run_t* r)
{
    int rc = check_common(m, r);
    if (rc) return rc;
    if (!lp || !ip || lp->kind != GLABC_DIST_DIAG_GAUSSIAN || ip->kind != GLABC_DIST_DIAG_GAUSSIAN)
        return GLABC_ERR_UNSUPPORTED;
    const int K = r->n_candidates;
    if (K < 1 || K > GLABC_MAX_K || !r->aux) return GLABC_ERR_INVALID;
    if (r->rng_mode == GLABC_RNG_REPLAY && !r->tape64) return GLABC_ERR_INVALID;
    sampler_job job = {m, lp, ip, r};
    parallel_chains(run_isir_range, &job, r->n_chains);
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * GLMALA — GLMALA.py:150-200
 *
 * dtype bookkeeping the reference does by accident and which defines its behaviour:
 *   - the MALA proposal adds a float64 gradient (GLMALA.py:43), so theta', y', their log-densities
 *     and log_acc are float64; once a local move is accepted Theta_old / y_old ARE float64 tensors
 *     ("wide"), and torch.cat keeps them float64 through later global switches;
 *   - log_weight_old is computed once, at the first global move (`local` is never set again,
 *     GLMALA.py:152-157 vs :195-199): if the state was already wide then, the iSIR weights are
 *     exponentiated / normalised in float64 for the rest of the run ("lw_wide"), otherwise float32.
 * Draw order (local): U_b, [first gradient: N[num,y] per k, + and - share them], N[1,d] (z),
 * gradient at theta' (N[num,y] per k), N[1,y] (simulator), U_a.
 * ------------------------------------------------------------------------------------------- */
static double torch_sum_f64(const double* v, int n)
{
    double p[4] = {0.0, 0.0, 0.0, 0.0};
    const int rows = n / 4;
    for (int r = 0; r < rows; ++r)
        for (int k = 0; k < 4; ++k) p[k] += v[r * 4 + k];
    for (int i = rows * 4; i < n; ++i) p[0] += v[i];
    for (int k = 1; k < 4; ++k) p[0] += p[k];
    return p[0];
}

static inline double half_log_2pi_f64(int d) { return -0.5 * (double)d * log(2.0 * M_PI); }

/* DiagGaussian.log_prob on a float64 tensor with float32 parameters (type promotion) */
static double diag_gauss_log_prob64(const double* z, const float* loc, const float* log_scale, const float* scale, int d)
{
    double t[GLABC_MAX_DIM];
    for (int i = 0; i < d; ++i) {
        const double r = (z[i] - (double)loc[i]) / (double)scale[i];
        t[i] = (double)log_scale[i] + 0.5 * (r * r);
    }
    return half_log_2pi_f64(d) - torch_sum_f64(t, d);
}

static double model_prior64(const glabc_model_t* m, const double* theta)
{
    return diag_gauss_log_prob64(theta, m->prior_loc, m->prior_log_scale, m->prior_scale, m->theta_dim);
}

static double model_log_kernel64(const glabc_model_t* m, const double* y)
{
    double t[GLABC_MAX_DIM];
    for (int i = 0; i < m->y_dim; ++i) {
        const double dy = y[i] - (double)m->y_obs[i];
        t[i] = dy * dy;
    }
    const double dis = sqrt(torch_sum_f64(t, m->y_dim));
    const double r = (dis - 0.0) / (double)m->eps_scale;
    return half_log_2pi_f64(1) - ((double)m->eps_log_scale + 0.5 * (r * r));
}

/* native-mode layout of the gradient normals: draw j of dimension k takes y_dim normals of block
 * slot0 + k*nblk + j/dpb, dpb = draws per Philox block (4 / y_dim; 1 for y_dim 3)               */
enum { SLOT_GRAD = 0x10000, SLOT_GRAD0 = 0x20000 };
static inline int grad_dpb(int yd) { return yd == 3 ? 1 : 4 / yd; }

static void native_grad_normals(uint64_t seed, uint64_t chain, uint32_t step, uint32_t slot0, int d, int yd, int num, float* eps)
{
    const int dpb = grad_dpb(yd), nblk = (num + dpb - 1) / dpb;
    for (int k = 0; k < d; ++k)
        for (int g = 0; g < nblk; ++g) {
            uint32_t w[4];
            float z[4];
            philox_block(seed, chain, step, slot0 + (uint32_t)(k * nblk + g), w);
            box_muller(w[0], w[1], &z[0], &z[1]);
            box_muller(w[2], w[3], &z[2], &z[3]);
            for (int t = 0; t < dpb && g * dpb + t < num; ++t)
                for (int q = 0; q < yd; ++q) eps[((size_t)k * num + g * dpb + t) * yd + q] = z[t * yd + q];
        }
}

/* numberical_gradient_logABC, GLMALA.py:46-95.  eps[d][num][y_dim]: the simulator normals (the same
 * for theta + 0.1 e_k and theta - 0.1 e_k: both calls follow the same manual_seed, :76-83).
 * Mean / unbiased variance of the discrepancies in float64 (:70-72,86-89) — accumulated in one pass
 * around the noise-free discrepancy (the kernels use the same formulation; torch's reduction order
 * differs at the 1e-16 level).                                                                     */
static void mala_gradient(const glabc_model_t* m, const double* theta_in, int num, const float* eps, double* grad, double* gstat)
{   /* gstat (optional): mu_plus[4], mu_minus[4], Sigma_plus[4], Sigma_minus[4] — torch.mean / torch.var of :86-89 */
    const int d = m->theta_dim, yd = m->y_dim;
    float th[GLABC_MAX_DIM], zero[GLABC_MAX_DIM] = {0};
    for (int i = 0; i < d; ++i) th[i] = (float)theta_in[i]; /* theta.float(), :60 */
    const float h = (float)1e-1;                             /* d * torch.eye in float32, :63 */
    const double eps2 = m->epsilon * m->epsilon;             /* ABCset.epsilon ** 2 */
    for (int k = 0; k < d; ++k) {
        double logp[2];
        for (int sgn = 0; sgn < 2; ++sgn) {
            float tp[GLABC_MAX_DIM], y[GLABC_MAX_DIM];
            memcpy(tp, th, sizeof(float) * d);
            tp[k] = sgn == 0 ? th[k] + h : th[k] - h;        /* :66-67 */
            model_simulate(m, tp, zero, y);
            const double c = (double)model_discrepancy(m, y);
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < num; ++j) {
                model_simulate(m, tp, eps + ((size_t)k * num + j) * yd, y);
                const double dx = (double)model_discrepancy(m, y) - c; /* :78-83 */
                s1 += dx;
                s2 = fma(dx, dx, s2);
            }
            const double n = (double)num;
            const double mu = c + s1 / n;
            const double var = (s2 - s1 * s1 / n) / (n - 1.0);
            if (gstat && k < 4) { gstat[sgn * 4 + k] = mu; gstat[8 + sgn * 4 + k] = var; }
            logp[sgn] = -0.5 * log(var + eps2) - 0.5 * (mu * mu) / (var + eps2); /* :90-93 */
        }
        float ta[GLABC_MAX_DIM], tb[GLABC_MAX_DIM];
        memcpy(ta, th, sizeof(float) * d);
        memcpy(tb, th, sizeof(float) * d);
        const float fd = (float)0.00001;                     /* eye[k,:] * 0.00001 in float32, :84-85 */
        ta[k] = th[k] + fd;
        tb[k] = th[k] - fd;
        const float gprior = (model_prior(m, ta) - model_prior(m, tb)) / (float)(2 * 0.00001);
        grad[k] = (logp[0] - logp[1]) / (2 * 1e-1) + (double)gprior; /* :94-95 */
    }
}

/* the gradient alone (unit-tested against the reference's recorded gradients) */
ORACLE_EXPORT void oracle_mala_gradient(const glabc_model_t* m, const double* theta, int num, const float* eps, double* grad)
{
    mala_gradient(m, theta, num, eps, grad, NULL);
}

typedef struct { const glabc_model_t* m; const glabc_dist_t* ip; const glabc_run_t* r; } mala_job;

static void run_mala_range(void* vctx, int64_t c_begin, int64_t c_end)
{
    const mala_job* job = (const mala_job*)vctx;
    const glabc_model_t* m = job->m;
    const glabc_dist_t* ip = job->ip;
    const glabc_run_t* r = job->r;
    const int K = r->n_candidates, num = r->num_grad;
    const int d = m->theta_dim, yd = m->y_dim;
    const int slots = GLABC_TAPE_MALA_SLOTS(d, yd, K, num);
    const int gslot0 = 2 + K * (d + yd);
    const int ns = GLABC_NSTATS(d);
    const float gf = r->global_frequency;
    const float tau_f = r->tau;              /* z * tau: float32 tensor times Python scalar */
    const double tau = r->tau64 != 0.0 ? r->tau64 : (double)r->tau; /* the Python float itself */
    const int64_t C = r->n_chains;
    const size_t ng = (size_t)d * num * yd;
    float* geps = (float*)malloc(sizeof(float) * ng);
    const float zeros[GLABC_MAX_DIM] = {0}, ones[GLABC_MAX_DIM] = {1, 1, 1, 1, 1, 1, 1, 1};

    for (int64_t c = c_begin; c < c_end; ++c) {
        double theta[GLABC_MAX_DIM], y[GLABC_MAX_DIM], grad[GLABC_MAX_DIM] = {0}, lw_old;
        float st[GLABC_NSTATS(GLABC_MAX_DIM)];
        memset(st, 0, sizeof(st));
        float* aux = r->aux + c * GLABC_AUX_SLOTS;
        double* s64 = r->state64 + c * GLABC_STATE64_SLOTS;
        int local = aux[GLABC_AUX_LOCAL] != 0.0f, wide = aux[GLABC_AUX_WIDE] != 0.0f;
        int lw_wide = aux[GLABC_AUX_LW_WIDE] != 0.0f, have_grad = aux[GLABC_AUX_HAVE_GRAD] != 0.0f;
        for (int k = 0; k < d; ++k) theta[k] = wide ? s64[GLABC_S64_THETA + k] : (double)r->theta[c * d + k];
        for (int k = 0; k < yd; ++k) y[k] = wide ? s64[GLABC_S64_Y + k] : (double)r->y[c * yd + k];
        for (int k = 0; k < d; ++k) grad[k] = s64[GLABC_S64_GRAD + k];
        lw_old = s64[GLABC_S64_LOGW];
        if (r->write_row0 && r->trace_layout != GLABC_TRACE_NONE)
            for (int k = 0; k < d; ++k) r->trace[trace_index(r, c, r->step_base, d) + k] = (float)theta[k];

        for (int64_t s = 0; s < r->n_steps; ++s) {
            const int64_t i = r->step_base + 1 + s;
            const uint64_t gid = (uint64_t)(r->chain_id_base + c);
            const float* t = r->rng_mode == GLABC_RNG_REPLAY ? r->tape32 + (size_t)s * slots * C + c : NULL;
            float u_b, u_a = 0.0f, zl[2 * GLABC_MAX_DIM + 4];
            if (t) u_b = t[0];
            else native_step_draws(r->seed, gid, (uint32_t)i, d + yd, zl, &u_b, &u_a);
            const int is_global = u_b < gf; /* GLMALA.py:151 */
            int changed = 0, ind = -1;
            double dbg[GLABC_DEBUG64_SLOTS];
            memset(dbg, 0, sizeof(dbg));
            float prev[GLABC_MAX_DIM], now[GLABC_MAX_DIM];
            for (int k = 0; k < d; ++k) prev[k] = (float)theta[k];

            if (is_global) { /* GLMALA.py:151-180: the iSIR move of GLMCMC.py:60-89 */
                float eps_p[GLABC_MAX_K * GLABC_MAX_DIM], eps_s[GLABC_MAX_K * GLABC_MAX_DIM];
                double u64;
                if (t) {
                    for (int k = 0; k < K * d; ++k) eps_p[k] = t[(size_t)(1 + k) * C];
                    for (int k = 0; k < K * yd; ++k) eps_s[k] = t[(size_t)(1 + K * d + k) * C];
                    u64 = r->tape64[(size_t)s * C + c];
                } else {
                    const int G = (d + yd + 3) / 4;
                    for (int j = 0; j < K; ++j) {
                        float z[2 * GLABC_MAX_DIM + 4];
                        native_normals(r->seed, gid, (uint32_t)i, SLOT_NORMAL + 8u + (uint32_t)(j * G), d + yd, z);
                        memcpy(eps_p + j * d, z, sizeof(float) * d);
                        memcpy(eps_s + j * yd, z + d, sizeof(float) * yd);
                    }
                    uint32_t w0[4], c0[4];
                    philox_block(r->seed, gid, (uint32_t)i, SLOT_STEP, w0);
                    philox_block(r->seed, gid, (uint32_t)i, SLOT_NORMAL + 8u, c0);
                    const uint64_t ua_step = (w0[1] & 0xFFFu) | ((w0[3] << 12) & 0xFFF000u);
                    const uint64_t ua_c0 = (c0[1] & 0xFFFu) | ((c0[3] << 12) & 0xFFF000u);
                    u64 = (double)((ua_step << 29) | (ua_c0 << 5) | ((c0[0] & 0xFFu) >> 3)) * 0x1p-53;
                }
                if (local) { /* GLMALA.py:152-156 — the only place log_weight_old is computed from the state */
                    if (wide) {
                        lw_old = (model_prior64(m, theta) + model_log_kernel64(m, y)) -
                                 diag_gauss_log_prob64(theta, ip->a, ip->b, ip->c, d);
                    } else {
                        float tf[GLABC_MAX_DIM], yf[GLABC_MAX_DIM];
                        for (int k = 0; k < d; ++k) tf[k] = (float)theta[k];
                        for (int k = 0; k < yd; ++k) yf[k] = (float)y[k];
                        lw_old = (double)((model_prior(m, tf) + model_log_kernel(m, yf)) - diag_gauss_log_prob(tf, ip->a, ip->b, ip->c, d));
                    }
                    lw_wide = wide;
                }
                local = 0;
                float th[(GLABC_MAX_K + 1) * GLABC_MAX_DIM], x[(GLABC_MAX_K + 1) * GLABC_MAX_DIM], lw[GLABC_MAX_K + 1];
                for (int j = 0; j < K; ++j) { /* GLMALA.py:158-165 */
                    float* tj = th + (j + 1) * d;
                    float* xj = x + (j + 1) * yd;
                    const float lq = diag_gauss_forward(eps_p + j * d, ip->a, ip->b, ip->c, d, tj);
                    model_simulate(m, tj, eps_s + j * yd, xj);
                    lw[j + 1] = (model_prior(m, tj) + model_log_kernel(m, xj)) - lq;
                }
                double S, w0n;
                if (lw_wide) { /* float64 weights: exp does not underflow near -104 (contrast B-1) */
                    double w[GLABC_MAX_K + 1];
                    w[0] = exp(lw_old);
                    for (int j = 1; j <= K; ++j) w[j] = exp((double)lw[j]);
                    for (int j = 0; j <= K; ++j) if (isnan(w[j])) w[j] = 0.0;
                    S = torch_sum_f64(w, K + 1);
                    double run = 0.0;
                    for (int j = 0; j <= K; ++j) {
                        w[j] = w[j] / S;
                        run += w[j];
                        if (ind < 0 && u64 < run) ind = j;
                    }
                    w0n = w[0];
                } else {
                    float w[GLABC_MAX_K + 1];
                    w[0] = expf((float)lw_old);
                    for (int j = 1; j <= K; ++j) w[j] = expf(lw[j]);
                    for (int j = 0; j <= K; ++j) if (isnan(w[j])) w[j] = 0.0f;
                    const float Sf = torch_sum_f32(w, K + 1);
                    for (int j = 0; j <= K; ++j) w[j] = w[j] / Sf;
                    ind = weight_sampling(w, K + 1, u64);
                    S = (double)Sf;
                    w0n = (double)w[0];
                }
                dbg[1] = lw_old;
                dbg[2] = S;
                dbg[3] = w0n;
                for (int j = 0; j < K; ++j) dbg[4 + j] = (double)lw[j + 1];
                if (ind > 0) { /* GLMALA.py:175-179: grad_logABC_Theta_old is NOT refreshed (B-6) */
                    for (int k = 0; k < d; ++k) theta[k] = (double)th[ind * d + k];
                    for (int k = 0; k < yd; ++k) y[k] = (double)x[ind * yd + k];
                    lw_old = (double)lw[ind];
                }
                for (int k = 0; k < d; ++k) changed |= ((float)theta[k] != prev[k]);
            } else { /* GLMALA.py:182-200 */
                float z[GLABC_MAX_DIM], eps_s[GLABC_MAX_DIM];
                if (!have_grad) { /* :183-184 */
                    if (t) for (size_t q = 0; q < ng; ++q) geps[q] = r->tape_grad0[q * C + c];
                    else native_grad_normals(r->seed, gid, (uint32_t)i, SLOT_GRAD0, d, yd, num, geps);
                    mala_gradient(m, theta, num, geps, grad, NULL);
                    have_grad = 1;
                }
                if (t) {
                    for (int k = 0; k < d; ++k) z[k] = t[(size_t)(1 + k) * C];
                    for (int k = 0; k < yd; ++k) eps_s[k] = t[(size_t)(1 + K * d + k) * C];
                    u_a = t[(size_t)(1 + K * (d + yd)) * C];
                    for (size_t q = 0; q < ng; ++q) geps[q] = t[(gslot0 + q) * C];
                } else {
                    memcpy(z, zl, sizeof(float) * d);
                    memcpy(eps_s, zl + d, sizeof(float) * yd);
                    native_grad_normals(r->seed, gid, (uint32_t)i, SLOT_GRAD, d, yd, num, geps);
                }
                /* Local_proposal_forward, GLMALA.py:25-44: DiagGaussian(d, [0], [0]).forward(1) */
                float zz[GLABC_MAX_DIM];
                const float lq_fwd = diag_gauss_forward(z, zeros, zeros, ones, d, zz);
                double theta_p[GLABC_MAX_DIM], y_p[GLABC_MAX_DIM], grad_p[GLABC_MAX_DIM];
                for (int k = 0; k < d; ++k) {
                    const float zt = zz[k] * tau_f;
                    const double a = wide ? (double)zt + theta[k] : (double)(zt + (float)theta[k]);
                    theta_p[k] = a + grad[k] * (tau * tau) / 2.0; /* :43 */
                }
                mala_gradient(m, theta_p, num, geps, grad_p, dbg + 20); /* :187 */
                for (int k = 0; k < yd; ++k) { /* :188-189: |theta'| (float64) + likelihood.sample (float32) */
                    const float noise = m->noise_loc[k] + m->noise_scale[k] * eps_s[k];
                    const double mean = m->family == GLABC_MODEL_ABS_NORMAL ? fabs(theta_p[k]) : theta_p[k];
                    y_p[k] = mean + (double)noise;
                }
                const double prior_p = model_prior64(m, theta_p), kern_p = model_log_kernel64(m, y_p);
                double rr[GLABC_MAX_DIM]; /* log_proposal(Theta_prop, grad_prop, Theta_old, tau), :97-116 */
                for (int k = 0; k < d; ++k) rr[k] = (theta[k] - theta_p[k] - grad_p[k] * (tau * tau) / 2.0) / tau;
                const double lq_rev = diag_gauss_log_prob64(rr, zeros, zeros, ones, d);
                double prior_o, kern_o;
                if (wide) {
                    prior_o = model_prior64(m, theta);
                    kern_o = model_log_kernel64(m, y);
                } else {
                    float tf[GLABC_MAX_DIM], yf[GLABC_MAX_DIM];
                    for (int k = 0; k < d; ++k) tf[k] = (float)theta[k];
                    for (int k = 0; k < yd; ++k) yf[k] = (float)y[k];
                    prior_o = (double)model_prior(m, tf);
                    kern_o = (double)model_log_kernel(m, yf);
                }
                const double log_acc = prior_p + kern_p + lq_rev - prior_o - kern_o - (double)lq_fwd; /* :190-193 */
                const int accept = (double)logf(u_a) < log_acc;
                dbg[1] = log_acc;
                for (int k = 0; k < d && k < 4; ++k) { dbg[2 + k] = theta_p[k]; dbg[10 + k] = grad_p[k]; }
                for (int k = 0; k < yd && k < 4; ++k) dbg[6 + k] = y_p[k];
                dbg[14] = prior_p; dbg[15] = kern_p; dbg[16] = lq_rev; dbg[17] = (double)lq_fwd;
                if (accept) { /* :195-199 */
                    memcpy(theta, theta_p, sizeof(double) * d);
                    memcpy(y, y_p, sizeof(double) * yd);
                    memcpy(grad, grad_p, sizeof(double) * d);
                    wide = 1;
                    changed = 1;
                }
            }
            dbg[0] = (double)(is_global | (changed << 1) | ((is_global ? ind + 1 : 0) << 8) | ((is_global && lw_wide) << 16));
            for (int k = 0; k < d; ++k) now[k] = (float)theta[k]; /* Theta_Re[i,:] = Theta_old (float32 buffer) */
            stats_update(st, d, now, prev);
            st[GLABC_STAT_GLOBAL_STEPS] += (float)is_global;
            st[is_global ? GLABC_STAT_ACC_GLOBAL : GLABC_STAT_ACC_LOCAL] += (float)changed;
            if (r->trace_layout != GLABC_TRACE_NONE)
                memcpy(r->trace + trace_index(r, c, i, d), now, sizeof(float) * d);
            if (r->debug64) {
                double* g = r->debug64 + (size_t)s * GLABC_DEBUG64_SLOTS * C + c;
                for (int k = 0; k < GLABC_DEBUG64_SLOTS; ++k) g[(size_t)k * C] = dbg[k];
            }
        }
        for (int k = 0; k < d; ++k) { r->theta[c * d + k] = (float)theta[k]; s64[GLABC_S64_THETA + k] = theta[k]; s64[GLABC_S64_GRAD + k] = grad[k]; }
        for (int k = 0; k < yd; ++k) { r->y[c * yd + k] = (float)y[k]; s64[GLABC_S64_Y + k] = y[k]; }
        s64[GLABC_S64_LOGW] = lw_old;
        aux[GLABC_AUX_LOCAL] = (float)local; aux[GLABC_AUX_WIDE] = (float)wide;
        aux[GLABC_AUX_LW_WIDE] = (float)lw_wide; aux[GLABC_AUX_HAVE_GRAD] = (float)have_grad;
        if (r->stats)
            for (int k = 0; k < ns; ++k) r->stats[c * ns + k] += st[k];
    }
    free(geps);
}

ORACLE_EXPORT int oracle_run_mala(const glabc_model_t* m, const glabc_dist_t* unused, const glabc_dist_t* ip,
                                  const glabc_run_t* r)
{
    (void)unused;
    int rc = check_common(m, r);
    if (rc) return rc;
    if (!ip || ip->kind != GLABC_DIST_DIAG_GAUSSIAN) return GLABC_ERR_UNSUPPORTED;
    if (m->theta_dim > 4) return GLABC_ERR_UNSUPPORTED;
    if (r->n_candidates < 1 || r->n_candidates > GLABC_MAX_K || !r->aux || !r->state64) return GLABC_ERR_INVALID;
    if (r->num_grad < 2 || r->num_grad > GLABC_MAX_NUM_GRAD || !(r->tau > 0.0f)) return GLABC_ERR_INVALID;
    if (r->rng_mode == GLABC_RNG_REPLAY && (!r->tape64 || !r->tape_grad0)) return GLABC_ERR_INVALID;
    mala_job job = {m, ip, r};
    parallel_chains(run_mala_range, &job, r->n_chains);
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * KernelDensity — kernel_density.py:22-177
 * ------------------------------------------------------------------------------------------- */
/* fit, :70-94 + _compute_bandwidth :22-37 + weighted_std :39-68.  Sums in float64 (torch reduces the
 * float32 tensors with a vectorised cascade; the two agree to float32 rounding).                     */
ORACLE_EXPORT int oracle_kde_fit(const float* X, const float* w, int64_t n, int32_t d, int32_t rule, float* weights, float* bw)
{
    if (!X || !weights || !bw || n < 1 || d < 1 || d > GLABC_MAX_DIM) return GLABC_ERR_INVALID;
    if (w) {
        double sw = 0.0;
        for (int64_t j = 0; j < n; ++j) sw += (double)w[j];
        const float swf = (float)sw;
        for (int64_t j = 0; j < n; ++j) weights[j] = w[j] / swf; /* :83 */
    } else {
        const float u = 1.0f / (float)n; /* :80 */
        for (int64_t j = 0; j < n; ++j) weights[j] = u;
    }
    const double h = rule == GLABC_BW_SILVERMAN ? pow((double)n * (d + 2) / 4., -1. / (d + 4)) : pow((double)n, -1. / (d + 4));
    double s2 = 0.0, sw2 = 0.0; /* weighted_std: w = weights / weights.sum() once more, :53 */
    for (int64_t j = 0; j < n; ++j) s2 += (double)weights[j];
    const float s2f = (float)s2;
    for (int64_t j = 0; j < n; ++j) { const float wj = weights[j] / s2f; sw2 += (double)(wj * wj); }
    float corr = 1.0f - (float)sw2; /* :64 */
    if (corr < 1e-10f) corr = 1e-10f;
    for (int i = 0; i < d; ++i) {
        double mean = 0.0, var = 0.0;
        for (int64_t j = 0; j < n; ++j) mean += (double)((weights[j] / s2f) * X[j * d + i]); /* :56 */
        const float mf = (float)mean;
        for (int64_t j = 0; j < n; ++j) {
            const float df = X[j * d + i] - mf;
            var += (double)((weights[j] / s2f) * (df * df)); /* :59-61 */
        }
        bw[i] = (float)h * sqrtf((float)var / corr); /* :36,65-68 */
    }
    return GLABC_OK;
}

/* log_prob, :96-128: max-shifted logsumexp over the n training points */
static float kde_log_prob_one(const float* X, const float* weights, const float* bw, int64_t n, int d, const float* x)
{
    float slog = 0.0f; /* torch.log(bandwidth).sum() */
    {
        float t[GLABC_MAX_DIM];
        for (int i = 0; i < d; ++i) t[i] = logf(bw[i]);
        slog = torch_sum_f32(t, d);
    }
    const float c = 0.5f * (float)d * logf((float)(2 * M_PI)); /* 0.5*dim*log(tensor(2*pi)), float32 */
    float mx = -INFINITY;
    for (int pass = 0; pass < 2; ++pass) {
        double acc = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            float t[GLABC_MAX_DIM];
            for (int i = 0; i < d; ++i) {
                const float df = (x[i] - X[j * d + i]) / bw[i];
                t[i] = df * df;
            }
            float lk = -0.5f * torch_sum_f32(t, d);
            lk = lk - c;
            lk = lk - slog;
            const float v = lk + logf(weights[j] + 1e-10f);
            if (pass == 0) { if (v > mx) mx = v; }
            else acc += (double)expf(v - mx);
        }
        if (pass == 1) return mx + logf((float)acc);
        if (isinf(mx)) mx = 0.0f; /* torch.logsumexp: infinite max is replaced by 0 */
    }
    return 0.0f;
}

ORACLE_EXPORT int oracle_kde_log_prob(const float* X, const float* weights, const float* bw, int64_t n, int32_t d,
                                      const float* x, int64_t m, float* out)
{
    if (!X || !weights || !bw || !x || !out || n < 1 || d < 1 || d > GLABC_MAX_DIM) return GLABC_ERR_INVALID;
    for (int64_t q = 0; q < m; ++q) out[q] = kde_log_prob_one(X, weights, bw, n, d, x + q * d);
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * AGLMCMC — AGLMCMC.py:84-272.  Draw order: init N[B,d], N[B,y]; per
 * iteration U_b; global: U64 (+ at an adaptation: multinomial[4B], N[4B,d], N[B,y]); local: N[1,d],
 * N[1,y], U_a.  Deviations from the reference, both on purpose: the chain is returned
 * for any num_ite and row 0 holds the initial theta.
 * ------------------------------------------------------------------------------------------- */
enum { SLOT_INIT = 0x30000, SLOT_AD_SAMPLE = 0x40000000, SLOT_AD_SIM = 0x48000000 };

typedef struct {
    const glabc_model_t* m; const glabc_dist_t* lp; const glabc_dist_t* ip; const glabc_run_t* r; const glabc_aglmcmc_t* ag;
} ag_job;

static int cmp_float(const void* a, const void* b)
{
    const float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* log N(dis; 0, eps) with eps given as a float32 value: DiagGaussian(1, 0, log(tensor([eps]))).log_prob, Mixture.py:47-53 */
static float log_kernel_dis_eps(float dis, float eps)
{
    const float ls = logf(eps);
    const float r = (dis - 0.0f) / expf(ls);
    return half_log_2pi_f32(1) - (ls + 0.5f * (r * r));
}

/* the block of one chain after its proposals theta0 / lq0 are known: simulate, discrepancy, weights */
static void ag_finish_block(const glabc_model_t* m, int B, const float* th0, const float* lq0, const float* eps_s,
                            float* x0, float* dis0, float* w0, int nan_to_zero)
{
    const int d = m->theta_dim, yd = m->y_dim;
    int all_nan = 1;
    for (int b = 0; b < B; ++b) {
        model_simulate(m, th0 + b * d, eps_s + b * yd, x0 + b * yd);
        dis0[b] = model_discrepancy(m, x0 + b * yd);
        if (!isnan(dis0[b])) all_nan = 0;
    }
    if (all_nan) for (int b = 0; b < B; ++b) dis0[b] = 1000000 - 5; /* AGLMCMC.py:100-101 (scalar torch.all) */
    for (int b = 0; b < B; ++b) {
        const float lw = (model_prior(m, th0 + b * d) + model_log_kernel_dis(m, dis0[b])) - lq0[b];
        float w = expf(lw);
        if ((nan_to_zero && isnan(w)) || all_nan) w = 0.0f; /* :110-112 (initial block) / :248-249 */
        w0[b] = w;
    }
}

static void run_aglmcmc_range(void* vctx, int64_t c_begin, int64_t c_end)
{
    const ag_job* job = (const ag_job*)vctx;
    const glabc_model_t* m = job->m;
    const glabc_dist_t* lp = job->lp;
    const glabc_dist_t* ip = job->ip;
    const glabc_run_t* r = job->r;
    const glabc_aglmcmc_t* ag = job->ag;
    const int K = r->n_candidates, S = ag->step_size, B = K * S;
    const int d = m->theta_dim, yd = m->y_dim;
    const int slots = GLABC_TAPE_GLOBAL_SLOTS(d, yd);
    const int ns = GLABC_NSTATS(d);
    const float gf = r->global_frequency;
    const int64_t C = r->n_chains;
    const int replay = r->rng_mode == GLABC_RNG_REPLAY;
    float* th0 = (float*)malloc(sizeof(float) * B * d);
    float* x0 = (float*)malloc(sizeof(float) * B * yd);
    float* lq0 = (float*)malloc(sizeof(float) * B);
    float* w0 = (float*)malloc(sizeof(float) * B);
    float* dis0 = (float*)malloc(sizeof(float) * B);
    float* eps_s = (float*)malloc(sizeof(float) * B * yd);
    float* kX = (float*)malloc(sizeof(float) * B * d);
    float* kw = (float*)malloc(sizeof(float) * B);
    float* kwn = (float*)malloc(sizeof(float) * B);
    float* smp = (float*)malloc(sizeof(float) * 4 * B * d);
    float* sorted = (float*)malloc(sizeof(float) * B);
    double* cdf = (double*)malloc(sizeof(double) * B);

    for (int64_t c = c_begin; c < c_end; ++c) {
        const uint64_t gid = (uint64_t)(r->chain_id_base + c);
        float theta[GLABC_MAX_DIM], y[GLABC_MAX_DIM], st[GLABC_NSTATS(GLABC_MAX_DIM)], kbw[GLABC_MAX_DIM] = {0};
        memcpy(theta, r->theta + c * d, sizeof(float) * d);
        memcpy(y, r->y + c * yd, sizeof(float) * yd);
        memset(st, 0, sizeof(st));
        if (r->write_row0 && r->trace_layout != GLABC_TRACE_NONE)
            memcpy(r->trace + trace_index(r, c, r->step_base, d), theta, sizeof(float) * d);

        /* ---- initial block, AGLMCMC.py:84-112 ---- */
        for (int b = 0; b < B; ++b) {
            float eps_p[GLABC_MAX_DIM];
            if (replay) {
                for (int k = 0; k < d; ++k) eps_p[k] = ag->init_p[(size_t)(b * d + k) * C + c];
                for (int k = 0; k < yd; ++k) eps_s[b * yd + k] = ag->init_s[(size_t)(b * yd + k) * C + c];
            } else {
                const int G = (d + yd + 3) / 4;
                float z[2 * GLABC_MAX_DIM + 4];
                native_normals(r->seed, gid, 0u, SLOT_INIT + (uint32_t)(b * G), d + yd, z);
                memcpy(eps_p, z, sizeof(float) * d);
                memcpy(eps_s + b * yd, z + d, sizeof(float) * yd);
            }
            lq0[b] = diag_gauss_forward(eps_p, ip->a, ip->b, ip->c, d, th0 + b * d);
        }
        ag_finish_block(m, B, th0, lq0, eps_s, x0, dis0, w0, 1);
        if (ag->init_w) for (int b = 0; b < B; ++b) ag->init_w[(size_t)b * C + c] = w0[b];
        int kk = 0, num_train = 0, kn = 0;
        float hat_eps = 1000000.0f;

        for (int64_t s = 0; s < r->n_steps; ++s) {
            const int64_t i = r->step_base + 1 + s;
            const float* t = replay ? r->tape32 + (size_t)s * slots * C + c : NULL;
            float u_b, u_a = 0.0f, zl[2 * GLABC_MAX_DIM + 4];
            double u64 = 0.0;
            if (t) {
                u_b = t[0];
                u64 = r->tape64[(size_t)s * C + c];
            } else {
                native_step_draws(r->seed, gid, (uint32_t)i, d + yd, zl, &u_b, &u_a);
                uint32_t w[4]; /* a global move does not use the step block's normals: their bits give the 53-bit uniform */
                philox_block(r->seed, gid, (uint32_t)i, SLOT_STEP, w);
                u64 = (double)(((uint64_t)(w[0] >> 8) << 29) | ((uint64_t)(w[2] >> 8) << 5) | (uint64_t)(w[1] >> 27)) * 0x1p-53;
            }
            const int is_global = u_b < gf; /* AGLMCMC.py:125-126 */
            int changed = 0, ind = -1;
            float dbg[4] = {0, 0, 0, 0}, prev[GLABC_MAX_DIM];
            memcpy(prev, theta, sizeof(float) * d);
            if (is_global) {
                /* :137-149 */
                const float lq_old = num_train == 0 ? diag_gauss_log_prob(theta, ip->a, ip->b, ip->c, d)
                                                    : kde_log_prob_one(kX, kwn, kbw, kn, d, theta);
                const float w_old = expf((model_prior(m, theta) + model_log_kernel(m, y)) - lq_old);
                float w[GLABC_MAX_K + 1];
                w[0] = w_old;
                for (int j = 0; j < K; ++j) w[j + 1] = w0[kk * K + j];
                const float Ssum = torch_sum_f32(w, K + 1); /* :155 */
                for (int j = 0; j <= K; ++j) w[j] = w[j] / Ssum;
                ind = weight_sampling(w, K + 1, u64); /* :158 */
                if (ind > 0) { /* :161-163 */
                    memcpy(theta, th0 + (kk * K + ind - 1) * d, sizeof(float) * d);
                    memcpy(y, x0 + (kk * K + ind - 1) * yd, sizeof(float) * yd);
                }
                for (int k = 0; k < d; ++k) changed |= theta[k] != prev[k];
                dbg[1] = lq_old; dbg[2] = w_old; dbg[3] = Ssum;
                kk += 1;
                if (kk == S) { /* ---- adaptation, :170-249 ---- */
                    kk = 0;
                    if (hat_eps > ag->hat_eps_T) { /* :174-196 */
                        int num_a = 0, nv = 0;
                        for (int b = 0; b < B; ++b) {
                            if (dis0[b] < hat_eps) ++num_a;
                            if (!isnan(dis0[b])) sorted[nv++] = dis0[b];
                        }
                        if (nv > 0) {
                            float q = (ag->alpha * (float)num_a) / (float)nv;
                            q = q < 0.0f ? 0.0f : (q > 1.0f ? 1.0f : q);
                            qsort(sorted, (size_t)nv, sizeof(float), cmp_float);
                            /* torch.quantile, linear interpolation: rank = q*(n-1); lerp(below, above, frac) */
                            const float rank = q * (float)(nv - 1);
                            const float lo = floorf(rank);
                            const float fr = rank - lo;
                            const float a = sorted[(int)lo], bq = sorted[(int)ceilf(rank)];
                            hat_eps = fr < 0.5f ? a + fr * (bq - a) : bq - (bq - a) * (1.0f - fr);
                        }
                        if (!(hat_eps > ag->hat_eps_T)) hat_eps = ag->hat_eps_T; /* torch.max, :196 */
                    }
                    kn = 0;
                    for (int b = 0; b < B; ++b) { /* :199-208 */
                        const float tw = expf((model_prior(m, th0 + b * d) + log_kernel_dis_eps(dis0[b], hat_eps)) - lq0[b]);
                        if (tw > 0.0f) {
                            memcpy(kX + kn * d, th0 + b * d, sizeof(float) * d);
                            kw[kn++] = tw;
                        }
                    }
                    {   /* :211 Train_weight / torch.sum(Train_weight) */
                        double sw = 0.0;
                        for (int j = 0; j < kn; ++j) sw += (double)kw[j];
                        for (int j = 0; j < kn; ++j) kw[j] = kw[j] / (float)sw;
                    }
                    oracle_kde_fit(kX, kw, kn, d, ag->kde_rule, kwn, kbw); /* :214-215 */
                    const int rr = num_train;
                    num_train += 1;
                    /* KDE.sample(4B), kernel_density.py:130-152; keep the first B with prior > log(1e-10), :220-226 */
                    if (!replay) {
                        double run = 0.0;
                        for (int j = 0; j < kn; ++j) { run += (double)kwn[j]; cdf[j] = run; }
                    }
                    int nb = 0;
                    for (int q = 0; q < 4 * B && nb < B; ++q) {
                        int idx;
                        float nz[GLABC_MAX_DIM + 4];
                        if (replay) {
                            idx = ag->ad_idx[((size_t)rr * 4 * B + q) * C + c];
                            for (int k = 0; k < d; ++k) nz[k] = ag->ad_noise[((size_t)rr * 4 * B * d + (size_t)q * d + k) * C + c];
                        } else {
                            uint32_t w[4];
                            philox_block(r->seed, gid, (uint32_t)rr, SLOT_AD_SAMPLE + 2u * (uint32_t)q, w);
                            const double u = (double)w[0] * 0x1p-32;
                            idx = kn - 1;
                            for (int j = 0; j < kn; ++j) if (u < cdf[j]) { idx = j; break; }
                            native_normals(r->seed, gid, (uint32_t)rr, SLOT_AD_SAMPLE + 2u * (uint32_t)q + 1u, d, nz);
                        }
                        float cand[GLABC_MAX_DIM];
                        for (int k = 0; k < d; ++k) cand[k] = kX[idx * d + k] + nz[k] * kbw[k];
                        if (model_prior(m, cand) > (float)log(1e-10)) memcpy(th0 + (nb++) * d, cand, sizeof(float) * d);
                    }
                    /* fewer than B valid samples: the reference raises IndexError; the remaining rows keep their old
                     * candidates here (cannot happen for the fused Gaussian priors short of |theta| > 6 sigma x 4B) */
                    oracle_kde_log_prob(kX, kwn, kbw, kn, d, th0, B, lq0); /* :229 */
                    for (int b = 0; b < B; ++b)
                        for (int k = 0; k < yd; ++k) {
                            if (replay) eps_s[b * yd + k] = ag->ad_sim[((size_t)rr * B * yd + (size_t)b * yd + k) * C + c];
                        }
                    if (!replay)
                        for (int b = 0; b < B; ++b) native_normals(r->seed, gid, (uint32_t)rr, SLOT_AD_SIM + (uint32_t)b, yd, eps_s + b * yd);
                    ag_finish_block(m, B, th0, lq0, eps_s, x0, dis0, w0, 0); /* :232-249 */
                    if (ag->ad_rec && rr < ag->dump_rounds) {
                        float* g = ag->ad_rec + (size_t)rr * GLABC_AG_REC_SLOTS * C + c;
                        g[0] = hat_eps; g[(size_t)1 * C] = (float)kn;
                        for (int k = 0; k < d && k < 4; ++k) g[(size_t)(2 + k) * C] = kbw[k];
                    }
                    if (ag->ad_blk && rr < ag->dump_rounds)
                        for (int b = 0; b < B; ++b) {
                            float* g = ag->ad_blk + ((size_t)rr * B + b) * (d + 3) * C + c;
                            for (int k = 0; k < d; ++k) g[(size_t)k * C] = th0[b * d + k];
                            g[(size_t)d * C] = lq0[b]; g[(size_t)(d + 1) * C] = w0[b]; g[(size_t)(d + 2) * C] = dis0[b];
                        }
                }
            } else { /* local RW-MH, :251-271 */
                float eps_p[GLABC_MAX_DIM], eps_l[GLABC_MAX_DIM], z[GLABC_MAX_DIM], theta_p[GLABC_MAX_DIM], y_p[GLABC_MAX_DIM];
                if (t) {
                    for (int k = 0; k < d; ++k) eps_p[k] = t[(size_t)(1 + k) * C];
                    for (int k = 0; k < yd; ++k) eps_l[k] = t[(size_t)(1 + d + k) * C];
                    u_a = t[(size_t)(1 + d + yd) * C];
                } else {
                    memcpy(eps_p, zl, sizeof(float) * d);
                    memcpy(eps_l, zl + d, sizeof(float) * yd);
                }
                (void)diag_gauss_forward(eps_p, lp->a, lp->b, lp->c, d, z);
                for (int k = 0; k < d; ++k) theta_p[k] = z[k] + theta[k];
                model_simulate(m, theta_p, eps_l, y_p);
                const float prior_p = model_prior(m, theta_p), kern_p = model_log_kernel(m, y_p);
                float log_acc = prior_p + kern_p;
                log_acc = log_acc - model_prior(m, theta);
                log_acc = log_acc - model_log_kernel(m, y);
                if (logf(u_a) < log_acc) {
                    memcpy(theta, theta_p, sizeof(float) * d);
                    memcpy(y, y_p, sizeof(float) * yd);
                    changed = 1;
                }
                dbg[1] = prior_p; dbg[2] = kern_p; dbg[3] = log_acc;
            }
            dbg[0] = (float)(is_global | (changed << 1) | ((is_global ? ind + 1 : 0) << 8));
            stats_update(st, d, theta, prev);
            st[GLABC_STAT_GLOBAL_STEPS] += (float)is_global;
            st[is_global ? GLABC_STAT_ACC_GLOBAL : GLABC_STAT_ACC_LOCAL] += (float)changed;
            if (r->trace_layout != GLABC_TRACE_NONE)
                memcpy(r->trace + trace_index(r, c, i, d), theta, sizeof(float) * d);
            if (r->debug) {
                float* g = r->debug + (size_t)s * GLABC_DEBUG_SLOTS * C + c;
                for (int k = 0; k < 4; ++k) g[(size_t)k * C] = dbg[k];
            }
        }
        memcpy(r->theta + c * d, theta, sizeof(float) * d);
        memcpy(r->y + c * yd, y, sizeof(float) * yd);
        if (r->stats)
            for (int k = 0; k < ns; ++k) r->stats[c * ns + k] += st[k];
    }
    free(th0); free(x0); free(lq0); free(w0); free(dis0); free(eps_s); free(kX); free(kw); free(kwn); free(smp);
    free(sorted); free(cdf);
}

ORACLE_EXPORT int oracle_run_aglmcmc(const glabc_model_t* m, const glabc_dist_t* lp, const glabc_dist_t* ip,
                                     const glabc_run_t* r, const glabc_aglmcmc_t* ag)
{
    int rc = check_common(m, r);
    if (rc) return rc;
    if (!ag || !lp || !ip || lp->kind != GLABC_DIST_DIAG_GAUSSIAN || ip->kind != GLABC_DIST_DIAG_GAUSSIAN) return GLABC_ERR_UNSUPPORTED;
    const int K = r->n_candidates;
    if (K < 1 || K > GLABC_MAX_K || ag->step_size < 1 || K * ag->step_size > GLABC_AG_MAX_BLOCK) return GLABC_ERR_INVALID;
    if (r->rng_mode == GLABC_RNG_REPLAY && (!r->tape64 || !ag->init_p || !ag->init_s || !ag->ad_idx || !ag->ad_noise || !ag->ad_sim))
        return GLABC_ERR_INVALID;
    ag_job job = {m, lp, ip, r, ag};
    parallel_chains(run_aglmcmc_range, &job, r->n_chains);
    return GLABC_OK;
}

/* -------------------------------------------------------------------------------------------
 * esjd — ESJD.py:17-24: det(D^T D / (N-1))^(1/d) in float32 (d <= 3 closed-form determinant;
 * torch.det goes through an LU factorisation, so agreement is to rounding, not bit-exact).
 * ------------------------------------------------------------------------------------------- */
ORACLE_EXPORT int oracle_esjd(const float* trace, int32_t layout, int64_t rows, int64_t chains, int32_t d, float* out)
{
    if (!trace || !out || d < 1 || d > 3 || rows < 2) return GLABC_ERR_INVALID;
    for (int64_t c = 0; c < chains; ++c) {
        float G[3][3] = {{0}};
        for (int64_t t = 1; t < rows; ++t) {
            float dl[3];
            for (int k = 0; k < d; ++k) {
                const size_t a = layout == GLABC_TRACE_CHAIN_MAJOR ? (size_t)((c * rows + t) * d + k) : (size_t)((t * chains + c) * d + k);
                const size_t b = layout == GLABC_TRACE_CHAIN_MAJOR ? (size_t)((c * rows + t - 1) * d + k) : (size_t)(((t - 1) * chains + c) * d + k);
                dl[k] = trace[a] - trace[b];
            }
            for (int p = 0; p < d; ++p)
                for (int q = 0; q < d; ++q) G[p][q] += dl[p] * dl[q];
        }
        const float n = (float)(rows - 1);
        for (int p = 0; p < d; ++p)
            for (int q = 0; q < d; ++q) G[p][q] = G[p][q] / n;
        float det;
        if (d == 1) det = G[0][0];
        else if (d == 2) det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
        else det = G[0][0] * (G[1][1] * G[2][2] - G[1][2] * G[2][1]) - G[0][1] * (G[1][0] * G[2][2] - G[1][2] * G[2][0]) +
                   G[0][2] * (G[1][0] * G[2][1] - G[1][1] * G[2][0]);
        out[c] = powf(det, 1.0f / (float)d);
    }
    return GLABC_OK;
}

ORACLE_EXPORT int oracle_num_threads(void) { return effective_threads(); }

/* n <= 0: all online cores */
ORACLE_EXPORT void oracle_set_num_threads(int n) { g_threads = n > 0 ? n : 0; }
