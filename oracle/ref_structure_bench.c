/*
 * ref_structure_bench.c -- the CPU baseline legs of bench.py.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see scorus_oracle.h).
 *
 * orc_ref_structure_run restates one sample_pt call (src/mcmc/ensemble_sample.rs:107-190)
 * with the reference's EXECUTION structure, because that -- not the arithmetic -- is what
 * sets its speed: every walker is its own heap vector
 * (Vec<LsVec<f64,Vec<f64>>>), pairing / z / flags / proposals / accept run on the calling
 * thread, scale_vec allocates three temporaries per walker (src/mcmc/utils.rs:55 via
 * src/linear_space/type_wrapper.rs:36-43,66-72), and only the log-density map (:161) fans
 * out over the host's threads (rayon global pool -> OpenMP here).  Results are identical
 * to orc_sample_pt for the same seed (checked in tests/).
 */
#define _GNU_SOURCE
#include "scorus_oracle.h"

#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int orc_max_threads(void)
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c > 0 ? (int)c : 1;
}

/* order-preserving parallel map over [0, n): worker threads pull chunks from an atomic
 * counter (stand-in for rayon's work-stealing global pool; pthreads so that it builds with
 * any gcc, with or without libgomp) */
typedef void (*range_fn)(size_t lo, size_t hi, void *ctx);
typedef struct { atomic_size_t next; size_t n, chunk; range_fn fn; void *ctx; } pf_job;
static void *pf_worker(void *arg)
{
    pf_job *j = (pf_job *)arg;
    for (;;) {
        size_t lo = atomic_fetch_add(&j->next, j->chunk);
        if (lo >= j->n) break;
        size_t hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
        j->fn(lo, hi, j->ctx);
    }
    return NULL;
}
static void parallel_for(size_t n, size_t chunk, int nthreads, range_fn fn, void *ctx)
{
    if (nthreads <= 1 || n <= chunk) { fn(0, n, ctx); return; }
    pf_job job;
    atomic_init(&job.next, 0);
    job.n = n; job.chunk = chunk; job.fn = fn; job.ctx = ctx;
    pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
}

typedef struct {
    int kind; const double *params; size_t nparams, dim;
    double **rows; double *out;
} lp_rows_ctx;
static void lp_rows_range(size_t lo, size_t hi, void *c)
{
    lp_rows_ctx *k = (lp_rows_ctx *)c;
    for (size_t i = lo; i < hi; ++i) k->out[i] = orc_logprob(k->kind, k->params, k->nparams, k->rows[i], k->dim);
}

typedef struct { size_t a, b; } pair_t;
static int cmp_pair(const void *x, const void *y)
{
    const pair_t *p = (const pair_t *)x, *q = (const pair_t *)y;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;
    return p->b < q->b ? -1 : (p->b > q->b);
}

/* &x * s : clone then scale in place (type_wrapper.rs:66-72) */
static double *vec_scaled(const double *x, size_t d, double s)
{
    double *r = (double *)malloc(d * sizeof(double));
    memcpy(r, x, d * sizeof(double));
    for (size_t i = 0; i < d; ++i) r[i] = r[i] * s;
    return r;
}
/* &a + &b : clone a then add b (type_wrapper.rs:36-43) */
static double *vec_added(const double *a, const double *b, size_t d)
{
    double *r = (double *)malloc(d * sizeof(double));
    memcpy(r, a, d * sizeof(double));
    for (size_t i = 0; i < d; ++i) r[i] = r[i] + b[i];
    return r;
}

double orc_ref_structure_run(int kind, const double *params, size_t nparams, double *ens,
                             double *lp, size_t n, size_t dim, const double *beta, size_t nbeta,
                             double a, int ufs_kind, double prob, uint64_t seed, size_t nsteps,
                             int nthreads, double *stage_seconds)
{
    if (nbeta == 0 || n == 0 || n % nbeta || (n / nbeta) % 2) return -1.0;
    size_t npb = n / nbeta;
    double stage[6] = {0, 0, 0, 0, 0, 0};
    orc_rng g;
    orc_rng_seed(&g, seed);
    uint64_t onehot = 0;
    if (nthreads <= 0) nthreads = orc_max_threads();

    /* caller-side state: one heap vector per walker */
    double **rows = (double **)malloc(n * sizeof(double *));
    for (size_t i = 0; i < n; ++i) {
        rows[i] = (double *)malloc(dim * sizeof(double));
        memcpy(rows[i], ens + i * dim, dim * sizeof(double));
    }

    double t_all = now_s();
    for (size_t step = 0; step < nsteps; ++step) {
        double t0 = now_s();
        /* :135-148 pairing: per-rung index vector, shuffle, chunk into both orientations, sort */
        pair_t *pair_id = (pair_t *)malloc(n * sizeof(pair_t));
        size_t np = 0;
        for (size_t r = 0; r < nbeta; ++r) {
            size_t *b = (size_t *)malloc(npb * sizeof(size_t));
            for (size_t k = 0; k < npb; ++k) b[k] = k + r * npb;
            for (size_t i = npb; i-- > 1;) {
                uint32_t j = orc_rng_below(&g, (uint32_t)(i + 1));
                size_t t = b[i]; b[i] = b[j]; b[j] = t;
            }
            for (size_t k = 0; k < npb; k += 2) { pair_id[np].a = b[k]; pair_id[np].b = b[k + 1]; ++np; }
            for (size_t k = 0; k < npb; k += 2) { pair_id[np].a = b[k + 1]; pair_id[np].b = b[k]; ++np; }
            free(b);
        }
        qsort(pair_id, n, sizeof(pair_t), cmp_pair);
        double t1 = now_s();

        /* :149 z_list */
        double *z_list = (double *)malloc(n * sizeof(double));
        double range = orc_z_range(a);
        for (size_t i = 0; i < n; ++i) z_list[i] = orc_z_from_p(orc_rng_uniform(&g) * range, a);
        double t2 = now_s();

        /* :150-153 flags: one Vec<bool> per walker */
        uint8_t **flags = (uint8_t **)malloc(n * sizeof(uint8_t *));
        for (size_t i = 0; i < n; ++i) {
            uint8_t *f = NULL;
            if (ufs_kind == ORC_UFS_PROB) {
                for (;;) {
                    f = (uint8_t *)malloc(dim);
                    int any = 0;
                    for (size_t k = 0; k < dim; ++k) { f[k] = orc_rng_uniform(&g) < prob; any |= f[k]; }
                    if (any) break;
                    free(f);
                }
            } else if (ufs_kind == ORC_UFS_ALL) {
                f = (uint8_t *)malloc(dim);
                for (size_t k = 0; k < dim; ++k) f[k] = 1;
            } else {
                f = (uint8_t *)malloc(dim);
                onehot = (onehot + 1) % dim;
                for (size_t k = 0; k < dim; ++k) f[k] = (k == onehot);
            }
            flags[i] = f;
        }
        double t3 = now_s();

        /* :155-159 proposals, three temporaries each (utils.rs:55) */
        double **prop = (double **)malloc(n * sizeof(double *));
        for (size_t i = 0; i < n; ++i) {
            double z = z_list[i];
            double *s1 = vec_scaled(rows[pair_id[i].a], dim, z);
            double *s2 = vec_scaled(rows[pair_id[i].b], dim, 1.0 - z);
            double *y = vec_added(s1, s2, dim);
            free(s1);
            free(s2);
            const double *p1 = rows[pair_id[i].a];
            for (size_t k = 0; k < dim; ++k)
                if (!flags[i][k]) y[k] = p1[k];
            prop[i] = y;
        }
        double t4 = now_s();

        /* :161 the only parallel region: order-preserving map over all host threads */
        double *new_lp = (double *)malloc(n * sizeof(double));
        lp_rows_ctx lc = {kind, params, nparams, dim, prop, new_lp};
        parallel_for(n, 256, nthreads, lp_rows_range, &lc);
        double t5 = now_s();

        /* :164-189 nphi + serial accept; an accepted proposal is moved in (pointer swap) */
        size_t *nphi = (size_t *)malloc(n * sizeof(size_t));
        for (size_t i = 0; i < n; ++i) {
            size_t c = 0;
            for (size_t k = 0; k < dim; ++k) c += flags[i][k];
            nphi[i] = c;
        }
        for (size_t i = 0; i < n; ++i) {
            double b = beta[i / npb];
            double delta_lp = new_lp[i] - lp[i];
            double q = exp((double)(nphi[i] - 1) * log(z_list[i]) + delta_lp * b);
            if (orc_rng_uniform(&g) < q) {
                free(rows[i]);
                rows[i] = prop[i];
                lp[i] = new_lp[i];
            } else {
                free(prop[i]);
            }
        }
        double t6 = now_s();

        for (size_t i = 0; i < n; ++i) free(flags[i]);
        free(flags); free(prop); free(new_lp); free(nphi); free(z_list); free(pair_id);
        stage[0] += t1 - t0; stage[1] += t2 - t1; stage[2] += t3 - t2;
        stage[3] += t4 - t3; stage[4] += t5 - t4; stage[5] += t6 - t5;
    }
    double total = now_s() - t_all;

    for (size_t i = 0; i < n; ++i) {
        memcpy(ens + i * dim, rows[i], dim * sizeof(double));
        free(rows[i]);
    }
    free(rows);
    if (stage_seconds) memcpy(stage_seconds, stage, sizeof(stage));
    return total;
}

typedef struct {
    int kind; const double *params; size_t nparams, dim, npb;
    double *ens, *lp; const double *beta; const uint32_t *partner; const double *z, *u;
    const uint8_t *flags; double *prop, *nlp;
} flat_ctx;
static void flat_propose_range(size_t lo, size_t hi, void *c)
{
    flat_ctx *k = (flat_ctx *)c;
    size_t dim = k->dim;
    for (size_t i = lo; i < hi; ++i) {
        orc_propose_move(k->ens + i * dim, k->ens + (size_t)k->partner[i] * dim, k->z[i],
                         k->flags + i * dim, dim, dim, k->prop + i * dim);
        k->nlp[i] = orc_logprob(k->kind, k->params, k->nparams, k->prop + i * dim, dim);
    }
}
static void flat_accept_range(size_t lo, size_t hi, void *c)
{
    flat_ctx *k = (flat_ctx *)c;
    size_t dim = k->dim;
    for (size_t i = lo; i < hi; ++i) {
        size_t cnt = 0;
        for (size_t d = 0; d < dim; ++d) cnt += k->flags[i * dim + d];
        double q = exp((double)(cnt - 1) * log(k->z[i]) + (k->nlp[i] - k->lp[i]) * k->beta[i / k->npb]);
        if (k->u[i] < q) {
            memcpy(k->ens + i * dim, k->prop + i * dim, dim * sizeof(double));
            k->lp[i] = k->nlp[i];
        }
    }
}

typedef struct {
    uint64_t seed, step, onehot; size_t n, dim, nbeta; double a; int ufs_kind; double prob;
    uint32_t *partner; double *z, *u; uint8_t *flags; int rc;
} stream_ctx;
static void stream_partner_range(size_t lo, size_t hi, void *c)
{
    stream_ctx *k = (stream_ctx *)c;  /* lo, hi count PAIRS */
    if (orc_philox_partner_range(k->seed, k->step, k->n, k->nbeta, 2 * lo, 2 * hi, k->partner)) k->rc = 1;
}
static void stream_walker_range(size_t lo, size_t hi, void *c)
{
    stream_ctx *k = (stream_ctx *)c;
    if (orc_philox_walker_range(k->seed, k->step, k->dim, k->a, k->ufs_kind, k->prob, k->onehot, lo, hi, k->z, k->flags, k->u))
        k->rc = 1;
}

double orc_flat_threaded_run(int kind, const double *params, size_t nparams, double *ens,
                             double *lp, size_t n, size_t dim, const double *beta, size_t nbeta,
                             double a, int ufs_kind, double prob, uint64_t seed,
                             uint64_t first_step, size_t nsteps, int nthreads)
{
    if (nbeta == 0 || n == 0 || n % nbeta || (n / nbeta) % 2) return -1.0;
    size_t npb = n / nbeta;
    if (nthreads <= 0) nthreads = orc_max_threads();
    uint32_t *partner = (uint32_t *)malloc(n * sizeof(uint32_t));
    double *z = (double *)malloc(n * sizeof(double));
    double *u = (double *)malloc(n * sizeof(double));
    uint8_t *flags = (uint8_t *)malloc(n * dim);
    double *prop = (double *)malloc(n * dim * sizeof(double));
    double *nlp = (double *)malloc(n * sizeof(double));
    uint64_t onehot = 0;
    double t0 = now_s();
    for (size_t s = 0; s < nsteps; ++s) {
        /* the streams are counter-based: generated on all threads too (large rungs; rungs of up to 1024 walkers are
         * shuffled exactly by a sort per rung and stay serial) */
        if (npb > 1024) {
            stream_ctx sc = {seed, first_step + s, onehot, n, dim, nbeta, a, ufs_kind, prob, partner, z, u, flags, 0};
            parallel_for(n / 2, 4096, nthreads, stream_partner_range, &sc);
            parallel_for(n, 1024, nthreads, stream_walker_range, &sc);
            if (sc.rc) { t0 = now_s() + 1.0; break; }
        } else if (orc_philox_sample_streams(seed, first_step + s, n, dim, nbeta, a, ufs_kind, prob, onehot,
                                             partner, z, flags, u)) { t0 = now_s() + 1.0; break; }
        onehot = (onehot + n) % dim;
        flat_ctx fc = {kind, params, nparams, dim, npb, ens, lp, beta, partner, z, u, flags, prop, nlp};
        parallel_for(n, 1024, nthreads, flat_propose_range, &fc);
        parallel_for(n, 1024, nthreads, flat_accept_range, &fc);
    }
    double total = now_s() - t0;
    free(partner); free(z); free(u); free(flags); free(prop); free(nlp);
    return total;
}
