/*
 * scorus_oracle.c -- CPU restatement of scorus's ensemble-sampler hot path.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see scorus_oracle.h).
 *
 * Layout: ensemble is flat walker-major double[n][dim]; rung r owns walkers
 * [r*npb, (r+1)*npb) as in src/mcmc/ensemble_sample.rs:137,181.
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile).
 */
#include "scorus_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* log-density plugins; sequential left-to-right sums  */
/* ------------------------------------------------------------------ */

static double lp_gaussian(const double *p, size_t np, const double *x, size_t d)
{
    /* examples/test_emcee.rs:18-24 (c = 1/2: `result -= i*i/2.0`);
     * examples/sample_high_dim.rs:22-28 (c = 1: `result -= i*i`).
     * (x*x)*0.5 == (x*x)/2.0 bit for bit, (x*x)*1.0 == x*x. */
    double c = np >= 1 ? p[0] : 0.5;
    double result = 0.0;
    for (size_t i = 0; i < d; ++i) {
        double sq = x[i] * x[i];
        result -= sq * c;
    }
    return result;
}

static double lp_bounded_gaussian(const double *p, size_t np, const double *x, size_t d)
{
    /* examples/sample_bimodal.rs:35-44: subtract first, then the early return */
    double limit = np >= 1 ? p[0] : 1.5;
    double result = 0.0;
    for (size_t i = 0; i < d; ++i) {
        result -= x[i] * x[i];
        if (fabs(x[i]) > limit) return -INFINITY;
    }
    return result;
}

static double lp_rosenbrock(const double *x, size_t d)
{
    /* examples/test_emcee.rs:26-32; src/main.rs:42-48.  powi(2) is x*x. */
    double result = 0.0;
    for (size_t i = 0; i + 1 < d; ++i) {
        double xi2 = x[i] * x[i];
        double t = x[i + 1] - xi2;
        double t2 = t * t;
        double s = 1.0 - x[i];
        double s2 = s * s;
        double term = 100.0 * t2 + s2;
        result += term;
    }
    return -result;
}

static double lp_rastrigin(const double *p, size_t np, const double *x, size_t d)
{
    /* examples/sample_rastrigin.rs:17-25 */
    double a = np >= 1 ? p[0] : 10.0;
    const double two_pi = 2.0 * 3.14159265358979323846; /* 2.0 * f64::PI() */
    double result = -((double)d * a);
    for (size_t i = 0; i < d; ++i) {
        double arg = two_pi * x[i];
        double c = a * cos(arg);
        double sq = x[i] * x[i];
        result += c - sq;
    }
    return result;
}

static double lp_bimodal2d(const double *p, size_t np, const double *x, size_t d)
{
    /* examples/sample_bimodal.rs:46-54; optional params override the constants:
     * {lo0, hi0, lo1, hi1, split, mu_a, sigma_a, mu_b, sigma_b} */
    double lo0 = -15.0, hi0 = 15.0, lo1 = 0.0, hi1 = 1.0, split = 0.5;
    double mu_a = -5.0, sg_a = 0.1, mu_b = 5.0, sg_b = 1.0;
    if (np >= 9) {
        lo0 = p[0]; hi0 = p[1]; lo1 = p[2]; hi1 = p[3]; split = p[4];
        mu_a = p[5]; sg_a = p[6]; mu_b = p[7]; sg_b = p[8];
    }
    if (d < 2) return NAN;
    if (x[0] < lo0 || x[0] > hi0 || x[1] < lo1 || x[1] > hi1) return -INFINITY;
    double mu = x[1] < split ? mu_a : mu_b;
    double sigma = x[1] < split ? sg_a : sg_b;
    double dx = x[0] - mu;
    double num = (-dx) * dx;
    double den = (2.0 * sigma) * sigma;
    return num / den - log(sigma);
}

static double lp_step2d(const double *x, size_t d)
{
    /* examples/sample_bimodal.rs:56-65 */
    if (d < 2) return NAN;
    if (x[0] < 0.0 || x[0] > 1.0 || x[1] < 0.0 || x[1] > 1.0) return -INFINITY;
    return x[0] > 0.5 ? log(0.1) : log(0.9);
}

static double lp_corr_gaussian(const double *p, size_t np, const double *x, size_t d)
{
    /* builder-defined: -0.5 * sum_i ( sum_j L[i][j]*(x[j]-mu[j]) )^2 */
    if (np < d + d * d) return NAN;
    const double *mu = p;
    const double *L = p + d;
    double s = 0.0;
    for (size_t i = 0; i < d; ++i) {
        double w = 0.0;
        for (size_t j = 0; j < d; ++j) {
            double c = x[j] - mu[j];
            w += L[i * d + j] * c;
        }
        s += w * w;
    }
    return -0.5 * s;
}

static double lp_mixture(const double *p, size_t np, const double *x, size_t d)
{
    /* builder-defined:
     * logsumexp( -|x-m|^2/(2 s1^2) - D ln s1 , -|x+m|^2/(2 s2^2) - D ln s2 ) */
    if (np < 2 + d) return NAN;
    double s1 = p[0], s2 = p[1];
    const double *m = p + 2;
    double q1 = 0.0, q2 = 0.0;
    for (size_t i = 0; i < d; ++i) {
        double a = x[i] - m[i];
        double b = x[i] + m[i];
        q1 += a * a;
        q2 += b * b;
    }
    double t1 = (-q1) / ((2.0 * s1) * s1) - (double)d * log(s1);
    double t2 = (-q2) / ((2.0 * s2) * s2) - (double)d * log(s2);
    double hi = t1 > t2 ? t1 : t2;
    double lo = t1 > t2 ? t2 : t1;
    if (hi == -INFINITY) return -INFINITY;
    return hi + log1p(exp(lo - hi));
}

double orc_logprob(int kind, const double *params, size_t nparams, const double *x, size_t dim)
{
    switch (kind) {
    case ORC_LP_GAUSSIAN: return lp_gaussian(params, nparams, x, dim);
    case ORC_LP_BOUNDED_GAUSSIAN: return lp_bounded_gaussian(params, nparams, x, dim);
    case ORC_LP_ROSENBROCK: return lp_rosenbrock(x, dim);
    case ORC_LP_RASTRIGIN: return lp_rastrigin(params, nparams, x, dim);
    case ORC_LP_BIMODAL2D: return lp_bimodal2d(params, nparams, x, dim);
    case ORC_LP_STEP2D: return lp_step2d(x, dim);
    case ORC_LP_CORR_GAUSSIAN: return lp_corr_gaussian(params, nparams, x, dim);
    case ORC_LP_MIXTURE: return lp_mixture(params, nparams, x, dim);
    default: return NAN;
    }
}

/* ------------------------------------------------------------------ */
/* move primitives                                                     */
/* ------------------------------------------------------------------ */

double orc_z_range(double a)
{
    /* src/mcmc/utils.rs:39-42: upper end of the uniform, two*(sqrt_a - unit/sqrt_a) */
    double sqrt_a = sqrt(a);
    return 2.0 * (sqrt_a - 1.0 / sqrt_a);
}

double orc_z_from_p(double p, double a)
{
    /* src/mcmc/utils.rs:43-44 */
    double sqrt_a = sqrt(a);
    double y = 1.0 / sqrt_a + p / 2.0;
    return y * y;
}

void orc_propose_move(const double *p1, const double *p2, double z, const uint8_t *flags,
                      size_t flag_len, size_t dim, double *out)
{
    /* scale_vec, src/mcmc/utils.rs:55: &(x1 * z) + &(x2 * (1 - z)), each op rounded
     * (src/linear_space/type_wrapper.rs:36-43,66-72) */
    double omz = 1.0 - z;
    for (size_t i = 0; i < dim; ++i) {
        double t1 = p1[i] * z;
        double t2 = p2[i] * omz;
        out[i] = t1 + t2;
    }
    /* propose_move, src/mcmc/ensemble_sample.rs:67-73: loop runs to update_flags.len() */
    if (flags) {
        for (size_t i = 0; i < flag_len && i < dim; ++i)
            if (!flags[i]) out[i] = p1[i];
    }
}

double orc_exchange_prob(double lp1, double lp2, double beta1, double beta2)
{
    /* src/mcmc/utils.rs:63-69 */
    double x = exp((beta2 - beta1) * (-lp2 + lp1));
    if (x > 1.0) return 1.0;
    return x;
}

void orc_init_logprob(int kind, const double *params, size_t nparams, const double *ens,
                      size_t n, size_t dim, double *lp)
{
    /* src/mcmc/ensemble_sample.rs:83-84 (order-preserving map) */
    for (size_t i = 0; i < n; ++i) lp[i] = orc_logprob(kind, params, nparams, ens + i * dim, dim);
}

static int check_sample_args(size_t n, size_t lp_len, size_t nbeta)
{
    /* src/mcmc/ensemble_sample.rs:127-133, in assert order */
    if (nbeta == 0) return ORC_ERR_INVALID_ARG;
    size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA;
    if (n == 0) return ORC_ERR_NWALKERS_IS_ZERO;
    if (npb % 2 != 0) return ORC_ERR_NWALKERS_IS_NOT_EVEN;
    if (n != lp_len) return ORC_ERR_LENGTH_MISMATCH;
    return ORC_OK;
}

int orc_sample_pt_fixed(int kind, const double *params, size_t nparams, double *ens, double *lp,
                        size_t n, size_t lp_len, size_t dim, const double *beta, size_t nbeta,
                        const uint32_t *partner, const double *z, const uint8_t *flags,
                        size_t flag_len, const double *u, uint8_t *accepted, double *proposed,
                        double *new_lp_out)
{
    int rc = check_sample_args(n, lp_len, nbeta);
    if (rc) return rc;
    if (flag_len > dim) return ORC_ERR_INVALID_ARG; /* reference panics on index, :70 */
    size_t npb = n / nbeta;

    /* the pairing of :135-148 is a fixed-point-free involution inside each rung */
    for (size_t i = 0; i < n; ++i) {
        size_t p = partner[i];
        if (p >= n || p == i || p / npb != i / npb || partner[p] != i) return ORC_ERR_INVALID_ARG;
    }
    /* nphi - 1 is usize arithmetic (:184): nphi == 0 is an error */
    size_t *nphi = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) {
        size_t c = 0;
        for (size_t k = 0; k < flag_len; ++k) c += flags[i * flag_len + k] ? 1 : 0;
        nphi[i] = c;
        if (c == 0) { free(nphi); return ORC_ERR_NPHI_ZERO; }
    }

    /* :155-159 -- every proposal reads the pre-step ensemble */
    double *prop = (double *)malloc(n * dim * sizeof(double));
    double *nlp = (double *)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; ++i)
        orc_propose_move(ens + i * dim, ens + (size_t)partner[i] * dim, z[i],
                         flags + i * flag_len, flag_len, dim, prop + i * dim);
    /* :161 */
    for (size_t i = 0; i < n; ++i) nlp[i] = orc_logprob(kind, params, nparams, prop + i * dim, dim);

    /* :169-189 */
    for (size_t i = 0; i < n; ++i) {
        size_t ibeta = i / npb;
        double b = beta[ibeta];
        double delta_lp = nlp[i] - lp[i];
        double q = exp((double)(nphi[i] - 1) * log(z[i]) + delta_lp * b);
        int acc = u[i] < q;
        if (acc) {
            memcpy(ens + i * dim, prop + i * dim, dim * sizeof(double));
            lp[i] = nlp[i];
        }
        if (accepted) accepted[i] = (uint8_t)acc;
    }
    if (proposed) memcpy(proposed, prop, n * dim * sizeof(double));
    if (new_lp_out) memcpy(new_lp_out, nlp, n * sizeof(double));
    free(prop);
    free(nlp);
    free(nphi);
    return ORC_OK;
}

int orc_swap_walkers_fixed(double *ens, double *lp, size_t n, size_t lp_len, size_t dim,
                           const double *beta, size_t nbeta, const uint32_t *jvec,
                           const double *r, uint8_t *swapped)
{
    if (nbeta == 0) return ORC_ERR_INVALID_ARG;
    size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA; /* src/mcmc/utils.rs:86 */
    if (n != lp_len) return ORC_OK;                                 /* silent no-op, :88 */
    /* :93-95 panics mid-sweep; checked up front here so a failed call changes nothing */
    for (size_t i = 1; i < nbeta; ++i)
        if (beta[i] > beta[i - 1]) return ORC_ERR_BETA_NOT_IN_DECR_ORD;
    for (size_t s = 0; s + 1 < nbeta; ++s)
        for (size_t k = 0; k < npb; ++k)
            if (jvec[s * npb + k] >= npb) return ORC_ERR_INVALID_ARG;

    double *tmp = (double *)malloc(dim * sizeof(double));
    for (size_t s = 0; s + 1 < nbeta; ++s) {
        size_t i = nbeta - 1 - s; /* :89 for i in (1..nbeta).rev() */
        double beta1 = beta[i], beta2 = beta[i - 1];
        for (size_t j2 = 0; j2 < npb; ++j2) { /* :99 */
            size_t j1 = jvec[s * npb + j2];
            size_t hot = i * npb + j1, cold = (i - 1) * npb + j2;
            double ep = orc_exchange_prob(lp[hot], lp[cold], beta1, beta2);
            int sw = r[s * npb + j2] < ep; /* :104-105 */
            if (sw) {
                memcpy(tmp, ens + hot * dim, dim * sizeof(double));
                memcpy(ens + hot * dim, ens + cold * dim, dim * sizeof(double));
                memcpy(ens + cold * dim, tmp, dim * sizeof(double));
                double t = lp[hot]; lp[hot] = lp[cold]; lp[cold] = t;
            }
            if (swapped) swapped[s * npb + j2] = (uint8_t)sw;
        }
    }
    free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* the oracle's own seeded generator + reference-order stream drawing  */
/* ------------------------------------------------------------------ */

static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void orc_rng_seed(orc_rng *g, uint64_t seed)
{
    for (int i = 0; i < 4; ++i) g->s[i] = splitmix64(&seed);
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static uint64_t rng_next(orc_rng *g)
{
    uint64_t *s = g->s;
    uint64_t result = rotl64(s[1] * 5, 7) * 9;
    uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

double orc_rng_uniform(orc_rng *g) { return (double)(rng_next(g) >> 11) * 0x1.0p-53; }

uint32_t orc_rng_below(orc_rng *g, uint32_t n)
{
    /* unbiased: reject the partial top bucket */
    uint64_t limit = (0x100000000ull / n) * n;
    for (;;) {
        uint64_t v = rng_next(g) >> 32;
        if (v < limit) return (uint32_t)(v % n);
    }
}

static void shuffle_u32(orc_rng *g, uint32_t *b, size_t len)
{
    /* rand 0.8 SliceRandom::shuffle: Fisher-Yates from the top */
    for (size_t i = len; i-- > 1;) {
        uint32_t j = orc_rng_below(g, (uint32_t)(i + 1));
        uint32_t t = b[i]; b[i] = b[j]; b[j] = t;
    }
}

int orc_draw_sample_streams(orc_rng *g, size_t n, size_t dim, size_t nbeta, double a,
                            int ufs_kind, double prob, uint64_t *onehot_counter,
                            uint32_t *partner, double *z, uint8_t *flags, double *u)
{
    int rc = check_sample_args(n, n, nbeta);
    if (rc) return rc;
    size_t npb = n / nbeta;
    uint32_t *b = (uint32_t *)malloc(npb * sizeof(uint32_t));
    /* src/mcmc/ensemble_sample.rs:135-148: shuffle each rung, pair neighbours, both orientations */
    for (size_t r = 0; r < nbeta; ++r) {
        for (size_t k = 0; k < npb; ++k) b[k] = (uint32_t)(k + r * npb);
        shuffle_u32(g, b, npb);
        for (size_t k = 0; k < npb; k += 2) {
            partner[b[k]] = b[k + 1];
            partner[b[k + 1]] = b[k];
        }
    }
    free(b);
    /* :149 */
    double range = orc_z_range(a);
    for (size_t i = 0; i < n; ++i) z[i] = orc_z_from_p(orc_rng_uniform(g) * range, a);
    /* :150-153 -> :38-54 */
    for (size_t i = 0; i < n; ++i) {
        uint8_t *f = flags + i * dim;
        if (ufs_kind == ORC_UFS_PROB) {
            for (;;) {
                int any = 0;
                for (size_t k = 0; k < dim; ++k) {
                    f[k] = orc_rng_uniform(g) < prob;
                    any |= f[k];
                }
                if (any) break;
            }
        } else if (ufs_kind == ORC_UFS_ALL) {
            for (size_t k = 0; k < dim; ++k) f[k] = 1;
        } else if (ufs_kind == ORC_UFS_ONE_HOT_CYCLE) {
            /* the cycling closure of examples/sample_high_dim.rs:59-65, one call per walker */
            *onehot_counter = (*onehot_counter + 1) % dim;
            for (size_t k = 0; k < dim; ++k) f[k] = (k == *onehot_counter);
        } else {
            return ORC_ERR_INVALID_ARG;
        }
    }
    /* :185 -- one uniform per walker, in order, unconditionally */
    for (size_t i = 0; i < n; ++i) u[i] = orc_rng_uniform(g);
    return ORC_OK;
}

int orc_draw_swap_streams(orc_rng *g, size_t n, size_t nbeta, uint32_t *jvec, double *r)
{
    if (nbeta == 0) return ORC_ERR_INVALID_ARG;
    size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA;
    uint32_t *jv = (uint32_t *)malloc((npb ? npb : 1) * sizeof(uint32_t));
    for (size_t k = 0; k < npb; ++k) jv[k] = (uint32_t)k; /* src/mcmc/utils.rs:87 */
    for (size_t s = 0; s + 1 < nbeta; ++s) {
        shuffle_u32(g, jv, npb); /* :97 -- cumulative, the vector is never reset */
        memcpy(jvec + s * npb, jv, npb * sizeof(uint32_t));
        for (size_t k = 0; k < npb; ++k) r[s * npb + k] = orc_rng_uniform(g); /* :104 */
    }
    free(jv);
    return ORC_OK;
}

int orc_sample_pt(int kind, const double *params, size_t nparams, double *ens, double *lp,
                  size_t n, size_t dim, orc_rng *g, double a, int ufs_kind, double prob,
                  uint64_t *onehot_counter, const double *beta, size_t nbeta)
{
    int rc = check_sample_args(n, n, nbeta);
    if (rc) return rc;
    uint32_t *partner = (uint32_t *)malloc(n * sizeof(uint32_t));
    double *z = (double *)malloc(n * sizeof(double));
    double *u = (double *)malloc(n * sizeof(double));
    uint8_t *flags = (uint8_t *)malloc(n * dim);
    rc = orc_draw_sample_streams(g, n, dim, nbeta, a, ufs_kind, prob, onehot_counter, partner, z,
                                 flags, u);
    if (!rc)
        rc = orc_sample_pt_fixed(kind, params, nparams, ens, lp, n, n, dim, beta, nbeta, partner,
                                 z, flags, dim, u, NULL, NULL, NULL);
    free(partner); free(z); free(u); free(flags);
    return rc;
}

int orc_swap_walkers(double *ens, double *lp, size_t n, size_t dim, orc_rng *g,
                     const double *beta, size_t nbeta)
{
    if (nbeta == 0) return ORC_ERR_INVALID_ARG;
    size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA;
    size_t m = (nbeta - 1) * npb;
    uint32_t *jvec = (uint32_t *)malloc((m ? m : 1) * sizeof(uint32_t));
    double *r = (double *)malloc((m ? m : 1) * sizeof(double));
    int rc = orc_draw_swap_streams(g, n, nbeta, jvec, r);
    if (!rc) rc = orc_swap_walkers_fixed(ens, lp, n, n, dim, beta, nbeta, jvec, r, NULL);
    free(jvec); free(r);
    return rc;
}

/* ------------------------------------------------------------------ */
/* moments and init                                                    */
/* ------------------------------------------------------------------ */

void orc_mean(const double *ens, size_t n, size_t dim, double *mean_out)
{
    /* src/linear_space/utils.rs:12-14: fold from x[0], then multiply by 1/N */
    for (size_t k = 0; k < dim; ++k) mean_out[k] = ens[k];
    for (size_t i = 1; i < n; ++i)
        for (size_t k = 0; k < dim; ++k) mean_out[k] = mean_out[k] + ens[i * dim + k];
    double inv = 1.0 / (double)n;
    for (size_t k = 0; k < dim; ++k) mean_out[k] = mean_out[k] * inv;
}

void orc_cov(const double *ens, size_t n, size_t dim, double *cov_out)
{
    /* src/linear_space/utils.rs:25-40: biased (/N) covariance about the mean */
    double *xm = (double *)malloc(dim * sizeof(double));
    double *c = (double *)malloc(dim * sizeof(double));
    orc_mean(ens, n, dim, xm);
    for (size_t k = 0; k < dim * dim; ++k) cov_out[k] = 0.0;
    for (size_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < dim; ++k) c[k] = ens[i * dim + k] - xm[k];
        for (size_t a = 0; a < dim; ++a)
            for (size_t b = 0; b < dim; ++b) cov_out[a * dim + b] = cov_out[a * dim + b] + c[a] * c[b];
    }
    for (size_t k = 0; k < dim * dim; ++k) cov_out[k] = cov_out[k] / (double)n;
    free(xm);
    free(c);
}

void orc_init_realization(orc_rng *g, const double *lo, const double *hi, size_t dim, double *out)
{
    /* src/mcmc/init_ensemble.rs:28-30: U[lo, hi) per coordinate */
    for (size_t k = 0; k < dim; ++k) out[k] = lo[k] + orc_rng_uniform(g) * (hi[k] - lo[k]);
}
