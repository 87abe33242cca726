/* hmc.c -- TEST INFRASTRUCTURE ONLY (see scorus_oracle.h).  CPU restatement of the reference's ensemble /
 * parallel-tempering Hamiltonian Monte Carlo: src/mcmc/hmc/utils.rs:6-30 (leapfrog),
 * src/mcmc/hmc/naive.rs:17-56 (HmcParam), :122-158 (sample_ensemble_pt), :160-292 (sample_ensemble_pt_impl);
 * the single-chain sample (:58-120) is the same arithmetic for one walker at beta = 1.
 * PARITY UNPINNED: the reference has no test or golden vector for this path and cannot be compiled here.
 *
 * The reference takes the gradient as a second user closure (grad_logprob: &G).  Here it is a property of the
 * log-density plugin; the formulas and their operation order are stated below (the Rosenbrock one follows the
 * reference's own example closure, examples/test_hmc.rs:27-37). */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "scorus_oracle.h"

/* gradient of the plugin's log-density; returns 0, or ORC_ERR_INVALID_ARG for a plugin without one */
int orc_grad_logprob(int kind, const double *params, size_t nparams, const double *x, size_t dim, double *g)
{
    switch (kind) {
    case ORC_LP_GAUSSIAN: { /* lp = -sum c x^2: g = (-2 c) x */
        const double c = nparams >= 1 ? params[0] : 0.5;
        const double k = -2.0 * c;
        for (size_t j = 0; j < dim; ++j) g[j] = k * x[j];
        return ORC_OK;
    }
    case ORC_LP_BOUNDED_GAUSSIAN: /* lp = -sum x^2 inside the box (-inf outside rejects the trajectory) */
        for (size_t j = 0; j < dim; ++j) g[j] = -2.0 * x[j];
        return ORC_OK;
    case ORC_LP_ROSENBROCK:
        /* examples/test_hmc.rs:27-37, keeping only the i that contribute (the others subtract 0.0):
         * g[j] = 0; i = j-1: g[j] -= 200 (x[j] - x[j-1]^2) * 1;  i = j: g[j] -= 200 (x[j+1] - x[j]^2) (0 - 2 x[j]) + 2 (x[j] - 1) */
        for (size_t j = 0; j < dim; ++j) {
            double r = 0.0;
            if (j >= 1) r -= (200.0 * (x[j] - x[j - 1] * x[j - 1])) * 1.0;
            if (j + 1 < dim) r -= (200.0 * (x[j + 1] - x[j] * x[j])) * (0.0 - 2.0 * x[j]) + 2.0 * (x[j] - 1.0);
            g[j] = r;
        }
        return ORC_OK;
    case ORC_LP_RASTRIGIN: { /* lp = -n a + sum [a cos(2 pi x) - x^2]: g = -(2 pi a) sin(2 pi x) - 2 x */
        const double a = nparams >= 1 ? params[0] : 10.0;
        const double two_pi = 2.0 * 3.14159265358979323846;
        const double k = two_pi * a;
        for (size_t j = 0; j < dim; ++j) {
            const double s = k * sin(two_pi * x[j]);
            g[j] = (0.0 - s) - 2.0 * x[j];
        }
        return ORC_OK;
    }
    case ORC_LP_MIXTURE: { /* logsumexp of two isotropic Gaussians at +m, -m: w1 (-(x-m)/s1^2) + w2 (-(x+m)/s2^2) */
        if (nparams < 2 + dim) return ORC_ERR_INVALID_ARG;
        const double s1 = params[0], s2 = params[1];
        const double *m = params + 2;
        double q1 = 0.0, q2 = 0.0;
        for (size_t i = 0; i < dim; ++i) {
            const double a = x[i] - m[i], b = x[i] + m[i];
            q1 += a * a;
            q2 += b * b;
        }
        const double t1 = (-q1) / ((2.0 * s1) * s1) - (double)dim * log(s1);
        const double t2 = (-q2) / ((2.0 * s2) * s2) - (double)dim * log(s2);
        const double lse = orc_logprob(kind, params, nparams, x, dim);
        const double r1 = exp(t1 - lse) / (s1 * s1), r2 = exp(t2 - lse) / (s2 * s2);
        for (size_t j = 0; j < dim; ++j) g[j] = (0.0 - r1 * (x[j] - m[j])) - r2 * (x[j] + m[j]);
        return ORC_OK;
    }
    case ORC_LP_CORR_GAUSSIAN: { /* builder-defined lp = -1/2 |L (x - mu)|^2: g = 0 - L^T (L (x - mu)), both
                                  * products summed left to right (the order scorus_b200/csrc/hmc_kernel.cuh keeps) */
        if (nparams < dim + dim * dim) return ORC_ERR_INVALID_ARG;
        const double *mu = params, *L = params + dim;
        double *v = (double *)malloc(dim * sizeof(double));
        for (size_t i = 0; i < dim; ++i) {
            double w = 0.0;
            for (size_t j = 0; j < dim; ++j) w += L[i * dim + j] * (x[j] - mu[j]);
            v[i] = w;
        }
        for (size_t k = 0; k < dim; ++k) {
            double s = 0.0;
            for (size_t i = 0; i < dim; ++i) s += L[i * dim + k] * v[i];
            g[k] = 0.0 - s;
        }
        free(v);
        return ORC_OK;
    }
    default:
        return ORC_ERR_INVALID_ARG;
    }
}

/* leapfrog, src/mcmc/hmc/utils.rs:17-29, with h_p(p) = p * m_inv (m_inv = 1) and h_q(q) = grad(q) * (-beta)
 * (naive.rs:224-225) */
static int leapfrog(int kind, const double *params, size_t nparams, double *q, double *p, double *lhq, double eps,
                    double beta, size_t dim, double *tmp)
{
    const double he = eps * 0.5; /* epsilon * half, :21 */
    for (size_t j = 0; j < dim; ++j) p[j] = p[j] - lhq[j] * he;          /* p_mid, :21 */
    for (size_t j = 0; j < dim; ++j) q[j] = q[j] + (p[j] * 1.0) * eps;   /* :23-25 */
    int rc = orc_grad_logprob(kind, params, nparams, q, dim, tmp);
    if (rc) return rc;
    for (size_t j = 0; j < dim; ++j) lhq[j] = tmp[j] * (0.0 - beta);      /* :27 */
    for (size_t j = 0; j < dim; ++j) p[j] = p[j] - lhq[j] * he;          /* :29 */
    return ORC_OK;
}

static double kinetic(const double *p, size_t dim)
{
    double s = 0.0; /* p.dot(p) / two * m_inv, naive.rs:194 */
    for (size_t j = 0; j < dim; ++j) s += p[j] * p[j];
    return s / 2.0 * 1.0;
}

/* sample_ensemble_pt (naive.rs:122-292) with the draws supplied: p0[n][dim] the momenta (:182-192), u1[n] the accept
 * uniform (:269, consumed only when dh is finite), u2[n] the adaptation uniform (:275 or :280).  q0, lp, epsilon
 * [nbeta] are updated in place; accepted[n] and accepted_cnt[nbeta] (out, may be NULL).  l leapfrog steps. */
int orc_hmc_ensemble_pt_fixed(int kind, const double *params, size_t nparams, double *q0, double *lp, size_t n,
                              size_t dim, double *epsilon, const double *beta, size_t nbeta, size_t l,
                              double target_accept_ratio, double adj_factor, const double *p0, const double *u1,
                              const double *u2, uint8_t *accepted, uint64_t *accepted_cnt)
{
    if (nbeta == 0 || n == 0) return ORC_ERR_INVALID_ARG;
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA; /* the index beta_list[i / n_per_beta] :222 would panic */
    double *q = (double *)malloc(dim * sizeof(double)), *p = (double *)malloc(dim * sizeof(double));
    double *lhq = (double *)malloc(dim * sizeof(double)), *tmp = (double *)malloc(dim * sizeof(double));
    double *qs = (double *)malloc(n * dim * sizeof(double)), *pu = (double *)malloc(n * sizeof(double));
    double *dh = (double *)malloc(n * sizeof(double));
    int rc = ORC_OK;
    if (accepted_cnt) memset(accepted_cnt, 0, nbeta * sizeof(uint64_t));
    /* trajectories: every walker from the state at entry (:197-229), then the energies (:231-253) */
    for (size_t i = 0; i < n && rc == ORC_OK; ++i) {
        const double b = beta[i / npb], e = epsilon[i / npb];
        memcpy(q, q0 + i * dim, dim * sizeof(double));
        memcpy(p, p0 + i * dim, dim * sizeof(double));
        const double k0 = kinetic(p, dim);
        rc = orc_grad_logprob(kind, params, nparams, q, dim, tmp); /* :139 */
        if (rc) break;
        for (size_t j = 0; j < dim; ++j) lhq[j] = tmp[j] * (-1.0 * b); /* :201-205 */
        for (size_t s = 0; s < l && rc == ORC_OK; ++s) rc = leapfrog(kind, params, nparams, q, p, lhq, e, b, dim, tmp);
        const double u0 = -lp[i];
        pu[i] = -orc_logprob(kind, params, nparams, q, dim);
        const double k1 = kinetic(p, dim);
        dh[i] = b * (u0 - pu[i]) + k0 - k1; /* :250 */
        memcpy(qs + i * dim, q, dim * sizeof(double));
    }
    /* accept loop, walkers in order (:255-290): epsilon of the rung adapts as it goes */
    for (size_t i = 0; i < n && rc == ORC_OK; ++i) {
        const size_t r = i / npb;
        int acc = 0;
        if (isfinite(dh[i]) && u1[i] < exp(dh[i])) {
            memcpy(q0 + i * dim, qs + i * dim, dim * sizeof(double));
            lp[i] = -pu[i];
            acc = 1;
            if (accepted_cnt) accepted_cnt[r] += 1;
            if (u2[i] < 1.0 - target_accept_ratio) epsilon[r] = epsilon[r] * (1.0 + adj_factor);
        } else if (u2[i] < target_accept_ratio) {
            epsilon[r] = epsilon[r] / (1.0 + adj_factor);
        }
        if (accepted) accepted[i] = (uint8_t)acc;
    }
    free(q); free(p); free(lhq); free(tmp); free(qs); free(pu); free(dh);
    return rc;
}

/* standard normal by Box-Muller from two uniforms of the oracle's generator (the reference draws rand_distr's
 * StandardNormal, a ziggurat: a different -- equally valid -- map from random bits to normals) */
static void normal_pair(orc_rng *g, double *a, double *b)
{
    double u = orc_rng_uniform(g), v = orc_rng_uniform(g);
    if (u < 1e-300) u = 1e-300;
    const double r = sqrt(-2.0 * log(u)), t = 2.0 * 3.14159265358979323846 * v;
    *a = r * cos(t);
    *b = r * sin(t);
}

/* RNG-driven wrapper in the reference's consumption order: all momenta first (walker-major, :182-192), then per
 * walker the accept uniform (only when dh is finite) and the adaptation uniform */
int orc_hmc_ensemble_pt(int kind, const double *params, size_t nparams, double *q0, double *lp, size_t n, size_t dim,
                        orc_rng *g, double *epsilon, const double *beta, size_t nbeta, size_t l,
                        double target_accept_ratio, double adj_factor, uint64_t *accepted_cnt)
{
    double *p0 = (double *)malloc((n * dim + 2) * sizeof(double));
    double *u1 = (double *)malloc((n + 1) * sizeof(double)), *u2 = (double *)malloc((n + 1) * sizeof(double));
    for (size_t k = 0; k < n * dim; k += 2) {
        double a, b;
        normal_pair(g, &a, &b);
        p0[k] = a;
        if (k + 1 < n * dim) p0[k + 1] = b;
    }
    /* dh does not depend on the accept draws, so both uniforms can be drawn ahead; a non-finite dh leaves u1 unused
     * (the reference would not have consumed it: an oracle-stream detail, not a property of the chain) */
    for (size_t i = 0; i < n; ++i) { u1[i] = orc_rng_uniform(g); u2[i] = orc_rng_uniform(g); }
    int rc = orc_hmc_ensemble_pt_fixed(kind, params, nparams, q0, lp, n, dim, epsilon, beta, nbeta, l, target_accept_ratio,
                                       adj_factor, p0, u1, u2, NULL, accepted_cnt);
    free(p0); free(u1); free(u2);
    return rc;
}
