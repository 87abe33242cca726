/* twalk.c -- TEST INFRASTRUCTURE ONLY (see scorus_oracle.h).  CPU restatement of the reference's ensemble /
 * parallel-tempering t-walk, src/mcmc/twalk.rs: TWalkParams::new :89-110,
 * TWalkKernal::random :40-63, gen_update_flags :259-272, sim_walk :274-304, sim_b :306-322,
 * sim_traverse :324-353, calc_a :463-503, propose_move :505-532, sample :586-648, sample1 :650-762.
 * PARITY UNPINNED: the reference has no test or golden vector for this path and cannot be compiled here.
 *
 * One deliberate reading.  sample() builds each rung's pairs from indices 0..n/B (:605-612) and sample1()
 * uses them unshifted when it READS walkers and log-probs (:686-690, :741-749) but shifted by
 * ibeta * nwalkers_per_beta when it WRITES (:757-758): with B > 1 every rung would propose from rung 0's
 * walkers.  With B = 1 there is no shift and nothing to decide.  This restatement (and the device path)
 * applies the shift on both sides -- the pairs of rung r are walkers of rung r -- which is the only reading
 * under which rung r samples its own tempered density.  DESIGN.md section 4.7 records it.  The literal reading
 * is kept beside it (orc_twalk_half_fixed_literal) so that the difference is pinned by a test, not by prose:
 * parity with the reference AS WRITTEN holds for B = 1 only. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "scorus_oracle.h"

/* TWalkParams::new(n), twalk.rs:89-110: aw = 1.5, at = 6, pphi = min(n, 4) / n, fw = cumsum [0.5, 0.5] */
void orc_twalk_params_default(size_t dim, orc_twalk_params *out)
{
    const double n = (double)dim, n1phi = 4.0;
    out->aw = 1.5;
    out->at = 6.0;
    out->pphi = (n < n1phi ? n : n1phi) / n;
    out->fw[0] = 0.0 + 0.5;
    out->fw[1] = out->fw[0] + 0.5;
}

/* sim_b with its two uniforms supplied, twalk.rs:314-321 */
double orc_twalk_b_from_uniforms(double x, double v, double at)
{
    if (x == 0.0) return x;
    if (v < (at - 1.0) / (2.0 * at)) return pow(x, 1.0 / (at + 1.0));
    return pow(x, 1.0 / (1.0 - at));
}

/* One half-pass (sample1, twalk.rs:650-762) with the draws supplied per pair k = 0..n/2-1 in rung-major pair
 * order: pair k moves walker order[2k + half] against order[2k + 1 - half].
 *   kernel[k] 0 = Walk, 1 = Traverse;  b[k] (Traverse);  flags[k][dim];  walk_u[k][dim] (Walk, flagged
 *   coordinates only);  u[k] accept uniform.  accepted[k] (out, may be NULL), proposed [k][dim] / new_lp[k]
 *   (out, may be NULL). */
static int twalk_half_impl(int literal_index, int kind, const double *params, size_t nparams, double *ens, double *lp,
                           size_t n, size_t dim, const double *beta, size_t nbeta, const orc_twalk_params *tw,
                           const uint32_t *order, int half, const uint8_t *kernel, const double *b,
                           const uint8_t *flags, const double *walk_u, const double *u, uint8_t *accepted,
                           double *proposed, double *new_lp)
{
    (void)tw;
    if (nbeta == 0 || n == 0) return ORC_ERR_INVALID_ARG;
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA; /* assert :675 */
    if (npb % 2) return ORC_ERR_NWALKERS_IS_NOT_EVEN;               /* a.chunks(2) ... a[1] panics, :610 */
    const size_t npair = n / 2;
    double *y = (double *)malloc(npair * dim * sizeof(double));
    double *ylp = (double *)malloc(npair * sizeof(double));
    /* proposals of every pair from the ensemble as it stands (:676-688), then their log-densities (:693-732) */
    /* literal_index: the indices sample1 READS with are the rung-local ones of :605-612, i.e. walkers of rung 0 */
#define RD(idx) (literal_index ? (size_t)(idx) % npb : (size_t)(idx))
    for (size_t k = 0; k < npair; ++k) {
        const double *x = ens + RD(order[2 * k + half]) * dim;        /* the walker that moves */
        const double *xp = ens + RD(order[2 * k + 1 - half]) * dim;   /* its partner */
        const uint8_t *f = flags + k * dim;
        double *yk = y + k * dim;
        if (kernel[k] == 0) {
            /* sim_walk :287-301: z = aw/(1+aw) * (aw u^2 + 2u - 1) on flagged coordinates, else 0 */
            for (size_t i = 0; i < dim; ++i) {
                double z = 0.0;
                if (f[i]) {
                    const double uu = walk_u[k * dim + i];
                    z = (tw->aw / (1.0 + tw->aw)) * (tw->aw * (uu * uu) + 2.0 * uu - 1.0);
                }
                yk[i] = x[i] + (x[i] - xp[i]) * z;
            }
        } else {
            /* sim_traverse :342-350 */
            for (size_t i = 0; i < dim; ++i) yk[i] = f[i] ? xp[i] + b[k] * (xp[i] - x[i]) : x[i];
        }
        ylp[k] = orc_logprob(kind, params, nparams, yk, dim);
    }
    /* accept loop :736-760, pairs in order */
    for (size_t k = 0; k < npair; ++k) {
        const size_t i1 = order[2 * k + half];                 /* where the move is WRITTEN (:757-758: shifted) */
        const size_t r1 = RD(i1), i2 = RD(order[2 * k + 1 - half]); /* what calc_a READS (:741-749) */
        const double *yk = y + k * dim, *xpart = ens + i2 * dim;
        const uint8_t *f = flags + k * dim;
        size_t nphi = 0;
        for (size_t i = 0; i < dim; ++i) nphi += f[i] ? 1 : 0;
        int all_diff = 1; /* all_different(yp, x) :226-240, x = the partner */
        for (size_t i = 0; i < dim; ++i)
            if (yk[i] == xpart[i]) { all_diff = 0; break; }
        double a;
        const double bt = beta[i1 / npb];
        if (nphi == 0 || !all_diff) {
            a = 0.0; /* :482-483 */
        } else if (kernel[k] == 0) {
            a = exp((ylp[k] - lp[r1]) * bt); /* :486 */
        } else {
            a = exp((ylp[k] - lp[r1]) * bt + (double)((long)nphi - 2) * log(b[k])); /* :487-489 */
        }
        const int acc = u[k] < a; /* :756 */
        if (acc) {
            memcpy(ens + i1 * dim, yk, dim * sizeof(double));
            lp[i1] = ylp[k];
        }
        if (accepted) accepted[k] = (uint8_t)acc;
    }
    if (proposed) memcpy(proposed, y, npair * dim * sizeof(double));
    if (new_lp) memcpy(new_lp, ylp, npair * sizeof(double));
    free(y);
    free(ylp);
    return ORC_OK;
#undef RD
}

int orc_twalk_half_fixed(int kind, const double *params, size_t nparams, double *ens, double *lp, size_t n,
                         size_t dim, const double *beta, size_t nbeta, const orc_twalk_params *tw,
                         const uint32_t *order, int half, const uint8_t *kernel, const double *b,
                         const uint8_t *flags, const double *walk_u, const double *u, uint8_t *accepted,
                         double *proposed, double *new_lp)
{
    return twalk_half_impl(0, kind, params, nparams, ens, lp, n, dim, beta, nbeta, tw, order, half, kernel, b, flags,
                           walk_u, u, accepted, proposed, new_lp);
}

/* The same half-pass following the reference's index arithmetic LITERALLY: pairs are rung-local indices
 * (twalk.rs:605-612), sample1 reads walkers and log-probs with them unshifted (:686-690 propose_move,
 * :741-749 calc_a) -- rung 0's walkers, whatever the rung -- and writes shifted by ibeta * nwalkers_per_beta
 * (:757-758).  Identical to orc_twalk_half_fixed for one rung; for B > 1 every rung proposes from, and is
 * accepted against, rung 0 (tests/test_twalk_cpu.py pins where the two readings part).  `order` holds global
 * ids as above; the rung-local index is id % (n / nbeta). */
int orc_twalk_half_fixed_literal(int kind, const double *params, size_t nparams, double *ens, double *lp, size_t n,
                                 size_t dim, const double *beta, size_t nbeta, const orc_twalk_params *tw,
                                 const uint32_t *order, int half, const uint8_t *kernel, const double *b,
                                 const uint8_t *flags, const double *walk_u, const double *u, uint8_t *accepted,
                                 double *proposed, double *new_lp)
{
    return twalk_half_impl(1, kind, params, nparams, ens, lp, n, dim, beta, nbeta, tw, order, half, kernel, b, flags,
                           walk_u, u, accepted, proposed, new_lp);
}

/* the draws of propose_move (:519-531) for one pair, in the reference's consumption order */
static void draw_proposal(orc_rng *g, size_t dim, const orc_twalk_params *tw, uint8_t *kernel, double *b, uint8_t *f,
                          double *wu)
{
    const double uk = orc_rng_uniform(g) * tw->fw[1]; /* Uniform::new(0, fw[1]) :57 */
    *kernel = uk < tw->fw[0] ? 0 : 1;
    *b = 0.0;
    if (*kernel == 1) { /* sim_b :314-321: the second uniform is drawn only when the first is not 0 */
        const double x = orc_rng_uniform(g);
        *b = x == 0.0 ? x : orc_twalk_b_from_uniforms(x, orc_rng_uniform(g), tw->at);
    }
    for (;;) { /* gen_update_flags :264-271 */
        int any = 0;
        for (size_t i = 0; i < dim; ++i) {
            f[i] = orc_rng_uniform(g) < tw->pphi;
            any |= f[i];
        }
        if (any) break;
    }
    for (size_t i = 0; i < dim; ++i) wu[i] = 0.0;
    if (*kernel == 0)
        for (size_t i = 0; i < dim; ++i)
            if (f[i]) wu[i] = orc_rng_uniform(g); /* :291, flagged coordinates only */
}

/* The draws of one twalk::sample call (:586-648) in the reference's consumption order: the rung shuffles
 * (:605-612), then per half-pass every pair's proposal draws (:676-688) followed by every pair's accept uniform
 * (:756).  None depends on the state, so they can be drawn ahead.  order[n]; the rest [2][n/2](...[dim]). */
int orc_twalk_draw_streams(orc_rng *g, size_t n, size_t dim, size_t nbeta, const orc_twalk_params *tw, uint32_t *order,
                           uint8_t *kernel, double *b, uint8_t *flags, double *walk_u, double *u)
{
    if (nbeta == 0 || n == 0) return ORC_ERR_INVALID_ARG;
    const size_t npb = n / nbeta, npair = n / 2;
    if (npb * nbeta != n) return ORC_ERR_NWALKERS_MISMATCHES_NBETA;
    if (npb % 2) return ORC_ERR_NWALKERS_IS_NOT_EVEN;
    for (size_t r = 0; r < nbeta; ++r) {
        for (size_t k = 0; k < npb; ++k) order[r * npb + k] = (uint32_t)(r * npb + k);
        for (size_t i = npb; i-- > 1;) { /* SliceRandom::shuffle: Fisher-Yates from the top */
            const uint32_t j = orc_rng_below(g, (uint32_t)(i + 1));
            const uint32_t t = order[r * npb + i];
            order[r * npb + i] = order[r * npb + j];
            order[r * npb + j] = t;
        }
    }
    for (size_t half = 0; half < 2; ++half) {
        for (size_t k = 0; k < npair; ++k) {
            const size_t e = half * npair + k;
            draw_proposal(g, dim, tw, kernel + e, b + e, flags + e * dim, walk_u + e * dim);
        }
        for (size_t k = 0; k < npair; ++k) u[half * npair + k] = orc_rng_uniform(g);
    }
    return ORC_OK;
}

/* twalk::sample = the draws above + the two half-passes */
int orc_twalk_sample(int kind, const double *params, size_t nparams, double *ens, double *lp, size_t n, size_t dim,
                     orc_rng *g, const orc_twalk_params *tw, const double *beta, size_t nbeta)
{
    const size_t npair = n / 2;
    uint32_t *order = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint8_t *kernel = (uint8_t *)malloc(2 * npair + 1), *flags = (uint8_t *)malloc(2 * npair * dim + 1);
    double *b = (double *)malloc((2 * npair + 1) * sizeof(double)), *wu = (double *)malloc((2 * npair * dim + 1) * sizeof(double));
    double *u = (double *)malloc((2 * npair + 1) * sizeof(double));
    int rc = orc_twalk_draw_streams(g, n, dim, nbeta, tw, order, kernel, b, flags, wu, u);
    for (int half = 0; half < 2 && rc == ORC_OK; ++half)
        rc = orc_twalk_half_fixed(kind, params, nparams, ens, lp, n, dim, beta, nbeta, tw, order, half,
                                  kernel + half * npair, b + half * npair, flags + half * npair * dim,
                                  wu + half * npair * dim, u + half * npair, NULL, NULL, NULL);
    free(order); free(kernel); free(flags); free(b); free(wu); free(u);
    return rc;
}
