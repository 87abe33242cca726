/* chain_stats.c -- TEST INFRASTRUCTURE ONLY (see scorus_oracle.h).  CPU restatement of the device-side
 * chain statistics: the running per-rung moments built on the reference's
 * mean / cov definitions (src/linear_space/utils.rs:4-15, 17-41) and the integrated autocorrelation time of
 * the series the example drivers record (examples/sample_bimodal.rs:137-139, examples/sample_high_dim.rs:84-86).
 * The reference has no autocorrelation code: the estimator is the one DESIGN.md section 4.6 states, so
 * parity here is against this restatement and an independent numpy one (PARITY UNPINNED by the reference). */
#include <stddef.h>
#include <stdint.h>

#include "scorus_oracle.h"

/* trace[t][c], t < T.  Same summation orders as the device kernels (stats_kernel.cuh: iat_*_kernel), so the
 * results are bit-identical: mean = (sum over k < 256 of (sum of x_t, t = k mod 256)) / T; centred series;
 * C(k) = (sum_{t < T-k} xc[t] xc[t+k]) / T summed left to right; Sokal window. */
void orc_trace_iat(const double *trace, size_t T, size_t ncap, double c_window, size_t max_lag, double *tau_out,
                   uint32_t *window_out, double *mean_out, double *var_out, double *xc /* scratch [T] */)
{
    const size_t L = (max_lag == 0 || max_lag > T - 1) ? T - 1 : max_lag;
    for (size_t c = 0; c < ncap; ++c) {
        double part[256];
        for (size_t k = 0; k < 256; ++k) {
            double s = 0.0;
            for (size_t t = k; t < T; t += 256) s += trace[t * ncap + c];
            part[k] = s;
        }
        double tot = 0.0;
        for (size_t k = 0; k < 256; ++k) tot += part[k];
        const double mu = tot / (double)T;
        for (size_t t = 0; t < T; ++t) xc[t] = trace[t * ncap + c] - mu;
        double c0 = 0.0, tau = 1.0;
        size_t m = L;
        for (size_t k = 0; k <= L; ++k) {
            double s = 0.0;
            for (size_t t = 0; t + k < T; ++t) s += xc[t] * xc[t + k];
            const double ck = s / (double)T;
            if (k == 0) { c0 = ck; continue; }
            tau += 2.0 * (ck / c0);
            if ((double)k >= c_window * tau) { m = k; break; }
        }
        tau_out[c] = tau;
        if (window_out) window_out[c] = (uint32_t)m;
        if (mean_out) mean_out[c] = mu;
        if (var_out) var_out[c] = c0;
    }
}

/* Running moments of one rung over a sequence of snapshots, samples[s][w][d] (s < nsnap, w < npb): mean and
 * biased covariance about that mean over all nsnap * npb rows, as mean / cov of utils.rs would give for the
 * concatenated rows (two passes, left to right). */
void orc_running_moments(const double *samples, size_t nsnap, size_t npb, size_t dim, double *mean_out, double *cov_out)
{
    const size_t m = nsnap * npb;
    for (size_t d = 0; d < dim; ++d) mean_out[d] = 0.0;
    for (size_t r = 0; r < m; ++r)
        for (size_t d = 0; d < dim; ++d) mean_out[d] += samples[r * dim + d];
    for (size_t d = 0; d < dim; ++d) mean_out[d] *= 1.0 / (double)m;
    if (!cov_out) return;
    for (size_t e = 0; e < dim * dim; ++e) cov_out[e] = 0.0;
    for (size_t r = 0; r < m; ++r)
        for (size_t i = 0; i < dim; ++i)
            for (size_t j = 0; j < dim; ++j)
                cov_out[i * dim + j] += (samples[r * dim + i] - mean_out[i]) * (samples[r * dim + j] - mean_out[j]);
    for (size_t e = 0; e < dim * dim; ++e) cov_out[e] /= (double)m;
}
