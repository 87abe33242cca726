// logprob.cuh -- device-side log-density plugins.
//
// These replace the `flogprob: &F` closure of the reference
// (src/mcmc/ensemble_sample.rs:88,102,161).  A plugin consumes the proposed point one
// coordinate at a time, in index order, so every per-walker sum is accumulated left to right
// exactly as the reference closures do -- with -fmad=false the results are
// bit-identical to the CPU oracle for the transcendental-free targets.
//
// Interface (all __device__):
//   void   init(const double* params, uint32_t nparams, uint32_t dim)
//   void   accum(uint32_t d, double y)        // coordinate d of the proposal, d ascending
//   double finish()
//   static constexpr bool kNeedsRow           // true: finish_row(row) re-reads the whole row
#pragma once
#include <cmath>
#include <cstdint>

namespace scorus {

// -sum c*x^2.  examples/test_emcee.rs:18-24 (c = 1/2), examples/sample_high_dim.rs:22-28 (c = 1)
struct LpGaussian {
    static constexpr bool kNeedsRow = false;
    double c, acc;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t)
    {
        c = np >= 1 ? __ldg(p) : 0.5;
        acc = 0.0;
    }
    __device__ __forceinline__ void accum(uint32_t, double y)
    {
        double sq = y * y;
        acc -= sq * c;
    }
    __device__ __forceinline__ double finish() const { return acc; }
};

// examples/sample_bimodal.rs:35-44: subtract, then -inf as soon as one |x| exceeds the limit
struct LpBoundedGaussian {
    static constexpr bool kNeedsRow = false;
    double limit, acc;
    bool dead;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t)
    {
        limit = np >= 1 ? __ldg(p) : 1.5;
        acc = 0.0;
        dead = false;
    }
    __device__ __forceinline__ void accum(uint32_t, double y)
    {
        acc -= y * y;
        dead |= fabs(y) > limit;
    }
    __device__ __forceinline__ double finish() const { return dead ? -INFINITY : acc; }
};

// examples/test_emcee.rs:26-32: -sum_i [100 (x[i+1] - x[i]^2)^2 + (1 - x[i])^2]
struct LpRosenbrock {
    static constexpr bool kNeedsRow = false;
    double prev, acc;
    __device__ __forceinline__ void init(const double *, uint32_t, uint32_t)
    {
        prev = 0.0;
        acc = 0.0;
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        if (d > 0) {
            double xi2 = prev * prev;
            double t = y - xi2;
            double t2 = t * t;
            double s = 1.0 - prev;
            double s2 = s * s;
            double term = 100.0 * t2 + s2;
            acc += term;
        }
        prev = y;
    }
    __device__ __forceinline__ double finish() const { return -acc; }
};

// examples/sample_rastrigin.rs:17-25: -(n a) + sum [a cos(2 pi x) - x^2]
struct LpRastrigin {
    static constexpr bool kNeedsRow = false;
    double a, acc;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t dim)
    {
        a = np >= 1 ? __ldg(p) : 10.0;
        acc = -((double)dim * a);
    }
    __device__ __forceinline__ void accum(uint32_t, double y)
    {
        const double two_pi = 2.0 * 3.14159265358979323846;
        double arg = two_pi * y;
        double c = a * cos(arg);
        double sq = y * y;
        acc += c - sq;
    }
    __device__ __forceinline__ double finish() const { return acc; }
};

// examples/sample_bimodal.rs:46-54
struct LpBimodal2d {
    static constexpr bool kNeedsRow = false;
    double lo0, hi0, lo1, hi1, split, mu_a, sg_a, mu_b, sg_b, x0, x1;
    uint32_t dim;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t d)
    {
        lo0 = -15.0; hi0 = 15.0; lo1 = 0.0; hi1 = 1.0; split = 0.5;
        mu_a = -5.0; sg_a = 0.1; mu_b = 5.0; sg_b = 1.0;
        if (np >= 9) {
            lo0 = __ldg(p); hi0 = __ldg(p + 1); lo1 = __ldg(p + 2); hi1 = __ldg(p + 3);
            split = __ldg(p + 4); mu_a = __ldg(p + 5); sg_a = __ldg(p + 6); mu_b = __ldg(p + 7);
            sg_b = __ldg(p + 8);
        }
        x0 = 0.0; x1 = 0.0; dim = d;
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        if (d == 0) x0 = y;
        if (d == 1) x1 = y;
    }
    __device__ __forceinline__ double finish() const
    {
        if (dim < 2) return NAN;
        if (x0 < lo0 || x0 > hi0 || x1 < lo1 || x1 > hi1) return -INFINITY;
        double mu = x1 < split ? mu_a : mu_b;
        double sigma = x1 < split ? sg_a : sg_b;
        double dx = x0 - mu;
        double num = (-dx) * dx;
        double den = (2.0 * sigma) * sigma;
        return num / den - log(sigma);
    }
};

// examples/sample_bimodal.rs:56-65
struct LpStep2d {
    static constexpr bool kNeedsRow = false;
    double x0, x1;
    uint32_t dim;
    __device__ __forceinline__ void init(const double *, uint32_t, uint32_t d)
    {
        x0 = 0.0; x1 = 0.0; dim = d;
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        if (d == 0) x0 = y;
        if (d == 1) x1 = y;
    }
    __device__ __forceinline__ double finish() const
    {
        if (dim < 2) return NAN;
        if (x0 < 0.0 || x0 > 1.0 || x1 < 0.0 || x1 > 1.0) return -INFINITY;
        return x0 > 0.5 ? log(0.1) : log(0.9);
    }
};

// builder-defined (BASELINE config 3): logsumexp of two isotropic Gaussians at +m and -m
struct LpMixture {
    static constexpr bool kNeedsRow = false;
    const double *m;
    double s1, s2, q1, q2;
    uint32_t dim;
    bool ok;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t d)
    {
        ok = np >= 2 + d;
        s1 = ok ? __ldg(p) : 1.0;
        s2 = ok ? __ldg(p + 1) : 1.0;
        m = p + 2;
        q1 = 0.0; q2 = 0.0; dim = d;
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        double md = ok ? __ldg(m + d) : 0.0;
        double a = y - md;
        double b = y + md;
        q1 += a * a;
        q2 += b * b;
    }
    __device__ __forceinline__ double finish() const
    {
        if (!ok) return NAN;
        double t1 = (-q1) / ((2.0 * s1) * s1) - (double)dim * log(s1);
        double t2 = (-q2) / ((2.0 * s2) * s2) - (double)dim * log(s2);
        double hi = t1 > t2 ? t1 : t2;
        double lo = t1 > t2 ? t2 : t1;
        if (hi == -INFINITY) return -INFINITY;
        return hi + log1p(exp(lo - hi));
    }
};

// builder-defined (BASELINE config 2): -0.5 |L (x - mu)|^2, general dense L.
// Scalar fallback evaluated from the finished row in shared memory; the tensor-core (DMMA)
// tile version lives in dense_kernel.cuh.
struct LpCorrGaussian {
    static constexpr bool kNeedsRow = true;
    const double *mu, *L;
    uint32_t dim;
    bool ok;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t d)
    {
        ok = np >= d + d * d;
        mu = p;
        L = p + d;
        dim = d;
    }
    __device__ __forceinline__ void accum(uint32_t, double) {}
    __device__ __forceinline__ double finish() const { return NAN; }
    __device__ __forceinline__ double finish_row(const double *row) const
    {
        if (!ok) return NAN;
        double s = 0.0;
        for (uint32_t i = 0; i < dim; ++i) {
            double w = 0.0;
            const double *Li = L + (size_t)i * dim;
            for (uint32_t j = 0; j < dim; ++j) {
                double c = row[j] - __ldg(mu + j);
                w += __ldg(Li + j) * c;
            }
            s += w * w;
        }
        return -0.5 * s;
    }
};

}  // namespace scorus
