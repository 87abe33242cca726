// logprob.cuh -- device-side log-density plugins.
//
// These replace the `flogprob: &F` closure of the reference
// (src/mcmc/ensemble_sample.rs:88,102,161).  Two evaluation interfaces per plugin:
//
// Sequential (step_seq_kernel, init_logprob_kernel: one lane walks one row):
//   void   init(const double* params, uint32_t nparams, uint32_t dim)
//   void   accum(uint32_t d, double y)        // coordinate d of the point, d ascending
//   double finish()
//   Every per-walker sum is accumulated left to right exactly as the reference closures do
//   (SURVEY App. B); with -fmad=false the transcendental-free targets are bit-identical to the
//   CPU oracle.
//
// Cooperative (step_coop_kernel: the lanes of a warp share a row, two coordinates per lane and
// trip; per-lane partial sums are combined by a fixed shuffle tree, so the summation ORDER
// differs from the reference's -- the values agree to a few ulp of the sum of |terms|, the bar
// is 1e-12 relative):
//   static constexpr int kAcc                 // partial accumulators per walker (1 or 2)
//   void   terms(d, y0, y1, nb, has1, hasnb, acc)   // coordinates d (and d+1 if has1);
//                                                   // nb = coordinate d+2 (valid if hasnb)
//   double finish_acc(const double* acc)      // acc = lane-reduced accumulators
//
//   static constexpr bool kNeedsRow           // true: finish_row(row) evaluates from the whole row
//
// Gradient (hmc_kernel.cuh; replaces the reference's second closure `grad_logprob: &G`,
// src/mcmc/hmc/naive.rs:124): plugins with kHasGrad provide
//   void grad2(d, xm1, x0, x1, xp2, hasprev, has1, hasnb, g0, g1)   // d lp / d x[d], d lp / d x[d+1];
//                                                                   // xm1 = x[d-1], xp2 = x[d+2]
// with the operation order oracle/hmc.c: orc_grad_logprob states.
#pragma once
#include <cmath>
#include <cstdint>
#include "rng.cuh"  // fma_t

namespace scorus {

// -sum c*x^2.  examples/test_emcee.rs:18-24 (c = 1/2), examples/sample_high_dim.rs:22-28 (c = 1)
struct LpGaussian {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 1;
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
    __device__ __forceinline__ void terms(uint32_t, double y0, double y1, double, bool has1, bool, double *a) const
    {
        a[0] -= (y0 * y0) * c;
        if (has1) a[0] -= (y1 * y1) * c;
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return a[0]; }
    static constexpr bool kHasGrad = true;
    static constexpr bool kGradNeedsSums = false;
    __device__ __forceinline__ void grad2(uint32_t, double, double x0, double x1, double, bool, bool, bool, double &g0,
                                          double &g1) const
    {
        const double k = -2.0 * c;
        g0 = k * x0;
        g1 = k * x1;
    }
};

// examples/sample_bimodal.rs:35-44: subtract, then -inf as soon as one |x| exceeds the limit
struct LpBoundedGaussian {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 2;  // [0] the sum, [1] number of coordinates beyond the limit
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
    __device__ __forceinline__ void terms(uint32_t, double y0, double y1, double, bool has1, bool, double *a) const
    {
        a[0] -= y0 * y0;
        a[1] += fabs(y0) > limit ? 1.0 : 0.0;
        if (has1) {
            a[0] -= y1 * y1;
            a[1] += fabs(y1) > limit ? 1.0 : 0.0;
        }
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return a[1] > 0.0 ? -INFINITY : a[0]; }
    static constexpr bool kHasGrad = true;  // inside the box; a trajectory that ends outside has lp = -inf and is rejected
    static constexpr bool kGradNeedsSums = false;
    __device__ __forceinline__ void grad2(uint32_t, double, double x0, double x1, double, bool, bool, bool, double &g0,
                                          double &g1) const
    {
        g0 = -2.0 * x0;
        g1 = -2.0 * x1;
    }
};

// examples/test_emcee.rs:26-32: -sum_i [100 (x[i+1] - x[i]^2)^2 + (1 - x[i])^2]
struct LpRosenbrock {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = true;   // term(x[d+1], x[d+2]) couples neighbouring chunks
    static constexpr int kAcc = 1;
    double prev, acc;
    __device__ __forceinline__ void init(const double *, uint32_t, uint32_t)
    {
        prev = 0.0;
        acc = 0.0;
    }
    static __device__ __forceinline__ double term(double xi, double xn)
    {
        double xi2 = xi * xi;
        double t = xn - xi2;
        double t2 = t * t;
        double s = 1.0 - xi;
        double s2 = s * s;
        return 100.0 * t2 + s2;
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        if (d > 0) acc += term(prev, y);
        prev = y;
    }
    __device__ __forceinline__ double finish() const { return -acc; }
    __device__ __forceinline__ void terms(uint32_t, double y0, double y1, double nb, bool has1, bool hasnb,
                                          double *a) const
    {
        if (has1) a[0] += term(y0, y1);
        if (hasnb) a[0] += term(y1, nb);
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return -a[0]; }
    static constexpr bool kHasGrad = true;
    static constexpr bool kGradNeedsSums = false;
    // the reference's own example closure, examples/test_hmc.rs:27-37, keeping the i that contribute:
    // g[j] = 0;  i = j-1: g[j] -= 200 (x[j] - x[j-1]^2) * 1;  i = j: g[j] -= 200 (x[j+1] - x[j]^2) (0 - 2 x[j]) + 2 (x[j] - 1)
    static __device__ __forceinline__ double grad1(double xm, double x, double xn, bool hasprev, bool hasnext)
    {
        double r = 0.0;
        if (hasprev) r -= (200.0 * (x - xm * xm)) * 1.0;
        if (hasnext) r -= (200.0 * (xn - x * x)) * (0.0 - 2.0 * x) + 2.0 * (x - 1.0);
        return r;
    }
    __device__ __forceinline__ void grad2(uint32_t, double xm1, double x0, double x1, double xp2, bool hasprev, bool has1,
                                          bool hasnb, double &g0, double &g1) const
    {
        g0 = grad1(xm1, x0, x1, hasprev, has1);
        g1 = grad1(x0, x1, xp2, true, hasnb);
    }
};

// examples/sample_rastrigin.rs:17-25: -(n a) + sum [a cos(2 pi x) - x^2]
struct LpRastrigin {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 1;
    double a, acc, base;
    // cos(2 pi y) without a general-purpose cos: y - rint(y) is exact and |r| <= 1/2, cos(2 pi r) = -cos(2 pi (1/2 - |r|))
    // beyond a quarter period, and on [0, 1/4] an even polynomial in b (Taylor in b^2 to b^22: truncation 2e-17) by
    // Horner -- twenty instructions and NO branch, so the twelve evaluations a lane makes per tile interleave; the
    // library cos is 56 instructions behind a slow-path branch that serialises them (27 % of the c4 step's instructions
    // and 29 % of its stall samples, profiles/r2b_c4_step_lines.txt).  Within 3.2e-16 of the true cos(2 pi y) (measured
    // over 2e5 arguments in [-6, 6]); the reference's cos(fl(2 pi y)) is itself 5e-15 away from it -- the rounding of
    // the argument -- so the two differ by that much per term, far inside the 1e-12 sum|terms| bar of the parity tests.
    static __device__ __forceinline__ double cos2pi(double y)
    {
        const double r = y - rint(y);
        const double ar = fabs(r);
        const bool flip = ar > 0.25;
        const double b = flip ? 0.5 - ar : ar;
        const double s = b * b;
        // (-1)^k (2 pi)^(2k) / (2k)!, k = 1..11 (typed constants: the f32 instantiation converts them)
        const double c1 = -0x1.3bd3cc9be45dep+4, c2 = 0x1.03c1f081b5ac4p+6, c3 = -0x1.55d3c7e3cbffap+6, c4 = 0x1.e1f506891babbp+5,
                     c5 = -0x1.a6d1f2a204a8cp+4, c6 = 0x1.f9d38a3763cc3p+2, c7 = -0x1.b6e24f44b128fp+0, c8 = 0x1.20c62c2f2d7f5p-2,
                     c9 = -0x1.2a0c591af8314p-5, c10 = 0x1.ef6e308d6d1c4p-9, c11 = -0x1.52ae4120fde27p-12, one = 1.0;
        double p;
        if constexpr (sizeof(double) == 8) {  // (the f32 instantiation reads sizeof(float): seven terms are its 1e-8)
            p = c11;
            p = fma_t(p, s, c10);
            p = fma_t(p, s, c9);
            p = fma_t(p, s, c8);
            p = fma_t(p, s, c7);
        } else {
            p = c7;
        }
        p = fma_t(p, s, c6);
        p = fma_t(p, s, c5);
        p = fma_t(p, s, c4);
        p = fma_t(p, s, c3);
        p = fma_t(p, s, c2);
        p = fma_t(p, s, c1);
        p = fma_t(p, s, one);
        return flip ? -p : p;
    }
    static __device__ __forceinline__ double term(double a, double y)
    {
        double c = a * cos2pi(y);  // examples/sample_rastrigin.rs:21: a * (2 pi x).cos()
        double sq = y * y;
        return c - sq;
    }
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t dim)
    {
        a = np >= 1 ? __ldg(p) : 10.0;
        base = -((double)dim * a);
        acc = base;
    }
    __device__ __forceinline__ void accum(uint32_t, double y) { acc += term(a, y); }
    __device__ __forceinline__ double finish() const { return acc; }
    __device__ __forceinline__ void terms(uint32_t, double y0, double y1, double, bool has1, bool, double *s) const
    {
        s[0] += term(a, y0);
        if (has1) s[0] += term(a, y1);
    }
    __device__ __forceinline__ double finish_acc(const double *s) const { return base + s[0]; }
    static constexpr bool kHasGrad = true;  // g = -(2 pi a) sin(2 pi x) - 2 x
    static constexpr bool kGradNeedsSums = false;
    __device__ __forceinline__ void grad2(uint32_t, double, double x0, double x1, double, bool, bool, bool, double &g0,
                                          double &g1) const
    {
        const double two_pi = 2.0 * 3.14159265358979323846;
        const double k = two_pi * a;
        const double s0 = k * sin(two_pi * x0), s1 = k * sin(two_pi * x1);
        g0 = (0.0 - s0) - 2.0 * x0;
        g1 = (0.0 - s1) - 2.0 * x1;
    }
};

// examples/sample_bimodal.rs:46-54
struct LpBimodal2d {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 2;  // [0] carries x0, [1] carries x1
    static constexpr bool kHasGrad = false;
    static constexpr bool kGradNeedsSums = false;
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
    __device__ __forceinline__ double eval(double a0, double a1) const
    {
        if (dim < 2) return NAN;
        if (a0 < lo0 || a0 > hi0 || a1 < lo1 || a1 > hi1) return -INFINITY;
        double mu = a1 < split ? mu_a : mu_b;
        double sigma = a1 < split ? sg_a : sg_b;
        double dx = a0 - mu;
        double num = (-dx) * dx;
        double den = (2.0 * sigma) * sigma;
        return num / den - log(sigma);
    }
    __device__ __forceinline__ double finish() const { return eval(x0, x1); }
    __device__ __forceinline__ void terms(uint32_t d, double y0, double y1, double, bool has1, bool, double *a) const
    {
        if (d == 0) {  // only the lane holding coordinates 0,1 contributes; x + 0 is exact
            a[0] += y0;
            if (has1) a[1] += y1;
        }
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return eval(a[0], a[1]); }
};

// examples/sample_bimodal.rs:56-65
struct LpStep2d {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 2;
    static constexpr bool kHasGrad = false;
    static constexpr bool kGradNeedsSums = false;
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
    __device__ __forceinline__ double eval(double a0, double a1) const
    {
        if (dim < 2) return NAN;
        if (a0 < 0.0 || a0 > 1.0 || a1 < 0.0 || a1 > 1.0) return -INFINITY;
        return a0 > 0.5 ? log(0.1) : log(0.9);
    }
    __device__ __forceinline__ double finish() const { return eval(x0, x1); }
    __device__ __forceinline__ void terms(uint32_t d, double y0, double y1, double, bool has1, bool, double *a) const
    {
        if (d == 0) {
            a[0] += y0;
            if (has1) a[1] += y1;
        }
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return eval(a[0], a[1]); }
};

// builder-defined (BASELINE config 3): logsumexp of two isotropic Gaussians at +m and -m
struct LpMixture {
    static constexpr bool kNeedsRow = false;
    static constexpr int kMaxWarps = 14;
    static constexpr bool kNeedsNext = false;  // terms() uses the coordinate after the lane's chunk
    static constexpr int kAcc = 2;
    static constexpr bool kHasGrad = true;
    static constexpr bool kGradNeedsSums = true;  // the component weights need |x - m|^2 and |x + m|^2 of the whole row
    const double *m;
    double s1, s2, q1, q2;
    double den1, den2, c1, c2;  // (2 s) s and D ln s, hoisted out of the per-walker tail (same roundings)
    double r1, r2;              // gradient: component weight / sigma^2 of the current point (grad_prepare)
    uint32_t dim;
    bool ok;
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t d)
    {
        ok = np >= 2 + d;
        s1 = ok ? __ldg(p) : 1.0;
        s2 = ok ? __ldg(p + 1) : 1.0;
        m = p + 2;
        q1 = 0.0; q2 = 0.0; dim = d;
        den1 = (2.0 * s1) * s1;
        den2 = (2.0 * s2) * s2;
        c1 = (double)d * log(s1);
        c2 = (double)d * log(s2);
    }
    __device__ __forceinline__ void accum(uint32_t d, double y)
    {
        double md = ok ? __ldg(m + d) : 0.0;
        double a = y - md;
        double b = y + md;
        q1 += a * a;
        q2 += b * b;
    }
    __device__ __forceinline__ double eval(double a1, double a2) const
    {
        if (!ok) return NAN;
        double t1 = (-a1) / den1 - c1;
        double t2 = (-a2) / den2 - c2;
        double hi = t1 > t2 ? t1 : t2;
        double lo = t1 > t2 ? t2 : t1;
        if (hi == -INFINITY) return -INFINITY;
        return hi + log1p(exp(lo - hi));
    }
    __device__ __forceinline__ double finish() const { return eval(q1, q2); }
    __device__ __forceinline__ void terms(uint32_t d, double y0, double y1, double, bool has1, bool, double *a) const
    {
        double m0 = ok ? __ldg(m + d) : 0.0;
        double u0 = y0 - m0, v0 = y0 + m0;
        a[0] += u0 * u0;
        a[1] += v0 * v0;
        if (has1) {
            double m1 = ok ? __ldg(m + d + 1) : 0.0;
            double u1 = y1 - m1, v1 = y1 + m1;
            a[0] += u1 * u1;
            a[1] += v1 * v1;
        }
    }
    __device__ __forceinline__ double finish_acc(const double *a) const { return eval(a[0], a[1]); }
    // gradient of the logsumexp: w1 (-(x - m) / s1^2) + w2 (-(x + m) / s2^2), w_i = exp(t_i - lse); a = the row's
    // (|x - m|^2, |x + m|^2) as terms() accumulates them
    __device__ __forceinline__ void grad_prepare(const double *a)
    {
        const double t1 = (-a[0]) / den1 - c1, t2 = (-a[1]) / den2 - c2;
        const double lse = eval(a[0], a[1]);
        r1 = exp(t1 - lse) / (s1 * s1);
        r2 = exp(t2 - lse) / (s2 * s2);
    }
    __device__ __forceinline__ void grad2(uint32_t d, double, double x0, double x1, double, bool, bool has1, bool, double &g0,
                                          double &g1) const
    {
        const double m0 = ok ? __ldg(m + d) : 0.0, m1 = ok && has1 ? __ldg(m + d + 1) : 0.0;
        g0 = (0.0 - r1 * (x0 - m0)) - r2 * (x0 + m0);
        g1 = (0.0 - r1 * (x1 - m1)) - r2 * (x1 + m1);
    }
};

#ifndef SCORUS_F32  // the dense target runs on the fp64 tensor cores: no f32 instantiation (gen_f32.py)
#ifndef SCORUS_K3_NG
#define SCORUS_K3_NG 4
#endif

// fp64 tensor-core MMA, D[8x8] += A[8x4] * B[4x8] (SASS: DMMA.8x8x4).  Lane l holds A[l/4][l%4],
// B[l%4][l/4] and C/D[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// builder-defined (BASELINE config 2): -0.5 |L (x - mu)|^2, general dense L.
// finish_row: scalar evaluation, one lane per walker, reference-style left-to-right sums (used by the
// sequential kernels and init_logprob).  tile_logprob: the K3 path of the register kernel -- the
// 32 proposed rows of a warp's tile times L^T as a batched fp64 tensor-core GEMM (this really is a
// dense contraction: 2 D^2 flops per 16 D bytes), row norms from the accumulators.
struct LpCorrGaussian {
    static constexpr bool kNeedsRow = true;
    static constexpr bool kTile = true;
    static constexpr int kMaxWarps = 8;  // register kernel: 256 threads -> up to 255 registers for the DMMA accumulators
    static constexpr bool kNeedsNext = false;
    static constexpr int kAcc = 1;
    // gradient -L^T L (x - mu): two matrix-vector products, evaluated by the HMC kernel's dense path (hmc_kernel.cuh:
    // the walker's coordinates are spread over a group of lanes, so the products go through shuffles, not grad2)
    static constexpr bool kHasGrad = true;
    static constexpr bool kGradNeedsSums = false;
    __device__ __forceinline__ void grad2(uint32_t, double, double, double, double, bool, bool, bool, double &g0, double &g1) const
    {
        g0 = g1 = 0.0;  // (never called: kNeedsRow plugins take the dense path)
    }
    const double *mu, *L;
    uint32_t dim;
    bool ok, upper;  // upper: L[n][k] == 0 for k < n (flagged by scorus_set_logprob after {mu, L})
    __device__ __forceinline__ void init(const double *p, uint32_t np, uint32_t d)
    {
        ok = np >= d + d * d;
        mu = p;
        L = p + d;
        dim = d;
        upper = np > d + d * d && __ldg(p + d + d * d) != 0.0;
    }
    __device__ __forceinline__ void accum(uint32_t, double) {}
    __device__ __forceinline__ double finish() const { return NAN; }
    __device__ __forceinline__ void terms(uint32_t, double, double, double, bool, bool, double *) const {}
    __device__ __forceinline__ double finish_acc(const double *) const { return NAN; }
    // One pass of the tile GEMM over NA n-tiles (8 NA columns of W = (Y - mu) L^T starting at n0): adds the
    // squares of my accumulator entries to ss[m-tile].  NA is a template parameter so that the last,
    // narrower pass has no branches between its DMMAs (a uniform branch there cost 2.4x).
    template <int NA>
    __device__ __forceinline__ void tile_pass(uint32_t n0, const double *const (&rows)[4], uint32_t g, uint32_t t,
                                              double (&ss)[4]) const
    {
        // (folding mu into the accumulator start, W = Y L^T - (L mu)^T, measured 10 % SLOWER than
        //  subtracting it from every A fragment: the DADDs fill issue slots the DMMA pipe leaves idle)
        double c[NA][4][2];
#pragma unroll
        for (int ng = 0; ng < NA; ++ng)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) c[ng][mt][0] = c[ng][mt][1] = 0.0;
        // columns n0.. of W only see L[n][k] with k >= n0 when L is upper triangular (n0 is a multiple of 8)
        for (uint32_t k0 = upper ? n0 : 0u; k0 < dim; k0 += 4) {
            const uint32_t k = k0 + t;
            const bool kin = k < dim;
            const double mu_k = kin ? __ldg(mu + k) : 0.0;
            double a[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = kin ? rows[mt][k] - mu_k : 0.0;
#pragma unroll
            for (int ng = 0; ng < NA; ++ng) {
                const uint32_t n = n0 + 8u * ng + g;  // W column = row of L:  B[k][n] = L[n][k]
                const double b = (kin && n < dim) ? __ldg(L + (size_t)n * dim + k) : 0.0;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) dmma_m8n8k4(c[ng][mt], a[mt], b);
            }
        }
#pragma unroll
        for (int ng = 0; ng < NA; ++ng)
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) ss[mt] += c[ng][mt][0] * c[ng][mt][0] + c[ng][mt][1] * c[ng][mt][1];
    }

    // log-density of row `lane` of a 32-row tile in shared memory (rows row_stride bytes apart)
    __device__ __forceinline__ double tile_logprob(const unsigned char *tile, uint32_t row_stride, uint32_t lane) const
    {
        if (!ok) return NAN;
        const uint32_t g = lane >> 2, t = lane & 3u;
        const double *rows[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
            rows[mt] = reinterpret_cast<const double *>(tile + (size_t)(mt * 8 + g) * row_stride);
        double ss[4] = {0.0, 0.0, 0.0, 0.0};  // sum over my columns of W[row mt*8+g][.]^2
        constexpr int NG = SCORUS_K3_NG;      // n-tiles per full pass: NG*8 accumulator doubles, NG*4 independent DMMAs per k-step
        const uint32_t ntiles = (dim + 7u) / 8u, full = ntiles / NG, rest = ntiles - full * NG;
        for (uint32_t ps = 0; ps < full; ++ps) tile_pass<NG>(ps * 8u * NG, rows, g, t, ss);
        if (rest == 1) tile_pass<1>(full * 8u * NG, rows, g, t, ss);
        else if (rest == 2) tile_pass<2>(full * 8u * NG, rows, g, t, ss);
        else if (rest == 3) tile_pass<(NG > 3 ? 3 : 1)>(full * 8u * NG, rows, g, t, ss);
        static_assert(NG == 4, "the remainder dispatch above assumes four n-tiles per full pass");
        double v = 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            double x = ss[mt];
            x += __shfl_xor_sync(0xffffffffu, x, 1);
            x += __shfl_xor_sync(0xffffffffu, x, 2);          // all columns of row mt*8+g
            x = __shfl_sync(0xffffffffu, x, (lane & 7u) * 4u);  // row (lane % 8) of m-tile mt
            if ((int)(lane >> 3) == mt) v = x;
        }
        return -0.5 * v;
    }
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
#endif  // SCORUS_F32

}  // namespace scorus
