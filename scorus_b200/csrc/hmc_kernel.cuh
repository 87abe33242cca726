// hmc_kernel.cuh -- ensemble / parallel-tempering Hamiltonian Monte Carlo:
// src/mcmc/hmc/naive.rs:160-292 (sample_ensemble_pt_impl as called by sample_ensemble_pt, :122-158: the
// gradient at entry is recomputed from q0) with src/mcmc/hmc/utils.rs:6-30 (leapfrog).
//
// Walkers are independent here (no pairing): a group of G lanes owns one walker, lane j holds double2
// chunk j (+ k G) of q, p and the force in registers for the whole trajectory; neighbour coordinates for
// coupled gradients (Rosenbrock) come from the adjacent lanes by shuffle; the kinetic energies and the
// final log-density are group reductions.  One row in, at most one row out per walker per step: the kernel
// is bound by fp64 arithmetic (l gradient evaluations per walker), not by HBM.
//
// Rounding follows the reference operation by operation (-fmad=false): p_mid = p - lhq (eps/2);
// q += (p_mid * 1) eps; lhq = grad(q) (-beta); p = p_mid - lhq (eps/2) (utils.rs:21-29).
//
// The per-rung step size adapts inside the reference's sequential accept loop (:269-283): this kernel
// writes one code per walker (+1 grow, -1 shrink, 0 keep); hmc_adapt_kernel replays them in walker order
// (bit-identical epsilon) for rungs of up to 4096 walkers, hmc_adapt_count_kernel applies their net
// effect for larger ones (see below).
#pragma once
#include "tile_common.cuh"

namespace scorus {

enum : uint32_t { ST_HMC_P = 11 };  // idx = walker, sub = d/2: Box-Muller pair -> momenta of coordinates d, d+1

struct HmcParams {
    StepParams s;          // ensemble, log-probs, beta, plugin params, seed / step, shard offsets
    const double *eps;     // [nbeta] step size per rung (device)
    const double *p_in;    // [n][dim] caller-supplied momenta, or nullptr -> device stream
    const double *u2_in;   // [n] caller-supplied adaptation uniforms (accept uniforms: s.u_in)
    int8_t *adapt_out;     // [n] +1 / -1 / 0
    double target_accept_ratio;
    uint32_t l;            // leapfrog steps
};

// momenta of coordinates d, d+1 of a walker: Box-Muller on one Philox draw
__device__ __forceinline__ double2 hmc_momentum(uint64_t seed, uint64_t step, uint32_t walker, uint32_t chunk)
{
    const U4 r = draw(seed, step, ST_HMC_P, walker, chunk);
    double u = u53(r.x, r.y);
    const double v = u53(r.z, r.w);
    if (u < 1e-300) u = 1e-300;
    const double rad = sqrt(-2.0 * log(u)), t = (2.0 * 3.14159265358979323846) * v;
    double sn, cs;
    sincos(t, &sn, &cs);
    return make_double2(rad * cs, rad * sn);
}

template <int G>
__device__ __forceinline__ double group_sum(double v)
{
#pragma unroll
    for (int h = 1; h < G; h <<= 1) v += __shfl_xor_sync(0xffffffffu, v, h);
    return v;
}

template <class Plugin, int G, int ITERS>
__global__ void __launch_bounds__(256) hmc_kernel(const HmcParams hp)
{
    const StepParams &p = hp.s;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t j = lane % G;
    constexpr uint32_t kGroups = 32u / G;
    const uint32_t gid = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kGroups + lane / G;
    const uint32_t ngroups = gridDim.x * (blockDim.x >> 5) * kGroups;
    const uint32_t nvec = p.pitch >> 1;
    unsigned long long n_acc = 0;

    Plugin plug;
    plug.init(p.pparams, p.nparams, p.dim);

    // every lane of a warp runs the same number of trips (shuffles are warp-wide)
    const uint32_t trips = (p.n + ngroups - 1u) / ngroups;
    for (uint32_t t = 0; t < trips; ++t) {
        const uint32_t wi = t * ngroups + gid;
        const bool act = wi < p.n;
        const uint32_t w = act ? wi : 0u;
        const uint32_t rung = w / p.npb;
        const double beta = __ldg(p.beta + rung), eps = __ldg(hp.eps + rung);
        const double he = eps * 0.5;           // epsilon * half, utils.rs:21
        const double nbeta = 0.0 - beta;       // h_q(q) = grad(q) * (-beta), naive.rs:225
        const double lp0 = __ldg(p.lp + w);

        double2 q[ITERS], mom[ITERS], f[ITERS];
        bool in[ITERS], has1[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c = (uint32_t)it * G + j, d = 2u * c;
            in[it] = c < nvec && d < p.dim;
            has1[it] = in[it] && d + 1u < p.dim;
            q[it] = in[it] ? reinterpret_cast<const double2 *>(p.ens + (size_t)w * p.pitch)[c] : make_double2(0.0, 0.0);
            double2 m = make_double2(0.0, 0.0);
            if (in[it]) {
                if (hp.p_in) {
                    m.x = __ldg(hp.p_in + (size_t)w * p.dim + d);
                    if (has1[it]) m.y = __ldg(hp.p_in + (size_t)w * p.dim + d + 1u);
                } else {
                    m = hmc_momentum(p.seed, p.step, w + p.w_off, c);
                    if (!has1[it]) m.y = 0.0;
                }
            }
            mom[it] = m;
        }

        // Dense targets (kNeedsRow: lp = -1/2 |L (x - mu)|^2): v = L (x - mu) and g = 0 - L^T v as matrix-vector products
        // over the group's lanes -- lane j owns the rows / columns of its coordinates, the other operand's chunks come
        // round by shuffle, in coordinate order, so that every sum is the left-to-right sum oracle/hmc.c states.  L is
        // read from L1 / L2 (D^2 doubles; 80 KB at D = 100): functional, not fast -- a tile GEMM over 32 walkers on the
        // fp64 tensor cores (as K3 does for the stretch move) is what would make HMC on this target quick.
        [[maybe_unused]] double2 vrow[ITERS];
        [[maybe_unused]] auto dense_apply = [&]() {  // vrow <- L (q - mu), my rows
            if constexpr (Plugin::kNeedsRow) {
                const uint32_t D = p.dim;
                double2 xm[ITERS];
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const uint32_t d = 2u * ((uint32_t)it * G + j);
                    xm[it].x = in[it] ? q[it].x - __ldg(plug.mu + d) : 0.0;
                    xm[it].y = has1[it] ? q[it].y - __ldg(plug.mu + d + 1u) : 0.0;
                    vrow[it] = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int itk = 0; itk < ITERS; ++itk)
                    for (uint32_t jk = 0; jk < (uint32_t)G; ++jk) {
                        const uint32_t k = 2u * ((uint32_t)itk * G + jk);
                        if (k >= D) break;  // (uniform over the group)
                        const double xk0 = __shfl_sync(0xffffffffu, xm[itk].x, jk, G), xk1 = __shfl_sync(0xffffffffu, xm[itk].y, jk, G);
                        const bool k1 = k + 1u < D;
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            const uint32_t i = 2u * ((uint32_t)it * G + j);
                            if (in[it]) {
                                vrow[it].x += __ldg(plug.L + (size_t)i * D + k) * xk0;
                                if (k1) vrow[it].x += __ldg(plug.L + (size_t)i * D + k + 1u) * xk1;
                            }
                            if (has1[it]) {
                                vrow[it].y += __ldg(plug.L + (size_t)(i + 1u) * D + k) * xk0;
                                if (k1) vrow[it].y += __ldg(plug.L + (size_t)(i + 1u) * D + k + 1u) * xk1;
                            }
                        }
                    }
            }
        };
        // force at a point: lhq = grad(q) * scale; neighbours x[d-1], x[d+2] from the adjacent lanes / trips
        auto force = [&](double scale) {
            if constexpr (Plugin::kNeedsRow) {
                const uint32_t D = p.dim;
                dense_apply();
                double2 s[ITERS];
#pragma unroll
                for (int it = 0; it < ITERS; ++it) s[it] = make_double2(0.0, 0.0);
#pragma unroll
                for (int iti = 0; iti < ITERS; ++iti)
                    for (uint32_t ji = 0; ji < (uint32_t)G; ++ji) {
                        const uint32_t i = 2u * ((uint32_t)iti * G + ji);
                        if (i >= D) break;
                        const double v0 = __shfl_sync(0xffffffffu, vrow[iti].x, ji, G), v1 = __shfl_sync(0xffffffffu, vrow[iti].y, ji, G);
                        const bool i1 = i + 1u < D;
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            const uint32_t k = 2u * ((uint32_t)it * G + j);
                            if (in[it]) {
                                s[it].x += __ldg(plug.L + (size_t)i * D + k) * v0;
                                if (i1) s[it].x += __ldg(plug.L + (size_t)(i + 1u) * D + k) * v1;
                            }
                            if (has1[it]) {
                                s[it].y += __ldg(plug.L + (size_t)i * D + k + 1u) * v0;
                                if (i1) s[it].y += __ldg(plug.L + (size_t)(i + 1u) * D + k + 1u) * v1;
                            }
                        }
                    }
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    f[it].x = in[it] ? (0.0 - s[it].x) * scale : 0.0;
                    f[it].y = has1[it] ? (0.0 - s[it].y) * scale : 0.0;
                }
                return;
            }
            if constexpr (Plugin::kGradNeedsSums) {  // row-wide sums first (mixture: the component weights)
                double a[Plugin::kAcc];
#pragma unroll
                for (int k = 0; k < Plugin::kAcc; ++k) a[k] = 0.0;
#pragma unroll
                for (int it = 0; it < ITERS; ++it)
                    if (in[it]) plug.terms(2u * ((uint32_t)it * G + j), q[it].x, q[it].y, 0.0, has1[it], false, a);
#pragma unroll
                for (int k = 0; k < Plugin::kAcc; ++k) a[k] = group_sum<G>(a[k]);
                plug.grad_prepare(a);
            }
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c = (uint32_t)it * G + j, d = 2u * c;
                double xm1 = __shfl_up_sync(0xffffffffu, q[it].y, 1, G);
                double xp2 = __shfl_down_sync(0xffffffffu, q[it].x, 1, G);
                if (ITERS > 1) {
                    const double prev_last = __shfl_sync(0xffffffffu, q[it > 0 ? it - 1 : 0].y, G - 1, G);
                    const double next_first = __shfl_sync(0xffffffffu, q[it + 1 < ITERS ? it + 1 : it].x, 0, G);
                    if (j == 0 && it > 0) xm1 = prev_last;
                    if (j == G - 1 && it + 1 < ITERS) xp2 = next_first;
                }
                double g0 = 0.0, g1 = 0.0;
                if constexpr (Plugin::kHasGrad) plug.grad2(d, xm1, q[it].x, q[it].y, xp2, d > 0u, has1[it], d + 2u < p.dim, g0, g1);
                f[it].x = in[it] ? g0 * scale : 0.0;
                f[it].y = has1[it] ? g1 * scale : 0.0;
            }
        };
        auto kinetic = [&]() {  // p.dot(p) / two * m_inv, naive.rs:194 (fixed tree instead of the left-to-right sum)
            double s = 0.0;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                s += mom[it].x * mom[it].x;
                s += mom[it].y * mom[it].y;
            }
            return group_sum<G>(s) / 2.0 * 1.0;
        };

        const double k0 = kinetic();
        force(-1.0 * beta);  // last_grad_logprob * (-1 * beta), :201-205, with the gradient recomputed at entry (:139)
        for (uint32_t s = 0; s < hp.l; ++s) {  // leapfrog, utils.rs:17-29
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                mom[it].x = mom[it].x - f[it].x * he;
                mom[it].y = mom[it].y - f[it].y * he;
                q[it].x = q[it].x + (mom[it].x * 1.0) * eps;
                q[it].y = q[it].y + (mom[it].y * 1.0) * eps;
            }
            force(nbeta);
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                mom[it].x = mom[it].x - f[it].x * he;
                mom[it].y = mom[it].y - f[it].y * he;
            }
        }

        // log-density of the end point
        double acc[Plugin::kAcc];
#pragma unroll
        for (int a = 0; a < Plugin::kAcc; ++a) acc[a] = 0.0;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t c = (uint32_t)it * G + j, d = 2u * c;
            double nb = __shfl_down_sync(0xffffffffu, q[it].x, 1, G);
            if (ITERS > 1) {
                const double next_first = __shfl_sync(0xffffffffu, q[it + 1 < ITERS ? it + 1 : it].x, 0, G);
                if (j == G - 1 && it + 1 < ITERS) nb = next_first;
            }
            if (in[it]) plug.terms(d, q[it].x, q[it].y, nb, has1[it], d + 2u < p.dim, acc);
        }
#pragma unroll
        for (int a = 0; a < Plugin::kAcc; ++a) acc[a] = group_sum<G>(acc[a]);
        double lp1 = plug.finish_acc(acc);
        if constexpr (Plugin::kNeedsRow) {  // -1/2 |L (q - mu)|^2 from the same product (fixed tree over the group's lanes)
            dense_apply();
            double ss = 0.0;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                if (in[it]) ss += vrow[it].x * vrow[it].x;
                if (has1[it]) ss += vrow[it].y * vrow[it].y;
            }
            lp1 = plug.ok ? -0.5 * group_sum<G>(ss) : NAN;
        }
        const double k1 = kinetic();

        // :231-253: current_u = -lp, proposed_u = -flogprob(q), dh = beta (u0 - u1) + k0 - k1
        const double u0 = -lp0, pu = -lp1;
        const double dh = beta * (u0 - pu) + k0 - k1;
        double ua, ub;
        if (p.u_in) {
            ua = __ldg(p.u_in + w);
            ub = __ldg(hp.u2_in + w);
        } else {
            const U4 r = draw(p.seed, p.step, ST_ZU, w + p.w_off, 0);
            ua = u53(r.x, r.y);
            ub = u53(r.z, r.w);
        }
        // :269: is_finite() first, then the strict comparison
        const bool accept = act && !isinf(dh) && !isnan(dh) && ua < exp(dh);
        if (accept) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
                if (in[it]) reinterpret_cast<double2 *>(p.ens + (size_t)w * p.pitch)[(uint32_t)it * G + j] = q[it];
        }
        if (act && j == 0) {
            if (accept) {
                p.lp[w] = -pu;  // :271
                if (p.dirty) p.dirty[w] = 1;
                ++n_acc;
                if (p.rung_acc) atomicAdd(p.rung_acc + rung, 1ull);
            }
            // :274-283: grow after an accept with probability 1 - target, shrink after a reject with probability target
            hp.adapt_out[w] = accept ? (ub < 1.0 - hp.target_accept_ratio ? 1 : 0) : (ub < hp.target_accept_ratio ? -1 : 0);
            if (p.accepted_out) p.accepted_out[w] = accept ? 1 : 0;
        }
    }
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

// epsilon[r] after the reference's accept loop (:269-283), which multiplies or divides it once per adapting walker, in
// walker order.  Rungs of up to kHmcExactAdapt walkers replay exactly that, one thread per rung: bit-identical to the
// sequential loop.  A chain of npb dependent roundings cannot be parallelised exactly, and for larger rungs it would
// cost more than the trajectories (2.6 ms against 0.7 ms at 64 x 16384 walkers; 0.4 s for one rung of 2^22), so there
// the same factors are applied at once, epsilon (1 + adj)^(grow - shrink) by repeated squaring: within npb 2^-53
// relative of the sequential result (1.8e-12 worst case at 16384, ~1e-14 typical).  adj_factor == 0
// (HmcParam::fixed) never changes epsilon and skips the launch.
constexpr uint32_t kHmcExactAdapt = 4096;

static __global__ void hmc_adapt_kernel(const int8_t *adapt, uint32_t npb, uint32_t nbeta, double adj_factor, double *eps)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nbeta) return;
    const double fct = 1.0 + adj_factor;
    double e = eps[r];
    const int8_t *a = adapt + (size_t)r * npb;
    for (uint32_t i = 0; i < npb; ++i) {
        const int c = a[i];
        if (c > 0) e = e * fct;
        else if (c < 0) e = e / fct;
    }
    eps[r] = e;
}

// one block per rung: net = #grow - #shrink, epsilon *= (1 + adj)^net
static __global__ void __launch_bounds__(256) hmc_adapt_count_kernel(const int8_t *adapt, uint32_t npb, double adj_factor, double *eps)
{
    __shared__ int part[8];
    const uint32_t r = blockIdx.x;
    const int8_t *a = adapt + (size_t)r * npb;
    int net = 0;
    for (uint32_t i = threadIdx.x; i < npb; i += blockDim.x) net += a[i];
    for (int off = 16; off; off >>= 1) net += __shfl_xor_sync(0xffffffffu, net, off);
    if ((threadIdx.x & 31u) == 0) part[threadIdx.x >> 5] = net;
    __syncthreads();
    if (threadIdx.x == 0) {
        net = 0;
        for (int k = 0; k < 8; ++k) net += part[k];
        double base = 1.0 + adj_factor, pw = 1.0;
        for (unsigned m = (unsigned)(net < 0 ? -net : net); m; m >>= 1) {
            if (m & 1u) pw = pw * base;
            base = base * base;
        }
        eps[r] = net < 0 ? eps[r] / pw : eps[r] * pw;
    }
}

}  // namespace scorus
