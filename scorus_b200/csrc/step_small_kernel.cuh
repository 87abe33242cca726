// step_small_kernel.cuh -- the stretch-move step for ensembles that fit one thread block: the whole
// ensemble lives in shared memory and any number of consecutive steps run inside ONE launch.
//
// The big kernels pay a launch per step; for a 16-walker ensemble (BASELINE config 1, the shape of the
// reference's own examples/test_emcee.rs) that launch IS the step: 4.2 us, what the reference-structured
// CPU port needs too.  Here thread t owns walker t: per step the rungs are shuffled (rank of 64-bit Philox
// keys, the same keys and tie-break as exact_order_kernel), every thread draws its walker's z / u / flags
// from the same streams as the other kernels, builds the proposal from the OLD rows of its
// walker and partner (src/mcmc/ensemble_sample.rs:155-159), evaluates the log-density with the plugin's
// sequential interface -- the reference's left-to-right sums, so log-probs are bit-identical to the
// oracle for the transcendental-free targets -- and, after a block barrier, commits accepted rows
// (:181-188).  Positions are the same chain, bit for bit, as the streaming kernels produce.
//
// Shared memory: X[n][dimp] (current rows), Y[n][dimp] (proposals), LP[n], sort keys, the order, a dirty
// byte per walker and the per-warp flag words; dimp = pitch | 1 doubles keeps thread-strided row access
// spread over the banks.  Eligible when that fits (e.g. 1024 walkers x 10-D, 256 x 40-D).
#pragma once
#include "tile_common.cuh"

namespace scorus {

struct SmallLayout {
    uint32_t dimp;       // row stride in doubles
    size_t off_y, off_lp, off_keys, off_ord, off_dirty, off_flags, total;
};

__host__ __device__ inline SmallLayout small_layout(uint32_t n, uint32_t pitch, uint32_t flag_words, uint32_t warps)
{
    SmallLayout L;
    L.dimp = pitch | 1u;
    const size_t rows = (size_t)n * L.dimp * sizeof(double);
    L.off_y = rows;
    L.off_lp = 2 * rows;
    L.off_keys = L.off_lp + (size_t)n * sizeof(double);
    L.off_ord = L.off_keys + (size_t)n * sizeof(uint64_t);
    L.off_dirty = L.off_ord + (size_t)n * sizeof(uint32_t);
    L.off_flags = (L.off_dirty + n + 127u) & ~(size_t)127u;
    L.total = L.off_flags + (size_t)warps * (((size_t)flag_words * 128u + 127u) & ~(size_t)127u);
    return L;
}

template <class Plugin>
__global__ void __launch_bounds__(1024, 1) step_small_kernel(const StepParams p, uint32_t inner_steps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const SmallLayout L = small_layout(p.n, p.pitch, p.flag_words, blockDim.x >> 5);
    double *X = reinterpret_cast<double *>(smem);
    double *Y = reinterpret_cast<double *>(smem + L.off_y);
    double *LP = reinterpret_cast<double *>(smem + L.off_lp);
    uint64_t *keys = reinterpret_cast<uint64_t *>(smem + L.off_keys);
    uint32_t *ord = reinterpret_cast<uint32_t *>(smem + L.off_ord);
    uint8_t *dirty = smem + L.off_dirty;
    uint32_t *flagw = reinterpret_cast<uint32_t *>(smem + L.off_flags + (size_t)warp * (((size_t)p.flag_words * 128u + 127u) & ~(size_t)127u));
    const uint32_t dimp = L.dimp;
    const bool act = t < p.n;
    unsigned long long n_acc = 0;

    // ensemble -> shared memory (coalesced over the padded global rows)
    for (uint32_t e = t; e < p.n * p.dim; e += blockDim.x) {
        const uint32_t w = e / p.dim, d = e - w * p.dim;
        X[(size_t)w * dimp + d] = p.ens[(size_t)w * p.pitch + d];
    }
    if (act) { LP[t] = p.lp[t]; dirty[t] = 0; }
    __syncthreads();

    const uint32_t dim_mod = (uint32_t)(p.n_global % p.dim);
    for (uint32_t s = 0; s < inner_steps; ++s) {
        const uint64_t step = p.step + s;
        const uint64_t onehot = (p.onehot + (uint64_t)s * dim_mod) % p.dim;  // the cycling closure advances by n per step

        // thread t = walker t.  Its draws (shuffle key, z, u, flags) are independent of the matching, so the three
        // Philox calls overlap instead of queueing behind the two barriers of the shuffle.
        const uint32_t w = t;
        uint64_t key = 0;
        if (!p.order && act) {
            const U4 o = draw(p.seed, step, ST_PERM_KEY, p.w_off + t, 0);
            key = ((uint64_t)o.y << 32) | o.x;
            keys[t] = key;
        }
        double z = 1.0, u = 1.0;
        if (p.z_in) {
            if (act) { z = __ldg(p.z_in + w); u = __ldg(p.u_in + w); }
        } else {
            const U4 r = draw(p.seed, step, ST_ZU, w + p.w_off, 0);
            z = z_from_uniform(u53(r.x, r.y), p.inv_sqrt_a, p.z_range);
            u = u53(r.z, r.w);
        }
        uint32_t nphi = p.dim;
        const bool all_flags = p.flag_kind == 1u;
        if (!all_flags) nphi = gen_flags(p, step, onehot, w, flagw, lane, act);
        const double lz = (double)(nphi - 1u) * log(z);  // ensemble_sample.rs:184, first summand

        // ---- partner: from the caller's matching (fixed streams) or the exact shuffle of every rung ----
        uint32_t partner = 0;
        if (p.order) {
            if (act) ord[__ldg(p.order + t)] = __ldg(p.order + (t ^ 1u));  // walker at position t -> its partner
            __syncthreads();
            if (act) partner = ord[w];
        } else {
            __syncthreads();
            uint32_t pos = 0;
            if (act) {
                const uint32_t base = (t / p.npb) * p.npb;
                uint32_t rank = 0;
                for (uint32_t jj = 0; jj < p.npb; ++jj) {
                    const uint64_t kj = keys[base + jj];
                    rank += (kj < key) || (kj == key && base + jj < t);
                }
                pos = base + rank;
                ord[pos] = t;
            }
            __syncthreads();
            if (act) partner = ord[pos ^ 1u];
        }

        // ---- proposal from the old rows + log-density, one pass over the row ----
        bool accept = false;
        double lp_new = 0.0;
        if (act) {
            const double *x1 = X + (size_t)w * dimp, *x2 = X + (size_t)partner * dimp;
            double *y = Y + (size_t)w * dimp;
            const double omz = 1.0 - z;
            Plugin plug;
            plug.init(p.pparams, p.nparams, p.dim);
            for (uint32_t d = 0; d < p.dim; ++d) {
                const bool f = all_flags || ((flagw[(d >> 5) * 32u + lane] >> (d & 31u)) & 1u);
                // scale_vec, src/mcmc/utils.rs:55: (x1*z) + (x2*(1-z)), three roundings; propose_move :69-71
                const double a = x1[d], t1 = a * z, t2 = x2[d] * omz;
                const double yd = f ? t1 + t2 : a;
                y[d] = yd;
                if constexpr (!Plugin::kNeedsRow) plug.accum(d, yd);
            }
            if constexpr (Plugin::kNeedsRow) lp_new = plug.finish_row(y);
            else lp_new = plug.finish();
            // Metropolis accept, ensemble_sample.rs:181-188
            const double beta = __ldg(p.beta + w / p.npb);
            const double q = exp(lz + (lp_new - LP[w]) * beta);
            accept = u < q;
        }
        __syncthreads();  // every thread has read the old rows it needs
        if (accept) {
            const double *y = Y + (size_t)w * dimp;
            double *x = X + (size_t)w * dimp;
            for (uint32_t d = 0; d < p.dim; ++d) x[d] = y[d];
            LP[w] = lp_new;
            dirty[w] = 1;
            ++n_acc;
        }
        if (p.accepted_out && act) p.accepted_out[w] = accept ? 1 : 0;
        if (p.rung_acc) count_rung_accepts(p.rung_acc, act ? w / p.npb : 0u, act, accept, lane);
        __syncthreads();
    }

    // shared memory -> ensemble
    for (uint32_t e = t; e < p.n * p.dim; e += blockDim.x) {
        const uint32_t w = e / p.dim, d = e - w * p.dim;
        if (dirty[w]) p.ens[(size_t)w * p.pitch + d] = X[(size_t)w * dimp + d];
    }
    if (act && dirty[t]) {
        p.lp[t] = LP[t];
        if (p.dirty) p.dirty[t] = 1;
    }
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

}  // namespace scorus
