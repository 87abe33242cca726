// stats_kernel.cuh -- the periphery of the path: K6 ensemble statistics, chain capture, box init.
//
//   mean / cov        src/linear_space/utils.rs:4-15, 17-41 (ensemble mean, biased /N covariance):
//                     the reference's own summary statistics, here as deterministic two-stage
//                     reductions over a walker range (a rung or the whole ensemble)
//   trace             the recording loops of the examples (`results[i].push(ensemble[i*n+0][0])`,
//                     examples/sample_high_dim.rs:84-86, examples/sample_bimodal.rs:137-139): selected
//                     (walker, coordinate) values captured on the device after every step
//   init              get_one_init_realization, src/mcmc/init_ensemble.rs:19-32: x[d] ~ U[lo[d], hi[d])
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rng.cuh"

namespace scorus {

// x[w][d] = lo[d] + u * (hi[d] - lo[d]), u from stream ST_INIT keyed by the GLOBAL walker id
__global__ void __launch_bounds__(256) init_ensemble_kernel(double *ens, uint32_t n, uint32_t dim, uint32_t pitch,
                                                            const double *lo, const double *hi, uint64_t seed,
                                                            uint32_t w_off)
{
    const uint32_t half = (dim + 1u) / 2u;
    const size_t total = (size_t)n * half;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(g / half), c = (uint32_t)(g - (size_t)w * half);
        const U4 r = draw(seed, 0, ST_INIT, w + w_off, c);
        const uint32_t d0 = 2u * c;
        double *row = ens + (size_t)w * pitch;
        row[d0] = __ldg(lo + d0) + u53(r.x, r.y) * (__ldg(hi + d0) - __ldg(lo + d0));
        if (d0 + 1u < dim) row[d0 + 1u] = __ldg(lo + d0 + 1u) + u53(r.z, r.w) * (__ldg(hi + d0 + 1u) - __ldg(lo + d0 + 1u));
    }
}

// stage 1 of the column sums: block (b, g) sums rows b, b+grid, ... of group g = [first + g*count, +count)
// -> part[g][b][dim].  A group is a rung (running moments) or the caller's range (gridDim.y == 1).  With
// `shift` (one row of dim per group) the terms are x - shift.
__global__ void __launch_bounds__(256) colsum_partial_kernel(const double *ens, uint32_t first, uint32_t count,
                                                             uint32_t dim, uint32_t pitch, const double *shift,
                                                             double *part)
{
    const uint32_t g = blockIdx.y;
    const size_t base = (size_t)first + (size_t)g * count;
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
        const double c = shift ? __ldg(shift + (size_t)g * dim + d) : 0.0;
        double s = 0.0;
        if (shift) {
            for (uint32_t r = blockIdx.x; r < count; r += gridDim.x) s += ens[(base + r) * pitch + d] - c;
        } else {
            for (uint32_t r = blockIdx.x; r < count; r += gridDim.x) s += ens[(base + r) * pitch + d];
        }
        part[((size_t)g * gridDim.x + blockIdx.x) * dim + d] = s;
    }
}

// stage 2: fixed-order sum over the partials of group blockIdx.y, then scale (mean: * (1/N) as utils.rs:14;
// cov: / N as :38); accumulate != 0 adds the unscaled sum to out (running moments)
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double *part, uint32_t nparts, uint32_t len,
                                                              double scale, int divide, int accumulate, double *out)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    const uint32_t g = blockIdx.y;
    part += (size_t)g * nparts * len;
    out += (size_t)g * len;
    double s = 0.0;
    for (uint32_t b = 0; b < nparts; ++b) s += part[(size_t)b * len + e];
    if (accumulate) out[e] += s;
    else out[e] = divide ? s / scale : s * scale;
}

// stage 1 of the covariance: block b accumulates sum_r c_r c_r^T over its rows (c = x - mean), rows staged
// through shared memory 32 at a time; thread t owns entries t, t+256, ... of the D x D matrix (D <= 128);
// blockIdx.y = group as above (its rows, its `mean` row, its partials)
constexpr int kCovMaxDim = 128;
__global__ void __launch_bounds__(256) cov_partial_kernel(const double *ens, uint32_t first, uint32_t count, uint32_t dim,
                                                          uint32_t pitch, const double *mean, double *part)
{
    extern __shared__ double rows[];  // [32][dim]
    constexpr int kPer = (kCovMaxDim * kCovMaxDim + 255) / 256;
    double acc[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) acc[k] = 0.0;
    const uint32_t len = dim * dim;
    first += blockIdx.y * count;
    mean += (size_t)blockIdx.y * dim;
    part += (size_t)blockIdx.y * gridDim.x * len;
    for (uint32_t r0 = blockIdx.x * 32u; r0 < count; r0 += gridDim.x * 32u) {
        const uint32_t nr = count - r0 < 32u ? count - r0 : 32u;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < nr * dim; e += blockDim.x) {
            const uint32_t r = e / dim, d = e - r * dim;
            rows[e] = ens[(size_t)(first + r0 + r) * pitch + d] - __ldg(mean + d);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const uint32_t e = threadIdx.x + 256u * (uint32_t)k;
            if (e < len) {
                const uint32_t i = e / dim, j = e - i * dim;
                double s = acc[k];
                for (uint32_t r = 0; r < nr; ++r) s += rows[r * dim + i] * rows[r * dim + j];
                acc[k] = s;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const uint32_t e = threadIdx.x + 256u * (uint32_t)k;
        if (e < len) part[(size_t)blockIdx.x * len + e] = acc[k];
    }
}

// Host-buffer call, download side: only walkers accepted since the upload changed, so only their rows
// and log-probs are written -- straight into the caller's pinned (device-mapped) host buffers at their
// final place.  One warp per walker; host rows are dim doubles apart (no padding).
__global__ void __launch_bounds__(256) scatter_download_kernel(const double *ens, const double *lp, const uint8_t *dirty,
                                                               uint32_t n, uint32_t dim, uint32_t pitch, double *host_ens,
                                                               double *host_lp)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = warp; w < n; w += nwarps) {
        if (!dirty[w]) continue;
        const double *src = ens + (size_t)w * pitch;
        double *dst = host_ens + (size_t)w * dim;
        if ((dim & 1u) == 0) {
            const double2 *s2 = reinterpret_cast<const double2 *>(src);
            double2 *d2 = reinterpret_cast<double2 *>(dst);
            for (uint32_t c = lane; c < dim / 2; c += 32) d2[c] = s2[c];
        } else {
            for (uint32_t c = lane; c < dim; c += 32) dst[c] = src[c];
        }
        if (lane == 0) host_lp[w] = lp[w];
    }
}

// out[slot][c] = ens[walker[c]][coord[c]]
__global__ void trace_kernel(const double *ens, uint32_t pitch, const uint32_t *walker, const uint32_t *coord,
                             uint32_t ncap, double *out_row)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncap) out_row[c] = ens[(size_t)walker[c] * pitch + coord[c]];
}

// ---- integrated autocorrelation time of the captured chains ----
// trace[t][c], t < T.  Per series c: mean (256 strided partial sums, combined in index order), centred copy
// xc[c][t], autocovariance C(k) = sum_{t < T-k} xc[t] xc[t+k] / T for k <= L (sequential sums), then Sokal's
// window: tau(M) = 1 + 2 sum_{k=1..M} C(k)/C(0), M = the first lag with M >= c_win * tau(M) (L if none).
__global__ void __launch_bounds__(256) iat_center_kernel(const double *trace, uint32_t ncap, uint32_t T, double *xc,
                                                         double *mean_out)
{
    __shared__ double part[256];
    const uint32_t c = blockIdx.x;
    double s = 0.0;
    for (uint32_t t = threadIdx.x; t < T; t += 256u) s += trace[(size_t)t * ncap + c];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < 256; ++k) tot += part[k];
        part[0] = tot / (double)T;
        mean_out[c] = part[0];
    }
    __syncthreads();
    const double mu = part[0];
    for (uint32_t t = threadIdx.x; t < T; t += 256u) xc[(size_t)c * T + t] = trace[(size_t)t * ncap + c] - mu;
}

__global__ void __launch_bounds__(128) iat_autocov_kernel(const double *xc, uint32_t T, uint32_t L, double *acov)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
    if (k > L) return;
    const double *x = xc + (size_t)c * T;
    double s = 0.0;
    for (uint32_t t = 0; t + k < T; ++t) s += x[t] * x[t + k];
    acov[(size_t)c * (L + 1u) + k] = s / (double)T;
}

__global__ void iat_window_kernel(const double *acov, uint32_t ncap, uint32_t L, double c_win, double *tau_out,
                                  uint32_t *window_out)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncap) return;
    const double *a = acov + (size_t)c * (L + 1u);
    const double c0 = a[0];
    double tau = 1.0;
    uint32_t m = L;
    for (uint32_t k = 1; k <= L; ++k) {
        tau += 2.0 * (a[k] / c0);
        if ((double)k >= c_win * tau) { m = k; break; }
    }
    tau_out[c] = tau;
    window_out[c] = m;
}

}  // namespace scorus
