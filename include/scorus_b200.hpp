// scorus_b200.hpp -- header-only C++ host layer over the C ABI (include/scorus_b200.h).
//
// Mirrors the reference's Rust interface for the hot path: same function names, argument order
// and error behaviour, with the two substitutions the device forces:
//
//   scorus::mcmc::ensemble_sample::init_logprob(flogprob, ensemble, logprob)                    // src/mcmc/ensemble_sample.rs:77
//   scorus::mcmc::ensemble_sample::sample(flogprob, ensemble, cached_logprob, rng, a, ufs)      // :87
//   scorus::mcmc::ensemble_sample::sample_pt(flogprob, ensemble, cached_logprob, rng, a, ufs, beta_list)   // :107
//   scorus::mcmc::utils::swap_walkers(ensemble, logprob, rng, beta_list)                        // src/mcmc/utils.rs:72
//   scorus::mcmc::utils::draw_z(rng, a)                                                         // :33
//
//   flogprob  : scorus::mcmc::LogProb      (a registered device plugin instead of a closure)
//   rng       : scorus::mcmc::DeviceRng    (seed of the device streams; caches the device context)
//   ensemble  : std::vector<std::vector<double>>  (the reference's Vec<LsVec<f64, Vec<f64>>>)
//
// Where the reference assert!s or panics these throw McmcError (src/mcmc/mcmc_errors.rs:1-7 variants)
// or std::invalid_argument.  Long runs should hold a Sampler and keep the state on the device.
#pragma once

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "scorus_b200.h"

namespace scorus {
namespace mcmc {

// src/mcmc/mcmc_errors.rs:1-7
enum class McmcErr { NWalkersIsZero = 1, NWalkersIsNotEven = 2, NWalkersMismatchesNBeta = 3, BetaNotInDecrOrd = 4 };

struct McmcError : std::runtime_error {
    int code;
    McmcError(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
    McmcErr variant() const { return static_cast<McmcErr>(code); }
};

inline void check(int code, const scorus_ctx *ctx = nullptr)
{
    if (code == SCORUS_OK) return;
    std::string msg = scorus_status_string(code);
    if (ctx) {
        const char *d = scorus_last_error(ctx);
        if (d && *d) msg += std::string(" (") + d + ")";
    }
    if (code >= 1 && code <= 4) throw McmcError(code, msg);
    if (code == SCORUS_ERR_CUDA || code == SCORUS_ERR_NO_DEVICE) throw std::runtime_error(msg);
    throw std::invalid_argument(msg);
}

// ---- log-density plugins (replace the `flogprob: &F` closure) ----
struct LogProb {
    int kind;
    std::vector<double> params;
    static LogProb gaussian(double c = 0.5) { return {SCORUS_LP_GAUSSIAN, {c}}; }             // examples/test_emcee.rs:18-24
    static LogProb bounded_gaussian(double limit = 1.5) { return {SCORUS_LP_BOUNDED_GAUSSIAN, {limit}}; }
    static LogProb rosenbrock() { return {SCORUS_LP_ROSENBROCK, {}}; }                          // examples/test_emcee.rs:26-32
    static LogProb rastrigin(double a = 10.0) { return {SCORUS_LP_RASTRIGIN, {a}}; }            // examples/sample_rastrigin.rs:17-25
    static LogProb bimodal2d() { return {SCORUS_LP_BIMODAL2D, {}}; }                            // examples/sample_bimodal.rs:46-54
    static LogProb step2d() { return {SCORUS_LP_STEP2D, {}}; }
    static LogProb corr_gaussian(const std::vector<double> &mu, const std::vector<double> &L)
    {
        LogProb p{SCORUS_LP_CORR_GAUSSIAN, mu};
        p.params.insert(p.params.end(), L.begin(), L.end());
        return p;
    }
    static LogProb mixture(double s1, double s2, const std::vector<double> &m)
    {
        LogProb p{SCORUS_LP_MIXTURE, {s1, s2}};
        p.params.insert(p.params.end(), m.begin(), m.end());
        return p;
    }
};

namespace ensemble_sample {

// enum UpdateFlagSpec { Prob(T), All, Func(&mut dyn FnMut() -> Vec<bool>) }, src/mcmc/ensemble_sample.rs:25-32
class UpdateFlagSpec {
public:
    static UpdateFlagSpec Prob(double p) { UpdateFlagSpec u; u.kind_ = SCORUS_UFS_PROB; u.prob_ = p; return u; }
    static UpdateFlagSpec All() { UpdateFlagSpec u; u.kind_ = SCORUS_UFS_ALL; return u; }
    static UpdateFlagSpec Func(std::function<std::vector<bool>()> f)
    {
        UpdateFlagSpec u; u.kind_ = SCORUS_UFS_MASK; u.func_ = std::move(f); return u;
    }
    // the cycling one-hot closure of examples/sample_high_dim.rs:59-65, kept on the device
    static UpdateFlagSpec OneHotCycle() { UpdateFlagSpec u; u.kind_ = SCORUS_UFS_ONE_HOT_CYCLE; return u; }

    // evaluates Func once per walker (as :150-153 does) and fills the C struct
    scorus_ufs materialize(size_t n, size_t dim)
    {
        scorus_ufs u{};
        u.kind = kind_;
        u.prob = prob_;
        if (kind_ == SCORUS_UFS_MASK) {
            mask_.clear();
            size_t len = 0;
            for (size_t i = 0; i < n; ++i) {
                std::vector<bool> f = func_();
                if (i == 0) len = f.size();
                if (f.size() != len) throw std::invalid_argument("Func must return flag vectors of one length per step");
                if (len > dim) throw std::out_of_range("update_flags longer than the walker (index panic, :70)");
                for (bool b : f) mask_.push_back(b ? 1 : 0);
            }
            u.mask = mask_.data();
            u.mask_len = len;
        }
        return u;
    }
    int kind() const { return kind_; }

private:
    int kind_ = SCORUS_UFS_ALL;
    double prob_ = 0.0;
    std::function<std::vector<bool>()> func_;
    std::vector<uint8_t> mask_;
};

}  // namespace ensemble_sample

// RAII handle over scorus_ctx: the device-resident (ensemble, logprob) and the RNG counters
class Sampler {
public:
    Sampler(const LogProb &f, size_t n_walkers, size_t dim, uint64_t seed = 0, int device = -1, uint32_t flags = 0)
        : n_(n_walkers), dim_(dim)
    {
        scorus_cfg cfg{};
        cfg.n_walkers = n_walkers;
        cfg.dim = static_cast<uint32_t>(dim);
        cfg.device = device;
        cfg.seed = seed;
        cfg.flags = flags;
        check(scorus_ctx_create(&ctx_, &cfg));
        set_logprob(f);
    }
    ~Sampler() { scorus_ctx_destroy(ctx_); }
    Sampler(const Sampler &) = delete;
    Sampler &operator=(const Sampler &) = delete;

    void set_logprob(const LogProb &f)
    {
        check(scorus_set_logprob(ctx_, f.kind, f.params.empty() ? nullptr : f.params.data(), f.params.size()), ctx_);
    }
    void upload(const double *ens, const double *lp) { check(scorus_upload(ctx_, ens, lp), ctx_); }
    void download(double *ens, double *lp) { check(scorus_download(ctx_, ens, lp), ctx_); }
    void init_logprob() { check(scorus_init_logprob(ctx_), ctx_); }
    void sample_pt(double a, ensemble_sample::UpdateFlagSpec &ufs, const std::vector<double> &beta, uint64_t nsteps = 1)
    {
        const uint64_t reps = ufs.kind() == SCORUS_UFS_MASK ? nsteps : 1;
        for (uint64_t r = 0; r < reps; ++r) {
            scorus_ufs u = ufs.materialize(n_, dim_);
            check(scorus_sample_pt(ctx_, a, &u, beta.data(), beta.size(), reps == 1 ? nsteps : 1), ctx_);
        }
    }
    void swap_walkers(const std::vector<double> &beta) { check(scorus_swap_walkers(ctx_, beta.data(), beta.size()), ctx_); }
    void sample_pt_host(double *ens, double *lp, size_t lp_len, double a, ensemble_sample::UpdateFlagSpec &ufs,
                        const std::vector<double> &beta)
    {
        scorus_ufs u = ufs.materialize(n_, dim_);
        check(scorus_sample_pt_host(ctx_, ens, lp, lp_len, a, &u, beta.data(), beta.size()), ctx_);
    }
    void swap_walkers_host(double *ens, double *lp, size_t lp_len, const std::vector<double> &beta)
    {
        check(scorus_swap_walkers_host(ctx_, ens, lp, lp_len, beta.data(), beta.size()), ctx_);
    }
    double draw_z(double a)
    {
        double z = 0.0;
        check(scorus_draw_z(ctx_, a, 1, &z), ctx_);
        return z;
    }
    // t-walk, src/mcmc/twalk.rs:586-648 (namespace twalk below mirrors the module)
    void twalk_sample(const scorus_twalk_params &param, const std::vector<double> &beta, uint64_t nsteps = 1)
    {
        check(scorus_twalk_sample(ctx_, &param, beta.data(), beta.size(), nsteps), ctx_);
    }
    void twalk_sample_host(double *ens, double *lp, size_t lp_len, const scorus_twalk_params &param,
                           const std::vector<double> &beta, uint64_t nsteps = 1)
    {
        check(scorus_twalk_sample_host(ctx_, ens, lp, lp_len, &param, beta.data(), beta.size(), nsteps), ctx_);
    }
    // ensemble / parallel-tempering HMC, src/mcmc/hmc/naive.rs:122-292 (namespace hmc below mirrors the module)
    std::vector<size_t> hmc_sample_ensemble_pt_host(double *q0, double *lp, size_t lp_len, std::vector<double> &epsilon,
                                                    const std::vector<double> &beta, size_t l, const scorus_hmc_param &param)
    {
        if (epsilon.size() != beta.size()) throw std::invalid_argument("assert_eq!(epsilon.len(), nbeta)");  // :179
        std::vector<uint64_t> cnt(beta.size());
        check(scorus_hmc_sample_ensemble_pt_host(ctx_, q0, lp, lp_len, epsilon.data(), beta.data(), beta.size(), l, &param,
                                                 cnt.data()), ctx_);
        return std::vector<size_t>(cnt.begin(), cnt.end());
    }
    // chain statistics on the device (what the example drivers compute from per-step downloads)
    struct RungStats { std::vector<uint64_t> accepted, proposed, swapped, swap_proposed; };
    RungStats rung_stats(size_t nrungs)
    {
        RungStats r{std::vector<uint64_t>(nrungs), std::vector<uint64_t>(nrungs), std::vector<uint64_t>(nrungs),
                    std::vector<uint64_t>(nrungs)};
        check(scorus_get_rung_stats(ctx_, nrungs, r.accepted.data(), r.proposed.data(), r.swapped.data(), r.swap_proposed.data()), ctx_);
        r.swapped.resize(nrungs - 1);
        r.swap_proposed.resize(nrungs - 1);
        return r;
    }
    void moments_enable(size_t nrungs, uint64_t every = 1, bool with_cov = true)
    {
        check(scorus_moments_enable(ctx_, nrungs, every, with_cov ? 1 : 0), ctx_);
    }
    uint64_t moments(size_t rung, std::vector<double> &mean, std::vector<double> *cov = nullptr)
    {
        uint64_t m = 0;
        mean.resize(dim_);
        if (cov) cov->resize(dim_ * dim_);
        check(scorus_moments_download(ctx_, rung, mean.data(), cov ? cov->data() : nullptr, &m), ctx_);
        return m;
    }
    void trace_enable(const std::vector<uint32_t> &walkers, const std::vector<uint32_t> &coords, size_t capacity_steps, uint64_t every = 1)
    {
        check(scorus_trace_enable(ctx_, walkers.data(), coords.data(), walkers.size(), capacity_steps), ctx_);
        check(scorus_trace_stride(ctx_, every), ctx_);
        ncapture_ = walkers.size();
    }
    std::vector<double> trace_iat(double c_window = 5.0, size_t max_lag = 0, std::vector<uint32_t> *window = nullptr)
    {
        std::vector<double> tau(ncapture_);
        if (window) window->resize(ncapture_);
        check(scorus_trace_iat(ctx_, c_window, max_lag, tau.data(), window ? window->data() : nullptr, nullptr, nullptr), ctx_);
        return tau;
    }
    void sync() { check(scorus_sync(ctx_), ctx_); }
    scorus_ctx *raw() { return ctx_; }
    size_t n_walkers() const { return n_; }
    size_t dim() const { return dim_; }

private:
    scorus_ctx *ctx_ = nullptr;
    size_t n_, dim_, ncapture_ = 0;
};

// Stands in for `rng: &mut U`: the seed of the device streams.  Caches one Sampler per ensemble shape.
class DeviceRng {
public:
    explicit DeviceRng(uint64_t seed = 0, int device = -1) : seed_(seed), device_(device) {}
    Sampler &sampler(const LogProb &f, size_t n, size_t dim)
    {
        auto key = std::make_pair(n, dim);
        auto it = cache_.find(key);
        if (it == cache_.end()) it = cache_.emplace(key, std::make_unique<Sampler>(f, n, dim, seed_, device_)).first;
        else it->second->set_logprob(f);
        return *it->second;
    }

private:
    uint64_t seed_;
    int device_;
    std::map<std::pair<size_t, size_t>, std::unique_ptr<Sampler>> cache_;
};

namespace detail {
using Ensemble = std::vector<std::vector<double>>;
inline std::vector<double> flatten(const Ensemble &e, size_t dim)
{
    std::vector<double> flat(e.size() * dim);
    for (size_t i = 0; i < e.size(); ++i) {
        if (e[i].size() != dim) throw std::invalid_argument("walkers of different dimension");
        std::copy(e[i].begin(), e[i].end(), flat.begin() + i * dim);
    }
    return flat;
}
inline void unflatten(const std::vector<double> &flat, Ensemble &e, size_t dim)
{
    for (size_t i = 0; i < e.size(); ++i) std::copy(flat.begin() + i * dim, flat.begin() + (i + 1) * dim, e[i].begin());
}
}  // namespace detail

namespace ensemble_sample {

// src/mcmc/ensemble_sample.rs:77-85
inline void init_logprob(const LogProb &flogprob, const detail::Ensemble &ensemble, std::vector<double> &logprob,
                         DeviceRng &rng)
{
    if (ensemble.empty()) throw McmcError(SCORUS_ERR_NWALKERS_IS_ZERO, "NWalkersIsZero");
    if (logprob.size() != ensemble.size()) throw std::invalid_argument("copy_from_slice: lengths differ");  // :84
    const size_t dim = ensemble[0].size();
    Sampler &s = rng.sampler(flogprob, ensemble.size(), dim);
    std::vector<double> flat = detail::flatten(ensemble, dim);
    s.upload(flat.data(), nullptr);
    s.download(nullptr, logprob.data());
}

// src/mcmc/ensemble_sample.rs:107-190, in place on the host containers
inline void sample_pt(const LogProb &flogprob, detail::Ensemble &ensemble, std::vector<double> &cached_logprob,
                      DeviceRng &rng, double a, UpdateFlagSpec &ufs, const std::vector<double> &beta_list)
{
    check(scorus_check_sample_args(ensemble.size(), cached_logprob.size(), beta_list.size()));  // :127-133
    const size_t dim = ensemble[0].size();
    Sampler &s = rng.sampler(flogprob, ensemble.size(), dim);
    std::vector<double> flat = detail::flatten(ensemble, dim);
    s.sample_pt_host(flat.data(), cached_logprob.data(), cached_logprob.size(), a, ufs, beta_list);
    detail::unflatten(flat, ensemble, dim);
}

// src/mcmc/ensemble_sample.rs:87-105
inline void sample(const LogProb &flogprob, detail::Ensemble &ensemble, std::vector<double> &cached_logprob,
                   DeviceRng &rng, double a, UpdateFlagSpec &ufs)
{
    sample_pt(flogprob, ensemble, cached_logprob, rng, a, ufs, {1.0});
}

}  // namespace ensemble_sample

namespace utils {

// src/mcmc/utils.rs:72-112; a length mismatch is the reference's silent no-op (:88)
inline void swap_walkers(detail::Ensemble &ensemble, std::vector<double> &logprob, DeviceRng &rng,
                         const std::vector<double> &beta_list, const LogProb &flogprob = LogProb::gaussian())
{
    const size_t n = ensemble.size(), nb = beta_list.size();
    if (nb == 0) throw std::invalid_argument("empty beta list");
    if ((n / nb) * nb != n) throw McmcError(SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");  // :86
    if (logprob.size() != n) return;
    const size_t dim = ensemble[0].size();
    Sampler &s = rng.sampler(flogprob, n, dim);
    std::vector<double> flat = detail::flatten(ensemble, dim);
    s.swap_walkers_host(flat.data(), logprob.data(), logprob.size(), beta_list);
    detail::unflatten(flat, ensemble, dim);
}

// src/mcmc/utils.rs:33-45
inline double draw_z(DeviceRng &rng, double a) { return rng.sampler(LogProb::gaussian(), 2, 2).draw_z(a); }

}  // namespace utils

namespace twalk {

// TWalkParams, src/mcmc/twalk.rs:75-130 (fw holds cumulative weights, as the reference stores them)
struct TWalkParams : scorus_twalk_params {
    static TWalkParams make(size_t n)  // TWalkParams::new(n), :89-110
    {
        TWalkParams p;
        check(scorus_twalk_params_default(n, &p));
        return p;
    }
    TWalkParams with_fw(double w0, double w1) const  // :112-124: plain weights in
    {
        TWalkParams p = *this;
        p.fw[0] = 0.0 + w0;
        p.fw[1] = p.fw[0] + w1;
        return p;
    }
    TWalkParams with_pphi(double v) const  // :126-128
    {
        TWalkParams p = *this;
        p.pphi = v;
        return p;
    }
};

// twalk::sample, :586-648, in place on the host containers; nthreads is accepted and ignored
inline void sample(const LogProb &flogprob, std::pair<detail::Ensemble, std::vector<double>> &ensemble_logprob,
                   const TWalkParams &param, DeviceRng &rng, const std::vector<double> &beta_list, size_t nthreads = 1)
{
    (void)nthreads;
    detail::Ensemble &ensemble = ensemble_logprob.first;
    std::vector<double> &logprob = ensemble_logprob.second;
    check(scorus_check_sample_args(ensemble.size(), logprob.size(), beta_list.size()));  // :675 and the pairs of :610
    const size_t dim = ensemble[0].size();
    Sampler &s = rng.sampler(flogprob, ensemble.size(), dim);
    std::vector<double> flat = detail::flatten(ensemble, dim);
    s.twalk_sample_host(flat.data(), logprob.data(), logprob.size(), param, beta_list);
    detail::unflatten(flat, ensemble, dim);
}

}  // namespace twalk

namespace hmc {
namespace naive {

// HmcParam, src/mcmc/hmc/naive.rs:17-56
struct HmcParam : scorus_hmc_param {
    HmcParam(double target_accept_ratio_, double adj_factor_) : scorus_hmc_param{target_accept_ratio_, adj_factor_} {}
    HmcParam() : scorus_hmc_param{0.99, 0.000001} {}                                   // Default, :49-56
    static HmcParam quick_adj(double t) { return HmcParam(t, 0.01); }                  // :36-38
    static HmcParam slow_adj(double t) { return HmcParam(t, 0.000001); }               // :40-42
    static HmcParam fixed(double t) { return HmcParam(t, 0.0); }                       // :44-46
};

// sample_ensemble_pt, :122-158, in place on the host containers; returns the accepted count per rung.  The
// gradient (the reference's second closure) is the log-density plugin's own.
inline std::vector<size_t> sample_ensemble_pt(const LogProb &flogprob, detail::Ensemble &q0, std::vector<double> &lp,
                                              DeviceRng &rng, std::vector<double> &epsilon,
                                              const std::vector<double> &beta_list, size_t l, const HmcParam &param)
{
    if (q0.empty()) throw McmcError(SCORUS_ERR_NWALKERS_IS_ZERO, "NWalkersIsZero");
    if (q0.size() != lp.size()) throw std::invalid_argument("assert_eq!(q0.len(), lp.len())");  // :177
    const size_t dim = q0[0].size();
    Sampler &s = rng.sampler(flogprob, q0.size(), dim);
    std::vector<double> flat = detail::flatten(q0, dim);
    std::vector<size_t> cnt = s.hmc_sample_ensemble_pt_host(flat.data(), lp.data(), lp.size(), epsilon, beta_list, l, param);
    detail::unflatten(flat, q0, dim);
    return cnt;
}

}  // namespace naive
}  // namespace hmc
}  // namespace mcmc
}  // namespace scorus
