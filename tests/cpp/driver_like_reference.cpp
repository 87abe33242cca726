// Host-layer check in C++: the driver loops of examples/test_emcee.rs:34-86 and
// examples/sample_rastrigin.rs:27-71, written against include/scorus_b200.hpp with the reference's
// names and argument order.  Prints the final state; tests/test_parity_gpu.py compares it with the
// same run made through the Python mirror (same seed => identical numbers).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "scorus_b200.hpp"

using namespace scorus::mcmc;
using ensemble_sample::UpdateFlagSpec;

static double lcg(unsigned long long &s)  // any deterministic init; the Python side repeats it
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv)
{
    const int niter = argc > 1 ? atoi(argv[1]) : 30;
    unsigned long long st = 12345;
    {   // test_emcee.rs: rosenbrock, sample(), Prob(0.5), a = 2
        const size_t ndim = 8, nwalkers = 16;
        std::vector<std::vector<double>> ensemble(nwalkers, std::vector<double>(ndim));
        for (auto &w : ensemble)
            for (auto &x : w) x = 0.9 + 0.2 * lcg(st);
        DeviceRng rng(7);
        LogProb lpf = LogProb::rosenbrock();
        std::vector<double> logprob(nwalkers);
        ensemble_sample::init_logprob(lpf, ensemble, logprob, rng);
        UpdateFlagSpec ufs = UpdateFlagSpec::Prob(0.5);
        for (int k = 0; k < niter; ++k) ensemble_sample::sample(lpf, ensemble, logprob, rng, 2.0, ufs);
        for (size_t i = 0; i < nwalkers; ++i) printf("A %zu %.17g %.17g %.17g\n", i, ensemble[i][0], ensemble[i][ndim - 1], logprob[i]);
    }
    {   // sample_rastrigin.rs: 8 rungs x 4 walkers, 20-D, swap every 10th iteration, Prob(0.2)
        const size_t nbetas = 8, per = 4, ndims = 20;
        std::vector<double> blist;
        for (size_t i = 0; i < nbetas; ++i) blist.push_back(1.0 / (double)(1u << i));
        std::vector<std::vector<double>> ensemble(nbetas * per, std::vector<double>(ndims));
        for (auto &w : ensemble)
            for (auto &x : w) x = -0.01 + 0.02 * lcg(st);
        DeviceRng rng(11);
        LogProb lpf = LogProb::rastrigin();
        std::vector<double> logprob(ensemble.size());
        ensemble_sample::init_logprob(lpf, ensemble, logprob, rng);
        UpdateFlagSpec ufs = UpdateFlagSpec::Prob(0.2);
        for (int k = 0; k < niter; ++k) {
            if (k % 10 == 0) utils::swap_walkers(ensemble, logprob, rng, blist, lpf);
            ensemble_sample::sample_pt(lpf, ensemble, logprob, rng, 2.0, ufs, blist);
        }
        for (size_t i = 0; i < ensemble.size(); ++i) printf("B %zu %.17g %.17g %.17g\n", i, ensemble[i][0], ensemble[i][ndims - 1], logprob[i]);
    }
    {   // the t-walk through the same host containers: twalk::sample(flogprob, &mut (ensemble, logprob), param, rng, beta, nthreads)
        const size_t nbetas = 2, per = 8, ndims = 6;
        std::vector<double> blist = {1.0, 0.5};
        std::pair<std::vector<std::vector<double>>, std::vector<double>> el;
        el.first.assign(nbetas * per, std::vector<double>(ndims));
        for (auto &w : el.first)
            for (auto &x : w) x = -1.0 + 2.0 * lcg(st);
        el.second.resize(el.first.size());
        DeviceRng rng(13);
        LogProb lpf = LogProb::rosenbrock();
        ensemble_sample::init_logprob(lpf, el.first, el.second, rng);
        twalk::TWalkParams param = twalk::TWalkParams::make(ndims).with_pphi(0.5);
        for (int k = 0; k < niter; ++k) twalk::sample(lpf, el, param, rng, blist, 4);
        for (size_t i = 0; i < el.first.size(); ++i) printf("T %zu %.17g %.17g %.17g\n", i, el.first[i][0], el.first[i][ndims - 1], el.second[i]);
    }
    {   // ensemble-PT HMC: hmc::naive::sample_ensemble_pt(flogprob, grad, q0, lp, rng, epsilon, beta_list, l, param)
        const size_t nbetas = 2, per = 5, ndims = 4;   // odd rungs: nothing pairs up in HMC
        std::vector<double> blist = {1.0, 0.5}, epsilon = {0.01, 0.01};
        std::vector<std::vector<double>> q0(nbetas * per, std::vector<double>(ndims));
        for (auto &w : q0)
            for (auto &x : w) x = 0.9 + 0.2 * lcg(st);
        std::vector<double> lp(q0.size());
        DeviceRng rng(17);
        LogProb lpf = LogProb::rosenbrock();
        ensemble_sample::init_logprob(lpf, q0, lp, rng);
        hmc::naive::HmcParam param = hmc::naive::HmcParam::quick_adj(0.7);
        size_t accepted = 0;
        for (int k = 0; k < niter; ++k)
            for (size_t c : hmc::naive::sample_ensemble_pt(lpf, q0, lp, rng, epsilon, blist, 3, param)) accepted += c;
        for (size_t i = 0; i < q0.size(); ++i) printf("H %zu %.17g %.17g %.17g\n", i, q0[i][0], q0[i][ndims - 1], lp[i]);
        printf("HE %.17g %.17g %zu\n", epsilon[0], epsilon[1], accepted);
    }
    {   // the reference's panics become exceptions
        std::vector<std::vector<double>> e(6, std::vector<double>(2, 0.1));
        std::vector<double> lp(6, 0.0);
        DeviceRng rng(1);
        UpdateFlagSpec all = UpdateFlagSpec::All();
        try { ensemble_sample::sample_pt(LogProb::gaussian(), e, lp, rng, 2.0, all, {1.0, 0.5}); printf("C no-throw\n"); }
        catch (const McmcError &err) { printf("C %d\n", err.code); }   // 3 walkers per rung -> NWalkersIsNotEven
        try { utils::swap_walkers(e, lp, rng, {0.5, 1.0}); printf("D no-throw\n"); }
        catch (const McmcError &err) { printf("D %d\n", err.code); }   // BetaNotInDecrOrd
    }
    return 0;
}
