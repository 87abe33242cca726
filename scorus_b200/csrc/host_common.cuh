// host_common.cuh -- what the translation units of the C ABI share: the context, error plumbing and the
// helpers that several entry points use.  Private to scorus_b200/csrc (the public surface is
// include/scorus_b200.h).  There is deliberately no CPU path anywhere behind it.
#pragma once
#include "../../include/scorus_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "launch.cuh"
#include "swap_params.cuh"
#include "tile_common.cuh"
#include "twalk_kernel.cuh"

using namespace scorus;

struct scorus_ctx {
    int device = 0;
    uint64_t n = 0;
    uint32_t dim = 0, pitch = 0;
    uint64_t seed = 0;
    uint64_t step = 0, swap_call = 0, onehot = 0;
    uint64_t w_off = 0, n_global = 0;  // shard position in the whole ensemble (scorus_set_shard)
    uint32_t rung_off = 0;
    uint64_t launches = 0, proposed = 0, swap_proposed = 0;
    int num_sms = 0, smem_optin = 0;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    double *d_ens = nullptr, *d_lp = nullptr, *d_scratch = nullptr;
    double *d_beta = nullptr;   // the beta list of the current call (one of beta_cache)
    struct BetaEntry { std::vector<double> host; double *dev = nullptr; uint64_t used = 0; };
    std::vector<BetaEntry> beta_cache;  // callers alternate between a few lists (local / global ladder)
    uint64_t beta_clock = 0;
    double *d_params = nullptr; size_t params_cap = 0;
    uint32_t *d_order = nullptr, *d_src = nullptr;
    uint32_t *d_gorder = nullptr; size_t gorder_cap = 0;   // global-matching mode: exact order of the whole ensemble
    uint32_t *d_pairs = nullptr;  size_t pairs_cap = 0;    // ... local pair list
    uint32_t *d_count = nullptr;
    // one-sided partner reads over NVLink (scorus_ipc_attach): peers' ensembles / log-probs mapped here
    int world = 1, rank = 0;
    double *peer_ens[16] = {nullptr}, *peer_lp[16] = {nullptr};
    double **d_peer_ens = nullptr, **d_peer_lp = nullptr;   // device copies of the tables
    unsigned long long bar_epoch = 0;                       // inter-rank flag barrier (api_shard.cu): calls so far
    // blocked matchings (scorus_set_matching_blocks): blocks per rung on "local" steps, key offset of block 0,
    // and the schedule: global_every = k >= 1 -> steps with step % k == 0 match whole rungs; 0 -> never
    uint32_t mblocks = 1, mblk_off = 0;
    uint64_t global_every = 0;
    // K6: statistics scratch and the chain trace
    double *d_part = nullptr; size_t part_cap = 0;
    double *d_stat = nullptr; size_t stat_cap = 0;
    uint32_t *d_tr_walker = nullptr, *d_tr_coord = nullptr;
    double *d_trace = nullptr;
    size_t tr_ncap = 0, tr_capacity = 0, tr_count = 0;
    uint64_t tr_every = 1, tr_tick = 0;                     // capture after every tr_every-th step
    // chain statistics: per-rung accepted moves [0, rung_cap) and per-stage accepted exchanges [rung_cap, 2 rung_cap)
    unsigned long long *d_rung_stats = nullptr; size_t rung_cap = 0;
    std::vector<uint64_t> rung_proposed, stage_proposed;
    // running per-rung moments: sums of (x - shift) and of its outer products, taken every mom_every-th step
    double *d_mom = nullptr;   // [shift: rungs*dim][s1: rungs*dim][s2: rungs*dim*dim]
    double *d_iat = nullptr; size_t iat_cap = 0;
    uint32_t mom_rungs = 0; bool mom_cov = false;
    uint64_t mom_every = 1, mom_tick = 0, mom_count = 0;
    unsigned long long *d_stats = nullptr;
    // scratch for caller-supplied streams
    double *d_z = nullptr, *d_u = nullptr;
    uint8_t *d_flags = nullptr; size_t flags_cap = 0;
    uint32_t *d_flag_cdf = nullptr; double flag_cdf_prob = -1.0;  // Prob(p) at 32 bits: thresholds of the first set flag
    uint8_t *d_acc = nullptr;
    uint8_t *d_dirty = nullptr;     // host-buffer call: walkers accepted since the upload
    unsigned long long *d_tile_counter = nullptr;  // dynamic tile scheduling of the register kernel
    unsigned long long tile_ticket = 0;
    bool track_dirty = false;
    uint32_t *d_jinv = nullptr; size_t jinv_cap = 0;
    double *d_r = nullptr;      size_t r_cap = 0;
    uint8_t *d_swapped = nullptr; size_t swapped_cap = 0;
    double *d_lp_all = nullptr; size_t lp_all_cap = 0;          // sharded ladder swap: the plan for all slots
    uint32_t *d_src_all = nullptr; size_t src_all_cap = 0;
    void *d_chain = nullptr; size_t chain_cap = 0;              // records of the ladder swap's sub-chains (swap_kernel.cuh)
    // A single-GPU f64 swap leaves its result interleaved in d_swap_out and the rows where they were: the stretch step
    // that follows reads through it and writes every row to its new slot itself (step_reg_kernel.cuh, kRemap); anything
    // else that looks at the ensemble or the log-probs first calls settle_swap (api_core.cu), which moves the rows.
    void *d_swap_out = nullptr; size_t swap_out_cap = 0;
    bool swap_pending = false;
    // Host-buffer call (api_host.cu) on pinned buffers: device pointers of the caller's ensemble / log-probs for the
    // FIRST step of the call, which streams the rows from there instead of an upload (step_reg_kernel.cuh, kRemap
    // variant); consumed (reset) by that step
    double *host_ens = nullptr, *host_lp = nullptr;
    int8_t *d_hmc_adapt = nullptr;                              // HMC: step-size adaptation codes, one per walker
    double *d_hmc_eps = nullptr; size_t hmc_eps_cap = 0;        // ... step size per rung
    double *d_hmc_p = nullptr; size_t hmc_p_cap = 0;            // ... caller-supplied momenta
    uint8_t *d_tw_kernel = nullptr; size_t tw_kernel_cap = 0;   // t-walk, caller-supplied draws
    double *d_tw_walk = nullptr;    size_t tw_walk_cap = 0;

    int lp_kind = -1;
    uint32_t nparams = 0;
    bool uploaded = false;
    bool sequential = false;  // SCORUS_CFG_SEQUENTIAL_SUM
    bool f32 = false;         // SCORUS_CFG_F32: d_ens / d_lp hold float, kernels of namespace scorus::f32 (api_f32.cu)
    uint32_t elem = 8;        // bytes per coordinate
    float *d_beta_f32 = nullptr; size_t beta_f32_cap = 0;   // f32 contexts: the beta list as float
    std::vector<float> beta_f32_host;
    std::vector<double> beta_f32_src;
    std::string err;
};

inline int fail(scorus_ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

// entry points without an f32 instantiation
#define F64_ONLY(ctx)                                                                                          \
    do {                                                                                                       \
        if ((ctx) && (ctx)->f32)                                                                               \
            return fail((ctx), SCORUS_ERR_INVALID_ARG, "%s is not available on an f32 context", __func__);     \
    } while (0)

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, SCORUS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

namespace scorus {
namespace host {

// api_core.cu
int env_int(const char *name, int dflt);
int ensure(scorus_ctx *ctx, void **buf, size_t *cap, size_t bytes);
int step_geometry(scorus_ctx *ctx, Geometry *g, bool tile_in_smem);
int upload_beta(scorus_ctx *ctx, const double *beta, size_t nbeta);
void stamp_tiles(scorus_ctx *ctx, StepParams *p, const Geometry &g);
void commit_tiles(scorus_ctx *ctx, const StepParams &p);  // after the launch succeeded: advance the host's ticket
// matching blocks of step `step` under the context's schedule -> p.mblk / mblk_bits / mblk_off
void set_matching(const scorus_ctx *ctx, StepParams *p, uint64_t step);
uint32_t flag_bits_for(double prob);
// UpdateFlagSpec::Prob(p) -> p.flag_bits / p.flag_thr (the one place that quantises p; INVALID_ARG unless p > 0)
int set_prob_flags(scorus_ctx *ctx, StepParams *p, double prob);
int ensure_rung_stats(scorus_ctx *ctx, size_t nbeta);
void count_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb, uint64_t w_lo, uint64_t w_hi, uint64_t steps);
void count_swap_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb);
// tile_in_smem: the kernel about to be launched stages a 32-row tile of proposals per warp (t-walk; the dense target)
// keep_pending_swap: the caller handles a pending swap itself (the stretch step: fused, or settled there)
int fill_common(scorus_ctx *ctx, StepParams *p, Geometry *g, size_t nbeta, bool tile_in_smem = false,
                bool keep_pending_swap = false);
// kernels of swap_kernel.cuh, launched from more than one translation unit
void launch_exact_order(scorus_ctx *ctx, uint64_t step, size_t nbeta, uint32_t npb, uint32_t *order, uint32_t w_off);
void launch_exact_jvec(scorus_ctx *ctx, size_t nbeta, size_t npb, uint32_t *jinv);
int launch_swap_chain(scorus_ctx *ctx, SwapParams &sp);
// rows and log-probs of a pending swap to their slots (no-op when nothing is pending)
int settle_swap(scorus_ctx *ctx);
// whether the stretch step of this context can take its rows from somewhere else and write all of them (the kRemap
// variant of the register kernel: after a swap, or streaming from a pinned host buffer)
bool stretch_step_moves_rows(scorus_ctx *ctx, size_t nbeta);
// api_f32.cu: the T = f32 paths behind the shared entry points
int f32_set_params(scorus_ctx *ctx, const double *params, size_t nparams);
int f32_init_logprob(scorus_ctx *ctx);
int f32_sample_pt(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta, size_t nbeta, uint64_t nsteps);
int f32_swap_walkers(scorus_ctx *ctx, const double *beta, size_t nbeta);
// api_stats.cu
int after_step(scorus_ctx *ctx);  // chain capture and running moments, after every step
int launch_scatter_download(scorus_ctx *ctx, double *host_ens, double *host_lp);

}  // namespace host
}  // namespace scorus

using namespace scorus::host;
