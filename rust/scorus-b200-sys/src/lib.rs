//! Raw bindings of `include/scorus_b200.h` (ABI version 1).  One declaration per C entry point;
//! the comment on each names the reference function it replaces.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct scorus_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct scorus_cfg {
    pub n_walkers: u64,
    pub dim: u32,
    pub device: i32,
    pub seed: u64,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct scorus_ufs {
    pub kind: i32,
    pub reserved: i32,
    pub prob: f64,
    pub mask: *const u8,
    pub mask_len: usize,
}

pub const SCORUS_OK: c_int = 0;
pub const SCORUS_UFS_PROB: i32 = 0;
pub const SCORUS_UFS_ALL: i32 = 1;
pub const SCORUS_UFS_ONE_HOT_CYCLE: i32 = 2;
pub const SCORUS_UFS_MASK: i32 = 3;
pub const SCORUS_CFG_SEQUENTIAL_SUM: u32 = 1;
pub const SCORUS_CFG_F32: u32 = 2;

/// HmcParam, src/mcmc/hmc/naive.rs:17-23
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct scorus_hmc_param { pub target_accept_ratio: f64, pub adj_factor: f64 }

/// TWalkParams, src/mcmc/twalk.rs:75-83 (fw cumulative)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct scorus_twalk_params { pub aw: f64, pub at: f64, pub pphi: f64, pub fw: [f64; 2] }

extern "C" {
    pub fn scorus_abi_version() -> c_int;
    pub fn scorus_status_string(status: c_int) -> *const c_char;
    pub fn scorus_device_count() -> c_int;
    /// asserts of src/mcmc/ensemble_sample.rs:127-133
    pub fn scorus_check_sample_args(n_walkers: usize, logprob_len: usize, nbeta: usize) -> c_int;
    /// panic of src/mcmc/utils.rs:91-95
    pub fn scorus_check_beta_order(beta_list: *const f64, nbeta: usize) -> c_int;
    pub fn scorus_ctx_create(out: *mut *mut scorus_ctx, cfg: *const scorus_cfg) -> c_int;
    pub fn scorus_ctx_destroy(ctx: *mut scorus_ctx);
    pub fn scorus_last_error(ctx: *const scorus_ctx) -> *const c_char;
    /// replaces the `flogprob: &F` argument (src/mcmc/ensemble_sample.rs:88,102)
    pub fn scorus_set_logprob(ctx: *mut scorus_ctx, kind: c_int, params: *const f64, nparams: usize) -> c_int;
    pub fn scorus_upload(ctx: *mut scorus_ctx, ensemble: *const f64, logprob: *const f64) -> c_int;
    pub fn scorus_download(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64) -> c_int;
    /// init_logprob, src/mcmc/ensemble_sample.rs:77-85
    pub fn scorus_init_logprob(ctx: *mut scorus_ctx) -> c_int;
    /// sample, src/mcmc/ensemble_sample.rs:87-105
    pub fn scorus_sample(ctx: *mut scorus_ctx, a: f64, ufs: *const scorus_ufs, nsteps: u64) -> c_int;
    /// sample_pt, src/mcmc/ensemble_sample.rs:107-190
    pub fn scorus_sample_pt(ctx: *mut scorus_ctx, a: f64, ufs: *const scorus_ufs, beta_list: *const f64,
                            nbeta: usize, nsteps: u64) -> c_int;
    pub fn scorus_sample_pt_fixed(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize, partner: *const u32,
                                  z: *const f64, flags: *const u8, flag_len: usize, u: *const f64,
                                  accepted_out: *mut u8) -> c_int;
    /// draw_z, src/mcmc/utils.rs:33-45
    pub fn scorus_draw_z(ctx: *mut scorus_ctx, a: f64, count: usize, out: *mut f64) -> c_int;
    /// swap_walkers, src/mcmc/utils.rs:72-112
    pub fn scorus_swap_walkers(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize) -> c_int;
    pub fn scorus_swap_walkers_fixed(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize, jvec: *const u32,
                                     r: *const f64, swapped_out: *mut u8) -> c_int;
    pub fn scorus_sample_pt_host(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64, logprob_len: usize,
                                 a: f64, ufs: *const scorus_ufs, beta_list: *const f64, nbeta: usize) -> c_int;
    pub fn scorus_swap_walkers_host(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64,
                                    logprob_len: usize, beta_list: *const f64, nbeta: usize) -> c_int;
    pub fn scorus_sample_pt_host_n(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64, logprob_len: usize,
                                   a: f64, ufs: *const scorus_ufs, beta_list: *const f64, nbeta: usize, nsteps: u64) -> c_int;
    /// get_one_init_realization, src/mcmc/init_ensemble.rs:19-32, for every walker
    pub fn scorus_init_ensemble(ctx: *mut scorus_ctx, lo: *const f64, hi: *const f64) -> c_int;
    /// mean / cov, src/linear_space/utils.rs:4-15, 17-41
    pub fn scorus_mean(ctx: *mut scorus_ctx, first: usize, count: usize, mean_out: *mut f64) -> c_int;
    pub fn scorus_cov(ctx: *mut scorus_ctx, first: usize, count: usize, cov_out: *mut f64) -> c_int;
    /// the recording loops of examples/sample_high_dim.rs:84-86, on the device
    pub fn scorus_trace_enable(ctx: *mut scorus_ctx, walkers: *const u32, coords: *const u32, ncapture: usize,
                               capacity_steps: usize) -> c_int;
    pub fn scorus_trace_download(ctx: *mut scorus_ctx, out: *mut f64, max_steps: usize, nrecorded: *mut usize) -> c_int;
    pub fn scorus_trace_stride(ctx: *mut scorus_ctx, every: u64) -> c_int;
    /// chain statistics on the device
    pub fn scorus_get_rung_stats(ctx: *mut scorus_ctx, nrungs: usize, accepted: *mut u64, proposed: *mut u64,
                                 swapped: *mut u64, swap_proposed: *mut u64) -> c_int;
    pub fn scorus_moments_enable(ctx: *mut scorus_ctx, nrungs: usize, every: u64, with_cov: c_int) -> c_int;
    pub fn scorus_moments_download(ctx: *mut scorus_ctx, rung: usize, mean_out: *mut f64, cov_out: *mut f64,
                                   nsamples: *mut u64) -> c_int;
    pub fn scorus_trace_iat(ctx: *mut scorus_ctx, c_window: f64, max_lag: usize, tau_out: *mut f64,
                            window_out: *mut u32, mean_out: *mut f64, var_out: *mut f64) -> c_int;
    /// the ensemble / parallel-tempering t-walk, src/mcmc/twalk.rs:586-762
    pub fn scorus_twalk_params_default(dim: usize, out: *mut scorus_twalk_params) -> c_int;
    pub fn scorus_twalk_sample(ctx: *mut scorus_ctx, tw: *const scorus_twalk_params, beta_list: *const f64, nbeta: usize,
                               nsteps: u64) -> c_int;
    pub fn scorus_twalk_half_fixed(ctx: *mut scorus_ctx, tw: *const scorus_twalk_params, beta_list: *const f64,
                                   nbeta: usize, order: *const u32, half: c_int, kernel: *const u8, b: *const f64,
                                   flags: *const u8, walk_u: *const f64, u: *const f64, accepted_out: *mut u8) -> c_int;
    pub fn scorus_twalk_sample_host(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64, logprob_len: usize,
                                    tw: *const scorus_twalk_params, beta_list: *const f64, nbeta: usize, nsteps: u64) -> c_int;
    pub fn scorus_twalk_draw_b(ctx: *mut scorus_ctx, at: f64, step: u64, count: usize, out: *mut f64) -> c_int;
    /// ensemble / parallel-tempering HMC, src/mcmc/hmc/naive.rs:122-292
    pub fn scorus_hmc_sample_ensemble_pt(ctx: *mut scorus_ctx, epsilon: *mut f64, beta_list: *const f64, nbeta: usize,
                                         l: usize, param: *const scorus_hmc_param, accepted_cnt: *mut u64, nsteps: u64) -> c_int;
    pub fn scorus_hmc_sample_ensemble_pt_fixed(ctx: *mut scorus_ctx, epsilon: *mut f64, beta_list: *const f64, nbeta: usize,
                                               l: usize, param: *const scorus_hmc_param, p0: *const f64, u1: *const f64,
                                               u2: *const f64, accepted_out: *mut u8, accepted_cnt: *mut u64) -> c_int;
    pub fn scorus_hmc_sample_ensemble_pt_host(ctx: *mut scorus_ctx, q0: *mut f64, logprob: *mut f64, logprob_len: usize,
                                              epsilon: *mut f64, beta_list: *const f64, nbeta: usize, l: usize,
                                              param: *const scorus_hmc_param, accepted_cnt: *mut u64) -> c_int;
    pub fn scorus_hmc_draw_momenta(ctx: *mut scorus_ctx, step: u64, count: usize, out: *mut f64) -> c_int;
    /// one process per GPU: shards of one ensemble
    pub fn scorus_set_shard(ctx: *mut scorus_ctx, walker_offset: u64, rung_offset: u32, n_walkers_global: u64) -> c_int;
    pub fn scorus_sample_pt_global(ctx: *mut scorus_ctx, a: f64, ufs: *const scorus_ufs, beta_list: *const f64,
                                   nbeta: usize, ens_all: *const f64, lp_all: *const f64) -> c_int;
    pub fn scorus_swap_plan_device(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize, lp_all: *mut f64,
                                   src_all: *mut u32) -> c_int;
    pub fn scorus_ipc_export(ctx: *mut scorus_ctx, handle_ens_64b: *mut c_void, handle_lp_64b: *mut c_void) -> c_int;
    pub fn scorus_ipc_attach(ctx: *mut scorus_ctx, world: c_int, rank: c_int, handles_ens: *const c_void,
                             handles_lp: *const c_void) -> c_int;
    pub fn scorus_sample_pt_peer(ctx: *mut scorus_ctx, a: f64, ufs: *const scorus_ufs, beta_list: *const f64,
                                 nbeta: usize, nsteps: u64) -> c_int;
    pub fn scorus_sample_pt_peer_host(ctx: *mut scorus_ctx, ensemble: *mut f64, logprob: *mut f64, logprob_len: usize,
                                      a: f64, ufs: *const scorus_ufs, beta_list: *const f64, nbeta: usize,
                                      nsteps: u64) -> c_int;
    pub fn scorus_peer_barrier(ctx: *mut scorus_ctx) -> c_int;
    pub fn scorus_hmc_adapt_net(ctx: *mut scorus_ctx, net_out: *mut i64) -> c_int;
    pub fn scorus_upload_f32(ctx: *mut scorus_ctx, ensemble: *const f32, logprob: *const f32) -> c_int;
    pub fn scorus_download_f32(ctx: *mut scorus_ctx, ensemble: *mut f32, logprob: *mut f32) -> c_int;
    pub fn scorus_sample_pt_fixed_f32(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize, partner: *const u32,
                                      z: *const f32, flags: *const u8, flag_len: usize, u: *const f32,
                                      accepted_out: *mut u8) -> c_int;
    pub fn scorus_swap_walkers_fixed_f32(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize, jvec: *const u32,
                                         r: *const f32, swapped_out: *mut u8) -> c_int;
    pub fn scorus_swap_walkers_peer(ctx: *mut scorus_ctx, beta_list: *const f64, nbeta: usize) -> c_int;
    pub fn scorus_peer_read_bandwidth(ctx: *mut scorus_ctx, peer: c_int, bytes: usize, iters: c_int,
                                      gbps_copy_engine: *mut f64, gbps_sm_loads: *mut f64, gbps_random_rows: *mut f64) -> c_int;
    pub fn scorus_set_matching_blocks(ctx: *mut scorus_ctx, nblocks: u32, block_offset: u32, global_every: u64) -> c_int;
    pub fn scorus_swap_apply_p2p(ctx: *mut scorus_ctx, src_all: *const u32, lp_all: *const f64, phase: c_int) -> c_int;
    pub fn scorus_sync(ctx: *mut scorus_ctx) -> c_int;
    pub fn scorus_set_stream(ctx: *mut scorus_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn scorus_reset_stream(ctx: *mut scorus_ctx) -> c_int;
    pub fn scorus_get_stream(ctx: *mut scorus_ctx) -> *mut c_void;
    pub fn scorus_device_buffers(ctx: *mut scorus_ctx, ensemble: *mut *mut f64, pitch_elems: *mut usize,
                                 logprob: *mut *mut f64) -> c_int;
    pub fn scorus_get_counters(ctx: *const scorus_ctx, step: *mut u64, swap_call: *mut u64, onehot: *mut u64) -> c_int;
    pub fn scorus_set_counters(ctx: *mut scorus_ctx, step: u64, swap_call: u64, onehot: u64) -> c_int;
    pub fn scorus_get_stats(ctx: *mut scorus_ctx, accepted: *mut u64, proposed: *mut u64, swapped: *mut u64,
                            swap_proposed: *mut u64) -> c_int;
    pub fn scorus_launch_count(ctx: *const scorus_ctx) -> u64;
    pub fn scorus_host_alloc(bytes: usize) -> *mut c_void;
    pub fn scorus_host_free(p: *mut c_void);
}
