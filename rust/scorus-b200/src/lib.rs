//! Safe wrapper with the reference's names and argument order (NOT compiled in this repo's image:
//! no Rust toolchain; see rust/README.md).
//!
//! ```text
//! reference (src/mcmc/ensemble_sample.rs:107)              this crate
//! sample_pt(&flogprob, &mut ensemble, &mut lp,             sample_pt(&flogprob, &mut ensemble, &mut lp,
//!           &mut rng, a, &mut ufs, &beta_list)                       &mut rng, a, &mut ufs, &beta_list)
//!   flogprob: &F where F: Fn(&V) -> T                        flogprob: &LogProb   (device plugin id + params)
//!   rng: &mut U where U: Rng                                 rng: &mut DeviceRng  (seed; owns the device context)
//! ```
use scorus::linear_space::IndexableLinearSpace;
use scorus_b200_sys as sys;
use std::collections::HashMap;
use std::ffi::CStr;

/// src/mcmc/mcmc_errors.rs:1-7 (the reference declares it and then panics instead)
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum McmcErr {
    NWalkersIsZero = 1,
    NWalkersIsNotEven = 2,
    NWalkersMismatchesNBeta = 3,
    BetaNotInDecrOrd = 4,
}

fn check(code: i32, ctx: *const sys::scorus_ctx) {
    if code == sys::SCORUS_OK {
        return;
    }
    // the reference panics (assert! at ensemble_sample.rs:130-133, panic! at utils.rs:93-95): so do we
    let what = unsafe { CStr::from_ptr(sys::scorus_status_string(code)) }.to_string_lossy().into_owned();
    let detail = if ctx.is_null() { String::new() } else {
        unsafe { CStr::from_ptr(sys::scorus_last_error(ctx)) }.to_string_lossy().into_owned()
    };
    panic!("scorus-b200: {what} {detail}");
}

/// Device-side log-density (replaces `F: Fn(&V) -> T + Send + Sync`).
#[derive(Clone, Debug)]
pub struct LogProb { pub kind: i32, pub params: Vec<f64> }
impl LogProb {
    pub fn gaussian(c: f64) -> Self { Self { kind: 0, params: vec![c] } }          // examples/test_emcee.rs:18-24
    pub fn bounded_gaussian(limit: f64) -> Self { Self { kind: 1, params: vec![limit] } }
    pub fn rosenbrock() -> Self { Self { kind: 2, params: vec![] } }               // examples/test_emcee.rs:26-32
    pub fn rastrigin(a: f64) -> Self { Self { kind: 3, params: vec![a] } }         // examples/sample_rastrigin.rs:17-25
    pub fn bimodal2d() -> Self { Self { kind: 4, params: vec![] } }                // examples/sample_bimodal.rs:46-54
}

/// src/mcmc/ensemble_sample.rs:25-32
pub enum UpdateFlagSpec<'a> {
    Prob(f64),
    All,
    Func(&'a mut dyn FnMut() -> Vec<bool>),
    /// the cycling closure of examples/sample_high_dim.rs:59-65, evaluated on the device
    OneHotCycle,
}

/// RAII owner of the device buffers (`scorus_ctx`).
pub struct Sampler { ctx: *mut sys::scorus_ctx, n: usize, dim: usize }
impl Sampler {
    pub fn new(flogprob: &LogProb, n: usize, dim: usize, seed: u64) -> Self {
        let cfg = sys::scorus_cfg { n_walkers: n as u64, dim: dim as u32, device: -1, seed, flags: 0, reserved: 0 };
        let mut ctx = std::ptr::null_mut();
        check(unsafe { sys::scorus_ctx_create(&mut ctx, &cfg) }, std::ptr::null());
        let s = Self { ctx, n, dim };
        s.set_logprob(flogprob);
        s
    }
    pub fn set_logprob(&self, f: &LogProb) {
        check(unsafe { sys::scorus_set_logprob(self.ctx, f.kind, f.params.as_ptr(), f.params.len()) }, self.ctx);
    }
    fn with_ufs<R>(&self, ufs: &mut UpdateFlagSpec, body: impl FnOnce(&sys::scorus_ufs) -> R) -> R {
        let mut mask: Vec<u8> = Vec::new();
        let mut c = sys::scorus_ufs { kind: sys::SCORUS_UFS_ALL, reserved: 0, prob: 0.0, mask: std::ptr::null(), mask_len: 0 };
        match ufs {
            UpdateFlagSpec::Prob(p) => { c.kind = sys::SCORUS_UFS_PROB; c.prob = *p; }
            UpdateFlagSpec::All => {}
            UpdateFlagSpec::OneHotCycle => { c.kind = sys::SCORUS_UFS_ONE_HOT_CYCLE; }
            UpdateFlagSpec::Func(f) => {
                // called once per walker, in walker order, as ensemble_sample.rs:150-153 does
                let mut len = 0;
                for i in 0..self.n {
                    let v = f();
                    if i == 0 { len = v.len(); }
                    assert!(v.len() == len && len <= self.dim);
                    mask.extend(v.into_iter().map(|b| b as u8));
                }
                c.kind = sys::SCORUS_UFS_MASK; c.mask = mask.as_ptr(); c.mask_len = len;
            }
        }
        body(&c)
    }
    pub fn sample_pt_host(&self, ens: &mut [f64], lp: &mut [f64], a: f64, ufs: &mut UpdateFlagSpec, beta: &[f64]) {
        let (ctx, n_lp) = (self.ctx, lp.len());
        let (pe, pl) = (ens.as_mut_ptr(), lp.as_mut_ptr());
        self.with_ufs(ufs, |c| check(unsafe {
            sys::scorus_sample_pt_host(ctx, pe, pl, n_lp, a, c, beta.as_ptr(), beta.len()) }, ctx));
    }
    pub fn twalk_sample_host(&self, ens: &mut [f64], lp: &mut [f64], param: &sys::scorus_twalk_params, beta: &[f64]) {
        check(unsafe { sys::scorus_twalk_sample_host(self.ctx, ens.as_mut_ptr(), lp.as_mut_ptr(), lp.len(), param,
                                                     beta.as_ptr(), beta.len(), 1) }, self.ctx);
    }
    pub fn hmc_sample_ensemble_pt_host(&self, q0: &mut [f64], lp: &mut [f64], epsilon: &mut [f64], beta: &[f64], l: usize,
                                       param: &sys::scorus_hmc_param) -> Vec<usize> {
        assert_eq!(epsilon.len(), beta.len());   // naive.rs:179
        let mut cnt = vec![0u64; beta.len()];
        check(unsafe { sys::scorus_hmc_sample_ensemble_pt_host(self.ctx, q0.as_mut_ptr(), lp.as_mut_ptr(), lp.len(),
                                                               epsilon.as_mut_ptr(), beta.as_ptr(), beta.len(), l, param,
                                                               cnt.as_mut_ptr()) }, self.ctx);
        cnt.into_iter().map(|c| c as usize).collect()
    }
    pub fn swap_walkers_host(&self, ens: &mut [f64], lp: &mut [f64], beta: &[f64]) {
        check(unsafe { sys::scorus_swap_walkers_host(self.ctx, ens.as_mut_ptr(), lp.as_mut_ptr(), lp.len(),
                                                     beta.as_ptr(), beta.len()) }, self.ctx);
    }
}
impl Drop for Sampler { fn drop(&mut self) { unsafe { sys::scorus_ctx_destroy(self.ctx) } } }

/// Replaces `rng: &mut U`: the seed of the counter-based device streams; caches one context per shape.
pub struct DeviceRng { pub seed: u64, cache: HashMap<(usize, usize), Sampler> }
impl DeviceRng {
    pub fn new(seed: u64) -> Self { Self { seed, cache: HashMap::new() } }
    fn sampler(&mut self, f: &LogProb, n: usize, dim: usize) -> &Sampler {
        let seed = self.seed;
        let s = self.cache.entry((n, dim)).or_insert_with(|| Sampler::new(f, n, dim, seed));
        s.set_logprob(f);
        s
    }
}

fn flatten<V: IndexableLinearSpace<f64>>(ensemble: &[V]) -> (Vec<f64>, usize) {
    let dim = ensemble[0].dimension();
    let mut flat = Vec::with_capacity(ensemble.len() * dim);
    for w in ensemble { for k in 0..dim { flat.push(w[k]); } }
    (flat, dim)
}
fn unflatten<V: IndexableLinearSpace<f64>>(flat: &[f64], ensemble: &mut [V], dim: usize) {
    for (i, w) in ensemble.iter_mut().enumerate() { for k in 0..dim { w[k] = flat[i * dim + k]; } }
}

/// src/mcmc/ensemble_sample.rs:107-190
pub fn sample_pt<'a, V>(flogprob: &LogProb, ensemble: &mut [V], cached_logprob: &mut [f64], rng: &mut DeviceRng,
                        a: f64, ufs: &mut UpdateFlagSpec<'a>, beta_list: &[f64])
where V: Clone + IndexableLinearSpace<f64> + Sync + Send {
    check(unsafe { sys::scorus_check_sample_args(ensemble.len(), cached_logprob.len(), beta_list.len()) }, std::ptr::null());
    let (mut flat, dim) = flatten(ensemble);
    rng.sampler(flogprob, ensemble.len(), dim).sample_pt_host(&mut flat, cached_logprob, a, ufs, beta_list);
    unflatten(&flat, ensemble, dim);
}

/// src/mcmc/ensemble_sample.rs:87-105
pub fn sample<'a, V>(flogprob: &LogProb, ensemble: &mut [V], cached_logprob: &mut [f64], rng: &mut DeviceRng, a: f64,
                     ufs: &mut UpdateFlagSpec<'a>)
where V: Clone + IndexableLinearSpace<f64> + Sync + Send {
    sample_pt(flogprob, ensemble, cached_logprob, rng, a, ufs, &[1.0])
}

/// src/mcmc/utils.rs:72-112 (silently does nothing when the lengths differ, :88)
pub fn swap_walkers<V>(ensemble: &mut [V], logprob: &mut [f64], rng: &mut DeviceRng, beta_list: &[f64])
where V: Clone + IndexableLinearSpace<f64> {
    let nbeta = beta_list.len();
    assert!(nbeta * (ensemble.len() / nbeta) == ensemble.len());
    if ensemble.len() != logprob.len() { return; }
    let (mut flat, dim) = flatten(ensemble);
    rng.sampler(&LogProb::gaussian(0.5), ensemble.len(), dim).swap_walkers_host(&mut flat, logprob, beta_list);
    unflatten(&flat, ensemble, dim);
}

/// `mcmc::twalk`, ensemble entry point (src/mcmc/twalk.rs:586-648)
pub mod twalk {
    use super::*;

    /// src/mcmc/twalk.rs:75-130; `fw` holds cumulative weights, as the reference stores them
    #[derive(Clone, Copy, Debug)]
    pub struct TWalkParams { pub aw: f64, pub at: f64, pub pphi: f64, pub fw: [f64; 2] }
    impl TWalkParams {
        pub fn new(n: usize) -> Self {
            let mut c = sys::scorus_twalk_params { aw: 0.0, at: 0.0, pphi: 0.0, fw: [0.0; 2] };
            check(unsafe { sys::scorus_twalk_params_default(n, &mut c) }, std::ptr::null());
            Self { aw: c.aw, at: c.at, pphi: c.pphi, fw: c.fw }
        }
        pub fn with_fw(mut self, fw: [f64; 2]) -> Self { self.fw = [fw[0], fw[0] + fw[1]]; self }
        pub fn with_pphi(self, pphi: f64) -> Self { Self { pphi, ..self } }
    }

    /// twalk::sample: `nthreads` is accepted and ignored (the log-densities are evaluated on the device)
    pub fn sample<V>(flogprob: &LogProb, ensemble_logprob: &mut (Vec<V>, Vec<f64>), param: &TWalkParams,
                     rng: &mut DeviceRng, beta_list: &[f64], _nthreads: usize)
    where V: Clone + IndexableLinearSpace<f64> + Sync + Send {
        let (ensemble, logprob) = (&mut ensemble_logprob.0, &mut ensemble_logprob.1);
        check(unsafe { sys::scorus_check_sample_args(ensemble.len(), logprob.len(), beta_list.len()) }, std::ptr::null());
        let (mut flat, dim) = flatten(ensemble);
        let c = sys::scorus_twalk_params { aw: param.aw, at: param.at, pphi: param.pphi, fw: param.fw };
        rng.sampler(flogprob, ensemble.len(), dim).twalk_sample_host(&mut flat, logprob, &c, beta_list);
        unflatten(&flat, ensemble, dim);
    }
}

/// `mcmc::hmc::naive`, ensemble entry point (src/mcmc/hmc/naive.rs:122-158)
pub mod hmc {
    pub mod naive {
        use super::super::*;

        /// naive.rs:17-56
        #[derive(Clone, Copy, Debug)]
        pub struct HmcParam { pub target_accept_ratio: f64, pub adj_factor: f64 }
        impl HmcParam {
            pub fn new(target_accept_ratio: f64, adj_factor: f64) -> Self { Self { target_accept_ratio, adj_factor } }
            pub fn quick_adj(t: f64) -> Self { Self::new(t, 0.01) }
            pub fn slow_adj(t: f64) -> Self { Self::new(t, 0.000_001) }
            pub fn fixed(t: f64) -> Self { Self::new(t, 0.0) }
        }
        impl Default for HmcParam { fn default() -> Self { Self::new(0.99, 0.000_001) } }

        /// sample_ensemble_pt: the gradient (the reference's `grad_logprob: &G`) is the plugin's own
        pub fn sample_ensemble_pt<V>(flogprob: &LogProb, q0: &mut [V], lp: &mut [f64], rng: &mut DeviceRng,
                                     epsilon: &mut [f64], beta_list: &[f64], l: usize, param: &HmcParam) -> Vec<usize>
        where V: Clone + IndexableLinearSpace<f64> + Sync + Send {
            assert_eq!(q0.len(), lp.len());
            let (mut flat, dim) = flatten(q0);
            let c = sys::scorus_hmc_param { target_accept_ratio: param.target_accept_ratio, adj_factor: param.adj_factor };
            let cnt = rng.sampler(flogprob, q0.len(), dim).hmc_sample_ensemble_pt_host(&mut flat, lp, epsilon, beta_list, l, &c);
            unflatten(&flat, q0, dim);
            cnt
        }
    }
}
