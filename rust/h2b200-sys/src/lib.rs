//! Raw + safe bindings of `include/h2b200.h`.
//!
//! NOTE: this crate is shipped as source.  The build image of this repository has no Rust
//! toolchain (no cargo/rustc, no network), so it has not been compiled here; the C ABI it binds
//! is exercised through the identical ctypes binding in `halo2-prover_b200/_ffi.py`.
//!
//! Layout contract (checked by the static asserts below): `Fr`/`Fq` = `[u64; 4]` Montgomery limbs,
//! `G1Affine` = 64 bytes `{x, y}`, `G1` = 96 bytes `{x, y, z}` -- what halo2curves 0.3.2 stores.
#![cfg(not(target_family = "wasm"))]

use core::ffi::{c_char, c_int, c_void};
use halo2curves::bn256::{Fr, G1Affine, G1};

const _: () = assert!(core::mem::size_of::<Fr>() == 32 && core::mem::align_of::<Fr>() == 8);
const _: () = assert!(core::mem::size_of::<G1Affine>() == 64);
const _: () = assert!(core::mem::size_of::<G1>() == 96);

#[repr(C)]
#[derive(Clone, Copy)]
pub struct H2bDomain {
    pub k: u32,
    pub extended_k: u32,
    pub j: u32,
    pub n_t: u32,
    pub omega: [u64; 4],
    pub omega_inv: [u64; 4],
    pub extended_omega: [u64; 4],
    pub extended_omega_inv: [u64; 4],
    pub g_coset: [u64; 4],
    pub g_coset_inv: [u64; 4],
    pub ifft_divisor: [u64; 4],
    pub extended_ifft_divisor: [u64; 4],
    pub t_evaluations: [u64; 128],
    pub extended_ifft_coset: [u64; 12],
}

extern "C" {
    pub fn h2b_init(device: c_int) -> c_int;
    pub fn h2b_init_devices(devices: *const c_int, count: c_int) -> c_int;
    pub fn h2b_device_count() -> c_int;
    pub fn h2b_srs_layout(srs: u64, parts: *mut u32, replicated: *mut u32, part_n: *mut usize) -> c_int;
    pub fn h2b_shutdown();
    pub fn h2b_last_error() -> *const c_char;
    pub fn h2b_abi_version() -> u32;
    pub fn h2b_best_multiexp(coeffs: *const u64, bases: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn h2b_srs_register(bases: *const u64, n: usize, handle: *mut u64) -> c_int;
    pub fn h2b_srs_release(handle: u64) -> c_int;
    pub fn h2b_commit(srs: u64, scalars: *const u64, n: usize, out: *mut u64) -> c_int;
    pub fn h2b_g1_fold(points: *const u64, count: usize, out: *mut u64) -> c_int;
    pub fn h2b_best_fft(a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn h2b_domain_new(j: u32, k: u32, out: *mut H2bDomain) -> c_int;
    pub fn h2b_lagrange_to_coeff(d: *const H2bDomain, a: *mut u64) -> c_int;
    pub fn h2b_coeff_to_extended(d: *const H2bDomain, input: *const u64, out: *mut u64) -> c_int;
    pub fn h2b_extended_to_coeff(d: *const H2bDomain, input: *const u64, out: *mut u64) -> c_int;
    pub fn h2b_divide_by_vanishing_poly(d: *const H2bDomain, a: *mut u64) -> c_int;
    pub fn h2b_commit_many(srs: u64, polys: *const *const u64, n: usize, m: usize, out: *mut u64) -> c_int;
    pub fn h2b_params_read(bytes: *const u8, len: usize, k: *mut u32, g: *mut u64, g_lagrange: *mut u64) -> c_int;
    pub fn h2b_params_write(k: u32, g: u64, g_lagrange: u64, g2_and_s_g2: *const u8, out: *mut u8, len: usize) -> c_int;
    pub fn h2b_lagrange_to_coeff_many(d: *const H2bDomain, cols: *const *mut u64, m: usize) -> c_int;
    pub fn h2b_coeff_to_extended_many(d: *const H2bDomain, input: *const *const u64, out: *const *mut u64, m: usize) -> c_int;
    pub fn h2b_dev_evaluate_h(d: *const H2bDomain, a: *const c_void /* h2b_eval_h, include/h2b200.h */, values: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_evaluate_h_lookup(d: *const H2bDomain, a: *const c_void, product: *const c_void, permuted_input: *const c_void,
                                     permuted_table: *const c_void, values: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_divide_by_vanishing_poly(d: *const H2bDomain, a: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_g_to_lagrange(g: *const u64, k: u32, out: *mut u64) -> c_int;
    pub fn h2b_g1_to_bytes(points: *const u64, m: usize, out: *mut u8) -> c_int;
    pub fn h2b_dev_msm(c: *const c_void, b: *const c_void, n: usize, out: *mut c_void, stream: *mut c_void) -> c_int;
    // device-resident entry points (polynomials stay in HBM between the phases of create_proof; `stream` = a CUDA stream or null)
    pub fn h2b_dev_srs_register(d_bases: *const c_void, n: usize, handle: *mut u64) -> c_int;
    pub fn h2b_srs_device_ptr(srs: u64, d_bases: *mut *mut c_void, n: *mut usize) -> c_int;
    pub fn h2b_srs_info(srs: u64, n: *mut usize, window_bits: *mut u32, windows: *mut u32, table_bytes: *mut usize) -> c_int;
    pub fn h2b_dev_commit(srs: u64, d_coeffs: *const c_void, n: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_commit_many(srs: u64, d_coeffs: *const c_void, n: usize, m: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_best_fft(d_a: *mut c_void, omega: *const u64, log_n: u32, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_lagrange_to_coeff(d: *const H2bDomain, d_a: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_coeff_to_extended(d: *const H2bDomain, d_in: *const c_void, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_extended_to_coeff(d: *const H2bDomain, d_in: *const c_void, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_lagrange_to_coeff_many(d: *const H2bDomain, d_a: *mut c_void, m: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_coeff_to_extended_many(d: *const H2bDomain, d_in: *const c_void, d_out: *mut c_void, m: usize, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_g1_fold(d_points: *const c_void, count: usize, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn h2b_dev_fixed_base_mul(d_scalars: *const c_void, n: usize, base: *const u64, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    // tuning (defaults are the measured optima)
    pub fn h2b_set_msm_window(c: u32) -> c_int;
    pub fn h2b_set_srs_precompute(enabled: c_int, c: u32) -> c_int;
    pub fn h2b_set_srs_table_stride(t: u32) -> c_int;
    pub fn h2b_set_h2d_bandwidth(gbs: f64) -> c_int;
    pub fn h2b_set_e2e_chunking(chunks: u32, min_n: usize) -> c_int;
    pub fn h2b_kernel_launches() -> u64;
}

fn check(rc: c_int, what: &str) {
    // upstream semantics are panics (assert_eq! at arithmetic.rs:148, :199; domain.rs:227, :244, :311)
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(h2b_last_error()) }.to_string_lossy().into_owned();
        panic!("{what}: h2b200 error {rc}: {msg}");
    }
}

/// `H2B200_DEVICES=0,1,2,3` makes one process use several GPUs (h2b_init_devices: SRS arrays of 2^21 points and more
/// are sharded by point range, smaller ones replicated and whole columns dealt to the devices); `H2B200_DEVICE=n`
/// (default 0) selects a single one.
fn ensure_init() {
    static ONCE: std::sync::Once = std::sync::Once::new();
    ONCE.call_once(|| {
        if let Ok(list) = std::env::var("H2B200_DEVICES") {
            let devs: Vec<c_int> = list.split(',').filter_map(|s| s.trim().parse().ok()).collect();
            assert!(!devs.is_empty(), "H2B200_DEVICES is set but names no device");
            check(unsafe { h2b_init_devices(devs.as_ptr(), devs.len() as c_int) }, "h2b_init_devices");
        } else {
            let dev = std::env::var("H2B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            check(unsafe { h2b_init(dev) }, "h2b_init");
        }
    });
}

/// Drop-in for `halo2_proofs::arithmetic::best_multiexp::<G1Affine>` (arithmetic.rs:147-180).
pub fn best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1 {
    assert_eq!(coeffs.len(), bases.len());
    ensure_init();
    let mut out = core::mem::MaybeUninit::<G1>::uninit();
    check(
        unsafe {
            h2b_best_multiexp(coeffs.as_ptr() as *const u64, bases.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr() as *mut u64)
        },
        "best_multiexp",
    );
    unsafe { out.assume_init() }
}

/// Drop-in for `halo2_proofs::arithmetic::best_fft::<Fr, Fr>` (arithmetic.rs:185-250).
pub fn best_fft(a: &mut [Fr], omega: Fr, log_n: u32) {
    assert_eq!(a.len(), 1 << log_n);
    ensure_init();
    check(unsafe { h2b_best_fft(a.as_mut_ptr() as *mut u64, &omega as *const Fr as *const u64, log_n) }, "best_fft");
}

/// Device-resident SRS for `ParamsKZG::{commit, commit_lagrange}` (kzg/commitment.rs:319, :363).
/// (`ParamsKZG` derives `Debug` and `Clone`: the patch keeps it behind an `Arc`.)
#[derive(Debug)]
pub struct Srs(u64, usize);
impl Srs {
    pub fn register(bases: &[G1Affine]) -> Self {
        ensure_init();
        let mut h = 0u64;
        check(unsafe { h2b_srs_register(bases.as_ptr() as *const u64, bases.len(), &mut h) }, "srs_register");
        Srs(h, bases.len())
    }
    pub fn len(&self) -> usize { self.1 }
    pub fn is_empty(&self) -> bool { self.1 == 0 }
    pub fn commit(&self, scalars: &[Fr]) -> G1 {
        assert!(self.1 >= scalars.len());  // commitment.rs:319 / :363: assert!(bases.len() >= size)
        let mut out = core::mem::MaybeUninit::<G1>::uninit();
        check(unsafe { h2b_commit(self.0, scalars.as_ptr() as *const u64, scalars.len(), out.as_mut_ptr() as *mut u64) }, "commit");
        unsafe { out.assume_init() }
    }
}
impl Srs {
    /// `polys.iter().map(|p| self.commit(p))` in one pass of the kernels: the advice columns, the
    /// permutation / lookup products or the h pieces of one create_proof phase (plonk/prover.rs).
    pub fn commit_many(&self, polys: &[&[Fr]]) -> Vec<G1> {
        let n = polys.first().map_or(0, |p| p.len());
        assert!(polys.iter().all(|p| p.len() == n));
        let ptrs: Vec<*const u64> = polys.iter().map(|p| p.as_ptr() as *const u64).collect();
        let mut out: Vec<G1> = Vec::with_capacity(polys.len());
        check(unsafe { h2b_commit_many(self.0, ptrs.as_ptr(), n, polys.len(), out.as_mut_ptr() as *mut u64) }, "commit_many");
        unsafe { out.set_len(polys.len()) };
        out
    }
    /// `ParamsKZG::read` (SerdeFormat::RawBytes): returns (k, g, g_lagrange) registered from the buffer.
    pub fn read_params(bytes: &[u8]) -> (u32, Srs, Srs) {
        ensure_init();
        let (mut k, mut g, mut gl) = (0u32, 0u64, 0u64);
        check(unsafe { h2b_params_read(bytes.as_ptr(), bytes.len(), &mut k, &mut g, &mut gl) }, "params_read");
        (k, Srs(g, 1 << k), Srs(gl, 1 << k))
    }
    /// `ParamsKZG::write` (SerdeFormat::RawBytes): k | g | g_lagrange read back from HBM | the caller's g2 | s_g2.
    pub fn write_params(k: u32, g: &Srs, g_lagrange: &Srs, g2_and_s_g2: &[u8; 256]) -> Vec<u8> {
        let mut out = vec![0u8; 4 + (128usize << k) + 256];
        check(unsafe { h2b_params_write(k, g.0, g_lagrange.0, g2_and_s_g2.as_ptr(), out.as_mut_ptr(), out.len()) }, "params_write");
        out
    }
}
impl Drop for Srs {
    fn drop(&mut self) {
        unsafe { h2b_srs_release(self.0) };
    }
}

/// Fused EvaluationDomain transforms (domain.rs:227, :244, :311).  `EvaluationDomain` derives `Clone` and `Debug`.
#[derive(Clone, Copy)]
pub struct Domain(H2bDomain);
impl core::fmt::Debug for Domain {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "h2b200::Domain {{ k: {}, extended_k: {}, j: {} }}", self.0.k, self.0.extended_k, self.0.j)
    }
}
impl Domain {
    pub fn new(j: u32, k: u32) -> Self {
        ensure_init();
        let mut d = core::mem::MaybeUninit::<H2bDomain>::uninit();
        check(unsafe { h2b_domain_new(j, k, d.as_mut_ptr()) }, "domain_new");
        Domain(unsafe { d.assume_init() })
    }
    pub fn raw(&self) -> &H2bDomain { &self.0 }
    pub fn lagrange_to_coeff(&self, a: &mut [Fr]) {
        assert_eq!(a.len(), 1 << self.0.k);
        check(unsafe { h2b_lagrange_to_coeff(&self.0, a.as_mut_ptr() as *mut u64) }, "lagrange_to_coeff");
    }
    pub fn coeff_to_extended(&self, a: &[Fr]) -> Vec<Fr> {
        assert_eq!(a.len(), 1 << self.0.k);
        let mut out = vec![Fr::zero(); 1 << self.0.extended_k];
        check(unsafe { h2b_coeff_to_extended(&self.0, a.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64) }, "coeff_to_extended");
        out
    }
    /// `for a in cols { self.lagrange_to_coeff(a) }` with one kernel launch per pass for all columns.
    pub fn lagrange_to_coeff_many(&self, cols: &mut [&mut [Fr]]) {
        assert!(cols.iter().all(|a| a.len() == 1 << self.0.k));
        let ptrs: Vec<*mut u64> = cols.iter_mut().map(|a| a.as_mut_ptr() as *mut u64).collect();
        check(unsafe { h2b_lagrange_to_coeff_many(&self.0, ptrs.as_ptr(), ptrs.len()) }, "lagrange_to_coeff_many");
    }
    /// `cols.iter().map(|a| self.coeff_to_extended(a))` in one call (evaluate_h extends every column).
    pub fn coeff_to_extended_many(&self, cols: &[&[Fr]]) -> Vec<Vec<Fr>> {
        assert!(cols.iter().all(|a| a.len() == 1 << self.0.k));
        let mut outs: Vec<Vec<Fr>> = cols.iter().map(|_| vec![Fr::zero(); 1 << self.0.extended_k]).collect();
        let pin: Vec<*const u64> = cols.iter().map(|a| a.as_ptr() as *const u64).collect();
        let pout: Vec<*mut u64> = outs.iter_mut().map(|a| a.as_mut_ptr() as *mut u64).collect();
        check(unsafe { h2b_coeff_to_extended_many(&self.0, pin.as_ptr(), pout.as_ptr(), pin.len()) }, "coeff_to_extended_many");
        outs
    }
    pub fn extended_to_coeff(&self, a: &[Fr]) -> Vec<Fr> {
        assert_eq!(a.len(), 1 << self.0.extended_k);
        let mut out = vec![Fr::zero(); ((self.0.j - 1) as usize) << self.0.k];
        check(unsafe { h2b_extended_to_coeff(&self.0, a.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64) }, "extended_to_coeff");
        out
    }
}
