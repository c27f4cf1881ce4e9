/*
 * h2b200.h -- C ABI of libh2b200.so: BN254 MSM + NTT proving hot path on B200 (sm_100a).
 *
 * Drop-in boundary for the leaf functions the reference reaches through
 * halo2_proofs::plonk::create_proof (/root/reference/circuits/src/utils.rs:83-91,
 * :105-120; keygen at :67-68).  The leaves live in the Cargo.lock-pinned,
 * un-vendored dependency halo2_proofs 0.2.0 @6b43b6b (circuits/Cargo.lock:836-838)
 * with field/curve types from halo2curves 0.3.2 @9f5c508 (:854-856); each entry
 * point below names the upstream item it replaces.  INTEGRATION.md shows the Rust
 * `-sys` binding and the patch that rewires the call sites.
 *
 * Data layout (identical to the Rust types, so buffers cross the FFI untouched):
 *   Fr, Fq      4 x uint64_t little-endian limbs, Montgomery form (R = 2^256), < modulus
 *   G1Affine    {x: Fq, y: Fq}       64 bytes, identity = (0, 0)
 *   G1          {x, y, z: Fq}        96 bytes, homogeneous projective (x = X/Z, y = Y/Z), identity z = 0
 *                                    (what halo2curves 0.3.2 stores; verified by running the reference binary)
 * All pointers are plain host pointers unless the function name contains `_dev`,
 * in which case they are CUDA device pointers on the library's device and `stream`
 * is a cudaStream_t passed as void* (NULL = the library's own stream).
 *
 * Errors: upstream panics (assert_eq!) on bad lengths; a C ABI cannot unwind, so
 * every function returns H2B_OK (0) or a negative code and never throws.  The
 * binding asserts on non-zero to keep the reference's behaviour.  There is no CPU
 * fallback: without a usable GPU every compute call returns H2B_ERR_CUDA.
 *
 * Threading: calls may come from any thread; the library serialises them on one
 * internal mutex (create_proof issues them sequentially from one thread anyway).
 * `_dev` pointers must be 16-byte aligned (field elements move as 128-bit transactions).
 */
#ifndef H2B200_H
#define H2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2B_OK 0
#define H2B_ERR_ARG (-1)    /* bad length / null pointer / unsupported size (upstream: assert panic) */
#define H2B_ERR_CUDA (-2)   /* CUDA runtime error, no device */
#define H2B_ERR_OOM (-3)    /* device allocation failed */
#define H2B_ERR_STATE (-4)  /* unknown handle / not initialised */

/* ---- lifecycle ------------------------------------------------------------------- */
/* Select the CUDA device and create the library context (stream, workspace pool).
 * Idempotent for the same device. */
int h2b_init(int device);
/* Several devices behind ONE process and one handle space (the reference's prover is a single process calling
 * create_proof, /root/reference/circuits/src/utils.rs:105-120): devices[0] is the primary device -- every `_dev`
 * pointer and stream belongs to it -- and the others take shares of the work:
 *   - a registered SRS of at least 2^21 points is sharded by contiguous point range, each device precomputes the
 *     window table of its own share, every commit is split the same way (each device copies its slice of the
 *     host scalars over its own PCIe link), and the partial sums (96 B each) travel device-to-device and are
 *     folded on the primary device: best_multiexp's own chunk + fold (arithmetic.rs:152-176) with a device in
 *     place of a thread;
 *   - a smaller SRS is replicated, and the independent columns of h2b_commit_many / h2b_lagrange_to_coeff_many /
 *     h2b_coeff_to_extended_many are dealt to the devices in contiguous blocks (a single NTT stays on one device).
 * Idempotent for the same list; h2b_init(d) is h2b_init_devices(&d, 1). */
int h2b_init_devices(const int *devices, int count);
int h2b_device_count(void);
void h2b_shutdown(void);
/* Human-readable description of the last error on the calling thread's last call. */
const char *h2b_last_error(void);
/* ABI version of this header. */
uint32_t h2b_abi_version(void);   /* 2: h2b_eval_h.flags, h2b_init_devices */

/* ---- MSM ------------------------------------------------------------------------- */
/* best_multiexp(coeffs: &[Fr], bases: &[G1Affine]) -> G1
 * replaces halo2_proofs src/arithmetic.rs:147-180 (and multiexp_serial :28-140).
 * coeffs: n x 4 u64, bases: n x 8 u64, out: 12 u64 (projective; any representative of the sum). */
int h2b_best_multiexp(const uint64_t *coeffs, const uint64_t *bases, size_t n, uint64_t out[12]);

/* SRS residency.  ParamsKZG keeps two static base arrays (g, g_lagrange); register each
 * once (copied to device memory; the host copy is not referenced after return).
 * replaces the `&self.g[0..n]` / `&self.g_lagrange[0..n]` arguments of
 * src/poly/kzg/commitment.rs:319 (commit_lagrange) and :363 (commit). */
int h2b_srs_register(const uint64_t *bases, size_t n, uint64_t *handle);
int h2b_srs_release(uint64_t handle);
/* ParamsKZG::commit / commit_lagrange: best_multiexp(scalars, &bases[0..n]) against a
 * registered SRS (n <= registered length; upstream asserts bases.len() >= size). */
int h2b_commit(uint64_t srs, const uint64_t *scalars, size_t n, uint64_t out[12]);

/* Sum of `count` projective points (fold of per-GPU partial MSM results;
 * arithmetic.rs:~176 `results.iter().fold(identity, |a, b| a + b)`). */
int h2b_g1_fold(const uint64_t *points /* count x 12 */, size_t count, uint64_t out[12]);

/* ---- NTT ------------------------------------------------------------------------- */
/* best_fft(a: &mut [Fr], omega: Fr, log_n: u32)   replaces src/arithmetic.rs:185-290.
 * In place, natural order in and out, a has 2^log_n elements, log_n <= 28. */
int h2b_best_fft(uint64_t *a, const uint64_t omega[4], uint32_t log_n);

/* EvaluationDomain<Fr> constants, replaces EvaluationDomain::new(j, k)
 * (src/poly/domain.rs:~40-140; the assert at :101 is checked). */
typedef struct h2b_domain {
    uint32_t k;
    uint32_t extended_k;
    uint32_t j;                 /* cs degree; quotient_poly_degree = j - 1 */
    uint32_t n_t;               /* 2^(extended_k - k) = t_evaluations.len() */
    uint64_t omega[4];
    uint64_t omega_inv[4];
    uint64_t extended_omega[4];
    uint64_t extended_omega_inv[4];
    uint64_t g_coset[4];        /* ZETA */
    uint64_t g_coset_inv[4];    /* ZETA^2 */
    uint64_t ifft_divisor[4];          /* 1 / 2^k */
    uint64_t extended_ifft_divisor[4]; /* 1 / 2^extended_k */
    uint64_t t_evaluations[32 * 4];    /* 1 / ((zeta * extended_omega^i)^n - 1), i < n_t <= 32 */
    /* derived, filled by h2b_domain_new: extended_ifft_divisor * {1, zeta^2, zeta}, the fused
     * store-side scaling of extended_to_coeff (ifft divisor, then distribute_powers_zeta(false)) */
    uint64_t extended_ifft_coset[3 * 4];
} h2b_domain;
int h2b_domain_new(uint32_t j, uint32_t k, h2b_domain *out);

/* EvaluationDomain::lagrange_to_coeff (domain.rs:227): a has 2^k elements, in place. */
int h2b_lagrange_to_coeff(const h2b_domain *d, uint64_t *a);
/* EvaluationDomain::coeff_to_extended (domain.rs:244): in 2^k elements -> out 2^extended_k. */
int h2b_coeff_to_extended(const h2b_domain *d, const uint64_t *in, uint64_t *out);
/* EvaluationDomain::extended_to_coeff (domain.rs:311): in 2^extended_k -> out 2^k * (j-1). */
int h2b_extended_to_coeff(const h2b_domain *d, const uint64_t *in, uint64_t *out);
/* EvaluationDomain::divide_by_vanishing_poly: a[i] *= t_evaluations[i % n_t], 2^extended_k, in place. */
int h2b_divide_by_vanishing_poly(const h2b_domain *d, uint64_t *a);

/* g_to_lagrange (halo2_proofs arithmetic.rs; the G = G1 instantiation of best_fft with omega_inv, the
 * 1/2^k scaling and the normalisation): g_lagrange from g, as ParamsKZG::{setup, from_parts} need it
 * (kzg/commitment.rs:68-114).  g_bases and out: 2^k x 8 u64 affine points. */
int h2b_g_to_lagrange(const uint64_t *g_bases, uint32_t k, uint64_t *out);
/* G1Affine::to_bytes (halo2curves 0.3.2) of m projective results: normalise, x as 32 little-endian canonical
 * bytes with bit 6 of byte 31 = y mod 2 (pinned on the reference's proof bytes), identity = zeros (not
 * exercised by those proofs) -- the bytes create_proof's transcript writes for a commitment.  points: m x 12 u64, out: m x 32 bytes. */
int h2b_g1_to_bytes(const uint64_t *points, size_t m, uint8_t *out);

/* m polynomials of n scalars each committed against bases[0..n] of ONE registered SRS in a single pass
 * (the A advice columns, the permutation products or the h pieces of create_proof, which upstream
 * commits one by one: plonk/prover.rs advice `.map(|poly| params.commit_lagrange(poly, blind))`).
 * polys[q] points at column q (n x 4 u64, host); out receives m x 12 u64.  Results are identical
 * to m calls of h2b_commit; the fixed latency of a small MSM is paid once per batch. */
int h2b_commit_many(uint64_t srs, const uint64_t *const *polys, size_t n, size_t m, uint64_t *out);

/* lagrange_to_coeff / coeff_to_extended over m columns in one call (one kernel launch per pass for
 * the whole batch).  cols[q] / in[q] / out[q] are host columns of 2^k (resp. 2^extended_k) elements;
 * results are identical to m single calls. */
int h2b_lagrange_to_coeff_many(const h2b_domain *d, uint64_t *const *cols, size_t m);
int h2b_coeff_to_extended_many(const h2b_domain *d, const uint64_t *const *in, uint64_t *const *out, size_t m);

/* ---- device-resident variants (inputs/outputs already in HBM) ---------------------- */
/* h2b_srs_register for bases that are already in HBM (n x 64 B on the library's device); the library keeps
 * its own copy, the caller's buffer may be freed afterwards. */
int h2b_dev_srs_register(const void *d_bases, size_t n, uint64_t *handle);
int h2b_dev_msm(const void *d_coeffs, const void *d_bases, size_t n, void *d_out /* 96 B */, void *stream);
/* ParamsKZG::commit with the polynomial already in HBM: d_coeffs (n x 32 B) against bases[0..n] of a
 * registered SRS (commitment.rs:319, :363); uses the SRS's precomputed window table when it has one. */
int h2b_dev_commit(uint64_t srs, const void *d_coeffs, size_t n, void *d_out /* 96 B */, void *stream);
/* h2b_commit_many with the m columns already in HBM, one after the other (m x n x 32 B); d_out: m x 96 B. */
int h2b_dev_commit_many(uint64_t srs, const void *d_coeffs, size_t n, size_t m, void *d_out, void *stream);
/* Device address of a registered SRS (n x 64 bytes) on the primary device; for an SRS sharded over several devices
 * (h2b_init_devices) the primary device's share and its length. */
int h2b_srs_device_ptr(uint64_t srs, void **d_bases, size_t *n);
int h2b_dev_best_fft(void *d_a, const uint64_t omega[4], uint32_t log_n, void *stream);
int h2b_dev_lagrange_to_coeff(const h2b_domain *d, void *d_a, void *stream);
int h2b_dev_coeff_to_extended(const h2b_domain *d, const void *d_in, void *d_out, void *stream);
int h2b_dev_extended_to_coeff(const h2b_domain *d, const void *d_in, void *d_out, void *stream);
int h2b_dev_divide_by_vanishing_poly(const h2b_domain *d, void *d_a /* 2^extended_k, in place */, void *stream);
/* m columns one after the other in HBM: d_a + q * 2^k (in place); d_in + q * 2^k -> d_out + q * 2^extended_k. */
int h2b_dev_lagrange_to_coeff_many(const h2b_domain *d, void *d_a, size_t m, void *stream);
int h2b_dev_coeff_to_extended_many(const h2b_domain *d, const void *d_in, void *d_out, size_t m, void *stream);
int h2b_dev_g1_fold(const void *d_points, size_t count, void *d_out, void *stream);
/* out[i] = [scalars[i]] * base as G1Affine (n x 64 B): the per-element fixed-base multiplication of
 * ParamsKZG::setup (src/poly/kzg/commitment.rs:68-114); builds synthetic SRS / benchmark bases in HBM. */
int h2b_dev_fixed_base_mul(const void *d_scalars, size_t n, const uint64_t base[8], void *d_out, void *stream);

/* ---- quotient numerator ---------------------------------------- */
/* Evaluator::evaluate_h (halo2_proofs@6b43b6b src/plonk/evaluation.rs) for circuits without lookups, with
 * every column already on the extended coset in HBM (the outputs of h2b_dev_coeff_to_extended(_many)): the
 * custom gates as upstream's compiled GraphEvaluator, then the permutation argument folded in with y.
 *
 * ValueSource (one u64): kind | a << 8 | b << 36 with kind 0 Constant(a) 1 Intermediate(a) 2 Fixed(column a,
 * rotation index b) 3 Advice(a, b) 4 Instance(a, b) 5 Challenge(a) 6 Beta 7 Gamma 8 Theta 9 Y 10 PreviousValue.
 * Calculation: header u64 = op | target << 8 | nparts << 40 followed by its ValueSources, op 0 Add(a, b)
 * 1 Sub(a, b) 2 Mul(a, b) 3 Square(a) 4 Double(a) 5 Negate(a) 6 Horner(start, factor, parts[nparts])
 * 7 Store(a); `target` is the intermediate it writes.  The value of a row is that of the last calculation. */
typedef struct h2b_eval_h {
    uint32_t num_fixed, num_advice, num_instance, num_challenges;
    const void *const *fixed;    /* host arrays of DEVICE pointers, each column 2^extended_k x 32 B */
    const void *const *advice;
    const void *const *instance;
    const uint64_t *challenges;  /* host, num_challenges x 4 */
    uint64_t beta[4], gamma[4], theta[4], y[4];
    uint32_t num_constants, num_rotations, num_calcs, num_intermediates;
    const uint64_t *constants;   /* host, num_constants x 4 */
    const int32_t *rotations;    /* host, num_rotations (<= 32) */
    const uint64_t *calcs;       /* host, calc_words u64 */
    size_t calc_words;
    /* permutation argument; num_perm_columns = 0 when the circuit has none */
    uint32_t num_perm_columns, chunk_len; /* chunk_len = cs.degree() - 2 */
    int32_t last_rotation;                /* -(cs.blinding_factors() + 1) */
    const uint8_t *perm_kind;             /* per column of cs.permutation: 0 advice, 1 fixed, 2 instance */
    const uint32_t *perm_index;
    const void *const *sigma_cosets;      /* DEVICE pointers: pk.permutation.cosets */
    const void *const *z_cosets;          /* DEVICE pointers: permutation_product_coset of every chunk */
    const void *l0, *l_last, *l_active_row; /* DEVICE pointers */
    uint32_t flags;                         /* H2B_EVALH_* */
} h2b_eval_h;
/* flags: start from the values already in d_values (PreviousValue of the graph = the running value upstream threads
 * through the circuits of one create_proof) instead of from zero. */
#define H2B_EVALH_ACCUMULATE 1u
/* d_values: 2^extended_k x 32 B in HBM, overwritten (upstream starts from domain.empty_extended()) unless
 * H2B_EVALH_ACCUMULATE is set.  Gates and permutation argument run as ONE pass over the rows; a graph's intermediates
 * are renumbered by liveness on the host and live in a per-thread array (an HBM scratch of live-slots x 2^extended_k
 * elements is used only when more than 40 are live at once). */
int h2b_dev_evaluate_h(const h2b_domain *d, const h2b_eval_h *a, void *d_values, void *stream);
/* One lookup argument folded into d_values after h2b_dev_evaluate_h (evaluation.rs, "Lookup constraints"; call once
 * per lookup, in cs.lookups order).  `a` carries the column tables, the scalars, l0 / l_last / l_active_row and, as
 * its graph, THAT lookup's GraphEvaluator ((compressed input + beta) * (compressed table + gamma)); its permutation
 * fields are ignored.  d_product / d_permuted_input / d_permuted_table: that lookup's extended cosets.
 * Parity of this entry point is not pinned on a reference record: none of the reference's circuits has a lookup. */
int h2b_dev_evaluate_h_lookup(const h2b_domain *d, const h2b_eval_h *a, const void *d_product, const void *d_permuted_input,
                              const void *d_permuted_table, void *d_values, void *stream);

/* ---- tuning / introspection ------------------------------------------------------- */
/* Override the MSM window (0 = automatic). */
int h2b_set_msm_window(uint32_t c);
/* ParamsKZG::read for SerdeFormat::RawBytes (the bytes ParamsKZG::write produces, wasm.rs:52/:79/:126):
 * k:u32 LE | g[2^k] x 64 B | g_lagrange[2^k] x 64 B | g2 128 B | s_g2 128 B.  Registers both base arrays
 * straight from the buffer and returns their handles (release each with h2b_srs_release). */
int h2b_params_read(const uint8_t *bytes, size_t len, uint32_t *k, uint64_t *g_handle, uint64_t *g_lagrange_handle);
/* ParamsKZG::write (same format, wasm.rs:52): the two registered base arrays (each 2^k points, e.g. built on the device
 * by h2b_dev_fixed_base_mul + h2b_g_to_lagrange and registered with h2b_dev_srs_register) are read back from HBM into
 * `out` (len = 4 + 128 * 2^k + 256) between the k header and the caller's 256 bytes of g2 | s_g2 (G2 arithmetic is not
 * on this path).  h2b_params_write(h2b_params_read(bytes)) reproduces the reference's own setup() bytes. */
int h2b_params_write(uint32_t k, uint64_t g_handle, uint64_t g_lagrange_handle, const uint8_t *g2_and_s_g2, uint8_t *out, size_t len);
/* Geometry of a registered SRS: its length and, when it has a precomputed window table, the window
 * width, the number of windows (= bucket additions per point of a commit) and the table's size in HBM
 * (zeros when there is no table). */
int h2b_srs_info(uint64_t srs, size_t *n, uint32_t *window_bits, uint32_t *windows, size_t *table_bytes);
/* How a registered SRS is laid out over the library's devices: the number of shares, whether every share is the
 * whole array (replicated) or a contiguous point range, and (part_n, `*parts` entries, may be NULL) their lengths. */
int h2b_srs_layout(uint64_t srs, uint32_t *parts, uint32_t *replicated, size_t *part_n);
/* What h2b_srs_register precomputes for the static bases.  enabled = 1 (default): SRS of up to 2^14 points get
 * every window multiple d * 2^(8w) * P_i (commits become bucket-free sums, msm_comb.cuh), larger ones the window
 * table 2^(c*w) * P_i (all windows of a commit share one bucket set); 2: the window table at every size;
 * 0: nothing.  `c` overrides the window table's width (0 = automatic) and implies the window table.
 * Applies to SRS registered after the call. */
int h2b_set_srs_precompute(int enabled, uint32_t c);
/* The pinned-host-to-device bandwidth (GB/s) this process can count on for its scalar copies.  The copy pieces grow by
 * (accumulation time per point) / (copy time per point): ~4 with a PCIe Gen5 x16 link to itself (the default, 55), less
 * when several devices or processes share the host's bandwidth (8 x B200 copying at once: 24 GB/s each).
 * h2b_init_devices measures it with all its devices copying at once; a process that shares the host with other
 * processes (one rank per GPU) can pass what it measured.  0 restores the default. */
int h2b_set_h2d_bandwidth(double gbs);
/* How much of the window table h2b_srs_register keeps: every t-th window power 2^(c*t*v) * P_i (t bucket sets per
 * commit and a Horner over their t sums at the end).  t = 1 is the whole table (one bucket set, no doubling at all);
 * t = 2 halves its HBM for one extra bucket reduction and c doublings per commit (measured at 2^24: 7.5 instead of
 * 14 GB for +0.25 % device-resident, +2.4 % end to end from host memory, where every copy piece sorts into t times the
 * buckets).  0 (default): the whole table while it fits a sixth of the free HBM, else the thinnest stride up to 4 that
 * does (2^26 points: t = 2, 27.5 instead of 51.5 GB).  Applies to SRS registered after the call; results are identical
 * for every t. */
int h2b_set_srs_table_stride(uint32_t t);
/* Host-buffer MSM entry points (h2b_best_multiexp, h2b_commit) split inputs of at least `min_n`
 * points into `chunks` contiguous pieces so the H2D copy of a piece overlaps the bucket accumulation
 * of the previous one (default from 2^21 points: 2 pieces below 2^23 points, 4 from there; a call fixes the count,
 * chunks = 0 returns to that rule). */
int h2b_set_e2e_chunking(uint32_t chunks, size_t min_n);
/* Number of kernel launches issued by the library since h2b_init (for bench accounting). */
uint64_t h2b_kernel_launches(void);
/* Dominant-kernel timing (MSM: the bucket-accumulation kernel; NTT: all passes of one transform).
 * While enabled every call brackets that kernel with a CUDA event pair on the launching stream
 * (no synchronisation is added).  h2b_kernel_time_collect waits for the recorded pairs, returns
 * the summed milliseconds and the number of calls, and clears them (at most 256 are kept). */
int h2b_set_kernel_timing(int enabled);
int h2b_kernel_time_collect(double *total_ms, uint32_t *calls);

/* ---- test hooks (element-wise device arithmetic, used by tests/ only) ---------------- */
/* op: 0 mul, 1 add, 2 sub, 3 mul (portable 64-bit path), 4 inverse of a, 5 from_mont(a), 6 square of a,
 * 7 a*b + (a+b)*(a-b) and 8 a*b - b*a through the fused two-product multiply, 9 lazy product (a < 4N as raw limbs,
 * b < N), 10 lazy difference and 11 lazy sum (a, b < 2N as raw limbs) of the NTT butterflies, each brought back to [0, N)
 * field: 0 Fr, 1 Fq.  a, b, out: n x 4 u64 host buffers. */
int h2b_test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* out[i] = a[i] + b[i] on affine inputs (n x 8 u64) through the XYZZ mixed-add path -> n x 12 u64. */
int h2b_test_g1_add_affine(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* Lowers the number of sorted entries one MSM pass may hold (2^log2; the product limit is 2^31, beyond which a
 * chunk is split by point range) so that the splitting can be exercised at test sizes.  0 restores the limit. */
int h2b_test_set_max_entries(uint32_t log2_entries);
/* Integer-pipe microbenchmark: returns measured 32-bit IMAD (mad.lo.u32) and IMAD.WIDE
 * (mad.wide.u32) throughput in G instr/s on the current device. */
int h2b_imad_peak(double *imad_gops, double *imad_wide_gops, double *sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* H2B200_H */
