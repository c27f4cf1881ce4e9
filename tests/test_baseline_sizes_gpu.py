"""GPU parity at BASELINE.json's own sizes, compared DIRECTLY with the oracle (oracle/libh2ref.so, the C restatement
of halo2_proofs@6b43b6b best_multiexp / best_fft / EvaluationDomain, all host threads): MSM 2^20, 2^22, 2^24 and
2^16 +- 1 through both paths (ParamsKZG::commit against a registered SRS with its window table -- c = 20, 13
windows at 2^24, the benchmarked geometry -- and best_multiexp with the caller's bases), best_fft k = 22 and 24,
coeff_to_extended for SURVEY.md section 8d's (k, extended_k) pairs and extended_to_coeff with j in {3, 4, 6}.
MSM results are compared after affine normalisation, NTT results limb for limb."""
import ctypes as C

import numpy as np
import pytest

from util import rand_fr

pytestmark = pytest.mark.gpu


def _device_bases(n, seed):
    """n distinct points [s_i] G built on the GPU (h2b_dev_fixed_base_mul, itself checked against the oracle in
    test_fixed_base_mul_vs_oracle) and read back, so that the oracle and the GPU see the same host array."""
    import torch
    import bn254
    from halo2_prover_b200 import _ffi
    gen = bn254.affine_to_array([bn254.G1_GENERATOR])[0]
    seeds = torch.from_numpy(rand_fr(n, seed).view(np.int64)).cuda()
    out = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream()
    _ffi.check(_ffi.lib().h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(n), _ffi.u64p(gen),
                                                 C.c_void_p(out.data_ptr()), C.c_void_p(s.cuda_stream or 1)))
    s.synchronize()
    return np.ascontiguousarray(out.cpu().numpy().view(np.uint64))


def test_fixed_base_mul_vs_oracle(h2b, spec, href):
    """h2b_dev_fixed_base_mul (ParamsKZG::setup's per-element multiplication, kzg/commitment.rs:68-114) against the
    oracle's g1_scalar_mul: random scalars plus 0, 1, 2, r - 1, r - 2, 2^253, and the identity as base."""
    import torch
    from halo2_prover_b200 import _ffi
    edge = [0, 1, 2, spec.R_MOD - 1, spec.R_MOD - 2, 1 << 253, (1 << 128) - 1, 1 << 64]
    sc = np.concatenate([spec.fr_array(edge), href.random_fr(56, 404)])
    n = sc.shape[0]
    gen = spec.affine_to_array([spec.G1_GENERATOR])[0]
    other = href.random_g1(1, 405)[0]
    s = torch.cuda.current_stream()
    for base in (gen, other, np.zeros(8, dtype=np.uint64)):
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        d_out = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        _ffi.check(_ffi.lib().h2b_dev_fixed_base_mul(C.c_void_p(d_sc.data_ptr()), C.c_size_t(n), _ffi.u64p(np.ascontiguousarray(base)),
                                                     C.c_void_p(d_out.data_ptr()), C.c_void_p(s.cuda_stream or 1)))
        s.synchronize()
        got = d_out.cpu().numpy().view(np.uint64)
        for i in range(n):
            want = href.g1_to_affine(href.g1_scalar_mul(base, sc[i]))
            assert (got[i] == want).all(), (i, base[:1])


@pytest.mark.parametrize("n", [(1 << 16) - 1, (1 << 16) + 1, 1 << 20, 1 << 22, 1 << 24])
def test_msm_at_baseline_sizes_vs_oracle(h2b, href, n):
    from halo2_prover_b200 import _ffi
    bases = _device_bases(n, 9000 + (n & 0xffff))
    scalars = rand_fr(n, 9100 + (n & 0xffff))
    want = href.g1_to_affine(href.best_multiexp(scalars, bases))
    # best_multiexp with the caller's bases (no table, one bucket set per window)
    assert (href.g1_to_affine(h2b.best_multiexp(scalars, bases)) == want).all(), "best_multiexp"
    # ParamsKZG::commit against the registered SRS (window table; the benchmarked path)
    k = int(n - 1).bit_length()
    params = h2b.ParamsKZG.__new__(h2b.ParamsKZG)
    params.k, params.n, params._handles = k, n, {}
    h = C.c_uint64(0)
    _ffi.check(_ffi.lib().h2b_srs_register(_ffi.u64p(bases), C.c_size_t(n), C.byref(h)))
    params._handles["g"] = h.value
    try:
        c, w = C.c_uint32(), C.c_uint32()
        _ffi.check(_ffi.lib().h2b_srs_info(h, None, C.byref(c), C.byref(w), None))
        if n == 1 << 24:
            assert (c.value, w.value) == (20, 13), "the geometry bench.py measures"
        assert (href.g1_to_affine(params.commit(scalars)) == want).all(), "commit"
        # a shorter polynomial against the same SRS uses bases[0..size]
        m = n // 2 + 3
        want_m = href.g1_to_affine(href.best_multiexp(np.ascontiguousarray(scalars[:m]), np.ascontiguousarray(bases[:m])))
        assert (href.g1_to_affine(params.commit(np.ascontiguousarray(scalars[:m]))) == want_m).all(), "commit (short)"
    finally:
        params.release()


def test_msm_entry_limit_split_vs_oracle(h2b, href):
    """A pass holds at most 2^31 sorted entries; beyond that a chunk is split by point range.  Lower the limit so
    that a 2^14-point MSM is split several times, through both paths."""
    from halo2_prover_b200 import _ffi
    n = 1 << 14
    sc, pts = href.random_fr(n, 61), href.random_g1(n, 62)
    want = href.g1_to_affine(href.best_multiexp(sc, pts))
    _ffi.check(_ffi.lib().h2b_test_set_max_entries(15))
    try:
        assert (href.g1_to_affine(h2b.best_multiexp(sc, pts)) == want).all()
        _ffi.check(_ffi.lib().h2b_set_srs_precompute(2, 0))
        params = h2b.ParamsKZG(14, pts)
        assert (href.g1_to_affine(params.commit(sc)) == want).all()
        params.release()
    finally:
        _ffi.check(_ffi.lib().h2b_set_srs_precompute(1, 0))
        _ffi.check(_ffi.lib().h2b_test_set_max_entries(0))


@pytest.mark.parametrize("k", [22, 24])
def test_best_fft_large_vs_oracle(h2b, spec, href, k):
    a = rand_fr(1 << k, 7000 + k)
    omega = spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]
    want = href.best_fft(a, omega, k)
    got = a.copy()
    h2b.best_fft(got, omega, k)
    assert (got == want).all()


# (j, k, extended_k): SURVEY.md section 8d's coset pairs
@pytest.mark.parametrize("j,k,ek", [(4, 18, 20), (6, 17, 20), (3, 19, 20), (4, 20, 22), (6, 21, 24)])
def test_coeff_to_extended_pairs_vs_oracle(h2b, href, j, k, ek):
    d, dc = h2b.EvaluationDomain(j, k), href.domain_new(j, k)
    assert d.extended_k == ek == dc.extended_k
    a = rand_fr(1 << k, 7100 + k + j)
    assert (d.coeff_to_extended(a) == href.coeff_to_extended(dc, a)).all()


@pytest.mark.parametrize("j,k,ek", [(3, 19, 20), (4, 18, 20), (6, 17, 20), (4, 20, 22), (3, 23, 24), (6, 21, 24)])
def test_extended_to_coeff_vs_oracle(h2b, href, j, k, ek):
    d, dc = h2b.EvaluationDomain(j, k), href.domain_new(j, k)
    assert d.extended_k == ek
    a = rand_fr(1 << ek, 7200 + k + j)
    got = d.extended_to_coeff(a)
    assert got.shape[0] == (j - 1) << k  # truncation to n * (j - 1), not a power of two for j = 4, 6
    assert (got == href.extended_to_coeff(dc, a)).all()


@pytest.mark.parametrize("k", [20, 22])
def test_lagrange_to_coeff_large_vs_oracle(h2b, href, k):
    d, dc = h2b.EvaluationDomain(3, k), href.domain_new(3, k)
    a = rand_fr(1 << k, 7300 + k)
    assert (d.lagrange_to_coeff(a.copy()) == href.lagrange_to_coeff(dc, a)).all()
