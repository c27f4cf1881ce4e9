"""CPU tests of the oracle itself: the C restatement against the big-integer spec and both
against the committed golden vectors.  No GPU, no product code."""
import numpy as np
import pytest

from util import load_golden, unhx


def test_constants(spec):
    o = spec
    assert pow(o.FR_GENERATOR, (o.R_MOD - 1) >> o.FR_S, o.R_MOD) == o.ROOT_OF_UNITY
    assert pow(o.ROOT_OF_UNITY, 1 << o.FR_S, o.R_MOD) == 1
    assert pow(o.ROOT_OF_UNITY, 1 << (o.FR_S - 1), o.R_MOD) != 1
    assert pow(o.ZETA, 3, o.R_MOD) == 1 and o.ZETA != 1
    assert o.g1_is_on_curve(o.G1_GENERATOR)
    assert o.g1_mul(o.G1_GENERATOR, o.R_MOD) is None  # group order r, cofactor 1
    # Montgomery forms quoted in SURVEY.md section 8 (recovered from the reference binary)
    assert hex(o.to_mont(o.ROOT_OF_UNITY, o.R_MOD)).startswith("0x1d69070d")
    assert hex(o.to_mont(o.ROOT_OF_UNITY, o.R_MOD)).endswith("b639feb8")
    # the immediate 0x059c805d... found in the reference binary's EvaluationDomain::new is ZETA^2 (g_coset_inv)
    assert hex(o.to_mont(o.ZETA * o.ZETA % o.R_MOD, o.R_MOD)).startswith("0x59c805d")
    assert o.ZETA == 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD


def test_random_streams_agree(spec, href):
    assert (href.random_fr(50, 7) == spec.fr_array(spec.random_fr(50, 7))).all()
    assert (href.random_g1(24, 9) == spec.affine_to_array(spec.random_g1(24, 9))).all()


def test_field_mul(spec, href):
    a, b = href.random_fr(200, 1), href.random_fr(200, 2)
    ai, bi = spec.fr_ints(a), spec.fr_ints(b)
    assert (href.fr_mul(a, b) == spec.fr_array([x * y % spec.R_MOD for x, y in zip(ai, bi)])).all()
    a, b = href.random_g1(100, 3)[:, :4].copy(), href.random_g1(100, 4)[:, 4:].copy()
    ai, bi = spec.array_to_ints(a, spec.Q_MOD), spec.array_to_ints(b, spec.Q_MOD)
    want = spec.ints_to_array([x * y % spec.Q_MOD for x, y in zip(ai, bi)], spec.Q_MOD)
    assert (href.fq_mul(a, b) == want).all()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 100])
def test_multiexp_c_vs_spec(spec, href, n):
    sc, pts = href.random_fr(n, n + 100), href.random_g1(n, n + 200)
    want = spec.msm_naive(spec.fr_ints(sc), spec.array_to_affine(pts))
    for threads in (1, 3, 8):
        got = spec.projective_array_to_affine(href.best_multiexp(sc, pts, threads))
        assert got == want
    # the spec's own restatement of multiexp_serial / best_multiexp agrees too
    assert spec.best_multiexp(spec.fr_ints(sc), spec.array_to_affine(pts), 3) == want


@pytest.mark.parametrize("k", range(0, 11))
def test_fft_c_vs_spec(spec, href, k):
    omega = pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)
    a = href.random_fr(1 << k, k)
    want = spec.best_fft(spec.fr_ints(a), omega, k)
    if k <= 6:
        assert want == spec.dft_naive(spec.fr_ints(a), omega)
    for threads in (1, 2, 4, 8):
        got = href.best_fft(a, spec.fr_array([omega])[0], k, threads)
        assert spec.fr_ints(got) == want


@pytest.mark.parametrize("j,k", [(3, 4), (4, 5), (6, 6), (2, 3), (5, 4), (9, 3)])
def test_domain_c_vs_spec(spec, href, j, k):
    dp, dc = spec.EvaluationDomain(j, k), href.domain_new(j, k)
    assert dc.extended_k == dp.extended_k
    for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                 "ifft_divisor", "extended_ifft_divisor"):
        assert spec.from_mont(spec.limbs_to_int(list(getattr(dc, name))), spec.R_MOD) == getattr(dp, name), name
    te = spec.fr_ints(np.array(list(dc.t_evaluations), dtype=np.uint64)[: 4 * dc.n_t])
    assert te == dp.t_evaluations
    a = href.random_fr(1 << k, 5)
    ai = spec.fr_ints(a)
    for threads in (1, 4):
        assert spec.fr_ints(href.lagrange_to_coeff(dc, a, threads)) == dp.lagrange_to_coeff(ai)
        e = href.coeff_to_extended(dc, a, threads)
        assert spec.fr_ints(e) == dp.coeff_to_extended(ai)
        assert spec.fr_ints(href.extended_to_coeff(dc, e, threads)) == dp.extended_to_coeff(spec.fr_ints(e))
        assert spec.fr_ints(href.divide_by_vanishing_poly(dc, e, threads)) == dp.divide_by_vanishing_poly(spec.fr_ints(e))
    # round trip: extended_to_coeff(coeff_to_extended(p)) == p zero-extended / truncated to n*(j-1)
    back = dp.extended_to_coeff(dp.coeff_to_extended(ai))
    keep = (1 << k) * (j - 1)
    assert back == (ai + [0] * keep)[:keep]


def test_golden_vectors_c_oracle(spec, href):
    g = load_golden("spec_vectors.json")
    for v in g["ntt"]:
        got = href.best_fft(unhx(v["a"], 4), unhx(v["omega"], 4)[0], v["log_n"], 4)
        assert (got == unhx(v["out"], 4)).all()
    for v in g["domain"]:
        d = href.domain_new(v["j"], v["k"])
        assert d.extended_k == v["extended_k"]
        a, e = unhx(v["a"], 4), unhx(v["e"], 4)
        assert (href.lagrange_to_coeff(d, a) == unhx(v["lagrange_to_coeff"], 4)).all()
        assert (href.coeff_to_extended(d, a) == unhx(v["coeff_to_extended"], 4)).all()
        assert (href.extended_to_coeff(d, e) == unhx(v["extended_to_coeff"], 4)).all()
        assert (href.divide_by_vanishing_poly(d, e) == unhx(v["divide_by_vanishing_poly"], 4)).all()
    for v in g["msm"]:
        jac = href.best_multiexp(unhx(v["scalars"], 4), unhx(v["bases"], 8), 4)
        assert (href.g1_to_affine(jac) == unhx(v["affine"], 8)[0]).all()


def test_poseidon_constants_against_the_reference_known_answers():
    """oracle/poseidon_spec.py (the constants behind the third evaluate_h pin) against the known answers the reference
    holds for the same generator over the Pallas base field: all 192 round constants, MDS, MDS^-1
    (circuits/src/poseidon/primitives/fp.rs; the reference's own test p128pow5t3.rs:116-148 makes this comparison), and
    its zcash permutation / hash vectors (primitives/test_vectors.rs:16, :420).  Fixture cut by oracle/make_poseidon_kat.py."""
    import hashlib
    import json
    import os

    import poseidon_spec as ps
    from util import GOLDEN
    kat = json.load(open(os.path.join(GOLDEN, "poseidon_pallas_kat.json")))
    p = int(kat["modulus"], 16)
    rc, mds, inv = ps.generate_constants(3, r_p=56, modulus=p, num_bits=kat["num_bits"])
    hx = lambda rows: [[hex(v) for v in row] for row in rows]
    assert len(rc) == 64 and hx(rc[:2]) == kat["round_constants_first"] and hx([rc[-1]])[0] == kat["round_constants_last"]
    assert hashlib.sha256(",".join(hex(v) for row in rc for v in row).encode()).hexdigest() == kat["round_constants_sha256"]
    assert hx(mds) == kat["mds"] and hx(inv) == kat["mds_inv"]
    for v in kat["permute_vectors"]:
        assert [hex(x) for x in ps.permute([int(x, 16) for x in v["initial"]], rc, mds, p)] == v["final"]
    for v in kat["hash_vectors"]:
        assert hex(ps.hash_constant_length([int(x, 16) for x in v["input"]], 3, (rc, mds, inv), p)) == v["output"]
