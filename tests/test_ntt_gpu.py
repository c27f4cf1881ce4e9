"""GPU parity for best_fft and the EvaluationDomain transforms (bit-exact against the oracle)."""
import numpy as np
import pytest

from util import load_golden, unhx

pytestmark = pytest.mark.gpu


def _omega(spec, k):
    return spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]


@pytest.mark.parametrize("k", list(range(0, 19)) + [20])
def test_best_fft_vs_oracle(h2b, spec, href, k):
    a = href.random_fr(1 << k, 1000 + k)
    om = _omega(spec, k)
    want = href.best_fft(a, om, k)
    got = a.copy()
    h2b.best_fft(got, om, k)
    assert (got == want).all()


def test_best_fft_small_vs_bigint_spec(h2b, spec):
    for k in range(0, 8):
        vals = spec.random_fr(1 << k, 77 + k)
        om = pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)
        a = spec.fr_array(vals)
        h2b.best_fft(a, spec.fr_array([om])[0], k)
        assert spec.fr_ints(a) == spec.dft_naive(vals, om)


def test_best_fft_arbitrary_omega_and_inverse(h2b, spec, href):
    # inverse root, then scaling by 1/n on the oracle side, returns the input
    k = 12
    a = href.random_fr(1 << k, 5)
    om = pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)
    f = a.copy()
    h2b.best_fft(f, spec.fr_array([om])[0], k)
    h2b.best_fft(f, spec.fr_array([pow(om, -1, spec.R_MOD)])[0], k)
    ninv = spec.fr_array([pow(1 << k, -1, spec.R_MOD)] * (1 << k))
    assert (href.fr_mul(f, ninv) == a).all()


def test_best_fft_edge_inputs(h2b, spec, href):
    k = 11
    om = _omega(spec, k)
    n = 1 << k
    zeros = np.zeros((n, 4), dtype=np.uint64)
    z = zeros.copy()
    h2b.best_fft(z, om, k)
    assert (z == 0).all()
    delta = zeros.copy()
    delta[0] = spec.fr_array([1])[0]
    h2b.best_fft(delta, om, k)
    assert (delta == spec.fr_array([1])[0]).all()  # transform of a delta is all ones
    top = spec.fr_array([spec.R_MOD - 1] * n)
    want = href.best_fft(top, om, k)
    h2b.best_fft(top, om, k)
    assert (top == want).all()


def test_bad_length_is_rejected(h2b, spec):
    with pytest.raises(AssertionError):
        h2b.best_fft(np.zeros((3, 4), dtype=np.uint64), _omega(spec, 2), 2)


@pytest.mark.parametrize("j,k", [(3, 1), (3, 4), (4, 5), (6, 6), (2, 3), (5, 4), (9, 3), (4, 10), (6, 11), (3, 13), (4, 14)])
def test_domain_vs_oracle(h2b, spec, href, j, k):
    d = h2b.EvaluationDomain(j, k)
    dc = href.domain_new(j, k)
    assert d.extended_k == dc.extended_k
    for mine, name in ((d.get_omega(), "omega"), (d.get_omega_inv(), "omega_inv"),
                       (d.get_extended_omega(), "extended_omega"), (d.extended_omega_inv, "extended_omega_inv"),
                       (d.g_coset, "g_coset"), (d.g_coset_inv, "g_coset_inv"), (d.ifft_divisor, "ifft_divisor"),
                       (d.extended_ifft_divisor, "extended_ifft_divisor")):
        assert (mine == np.array(list(getattr(dc, name)), dtype=np.uint64)).all(), name
    assert (d.t_evaluations.reshape(-1) == np.array(list(dc.t_evaluations), dtype=np.uint64)[: 4 * dc.n_t]).all()
    a = href.random_fr(1 << k, 31 * j + k)
    assert (d.lagrange_to_coeff(a.copy()) == href.lagrange_to_coeff(dc, a)).all()
    ext = d.coeff_to_extended(a)
    assert (ext == href.coeff_to_extended(dc, a)).all()
    e = href.random_fr(1 << d.extended_k, 17 * j + k)
    assert (d.extended_to_coeff(e) == href.extended_to_coeff(dc, e)).all()
    assert (d.divide_by_vanishing_poly(e.copy()) == href.divide_by_vanishing_poly(dc, e)).all()
    # round trip through the coset: coefficients come back, zero-extended / truncated to n*(j-1)
    back = d.extended_to_coeff(ext)
    keep = (1 << k) * (j - 1)
    ref = np.zeros((keep, 4), dtype=np.uint64)
    m = min(keep, 1 << k)
    ref[:m] = a[:m]
    assert (back == ref).all()


@pytest.mark.parametrize("j,k,m", [(3, 1, 3), (4, 5, 5), (6, 9, 7), (4, 12, 3), (6, 14, 4), (3, 21, 2)])
def test_batched_column_transforms_equal_single_calls(h2b, spec, href, j, k, m):
    """lagrange_to_coeff_many / coeff_to_extended_many (one launch per pass for all columns) == single calls
    == the oracle, for one-, two- and three-pass sizes."""
    d = h2b.EvaluationDomain(j, k)
    dc = href.domain_new(j, k)
    cols = [href.random_fr(1 << k, 500 + 7 * k + q) for q in range(m)]
    cols[-1][:] = 0
    got = d.lagrange_to_coeff_many([c.copy() for c in cols])
    ext = d.coeff_to_extended_many(cols)
    for q in range(m):
        if k <= 14 or q == 0:
            assert (got[q] == href.lagrange_to_coeff(dc, cols[q])).all(), q
            assert (ext[q] == href.coeff_to_extended(dc, cols[q])).all(), q
        else:
            assert (got[q] == d.lagrange_to_coeff(cols[q].copy())).all(), q
            assert (ext[q] == d.coeff_to_extended(cols[q])).all(), q
    assert d.lagrange_to_coeff_many([]) == [] and d.coeff_to_extended_many([]) == []
    with pytest.raises(AssertionError):
        d.coeff_to_extended_many([np.zeros((3, 4), dtype=np.uint64)])


def test_domain_length_asserts(h2b, href):
    d = h2b.EvaluationDomain(4, 5)
    with pytest.raises(AssertionError):
        d.lagrange_to_coeff(np.zeros((31, 4), dtype=np.uint64))
    with pytest.raises(AssertionError):
        d.coeff_to_extended(np.zeros((64, 4), dtype=np.uint64))
    with pytest.raises(AssertionError):
        d.extended_to_coeff(np.zeros((32, 4), dtype=np.uint64))


def test_golden_vectors(h2b):
    g = load_golden("spec_vectors.json")
    for v in g["ntt"]:
        a = unhx(v["a"], 4)
        h2b.best_fft(a, unhx(v["omega"], 4)[0], v["log_n"])
        assert (a == unhx(v["out"], 4)).all(), v["log_n"]
    for v in g["domain"]:
        d = h2b.EvaluationDomain(v["j"], v["k"])
        assert d.extended_k == v["extended_k"]
        assert (d.get_omega() == unhx(v["omega"], 4)[0]).all()
        assert (d.get_extended_omega() == unhx(v["extended_omega"], 4)[0]).all()
        assert (d.t_evaluations == unhx(v["t_evaluations"], 4)).all()
        a, e = unhx(v["a"], 4), unhx(v["e"], 4)
        assert (d.lagrange_to_coeff(a.copy()) == unhx(v["lagrange_to_coeff"], 4)).all()
        assert (d.coeff_to_extended(a) == unhx(v["coeff_to_extended"], 4)).all()
        assert (d.extended_to_coeff(e) == unhx(v["extended_to_coeff"], 4)).all()
        assert (d.divide_by_vanishing_poly(e.copy()) == unhx(v["divide_by_vanishing_poly"], 4)).all()


@pytest.mark.parametrize("k,ext", [(18, 20), (20, 22), (21, 24)])
def test_full_size_properties(h2b, spec, href, k, ext):
    """BASELINE sizes, checked through size-independent properties (the oracle would take minutes):
    coset round trip, and linearity of the transform on a random pair."""
    j = {2: 4, 3: 6}[ext - k]  # extended_k - k = 2 -> j-1 = 3..4 ; 3 -> j-1 = 5..8
    d = h2b.EvaluationDomain(j, k)
    assert d.extended_k == ext
    n = 1 << k
    a = href.random_fr(n, 900 + k)
    extd = d.coeff_to_extended(a)
    back = d.extended_to_coeff(extd)
    assert back.shape[0] == n * (j - 1)
    assert (back[:n] == a).all() and (back[n:] == 0).all()
    # X[0] = sum of the coset-scaled inputs: check one output against an independent reduction
    # linearity: T(a) + T(b) == T(a + b) on a sample of positions
    b = href.random_fr(n, 901 + k)
    s_ab = spec.fr_array([(x + y) % spec.R_MOD for x, y in zip(spec.fr_ints(a[:64]), spec.fr_ints(b[:64]))])
    ab = np.zeros_like(a)
    ab[:64] = s_ab
    a2, b2 = np.zeros_like(a), np.zeros_like(a)
    a2[:64], b2[:64] = a[:64], b[:64]
    ta, tb, tab = d.coeff_to_extended(a2), d.coeff_to_extended(b2), d.coeff_to_extended(ab)
    idx = np.linspace(0, (1 << ext) - 1, 97).astype(np.int64)
    sa, sb, sab = spec.fr_ints(ta[idx]), spec.fr_ints(tb[idx]), spec.fr_ints(tab[idx])
    assert [(x + y) % spec.R_MOD for x, y in zip(sa, sb)] == sab
