"""GPU parity for best_multiexp / ParamsKZG::commit (compared after affine normalisation,
SURVEY.md section 8b: the Jacobian representative is free)."""
import numpy as np
import pytest

from util import load_golden, unhx

pytestmark = pytest.mark.gpu


def _affine(href, jac):
    return href.g1_to_affine(np.ascontiguousarray(jac))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 255, 256, 257, 1000, 4096, (1 << 14) - 1, 1 << 16])
def test_msm_vs_oracle(h2b, href, n):
    sc, pts = href.random_fr(n, 5000 + n), href.random_g1(n, 6000 + n)
    want = _affine(href, href.best_multiexp(sc, pts))
    got = _affine(href, h2b.best_multiexp(sc, pts))
    assert (got == want).all()


def test_msm_empty_is_identity(h2b, spec):
    out = h2b.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64))
    assert (out[:4] == 0).all() and (out[8:] == 0).all()
    assert (out[4:8] == spec.ints_to_array([1], spec.Q_MOD)[0]).all()  # (0, R, 0)


def test_msm_length_mismatch_asserts(h2b):
    with pytest.raises(AssertionError):
        h2b.best_multiexp(np.zeros((2, 4), dtype=np.uint64), np.zeros((3, 8), dtype=np.uint64))


def test_msm_edge_scalars(h2b, spec, href):
    n = 777
    pts = href.random_g1(n, 42)
    for name, vals in (("zeros", [0] * n), ("ones", [1] * n), ("r-1", [spec.R_MOD - 1] * n),
                       ("mixed", [0, 1, spec.R_MOD - 1, 2, (1 << 253), (1 << 128) - 1, 1 << 16, (1 << 16) - 1] * 97 + [5])):
        sc = spec.fr_array(vals[:n])
        want = _affine(href, href.best_multiexp(sc, pts))
        got = _affine(href, h2b.best_multiexp(sc, pts))
        assert (got == want).all(), name
    # unit scalars = plain point sum
    acc = None
    for p in spec.array_to_affine(pts[:50]):
        acc = spec.g1_add(acc, p)
    got = spec.projective_array_to_affine(h2b.best_multiexp(spec.fr_array([1] * 50), pts[:50].copy()))
    assert got == acc


def test_msm_degenerate_bases(h2b, spec, href):
    n = 600
    sc = href.random_fr(n, 9)
    pts = href.random_g1(n, 10)
    pts[::7] = 0                      # identity bases (0,0)
    pts[1::7] = pts[1]                # one point repeated many times (same bucket doubling chains)
    neg = spec.affine_to_array([spec.g1_neg(p) for p in spec.array_to_affine(pts[2:3])])
    pts[2::7] = neg[0]                # -P next to P with equal scalars below
    sc[2::7] = sc[1]
    sc[1::7] = sc[1]
    want = _affine(href, href.best_multiexp(sc, pts))
    got = _affine(href, h2b.best_multiexp(sc, pts))
    assert (got == want).all()
    # everything cancels: sum s*P + s*(-P) = O
    half = href.random_g1(64, 3)
    negs = spec.affine_to_array([spec.g1_neg(p) for p in spec.array_to_affine(half)])
    s = href.random_fr(64, 4)
    out = h2b.best_multiexp(np.concatenate([s, s]), np.concatenate([half, negs]))
    assert spec.projective_array_to_affine(out) is None


def test_msm_witness_like_sparse(h2b, spec, href):
    """~5% non-zero small values plus a few random tail rows, like an advice column."""
    n = 1 << 14
    rng = np.random.default_rng(7)
    vals = [0] * n
    for i in rng.choice(n - 8, size=n // 20, replace=False):
        vals[int(i)] = int(rng.integers(1, 1 << 16))
    tail = spec.random_fr(6, 99)
    for t, v in enumerate(tail):
        vals[n - 6 + t] = v
    sc = spec.fr_array(vals)
    pts = href.random_g1(n, 12)
    want = _affine(href, href.best_multiexp(sc, pts))
    assert (_affine(href, h2b.best_multiexp(sc, pts)) == want).all()
    # heavy skew: every scalar equal (one bucket per window takes all points)
    sc = spec.fr_array([0xDEADBEEFCAFE] * n)
    want = _affine(href, href.best_multiexp(sc, pts))
    assert (_affine(href, h2b.best_multiexp(sc, pts)) == want).all()


@pytest.mark.parametrize("c", [4, 7, 10, 13, 16, 18])
def test_msm_every_window_size(h2b, href, c):
    from halo2_prover_b200 import _ffi
    n = 3000
    sc, pts = href.random_fr(n, 77), href.random_g1(n, 78)
    want = _affine(href, href.best_multiexp(sc, pts))
    _ffi.check(_ffi.lib().h2b_set_msm_window(c))
    try:
        got = _affine(href, h2b.best_multiexp(sc, pts))
    finally:
        _ffi.check(_ffi.lib().h2b_set_msm_window(0))
    assert (got == want).all()


def test_commit_against_registered_srs(h2b, spec, href):
    k = 10
    n = 1 << k
    g, gl = href.random_g1(n, 1), href.random_g1(n, 2)
    params = h2b.ParamsKZG(k, g, gl)
    poly = href.random_fr(n, 3)
    assert (_affine(href, params.commit(poly)) == _affine(href, href.best_multiexp(poly, g))).all()
    assert (_affine(href, params.commit_lagrange(poly)) == _affine(href, href.best_multiexp(poly, gl))).all()
    short = poly[:100].copy()  # commit of a shorter polynomial uses bases[0..size]
    assert (_affine(href, params.commit(short)) == _affine(href, href.best_multiexp(short, g[:100].copy()))).all()
    with pytest.raises(AssertionError):
        params.commit(href.random_fr(n + 1, 4))
    params.release()


@pytest.mark.parametrize("mode,srs_c", [(1, 0), (2, 0), (1, 7), (1, 13), (1, 16), (1, 20), (1, 22)])
def test_commit_precomputed_window_table(h2b, spec, href, mode, srs_c):
    """Registered bases get a precomputed table: mode 1 / c = 0 the bucket-free table of all window multiples
    (small SRS), mode 2 or an explicit c the window table 2^(c*w) * P_i whose windows share one bucket set.
    The result must equal best_multiexp on the plain bases in every mode, with identity bases, repeated
    scalars, r-1 / 0 / 1 scalars, for a prefix of the SRS, and with the tables disabled."""
    import ctypes as C
    from halo2_prover_b200 import _ffi
    n = 1 << 12
    g = href.random_g1(n, 61)
    g[5::97] = 0  # identity bases inside the SRS
    poly = href.random_fr(n, 62)
    poly[::7] = poly[3]
    poly[1] = spec.fr_array([spec.R_MOD - 1])[0]
    poly[2] = 0
    poly[4] = spec.fr_array([1])[0]
    want = _affine(href, href.best_multiexp(poly, g))
    want_short = _affine(href, href.best_multiexp(poly[:1500].copy(), g[:1500].copy()))
    try:
        _ffi.check(_ffi.lib().h2b_set_srs_precompute(mode, srs_c))
        params = h2b.ParamsKZG(12, g)
        assert (_affine(href, params.commit(poly)) == want).all()
        assert (_affine(href, params.commit(poly[:1500].copy())) == want_short).all()
        many = params.commit_many([poly, poly[::-1].copy(), np.zeros_like(poly)])
        assert (_affine(href, many[0]) == want).all()
        assert (_affine(href, many[1]) == _affine(href, href.best_multiexp(poly[::-1].copy(), g))).all()
        assert (_affine(href, many[2]) == 0).all()
        params.release()
        _ffi.check(_ffi.lib().h2b_set_srs_precompute(0, 0))
        params = h2b.ParamsKZG(12, g)
        assert (_affine(href, params.commit(poly)) == want).all()
        params.release()
    finally:
        _ffi.check(_ffi.lib().h2b_set_srs_precompute(1, 0))


def test_commit_table_collisions_and_odd_srs_length(h2b, spec, href):
    """Table path with an SRS whose length is not a power of two, duplicated bases (P + P inside a bucket),
    a base and its negative under equal scalars (P - P), and scalars that differ only in sign."""
    import ctypes as C
    from halo2_prover_b200 import _ffi
    n = 777
    g = href.random_g1(n, 401)
    g[11] = g[10]
    neg = spec.array_to_affine(g[13:14])[0]
    g[12] = spec.affine_to_array([(neg[0], spec.Q_MOD - neg[1])])[0]
    sc = href.random_fr(n, 402)
    sc[11] = sc[10]
    sc[12] = sc[13]
    r_minus = lambda a: spec.fr_array([(spec.R_MOD - v) % spec.R_MOD for v in spec.fr_ints(a.reshape(1, 4))])[0]
    sc[20] = r_minus(sc[21])
    g[20] = g[21]                      # s * P + (-s) * P
    h = C.c_uint64(0)
    _ffi.check(_ffi.lib().h2b_srs_register(_ffi.u64p(g), C.c_size_t(n), C.byref(h)))
    try:
        for m in (n, 500, 1):
            out = np.zeros(12, dtype=np.uint64)
            _ffi.check(_ffi.lib().h2b_commit(h, _ffi.u64p(np.ascontiguousarray(sc[:m])), C.c_size_t(m), _ffi.u64p(out)))
            assert (_affine(href, out) == _affine(href, href.best_multiexp(sc[:m].copy(), g[:m].copy()))).all(), m
        out = np.zeros(12, dtype=np.uint64)
        rc = _ffi.lib().h2b_commit(h, _ffi.u64p(np.ascontiguousarray(href.random_fr(n + 1, 5))), C.c_size_t(n + 1), _ffi.u64p(out))
        assert rc != 0  # bases.len() < size
    finally:
        _ffi.check(_ffi.lib().h2b_srs_release(h))


def test_dev_commit_matches_best_multiexp_2p18(h2b, spec, href):
    """Device-resident commit (h2b_dev_commit) over a 2^18-point SRS == the oracle's best_multiexp."""
    import ctypes as C
    import torch
    from halo2_prover_b200 import _ffi
    n = 1 << 18
    g = np.tile(href.random_g1(1 << 14, 71), (n >> 14, 1))
    poly = href.random_fr(n, 72)
    want = _affine(href, href.best_multiexp(poly, g))
    params = h2b.ParamsKZG(18, g)
    d_poly = torch.from_numpy(poly.view(np.int64)).cuda()
    out = torch.empty(12, dtype=torch.int64, device="cuda")
    params.dev_commit(d_poly, out)          # torch's default stream
    torch.cuda.current_stream().synchronize()
    assert (_affine(href, out.cpu().numpy().view(np.uint64)) == want).all()
    out.zero_()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    params.dev_commit(d_poly[: n // 2], out, stream=s)   # a prefix of the SRS, caller's stream
    s.synchronize()
    want = _affine(href, href.best_multiexp(poly[: n // 2].copy(), g[: n // 2].copy()))
    assert (_affine(href, out.cpu().numpy().view(np.uint64)) == want).all()
    want = _affine(href, href.best_multiexp(poly, g))
    assert (_affine(href, params.commit(poly)) == want).all()
    params.release()


@pytest.mark.parametrize("k,m", [(4, 3), (10, 9), (13, 5)])
def test_commit_many_equals_single_commits(h2b, spec, href, k, m):
    """h2b_commit_many (every column its own bucket set inside one pass) == m x h2b_commit == the oracle,
    including an all-zero column, a column of one repeated scalar and identity bases."""
    n = 1 << k
    g = href.random_g1(n, 80 + k)
    if n > 8:
        g[3::17] = 0
    cols = [href.random_fr(n, 90 + q) for q in range(m)]
    cols[1][:] = 0
    cols[2][:] = cols[2][0]
    params = h2b.ParamsKZG(k, g, g)
    got = params.commit_many(cols)
    for q in range(m):
        want = _affine(href, href.best_multiexp(cols[q], g))
        assert (_affine(href, got[q]) == want).all(), q
        assert (_affine(href, params.commit(cols[q])) == want).all(), q
    short = [c[: n // 2 + 1].copy() for c in cols[:2]]
    got = params.commit_lagrange_many(short)
    for q in range(2):
        assert (_affine(href, got[q]) == _affine(href, href.best_multiexp(short[q], g[: n // 2 + 1].copy()))).all()
    assert params.commit_many([]).shape == (0, 12)
    params.release()


def test_commit_is_p_of_s_times_g(h2b, spec, href):
    """Synthetic SRS with known s: commit(p) == [p(s)] G."""
    k, s = 6, 0x1234567
    n = 1 << k
    G = spec.affine_to_array([spec.G1_GENERATOR])[0]
    srs = np.zeros((n, 8), dtype=np.uint64)
    for i in range(n):
        srs[i] = href.g1_to_affine(href.g1_scalar_mul(G, spec.fr_array([pow(s, i, spec.R_MOD)])[0]))
    coeffs = spec.random_fr(n, 8)
    params = h2b.ParamsKZG(k, srs)
    p_s = sum(c * pow(s, i, spec.R_MOD) for i, c in enumerate(coeffs)) % spec.R_MOD
    want = href.g1_to_affine(href.g1_scalar_mul(G, spec.fr_array([p_s])[0]))
    assert (_affine(href, params.commit(spec.fr_array(coeffs))) == want).all()
    params.release()


def test_golden_vectors(h2b, href):
    for v in load_golden("spec_vectors.json")["msm"]:
        jac = h2b.best_multiexp(unhx(v["scalars"], 4), unhx(v["bases"], 8))
        assert (_affine(href, jac) == unhx(v["affine"], 8)[0]).all(), v["n"]


def test_g1_fold(h2b, spec, href):
    pts = href.random_g1(5, 31)
    jac = np.zeros((6, 12), dtype=np.uint64)
    one = spec.ints_to_array([1], spec.Q_MOD)[0]
    for i in range(5):
        jac[i, :8] = pts[i]
        jac[i, 8:] = one
    jac[5, 4:8] = one  # identity (0, R, 0)
    acc = None
    for p in spec.array_to_affine(pts):
        acc = spec.g1_add(acc, p)
    assert spec.projective_array_to_affine(h2b.g1_fold(jac)) == acc


def test_large_msm_linearity(h2b, spec, href):
    """2^20 points: MSM(a, P) + MSM(b, P) == MSM(a + b, P), and range-split partials fold to the whole."""
    n = 1 << 20
    pts = np.tile(href.random_g1(1 << 12, 50), (n >> 12, 1))
    a, b = np.tile(href.random_fr(1 << 14, 51), (n >> 14, 1)), href.random_fr(n, 52)
    ra, rb = h2b.best_multiexp(a, pts), h2b.best_multiexp(b, pts)
    # a + b elementwise via the device test hook
    import ctypes as C
    from halo2_prover_b200 import _ffi
    s = np.zeros_like(a)
    _ffi.check(_ffi.lib().h2b_test_field_op(0, 1, _ffi.u64p(a), _ffi.u64p(b), _ffi.u64p(s), C.c_size_t(n)))
    rs = h2b.best_multiexp(s, pts)
    lhs = href.g1_to_affine(href.g1_add(ra, rb))
    assert (lhs == href.g1_to_affine(rs)).all()
    lo = h2b.best_multiexp(b[: n // 2].copy(), pts[: n // 2].copy())
    hi = h2b.best_multiexp(b[n // 2:].copy(), pts[n // 2:].copy())
    assert (href.g1_to_affine(h2b.g1_fold(np.stack([lo, hi]))) == href.g1_to_affine(rb)).all()


@pytest.mark.parametrize("chunks", [2, 3, 7])
def test_chunked_host_path_matches_oracle(h2b, spec, href, chunks):
    """h2b_best_multiexp / h2b_commit overlap H2D with compute by splitting the points into chunks that
    accumulate into the same buckets; force that path at a size the oracle checks quickly."""
    import ctypes as C
    from halo2_prover_b200 import _ffi
    n = 6000
    sc, pts = href.random_fr(n, 91), href.random_g1(n, 92)
    sc[::5] = sc[0]          # skew: repeated scalar values straddle chunk boundaries
    pts[7::11] = 0           # identity bases
    want = _affine(href, href.best_multiexp(sc, pts))
    _ffi.check(_ffi.lib().h2b_set_e2e_chunking(chunks, C.c_size_t(1000)))
    try:
        assert (_affine(href, h2b.best_multiexp(sc, pts)) == want).all()
        params = h2b.ParamsKZG(13, np.concatenate([pts, href.random_g1((1 << 13) - n, 93)]))
        assert (_affine(href, params.commit(sc)) == want).all()
        params.release()
    finally:
        _ffi.check(_ffi.lib().h2b_set_e2e_chunking(0, C.c_size_t(1 << 21)))


def test_max_size_msm_2p26_linearity(h2b, spec, href):
    """BASELINE's largest MSM (2^26 points: 4 GiB of bases, 2 GiB of scalars), device-resident, checked through
    linearity: MSM(a, P) + MSM(b, P) == MSM(a + b, P), and a range split folds to the whole."""
    import ctypes as C
    import torch
    from halo2_prover_b200 import _ffi, arithmetic
    n = 1 << 26
    L = _ffi.lib()
    rng = np.random.default_rng(5)

    def rand(nn):
        a = rng.integers(0, np.iinfo(np.uint64).max, size=(nn, 4), dtype=np.uint64, endpoint=True)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        return a

    s = torch.cuda.Stream()
    gen = spec.affine_to_array([spec.G1_GENERATOR])[0]
    with torch.cuda.stream(s):
        seeds = torch.from_numpy(rand(n).view(np.int64)).cuda()
        bases = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        _ffi.check(L.h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(n), _ffi.u64p(gen),
                                            C.c_void_p(bases.data_ptr()), C.c_void_p(s.cuda_stream)))
        del seeds
        a_h, b_h = rand(n), rand(n)
        a, b = torch.from_numpy(a_h.view(np.int64)).cuda(), torch.from_numpy(b_h.view(np.int64)).cuda()
        # a + b < 2^61 * 2^192 < r: plain integer addition of the limbs IS the field addition here
        ab_h = np.zeros_like(a_h)
        carry = np.zeros(n, dtype=np.uint64)
        for limb in range(4):
            t = a_h[:, limb] + b_h[:, limb]
            c1 = (t < a_h[:, limb]).astype(np.uint64)
            t2 = t + carry
            c2 = (t2 < t).astype(np.uint64)
            ab_h[:, limb] = t2
            carry = c1 + c2
        ab = torch.from_numpy(ab_h.view(np.int64)).cuda()
        outs = torch.empty((5, 12), dtype=torch.int64, device="cuda")
        arithmetic.dev_msm(a, bases, outs[0], stream=s)
        arithmetic.dev_msm(b, bases, outs[1], stream=s)
        arithmetic.dev_msm(ab, bases, outs[2], stream=s)
        arithmetic.dev_msm(b[: n // 2], bases[: n // 2], outs[3], stream=s)
        arithmetic.dev_msm(b[n // 2:], bases[n // 2:], outs[4], stream=s)
        s.synchronize()
        # the same MSM as ParamsKZG::commit issues it: bases registered (from HBM), 2^26-point window table
        from halo2_prover_b200 import kzg
        s.synchronize()
        params = kzg.ParamsKZG.from_device(26, bases)
        outc = torch.empty((2, 12), dtype=torch.int64, device="cuda")
        params.dev_commit(ab, outc[0], stream=s)
        params.dev_commit(b[: n // 2], outc[1], stream=s)
        s.synchronize()
        params.release()
    o = outs.cpu().numpy().view(np.uint64)
    oc = outc.cpu().numpy().view(np.uint64)
    assert (href.g1_to_affine(oc[0]) == href.g1_to_affine(o[2])).all()
    assert (href.g1_to_affine(oc[1]) == href.g1_to_affine(o[3])).all()
    assert (href.g1_to_affine(href.g1_add(o[0], o[1])) == href.g1_to_affine(o[2])).all()
    assert (href.g1_to_affine(href.g1_add(o[3], o[4])) == href.g1_to_affine(o[1])).all()
    assert spec.g1_is_on_curve(spec.projective_array_to_affine(o[2]))


@pytest.mark.parametrize("k", [0, 1, 3, 8])
def test_g_to_lagrange_vs_oracle_msm(h2b, spec, href, k):
    """g_lagrange[i] = sum_j [omega^(-ij) / n] g[j] on random bases (with an identity among them)."""
    n = 1 << k
    g = href.random_g1(n, 700 + k)
    if n > 4:
        g[2] = 0
    got = h2b.g_to_lagrange(g, k)
    w_inv = pow(pow(spec.ROOT_OF_UNITY, 1 << (28 - k), spec.R_MOD), -1, spec.R_MOD)
    n_inv = pow(n, -1, spec.R_MOD)
    for i in sorted({0, 1 % n, (n // 2 + 1) % n, n - 1}):
        sc = spec.fr_array([pow(w_inv, i * j, spec.R_MOD) * n_inv % spec.R_MOD for j in range(n)])
        assert (_affine(href, href.best_multiexp(sc, g)) == got[i]).all(), i


@pytest.mark.parametrize("t", [2, 3, 4])
def test_thinned_window_table_vs_oracle(h2b, href, t):
    """h2b_set_srs_table_stride: only every t-th window power is tabulated (1/t of the table's HBM), commits use t bucket
    sets and a Horner over their sums.  Single commits, a shorter polynomial, a batch, and the chunked host path must
    equal the oracle (and so the t = 1 results)."""
    import ctypes as C
    from halo2_prover_b200 import _ffi
    L = _ffi.lib()
    n = 1 << 13
    g = href.random_g1(n, 300 + t)
    g[5] = 0
    cols = [href.random_fr(n, 310 + 10 * t + q) for q in range(3)]
    cols[1][::2] = cols[1][0]
    _ffi.check(L.h2b_set_srs_precompute(2, 0))   # window table at this size too
    _ffi.check(L.h2b_set_srs_table_stride(t))
    try:
        params = h2b.ParamsKZG(13, g)
        c, w, tb = C.c_uint32(), C.c_uint32(), C.c_size_t()
        _ffi.check(L.h2b_srs_info(C.c_uint64(params._handles["g"]), None, C.byref(c), C.byref(w), C.byref(tb)))
        assert tb.value == ((w.value + t - 1) // t) * n * 64
        want = [_affine(href, href.best_multiexp(p, g)) for p in cols]
        for q in range(3):
            assert (_affine(href, params.commit(cols[q])) == want[q]).all(), q
        many = params.commit_many(cols)
        for q in range(3):
            assert (_affine(href, many[q]) == want[q]).all(), ("many", q)
        short = cols[0][:3001].copy()
        assert (_affine(href, params.commit(short)) == _affine(href, href.best_multiexp(short, g[:3001].copy()))).all()
        _ffi.check(L.h2b_set_e2e_chunking(3, C.c_size_t(1000)))
        assert (_affine(href, params.commit(cols[2])) == want[2]).all(), "chunked"
        params.release()
    finally:
        _ffi.check(L.h2b_set_e2e_chunking(0, C.c_size_t(1 << 21)))
        _ffi.check(L.h2b_set_srs_table_stride(0))
        _ffi.check(L.h2b_set_srs_precompute(1, 0))


def test_concurrent_host_threads(h2b, spec, href):
    """The library serialises its entry points on one lock (one workspace per device): calls from several host threads
    at once -- rayon workers in the reference's create_proof, plonk/prover.rs -- must all return the right answer."""
    import threading

    jobs = [(href.random_fr(n, 7100 + n), href.random_g1(n, 7200 + n)) for n in (300, 4096, 5000, 1 << 14)]
    want = [_affine(href, href.best_multiexp(s, p)) for s, p in jobs]
    polys = [href.random_fr(1 << k, 7300 + k) for k in (10, 12, 13)]
    omega = lambda k: spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]  # noqa: E731
    want_fft = [href.best_fft(a, omega(a.shape[0].bit_length() - 1), a.shape[0].bit_length() - 1) for a in polys]
    got, got_fft, errs = [None] * len(jobs), [None] * len(polys), []

    def msm(i):
        try:
            for _ in range(3):
                got[i] = _affine(href, h2b.best_multiexp(jobs[i][0], jobs[i][1]))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    def fft(i):
        try:
            for _ in range(3):
                a = polys[i].copy()
                k = a.shape[0].bit_length() - 1
                h2b.best_fft(a, omega(k), k)
                got_fft[i] = a
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=msm, args=(i,)) for i in range(len(jobs))]
    ts += [threading.Thread(target=fft, args=(i,)) for i in range(len(polys))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for g, w in zip(got, want):
        assert (g == w).all()
    for g, w in zip(got_fft, want_fft):
        assert (g == w).all()


def test_params_setup_vs_oracle(h2b, spec, href):
    """ParamsKZG.setup(k, s) (the reference's generate_params with the secret given): g[i] = [s^i] G against the oracle's
    scalar multiplication, and g_lagrange through its defining property -- the commitment of a polynomial from its
    coefficients equals the commitment from its evaluations."""
    k, s = 6, 0x1234567_89ABCDEF_0FEDCBA9
    n = 1 << k
    params = h2b.ParamsKZG.setup(k, s)
    blob = params.write(bytes(256))
    g = np.frombuffer(blob, dtype=np.uint64, count=8 * n, offset=4).reshape(n, 8)
    gen = spec.affine_to_array([spec.G1_GENERATOR])[0]
    for i in (0, 1, 2, n - 1):
        want = href.g1_to_affine(href.g1_scalar_mul(gen, spec.fr_array([pow(s, i, spec.R_MOD)])[0]))
        assert (g[i] == want).all(), i
    coeffs = href.random_fr(n, 4242)
    omega = spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]
    evals = href.best_fft(coeffs, omega, k)
    assert (_affine(href, params.commit(coeffs)) == _affine(href, params.commit_lagrange(evals))).all()
    assert (_affine(href, params.commit(coeffs)) == _affine(href, href.best_multiexp(coeffs, np.ascontiguousarray(g)))).all()
    params.release()
