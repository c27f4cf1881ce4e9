"""GPU parity: device field arithmetic and the G1 mixed-add path against the oracle (bit-exact)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _field_op(h2b, field, op, a, b=None):
    from halo2_prover_b200 import _ffi
    out = np.zeros_like(a)
    bp = _ffi.u64p(b) if b is not None else None
    _ffi.check(_ffi.lib().h2b_test_field_op(field, op, _ffi.u64p(a), bp, _ffi.u64p(out), C.c_size_t(a.shape[0])))
    return out


def _edge(spec, mod, n, seed):
    """random values plus 0, 1, p-1, p-2, 2^k patterns, in Montgomery limbs"""
    vals = [0, 1, 2, mod - 1, mod - 2, (1 << 253), (1 << 253) - 1, (1 << 32) - 1, (1 << 64), mod >> 1]
    g = spec.SplitMix64(seed)
    while len(vals) < n:
        v = sum(g.next() << (64 * i) for i in range(4)) % mod
        vals.append(v)
    return vals


@pytest.mark.parametrize("field", [0, 1])
def test_mul_add_sub(h2b, spec, field):
    mod = spec.R_MOD if field == 0 else spec.Q_MOD
    n = 4096
    av, bv = _edge(spec, mod, n, 1), list(reversed(_edge(spec, mod, n, 2)))
    a, b = spec.ints_to_array(av, mod), spec.ints_to_array(bv, mod)
    want_mul = spec.ints_to_array([x * y % mod for x, y in zip(av, bv)], mod)
    assert (_field_op(h2b, field, 0, a, b) == want_mul).all()
    assert (_field_op(h2b, field, 3, a, b) == want_mul).all()  # portable path agrees
    assert (_field_op(h2b, field, 1, a, b) == spec.ints_to_array([(x + y) % mod for x, y in zip(av, bv)], mod)).all()
    assert (_field_op(h2b, field, 2, a, b) == spec.ints_to_array([(x - y) % mod for x, y in zip(av, bv)], mod)).all()


@pytest.mark.parametrize("field", [0, 1])
def test_inverse_and_from_mont(h2b, spec, field):
    mod = spec.R_MOD if field == 0 else spec.Q_MOD
    av = _edge(spec, mod, 256, 3)
    a = spec.ints_to_array(av, mod)
    inv = _field_op(h2b, field, 4, a)
    assert (inv == spec.ints_to_array([pow(x, -1, mod) if x else 0 for x in av], mod)).all()
    canon = _field_op(h2b, field, 5, a)
    assert (canon == spec.ints_to_array(av, None)).all()


@pytest.mark.parametrize("field", [0, 1])
def test_dedicated_square(h2b, spec, field):
    """Field::sqr forms every cross product once, doubled (36 limb products instead of 64): must equal a * a for
    random elements and for limb patterns that maximise the doubled partial sums."""
    mod = spec.R_MOD if field == 0 else spec.Q_MOD
    top = mod >> 224
    patterns = [mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1,
                ((top - 1) << 224) | ((1 << 224) - 1),            # every lower limb 0xffffffff
                ((top - 1) << 224) | int("80000000" * 7, 16),     # top bit of every lower limb (the funnel shifts)
                int("7fffffff" * 7, 16), int("ffffffff" * 7, 16), (1 << 253) + (1 << 31), (1 << 224) - 1, 1 << 223]
    vals = _edge(spec, mod, 8192, 11) + [v % mod for v in patterns]
    a = spec.ints_to_array(vals, mod)          # Montgomery form of vals
    want = spec.ints_to_array([v * v % mod for v in vals], mod)
    assert (_field_op(h2b, field, 6, a) == want).all()
    # the limb patterns themselves as Montgomery REPRESENTATIVES (what the multiplier actually sees)
    raw = spec.ints_to_array([v % mod for v in patterns], None)
    assert (_field_op(h2b, field, 6, raw) == _field_op(h2b, field, 0, raw, raw)).all()


@pytest.mark.parametrize("field", [0, 1])
def test_fused_two_product_multiply(h2b, spec, field):
    """Field::mul2_add: (a b + c d) R^-1 under one Montgomery reduction, and mul2_sub through neg()."""
    mod = spec.R_MOD if field == 0 else spec.Q_MOD
    n = 8192
    av, bv = _edge(spec, mod, n, 21), list(reversed(_edge(spec, mod, n, 22)))
    av[:4], bv[:4] = [mod - 1, mod - 1, 0, mod - 2], [mod - 1, 1, mod - 1, mod - 1]
    a, b = spec.ints_to_array(av, mod), spec.ints_to_array(bv, mod)
    want = spec.ints_to_array([(x * y + (x + y) * (x - y)) % mod for x, y in zip(av, bv)], mod)
    assert (_field_op(h2b, field, 7, a, b) == want).all()
    assert (_field_op(h2b, field, 8, a, b) == 0).all()
    # maximal limb patterns as raw representatives
    top = mod >> 224
    raw = spec.ints_to_array([((top - 1) << 224) | ((1 << 224) - 1), mod - 1, ((top - 1) << 224) | int("80000000" * 7, 16)], None)
    rb = raw[::-1].copy()
    x, y = _field_op(h2b, field, 1, raw, rb), _field_op(h2b, field, 2, raw, rb)
    want = _field_op(h2b, field, 1, _field_op(h2b, field, 0, raw, rb), _field_op(h2b, field, 0, x, y))
    assert (_field_op(h2b, field, 7, raw, rb) == want).all()


def test_mul_against_c_oracle_large(h2b, href):
    a, b = href.random_fr(1 << 16, 5), href.random_fr(1 << 16, 6)
    assert (_field_op(h2b, 0, 0, a, b) == href.fr_mul(a, b)).all()
    pa, pb = href.random_g1(1 << 15, 7), href.random_g1(1 << 15, 8)
    a, b = pa.reshape(-1, 4).copy(), pb.reshape(-1, 4).copy()
    assert (_field_op(h2b, 1, 0, a, b) == href.fq_mul(a, b)).all()


def test_g1_mixed_add_special_cases(h2b, spec, href):
    from halo2_prover_b200 import _ffi
    pts = spec.random_g1(64, 21)
    a = list(pts[:32])
    b = list(pts[32:])
    a[0], b[0] = None, pts[1]            # O + P
    a[1], b[1] = pts[2], None            # P + O
    a[2], b[2] = None, None              # O + O
    a[3], b[3] = pts[3], pts[3]          # P + P (doubling)
    a[4], b[4] = pts[4], spec.g1_neg(pts[4])  # P + (-P)
    A, B = spec.affine_to_array(a), spec.affine_to_array(b)
    out = np.zeros((32, 12), dtype=np.uint64)
    _ffi.check(_ffi.lib().h2b_test_g1_add_affine(_ffi.u64p(A), _ffi.u64p(B), _ffi.u64p(out), C.c_size_t(32)))
    for i in range(32):
        assert spec.projective_array_to_affine(out[i]) == spec.g1_add(a[i], b[i]), i
    # identity is reported as (0, R, 0) like G1::identity()
    assert out[2][8:].sum() == 0 and (out[2][4:8] == spec.ints_to_array([1], spec.Q_MOD)[0]).all()


@pytest.mark.parametrize("field", [0, 1])
def test_lazy_butterfly_arithmetic(h2b, spec, field):
    """The NTT butterflies keep values in [0, 2N) and multiply values below 4N by canonical twiddles without the final
    subtraction (field.cuh: mul_lazy, add_2n, sub_2n, reduce_2n).  Raw (non-canonical) limb patterns up to the
    bounds the kernel can produce, including the extremes, must give the same field elements as big integers."""
    mod = spec.R_MOD if field == 0 else spec.Q_MOD
    rinv = pow(1 << 256, -1, mod)
    g = spec.SplitMix64(77 + field)

    def raw(vals):
        return np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64)

    def rnd(bound):
        return sum(g.next() << (64 * i) for i in range(4)) % bound

    n = 4096
    # product: a < 4N raw, b < N raw -> a * b * R^-1 mod N
    av = [4 * mod - 1, 4 * mod - 2, 3 * mod, 2 * mod, 2 * mod - 1, mod, 0, 1, 4 * mod - 3]
    bv = [mod - 1, mod - 2, mod - 1, mod - 1, 1, mod - 1, mod - 1, 0, mod - 1]
    while len(av) < n:
        av.append(rnd(4 * mod))
        bv.append(rnd(mod))
    assert max(av) < min(4 * mod, 1 << 256)
    got = _field_op(h2b, field, 9, raw(av), raw(bv))
    assert (got == raw([x * y * rinv % mod for x, y in zip(av, bv)])).all()
    # difference / sum: a, b < 2N raw
    xv = [2 * mod - 1, 0, 2 * mod - 1, 0, mod, mod - 1, 1, 2 * mod - 2]
    yv = [0, 2 * mod - 1, 2 * mod - 1, 0, mod, mod, 2 * mod - 1, 1]
    while len(xv) < n:
        xv.append(rnd(2 * mod))
        yv.append(rnd(2 * mod))
    assert (_field_op(h2b, field, 10, raw(xv), raw(yv)) == raw([(x - y) % mod for x, y in zip(xv, yv)])).all()
    assert (_field_op(h2b, field, 11, raw(xv), raw(yv)) == raw([(x + y) % mod for x, y in zip(xv, yv)])).all()
