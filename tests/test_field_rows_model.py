"""Host model of the device Montgomery routines' accumulator discipline (csrc/field.cuh: mul, sqr, mul2_add).

The device code keeps two staggered 8-limb accumulators (E at limb 0, O at limb 1) and several of its carry chains
end WITHOUT a carry out because "the running value fits".  This model replays the same rows on Python integers and
asserts exactly those claims (no dropped carry anywhere) for random and for maximal inputs, and that the result is the
Montgomery product.  It is the written-out version of the bound argument in DESIGN.md section 4.1; the bit-exact GPU
checks are tests/test_field_gpu.py.
"""
import random

import pytest

B = 1 << 32
R = 1 << 256
MODS = {
    "Fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "Fq": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
}


def limbs(v):
    return [(v >> (32 * i)) & (B - 1) for i in range(8)]


class Acc:
    """An 8-limb accumulator as an integer below 2^256."""
    def __init__(self):
        self.v = 0


def cmad_row(acc, xs, m):
    """acc += x0 m + x1 m 2^64 + x2 m 2^128 + x3 m 2^192, no carry out allowed."""
    acc.v += sum(x * m << (64 * k) for k, x in enumerate(xs))
    assert acc.v < R, "cmad_row dropped a carry"


def cmad_row_fold(acc, other, xs, m):
    """Same, with the carry out added to the top limb of the other accumulator (one limb above acc's top)."""
    acc.v += sum(x * m << (64 * k) for k, x in enumerate(xs))
    carry, acc.v = acc.v >> 256, acc.v & (R - 1)
    assert carry <= 1
    top = (other.v >> 224) + carry
    assert top < B, "fold overflowed the top limb"
    other.v += carry << 224


def shift_mad_row(e, o, xs, m):
    """e[0] += o[1]; o <- (o >> 64) + products, the carry of the first addition continuing into o."""
    assert o.v & (B - 1) == 0, "limb 0 of the shifted accumulator must have been cleared by the reduction"
    e0 = (e.v & (B - 1)) + ((o.v >> 32) & (B - 1))
    e.v = (e.v & ~(B - 1)) | (e0 & (B - 1))
    o.v = (o.v >> 64) + (e0 >> 32) + sum(x * m << (64 * k) for k, x in enumerate(xs))
    assert o.v < R, "shift_mad_row dropped a carry"


def redc_step(e, o, n, m0):
    m = (e.v & (B - 1)) * m0 & (B - 1)
    cmad_row(o, n[1::2], m)
    cmad_row_fold(e, o, n[0::2], m)
    assert e.v & (B - 1) == 0


def run(mod, rows):
    """rows[i] = list of (multiplicand limbs, multiplier) pairs added in row i."""
    n, m0 = limbs(mod), (-pow(mod, -1, B)) % B
    ev, od = Acc(), Acc()
    e, o = ev, od
    for i, terms in enumerate(rows):
        if i:
            e, o = o, e
            (xs, m), rest = terms[0], terms[1:]
            shift_mad_row(e, o, xs[1::2], m)
            cmad_row_fold(e, o, xs[0::2], m)
        else:
            (xs, m), rest = terms[0], terms[1:]
            e.v = sum(x * m << (64 * k) for k, x in enumerate(xs[0::2]))
            o.v = sum(x * m << (64 * k) for k, x in enumerate(xs[1::2]))
        for xs, m in rest:
            cmad_row(o, xs[1::2], m)
            cmad_row_fold(e, o, xs[0::2], m)
        redc_step(e, o, n, m0)
    # after the last row E = od, O = ev: result = O + (E >> 32)
    r = o.v + (e.v >> 32)
    assert r < R, "final addition dropped a carry"
    assert r < 2 * mod, "one conditional subtraction must suffice"
    return r - mod if r >= mod else r


def model_mul(mod, a, b):
    al, bl = limbs(a), limbs(b)
    return run(mod, [[(al, bl[i])] for i in range(8)])


def model_mul2_add(mod, a, b, c, d):
    al, bl, cl, dl = limbs(a), limbs(b), limbs(c), limbs(d)
    return run(mod, [[(al, bl[i]), (cl, dl[i])] for i in range(8)])


def model_sqr(mod, a):
    al = limbs(a)
    t = limbs((2 * a) % R)
    u = [(x << 1) & (B - 1) for x in al]
    rows = []
    for i in range(8):
        xs = [0 if j < i else al[j] if j == i else u[j] if j == i + 1 else t[j] for j in range(8)]
        rows.append([(xs, al[i])])
    return run(mod, rows)


def _inputs(mod, count, seed):
    rng = random.Random(seed)
    top = mod >> 224
    fixed = [0, 1, mod - 1, mod - 2, mod, mod >> 1, ((top - 1) << 224) | ((1 << 224) - 1),
             ((top - 1) << 224) | int("80000000" * 7, 16), (1 << 224) - 1, int("ffffffff00000000" * 3 + "ffffffff", 16)]
    high = [((top - rng.randrange(2)) << 224 | sum(rng.randrange(B >> 1, B) << (32 * i) for i in range(7))) % mod
            for _ in range(count)]  # every lower limb in [2^31, 2^32)
    return fixed + high + [rng.randrange(mod) for _ in range(count)]


@pytest.mark.parametrize("name", ["Fr", "Fq"])
def test_rows_never_drop_a_carry(name):
    mod = MODS[name]
    rinv = pow(R, -1, mod)
    vals = _inputs(mod, 150, 7)
    rng = random.Random(11)
    for a in vals:
        assert model_sqr(mod, a) == a * a * rinv % mod
        b, c, d = rng.choice(vals), rng.choice(vals), rng.choice(vals)
        assert model_mul(mod, a, b) == a * b * rinv % mod
        assert model_mul2_add(mod, a, b, c, d) == (a * b + c * d) * rinv % mod
    big = [v for v in vals[:10]]
    for a in big:          # every combination of the maximal patterns through the two-product rows
        for b in big:
            assert model_mul2_add(mod, a, b, b, a) == 2 * a * b * rinv % mod
            assert model_mul2_add(mod, a, a, b, b) == (a * a + b * b) * rinv % mod
