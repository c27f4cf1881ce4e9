"""scripts/field52.cuh (5 x 52-bit limbs, products formed by FMA pairs) compiled for the host and checked
against big-integer arithmetic.  The device runs the same source with __fma_rz; tests/test_field_gpu.py
repeats the comparison there."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "field52_host.cpp")
HDR = os.path.join(HERE, "..", "scripts", "field52.cuh")
LIB = os.path.join(HERE, "native", "libfield52_host.so")
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
M52 = (1 << 52) - 1
R260_INV = pow(1 << 260, -1, Q)


@pytest.fixture(scope="module")
def f52():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O1", "-mfma", "-frounding-math", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return C.CDLL(LIB)


def limbs(x):
    return [(x >> (52 * i)) & M52 for i in range(5)]


def arr(vals):
    return np.array([limbs(v) for v in vals], dtype=np.uint64)


def ints(a):
    return [sum(int(a[i, j]) << (52 * j) for j in range(5)) for i in range(a.shape[0])]


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def _operands(rng, n, bound):
    vals = [rng.randrange(bound) for _ in range(n)]
    edge = [0, 1, Q - 1, Q, Q + 1, 2 * Q, bound - 1, (1 << 208) - 1, 1 << 208, M52, (1 << 260) - 1 if bound >= 1 << 260 else bound - 1]
    return [v for v in edge if v < bound] + vals


@pytest.mark.parametrize("bits", [254, 257, 259])
def test_mul_and_sqr_match_big_integers(f52, bits):
    rng = random.Random(bits)
    a = _operands(rng, 3000, 1 << bits)
    b = list(reversed(_operands(rng, 3000, 1 << bits)))
    A, B = arr(a), arr(b)
    out = np.zeros_like(A)
    f52.f52h_mul(p(A), p(B), p(out), C.c_size_t(len(a)))
    assert (out <= M52).all(), "limbs must come back normalised"
    for x, y, r in zip(a, b, ints(out)):
        assert r % Q == x * y * R260_INV % Q
        assert r <= (x * y >> 260) + Q  # the bound the lazy reduction relies on
    f52.f52h_sqr(p(A), p(out), C.c_size_t(len(a)))
    assert (out <= M52).all()
    for x, r in zip(a, ints(out)):
        assert r % Q == x * x * R260_INV % Q
        assert r <= (x * x >> 260) + Q


def test_all_ones_limbs_and_carry_extremes(f52):
    # every limb 2^52 - 1 on both sides drives every column accumulator to its maximum
    a = [(1 << 260) - 1, (1 << 260) - 1, M52 << 208, M52, sum(M52 << (104 * i) for i in range(3))]
    b = [(1 << 260) - 1, M52 << 208, M52 << 208, (1 << 260) - 1, sum(M52 << (52 + 104 * i) for i in range(2))]
    A, B = arr(a), arr(b)
    out = np.zeros_like(A)
    f52.f52h_mul(p(A), p(B), p(out), C.c_size_t(len(a)))
    # the top limb may exceed 52 bits here (inputs at 2^260 are outside the documented domain); value must still be right
    for x, y, r in zip(a, b, ints(out)):
        assert r % Q == x * y * R260_INV % Q


def test_sub_add_lazy(f52):
    rng = random.Random(7)
    n = 2000
    for k in (2, 4, 6, 8):
        a = [rng.randrange(1 << 257) for _ in range(n)] + [0, 0]
        b = [rng.randrange(k * Q + 1) for _ in range(n)] + [k * Q, 0]
        A, B = arr(a), arr(b)
        out = np.zeros_like(A)
        f52.f52h_sub(k, p(A), p(B), p(out), C.c_size_t(len(a)))
        assert (out <= M52).all()
        assert ints(out) == [x - y + k * Q for x, y in zip(a, b)]
    a = [rng.randrange(1 << 256) for _ in range(n)]
    b = [rng.randrange(2 * Q) for _ in range(n)]
    c = [rng.randrange(Q) for _ in range(n)]
    A, B, Cc = arr(a), arr(b), arr(c)
    out = np.zeros_like(A)
    f52.f52h_sub_b_2c(p(A), p(B), p(Cc), p(out), C.c_size_t(n))
    assert ints(out) == [x - y - 2 * z + 4 * Q for x, y, z in zip(a, b, c)]
    f52.f52h_add(p(A), p(B), p(out), C.c_size_t(n))
    assert ints(out) == [x + y for x, y in zip(a, b)]


def test_pack_unpack_and_zero_test(f52):
    rng = random.Random(9)
    vals = [0, 1, Q, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(500)]
    w = np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for v in vals], dtype=np.uint32)
    out = np.zeros((len(vals), 5), dtype=np.uint64)
    f52.f52h_unpack(p(w), p(out), C.c_size_t(len(vals)))
    assert ints(out) == vals
    back = np.zeros_like(w)
    f52.f52h_pack(p(out), p(back), C.c_size_t(len(vals)))
    assert (back == w).all()
    for v, want in ((0, 1), (Q, 1), (1, 0), (Q - 1, 0), (Q + 1, 0), (2 * Q - 1, 0)):
        assert f52.f52h_is_zero(p(arr([v]))) == want
