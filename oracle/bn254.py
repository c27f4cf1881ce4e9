"""Big-integer specification oracle for the BN254 MSM / NTT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``halo2-prover_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and only
as the checker.

Parity status: **pinned by uniqueness + constants, not by reference KATs** for
the Python functions in this file (the reference's own tests hold no MSM/NTT
vectors).  The vectors captured from the reference's compiled
prover (``tests/golden/wasm_*.json``, produced by ``oracle/wasm/``) are what pin
the oracle against the reference's *execution*; see DESIGN.md §3.

What is restated here (upstream = the un-vendored, Cargo.lock-pinned
dependencies of /root/reference/circuits, Cargo.lock:836-838 and :854-856):

* ``halo2curves 0.3.2 @9f5c508 src/bn256/{fr,fq,curve}.rs`` -- field moduli,
  Montgomery form (R = 2^256, 4 x u64 little-endian limbs), y^2 = x^3 + 3.
* ``halo2_proofs @6b43b6b src/arithmetic.rs:28-140`` multiexp_serial,
  ``:147-180`` best_multiexp, ``:185-290`` best_fft.
* ``halo2_proofs @6b43b6b src/poly/domain.rs`` EvaluationDomain::new and the
  three transforms (lagrange_to_coeff :227, coeff_to_extended :244,
  extended_to_coeff :311).

Everything is computed on Python integers in *canonical* form; ``to_mont`` /
``from_mont`` convert to the memory layout at the FFI boundary.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --- constants (SURVEY.md section 8 header; verified against the reference binary) ------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
MONT_R = 1 << 256
FR_S = 28
FR_GENERATOR = 7
ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
# Fr::ZETA of halo2curves 0.3.2 (the coset generator of EvaluationDomain).  SURVEY.md quotes its SQUARE
# (0x30644e72...36636f23, found as an immediate in EvaluationDomain::new: that is g_coset_inv); the reference's
# recorded coeff_to_extended calls multiply coefficient 1 by THIS value (tests/test_wasm_golden.py,
# tests/test_evaluate_h.py), which settles it.
ZETA = 0xB3C4D79D41A917585BFC41088D8DAAA78B17EA66B99C90DD
INV_R = 0xC2E1F593EFFFFFFF  # -r^{-1} mod 2^64
INV_Q = 0x87D20782E4866389  # -q^{-1} mod 2^64
CURVE_B = 3
G1_GENERATOR = (1, 2)

assert pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD) == ROOT_OF_UNITY
assert pow(ZETA, 3, R_MOD) == 1 and ZETA != 1
assert (-pow(R_MOD, -1, 1 << 64)) % (1 << 64) == INV_R
assert (-pow(Q_MOD, -1, 1 << 64)) % (1 << 64) == INV_Q


# --- Montgomery layout helpers -------------------------------------------------------------
def to_mont(x: int, mod: int) -> int:
    return (x * MONT_R) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * pow(MONT_R, -1, mod)) % mod


def int_to_limbs(x: int) -> List[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def limbs_to_int(l: Sequence[int]) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def ints_to_array(vals: Sequence[int], mod: Optional[int]) -> np.ndarray:
    """Canonical ints -> (n,4) uint64 array; Montgomery-encoded when ``mod`` given."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        if mod is not None:
            v = to_mont(v, mod)
        out[i] = int_to_limbs(v)
    return out


def array_to_ints(arr: np.ndarray, mod: Optional[int]) -> List[int]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    out = []
    for row in arr:
        v = limbs_to_int(row)
        if mod is not None:
            v = from_mont(v, mod)
        out.append(v)
    return out


def fr_array(vals: Sequence[int]) -> np.ndarray:
    return ints_to_array(vals, R_MOD)


def fr_ints(arr: np.ndarray) -> List[int]:
    return array_to_ints(arr, R_MOD)


# --- G1 (y^2 = x^3 + 3 over Fq), affine with None = identity -----------------------------------
Affine = Optional[Tuple[int, int]]


def g1_is_on_curve(p: Affine) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(p: Affine) -> Affine:
    if p is None:
        return None
    return (p[0], (-p[1]) % Q_MOD)


# Jacobian arithmetic on ints for speed (one inversion at the end).
def _jac_double(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A = X * X % Q_MOD
    B = Y * Y % Q_MOD
    C = B * B % Q_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % Q_MOD
    E = 3 * A % Q_MOD
    F = E * E % Q_MOD
    X3 = (F - 2 * D) % Q_MOD
    Y3 = (E * (D - X3) - 8 * C) % Q_MOD
    Z3 = 2 * Y * Z % Q_MOD
    return (X3, Y3, Z3)


def _jac_add(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % Q_MOD
    Z2Z2 = Z2 * Z2 % Q_MOD
    U1 = X1 * Z2Z2 % Q_MOD
    U2 = X2 * Z1Z1 % Q_MOD
    S1 = Y1 * Z2 * Z2Z2 % Q_MOD
    S2 = Y2 * Z1 * Z1Z1 % Q_MOD
    if U1 == U2:
        if S1 == S2:
            return _jac_double(p)
        return (0, 1, 0)
    H = (U2 - U1) % Q_MOD
    Rr = (S2 - S1) % Q_MOD
    HH = H * H % Q_MOD
    HHH = H * HH % Q_MOD
    V = U1 * HH % Q_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % Q_MOD
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q_MOD
    Z3 = Z1 * Z2 * H % Q_MOD
    return (X3, Y3, Z3)


def _to_jac(p: Affine):
    return (0, 1, 0) if p is None else (p[0], p[1], 1)


def _from_jac(p) -> Affine:
    X, Y, Z = p
    if Z % Q_MOD == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g1_add(p: Affine, q: Affine) -> Affine:
    return _from_jac(_jac_add(_to_jac(p), _to_jac(q)))


def g1_mul(p: Affine, k: int) -> Affine:
    k %= R_MOD
    acc = (0, 1, 0)
    base = _to_jac(p)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(base)
        k >>= 1
    return _from_jac(acc)


def msm_naive(scalars: Sequence[int], points: Sequence[Affine]) -> Affine:
    """sum_i scalars[i] * points[i]; double-and-add per term (small n only)."""
    assert len(scalars) == len(points)
    acc = (0, 1, 0)
    for k, p in zip(scalars, points):
        k %= R_MOD
        if p is None or k == 0:
            continue
        base = _to_jac(p)
        t = (0, 1, 0)
        while k:
            if k & 1:
                t = _jac_add(t, base)
            base = _jac_double(base)
            k >>= 1
        acc = _jac_add(acc, t)
    return _from_jac(acc)


def multiexp_window(n: int) -> int:
    """Window rule of multiexp_serial (h2p@6b43b6b src/arithmetic.rs:~33-41)."""
    if n < 4:
        return 1
    if n < 32:
        return 3
    return int(math.ceil(math.log(float(n))))


def multiexp_serial(scalars: Sequence[int], points: Sequence[Affine], acc=(0, 1, 0)):
    """Pippenger exactly as arithmetic.rs:28-140: unsigned c-bit windows over the
    256-bit canonical little-endian repr, segments = 256/c + 1, high to low, c
    doublings of the accumulator per segment, 2^c - 1 buckets, running-sum
    reduction.  Returns a Jacobian triple of ints."""
    n = len(points)
    c = multiexp_window(n)
    segments = 256 // c + 1
    reprs = [int(s % R_MOD).to_bytes(32, "little") for s in scalars]

    def get_at(segment: int, rep: bytes) -> int:
        skip_bits = segment * c
        skip_bytes = skip_bits // 8
        if skip_bytes >= 32:
            return 0
        v = rep[skip_bytes:skip_bytes + 8].ljust(8, b"\0")
        tmp = int.from_bytes(v, "little") >> (skip_bits - skip_bytes * 8)
        return tmp % (1 << c)

    for seg in range(segments - 1, -1, -1):
        for _ in range(c):
            acc = _jac_double(acc)
        buckets = [(0, 1, 0)] * ((1 << c) - 1)
        for rep, p in zip(reprs, points):
            w = get_at(seg, rep)
            if w != 0 and p is not None:
                buckets[w - 1] = _jac_add(buckets[w - 1], _to_jac(p))
        running = (0, 1, 0)
        for b in reversed(buckets):
            running = _jac_add(running, b)
            acc = _jac_add(acc, running)
    return acc


def best_multiexp(scalars: Sequence[int], points: Sequence[Affine], num_threads: int = 1) -> Affine:
    """arithmetic.rs:147-180: contiguous chunks of len/num_threads, fold partials."""
    n = len(scalars)
    assert n == len(points)  # arithmetic.rs:148
    if n > num_threads:
        chunk = n // num_threads
        acc = (0, 1, 0)
        for s in range(0, n, chunk):
            acc = _jac_add(acc, multiexp_serial(scalars[s:s + chunk], points[s:s + chunk]))
        return _from_jac(acc)
    return _from_jac(multiexp_serial(scalars, points))


def affine_to_array(points: Sequence[Affine]) -> np.ndarray:
    """(n,8) uint64: x limbs then y limbs, Montgomery; identity = all zero."""
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            continue
        out[i, :4] = int_to_limbs(to_mont(p[0], Q_MOD))
        out[i, 4:] = int_to_limbs(to_mont(p[1], Q_MOD))
    return out


def array_to_affine(arr: np.ndarray) -> List[Affine]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out: List[Affine] = []
    for row in arr:
        x = limbs_to_int(row[:4])
        y = limbs_to_int(row[4:])
        if x == 0 and y == 0:
            out.append(None)
        else:
            out.append((from_mont(x, Q_MOD), from_mont(y, Q_MOD)))
    return out


def projective_array_to_affine(arr: np.ndarray) -> Affine:
    """12 x u64 (x,y,z Montgomery, HOMOGENEOUS projective x = X/Z, y = Y/Z -- the reference's G1,
    established by executing its compiled prover; identity z=0) -> affine ints."""
    arr = np.asarray(arr, dtype=np.uint64).reshape(12)
    X = from_mont(limbs_to_int(arr[0:4]), Q_MOD)
    Y = from_mont(limbs_to_int(arr[4:8]), Q_MOD)
    Z = from_mont(limbs_to_int(arr[8:12]), Q_MOD)
    if Z == 0:
        return None
    zi = pow(Z, -1, Q_MOD)
    return (X * zi % Q_MOD, Y * zi % Q_MOD)


# --- NTT -----------------------------------------------------------------------------------------
def bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(a: List[int], omega: int, log_n: int) -> List[int]:
    """arithmetic.rs:185-250 (iterative branch): bit-reverse, sequential twiddle
    scan, radix-2 DIT; natural order in, natural order out.  Returns a new list."""
    n = len(a)
    assert n == 1 << log_n  # arithmetic.rs:199
    a = list(a)
    for k in range(n):
        rk = bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    tw = [1] * max(n // 2, 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R_MOD
    chunk, tchunk = 2, n // 2
    for _ in range(log_n):
        half = chunk // 2
        for s in range(0, n, chunk):
            for i in range(half):
                t = a[s + half + i] * tw[i * tchunk] % R_MOD
                u = a[s + i]
                a[s + i] = (u + t) % R_MOD
                a[s + half + i] = (u - t) % R_MOD
        chunk *= 2
        tchunk //= 2
    return a


def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


class EvaluationDomain:
    """poly/domain.rs EvaluationDomain::new(j, k) and the three transforms."""

    def __init__(self, j: int, k: int):
        self.k = k
        self.j = j
        self.quotient_poly_degree = j - 1
        self.n = 1 << k
        ext_k = k
        while (1 << ext_k) < self.n * self.quotient_poly_degree:
            ext_k += 1
        self.extended_k = ext_k
        w = ROOT_OF_UNITY
        for _ in range(ext_k, FR_S):
            w = w * w % R_MOD
        self.extended_omega = w
        self.extended_omega_inv = pow(w, -1, R_MOD)
        for _ in range(k, ext_k):
            w = w * w % R_MOD
        self.omega = w
        self.omega_inv = pow(w, -1, R_MOD)
        self.g_coset = ZETA
        self.g_coset_inv = ZETA * ZETA % R_MOD
        orig = pow(ZETA, self.n, R_MOD)
        step = pow(self.extended_omega, self.n, R_MOD)
        t = []
        cur = orig
        while True:
            t.append(cur)
            cur = cur * step % R_MOD
            if cur == orig:
                break
        assert len(t) == 1 << (ext_k - k)  # domain.rs:101
        self.t_evaluations = [pow((x - 1) % R_MOD, -1, R_MOD) for x in t]
        self.ifft_divisor = pow(1 << k, -1, R_MOD)
        self.extended_ifft_divisor = pow(1 << ext_k, -1, R_MOD)
        self.barycentric_weight = self.ifft_divisor

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def _zeta(self, a: List[int], into_coset: bool) -> List[int]:
        cp = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        return [x if i % 3 == 0 else x * cp[i % 3 - 1] % R_MOD for i, x in enumerate(a)]

    def lagrange_to_coeff(self, a: List[int]) -> List[int]:
        assert len(a) == 1 << self.k  # domain.rs:227
        out = best_fft(a, self.omega_inv, self.k)
        return [x * self.ifft_divisor % R_MOD for x in out]

    def coeff_to_lagrange(self, a: List[int]) -> List[int]:
        assert len(a) == 1 << self.k
        return best_fft(a, self.omega, self.k)

    def coeff_to_extended(self, a: List[int]) -> List[int]:
        assert len(a) == 1 << self.k  # domain.rs:244
        b = self._zeta(list(a), True)
        b += [0] * (self.extended_len() - len(b))
        return best_fft(b, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a: List[int]) -> List[int]:
        assert len(a) == self.extended_len()  # domain.rs:311
        b = best_fft(a, self.extended_omega_inv, self.extended_k)
        b = [x * self.extended_ifft_divisor % R_MOD for x in b]
        b = self._zeta(b, False)
        return b[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a: List[int]) -> List[int]:
        assert len(a) == self.extended_len()
        m = len(self.t_evaluations)
        return [x * self.t_evaluations[i % m] % R_MOD for i, x in enumerate(a)]


# --- deterministic synthetic inputs ------------------------------------------
MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)


def random_fr(n: int, seed: int) -> List[int]:
    """Uniform in [0, r): 4 limbs from splitmix64, top limb masked to 254 bits, reject >= r."""
    g = SplitMix64(seed)
    out = []
    while len(out) < n:
        l = [g.next() for _ in range(4)]
        l[3] &= (1 << 62) - 1
        v = limbs_to_int(l)
        if v < R_MOD:
            out.append(v)
    return out


def sqrt_fq(a: int) -> Optional[int]:
    y = pow(a, (Q_MOD + 1) // 4, Q_MOD)  # q = 3 mod 4
    return y if y * y % Q_MOD == a % Q_MOD else None


def random_g1(n: int, seed: int) -> List[Affine]:
    """Deterministic try-and-increment; cofactor 1 so every curve point is in G1."""
    g = SplitMix64(seed)
    out: List[Affine] = []
    while len(out) < n:
        l = [g.next() for _ in range(4)]
        l[3] &= (1 << 62) - 1
        x = limbs_to_int(l) % Q_MOD
        sign = g.next() & 1
        while True:
            y = sqrt_fq((x * x * x + CURVE_B) % Q_MOD)
            if y is not None:
                break
            x = (x + 1) % Q_MOD
        if (y & 1) != sign:
            y = Q_MOD - y
        out.append((x, y))
    return out
