"""ctypes binding of oracle/libh2ref.so (the C restatement of the reference CPU path).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/h2ref.c.  Arrays are numpy
uint64 in the FFI layout: Fr/Fq = 4 limbs little-endian Montgomery; G1Affine = 8
limbs (x,y), identity (0,0); G1 = 12 limbs homogeneous projective (x = X/Z, y = Y/Z), identity z = 0.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libh2ref.so")
    src = os.path.join(_HERE, "h2ref.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        # -march=native is avoided: the .so is built here and travels to the GPU box.
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "CFLAGS=-O3 -fPIC -Wall -Wextra -std=gnu11"])
    return so


class Domain(C.Structure):
    _fields_ = [
        ("k", C.c_uint32), ("extended_k", C.c_uint32), ("j", C.c_uint32), ("n_t", C.c_uint32),
        ("omega", C.c_uint64 * 4), ("omega_inv", C.c_uint64 * 4),
        ("extended_omega", C.c_uint64 * 4), ("extended_omega_inv", C.c_uint64 * 4),
        ("g_coset", C.c_uint64 * 4), ("g_coset_inv", C.c_uint64 * 4),
        ("ifft_divisor", C.c_uint64 * 4), ("extended_ifft_divisor", C.c_uint64 * 4),
        ("t_evaluations", C.c_uint64 * 32),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.h2ref_domain_new.restype = C.c_int
    return _LIB


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def default_threads() -> int:
    return os.cpu_count() or 1


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray, threads: int | None = None) -> np.ndarray:
    n = coeffs.shape[0]
    assert bases.shape[0] == n
    out = np.zeros(12, dtype=np.uint64)
    lib().h2ref_best_multiexp(_p(coeffs), _p(bases), C.c_size_t(n), C.c_int(threads or default_threads()), _p(out))
    return out


def g1_to_affine(jac: np.ndarray) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().h2ref_g1_to_affine(_p(np.ascontiguousarray(jac)), _p(out))
    return out


def g1_add(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().h2ref_g1_add(_p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out


def g1_scalar_mul(p_affine: np.ndarray, k_mont: np.ndarray) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().h2ref_g1_scalar_mul(_p(np.ascontiguousarray(p_affine)), _p(np.ascontiguousarray(k_mont)), _p(out))
    return out


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int | None = None) -> np.ndarray:
    """Returns a transformed copy (the C function works in place)."""
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    assert a.shape[0] == 1 << log_n
    lib().h2ref_best_fft(_p(a), _p(np.ascontiguousarray(omega)), C.c_uint32(log_n), C.c_int(threads or default_threads()))
    return a


def domain_new(j: int, k: int) -> Domain:
    d = Domain()
    rc = lib().h2ref_domain_new(C.c_uint32(j), C.c_uint32(k), C.byref(d))
    if rc != 0:
        raise ValueError(f"h2ref_domain_new({j},{k}) -> {rc}")
    return d


def lagrange_to_coeff(d: Domain, a: np.ndarray, threads: int | None = None) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    assert a.shape[0] == 1 << d.k
    lib().h2ref_lagrange_to_coeff(C.byref(d), _p(a), C.c_int(threads or default_threads()))
    return a


def coeff_to_extended(d: Domain, a: np.ndarray, threads: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[0] == 1 << d.k
    out = np.zeros((1 << d.extended_k, 4), dtype=np.uint64)
    lib().h2ref_coeff_to_extended(C.byref(d), _p(a), _p(out), C.c_int(threads or default_threads()))
    return out


def extended_to_coeff(d: Domain, a: np.ndarray, threads: int | None = None) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    assert a.shape[0] == 1 << d.extended_k
    out = np.zeros(((d.j - 1) << d.k, 4), dtype=np.uint64)
    lib().h2ref_extended_to_coeff(C.byref(d), _p(a), _p(out), C.c_int(threads or default_threads()))
    return out


def divide_by_vanishing_poly(d: Domain, a: np.ndarray, threads: int | None = None) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    lib().h2ref_divide_by_vanishing_poly(C.byref(d), _p(a), C.c_int(threads or default_threads()))
    return a


def random_fr(n: int, seed: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().h2ref_random_fr(_p(out), C.c_size_t(n), C.c_uint64(seed))
    return out


def random_g1(n: int, seed: int) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().h2ref_random_g1(_p(out), C.c_size_t(n), C.c_uint64(seed))
    return out


def progression_g1(n: int, seed: int, threads: int | None = None) -> np.ndarray:
    """n distinct points B + i*G (synthetic MSM bases at benchmark sizes; see h2ref_progression_g1)."""
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().h2ref_progression_g1(_p(out), C.c_size_t(n), C.c_uint64(seed), C.c_int(threads or default_threads()))
    return out


def fr_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros_like(a)
    lib().h2ref_fr_mul(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def fq_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.zeros_like(a)
    lib().h2ref_fq_mul(_p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def to_mont(a: np.ndarray, fq: bool = False) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    lib().h2ref_to_mont(_p(a), C.c_size_t(a.size // 4), C.c_int(1 if fq else 0))
    return a


def from_mont(a: np.ndarray, fq: bool = False) -> np.ndarray:
    a = np.array(a, dtype=np.uint64, copy=True, order="C")
    lib().h2ref_from_mont(_p(a), C.c_size_t(a.size // 4), C.c_int(1 if fq else 0))
    return a
