"""Host mirror of halo2_proofs::poly::EvaluationDomain<Fr> (src/poly/domain.rs @6b43b6b).

Same constructor arguments and method names as upstream; polynomials are (len,4)
uint64 arrays (Montgomery Fr).  All arithmetic, including the derivation of the
domain constants, runs on the GPU through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def _fr(limbs) -> np.ndarray:
    return np.array(list(limbs), dtype=np.uint64)


class EvaluationDomain:
    def __init__(self, j: int, k: int):
        """EvaluationDomain::new(j, k) (domain.rs:~40-140)."""
        _ffi.init()
        self._d = _ffi.Domain()
        _ffi.check(_ffi.lib().h2b_domain_new(C.c_uint32(j), C.c_uint32(k), C.byref(self._d)))
        self.k = self._d.k
        self.extended_k = self._d.extended_k
        self.j = j
        self.n = 1 << k
        self.quotient_poly_degree = j - 1

    # constants, as (4,) uint64 Montgomery limbs
    def get_omega(self): return _fr(self._d.omega)
    def get_omega_inv(self): return _fr(self._d.omega_inv)
    def get_extended_omega(self): return _fr(self._d.extended_omega)
    @property
    def extended_omega_inv(self): return _fr(self._d.extended_omega_inv)
    @property
    def g_coset(self): return _fr(self._d.g_coset)
    @property
    def g_coset_inv(self): return _fr(self._d.g_coset_inv)
    @property
    def ifft_divisor(self): return _fr(self._d.ifft_divisor)
    @property
    def extended_ifft_divisor(self): return _fr(self._d.extended_ifft_divisor)
    @property
    def t_evaluations(self): return _fr(self._d.t_evaluations)[: 4 * self._d.n_t].reshape(-1, 4)

    def extended_len(self) -> int:
        return 1 << self.extended_k

    def lagrange_to_coeff(self, a: np.ndarray) -> np.ndarray:
        """domain.rs:227 -- consumes ``a`` (transformed in place) and returns it."""
        a = _ffi.as_u64(a, 4)
        assert a.shape[0] == 1 << self.k, "assert_eq!(a.values.len(), 1 << self.k)"
        _ffi.check(_ffi.lib().h2b_lagrange_to_coeff(C.byref(self._d), _ffi.u64p(a)))
        return a

    def coeff_to_extended(self, a: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """domain.rs:244 -- 2^k coefficients -> 2^extended_k evaluations on the zeta-coset
        (``out``: optional preallocated (2^extended_k, 4) uint64 result buffer)."""
        a = _ffi.as_u64(a, 4)
        assert a.shape[0] == 1 << self.k, "assert_eq!(a.values.len(), 1 << self.k)"
        if out is None:
            out = np.empty((self.extended_len(), 4), dtype=np.uint64)
        assert out.shape == (self.extended_len(), 4)
        _ffi.check(_ffi.lib().h2b_coeff_to_extended(C.byref(self._d), _ffi.u64p(a), _ffi.u64p(out)))
        return out

    def lagrange_to_coeff_many(self, cols) -> list:
        """[lagrange_to_coeff(a) for a in cols] in one call (in place)."""
        cols = [_ffi.as_u64(a, 4) for a in cols]
        for a in cols:
            assert a.shape[0] == 1 << self.k, "assert_eq!(a.values.len(), 1 << self.k)"
        m = len(cols)
        if m:
            ptrs = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(a) for a in cols])
            _ffi.check(_ffi.lib().h2b_lagrange_to_coeff_many(C.byref(self._d), ptrs, C.c_size_t(m)))
        return cols

    def coeff_to_extended_many(self, cols, outs: list | None = None) -> list:
        """[coeff_to_extended(a) for a in cols] in one call (``outs``: optional preallocated results)."""
        cols = [_ffi.as_u64(a, 4) for a in cols]
        for a in cols:
            assert a.shape[0] == 1 << self.k, "assert_eq!(a.values.len(), 1 << self.k)"
        m = len(cols)
        if outs is None:
            outs = [np.empty((self.extended_len(), 4), dtype=np.uint64) for _ in range(m)]
        assert len(outs) == m and all(o.shape == (self.extended_len(), 4) for o in outs)
        if m:
            pin = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(a) for a in cols])
            pout = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(a) for a in outs])
            _ffi.check(_ffi.lib().h2b_coeff_to_extended_many(C.byref(self._d), pin, pout, C.c_size_t(m)))
        return outs

    def extended_to_coeff(self, a: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """domain.rs:311 -- 2^extended_k coset evaluations -> n * (j-1) coefficients."""
        a = _ffi.as_u64(a, 4)
        assert a.shape[0] == self.extended_len(), "assert_eq!(a.values.len(), self.extended_len())"
        if out is None:
            out = np.empty((self.n * self.quotient_poly_degree, 4), dtype=np.uint64)
        assert out.shape == (self.n * self.quotient_poly_degree, 4)
        _ffi.check(_ffi.lib().h2b_extended_to_coeff(C.byref(self._d), _ffi.u64p(a), _ffi.u64p(out)))
        return out

    def divide_by_vanishing_poly(self, a: np.ndarray) -> np.ndarray:
        a = _ffi.as_u64(a, 4)
        assert a.shape[0] == self.extended_len()
        _ffi.check(_ffi.lib().h2b_divide_by_vanishing_poly(C.byref(self._d), _ffi.u64p(a)))
        return a

    # device-resident (torch tensors)
    def dev_lagrange_to_coeff(self, a_t, stream=None) -> None:
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_lagrange_to_coeff(C.byref(self._d), _ptr(a_t), _stream_ptr(stream)))

    def dev_coeff_to_extended(self, in_t, out_t, stream=None) -> None:
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_coeff_to_extended(C.byref(self._d), _ptr(in_t), _ptr(out_t), _stream_ptr(stream)))

    def dev_divide_by_vanishing_poly(self, a_t, stream=None) -> None:
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_divide_by_vanishing_poly(C.byref(self._d), _ptr(a_t), _stream_ptr(stream)))

    def dev_lagrange_to_coeff_many(self, a_t, m: int, stream=None) -> None:
        """m columns one after the other in ``a_t`` ((m * 2^k, 4)), in place."""
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_lagrange_to_coeff_many(C.byref(self._d), _ptr(a_t), C.c_size_t(m), _stream_ptr(stream)))

    def dev_coeff_to_extended_many(self, in_t, out_t, m: int, stream=None) -> None:
        """m columns: in_t (m * 2^k, 4) -> out_t (m * 2^extended_k, 4)."""
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_coeff_to_extended_many(C.byref(self._d), _ptr(in_t), _ptr(out_t), C.c_size_t(m),
                                                            _stream_ptr(stream)))

    def dev_extended_to_coeff(self, in_t, out_t, stream=None) -> None:
        from .arithmetic import _ptr, _stream_ptr
        _ffi.check(_ffi.lib().h2b_dev_extended_to_coeff(C.byref(self._d), _ptr(in_t), _ptr(out_t), _stream_ptr(stream)))
