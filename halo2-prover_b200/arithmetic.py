"""Host mirror of halo2_proofs::arithmetic for the hot path (same names and argument meaning).

best_multiexp  <- halo2_proofs @6b43b6b src/arithmetic.rs:147-180
best_fft       <- src/arithmetic.rs:185-250
Upstream panics via assert_eq! on bad lengths; here that is an AssertionError raised
before the FFI call, and any non-zero status from the library raises H2BError.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray) -> np.ndarray:
    """sum_i coeffs[i] * bases[i] -> G1 (12 x u64, homogeneous projective x = X/Z, y = Y/Z).  coeffs (n,4), bases (n,8)."""
    coeffs = _ffi.as_u64(coeffs, 4)
    bases = _ffi.as_u64(bases, 8)
    assert coeffs.shape[0] == bases.shape[0], "assert_eq!(coeffs.len(), bases.len())"  # arithmetic.rs:148
    _ffi.init()
    out = np.zeros(12, dtype=np.uint64)
    _ffi.check(_ffi.lib().h2b_best_multiexp(_ffi.u64p(coeffs), _ffi.u64p(bases), C.c_size_t(coeffs.shape[0]),
                                            _ffi.u64p(out)))
    return out


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int) -> None:
    """In-place forward transform of ``a`` ((2^log_n, 4) uint64), natural order in and out."""
    if a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("best_fft works in place: pass a C-contiguous uint64 array")
    assert a.size == 4 << log_n, "assert_eq!(a.len(), 1 << log_n)"  # arithmetic.rs:199
    omega = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
    _ffi.init()
    _ffi.check(_ffi.lib().h2b_best_fft(_ffi.u64p(a), _ffi.u64p(omega), C.c_uint32(log_n)))


def g1_fold(points: np.ndarray) -> np.ndarray:
    """Sum of projective points ((m,12) uint64) -- the fold of per-chunk partial results."""
    points = _ffi.as_u64(points, 12)
    _ffi.init()
    out = np.zeros(12, dtype=np.uint64)
    _ffi.check(_ffi.lib().h2b_g1_fold(_ffi.u64p(points), C.c_size_t(points.shape[0]), _ffi.u64p(out)))
    return out


def g_to_lagrange(g: np.ndarray, k: int) -> np.ndarray:
    """arithmetic::g_to_lagrange: the Lagrange-basis SRS ((2^k, 8) uint64 affine) from the monomial one."""
    g = _ffi.as_u64(g, 8)
    assert g.shape[0] == 1 << k, "assert_eq!(g.len(), 1 << k)"
    _ffi.init()
    out = np.zeros_like(g)
    _ffi.check(_ffi.lib().h2b_g_to_lagrange(_ffi.u64p(g), C.c_uint32(k), _ffi.u64p(out)))
    return out


def g1_to_bytes(points: np.ndarray) -> bytes:
    """G1Affine::to_bytes of each projective point ((m,12) uint64): the 32-byte transcript encoding."""
    points = _ffi.as_u64(points, 12)
    _ffi.init()
    out = np.zeros(points.shape[0] * 32, dtype=np.uint8)
    _ffi.check(_ffi.lib().h2b_g1_to_bytes(_ffi.u64p(points), C.c_size_t(points.shape[0]),
                                          out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out.tobytes()


# ---- device-resident variants (torch tensors as raw 64-bit containers) ------------------------------
def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _stream_ptr(stream) -> C.c_void_p:
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    # torch's default stream is handle 0, which this ABI reads as "the library's own stream": name the
    # legacy default stream explicitly (cudaStreamLegacy) so the caller's synchronize() covers the work
    return C.c_void_p(stream.cuda_stream or 1)


def dev_msm(coeffs_t, bases_t, out_t, n: int | None = None, stream=None) -> None:
    """coeffs_t: (n,4) int64 cuda tensor, bases_t: (>=n,8), out_t: (12,) -- all on the library's device."""
    if n is None:
        n = coeffs_t.shape[0]
    _ffi.init(coeffs_t.device.index)
    _ffi.check(_ffi.lib().h2b_dev_msm(_ptr(coeffs_t), _ptr(bases_t), C.c_size_t(n), _ptr(out_t), _stream_ptr(stream)))


def dev_best_fft(a_t, omega: np.ndarray, log_n: int, stream=None) -> None:
    omega = np.ascontiguousarray(omega, dtype=np.uint64).reshape(4)
    assert a_t.numel() == 4 << log_n
    _ffi.init(a_t.device.index)
    _ffi.check(_ffi.lib().h2b_dev_best_fft(_ptr(a_t), _ffi.u64p(omega), C.c_uint32(log_n), _stream_ptr(stream)))


def dev_g1_fold(points_t, out_t, stream=None) -> None:
    _ffi.init(points_t.device.index)
    _ffi.check(_ffi.lib().h2b_dev_g1_fold(_ptr(points_t), C.c_size_t(points_t.numel() // 12), _ptr(out_t),
                                          _stream_ptr(stream)))
