"""Host mirror of the caller-side boundary: ParamsKZG::{commit, commit_lagrange}
(halo2_proofs @6b43b6b src/poly/kzg/commitment.rs:319, :363).

Upstream copies the polynomial into a Vec and calls
``best_multiexp(&scalars, &self.g[0..n])`` (resp. ``g_lagrange``); the blind is ignored
for KZG.  Here the two base arrays are registered once (device-resident SRS) and each
commit moves only the 32*n bytes of scalars.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


class ParamsKZG:
    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray | None = None):
        """``g`` / ``g_lagrange``: (2^k, 8) uint64 G1Affine arrays (the body of ParamsKZG::write)."""
        self.k = k
        self.n = 1 << k
        _ffi.init()
        self._handles = {}
        for name, arr in (("g", g), ("g_lagrange", g_lagrange)):
            if arr is None:
                continue
            arr = _ffi.as_u64(arr, 8)
            assert arr.shape[0] >= self.n
            h = C.c_uint64(0)
            _ffi.check(_ffi.lib().h2b_srs_register(_ffi.u64p(arr), C.c_size_t(arr.shape[0]), C.byref(h)))
            self._handles[name] = h.value

    @classmethod
    def from_device(cls, k: int, g_t, g_lagrange_t=None) -> "ParamsKZG":
        """Bases already in HBM ((>= 2^k, 8) int64 cuda tensors): h2b_dev_srs_register."""
        _ffi.init(g_t.device.index)
        self = cls.__new__(cls)
        self.k = k
        self.n = 1 << k
        self._handles = {}
        for name, t in (("g", g_t), ("g_lagrange", g_lagrange_t)):
            if t is None:
                continue
            assert t.shape[0] >= self.n
            h = C.c_uint64(0)
            _ffi.check(_ffi.lib().h2b_dev_srs_register(C.c_void_p(t.data_ptr()), C.c_size_t(t.shape[0]), C.byref(h)))
            self._handles[name] = h.value
        return self

    @classmethod
    def setup(cls, k: int, s: int) -> "ParamsKZG":
        """ParamsKZG::setup(k, rng) (kzg/commitment.rs:68-114; the reference's generate_params, utils.rs:59-61) with the
        secret scalar given instead of drawn: g[i] = [s^i] G by the per-element fixed-base multiplication on the device
        (h2b_dev_fixed_base_mul), g_lagrange = g_to_lagrange(g) (h2b_g_to_lagrange), both registered.  The G2 half of the
        parameters (g2, s_g2) is not on this path; write() takes those 256 bytes from the caller."""
        import torch
        n = 1 << k
        r = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001   # Fr modulus
        q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47   # Fq modulus
        limbs = lambda v: [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
        powers, acc = np.zeros((n, 4), dtype=np.uint64), 1
        for i in range(n):                      # Montgomery form: s^i * 2^256 mod r
            powers[i] = limbs(acc * (1 << 256) % r)
            acc = acc * s % r
        gen = np.array(limbs((1 << 256) % q) + limbs(2 * (1 << 256) % q), dtype=np.uint64)   # G = (1, 2)
        _ffi.init()
        d_s = torch.from_numpy(powers.view(np.int64)).cuda()
        d_g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream()
        _ffi.check(_ffi.lib().h2b_dev_fixed_base_mul(C.c_void_p(d_s.data_ptr()), C.c_size_t(n), _ffi.u64p(gen),
                                                     C.c_void_p(d_g.data_ptr()), C.c_void_p(st.cuda_stream or 1)))
        st.synchronize()
        g = np.ascontiguousarray(d_g.cpu().numpy().view(np.uint64))
        from .arithmetic import g_to_lagrange
        return cls(k, g, g_to_lagrange(g, k))

    @classmethod
    def read(cls, data) -> "ParamsKZG":
        """ParamsKZG::read (SerdeFormat::RawBytes): ``data`` is the byte string ParamsKZG::write produced
        (k | g | g_lagrange | g2 | s_g2); both base arrays are registered straight from it."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        _ffi.init()
        k, hg, hl = C.c_uint32(), C.c_uint64(), C.c_uint64()
        _ffi.check(_ffi.lib().h2b_params_read(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(buf.size),
                                              C.byref(k), C.byref(hg), C.byref(hl)))
        self = cls.__new__(cls)
        self.k = k.value
        self.n = 1 << k.value
        self._handles = {"g": hg.value, "g_lagrange": hl.value}
        self.g2_and_s_g2 = bytes(data)[-256:]
        return self

    def write(self, g2_and_s_g2: bytes | None = None) -> bytes:
        """ParamsKZG::write (SerdeFormat::RawBytes): k | g | g_lagrange read back from HBM | g2 | s_g2.  The 256 bytes
        of G2 points are the caller's (kept from read(); G2 arithmetic is not on this path)."""
        tail = g2_and_s_g2 if g2_and_s_g2 is not None else getattr(self, "g2_and_s_g2", None)
        assert tail is not None and len(tail) == 256, "write: 256 bytes of g2 | s_g2 are needed"
        assert "g" in self._handles and "g_lagrange" in self._handles
        out = np.zeros(4 + 128 * self.n + 256, dtype=np.uint8)
        t = np.frombuffer(tail, dtype=np.uint8)
        _ffi.check(_ffi.lib().h2b_params_write(C.c_uint32(self.k), C.c_uint64(self._handles["g"]),
                                               C.c_uint64(self._handles["g_lagrange"]), t.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size)))
        return out.tobytes()

    def _commit(self, which: str, poly: np.ndarray) -> np.ndarray:
        poly = _ffi.as_u64(poly, 4)
        size = poly.shape[0]
        assert self.n >= size, "assert!(bases.len() >= size)"  # commitment.rs:319 / :363
        out = np.zeros(12, dtype=np.uint64)
        _ffi.check(_ffi.lib().h2b_commit(C.c_uint64(self._handles[which]), _ffi.u64p(poly), C.c_size_t(size),
                                         _ffi.u64p(out)))
        return out

    def commit(self, poly: np.ndarray, _blind=None) -> np.ndarray:
        return self._commit("g", poly)

    def commit_lagrange(self, poly: np.ndarray, _blind=None) -> np.ndarray:
        return self._commit("g_lagrange", poly)

    def _commit_many(self, which: str, polys) -> np.ndarray:
        polys = [_ffi.as_u64(p, 4) for p in polys]
        m = len(polys)
        out = np.zeros((m, 12), dtype=np.uint64)
        if m == 0:
            return out
        size = polys[0].shape[0]
        assert all(p.shape[0] == size for p in polys), "commit_many: columns of one batch have one length"
        assert self.n >= size, "assert!(bases.len() >= size)"  # commitment.rs:319 / :363
        ptrs = (C.POINTER(C.c_uint64) * m)(*[_ffi.u64p(p) for p in polys])
        _ffi.check(_ffi.lib().h2b_commit_many(C.c_uint64(self._handles[which]), ptrs, C.c_size_t(size), C.c_size_t(m),
                                              _ffi.u64p(out)))
        return out

    def commit_many(self, polys) -> np.ndarray:
        """[commit(p) for p in polys] in one pass; (m, 12) uint64."""
        return self._commit_many("g", polys)

    def commit_lagrange_many(self, polys) -> np.ndarray:
        return self._commit_many("g_lagrange", polys)

    def dev_commit(self, coeffs_t, out_t, which: str = "g", stream=None) -> None:
        """Device-resident commit: coeffs_t (n,4) int64 cuda tensor, out_t (12,) -- h2b_dev_commit."""
        from .arithmetic import _ptr, _stream_ptr
        size = coeffs_t.shape[0]
        assert self.n >= size, "assert!(bases.len() >= size)"  # commitment.rs:319 / :363
        _ffi.check(_ffi.lib().h2b_dev_commit(C.c_uint64(self._handles[which]), _ptr(coeffs_t), C.c_size_t(size),
                                             _ptr(out_t), _stream_ptr(stream)))

    def device_bases(self, which: str = "g"):
        """(device pointer, length) of a registered base array."""
        p = C.c_void_p()
        n = C.c_size_t()
        _ffi.check(_ffi.lib().h2b_srs_device_ptr(C.c_uint64(self._handles[which]), C.byref(p), C.byref(n)))
        return p.value, n.value

    def release(self) -> None:
        for h in self._handles.values():
            _ffi.lib().h2b_srs_release(C.c_uint64(h))
        self._handles = {}
