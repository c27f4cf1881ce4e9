"""h2b_shutdown followed by h2b_init gives a working library again (runs last: it invalidates every handle).

Kernels that need more than 48 KiB of dynamic shared memory (the NTT passes, the bucket-reduction tree) set a
per-function attribute that belongs to the context; it must be set again for the new one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _affine(href, p):
    return href.g1_to_affine(np.ascontiguousarray(p))


def test_reinit_after_shutdown(h2b, spec, href):
    from halo2_prover_b200 import _ffi
    k = 16
    n = 1 << k
    g = np.tile(href.random_g1(1 << 12, 7), (n >> 12, 1))
    poly = href.random_fr(n, 8)
    om = spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]

    def run():
        params = h2b.ParamsKZG(k, g)
        c = _affine(href, params.commit(poly))          # window table + bucket-reduction tree (64 KiB shared memory)
        m = _affine(href, h2b.best_multiexp(poly, g))   # generic path
        a = poly.copy()
        h2b.best_fft(a, om, k)                          # two radix-2^8 passes (> 48 KiB shared memory)
        params.release()
        return c, m, a

    c0, m0, a0 = run()
    assert (c0 == m0).all()
    assert (a0 == href.best_fft(poly.copy(), om, k)).all()
    _ffi.shutdown()
    _ffi.init(0)
    c1, m1, a1 = run()
    assert (c1 == c0).all() and (m1 == m0).all() and (a1 == a0).all()
