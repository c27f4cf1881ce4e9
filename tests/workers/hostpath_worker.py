"""Subprocess body of tests/test_hostpath_gpu.py: pageable-buffer transforms and commits under the copy policy the
environment selects (read once by h2b_init), compared with the oracle.  Prints OK <n checks>."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import bn254 as spec  # noqa: E402
import h2ref as href  # noqa: E402
import halo2_prover_b200 as h2b  # noqa: E402


def main():
    from halo2_prover_b200 import _ffi
    _ffi.init(0)
    checks = 0
    # sizes on both sides of every threshold: 128 KiB (driver), 512 KiB (one ring chunk), 2 / 8 / 32 MiB (several chunks,
    # ring growth 4 -> 8 -> 32 MiB, workers from 8 MiB by default)
    for k in (12, 14, 16, 18, 20, 13):
        a = href.random_fr(1 << k, 900 + k)
        om = spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - k), spec.R_MOD)])[0]
        want = href.best_fft(a, om, k)
        got = a.copy()
        h2b.best_fft(got, om, k)
        assert (got == want).all(), f"best_fft 2^{k}"
        checks += 1
    k = 18
    g = np.tile(href.random_g1(1 << 10, 5), (1 << (k - 10), 1))
    params = h2b.ParamsKZG(k, g)
    for n in (1 << 14, 1 << 18):      # 512 KiB and 8 MiB of scalars
        poly = href.random_fr(n, 77 + n)
        want = href.g1_to_affine(np.ascontiguousarray(href.best_multiexp(poly, g[:n])))
        got = href.g1_to_affine(np.ascontiguousarray(params.commit(poly)))
        assert (got == want).all(), f"commit {n}"
        checks += 1
    print("OK", checks)


if __name__ == "__main__":
    main()
