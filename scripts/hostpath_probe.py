"""Host-buffer (pageable) entry points at the sizes of a k = 14 proof: ms per call.  H2B_MIN_STAGED_KB sets from which
total size the copies go through the library's pinned ring instead of the driver's own staging."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi  # noqa: E402
import bn254 as o  # noqa: E402

_ffi.init(0)
rng = np.random.default_rng(3)


def rnd(n):
    a = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 4), dtype=np.uint64, endpoint=True)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


for k in (12, 14, 16, 17, 18, 20):
    n = 1 << k
    om = o.fr_array([pow(o.ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD)])[0]
    a = rnd(n)
    for _ in range(3):
        h2b.best_fft(a, om, k)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        h2b.best_fft(a, om, k)
    print(f"best_fft host 2^{k} ({n * 32 >> 10} KiB each way): {(time.perf_counter() - t0) / reps * 1e3:.3f} ms", flush=True)
d = h2b.EvaluationDomain(6, 14)
cols = [rnd(1 << 14) for _ in range(7)]
outs = [np.zeros((1 << 17, 4), dtype=np.uint64) for _ in range(7)]
for _ in range(2):
    d.coeff_to_extended_many(cols, outs)
t0 = time.perf_counter()
for _ in range(10):
    d.coeff_to_extended_many(cols, outs)
print(f"coeff_to_extended_many 7 x 14->17: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms", flush=True)
t0 = time.perf_counter()
for _ in range(10):
    for q in range(7):
        d.coeff_to_extended(cols[q], outs[q])
print(f"coeff_to_extended x 7 14->17: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms", flush=True)

import h2ref  # noqa: E402
for k in (13, 14, 16):
    n = 1 << k
    g = h2ref.random_g1(1 << 10, 5)
    g = np.ascontiguousarray(np.tile(g, (n >> 10, 1)))
    params = h2b.ParamsKZG(k, g)
    cs = [rnd(n) for _ in range(8)]
    for _ in range(3):
        params.commit(cs[0]); params.commit_many(cs)
    t0 = time.perf_counter()
    for _ in range(20):
        for c in cs:
            params.commit(c)
    t1 = time.perf_counter()
    for _ in range(20):
        params.commit_many(cs)
    t2 = time.perf_counter()
    print(f"k={k}: 8 commits {(t1 - t0) / 20 * 1e3:.3f} ms, commit_many(8) {(t2 - t1) / 20 * 1e3:.3f} ms", flush=True)
    params.release()
