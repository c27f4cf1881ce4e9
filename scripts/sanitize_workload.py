"""Small hot-path workload for compute-sanitizer: MSM at 2^12 (bucket-free table) and 2^16
(window table and generic path), NTT / coset transforms at k = 12 and 16, each checked against the oracle.

    compute-sanitizer --tool memcheck  python scripts/sanitize_workload.py
    compute-sanitizer --tool racecheck python scripts/sanitize_workload.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import h2ref  # noqa: E402
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi  # noqa: E402

_ffi.init(0)
for k in (12, 16):
    n = 1 << k
    bases, scalars = h2ref.random_g1(n, 1), h2ref.random_fr(n, 2)
    want = h2ref.g1_to_affine(h2ref.best_multiexp(scalars, bases))
    params = h2b.ParamsKZG(k, bases)
    assert (h2ref.g1_to_affine(params.commit(scalars)) == want).all()
    cols = [h2ref.random_fr(n, 10 + q) for q in range(3)]
    many = params.commit_many(cols)
    assert (h2ref.g1_to_affine(many[1]) == h2ref.g1_to_affine(h2ref.best_multiexp(cols[1], bases))).all()
    params.release()
    assert (h2ref.g1_to_affine(h2b.best_multiexp(scalars, bases)) == want).all()
    d, dc = h2b.EvaluationDomain(4, k), h2ref.domain_new(4, k)
    ext = d.coeff_to_extended(scalars)
    assert (ext == h2ref.coeff_to_extended(dc, scalars)).all()
    assert (d.extended_to_coeff(ext) == h2ref.extended_to_coeff(dc, ext)).all()
    assert (d.lagrange_to_coeff(scalars.copy()) == h2ref.lagrange_to_coeff(dc, scalars)).all()
    print("ok k =", k, flush=True)
_ffi.shutdown()
print("sanitize workload done")
