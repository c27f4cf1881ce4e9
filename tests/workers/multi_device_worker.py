"""Run by tests/test_multi_device_gpu.py in a fresh process: ONE process, several GPUs through the C ABI
(h2b_init_devices), every result compared with the oracle.  H2B_SHARD_MIN_LOG (environment) lowers the size from
which a registered SRS is sharded by point range so that both layouts are exercised at test sizes."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import h2ref  # noqa: E402
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi  # noqa: E402

ndev = int(sys.argv[1])
_ffi.init_devices(list(range(ndev)))
L = _ffi.lib()
assert L.h2b_device_count() == ndev
aff = h2ref.g1_to_affine


def layout(params):
    parts, repl = C.c_uint32(), C.c_uint32()
    sizes = (C.c_size_t * 64)()
    _ffi.check(L.h2b_srs_layout(C.c_uint64(params._handles["g"]), C.byref(parts), C.byref(repl), sizes))
    return parts.value, bool(repl.value), list(sizes[: parts.value])


# ---- sharded SRS (n >= 2^H2B_SHARD_MIN_LOG): every commit is split by point range and folded on the primary device
k = 15
n = 1 << k
bases, scalars = h2ref.random_g1(n, 1), h2ref.random_fr(n, 2)
params = h2b.ParamsKZG(k, bases)
parts, repl, sizes = layout(params)
assert parts == ndev and not repl and sum(sizes) == n, (parts, repl, sizes)
want = aff(h2ref.best_multiexp(scalars, bases))
assert (aff(params.commit(scalars)) == want).all(), "sharded commit"
for m in (1, 5, n // ndev, n // ndev + 1, n - 1):  # shorter polynomials use only the shares they reach
    sc = np.ascontiguousarray(scalars[:m])
    assert (aff(params.commit(sc)) == aff(h2ref.best_multiexp(sc, np.ascontiguousarray(bases[:m])))).all(), m
cols = [h2ref.random_fr(n, 10 + q) for q in range(5)]
cols[2][::3] = 0
many = params.commit_many(cols)
for q in range(5):
    assert (aff(many[q]) == aff(h2ref.best_multiexp(cols[q], bases))).all(), ("sharded commit_many", q)
# scalars resident on the primary device: the other devices pull their slices device-to-device
import torch  # noqa: E402
torch.cuda.set_device(0)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ds = torch.from_numpy(np.concatenate(cols).view(np.int64)).cuda()
    out = torch.empty((5, 12), dtype=torch.int64, device="cuda")
    sp = C.c_void_p(s.cuda_stream)
    _ffi.check(L.h2b_dev_commit_many(C.c_uint64(params._handles["g"]), C.c_void_p(ds.data_ptr()), C.c_size_t(n), C.c_size_t(5),
                                     C.c_void_p(out.data_ptr()), sp))
    one = torch.empty(12, dtype=torch.int64, device="cuda")
    _ffi.check(L.h2b_dev_commit(C.c_uint64(params._handles["g"]), C.c_void_p(ds[3 * n:].data_ptr()), C.c_size_t(n),
                                C.c_void_p(one.data_ptr()), sp))
    s.synchronize()
o = out.cpu().numpy().view(np.uint64)
for q in range(5):
    assert (aff(o[q]) == aff(many[q])).all(), ("dev_commit_many", q)
assert (aff(one.cpu().numpy().view(np.uint64)) == aff(many[3])).all(), "dev_commit"
params.release()
# ParamsKZG::write gathers the shares of both arrays back into the reference's byte layout
lag_bases = h2ref.random_g1(n, 4)
params = h2b.ParamsKZG(k, bases, lag_bases)
tail = bytes(range(256))
blob = params.write(tail)
assert blob == k.to_bytes(4, "little") + bases.tobytes() + lag_bases.tobytes() + tail, "sharded params write"
params.release()

# ---- replicated SRS (small): whole columns are dealt to the devices
k = 10
n = 1 << k
bases = h2ref.random_g1(n, 3)
params = h2b.ParamsKZG(k, bases)
parts, repl, sizes = layout(params)
assert parts == ndev and repl and all(sz == n for sz in sizes), (parts, repl, sizes)
for m in (1, 2, ndev, 2 * ndev + 1):
    cols = [h2ref.random_fr(n, 100 + 7 * m + q) for q in range(m)]
    many = params.commit_many(cols)
    for q in range(m):
        assert (aff(many[q]) == aff(h2ref.best_multiexp(cols[q], bases))).all(), ("dealt commit_many", m, q)
assert (aff(params.commit(cols[0])) == aff(h2ref.best_multiexp(cols[0], bases))).all()
params.release()

# ---- independent column transforms are dealt to the devices (a single NTT stays on one device)
for k in (9, 13):
    d, dc = h2b.EvaluationDomain(4, k), h2ref.domain_new(4, k)
    for m in (1, 3, 2 * ndev + 1):
        cols = [h2ref.random_fr(1 << k, 500 + 11 * m + q) for q in range(m)]
        ext = d.coeff_to_extended_many(cols)
        lag = d.lagrange_to_coeff_many([c.copy() for c in cols])
        for q in range(m):
            assert (ext[q] == h2ref.coeff_to_extended(dc, cols[q])).all(), ("coeff_to_extended_many", k, m, q)
            assert (lag[q] == h2ref.lagrange_to_coeff(dc, cols[q])).all(), ("lagrange_to_coeff_many", k, m, q)
_ffi.shutdown()
print("multi-device ok on", ndev, "devices")
