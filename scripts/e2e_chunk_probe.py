"""End-to-end commit (pinned host scalars -> h2b_commit) at a device's share of a sharded 2^24-point commit, by the
number of pieces the copy is pipelined in.  PROBE_N = log2 sizes."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from halo2_prover_b200 import _ffi  # noqa: E402
import bn254  # noqa: E402

_ffi.init(0)
L = _ffi.lib()
gen = bn254.affine_to_array([bn254.G1_GENERATOR])[0]
rng = np.random.default_rng(1)
for lg in [int(x) for x in os.environ.get("PROBE_N", "21,22,23,24").split(",")]:
    n = 1 << lg
    a = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 4), dtype=np.uint64, endpoint=True)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    hs = torch.from_numpy(a.view(np.int64)).pin_memory()
    s = torch.cuda.current_stream()
    seeds = hs.cuda()
    db = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    _ffi.check(L.h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(n), _ffi.u64p(gen), C.c_void_p(db.data_ptr()), C.c_void_p(1)))
    torch.cuda.synchronize()
    h = C.c_uint64(0)
    _ffi.check(L.h2b_dev_srs_register(C.c_void_p(db.data_ptr()), C.c_size_t(n), C.byref(h)))
    res = np.zeros(12, dtype=np.uint64)
    hp = C.cast(C.c_void_p(hs.data_ptr()), C.POINTER(C.c_uint64))
    out = torch.empty(12, dtype=torch.int64, device="cuda")
    for _ in range(3):
        _ffi.check(L.h2b_dev_commit(h, C.c_void_p(seeds.data_ptr()), C.c_size_t(n), C.c_void_p(out.data_ptr()), C.c_void_p(1)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        _ffi.check(L.h2b_dev_commit(h, C.c_void_p(seeds.data_ptr()), C.c_size_t(n), C.c_void_p(out.data_ptr()), C.c_void_p(1)))
    torch.cuda.synchronize()
    line = f"2^{lg}: resident {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms; e2e by pieces:"
    for chunks in (1, 2, 3, 4, 8):
        _ffi.check(L.h2b_set_e2e_chunking(chunks, C.c_size_t(1 << 16)))
        for _ in range(2):
            _ffi.check(L.h2b_commit(h, hp, C.c_size_t(n), _ffi.u64p(res)))
        t0 = time.perf_counter()
        for _ in range(5):
            _ffi.check(L.h2b_commit(h, hp, C.c_size_t(n), _ffi.u64p(res)))
        line += f"  {chunks}: {(time.perf_counter() - t0) / 5 * 1e3:.3f}"
    print(line, flush=True)
    _ffi.check(L.h2b_srs_release(h))
    del db, seeds, hs
