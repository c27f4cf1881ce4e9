"""Developer probe (GPU box): h2b_commit of 2^LG points with PAGEABLE host scalars (what a Rust Vec is)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import halo2_prover_b200 as h2b
from halo2_prover_b200 import _ffi
import bench, torch, bn254
_ffi.init(0)
L = _ffi.lib()
lg = int(os.environ.get("LG", "24")); n = 1 << lg
s = torch.cuda.Stream()
gen = bn254.affine_to_array([bn254.G1_GENERATOR])[0]
with torch.cuda.stream(s):
    seeds = torch.from_numpy(bench.rand_fr_np(n, 2).view(np.int64)).cuda()
    bases = torch.empty((n, 8), dtype=torch.int64, device="cuda")
    _ffi.check(L.h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(n), _ffi.u64p(gen), C.c_void_p(bases.data_ptr()), C.c_void_p(s.cuda_stream)))
    s.synchronize()
params = h2b.ParamsKZG(lg, bases.cpu().numpy().view(np.uint64))
sc = bench.rand_fr_np(n, 1)          # ordinary numpy memory: pageable
for _ in range(2): params.commit(sc)
t = time.perf_counter()
for _ in range(3): params.commit(sc)
dt = (time.perf_counter() - t) / 3
print(f"commit 2^{lg} pageable scalars, H2B_COPY_THREADS={os.environ.get('H2B_COPY_THREADS','default')}: {dt*1e3:.2f} ms ({n/dt:.3e} pts/s)")
