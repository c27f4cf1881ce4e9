import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import halo2_prover_b200 as h2b
from halo2_prover_b200 import _ffi
import bench, h2ref
_ffi.init(0)
k = int(os.environ.get("K", "14")); m = int(os.environ.get("M", "8"))
n = 1 << k
g = h2ref.random_g1(1 << 10, 5)
g = np.ascontiguousarray(np.tile(g, (n >> 10, 1))) if n >= 1024 else g[:n].copy()
params = h2b.ParamsKZG(k, g)
cols = [bench.rand_fr_np(n, 300 + i) for i in range(m)]
for _ in range(int(os.environ.get("REPS", "3"))):
    t = time.perf_counter(); params.commit_many(cols); print("commit_many ms", (time.perf_counter() - t) * 1e3)
