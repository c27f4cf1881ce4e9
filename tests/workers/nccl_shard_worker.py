"""Run under torchrun by tests/test_multi_device_gpu.py: one process per GPU, the point-range sharded commit of
multi_gpu.sharded_commit over NCCL, the folded result compared with the oracle's MSM over ALL ranks' points."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import h2ref  # noqa: E402
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi, multi_gpu  # noqa: E402

_ffi.init(local)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
bases, scalars = h2ref.random_g1(n, 1), h2ref.random_fr(n, 2)  # the same on every rank
lo, hi = multi_gpu.shard_range(n, rank, world)
params = h2b.ParamsKZG.__new__(h2b.ParamsKZG)
import ctypes as C  # noqa: E402
h = C.c_uint64(0)
mine = np.ascontiguousarray(bases[lo:hi])
_ffi.check(_ffi.lib().h2b_srs_register(_ffi.u64p(mine), C.c_size_t(hi - lo), C.byref(h)))
params.k, params.n, params._handles = 0, hi - lo, {"g": h.value}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ds = torch.from_numpy(np.ascontiguousarray(scalars[lo:hi]).view(np.int64)).cuda()
    res = multi_gpu.sharded_commit(params, ds, stream=s)
    s.synchronize()
got = h2ref.g1_to_affine(res.cpu().numpy().view(np.uint64))
want = h2ref.g1_to_affine(h2ref.best_multiexp(scalars, bases))
assert (got == want).all(), f"rank {rank}: folded result differs from the oracle"
# generic path (caller's bases, no table)
with torch.cuda.stream(s):
    db = torch.from_numpy(mine.view(np.int64)).cuda()
    res2 = multi_gpu.sharded_multiexp(ds, db, stream=s)
    s.synchronize()
assert (h2ref.g1_to_affine(res2.cpu().numpy().view(np.uint64)) == want).all()
params.release()
dist.barrier()
if rank == 0:
    print("nccl sharded commit ok on", world, "ranks, n =", n)
dist.destroy_process_group()
