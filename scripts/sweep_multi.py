"""Strong-scaling sweep of the point-range sharded commit (GPU box, under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/sweep_multi.py

For every total size 2^lg (SWEEP_LG, default 20,22,24,26) rank g of N owns points [g n/N, (g+1) n/N): it registers that
slice of a synthetic SRS (window table included), and one step is multi_gpu.sharded_commit = this rank's
h2b_dev_commit + all-gather of N x 96 B + fold.  Timed with CUDA events between barriers, max over ranks.
Rank 0 writes gpurun_out/sweep_multi_N.json.  Not the bench.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from probe import rand_fr_np  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import halo2_prover_b200 as h2b
    from halo2_prover_b200 import _ffi, multi_gpu
    import bn254
    _ffi.init(local)
    L = _ffi.lib()
    sizes = [int(x) for x in os.environ.get("SWEEP_LG", "20,22,24,26").split(",")]
    stream = torch.cuda.Stream()
    gen = bn254.affine_to_array([bn254.G1_GENERATOR])[0]
    nmax = (1 << max(sizes)) // world
    res = {"n_gpus": world, "sizes": {}}
    with torch.cuda.stream(stream):
        scal = torch.from_numpy(rand_fr_np(nmax, 100 + rank).view(np.int64)).cuda()
        seeds = torch.from_numpy(rand_fr_np(nmax, 200 + rank).view(np.int64)).cuda()
        bases = torch.empty((nmax, 8), dtype=torch.int64, device="cuda")
        _ffi.check(L.h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(nmax), _ffi.u64p(gen),
                                            C.c_void_p(bases.data_ptr()), C.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        del seeds
        for lg in sizes:
            n = (1 << lg) // world
            lg_local = max(n.bit_length() - 1, 0)
            params = h2b.ParamsKZG.from_device(lg_local, bases[:n])
            for _ in range(3):
                out = multi_gpu.sharded_commit(params, scal[:n], stream=stream)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            reps = 5
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(reps):
                out = multi_gpu.sharded_commit(params, scal[:n], stream=stream)
            e1.record(stream)
            stream.synchronize()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            # every rank must hold the same folded point
            chk = out.clone()
            if world > 1:
                dist.broadcast(chk, src=0)
            same = bool((chk == out).all().item())
            ms = float(t.item())
            res["sizes"][str(lg)] = {"ms": ms, "points_per_s": (1 << lg) / ms * 1e3, "per_rank_points": n, "ranks_agree": same}
            if rank == 0:
                print(f"N={world} total 2^{lg}: {ms:.3f} ms ({(1 << lg) / ms * 1e3:.3e} pts/s) agree={same}", flush=True)
            params.release()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"sweep_multi_{world}.json"), "w") as f:
            json.dump(res, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
