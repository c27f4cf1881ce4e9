"""Host-side logic on CPU: shard ranges and the world_size-2 gather path over gloo."""
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions():
    from halo2_prover_b200.multi_gpu import shard_range
    for n in (0, 1, 7, 8, 1000, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch
    import torch.distributed as dist
    import h2ref
    from halo2_prover_b200.multi_gpu import gather_partials, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc, pts = h2ref.random_fr(n, 11), h2ref.random_g1(n, 12)
    lo, hi = shard_range(n, rank, world)
    # the per-rank MSM is the GPU's job; here the checker stands in so the exchange can be tested on CPU
    part = h2ref.best_multiexp(sc[lo:hi].copy(), pts[lo:hi].copy(), 1)
    t = torch.from_numpy(part.view(np.int64).copy())
    parts = gather_partials(t).numpy().view(np.uint64)
    acc = parts[0].copy()
    for r in range(1, world):
        acc = h2ref.g1_add(acc, parts[r].copy())
    q.put((rank, h2ref.g1_to_affine(acc).tolist()))
    dist.destroy_process_group()


def test_sharded_msm_gather_gloo_world2():
    import torch.multiprocessing as mp
    import h2ref
    n, world = 96, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc, pts = h2ref.random_fr(n, 11), h2ref.random_g1(n, 12)
    want = h2ref.g1_to_affine(h2ref.best_multiexp(sc, pts, 1)).tolist()
    assert res[0] == want and res[1] == want


def _col_worker(rank, world, port, m, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import torch.distributed as dist
    import h2ref
    from halo2_prover_b200.multi_gpu import commit_columns
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 32
    pts = h2ref.random_g1(n, 21)
    cols = [h2ref.random_fr(n, 30 + c) for c in range(m)]
    # the per-rank batched commit is the GPU's job; the checker stands in so the scheduling is tested on CPU
    fn = lambda cs: np.stack([h2ref.best_multiexp(c, pts, 1) for c in cs]) if cs else np.zeros((0, 12), dtype=np.uint64)
    res = commit_columns(fn, cols)
    q.put((rank, [h2ref.g1_to_affine(r).tolist() for r in res]))
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [1, 5])
def test_columns_round_robin_gloo_world2(m):
    """Independent columns dealt round-robin over two ranks: every rank ends with every commitment, in order."""
    import torch.multiprocessing as mp
    import h2ref
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_col_worker, args=(r, world, port, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts = h2ref.random_g1(32, 21)
    want = [h2ref.g1_to_affine(h2ref.best_multiexp(h2ref.random_fr(32, 30 + c), pts, 1)).tolist() for c in range(m)]
    assert res[0] == want and res[1] == want
