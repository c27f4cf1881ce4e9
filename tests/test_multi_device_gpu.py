"""Multi-GPU parity on hardware (skipped with fewer than two GPUs): the single-process multi-device paths of the C
ABI (h2b_init_devices: sharded and replicated SRS, dealt columns) and the one-process-per-GPU NCCL path
(multi_gpu.sharded_commit), each against the oracle.  Both run in fresh processes: the library context of this
pytest process is bound to cuda:0."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [2, 4, 8])
def test_single_process_multi_device_vs_oracle(ndev):
    if _gpus() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    env = dict(os.environ, H2B_SHARD_MIN_LOG="13")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "workers", "multi_device_worker.py"), str(ndev)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi-device ok" in r.stdout


@pytest.mark.parametrize("world", [2, 8])
def test_nccl_sharded_commit_vs_oracle(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 2000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "workers", "nccl_shard_worker.py"), str(1 << 16)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "nccl sharded commit ok" in r.stdout


def test_single_device_through_init_devices():
    """h2b_init_devices with one device is h2b_init (runs on any GPU box)."""
    r = subprocess.run([sys.executable, "-c",
                        "import sys; sys.path[:0]=[%r, %r]\n"
                        "import h2ref, halo2_prover_b200 as h2b\n"
                        "from halo2_prover_b200 import _ffi\n"
                        "_ffi.init_devices([0]); assert _ffi.lib().h2b_device_count() == 1\n"
                        "b, s = h2ref.random_g1(4096, 1), h2ref.random_fr(4096, 2)\n"
                        "p = h2b.ParamsKZG(12, b)\n"
                        "assert (h2ref.g1_to_affine(p.commit(s)) == h2ref.g1_to_affine(h2ref.best_multiexp(s, b))).all()\n"
                        "p.release(); _ffi.shutdown(); print('ok')\n" % (ROOT, os.path.join(ROOT, "oracle"))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
