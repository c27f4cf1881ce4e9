"""Developer probe (GPU box): per-call-type time of the proof-shaped replay (bench.py: bench_proof_replay)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import bench  # noqa: E402
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi  # noqa: E402

_ffi.init(0)
for k in [int(x) for x in os.environ.get("REPLAY_K", "10,14").split(",")]:
    args = types.SimpleNamespace(proof_k=k)
    r = bench.bench_proof_replay(args, h2b, _ffi)
    print(json.dumps({"k": k, "gpu_ms": r["gpu_ms"], "by_call": r["gpu_ms_by_call"], "batched_ms": r["gpu_batched_ms"], "batched_by_call": r["gpu_batched_ms_by_call"], "beq": r["batched_equals_single"], "cpu_ms": r["cpu_ms"], "equal": r["commitments_equal"]}))
