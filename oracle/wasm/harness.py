"""Runs the reference's shipped prover under oracle/wasm/wasmrun and parses what it wrote.

TEST INFRASTRUCTURE ONLY (oracle/).  `run()` executes setup(k) -> wasm_generate_proof -> wasm_verify_proof of one
of the reference's circuits with a seeded RNG; `hot` selects who answers the prover's best_multiexp / best_fft
calls: the interpreted module itself (None), the library under test ("gpu"), or the C restatement ("cpu")."""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
WASMRUN = os.path.join(HERE, "wasmrun")
LIB_GPU = os.path.join(ROOT, "halo2-prover_b200", "csrc", "libh2b200.so")
LIB_CPU = os.path.join(ROOT, "oracle", "libh2ref.so")

# name -> (circuit index in wasm.rs:82-119, input JSON); the seeds / k of the committed fixtures are in
# tests/golden/wasm_manifest.json
CIRCUITS = {
    "arithmetic": (1, '{"x": 6, "y": 9, "constant": 7, "z": 2923}'),
    "poseidon": (2, '{"x": [1, 2]@SIMULATE@}'),
    "collatz": (0, '{ "x": [5, 16, 8, 4, 2, 1]}'),
}


def wasm_path() -> str | None:
    for p in (os.path.join(ROOT, "oracle", "_ref", "halo2_prover_bg.wasm"),
              "/root/reference/src/lib/wasm/halo2_prover_bg.wasm"):
        if os.path.exists(p):
            return p
    return None


def available() -> bool:
    return wasm_path() is not None and os.path.exists(WASMRUN)


def parse_meta(path: str) -> dict:
    """params / proof / verify flag / call counts of a wasmrun output file (records of the hot calls are skipped)."""
    data = open(path, "rb").read()
    off, meta = 0, {}
    while off < len(data):
        kind, n = struct.unpack_from("<II", data, off)
        off += 8
        if kind == 1:
            off += 96 * n + 96
        elif kind == 2:
            off += 32 + 64 * (1 << n)
        elif kind in (10, 11, 15):
            meta[{10: "params", 11: "proof", 15: "input"}[kind]] = data[off:off + n]
            off += n
        elif kind in (12, 13, 14):
            meta[{12: "verify_ok", 13: "msm_calls_prove", 14: "fft_calls_prove"}[kind]] = n
        else:
            raise ValueError(f"bad record kind {kind}")
    return meta


def run(circuit: str, k: int, seed: int, hot: str | None = None, threads: int | None = None, record: bool = False,
        timeout: int = 1800, keep: str | None = None, repeat: int = 0) -> tuple[dict, dict]:
    """-> (meta, stats): meta as parse_meta, stats the harness's JSON line (times inside the dispatched calls).
    repeat > 0 proves that many more times in the same process (same random stream; the harness insists on identical
    bytes) and reports the last repeat's hot-path time as steady_*."""
    idx, inp = CIRCUITS[circuit]
    env = dict(os.environ)
    env["WASMRUN_REPEAT"] = str(repeat)
    env["WASMRUN_RECORD"] = "1" if record else "0"
    if hot == "gpu":
        env["WASMRUN_HOT"] = "gpu:" + LIB_GPU
    elif hot == "cpu":
        env["WASMRUN_HOT"] = "cpu:" + LIB_CPU
        env["WASMRUN_CPU_THREADS"] = str(threads or os.cpu_count() or 1)
    else:
        env.pop("WASMRUN_HOT", None)
    out = keep or tempfile.mktemp(suffix=".bin", prefix="wasmrun_")
    try:
        cmd = f"ulimit -s unlimited; exec '{WASMRUN}' '{wasm_path()}' '{out}' {k} {idx} '{inp}' {seed}"
        r = subprocess.run(["bash", "-c", cmd], env=env, capture_output=True, text=True, timeout=timeout)
        if r.returncode not in (0, 1):
            raise RuntimeError(f"wasmrun failed ({r.returncode}): {r.stderr[-2000:]}")
        stats = json.loads(r.stdout.strip().splitlines()[-1])
        return parse_meta(out), stats
    finally:
        if not keep and os.path.exists(out):
            os.remove(out)
