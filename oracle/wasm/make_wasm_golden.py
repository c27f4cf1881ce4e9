"""Runs the reference's own compiled prover under oracle/wasm/wasmrun and turns what it records at
best_multiexp (wasm func 347) / best_fft (wasm func 80) into committed fixtures:

    python oracle/wasm/make_wasm_golden.py            # needs /root/reference (build container only)

Writes tests/golden/wasm_<circuit>_k<k>.npz (+ a manifest JSON).  Every record is *also* checked
here against the big-integer spec / C restatement, which is what pins the oracle to the
reference's execution; the tests then replay the same records against the oracle (CPU) and the
CUDA path (GPU) without needing the wasm.

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import bn254 as spec  # noqa: E402
import h2ref  # noqa: E402

WASM = "/root/reference/src/lib/wasm/halo2_prover_bg.wasm"
GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, circuit index in wasm.rs:82-119, k, input JSON, RNG seed, keep-all?)
# Collatz needs k = 10 (NotEnoughRowsAvailable below that; the web demo uses setup(10),
# src/components/Circuits.tsx:90); its 7 MB of records are all CHECKED here but only a sample is
# committed (every MSM against the SRS, first three MSMs of each size with their own bases, first two
# FFTs of each log_n) to keep fixtures small.
RUNS = [
    ("arithmetic", 1, 4, '{"x": 6, "y": 9, "constant": 7, "z": 2923}', 12345, True),
    ("poseidon", 2, 7, '{"x": [1, 2]@SIMULATE@}', 4242, True),
    ("collatz", 0, 10, '{ "x": [5, 16, 8, 4, 2, 1]}', 777, False),
]


def parse(path):
    data = open(path, "rb").read()
    off = 0
    recs = []
    meta = {}
    while off < len(data):
        kind, n = struct.unpack_from("<II", data, off)
        off += 8
        if kind == 1:
            sc = np.frombuffer(data, dtype=np.uint64, count=4 * n, offset=off).reshape(n, 4); off += 32 * n
            bs = np.frombuffer(data, dtype=np.uint64, count=8 * n, offset=off).reshape(n, 8); off += 64 * n
            out = np.frombuffer(data, dtype=np.uint64, count=12, offset=off); off += 96
            recs.append(("msm", sc.copy(), bs.copy(), out.copy()))
        elif kind == 2:
            m = 1 << n
            om = np.frombuffer(data, dtype=np.uint64, count=4, offset=off); off += 32
            a = np.frombuffer(data, dtype=np.uint64, count=4 * m, offset=off).reshape(m, 4); off += 32 * m
            b = np.frombuffer(data, dtype=np.uint64, count=4 * m, offset=off).reshape(m, 4); off += 32 * m
            recs.append(("fft", n, om.copy(), a.copy(), b.copy()))
        elif kind in (10, 11, 15):
            meta[{10: "params", 11: "proof", 15: "input"}[kind]] = data[off:off + n]
            off += n
        elif kind in (12, 13, 14):
            meta[{12: "verify_ok", 13: "msm_calls_prove", 14: "fft_calls_prove"}[kind]] = n
        else:
            raise ValueError(f"bad record kind {kind}")
    return recs, meta


def main():
    subprocess.check_call(["make", "-C", HERE, "-s"])
    manifest = {}
    for name, circuit, k, inp, seed, keep_all in RUNS:
        out_bin = f"/tmp/wasm_{name}_k{k}.bin"
        if not os.path.exists(out_bin) or os.environ.get("WASM_GOLDEN_RERUN"):
            print(f"running the reference prover: {name} k={k} (interpreted; this takes minutes)", flush=True)
            subprocess.check_call(f"ulimit -s unlimited; {HERE}/wasmrun {WASM} {out_bin} {k} {circuit} '{inp}' {seed}",
                                  shell=True, executable="/bin/bash")
        recs, meta = parse(out_bin)
        assert meta["verify_ok"] == 1, "the reference verifier rejected the reference proof"
        params = np.frombuffer(meta["params"], dtype=np.uint8)
        n = 1 << k
        # ParamsKZG::write (RawBytes): k u32 LE | g[n] x 64 B | g_lagrange[n] x 64 B | g2 128 B | s_g2 128 B
        assert struct.unpack_from("<I", meta["params"], 0)[0] == k and len(meta["params"]) == 4 + 128 * n + 256
        g = np.frombuffer(meta["params"], dtype=np.uint64, count=8 * n, offset=4).reshape(n, 8)
        gl = np.frombuffer(meta["params"], dtype=np.uint64, count=8 * n, offset=4 + 64 * n).reshape(n, 8)
        arrays = {"params": params, "proof": np.frombuffer(meta["proof"], dtype=np.uint8)}
        msm_list, fft_list = [], []
        n_msm = n_fft = 0
        seen_msm, seen_fft = {}, {}
        for r in recs:
            if r[0] == "msm":
                _, sc, bs, out = r
                m = sc.shape[0]
                # pin: the reference's result == the oracle's, as group elements
                want_aff = h2ref.g1_to_affine(np.ascontiguousarray(out))
                got = h2ref.g1_to_affine(h2ref.best_multiexp(np.ascontiguousarray(sc), np.ascontiguousarray(bs), 1))
                assert (got == want_aff).all(), f"{name}: MSM record {n_msm} disagrees with the C restatement"
                if m <= 64:
                    sp = spec.msm_naive(spec.fr_ints(sc), spec.array_to_affine(bs))
                    assert (spec.affine_to_array([sp])[0] == want_aff).all(), "big-integer spec disagrees"
                seen_msm[m] = seen_msm.get(m, 0) + 1
                if m <= n and (bs == g[:m]).all():
                    src = "g"
                elif m <= n and (bs == gl[:m]).all():
                    src = "g_lagrange"
                else:
                    src = "explicit"
                # records against the SRS cost only their scalars: keep them all (the proof-byte tests need every
                # commitment); records with their own bases are sampled when the run is large
                if not keep_all and src == "explicit" and seen_msm[m] > 3:
                    n_msm += 1
                    continue
                if src == "explicit":
                    arrays[f"msm{n_msm}_bases"] = bs
                arrays[f"msm{n_msm}_scalars"] = sc
                arrays[f"msm{n_msm}_affine"] = want_aff
                msm_list.append({"i": n_msm, "n": int(m), "bases": src})
                n_msm += 1
            else:
                _, logn, om, a, b = r
                got = h2ref.best_fft(a, om, logn, 1)
                assert (got == b).all(), f"{name}: FFT record {n_fft} disagrees with the C restatement"
                if logn <= 7:
                    sp = spec.best_fft(spec.fr_ints(a), spec.fr_ints(om.reshape(1, 4))[0], logn)
                    assert spec.fr_array(sp).tolist() == b.tolist(), "big-integer spec disagrees"
                seen_fft[logn] = seen_fft.get(logn, 0) + 1
                if not keep_all and seen_fft[logn] > 2:
                    n_fft += 1
                    continue
                arrays[f"fft{n_fft}_omega"] = om
                arrays[f"fft{n_fft}_in"] = a
                arrays[f"fft{n_fft}_out"] = b
                fft_list.append({"i": n_fft, "log_n": int(logn)})
                n_fft += 1
        out_npz = os.path.join(GOLDEN, f"wasm_{name}_k{k}.npz")
        np.savez_compressed(out_npz, **arrays)
        manifest[name] = {
            "file": os.path.basename(out_npz), "k": k, "circuit_index": circuit, "input": meta["input"].decode(),
            "rng_seed": seed, "proof_bytes": len(meta["proof"]), "verified_by_reference_verifier": True,
            "msm_calls_total": n_msm, "fft_calls_total": n_fft, "all_records_committed": keep_all,
            "msm_calls_keygen_and_prove": meta["msm_calls_prove"], "fft_calls_keygen_and_prove": meta["fft_calls_prove"],
            "msm": msm_list, "fft": fft_list,
        }
        print(f"{name}: {n_msm} MSM + {n_fft} FFT records pinned against the oracle -> {out_npz} "
              f"({os.path.getsize(out_npz) / 1024:.0f} KiB)", flush=True)
    with open(os.path.join(GOLDEN, "wasm_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
