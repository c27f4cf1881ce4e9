"""Developer probe (GPU box): time inside the dispatched best_fft / commit calls of a REAL Poseidon proof (the reference's
prover under oracle/wasm/wasmrun) with the library's pinned-ring staging on and off.  Calls in a real prover are
separated by long stretches of host work, so the device, the link and the copy threads are cold at every call -- a
different regime from a tight loop (scripts/hostpath_probe.py)."""
import collections
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import harness  # noqa: E402


def run(k, extra_env):
    idx, inp = harness.CIRCUITS["poseidon"]
    env = dict(os.environ, WASMRUN_RECORD="0", WASMRUN_HOT="gpu:" + harness.LIB_GPU, WASMRUN_TRACE="1", **extra_env)
    out = tempfile.mktemp(suffix=".bin")
    cmd = f"ulimit -s unlimited; exec '{harness.WASMRUN}' '{harness.wasm_path()}' '{out}' {k} {idx} '{inp}' 4242"
    r = subprocess.run(["bash", "-c", cmd], env=env, capture_output=True, text=True, timeout=1500)
    os.remove(out)
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    per = collections.defaultdict(list)
    for m in re.finditer(r"hot (fft log_n|commit n)=(\d+) ([\d.]+) ms", r.stderr):
        per[(m.group(1).split()[0], int(m.group(2)))].append(float(m.group(3)))
    return stats, per


def main():
    ks = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12").split(",")]
    variants = [("default", {}), ("driver staging (H2B_COPY_THREADS=0)", {"H2B_COPY_THREADS": "0"})]
    for extra in sys.argv[2:]:
        variants.append((extra, dict(kv.split("=") for kv in extra.split(","))))
    for k in ks:
        for name, env in variants:
            stats, per = run(k, env)
            print(f"k={k} {name}: prove fft {stats['hot_fft_ms_prove']:.2f} ms, msm {stats['hot_msm_ms_prove']:.2f} ms, "
                  f"total fft {stats['hot_fft_ms_total']:.2f} ms, register {stats['srs_register_ms']:.1f} ms", flush=True)
            for key in sorted(per):
                v = sorted(per[key])
                print(f"    {key[0]} {key[1]}: {len(v)} calls, min {v[0]:.3f} median {v[len(v) // 2]:.3f} max {v[-1]:.3f} ms")


if __name__ == "__main__":
    main()
