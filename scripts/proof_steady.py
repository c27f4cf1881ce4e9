"""Developer probe (GPU box): one real Poseidon proof of the reference prover through libh2b200.so at the given k, then
the same proof again in the same process (steady state).  Prints the harness's timing line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "wasm"))
import harness  # noqa: E402

for k in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "14").split(",")]:
    meta, st = harness.run("poseidon", k, 4242, hot=sys.argv[2] if len(sys.argv) > 2 else "gpu", repeat=1)
    st["proof_bytes"] = len(meta["proof"])
    st["verify_ok_meta"] = meta["verify_ok"]
    print(json.dumps(st), flush=True)
