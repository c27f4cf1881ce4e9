"""The pageable-host-memory copy policies of csrc/hostcopy.h (chosen by environment at h2b_init, so each runs in its
own process): results must not depend on which path moved the bytes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

POLICIES = {
    "default": {},
    "driver staging only": {"H2B_COPY_THREADS": "0"},
    "ring + workers for everything": {"H2B_MIN_STAGED_H2D_KB": "64", "H2B_MIN_STAGED_D2H_KB": "64", "H2B_MIN_STAGED_RT_KB": "64",
                                      "H2B_COPY_WORKERS_MIN_KB": "64"},
    "ring, calling thread only": {"H2B_MIN_STAGED_H2D_KB": "64", "H2B_MIN_STAGED_D2H_KB": "64", "H2B_MIN_STAGED_RT_KB": "64",
                                  "H2B_COPY_WORKERS_MIN_KB": "1048576"},
}


@pytest.mark.parametrize("policy", list(POLICIES))
def test_copy_policy_does_not_change_results(policy):
    env = dict(os.environ, **POLICIES[policy])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "workers", "hostpath_worker.py")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.strip().splitlines()[-1] == "OK 8"
