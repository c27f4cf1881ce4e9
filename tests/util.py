import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unhx(s: str, cols: int) -> np.ndarray:
    vals = [int(s[i:i + 16], 16) for i in range(0, len(s), 16)]
    return np.array(vals, dtype=np.uint64).reshape(-1, cols)


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rand_fr(n: int, seed: int) -> np.ndarray:
    """n values uniform in [0, r) as (n, 4) uint64 limbs -- each is the Montgomery form of exactly one field
    element, so they are valid Fr inputs as they are (numpy: 2^24 values in about a second)."""
    rng = np.random.default_rng(seed)
    top = np.uint64(0x30644E72E131A029)  # top limb of r
    a = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 4), dtype=np.uint64, endpoint=True)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    bad = a[:, 3] >= top
    while bad.any():
        m = int(bad.sum())
        a[bad] = rng.integers(0, np.iinfo(np.uint64).max, size=(m, 4), dtype=np.uint64, endpoint=True)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        bad = a[:, 3] >= top
    return a
