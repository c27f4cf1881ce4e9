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
