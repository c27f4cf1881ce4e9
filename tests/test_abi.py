"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the
header declares, and refuses to compute without a GPU (no fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "h2b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(h2b_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from halo2_prover_b200 import _ffi
    assert _header_functions() == sorted(_ffi.SYMBOLS)


def test_library_exports_every_symbol():
    from halo2_prover_b200 import _ffi
    L = _ffi.lib()
    for name in _header_functions():
        assert hasattr(L, name), name
    assert L.h2b_abi_version() == 2


def test_struct_layout_matches_header():
    import ctypes as C
    from halo2_prover_b200 import _ffi
    # 4 x u32 + 8 x 32 B constants + 32 x 32 B t_evaluations + 3 x 32 B derived
    assert C.sizeof(_ffi.Domain) == 16 + 8 * 32 + 32 * 32 + 3 * 32


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU h2b_init must fail and compute calls must raise."""
    import numpy as np
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    import halo2_prover_b200 as pkg
    from halo2_prover_b200 import _ffi
    with pytest.raises(_ffi.H2BError):
        _ffi.init(0)
    with pytest.raises(_ffi.H2BError):
        pkg.best_multiexp(np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 8), dtype=np.uint64))
    # compute entry points called without init report H2B_ERR_STATE, not a result
    out = np.zeros(12, dtype=np.uint64)
    rc = _ffi.lib().h2b_g1_fold(_ffi.u64p(out), 0, _ffi.u64p(out))
    assert rc == -4


def test_product_does_not_import_oracle():
    """The product package must not reference oracle/ (checker only)."""
    pkg_dir = os.path.join(ROOT, "halo2-prover_b200")
    for base, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f), errors="replace").read()
                assert "h2ref" not in txt and "import bn254" not in txt and "oracle/" not in txt, f
