"""ctypes binding of the C ABI declared in include/h2b200.h.

The library is looked up in-tree only (``csrc/libh2b200.so``).  If it is missing the
loader raises -- there is deliberately no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libh2b200.so")

H2B_OK = 0
ERR_NAMES = {-1: "H2B_ERR_ARG", -2: "H2B_ERR_CUDA", -3: "H2B_ERR_OOM", -4: "H2B_ERR_STATE"}

# every symbol include/h2b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "h2b_init", "h2b_init_devices", "h2b_device_count", "h2b_shutdown", "h2b_last_error", "h2b_abi_version",
    "h2b_best_multiexp", "h2b_srs_register", "h2b_srs_release", "h2b_commit", "h2b_g1_fold", "h2b_g1_to_bytes", "h2b_g_to_lagrange", "h2b_dev_evaluate_h", "h2b_dev_evaluate_h_lookup",
    "h2b_best_fft", "h2b_domain_new", "h2b_lagrange_to_coeff", "h2b_coeff_to_extended",
    "h2b_extended_to_coeff", "h2b_divide_by_vanishing_poly", "h2b_lagrange_to_coeff_many", "h2b_coeff_to_extended_many",
    "h2b_dev_lagrange_to_coeff_many", "h2b_dev_coeff_to_extended_many", "h2b_dev_divide_by_vanishing_poly",
    "h2b_dev_srs_register", "h2b_dev_msm", "h2b_dev_commit", "h2b_dev_commit_many", "h2b_commit_many", "h2b_srs_device_ptr", "h2b_dev_best_fft", "h2b_dev_lagrange_to_coeff",
    "h2b_dev_coeff_to_extended", "h2b_dev_extended_to_coeff", "h2b_dev_g1_fold", "h2b_dev_fixed_base_mul",
    "h2b_set_msm_window", "h2b_srs_info", "h2b_srs_layout", "h2b_test_set_max_entries", "h2b_params_read", "h2b_params_write", "h2b_set_srs_precompute", "h2b_set_srs_table_stride", "h2b_set_h2d_bandwidth", "h2b_set_e2e_chunking", "h2b_kernel_launches", "h2b_set_kernel_timing", "h2b_kernel_time_collect",
    "h2b_test_field_op", "h2b_test_g1_add_affine", "h2b_imad_peak",
]


class H2BError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Domain(C.Structure):
    """struct h2b_domain."""
    _fields_ = [
        ("k", C.c_uint32), ("extended_k", C.c_uint32), ("j", C.c_uint32), ("n_t", C.c_uint32),
        ("omega", C.c_uint64 * 4), ("omega_inv", C.c_uint64 * 4),
        ("extended_omega", C.c_uint64 * 4), ("extended_omega_inv", C.c_uint64 * 4),
        ("g_coset", C.c_uint64 * 4), ("g_coset_inv", C.c_uint64 * 4),
        ("ifft_divisor", C.c_uint64 * 4), ("extended_ifft_divisor", C.c_uint64 * 4),
        ("t_evaluations", C.c_uint64 * 128),
        ("extended_ifft_coset", C.c_uint64 * 12),
    ]


_lib = None
_lock = threading.Lock()
_inited_device = None


def lib() -> C.CDLL:
    """Load libh2b200.so (no GPU needed just to load and inspect symbols)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                    "(nvcc, sm_100a).  halo2-prover_b200 has no CPU fallback."
                )
            L = C.CDLL(LIB_PATH)
            L.h2b_last_error.restype = C.c_char_p
            L.h2b_abi_version.restype = C.c_uint32
            L.h2b_kernel_launches.restype = C.c_uint64
            for name in SYMBOLS:
                fn = getattr(L, name)
                if fn.restype is C.c_int and name != "h2b_shutdown":
                    fn.restype = C.c_int
            L.h2b_shutdown.restype = None
            _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != H2B_OK:
        raise H2BError(rc, lib().h2b_last_error().decode(errors="replace"))


def init(device: int | None = None) -> None:
    """h2b_init on ``device`` (default: LOCAL_RANK or 0).  Raises if no GPU is usable."""
    global _inited_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _inited_device == device:
        return
    check(lib().h2b_init(C.c_int(device)))
    _inited_device = device


def init_devices(devices) -> None:
    """h2b_init_devices: one process, several GPUs (devices[0] is the primary device)."""
    global _inited_device
    devices = [int(d) for d in devices]
    arr = (C.c_int * len(devices))(*devices)
    check(lib().h2b_init_devices(arr, C.c_int(len(devices))))
    _inited_device = devices[0]


def shutdown() -> None:
    global _inited_device
    if _lib is not None:
        _lib.h2b_shutdown()
    _inited_device = None


def u64p(a: np.ndarray):
    if a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("expected a C-contiguous uint64 array")
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def as_u64(a, cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2 or a.shape[1] != cols:
        a = a.reshape(-1, cols)
    return a
