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


def _param_counts_header():
    txt = open(os.path.join(ROOT, "include", "h2b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(h2b_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_rust_binding_declares_header_functions_with_the_same_arity():
    """rust/h2b200-sys cannot be compiled in this image (no rustc): at least every `pub fn h2b_*` of its extern block
    must name a function of the header and take as many parameters."""
    src = open(os.path.join(ROOT, "rust", "h2b200-sys", "src", "lib.rs")).read()
    block = src[src.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    block = re.sub(r"/\*.*?\*/", "", block, flags=re.S)
    block = re.sub(r"//[^\n]*", "", block)
    decls = re.findall(r"pub fn (h2b_[a-z0-9_]+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", block, flags=re.S)
    want = _param_counts_header()
    assert len(decls) >= 40
    for name, params in decls:
        assert name in want, name
        n = 0 if not params.strip() else params.count(":")
        assert n == want[name], (name, n, want[name])


def test_halo2_proofs_patch_uses_only_defined_items():
    """The round-1 patch called helpers nothing defined.  Every `h2b200_sys::item` the patch uses must exist in the
    crate, and every `b200_*` helper it calls must be defined by one of its own hunks."""
    patch = open(os.path.join(ROOT, "rust", "patches", "halo2_proofs-6b43b6b.patch")).read()
    crate = open(os.path.join(ROOT, "rust", "h2b200-sys", "src", "lib.rs")).read()
    added = "\n".join(line[1:] for line in patch.splitlines() if line.startswith("+") and not line.startswith("+++"))
    free_fns = set(re.findall(r"^pub fn ([a-z0-9_]+)", crate, flags=re.M))
    types = set(re.findall(r"^pub struct ([A-Za-z0-9_]+)", crate, flags=re.M))
    methods = set(re.findall(r"^\s+pub fn ([a-z0-9_]+)", crate, flags=re.M))
    for path in set(re.findall(r"h2b200_sys::([A-Za-z0-9_]+(?:::[a-z0-9_]+)?)", added)):
        head, _, tail = path.partition("::")
        assert head in free_fns | types, path
        if tail:
            assert tail in methods, path
    called = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", added))
    defined = set(re.findall(r"fn (b200_[a-z0-9_]+)", added))
    assert called and called <= defined, called - defined
    for m in set(re.findall(r"\bsrs\.([a-z0-9_]+)\(|\bd\.([a-z0-9_]+)\(", added)):
        name = m[0] or m[1]
        assert name in methods, name


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


def _build_c_test():
    import subprocess
    exe = os.path.join(ROOT, "tests", "native", "abi_c_test")
    src = exe + ".c"
    libdir = os.path.join(ROOT, "halo2-prover_b200", "csrc")
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, src,
                               "-L" + libdir, "-lh2b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_program_links_against_the_header_and_gets_error_codes_without_a_gpu():
    """A plain C program with Rust-layout structs compiles against include/h2b200.h, links libh2b200.so, and --
    where there is no GPU -- gets status codes, never a crash or a CPU result."""
    import subprocess
    exe = _build_c_test()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by test_c_program_computes_through_the_abi")
    r = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "nogpu ok" in r.stdout


@pytest.mark.gpu
def test_c_program_computes_through_the_abi(tmp_path, href, spec):
    """The same C program on a GPU: Rust-layout buffers in, results compared with the oracle."""
    import struct
    import subprocess
    import numpy as np
    exe = _build_c_test()
    n, log_n = 3000, 11
    sc, bases = href.random_fr(n, 81), href.random_g1(n, 82)
    a = href.random_fr(1 << log_n, 83)
    omega = spec.fr_array([pow(spec.ROOT_OF_UNITY, 1 << (spec.FR_S - log_n), spec.R_MOD)])[0]
    inp, out = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<QQ", n, log_n))
        f.write(sc.tobytes()); f.write(bases.tobytes()); f.write(omega.tobytes()); f.write(a.tobytes())
    r = subprocess.run([exe, "run", str(inp), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "abi_c_test ok" in r.stdout, r.stdout + r.stderr
    res = np.fromfile(out, dtype=np.uint64)
    m = 1 << log_n
    msm, commit, half = res[:12], res[12:24], res[24:36]
    off = 36
    fft = res[off:off + 4 * m].reshape(m, 4); off += 4 * m
    lag = res[off:off + 4 * m].reshape(m, 4); off += 4 * m
    ext = res[off:off + 16 * m].reshape(4 * m, 4)
    want = href.g1_to_affine(href.best_multiexp(sc, bases))
    assert (href.g1_to_affine(msm) == want).all() and (href.g1_to_affine(commit) == want).all()
    h = n // 2
    assert (href.g1_to_affine(half) == href.g1_to_affine(href.best_multiexp(sc[:h].copy(), bases[:h].copy()))).all()
    assert (fft == href.best_fft(a, omega, log_n)).all()
    dc = href.domain_new(4, log_n)
    assert (lag == href.lagrange_to_coeff(dc, a)).all()
    assert (ext == href.coeff_to_extended(dc, a)).all()
