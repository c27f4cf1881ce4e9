"""Developer probe (GPU box): integer-pipe peak, MSM and NTT timings at several sizes.  Not the bench."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import halo2_prover_b200 as h2b  # noqa: E402
from halo2_prover_b200 import _ffi, arithmetic  # noqa: E402


def rand_fr_np(n, seed):
    rng = np.random.default_rng(seed)
    r3 = np.uint64(0x30644E72E131A029)
    a = rng.integers(0, np.iinfo(np.uint64).max, size=(n, 4), dtype=np.uint64, endpoint=True)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    bad = a[:, 3] >= r3
    while bad.any():
        m = int(bad.sum())
        a[bad] = rng.integers(0, np.iinfo(np.uint64).max, size=(m, 4), dtype=np.uint64, endpoint=True)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        bad = a[:, 3] >= r3
    return a


def main():
    _ffi.init(0)
    L = _ffi.lib()
    res = {}
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _ffi.check(L.h2b_imad_peak(C.byref(a), C.byref(b), C.byref(c)))
    res["imad_gops"], res["imad_wide_gops"], res["sm_mhz_attr"] = a.value, b.value, c.value
    print("IMAD peak G/s", a.value, "IMAD.WIDE G/s", b.value, flush=True)
    s = torch.cuda.Stream()
    sizes = [int(x) for x in os.environ.get("PROBE_MSM", "16,20,22,24").split(",") if x not in ("", "none")]
    gen = np.array([1, 0, 0, 0, 2, 0, 0, 0], dtype=np.uint64)
    import bn254
    gen = bn254.affine_to_array([bn254.G1_GENERATOR])[0]
    with torch.cuda.stream(s):
        csizes = [int(x) for x in os.environ.get("PROBE_COMMIT", "").split(",") if x]
        nmax = 1 << max(sizes + csizes + [4])
        sc_all = torch.from_numpy(rand_fr_np(nmax, 1).view(np.int64)).cuda()
        seeds = torch.from_numpy(rand_fr_np(nmax, 2).view(np.int64)).cuda()
        bases = torch.empty((nmax, 8), dtype=torch.int64, device="cuda")
        t0 = time.time()
        _ffi.check(L.h2b_dev_fixed_base_mul(C.c_void_p(seeds.data_ptr()), C.c_size_t(nmax), _ffi.u64p(gen),
                                            C.c_void_p(bases.data_ptr()), C.c_void_p(s.cuda_stream)))
        s.synchronize()
        print("generated", nmax, "bases in", time.time() - t0, "s", flush=True)
        out = torch.empty(12, dtype=torch.int64, device="cuda")
        for lg in sizes:
            n = 1 << lg
            for cwin in [int(x) for x in os.environ.get("PROBE_C", "0").split(",")]:
                _ffi.check(L.h2b_set_msm_window(cwin))
                for _ in range(2):
                    arithmetic.dev_msm(sc_all[:n], bases, out, n=n, stream=s)
                s.synchronize()
                _ffi.check(L.h2b_set_kernel_timing(1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record(s)
                for _ in range(reps):
                    arithmetic.dev_msm(sc_all[:n], bases, out, n=n, stream=s)
                e1.record(s)
                s.synchronize()
                tot, calls = C.c_double(), C.c_uint32()
                _ffi.check(L.h2b_kernel_time_collect(C.byref(tot), C.byref(calls)))
                _ffi.check(L.h2b_set_kernel_timing(0))
                ms = e0.elapsed_time(e1) / reps
                kms = tot.value / max(calls.value, 1)
                print(f"MSM 2^{lg} c={cwin}: {ms:.3f} ms/step ({n / ms * 1e3:.3e} pts/s), accumulate {kms:.3f} ms", flush=True)
                res[f"msm_{lg}_c{cwin}"] = {"ms": ms, "acc_ms": kms}
        _ffi.check(L.h2b_set_msm_window(0))
        # commit against a registered SRS (precomputed window table, shared buckets), device-resident scalars
        for lg in [int(x) for x in os.environ.get("PROBE_COMMIT", "").split(",") if x]:
            n = 1 << lg
            hb = bases[:n].cpu().numpy().view(np.uint64)
            for cwin in [int(x) for x in os.environ.get("PROBE_SRS_C", "0").split(",")]:
                _ffi.check(L.h2b_set_srs_precompute(1, cwin))
                h = C.c_uint64(0)
                t0 = time.time()
                _ffi.check(L.h2b_srs_register(_ffi.u64p(hb), C.c_size_t(n), C.byref(h)))
                treg = time.time() - t0
                for _ in range(2):
                    _ffi.check(L.h2b_dev_commit(h, C.c_void_p(sc_all.data_ptr()), C.c_size_t(n), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(s.cuda_stream)))
                s.synchronize()
                _ffi.check(L.h2b_set_kernel_timing(1))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record(s)
                for _ in range(reps):
                    _ffi.check(L.h2b_dev_commit(h, C.c_void_p(sc_all.data_ptr()), C.c_size_t(n), C.c_void_p(out.data_ptr()),
                                                C.c_void_p(s.cuda_stream)))
                e1.record(s)
                s.synchronize()
                tot, calls = C.c_double(), C.c_uint32()
                _ffi.check(L.h2b_kernel_time_collect(C.byref(tot), C.byref(calls)))
                _ffi.check(L.h2b_set_kernel_timing(0))
                ms = e0.elapsed_time(e1) / reps
                kms = tot.value / max(calls.value, 1)
                print(f"COMMIT 2^{lg} srs_c={cwin}: {ms:.3f} ms/step ({n / ms * 1e3:.3e} pts/s), accumulate {kms:.3f} ms, "
                      f"register {treg:.2f} s", flush=True)
                res[f"commit_{lg}_c{cwin}"] = {"ms": ms, "acc_ms": kms, "register_s": treg}
                _ffi.check(L.h2b_srs_release(h))
            del hb
        _ffi.check(L.h2b_set_srs_precompute(1, 0))
        del bases, seeds, sc_all
        import bn254 as o
        for k in [int(x) for x in os.environ.get("PROBE_NTT", "10,14,16,18,20,22,24").split(",") if x not in ("", "none")]:
            n = 1 << k
            om = o.fr_array([pow(o.ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD)])[0]
            a_t = torch.from_numpy(rand_fr_np(n, 3).view(np.int64)).cuda()
            for _ in range(2):
                arithmetic.dev_best_fft(a_t, om, k, stream=s)
            s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record(s)
            for _ in range(reps):
                arithmetic.dev_best_fft(a_t, om, k, stream=s)
            e1.record(s)
            s.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"NTT 2^{k}: {ms:.4f} ms ({n / ms * 1e3:.3e} elems/s, {n * k / 2 / ms * 1e-6:.2f} G modmul/s)", flush=True)
            res[f"ntt_{k}"] = ms
            del a_t
        # the EvaluationDomain transforms, device-resident (PROBE_DOMAIN = list of k; j = 5, extended_k = k + 2)
        for k in [int(x) for x in os.environ.get("PROBE_DOMAIN", "").split(",") if x]:
            d = h2b.EvaluationDomain(5, k)
            ek = d.extended_k
            a_t = torch.from_numpy(rand_fr_np(1 << k, 5).view(np.int64)).cuda()
            e_t = torch.empty((1 << ek, 4), dtype=torch.int64, device="cuda")
            c_t = torch.empty((d.extended_len() if hasattr(d, "extended_len") else 1 << ek, 4), dtype=torch.int64, device="cuda")
            fns = {"lagrange_to_coeff": lambda: d.dev_lagrange_to_coeff(a_t, stream=s),
                   "coeff_to_extended": lambda: d.dev_coeff_to_extended(a_t, e_t, stream=s),
                   "extended_to_coeff": lambda: d.dev_extended_to_coeff(e_t, c_t, stream=s)}
            for name, fn in fns.items():
                for _ in range(2):
                    fn()
                s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record(s)
                for _ in range(reps):
                    fn()
                e1.record(s)
                s.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"DOMAIN k={k} ext={ek} {name}: {ms:.4f} ms", flush=True)
                res[f"{name}_{k}"] = ms
            del a_t, e_t, c_t
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
