"""Generates tests/golden/spec_vectors.json from the big-integer specification oracle
(oracle/bn254.py).  Deterministic; re-run to regenerate:  python tests/golden/make_golden.py

These vectors pin the *layout and semantics* at the FFI boundary (Montgomery limbs in,
Montgomery limbs out) independently of the C restatement and of the CUDA code.  Vectors
captured from the reference's own compiled prover live beside this file as wasm_*.json
(see oracle/wasm/).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import bn254 as o  # noqa: E402


def hx(arr):
    """u64 limbs, 16 hex digits each, concatenated."""
    return "".join(format(int(v), "016x") for v in arr.reshape(-1))


def main():
    out = {"ntt": [], "domain": [], "msm": []}
    for k in (0, 1, 2, 3, 6, 9):
        omega = pow(o.ROOT_OF_UNITY, 1 << (o.FR_S - k), o.R_MOD)
        a = o.random_fr(1 << k, 0x4E5454 + k)
        out["ntt"].append({
            "log_n": k, "omega": hx(o.fr_array([omega])), "a": hx(o.fr_array(a)),
            "out": hx(o.fr_array(o.best_fft(a, omega, k))),
        })
    for j, k in ((3, 3), (4, 4), (6, 5), (2, 4), (5, 3)):
        d = o.EvaluationDomain(j, k)
        a = o.random_fr(1 << k, 0x444F4D + 16 * j + k)
        ext = d.coeff_to_extended(a)
        e = o.random_fr(d.extended_len(), 0x455854 + 16 * j + k)
        out["domain"].append({
            "j": j, "k": k, "extended_k": d.extended_k,
            "omega": hx(o.fr_array([d.omega])), "omega_inv": hx(o.fr_array([d.omega_inv])),
            "extended_omega": hx(o.fr_array([d.extended_omega])),
            "extended_omega_inv": hx(o.fr_array([d.extended_omega_inv])),
            "g_coset": hx(o.fr_array([d.g_coset])), "g_coset_inv": hx(o.fr_array([d.g_coset_inv])),
            "ifft_divisor": hx(o.fr_array([d.ifft_divisor])),
            "extended_ifft_divisor": hx(o.fr_array([d.extended_ifft_divisor])),
            "t_evaluations": hx(o.fr_array(d.t_evaluations)),
            "a": hx(o.fr_array(a)),
            "lagrange_to_coeff": hx(o.fr_array(d.lagrange_to_coeff(a))),
            "coeff_to_extended": hx(o.fr_array(ext)),
            "e": hx(o.fr_array(e)),
            "extended_to_coeff": hx(o.fr_array(d.extended_to_coeff(e))),
            "divide_by_vanishing_poly": hx(o.fr_array(d.divide_by_vanishing_poly(e))),
        })
    for n in (1, 2, 3, 4, 5, 31, 32, 33, 64):
        sc = o.random_fr(n, 0x4D534D + n)
        pts = o.random_g1(n, 0x505453 + n)
        if n >= 4:  # structure: zero scalar, r-1, identity base, duplicate point, P and -P
            sc[0] = 0
            sc[1] = o.R_MOD - 1
            pts[2] = None
            pts[3] = pts[1]
        if n >= 32:
            pts[5] = o.g1_neg(pts[4])
            sc[5] = sc[4]
        res = o.msm_naive(sc, pts)
        out["msm"].append({
            "n": n, "scalars": hx(o.fr_array(sc)), "bases": hx(o.affine_to_array(pts)),
            "affine": hx(o.affine_to_array([res])),
        })
    with open(os.path.join(HERE, "spec_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote spec_vectors.json", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
