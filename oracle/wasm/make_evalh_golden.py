#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Extracts, from the recorded execution of the reference's Collatz proof
(/tmp/wasm_collatz_k10.bin, written by make_wasm_golden.py), exactly the arrays Evaluator::evaluate_h consumed and
produced, and commits them as tests/golden/wasm_collatz_k10_evalh.npz (the full record set is 7 MB; these are 1.4 MB).

Order of the recorded best_fft calls during keygen_pk + create_proof of that circuit (k = 10, extended_k = 12):
0-1 fixed lagrange_to_coeff (the two selector columns), 2-3 their coeff_to_extended, 4-5 the permutation polynomial,
6-11 l0, l_blind, l_last, 12-13 the permutation product z, 14-16 advice lagrange_to_coeff (witness, is_odd, is_one),
17-19 their coeff_to_extended inside evaluate_h, 20 extended_to_coeff of the quotient."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
spec_ = importlib.util.spec_from_file_location("mk", os.path.join(HERE, "make_wasm_golden.py"))
mk = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(mk)


def main():
    src = "/tmp/wasm_collatz_k10.bin"
    if not os.path.exists(src):
        sys.exit("run make_wasm_golden.py first (it executes the reference prover and writes " + src + ")")
    recs, meta = mk.parse(src)
    ff = [r for r in recs if r[0] == "fft"]
    assert [r[1] for r in ff[:21]] == [10, 10, 12, 12, 10, 12, 10, 12, 10, 12, 10, 12, 10, 12, 10, 10, 10, 12, 12, 12, 12]
    out = lambda i: ff[i][4]
    arrays = {
        "fixed0": out(2), "fixed1": out(3), "sigma0": out(5), "l0": out(7), "l_blind": out(9), "l_last": out(11), "z0": out(13),
        "advice0": out(17), "advice1": out(18), "advice2": out(19), "quotient_in": ff[20][3], "quotient_out": out(20),
        "advice0_coeff_fft": out(14),   # best_fft output of the witness column's lagrange_to_coeff (coefficients * n)
        "advice0_ext_in": ff[17][3],    # the scaled, zero-padded coefficients the reference fed to best_fft
    }
    path = os.path.join(ROOT, "tests", "golden", "wasm_collatz_k10_evalh.npz")
    np.savez_compressed(path, **arrays)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
