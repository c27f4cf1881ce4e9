"""evaluate_h: the oracle's restatement pinned against the reference's execution, and
the CUDA path against both.

In the recorded arithmetic-circuit proof (k = 4, extended_k = 5) the recorded best_fft calls of keygen_pk and
create_proof are, in order: 0-4 fixed lagrange_to_coeff (sm, sl, sr, so, sc), 5-9 their coeff_to_extended,
10-13 / 14-17 the four permutation polynomials, 18-23 l0, l_blind, l_last (each lagrange_to_coeff then
coeff_to_extended), 24 instance, 25-32 the four permutation products z_i (lagrange_to_coeff, coeff_to_extended),
33-35 advice l, r, o, 36-39 coeff_to_extended of advice and instance inside evaluate_h, 40 extended_to_coeff of
the quotient.  So every input of evaluate_h is a recorded output, and its result is input 40 divided by the
t_evaluations."""
import json

import numpy as np
import pytest

from util import GOLDEN

MANIFEST = json.load(open(f"{GOLDEN}/wasm_manifest.json"))
K, EXT_K = 4, 5
N, EN = 1 << K, 1 << EXT_K


def _fixture(spec):
    ent = MANIFEST["arithmetic"]
    z = np.load(f"{GOLDEN}/{ent['file']}")
    R = spec.R_MOD
    out = lambda i: spec.fr_ints(z[f"fft{i}_out"])
    inp = lambda i: spec.fr_ints(z[f"fft{i}_in"])
    ext_omega = pow(spec.ROOT_OF_UNITY, 1 << (28 - EXT_K), R)
    # the coset generator: recorded coset evaluations are p(zeta * omega^i); check it on l0 = L_0
    zeta = spec.fr_ints(z["fft36_in"][1:2])[0] * pow(spec.fr_ints(z["fft33_out"][1:2])[0] * pow(N, -1, R), -1, R) % R
    d = {
        "fixed": [out(i) for i in range(5, 10)], "sigma": [out(i) for i in range(14, 18)],
        "l0": out(19), "l_blind": out(21), "l_last": out(23), "z": [out(i) for i in (26, 28, 30, 32)],
        "advice": [out(i) for i in (36, 37, 38)], "instance": [out(39)], "ext_omega": ext_omega, "zeta": zeta,
    }
    d["l_active"] = [(1 - a - b) % R for a, b in zip(d["l_last"], d["l_blind"])]
    tev = [pow((pow(zeta * pow(ext_omega, i, R) % R, N, R) - 1) % R, -1, R) for i in range(EN // N)]
    d["values"] = [v * pow(tev[i % len(tev)], -1, R) % R for i, v in enumerate(inp(40))]   # undo divide_by_vanishing_poly
    return d


def _gate_graph(ev, spec):
    """The compiled "plonk" gate of circuits/src/arithmetic_circuit.rs:205-217:
    l*sl + r*sr + l*r*sm + (o*so*(-1)) + sc with fixed columns created in the order sm, sl, sr, so, sc (:196-200)."""
    g = ev.Graph()
    r0 = g.add_rotation(0)
    l, r, o = (ev.ADVICE, 0, r0), (ev.ADVICE, 1, r0), (ev.ADVICE, 2, r0)
    sm, sl, sr, so, sc = [(ev.FIXED, i, r0) for i in range(5)]
    t = g.add_calc(ev.ADD, g.add_calc(ev.MUL, l, sl), g.add_calc(ev.MUL, r, sr))
    t = g.add_calc(ev.ADD, t, g.add_calc(ev.MUL, g.add_calc(ev.MUL, l, r), sm))
    t = g.add_calc(ev.ADD, t, g.add_calc(ev.MUL, g.add_calc(ev.MUL, o, so), g.add_constant(spec.R_MOD - 1)))
    t = g.add_calc(ev.ADD, t, sc)
    g.add_calc(ev.HORNER, (ev.PREVIOUS, 0, 0), [t], (ev.Y, 0, 0))
    return g


def _solve(A, b, R):
    m, nv = len(A), len(A[0])
    A = [row[:] + [bb] for row, bb in zip(A, b)]
    piv, rr = [], 0
    for c in range(nv):
        p = next((i for i in range(rr, m) if A[i][c] % R), None)
        if p is None:
            piv.append(None)
            continue
        A[rr], A[p] = A[p], A[rr]
        inv = pow(A[rr][c], -1, R)
        A[rr] = [v * inv % R for v in A[rr]]
        for i in range(m):
            if i != rr and A[i][c] % R:
                f = A[i][c]
                A[i] = [(v - f * w) % R for v, w in zip(A[i], A[rr])]
        piv.append(rr)
        rr += 1
    residual = sum(1 for i in range(rr, m) if A[i][nv] % R)
    return [A[p][nv] if p is not None else None for p in piv], rr, residual


def _recover_challenges(d, spec):
    """values[idx] is linear in the 17 monomials y^9..y, beta*y^3..beta, gamma*y^3..gamma (10 Horner terms, the last
    four are A + beta*B + gamma*C): solve the 32 x 17 system and insist on zero residual and on monomials that are
    powers / products of one (y, beta, gamma)."""
    import evaluate_h as ev
    R = spec.R_MOD
    sm, sl, sr, so, sc = d["fixed"]
    adv, z, sig = d["advice"], d["z"], d["sigma"]
    cols = adv + d["instance"]
    rows, rhs = [], []
    for idx in range(EN):
        l, r, o = adv[0][idx], adv[1][idx], adv[2][idx]
        gate = (l * sl[idx] + r * sr[idx] + l * r * sm[idx] - o * so[idx] + sc[idx]) % R
        r_next, r_last = (idx + 2) % EN, (idx - 12) % EN
        terms = [gate, (1 - z[0][idx]) * d["l0"][idx], (z[3][idx] ** 2 - z[3][idx]) * d["l_last"][idx]]
        terms += [(z[i][idx] - z[i - 1][r_last]) * d["l0"][idx] for i in (1, 2, 3)]
        X = d["zeta"] * pow(d["ext_omega"], idx, R) % R
        row, const = [t % R for t in terms] + [0] * 11, 0
        for i in range(4):
            zn, zz, c, la = z[i][r_next], z[i][idx], cols[i][idx], d["l_active"][idx]
            A, B, Cc = (zn - zz) * c * la, (zn * sig[i][idx] - zz * pow(ev.DELTA, i, R) * X) * la, (zn - zz) * la
            if i < 3:
                row[6 + i] = (row[6 + i] + A) % R      # y^(3-i) sits at index 9 - (3 - i)
            else:
                const = A % R
            row[9 + i], row[13 + i] = B % R, Cc % R
        rows.append(row)
        rhs.append((d["values"][idx] - const) % R)
    sol, rank, residual = _solve(rows, rhs, R)
    return sol, rank, residual


def test_oracle_evaluate_h_is_pinned_by_the_reference_execution(spec):
    import evaluate_h as ev
    R = spec.R_MOD
    d = _fixture(spec)
    # the labelling itself: l0 is L_0 on the coset zeta * <extended_omega>, the gate vanishes on H
    for idx in range(EN):
        X = d["zeta"] * pow(d["ext_omega"], idx, R) % R
        assert d["l0"][idx] == (pow(X, N, R) - 1) * pow(N * (X - 1) % R, -1, R) % R
    assert pow(d["zeta"], 3, R) == 1 and d["zeta"] != 1
    sol, rank, residual = _recover_challenges(d, spec)
    assert rank == 17 and residual == 0, "the recorded quotient is not consistent with this restatement"
    y, beta, gamma = sol[8], sol[12], sol[16]
    assert all(sol[9 - p] == pow(y, p, R) for p in range(1, 10))
    assert all(sol[9 + i] == beta * pow(y, 3 - i, R) % R and sol[13 + i] == gamma * pow(y, 3 - i, R) % R for i in range(4))
    perm = ev.Permutation(columns=[(ev.ADVICE, 0), (ev.ADVICE, 1), (ev.ADVICE, 2), (ev.INSTANCE, 0)], sigma_cosets=d["sigma"],
                          z_cosets=d["z"], chunk_len=1, last_rotation=-6, l0=d["l0"], l_last=d["l_last"], l_active_row=d["l_active"])
    sc = ev.Scalars(challenges=[], beta=beta, gamma=gamma, theta=0, y=y)
    got = ev.evaluate_h(_gate_graph(ev, spec), d["fixed"], d["advice"], d["instance"], sc, perm, K, EXT_K, d["ext_omega"], d["zeta"])
    assert got == d["values"]
    # and only this restatement: another rotation of z_{i-1} or no coset generator in delta leaves a residual
    perm_bad = ev.Permutation(**{**perm.__dict__, "last_rotation": -5})
    assert ev.evaluate_h(_gate_graph(ev, spec), d["fixed"], d["advice"], d["instance"], sc, perm_bad, K, EXT_K, d["ext_omega"],
                         d["zeta"]) != d["values"]


def _to_product_graph(g, spec, evaluation):
    return evaluation.GraphEvaluator(constants=spec.fr_array(g.constants) if g.constants else np.zeros((0, 4), dtype=np.uint64),
                                     rotations=list(g.rotations), calculations=list(g.calcs), num_intermediates=g.num_intermediates)


def _run_gpu(h2b, spec, evaluation, g, fixed, advice, instance, sc, perm, k, j):
    import torch
    up = lambda col: torch.from_numpy(spec.fr_array(col).view(np.int64)).cuda()
    dom = h2b.EvaluationDomain(j, k)
    f_t, a_t, i_t = [up(c) for c in fixed], [up(c) for c in advice], [up(c) for c in instance]
    pd = None
    if perm is not None:
        pd = evaluation.PermutationData(columns=perm.columns, sigma_cosets=[up(c) for c in perm.sigma_cosets],
                                        z_cosets=[up(c) for c in perm.z_cosets], chunk_len=perm.chunk_len,
                                        last_rotation=perm.last_rotation, l0=up(perm.l0), l_last=up(perm.l_last),
                                        l_active_row=up(perm.l_active_row))
    values = torch.empty((1 << dom.extended_k, 4), dtype=torch.int64, device="cuda")
    fr1 = lambda v: spec.fr_array([v])[0]
    ch = spec.fr_array(list(sc.challenges)) if len(sc.challenges) else np.zeros((0, 4), dtype=np.uint64)
    evaluation.dev_evaluate_h(dom, _to_product_graph(g, spec, evaluation), f_t, a_t, i_t, ch, fr1(sc.beta), fr1(sc.gamma),
                              fr1(sc.theta), fr1(sc.y), pd, values)
    torch.cuda.synchronize()
    return spec.fr_ints(values.cpu().numpy().view(np.uint64)), dom


@pytest.mark.gpu
def test_cuda_evaluate_h_reproduces_the_reference_quotient(h2b, spec):
    """h2b_dev_evaluate_h on the recorded cosets, with the challenges recovered from the record, gives the recorded
    quotient numerator row for row; dividing by the vanishing polynomial and extended_to_coeff on the device then
    reproduces the recorded best_fft input / output of the reference's own extended_to_coeff call."""
    import evaluate_h as ev
    from halo2_prover_b200 import evaluation
    d = _fixture(spec)
    sol, rank, residual = _recover_challenges(d, spec)
    assert residual == 0
    sc = ev.Scalars(challenges=[], beta=sol[12], gamma=sol[16], theta=0, y=sol[8])
    perm = ev.Permutation(columns=[(ev.ADVICE, 0), (ev.ADVICE, 1), (ev.ADVICE, 2), (ev.INSTANCE, 0)], sigma_cosets=d["sigma"],
                          z_cosets=d["z"], chunk_len=1, last_rotation=-6, l0=d["l0"], l_last=d["l_last"], l_active_row=d["l_active"])
    got, dom = _run_gpu(h2b, spec, evaluation, _gate_graph(ev, spec), d["fixed"], d["advice"], d["instance"], sc, perm, K, 3)
    assert dom.extended_k == EXT_K
    assert got == d["values"]
    z = np.load(f"{GOLDEN}/{MANIFEST['arithmetic']['file']}")
    h = dom.divide_by_vanishing_poly(spec.fr_array(got))
    assert (h == z["fft40_in"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2])
def test_cuda_evaluate_h_random_graph_vs_oracle(h2b, spec, seed):
    """Random expression graphs over every ValueSource and Calculation kind, random columns, rotations in both
    directions, two permutation chunks of unequal length: CUDA == oracle."""
    import random
    import evaluate_h as ev
    from halo2_prover_b200 import evaluation
    rng = random.Random(seed)
    R = spec.R_MOD
    k, j = 5, 4
    ext_k, en = 7, 128
    rnd_col = lambda: [rng.randrange(R) for _ in range(en)]
    fixed, advice, instance = [rnd_col() for _ in range(3)], [rnd_col() for _ in range(4)], [rnd_col() for _ in range(2)]
    g = ev.Graph()
    rots = [g.add_rotation(r) for r in (0, 1, -1, 3, -7)]
    leaves = [g.add_constant(rng.randrange(R)) for _ in range(3)]
    leaves += [(ev.FIXED, c, rng.choice(rots)) for c in range(3)] + [(ev.ADVICE, c, rng.choice(rots)) for c in range(4)]
    leaves += [(ev.INSTANCE, c, rng.choice(rots)) for c in range(2)]
    leaves += [(ev.CHALLENGE, 0, 0), (ev.CHALLENGE, 1, 0), (ev.BETA, 0, 0), (ev.GAMMA, 0, 0), (ev.THETA, 0, 0), (ev.Y, 0, 0)]
    nodes = list(leaves)
    for _ in range(40):
        kind = rng.choice([ev.ADD, ev.SUB, ev.MUL, ev.SQUARE, ev.DOUBLE, ev.NEGATE, ev.STORE, ev.HORNER])
        if kind in (ev.ADD, ev.SUB, ev.MUL):
            nodes.append(g.add_calc(kind, rng.choice(nodes), rng.choice(nodes)))
        elif kind == ev.HORNER:
            nodes.append(g.add_calc(kind, rng.choice(nodes), [rng.choice(nodes) for _ in range(rng.randrange(0, 4))], rng.choice(nodes)))
        else:
            nodes.append(g.add_calc(kind, rng.choice(nodes)))
    g.add_calc(ev.HORNER, (ev.PREVIOUS, 0, 0), nodes[-3:], (ev.Y, 0, 0))
    sc = ev.Scalars(challenges=[rng.randrange(R), rng.randrange(R)], beta=rng.randrange(R), gamma=rng.randrange(R),
                    theta=rng.randrange(R), y=rng.randrange(R))
    perm = ev.Permutation(columns=[(ev.ADVICE, 1), (ev.FIXED, 2), (ev.INSTANCE, 0), (ev.ADVICE, 3), (ev.ADVICE, 0)],
                          sigma_cosets=[rnd_col() for _ in range(5)], z_cosets=[rnd_col() for _ in range(3)], chunk_len=2,
                          last_rotation=-6, l0=rnd_col(), l_last=rnd_col(), l_active_row=rnd_col())
    dom_probe = h2b.EvaluationDomain(j, k)
    assert dom_probe.extended_k == ext_k
    ext_omega = spec.fr_ints(dom_probe.get_extended_omega().reshape(1, 4))[0]
    zeta = spec.fr_ints(dom_probe.g_coset.reshape(1, 4))[0]
    want = ev.evaluate_h(g, fixed, advice, instance, sc, perm, k, ext_k, ext_omega, zeta)
    got, _ = _run_gpu(h2b, spec, evaluation, g, fixed, advice, instance, sc, perm, k, j)
    assert got == want
    # no permutation argument, empty graph
    assert _run_gpu(h2b, spec, evaluation, ev.Graph(), fixed, advice, instance, sc, None, k, j)[0] == [0] * en
    # a lookup argument folded in afterwards (parity not pinned on a reference record: no reference circuit has one)
    import torch
    lg = ev.Graph()
    r0, r1 = lg.add_rotation(0), lg.add_rotation(-1)
    cin = lg.add_calc(ev.HORNER, lg.add_constant(0), [(ev.ADVICE, 0, r0), (ev.ADVICE, 2, r1)], (ev.THETA, 0, 0))
    ctab = lg.add_calc(ev.HORNER, lg.add_constant(0), [(ev.FIXED, 0, r0), (ev.FIXED, 1, r0)], (ev.THETA, 0, 0))
    lg.add_calc(ev.MUL, lg.add_calc(ev.ADD, cin, (ev.BETA, 0, 0)), lg.add_calc(ev.ADD, ctab, (ev.GAMMA, 0, 0)))
    lk = ev.Lookup(graph=lg, product_coset=rnd_col(), permuted_input_coset=rnd_col(), permuted_table_coset=rnd_col())
    want2 = ev.fold_lookup(want, lk, fixed, advice, instance, sc, perm.l0, perm.l_last, perm.l_active_row, k, ext_k)
    up = lambda col: torch.from_numpy(spec.fr_array(col).view(np.int64)).cuda()
    dom = h2b.EvaluationDomain(j, k)
    values = up(want)
    fr1 = lambda v: spec.fr_array([v])[0]
    evaluation.dev_evaluate_h_lookup(dom, _to_product_graph(lg, spec, evaluation), [up(c) for c in fixed], [up(c) for c in advice],
                                     [up(c) for c in instance], spec.fr_array(list(sc.challenges)), fr1(sc.beta), fr1(sc.gamma),
                                     fr1(sc.theta), fr1(sc.y), up(perm.l0), up(perm.l_last), up(perm.l_active_row),
                                     up(lk.product_coset), up(lk.permuted_input_coset), up(lk.permuted_table_coset), values)
    torch.cuda.synchronize()
    assert spec.fr_ints(values.cpu().numpy().view(np.uint64)) == want2


@pytest.mark.gpu
def test_cuda_resident_quotient_pipeline_reproduces_the_h_commitments(h2b, spec, href):
    """The whole quotient step with nothing leaving HBM: coefficient polynomials of the recorded proof ->
    coeff_to_extended (batched) -> evaluate_h -> divide_by_vanishing_poly -> extended_to_coeff -> commit of the h
    pieces.  The h pieces equal the scalars of the reference's two h commitments (MSM records 17, 18) and the
    commitments are the bytes at words 8 and 9 of the reference's proof."""
    import torch
    import evaluate_h as ev
    from halo2_prover_b200 import evaluation
    d = _fixture(spec)
    sol, _, residual = _recover_challenges(d, spec)
    assert residual == 0
    z = np.load(f"{GOLDEN}/{MANIFEST['arithmetic']['file']}")
    R = spec.R_MOD
    n_inv = pow(N, -1, R)
    up = lambda col: torch.from_numpy(spec.fr_array(col).view(np.int64)).cuda()
    dom = h2b.EvaluationDomain(3, K)
    # advice l, r, o and the instance column in coefficient form (recorded lagrange_to_coeff outputs / n), one after the other
    coeffs = np.concatenate([spec.fr_array([v * n_inv % R for v in spec.fr_ints(z[f"fft{i}_out"])]) for i in (33, 34, 35, 24)])
    c_t = torch.from_numpy(coeffs.view(np.int64)).cuda()
    ext_t = torch.empty((4 * EN, 4), dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    dom.dev_coeff_to_extended_many(c_t, ext_t, 4, stream=s)
    cols = [ext_t[i * EN:(i + 1) * EN] for i in range(4)]
    g = _to_product_graph(_gate_graph(ev, spec), spec, evaluation)
    fixed_t = [up(c) for c in d["fixed"]]
    pd = evaluation.PermutationData(columns=[(ev.ADVICE, 0), (ev.ADVICE, 1), (ev.ADVICE, 2), (ev.INSTANCE, 0)],
                                    sigma_cosets=[up(c) for c in d["sigma"]], z_cosets=[up(c) for c in d["z"]], chunk_len=1,
                                    last_rotation=-6, l0=up(d["l0"]), l_last=up(d["l_last"]), l_active_row=up(d["l_active"]))
    torch.cuda.synchronize()
    values = torch.empty((EN, 4), dtype=torch.int64, device="cuda")
    fr1 = lambda v: spec.fr_array([v])[0]
    evaluation.dev_evaluate_h(dom, g, fixed_t, cols[:3], cols[3:], np.zeros((0, 4), dtype=np.uint64), fr1(sol[12]), fr1(sol[16]),
                              fr1(0), fr1(sol[8]), pd, values, stream=s)
    dom.dev_divide_by_vanishing_poly(values, stream=s)
    h_t = torch.empty((N * 2, 4), dtype=torch.int64, device="cuda")
    dom.dev_extended_to_coeff(values, h_t, stream=s)
    params = h2b.ParamsKZG.read(z["params"].tobytes())
    out_t = torch.empty((2, 12), dtype=torch.int64, device="cuda")
    from halo2_prover_b200 import _ffi
    import ctypes as C
    _ffi.check(_ffi.lib().h2b_dev_commit_many(C.c_uint64(params._handles["g"]), C.c_void_p(h_t.data_ptr()), C.c_size_t(N),
                                              C.c_size_t(2), C.c_void_p(out_t.data_ptr()), C.c_void_p(s.cuda_stream)))
    s.synchronize()
    h = h_t.cpu().numpy().view(np.uint64)
    assert (h[:N] == z["msm17_scalars"]).all() and (h[N:] == z["msm18_scalars"]).all()
    enc = h2b.g1_to_bytes(out_t.cpu().numpy().view(np.uint64))
    proof = z["proof"].tobytes()
    assert enc == proof[32 * 8: 32 * 10]
    params.release()


# ---------------------------------------------------------------------------------------------- Collatz (k = 10)
# Second pin: four gates (Horner over several gate polynomials), Rotation::next inside the gates with
# rot_scale = 4, selectors turned into fixed columns, one permutation column in a chunk of length 2.
# Fixture: tests/golden/wasm_collatz_k10_evalh.npz (oracle/wasm/make_evalh_golden.py).
CK, CEXT = 10, 12
CN, CEN = 1 << CK, 1 << CEXT


def _collatz_fixture(spec):
    z = np.load(f"{GOLDEN}/wasm_collatz_k10_evalh.npz")
    R = spec.R_MOD
    A = lambda name: spec.fr_ints(z[name])
    d = {"fixed": [A("fixed0"), A("fixed1")], "sigma": [A("sigma0")], "l0": A("l0"), "l_blind": A("l_blind"), "l_last": A("l_last"),
         "z": [A("z0")], "advice": [A("advice0"), A("advice1"), A("advice2")], "instance": []}
    d["ext_omega"] = pow(spec.ROOT_OF_UNITY, 1 << (28 - CEXT), R)
    # coset generator read off the record: coefficient 1 of the witness column is multiplied by it
    d["zeta"] = spec.fr_ints(z["advice0_ext_in"][1:2])[0] * pow(spec.fr_ints(z["advice0_coeff_fft"][1:2])[0] * pow(CN, -1, R), -1, R) % R
    d["l_active"] = [(1 - a - b) % R for a, b in zip(d["l_last"], d["l_blind"])]
    tev = [pow((pow(d["zeta"] * pow(d["ext_omega"], i, R) % R, CN, R) - 1) % R, -1, R) for i in range(CEN // CN)]
    d["values"] = [v * pow(tev[i % len(tev)], -1, R) % R for i, v in enumerate(A("quotient_in"))]
    d["quotient_in"] = z["quotient_in"]
    return d


def _collatz_graph(ev, spec):
    """circuits/src/collatz.rs:36-82: is_even, is_odd, is_one, final_element, in that order; the selectors become the
    fixed columns 1 (`selector`) and 0 (`final_entry`) (the only assignment consistent with the record)."""
    g = ev.Graph()
    cur, nxt = g.add_rotation(0), g.add_rotation(1)
    x, y = (ev.ADVICE, 0, cur), (ev.ADVICE, 0, nxt)
    is_odd, is_one = (ev.ADVICE, 1, cur), (ev.ADVICE, 2, cur)
    sel, fin = (ev.FIXED, 1, cur), (ev.FIXED, 0, cur)
    one, two, three = g.add_constant(1), g.add_constant(2), g.add_constant(3)
    c = g.add_calc
    g1 = c(ev.MUL, sel, c(ev.MUL, c(ev.SUB, one, is_odd), c(ev.SUB, x, c(ev.MUL, two, y))))
    g2 = c(ev.MUL, c(ev.MUL, sel, c(ev.SUB, one, is_one)), c(ev.MUL, is_odd, c(ev.SUB, c(ev.ADD, c(ev.MUL, three, x), one), y)))
    g3 = c(ev.MUL, c(ev.MUL, sel, is_one), c(ev.ADD, c(ev.SUB, x, y), c(ev.SUB, x, one)))
    g4 = c(ev.MUL, fin, c(ev.SUB, one, x))
    c(ev.HORNER, (ev.PREVIOUS, 0, 0), [g1, g2, g3, g4], (ev.Y, 0, 0))
    return g


def _collatz_challenges(d, spec):
    """7 Horner terms (4 gates, l0 (1 - z), l_last (z^2 - z), the product term A + beta B + gamma C): 8 unknown monomials
    y^6..y, beta, gamma; 96 sampled rows."""
    import random
    R = spec.R_MOD
    adv, fx, zc, sig = d["advice"], d["fixed"], d["z"][0], d["sigma"][0]
    rows, rhs = [], []
    for idx in random.Random(3).sample(range(CEN), 96):
        nx = (idx + 4) % CEN
        x, y, odd, one = adv[0][idx], adv[0][nx], adv[1][idx], adv[2][idx]
        sel, fin = fx[1][idx], fx[0][idx]
        terms = [sel * (1 - odd) * (x - 2 * y), sel * (1 - one) * odd * (3 * x + 1 - y), sel * one * ((x - y) + (x - 1)), fin * (1 - x),
                 (1 - zc[idx]) * d["l0"][idx], (zc[idx] ** 2 - zc[idx]) * d["l_last"][idx]]
        X = d["zeta"] * pow(d["ext_omega"], idx, R) % R
        zn, zz, cval, la = zc[nx], zc[idx], adv[0][idx], d["l_active"][idx]
        rows.append([t % R for t in terms] + [(zn * sig[idx] - zz * X) * la % R, (zn - zz) * la % R])
        rhs.append((d["values"][idx] - (zn - zz) * cval * la) % R)
    return _solve(rows, rhs, R)


def test_oracle_evaluate_h_collatz_pin(spec):
    import evaluate_h as ev
    R = spec.R_MOD
    d = _collatz_fixture(spec)
    assert d["zeta"] == spec.ZETA
    sol, rank, residual = _collatz_challenges(d, spec)
    assert rank == 8 and residual == 0
    y, beta, gamma = sol[5], sol[6], sol[7]
    assert all(sol[6 - p] == pow(y, p, R) for p in range(1, 7))
    perm = ev.Permutation(columns=[(ev.ADVICE, 0)], sigma_cosets=d["sigma"], z_cosets=d["z"], chunk_len=2, last_rotation=-6,
                          l0=d["l0"], l_last=d["l_last"], l_active_row=d["l_active"])
    sc = ev.Scalars(challenges=[], beta=beta, gamma=gamma, theta=0, y=y)
    got = ev.evaluate_h(_collatz_graph(ev, spec), d["fixed"], d["advice"], [], sc, perm, CK, CEXT, d["ext_omega"], d["zeta"])
    assert got == d["values"]   # all 4096 rows, 8 of them used up by the unknowns


@pytest.mark.gpu
def test_cuda_evaluate_h_collatz_reproduces_the_reference_quotient(h2b, spec):
    import evaluate_h as ev
    from halo2_prover_b200 import evaluation
    d = _collatz_fixture(spec)
    sol, rank, residual = _collatz_challenges(d, spec)
    assert residual == 0
    perm = ev.Permutation(columns=[(ev.ADVICE, 0)], sigma_cosets=d["sigma"], z_cosets=d["z"], chunk_len=2, last_rotation=-6,
                          l0=d["l0"], l_last=d["l_last"], l_active_row=d["l_active"])
    sc = ev.Scalars(challenges=[], beta=sol[6], gamma=sol[7], theta=0, y=sol[5])
    got, dom = _run_gpu(h2b, spec, evaluation, _collatz_graph(ev, spec), d["fixed"], d["advice"], [], sc, perm, CK, 4)
    assert dom.extended_k == CEXT
    assert got == d["values"]
    assert (dom.divide_by_vanishing_poly(spec.fr_array(got)) == d["quotient_in"]).all()


@pytest.mark.gpu
def test_cuda_evaluate_h_rejects_malformed_graphs(h2b, spec):
    """The serialised graph is validated before it is trusted with device pointers: out-of-range columns,
    rotations, intermediates or a truncated stream are argument errors (upstream would panic on the index)."""
    import torch
    import evaluate_h as ev
    from halo2_prover_b200 import _ffi, evaluation
    dom = h2b.EvaluationDomain(3, 4)
    en = 1 << dom.extended_k
    col = torch.zeros((en, 4), dtype=torch.int64, device="cuda")
    values = torch.zeros((en, 4), dtype=torch.int64, device="cuda")
    zero = np.zeros(4, dtype=np.uint64)
    none = np.zeros((0, 4), dtype=np.uint64)

    def run(calcs, num_inter=1, rotations=(0,)):
        g = evaluation.GraphEvaluator(rotations=list(rotations), calculations=calcs, num_intermediates=num_inter)
        evaluation.dev_evaluate_h(dom, g, [col], [col], [], none, zero, zero, zero, zero, None, values)

    run([(ev.STORE, 0, (ev.ADVICE, 0, 0))])                          # well-formed
    for bad in ([(ev.STORE, 0, (ev.ADVICE, 1, 0))],                    # advice column 1 of 1
                [(ev.STORE, 0, (ev.FIXED, 0, 1))],                     # rotation index 1 of 1
                [(ev.STORE, 1, (ev.ADVICE, 0, 0))],                    # target beyond num_intermediates
                [(ev.ADD, 0, (ev.INTERMEDIATE, 5, 0), (ev.Y, 0, 0))],  # intermediate 5 of 1
                [(ev.STORE, 0, (ev.INSTANCE, 0, 0))],                  # no instance columns
                [(ev.STORE, 0, (ev.CHALLENGE, 0, 0))]):                # no challenges
        with pytest.raises(_ffi.H2BError):
            run(bad)
    with pytest.raises(_ffi.H2BError):
        run([(ev.STORE, 0, (ev.ADVICE, 0, 0))], rotations=tuple(range(33)))   # more than 32 rotations


# ---------------------------------------------------------------------------------------------- Poseidon (k = 7)
# Third pin: the Pow5 chip (halo2_gadgets, the gates restated from circuits/src/poseidon/pow5.rs:57-201): ten gate
# polynomials of degree up to 6 behind three selector columns, Rotation::prev / next with rot_scale = 8, MDS constants
# inside the expressions, and a permutation argument over SEVEN columns of all three kinds (instance, fixed, advice) in
# two chunks of length cs.degree() - 2 = 4, so the products are genuinely multi-column.
# Fixture: tests/golden/wasm_poseidon_k7.npz, the complete record of that proof.  Order of its best_fft calls in
# keygen_pk + create_proof: 0-8 fixed lagrange_to_coeff (rc_a x3, rc_b x3, s_full, s_partial, s_pad_and_add), 9-17 their
# coeff_to_extended, 18-24 / 25-31 the seven permutation polynomials, 32-37 l0, l_blind, l_last, 38 instance, 39-42 the two
# permutation products, 43-46 advice (state x3, partial_sbox), 47-51 coeff_to_extended of advice + instance inside
# evaluate_h, 52 extended_to_coeff of the quotient.
PK, PEXT = 7, 10
PN, PEN = 1 << PK, 1 << PEXT


def _poseidon_fixture(spec):
    z = np.load(f"{GOLDEN}/{MANIFEST['poseidon']['file']}")
    R = spec.R_MOD
    out = lambda i: spec.fr_ints(z[f"fft{i}_out"])
    assert [z[f"fft{i}_in"].shape[0] for i in range(53)] == [PN] * 9 + [PEN] * 9 + [PN] * 7 + [PEN] * 7 + [PN, PEN] * 3 + \
        [PN, PN, PEN, PN, PEN] + [PN] * 4 + [PEN] * 6
    d = {"fixed": [out(i) for i in range(9, 18)], "fixed_lagrange": [spec.fr_ints(z[f"fft{i}_in"]) for i in range(9)],
         "sigma": [out(i) for i in range(25, 32)], "l0": out(33), "l_blind": out(35), "l_last": out(37),
         "z": [out(40), out(42)], "advice": [out(i) for i in range(47, 51)], "instance": [out(51)]}
    d["ext_omega"] = pow(spec.ROOT_OF_UNITY, 1 << (28 - PEXT), R)
    # coset generator read off the record: coefficient 1 of the first state column is multiplied by it
    d["zeta"] = spec.fr_ints(z["fft47_in"][1:2])[0] * pow(spec.fr_ints(z["fft43_out"][1:2])[0] * pow(PN, -1, R), -1, R) % R
    d["l_active"] = [(1 - a - b) % R for a, b in zip(d["l_last"], d["l_blind"])]
    tev = [pow((pow(d["zeta"] * pow(d["ext_omega"], i, R) % R, PN, R) - 1) % R, -1, R) for i in range(PEN // PN)]
    d["values"] = [v * pow(tev[i % len(tev)], -1, R) % R for i, v in enumerate(spec.fr_ints(z["fft52_in"]))]
    d["quotient_in"] = z["fft52_in"]
    return d


def _poseidon_columns(ev):
    """cs.permutation in enable_equality order: the instance column (poseidon_circuit.rs:70), rc_b[0] through
    enable_constant (:77), then state[0..3] and rc_b[0..3] (pow5.rs:79-84; rc_b[0] is already there)."""
    return [(ev.INSTANCE, 0), (ev.FIXED, 3), (ev.ADVICE, 0), (ev.ADVICE, 1), (ev.ADVICE, 2), (ev.FIXED, 4), (ev.FIXED, 5)]


def _poseidon_gate_polys(spec, fixed, advice, idx, rot_scale, size):
    """The ten gate polynomials at one row, straight from their definition (independent of the graph encoding)."""
    import poseidon_spec as ps
    R = spec.R_MOD
    _, m, minv = _poseidon_constants()
    cur, nxt, prv = idx, (idx + rot_scale) % size, (idx - rot_scale) % size
    st = lambda j, r: advice[j][r]
    sbox = advice[3][cur]
    rc_a = [fixed[j][cur] for j in range(3)]
    rc_b = [fixed[3 + j][cur] for j in range(3)]
    s_full, s_partial, s_pad = fixed[6][cur], fixed[7][cur], fixed[8][cur]
    p5 = lambda v: pow(v, 5, R)
    polys = [s_full * (sum(p5(st(j, cur) + rc_a[j]) * m[i][j] for j in range(3)) - st(i, nxt)) for i in range(3)]
    mid = lambda i: sbox * m[i][0] + sum((st(j, cur) + rc_a[j]) * m[i][j] for j in (1, 2))
    nx = lambda i: sum(st(j, nxt) * minv[i][j] for j in range(3))
    polys.append(s_partial * (p5(st(0, cur) + rc_a[0]) - sbox))
    polys.append(s_partial * (p5(mid(0) + rc_b[0]) - nx(0)))
    polys += [s_partial * (mid(i) + rc_b[i] - nx(i)) for i in (1, 2)]
    polys += [s_pad * (st(i, prv) + st(i, cur) - st(i, nxt)) for i in (0, 1)]
    polys.append(s_pad * (st(2, prv) - st(2, nxt)))
    assert ps.R == R
    return [p % R for p in polys]


_P_CONST = []


def _poseidon_constants():
    import poseidon_spec as ps
    if not _P_CONST:
        _P_CONST.append(ps.generate_constants(3))
    return _P_CONST[0]


def _poseidon_graph(ev, spec):
    """The same ten polynomials as a GraphEvaluator program: one Horner over them with y."""
    _, m, minv = _poseidon_constants()
    g = ev.Graph()
    cur, nxt, prv = g.add_rotation(0), g.add_rotation(1), g.add_rotation(-1)
    c = g.add_calc
    st = lambda j, r: (ev.ADVICE, j, r)
    sbox = (ev.ADVICE, 3, cur)
    rc_a = [(ev.FIXED, j, cur) for j in range(3)]
    rc_b = [(ev.FIXED, 3 + j, cur) for j in range(3)]
    s_full, s_partial, s_pad = [(ev.FIXED, 6 + j, cur) for j in range(3)]
    M = [[g.add_constant(v) for v in row] for row in m]
    MI = [[g.add_constant(v) for v in row] for row in minv]

    def p5(v):
        v2 = c(ev.SQUARE, v)
        return c(ev.MUL, c(ev.SQUARE, v2), v)

    def total(terms):
        acc = terms[0]
        for t in terms[1:]:
            acc = c(ev.ADD, acc, t)
        return acc

    pows = [p5(c(ev.ADD, st(j, cur), rc_a[j])) for j in range(3)]
    polys = [c(ev.MUL, s_full, c(ev.SUB, total([c(ev.MUL, pows[j], M[i][j]) for j in range(3)]), st(i, nxt))) for i in range(3)]
    lin = [c(ev.ADD, st(j, cur), rc_a[j]) for j in range(3)]
    mid = lambda i: total([c(ev.MUL, sbox, M[i][0])] + [c(ev.MUL, lin[j], M[i][j]) for j in (1, 2)])
    nx = lambda i: total([c(ev.MUL, st(j, nxt), MI[i][j]) for j in range(3)])
    polys.append(c(ev.MUL, s_partial, c(ev.SUB, pows[0], sbox)))
    polys.append(c(ev.MUL, s_partial, c(ev.SUB, p5(c(ev.ADD, mid(0), rc_b[0])), nx(0))))
    polys += [c(ev.MUL, s_partial, c(ev.SUB, c(ev.ADD, mid(i), rc_b[i]), nx(i))) for i in (1, 2)]
    polys += [c(ev.MUL, s_pad, c(ev.SUB, c(ev.ADD, st(i, prv), st(i, cur)), st(i, nxt))) for i in (0, 1)]
    polys.append(c(ev.MUL, s_pad, c(ev.SUB, st(2, prv), st(2, nxt))))
    c(ev.HORNER, (ev.PREVIOUS, 0, 0), polys, (ev.Y, 0, 0))
    return g


def _bivariate_product(factors, R):
    """prod (c + beta s + gamma) as {(a, b): coefficient of beta^a gamma^b}"""
    poly = {(0, 0): 1}
    for cst, s in factors:
        nxt = {}
        for (a, b), v in poly.items():
            for key, w in (((a, b), cst), ((a + 1, b), s), ((a, b + 1), 1)):
                nxt[key] = (nxt.get(key, 0) + v * w) % R
        poly = nxt
    return poly


def _poseidon_challenges(d, spec, ev):
    """values[idx] = sum_{i<13} y^(14-i) T_i + y C_0(beta, gamma) + C_1(beta, gamma) with the 13 simple Horner terms T (10
    gates, l0 (1 - z0), l_last (z1^2 - z1), l0 (z1 - z0(last))) and the two chunk terms, polynomials of degree 4 and 3 in
    (beta, gamma): linear in 13 + 15 + 9 = 37 monomials (the constant of C_1 is known).  Solved on 200 sampled rows."""
    import random
    R = spec.R_MOD
    cols = _poseidon_columns(ev)
    col_of = {ev.ADVICE: d["advice"], ev.FIXED: d["fixed"], ev.INSTANCE: d["instance"]}
    mono4 = sorted((a, b) for a in range(5) for b in range(5) if a + b <= 4)
    mono3 = sorted((a, b) for a in range(4) for b in range(4) if 0 < a + b <= 3)
    rows, rhs = [], []
    for idx in random.Random(11).sample(range(PEN), 200):
        nx, last = (idx + 8) % PEN, (idx - 6 * 8) % PEN
        z0, z1 = d["z"]
        terms = _poseidon_gate_polys(spec, d["fixed"], d["advice"], idx, 8, PEN)
        terms += [(1 - z0[idx]) * d["l0"][idx], (z1[idx] ** 2 - z1[idx]) * d["l_last"][idx], (z1[idx] - z0[last]) * d["l0"][idx]]
        X = d["zeta"] * pow(d["ext_omega"], idx, R) % R
        chunk = []
        for ci, zc in enumerate(d["z"]):
            part = list(enumerate(cols))[ci * 4:(ci + 1) * 4]
            left = _bivariate_product([(col_of[kd][ix][idx], d["sigma"][j][idx]) for j, (kd, ix) in part], R)
            right = _bivariate_product([(col_of[kd][ix][idx], pow(ev.DELTA, j, R) * X % R) for j, (kd, ix) in part], R)
            chunk.append({key: (zc[nx] * left.get(key, 0) - zc[idx] * right.get(key, 0)) * d["l_active"][idx] % R
                          for key in set(left) | set(right)})
        rows.append([t % R for t in terms] + [chunk[0].get(mn, 0) for mn in mono4] + [chunk[1].get(mn, 0) for mn in mono3])
        rhs.append((d["values"][idx] - chunk[1].get((0, 0), 0)) % R)
    sol, rank, residual = _solve(rows, rhs, R)
    y = sol[13 + mono4.index((0, 0))]
    beta, gamma = sol[28 + mono3.index((1, 0))], sol[28 + mono3.index((0, 1))]
    consistent = all(sol[i] == pow(y, 14 - i, R) for i in range(13)) and \
        all(sol[13 + i] == y * pow(beta, a, R) * pow(gamma, b, R) % R for i, (a, b) in enumerate(mono4)) and \
        all(sol[28 + i] == pow(beta, a, R) * pow(gamma, b, R) % R for i, (a, b) in enumerate(mono3))
    return (y, beta, gamma), rank, residual, consistent


def _poseidon_perm(ev, d):
    return ev.Permutation(columns=_poseidon_columns(ev), sigma_cosets=d["sigma"], z_cosets=d["z"], chunk_len=4, last_rotation=-6,
                          l0=d["l0"], l_last=d["l_last"], l_active_row=d["l_active"])


def test_poseidon_fixed_columns_hold_the_generated_round_constants(spec):
    """oracle/poseidon_spec.py against the record: the public output for the recorded input, and the proving key's rc_a /
    rc_b columns (full rounds: one round per row; partial rounds: rc_a = round 2i, rc_b = round 2i + 1)."""
    import poseidon_spec as ps
    ent = MANIFEST["poseidon"]
    inp = json.loads(ent["input"])
    assert ps.hash_constant_length(inp["x"]) == int(inp["output"], 16)
    rc, m, minv = _poseidon_constants()
    assert all(sum(m[i][k] * minv[k][j] for k in range(3)) % spec.R_MOD == int(i == j) for i in range(3) for j in range(3))
    fx = _poseidon_fixture(spec)["fixed_lagrange"]
    full = [r for r in range(PN) if fx[6][r]]
    partial = [r for r in range(PN) if fx[7][r]]
    assert len(full) == 8 and len(partial) == 30 and sum(fx[8]) == 1
    rounds = iter(range(68))
    for row in full[:4]:
        assert [fx[j][row] for j in range(3)] == rc[next(rounds)]
    for row in partial:
        assert [fx[j][row] for j in range(3)] == rc[next(rounds)]
        assert [fx[3 + j][row] for j in range(3)] == rc[next(rounds)]
    for row in full[4:]:
        assert [fx[j][row] for j in range(3)] == rc[next(rounds)]


def test_oracle_evaluate_h_poseidon_pin(spec):
    import evaluate_h as ev
    d = _poseidon_fixture(spec)
    assert d["zeta"] == spec.ZETA
    (y, beta, gamma), rank, residual, consistent = _poseidon_challenges(d, spec, ev)
    assert rank == 37 and residual == 0 and consistent
    sc = ev.Scalars(challenges=[], beta=beta, gamma=gamma, theta=0, y=y)
    got = ev.evaluate_h(_poseidon_graph(ev, spec), d["fixed"], d["advice"], d["instance"], sc, _poseidon_perm(ev, d), PK, PEXT,
                        d["ext_omega"], d["zeta"])
    assert got == d["values"]   # all 1024 rows, 37 of them used up by the unknowns


@pytest.mark.gpu
def test_cuda_evaluate_h_poseidon_reproduces_the_reference_quotient(h2b, spec):
    import evaluate_h as ev
    from halo2_prover_b200 import evaluation
    d = _poseidon_fixture(spec)
    (y, beta, gamma), rank, residual, consistent = _poseidon_challenges(d, spec, ev)
    assert residual == 0 and consistent
    sc = ev.Scalars(challenges=[], beta=beta, gamma=gamma, theta=0, y=y)
    got, dom = _run_gpu(h2b, spec, evaluation, _poseidon_graph(ev, spec), d["fixed"], d["advice"], d["instance"], sc,
                        _poseidon_perm(ev, d), PK, 6)
    assert dom.extended_k == PEXT
    assert got == d["values"]
    assert (dom.divide_by_vanishing_poly(spec.fr_array(got)) == d["quotient_in"]).all()
