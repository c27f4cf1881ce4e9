"""Vectors captured from the reference's OWN compiled prover (src/lib/wasm/halo2_prover_bg.wasm executed
under oracle/wasm/wasmrun, inputs/outputs of its best_multiexp = wasm func 347 and best_fft = wasm func 80,
during keygen + create_proof + verify_proof of the reference's circuits; the proofs were accepted by the
reference verifier).  CPU tests replay them against the oracle, GPU tests against the CUDA path."""
import struct

import numpy as np
import pytest

from util import GOLDEN, load_golden

MANIFEST = load_golden("wasm_manifest.json")


def _load(name):
    ent = MANIFEST[name]
    z = np.load(f"{GOLDEN}/{ent['file']}")
    k = ent["k"]
    n = 1 << k
    params = z["params"].tobytes()
    assert struct.unpack_from("<I", params, 0)[0] == k
    g = np.frombuffer(params, dtype=np.uint64, count=8 * n, offset=4).reshape(n, 8).copy()
    gl = np.frombuffer(params, dtype=np.uint64, count=8 * n, offset=4 + 64 * n).reshape(n, 8).copy()
    return ent, z, g, gl


def _bases(ent_msm, z, g, gl):
    i, n = ent_msm["i"], ent_msm["n"]
    if ent_msm["bases"] == "g":
        return g[:n].copy()
    if ent_msm["bases"] == "g_lagrange":
        return gl[:n].copy()
    return np.ascontiguousarray(z[f"msm{i}_bases"])


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_manifest_is_a_verified_reference_proof(name):
    ent = MANIFEST[name]
    assert ent["verified_by_reference_verifier"] is True
    if ent["all_records_committed"]:
        assert ent["msm_calls_total"] == len(ent["msm"]) and ent["fft_calls_total"] == len(ent["fft"])
    else:
        assert 0 < len(ent["msm"]) <= ent["msm_calls_total"] and 0 < len(ent["fft"]) <= ent["fft_calls_total"]
    assert ent["proof_bytes"] > 0


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_oracle_matches_reference_execution(name, href, spec):
    ent, z, g, gl = _load(name)
    for m in ent["msm"]:
        sc = np.ascontiguousarray(z[f"msm{m['i']}_scalars"])
        got = href.g1_to_affine(href.best_multiexp(sc, _bases(m, z, g, gl), 3))
        assert (got == z[f"msm{m['i']}_affine"]).all(), (name, m)
    for f in ent["fft"]:
        i = f["i"]
        got = href.best_fft(np.ascontiguousarray(z[f"fft{i}_in"]), z[f"fft{i}_omega"], f["log_n"], 4)
        assert (got == z[f"fft{i}_out"]).all(), (name, f)
    # the SRS inside the reference's params is a set of curve points in the layout we assume
    for p in spec.array_to_affine(g[:4]) + spec.array_to_affine(gl[:4]):
        assert spec.g1_is_on_curve(p)


def test_spec_matches_reference_execution_small(spec):
    ent, z, g, gl = _load("arithmetic")
    for m in ent["msm"][:6]:
        sc = z[f"msm{m['i']}_scalars"]
        want = spec.array_to_affine(z[f"msm{m['i']}_affine"].reshape(1, 8))[0]
        assert spec.msm_naive(spec.fr_ints(sc), spec.array_to_affine(_bases(m, z, g, gl))) == want
    for f in ent["fft"][:12]:
        i = f["i"]
        om = spec.fr_ints(z[f"fft{i}_omega"].reshape(1, 4))[0]
        assert spec.fr_array(spec.best_fft(spec.fr_ints(z[f"fft{i}_in"]), om, f["log_n"])).tolist() == z[f"fft{i}_out"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_cuda_matches_reference_execution(name, h2b, href):
    ent, z, g, gl = _load(name)
    params = h2b.ParamsKZG(ent["k"], g, gl)
    for m in ent["msm"]:
        sc = np.ascontiguousarray(z[f"msm{m['i']}_scalars"])
        want = z[f"msm{m['i']}_affine"]
        if m["bases"] == "g":
            out = params.commit(sc)                 # ParamsKZG::commit, resident SRS
        elif m["bases"] == "g_lagrange":
            out = params.commit_lagrange(sc)        # ParamsKZG::commit_lagrange
        else:
            out = h2b.best_multiexp(sc, _bases(m, z, g, gl))   # verifier-side MSMKZG::eval
        assert (href.g1_to_affine(out) == want).all(), (name, m)
    for f in ent["fft"]:
        i = f["i"]
        a = np.ascontiguousarray(z[f"fft{i}_in"]).copy()
        h2b.best_fft(a, z[f"fft{i}_omega"], f["log_n"])
        assert (a == z[f"fft{i}_out"]).all(), (name, f)
    params.release()


@pytest.mark.gpu
def test_cuda_domain_transforms_match_reference_fft_records(h2b, href):
    """The reference reaches best_fft through EvaluationDomain; replay recorded calls through the fused
    domain entry points: a record whose omega is the domain's omega_inv is a lagrange_to_coeff call whose
    final output is record_out * 1/2^k."""
    ent, z, g, gl = _load("poseidon")
    k = ent["k"]
    d = h2b.EvaluationDomain(6, k)   # Poseidon pow5: degree 6 -> extended_k = k + 3
    dc = href.domain_new(6, k)
    hit = 0
    for f in ent["fft"]:
        i = f["i"]
        if f["log_n"] == k and (z[f"fft{i}_omega"] == d.get_omega_inv()).all():
            a = np.ascontiguousarray(z[f"fft{i}_in"])
            assert (d.lagrange_to_coeff(a.copy()) == href.lagrange_to_coeff(dc, a)).all()
            div = np.tile(d.ifft_divisor, (1 << k, 1))
            assert (d.lagrange_to_coeff(a.copy()) == href.fr_mul(np.ascontiguousarray(z[f"fft{i}_out"]), div)).all()
            hit += 1
    assert hit > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_params_read_from_reference_setup_bytes(name, h2b, href):
    """ParamsKZG::read on the bytes the reference's own `setup(k)` wrote (SerdeFormat::RawBytes), then the
    recorded commits of the reference's proof through the SRS registered from that buffer -- one by one and
    as one batch per base array."""
    ent, z, g, gl = _load(name)
    params = h2b.ParamsKZG.read(z["params"].tobytes())
    assert params.k == ent["k"]
    batch = {"g": [], "g_lagrange": []}
    for m in ent["msm"]:
        if m["bases"] not in batch or m["n"] != params.n:
            continue
        sc = np.ascontiguousarray(z[f"msm{m['i']}_scalars"])
        out = params.commit(sc) if m["bases"] == "g" else params.commit_lagrange(sc)
        assert (href.g1_to_affine(out) == z[f"msm{m['i']}_affine"]).all(), (name, m)
        batch[m["bases"]].append((sc, z[f"msm{m['i']}_affine"]))
    assert batch["g"] or batch["g_lagrange"]
    for which, items in batch.items():
        if not items:
            continue
        outs = params.commit_many([sc for sc, _ in items]) if which == "g" else params.commit_lagrange_many([sc for sc, _ in items])
        for out, (_, want) in zip(outs, items):
            assert (href.g1_to_affine(out) == want).all(), (name, which)
    # ParamsKZG::write: both arrays read back from HBM reproduce the reference's bytes
    assert params.write() == z["params"].tobytes()
    params.release()
    for bad in (b"", b"\x04\x00\x00", z["params"].tobytes()[:-1]):
        with pytest.raises(Exception):
            h2b.ParamsKZG.read(bad)


def _compress(spec, aff8):
    """G1Affine::to_bytes restated: x little-endian, bit 6 of byte 31 = y & 1, identity = zeros."""
    p = spec.array_to_affine(np.asarray(aff8).reshape(1, 8))[0]
    if p is None:
        return bytes(32)
    b = bytearray(p[0].to_bytes(32, "little"))
    b[31] |= (p[1] & 1) << 6
    return bytes(b)


# (circuit, first MSM record, number of records, first 32-byte chunk of the proof they fill): the commitments
# create_proof writes before the evaluations, and the opening-proof points after them
PROOF_POINT_RUNS = [("arithmetic", 9, 10, 0), ("arithmetic", 19, 3, 34), ("poseidon", 16, 12, 0), ("poseidon", 28, 4, 44),
                    ("collatz", 3, 8, 0), ("collatz", 11, 2, 18)]  # collatz: SHPLONK, 10 points + 10 scalars = 640 bytes


@pytest.mark.parametrize("name,first,count,chunk", PROOF_POINT_RUNS)
def test_reference_proof_bytes_are_the_recorded_commitments(name, first, count, chunk, spec):
    """Every group element in the reference's proof is the compression of a recorded best_multiexp result, in
    call order: the proof bytes are determined by this path's outputs (plus field evaluations)."""
    ent, z, g, gl = _load(name)
    proof = z["proof"].tobytes()
    want = proof[32 * chunk: 32 * (chunk + count)]
    got = b"".join(_compress(spec, z[f"msm{i}_affine"]) for i in range(first, first + count))
    assert got == want
    if name == "arithmetic":  # 13 points + 24 scalars = the whole 1184-byte proof
        assert len(proof) == 32 * 37
    if name == "collatz":
        assert len(proof) == 32 * 20


@pytest.mark.gpu
@pytest.mark.parametrize("name,first,count,chunk", PROOF_POINT_RUNS)
def test_cuda_reproduces_reference_proof_bytes(name, first, count, chunk, h2b, href):
    """The same bytes from the CUDA path: SRS read from the reference's params bytes, the recorded scalars
    committed on the GPU (one batch per base array), results encoded by h2b_g1_to_bytes."""
    ent, z, g, gl = _load(name)
    params = h2b.ParamsKZG.read(z["params"].tobytes())
    recs = {m["i"]: m for m in ent["msm"]}
    proof = z["proof"].tobytes()
    out = {}
    by_kind = {}
    for i in range(first, first + count):
        by_kind.setdefault((recs[i]["bases"], recs[i]["n"]), []).append(i)
    for (kind, n), idxs in by_kind.items():
        cols = [np.ascontiguousarray(z[f"msm{i}_scalars"]) for i in idxs]
        pts = params.commit_many(cols) if kind == "g" else params.commit_lagrange_many(cols)
        enc = h2b.g1_to_bytes(pts)
        for j, i in enumerate(idxs):
            out[i] = enc[32 * j: 32 * j + 32]
    got = b"".join(out[i] for i in range(first, first + count))
    assert got == proof[32 * chunk: 32 * (chunk + count)]
    assert h2b.g1_to_bytes(np.array([[0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0]], dtype=np.uint64)) == bytes(32)  # identity
    params.release()


def test_reference_g_lagrange_is_the_group_ifft_of_g(href, spec):
    """In the reference's setup() bytes, g_lagrange[i] = sum_j [omega^(-ij) / n] g[j] (checked for a few i with the
    oracle's MSM): the relation g_to_lagrange computes."""
    ent, z, g, gl = _load("arithmetic")
    k = ent["k"]
    n = 1 << k
    w_inv = pow(pow(spec.ROOT_OF_UNITY, 1 << (28 - k), spec.R_MOD), -1, spec.R_MOD)
    n_inv = pow(n, -1, spec.R_MOD)
    for i in (0, 1, 5, n - 1):
        sc = spec.fr_array([pow(w_inv, i * j, spec.R_MOD) * n_inv % spec.R_MOD for j in range(n)])
        assert (href.g1_to_affine(href.best_multiexp(sc, g)) == gl[i]).all(), i


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_cuda_g_to_lagrange_reproduces_reference_params(name, h2b):
    """h2b_g_to_lagrange on the g of the reference's params bytes == the g_lagrange in the same bytes
    (k = 4, 7 and 10), limb for limb."""
    ent, z, g, gl = _load(name)
    assert (h2b.g_to_lagrange(g, ent["k"]) == gl).all()


# (circuit, j = cs.degree(), [(lagrange_to_coeff record, coeff_to_extended record)], extended_to_coeff record)
# The reference reaches best_fft through EvaluationDomain: a coeff_to_extended call shows up as a best_fft record
# whose INPUT is the scaled, zero-padded coefficient vector; the coefficients themselves are the (1/n-scaled)
# output of the column's earlier lagrange_to_coeff record.
DOMAIN_RUNS = [("arithmetic", 3, [(33, 36), (34, 37), (35, 38), (24, 39), (25, 26), (31, 32)], 40)]


def _domain_case(spec, name, j, pairs, e2c):
    ent, z, g, gl = _load(name)
    k = ent["k"]
    n_inv = pow(1 << k, -1, spec.R_MOD)
    cases = []
    for l2c, c2e in pairs:
        coeffs = spec.fr_array([v * n_inv % spec.R_MOD for v in spec.fr_ints(z[f"fft{l2c}_out"])])
        cases.append((coeffs, np.ascontiguousarray(z[f"fft{c2e}_in"]), np.ascontiguousarray(z[f"fft{c2e}_out"])))
    return k, cases, np.ascontiguousarray(z[f"fft{e2c}_in"]), np.ascontiguousarray(z[f"fft{e2c}_out"])


@pytest.mark.parametrize("name,j,pairs,e2c", DOMAIN_RUNS)
def test_oracle_coset_transforms_match_reference_execution(name, j, pairs, e2c, href, spec):
    """coeff_to_extended / extended_to_coeff of the oracle on the reference's own polynomials == what the reference
    computed (this is the check that fixes the coset generator: Fr::ZETA, not its square)."""
    k, cases, e_in, e_out = _domain_case(spec, name, j, pairs, e2c)
    dc = href.domain_new(j, k)
    d = spec.EvaluationDomain(j, k)
    for coeffs, ref_in, ref_out in cases:
        assert (href.coeff_to_extended(dc, coeffs) == ref_out).all()
        assert spec.fr_array(d.coeff_to_extended(spec.fr_ints(coeffs))).tolist() == ref_out.tolist()
        # the recorded input is the coefficient vector after distribute_powers_zeta, zero-padded
        zeta = d.g_coset
        want_in = [c * pow(zeta, i % 3, spec.R_MOD) % spec.R_MOD for i, c in enumerate(spec.fr_ints(coeffs))]
        assert spec.fr_ints(ref_in)[: len(want_in)] == want_in and not ref_in[len(want_in):].any()
    # extended_to_coeff = recorded best_fft output * 1/2^ext_k * {1, zeta^-1, zeta^-2}[i % 3], truncated to n (j - 1)
    ext_n = e_in.shape[0]
    div = pow(ext_n, -1, spec.R_MOD)
    zi = pow(d.g_coset, -1, spec.R_MOD)
    want = [v * div % spec.R_MOD * pow(zi, i % 3, spec.R_MOD) % spec.R_MOD for i, v in enumerate(spec.fr_ints(e_out))]
    want = spec.fr_array(want[: (1 << k) * (j - 1)])
    assert (href.extended_to_coeff(dc, e_in) == want).all()
    # divide_by_vanishing_poly: the recorded extended_to_coeff input is h / (X^n - 1); multiplying the numerator
    # back (values = input / t_evaluations) and dividing again must return it, and t_evaluations must be
    # 1 / ((zeta omega_ext^i)^n - 1) for the reference's zeta
    tev = spec.fr_ints(np.array(list(dc.t_evaluations), dtype=np.uint64)[: 4 * dc.n_t].reshape(-1, 4))
    ext_omega = pow(spec.ROOT_OF_UNITY, 1 << (28 - dc.extended_k), spec.R_MOD)
    assert tev == [pow((pow(d.g_coset * pow(ext_omega, i, spec.R_MOD) % spec.R_MOD, 1 << k, spec.R_MOD) - 1) % spec.R_MOD, -1,
                       spec.R_MOD) for i in range(dc.n_t)]
    numer = spec.fr_array([v * pow(tev[i % dc.n_t], -1, spec.R_MOD) % spec.R_MOD for i, v in enumerate(spec.fr_ints(e_in))])
    assert (href.divide_by_vanishing_poly(dc, numer) == e_in).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name,j,pairs,e2c", DOMAIN_RUNS)
def test_cuda_coset_transforms_match_reference_execution(name, j, pairs, e2c, h2b, href, spec):
    k, cases, e_in, e_out = _domain_case(spec, name, j, pairs, e2c)
    d = h2b.EvaluationDomain(j, k)
    for coeffs, _, ref_out in cases:
        assert (d.coeff_to_extended(coeffs) == ref_out).all()
    outs = d.coeff_to_extended_many([c for c, _, _ in cases])
    assert all((o == ref_out).all() for o, (_, _, ref_out) in zip(outs, cases))
    dc = href.domain_new(j, k)
    assert (d.extended_to_coeff(e_in) == href.extended_to_coeff(dc, e_in)).all()
