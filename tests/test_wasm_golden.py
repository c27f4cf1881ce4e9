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
    params.release()
    for bad in (b"", b"\x04\x00\x00", z["params"].tobytes()[:-1]):
        with pytest.raises(Exception):
            h2b.ParamsKZG.read(bad)
