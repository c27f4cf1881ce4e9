"""TEST INFRASTRUCTURE ONLY -- the constants and the permutation of the reference's Poseidon circuit, restated.

The reference's PoseidonSpec (circuits/src/poseidon_circuit.rs:126-149: 8 full rounds, N_ROUNDS_P[WIDTH] partial
rounds, x^5, secure_mds = 0) takes its round constants and MDS matrix from halo2_gadgets' `generate_constants`
(an un-vendored git dependency, circuits/Cargo.toml:17; the reference carries a source copy of the same module at
circuits/src/poseidon/primitives.rs:57-84, primitives/grain.rs, primitives/mds.rs, which is what is restated here):

  * `Grain`: the 80-bit Grain LFSR of the Poseidon paper in self-shrinking mode, seeded with the field type, the
    S-box type, the field size in bits, t, R_F and R_P; the first 160 bits are discarded;
  * round constants: (R_F + R_P) * t field elements by rejection sampling of NUM_BITS-bit big-endian strings;
  * MDS: 2t elements without rejection (reduced mod r), xs = first t, ys = last t, m[i][j] = 1 / (xs[i] + ys[j]); the
    inverse is computed here by Gaussian elimination (the reference uses the closed form for Cauchy matrices).

PINNED by the recorded execution of the reference's Poseidon proof: `hash([1, 2])` below equals the public output the
reference computed for that input (tests/golden/wasm_manifest.json, "poseidon".input.output), and the fixed columns
rc_a / rc_b of the recorded proving key hold exactly these round constants (tests/test_evaluate_h.py); and by the
reference's own known answers for the Pallas field (tests/golden/poseidon_pallas_kat.json, cut from
circuits/src/poseidon/primitives/fp.rs and test_vectors.rs by oracle/make_poseidon_kat.py): all 192 round constants, MDS,
MDS^-1, permutation and hash vectors (tests/test_oracle.py).
"""
from __future__ import annotations

import bn254 as spec

R = spec.R_MOD
NUM_BITS = 254
FULL_ROUNDS = 8
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]   # poseidon_circuit.rs:20-22


class Grain:
    def __init__(self, t: int, r_f: int, r_p: int, sbox_tag: int = 0, field_tag: int = 1, modulus: int = R,
                 num_bits: int = NUM_BITS):
        self.modulus, self.num_bits = modulus, num_bits
        bits = []
        for value, length in ((field_tag, 2), (sbox_tag, 4), (num_bits, 12), (t, 12), (r_f, 10), (r_p, 10)):
            bits += [(value >> (length - 1 - i)) & 1 for i in range(length)]
        bits += [1] * 30
        assert len(bits) == 80
        self.state = bits
        for _ in range(160):
            self._raw()

    def _raw(self) -> int:
        s = self.state
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self) -> int:
        """self-shrinking: a pair (1, b) yields b, a pair (0, _) yields nothing"""
        while True:
            first, second = self._raw(), self._raw()
            if first:
                return second

    def _int(self) -> int:
        v = 0
        for _ in range(self.num_bits):
            v = (v << 1) | self.bit()
        return v

    def field_element(self) -> int:
        while True:
            v = self._int()
            if v < self.modulus:
                return v

    def field_element_without_rejection(self) -> int:
        return self._int() % self.modulus


def _inverse_matrix(m, R=R):
    t = len(m)
    a = [row[:] + [int(i == j) for j in range(t)] for i, row in enumerate(m)]
    for c in range(t):
        p = next(i for i in range(c, t) if a[i][c] % R)
        a[c], a[p] = a[p], a[c]
        inv = pow(a[c][c], -1, R)
        a[c] = [v * inv % R for v in a[c]]
        for i in range(t):
            if i != c and a[i][c]:
                f = a[i][c]
                a[i] = [(v - f * w) % R for v, w in zip(a[i], a[c])]
    return [row[t:] for row in a]


def generate_constants(t: int, r_p: int | None = None, modulus: int = R, num_bits: int = NUM_BITS):
    """-> (round_constants[R_F + R_P][t], mds[t][t], mds_inv[t][t]).  Defaults: the reference's PoseidonSpec over BN254
    Fr; other fields only for the known-answer test against the reference's Pallas constants
    (circuits/src/poseidon/primitives/fp.rs, `generate_parameters_grain.sage 1 0 255 3 8 56 p`)."""
    R = modulus
    r_f, r_p = FULL_ROUNDS, (N_ROUNDS_P[t] if r_p is None else r_p)
    g = Grain(t, r_f, r_p, modulus=modulus, num_bits=num_bits)
    rc = [[g.field_element() for _ in range(t)] for _ in range(r_f + r_p)]
    while True:
        vals = [g.field_element_without_rejection() for _ in range(2 * t)]
        if len(set(vals)) == len(vals):
            break
    xs, ys = vals[:t], vals[t:]
    mds = [[pow((xs[i] + ys[j]) % R, -1, R) for j in range(t)] for i in range(t)]
    return rc, mds, _inverse_matrix(mds, R)


def permute(state, rc, mds, R=R):
    t = len(state)
    half, r_p = FULL_ROUNDS // 2, len(rc) - FULL_ROUNDS
    mix = lambda s: [sum(mds[i][j] * s[j] for j in range(t)) % R for i in range(t)]
    r = 0
    for phase, count in (("full", half), ("partial", r_p), ("full", half)):
        for _ in range(count):
            state = [(s + c) % R for s, c in zip(state, rc[r])]
            if phase == "full":
                state = [pow(s, 5, R) for s in state]
            else:
                state[0] = pow(state[0], 5, R)
            state = mix(state)
            r += 1
    return state


def hash_constant_length(message, t: int = 3, constants=None, R=R):
    """poseidon::Hash::<_, S, ConstantLength<L>, WIDTH, RATE>::init().hash(message), RATE = WIDTH - 1
    (poseidon_circuit.rs:292-299): capacity word L * 2^64, zero padding up to a multiple of the rate."""
    rate = t - 1
    rc, mds, _ = constants if constants is not None else generate_constants(t)
    msg = list(message) + [0] * (-len(message) % rate)
    state = [0] * rate + [(len(message) << 64) % R]
    for off in range(0, len(msg), rate):
        state = [(s + m) % R for s, m in zip(state, msg[off:off + rate])] + state[rate:]
        state = permute(state, rc, mds, R)
    return state[0]
