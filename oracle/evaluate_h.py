"""TEST INFRASTRUCTURE ONLY -- big-integer restatement of the quotient-numerator evaluation of create_proof.

Restates halo2_proofs @6b43b6b src/plonk/evaluation.rs (an un-vendored dependency of the reference,
/root/reference/circuits/Cargo.lock:836-838; reached from circuits/src/utils.rs:83-91, :105-120 through
create_proof -> Evaluator::evaluate_h, wasm func 39 in SURVEY.md Appendix A):

  * ``GraphEvaluator::evaluate``: the compiled expression graph (ValueSource / Calculation) evaluated at one
    row of the extended domain, rotations resolved by ``get_rotation_idx(idx, rot, rot_scale, isize)``;
  * ``Evaluator::evaluate_h`` for circuits without lookups (none of the reference's three circuits has one):
    values[idx] = custom_gates(previous = values[idx]), then the permutation argument folded in with y:
    l_0 (1 - z_0), l_last (z_l^2 - z_l), l_0 (z_i - z_{i-1}(omega^last X)) for i > 0, and per column chunk
    (1 - l_last - l_blind) (z_i(omega X) prod (v + beta s + gamma) - z_i(X) prod (v + delta^j beta X + gamma)).

PINNED against the reference's own execution (tests/test_evaluate_h.py): in the recorded arithmetic-circuit
proof every input of evaluate_h is the output of a recorded coeff_to_extended call and its result is the
input of the recorded extended_to_coeff call; the 32 rows give 32 equations that are linear in the 17
monomials y^a, beta y^a, gamma y^a.  The system is consistent only for this restatement (rank 17, zero
residual, and the recovered monomials are multiplicatively consistent), which fixes the term order, the
rotation of z_{i-1}, the blinding-factor count, the coset generator and DELTA.

Scalars are Python ints mod r; columns are lists of ints over the extended domain.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import bn254 as spec

R = spec.R_MOD

# ValueSource kinds (same order as upstream's enum)
CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
# Calculation kinds
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)

DELTA = pow(7, 1 << 28, R)  # Fr::DELTA = MULTIPLICATIVE_GENERATOR^(2^S)


def get_rotation_idx(idx: int, rot: int, rot_scale: int, isize: int) -> int:
    """evaluation.rs: (((idx as i32) + (rot * rot_scale)).rem_euclid(isize)) as usize."""
    return (idx + rot * rot_scale) % isize


@dataclass
class Graph:
    """GraphEvaluator: constants, rotations, and the calculation list (each: kind, target, operands).
    Operands are ValueSources ``(kind, a, b)``: Fixed/Advice/Instance carry (column, rotation index)."""
    constants: List[int] = field(default_factory=list)
    rotations: List[int] = field(default_factory=list)
    calcs: List[Tuple] = field(default_factory=list)   # (ADD, target, src_a, src_b) ... (HORNER, target, start, [parts], factor)
    num_intermediates: int = 0

    def add_constant(self, c: int) -> Tuple:
        c %= R
        if c not in self.constants:
            self.constants.append(c)
        return (CONSTANT, self.constants.index(c), 0)

    def add_rotation(self, rot: int) -> int:
        if rot not in self.rotations:
            self.rotations.append(rot)
        return self.rotations.index(rot)

    def add_calc(self, kind: int, *operands) -> Tuple:
        target = self.num_intermediates
        self.num_intermediates += 1
        self.calcs.append((kind, target) + tuple(operands))
        return (INTERMEDIATE, target, 0)


@dataclass
class Scalars:
    challenges: Sequence[int]
    beta: int
    gamma: int
    theta: int
    y: int


def _get(src, g: Graph, rot_idx: List[int], inter: List[int], fixed, advice, instance, sc: Scalars, prev: int) -> int:
    kind, a, b = src
    if kind == CONSTANT:
        return g.constants[a]
    if kind == INTERMEDIATE:
        return inter[a]
    if kind == FIXED:
        return fixed[a][rot_idx[b]]
    if kind == ADVICE:
        return advice[a][rot_idx[b]]
    if kind == INSTANCE:
        return instance[a][rot_idx[b]]
    if kind == CHALLENGE:
        return sc.challenges[a]
    return {BETA: sc.beta, GAMMA: sc.gamma, THETA: sc.theta, Y: sc.y, PREVIOUS: prev}[kind]


def graph_evaluate(g: Graph, fixed, advice, instance, sc: Scalars, prev: int, idx: int, rot_scale: int, isize: int) -> int:
    """GraphEvaluator::evaluate: the value of the last calculation (0 for an empty graph)."""
    rot_idx = [get_rotation_idx(idx, r, rot_scale, isize) for r in g.rotations]
    inter = [0] * g.num_intermediates
    get = lambda s: _get(s, g, rot_idx, inter, fixed, advice, instance, sc, prev)
    for c in g.calcs:
        kind, target = c[0], c[1]
        if kind == ADD:
            v = get(c[2]) + get(c[3])
        elif kind == SUB:
            v = get(c[2]) - get(c[3])
        elif kind == MUL:
            v = get(c[2]) * get(c[3])
        elif kind == SQUARE:
            v = get(c[2]) ** 2
        elif kind == DOUBLE:
            v = 2 * get(c[2])
        elif kind == NEGATE:
            v = -get(c[2])
        elif kind == HORNER:
            v = get(c[2])
            factor = get(c[4])
            for part in c[3]:
                v = v * factor + get(part)
        elif kind == STORE:
            v = get(c[2])
        else:
            raise ValueError(kind)
        inter[target] = v % R
    return inter[g.calcs[-1][1]] if g.calcs else 0


@dataclass
class Permutation:
    """permutation::Argument + its proving-key / prover data on the extended coset."""
    columns: List[Tuple[int, int]]     # (kind, index) with kind in {ADVICE, FIXED, INSTANCE}, in cs.permutation order
    sigma_cosets: List[Sequence[int]]  # pk.permutation.cosets, one per column
    z_cosets: List[Sequence[int]]      # permutation_product_coset, one per chunk of `chunk_len` columns
    chunk_len: int                     # cs.degree() - 2
    last_rotation: int                 # -(blinding_factors + 1)
    l0: Sequence[int]
    l_last: Sequence[int]
    l_active_row: Sequence[int]


@dataclass
class Lookup:
    """lookup::Argument on the extended coset.  ``graph`` is upstream's per-lookup GraphEvaluator, whose value is
    (theta-compressed input + beta) * (theta-compressed table + gamma).  PARITY UNPINNED: none of the reference's
    circuits has a lookup argument, so this part rests on the upstream source as restated, not on a record."""
    graph: Graph
    product_coset: Sequence[int]
    permuted_input_coset: Sequence[int]
    permuted_table_coset: Sequence[int]


def fold_lookup(values: List[int], lk: Lookup, fixed, advice, instance, sc: Scalars, l0, l_last, l_active_row,
                k: int, extended_k: int) -> List[int]:
    """evaluation.rs, "Lookup constraints": five terms folded into ``values`` with y."""
    size = 1 << extended_k
    rot_scale = 1 << (extended_k - k)
    y, beta, gamma = sc.y, sc.beta, sc.gamma
    out = list(values)
    for idx in range(size):
        table_value = graph_evaluate(lk.graph, fixed, advice, instance, sc, 0, idx, rot_scale, size)
        r_next = get_rotation_idx(idx, 1, rot_scale, size)
        r_prev = get_rotation_idx(idx, -1, rot_scale, size)
        z, a, s_ = lk.product_coset, lk.permuted_input_coset, lk.permuted_table_coset
        a_minus_s = a[idx] - s_[idx]
        v = out[idx]
        v = (v * y + (1 - z[idx]) * l0[idx]) % R
        v = (v * y + (z[idx] * z[idx] - z[idx]) * l_last[idx]) % R
        v = (v * y + (z[r_next] * (a[idx] + beta) * (s_[idx] + gamma) - z[idx] * table_value) * l_active_row[idx]) % R
        v = (v * y + a_minus_s * l0[idx]) % R
        v = (v * y + a_minus_s * (a[idx] - a[r_prev]) * l_active_row[idx]) % R
        out[idx] = v
    return out


def evaluate_h(graph: Graph, fixed, advice, instance, sc: Scalars, perm: Permutation | None,
               k: int, extended_k: int, extended_omega: int, zeta: int) -> List[int]:
    """Evaluator::evaluate_h for one circuit instance without lookups; ``zeta`` is the domain's coset generator
    (Fr::ZETA as the reference uses it: the recorded cosets are evaluations on zeta * <extended_omega>)."""
    size = 1 << extended_k
    rot_scale = 1 << (extended_k - k)
    values = [0] * size
    for idx in range(size):
        values[idx] = graph_evaluate(graph, fixed, advice, instance, sc, values[idx], idx, rot_scale, size)
    if perm is not None and perm.z_cosets:
        sets = perm.z_cosets
        y, beta, gamma = sc.y, sc.beta, sc.gamma
        delta_start = beta * zeta % R
        col_of = {ADVICE: advice, FIXED: fixed, INSTANCE: instance}
        for idx in range(size):
            v = values[idx]
            r_next = get_rotation_idx(idx, 1, rot_scale, size)
            r_last = get_rotation_idx(idx, perm.last_rotation, rot_scale, size)
            v = (v * y + (1 - sets[0][idx]) * perm.l0[idx]) % R
            zl = sets[-1][idx]
            v = (v * y + (zl * zl - zl) * perm.l_last[idx]) % R
            for i in range(1, len(sets)):
                v = (v * y + (sets[i][idx] - sets[i - 1][r_last]) * perm.l0[idx]) % R
            current_delta = delta_start * pow(extended_omega, idx, R) % R
            for i, zc in enumerate(sets):
                cols = perm.columns[i * perm.chunk_len:(i + 1) * perm.chunk_len]
                sig = perm.sigma_cosets[i * perm.chunk_len:(i + 1) * perm.chunk_len]
                left = zc[r_next]
                for (kind, index), s in zip(cols, sig):
                    left = left * (col_of[kind][index][idx] + beta * s[idx] + gamma) % R
                right = zc[idx]
                for (kind, index) in cols:
                    right = right * (col_of[kind][index][idx] + current_delta + gamma) % R
                    current_delta = current_delta * DELTA % R
                v = (v * y + (left - right) * perm.l_active_row[idx]) % R
            values[idx] = v
    return values
