// evalh.cuh -- the quotient numerator of create_proof on the extended domain.
//
// Device counterpart of halo2_proofs @6b43b6b src/plonk/evaluation.rs (Evaluator::evaluate_h, GraphEvaluator::
// evaluate; wasm func 39 of the reference binary), the step BETWEEN coeff_to_extended and extended_to_coeff:
// with it the extended columns never leave HBM.  Covers the custom gates (the compiled expression graph, same
// ValueSource / Calculation vocabulary as upstream, serialised by the caller) and the permutation argument in ONE
// pass over the rows (intermediates in a per-thread array after the host has renumbered them by liveness; the
// coset point X of a row from the cached twiddle table); a lookup argument is a separate fold (evalh_lookup_kernel).  The tests check it against a CPU
// restatement that is itself pinned on the reference's recorded execution (tests/test_evaluate_h.py).
#pragma once
#include "field.cuh"

namespace h2b {

// ValueSource: kind | a << 8 | b << 36   (a: constant / intermediate / column / challenge index, b: rotation index)
enum : uint32_t { VS_CONSTANT, VS_INTERMEDIATE, VS_FIXED, VS_ADVICE, VS_INSTANCE, VS_CHALLENGE, VS_BETA, VS_GAMMA, VS_THETA,
                  VS_Y, VS_PREVIOUS };
// Calculation header: op | target << 8 | nparts << 40, followed by its operands (Horner: start, factor, parts...)
enum : uint32_t { CALC_ADD, CALC_SUB, CALC_MUL, CALC_SQUARE, CALC_DOUBLE, CALC_NEGATE, CALC_HORNER, CALC_STORE };

struct EvalGates {
    const Fe *const *fixed;     // device arrays of device pointers to extended cosets
    const Fe *const *advice;
    const Fe *const *instance;
    const Fe *challenges;
    const Fe *constants;
    const int32_t *rotations;
    const uint64_t *calcs;      // the serialised calculation list, intermediates renumbered to `slots` reusable slots
    Fe *scratch;                // slots in HBM (scratch[slot * size + idx]) when they do not fit the per-thread array
    uint32_t num_rotations, num_calcs, size, rot_scale;
    Fe beta, gamma, theta, y;
};

constexpr uint32_t kMaxRotations = 32;
// Intermediates of one row live in a per-thread array (local memory, L1-resident) when the graph needs at most this
// many at a time; the host renumbers them by liveness first (a slot is reused once its value has been read for the
// last time), so a graph with hundreds of intermediates typically needs a dozen slots.
constexpr uint32_t kLocalSlots = 40;

H2B_DI uint32_t rotation_idx(uint32_t idx, int32_t rot, uint32_t rot_scale, uint32_t size) {
    // size is a power of two: rem_euclid is a mask
    return (uint32_t)((int32_t)idx + rot * (int32_t)rot_scale) & (size - 1);
}

template <bool LOCAL>
H2B_DI Fe eval_source(uint64_t src, const EvalGates &g, const uint32_t *rot_idx, uint32_t idx, const Fe &prev, const Fe *inter) {
    const uint32_t kind = (uint32_t)(src & 0xff), a = (uint32_t)((src >> 8) & 0xfffffff), b = (uint32_t)(src >> 36);
    switch (kind) {
        case VS_CONSTANT: return load_fe_ro(&g.constants[a]);
        case VS_INTERMEDIATE: return LOCAL ? inter[a] : load_fe(&g.scratch[(size_t)a * g.size + idx]);
        case VS_FIXED: return load_fe_ro(&g.fixed[a][rot_idx[b]]);
        case VS_ADVICE: return load_fe_ro(&g.advice[a][rot_idx[b]]);
        case VS_INSTANCE: return load_fe_ro(&g.instance[a][rot_idx[b]]);
        case VS_CHALLENGE: return load_fe_ro(&g.challenges[a]);
        case VS_BETA: return g.beta;
        case VS_GAMMA: return g.gamma;
        case VS_THETA: return g.theta;
        case VS_Y: return g.y;
        default: return prev;
    }
}

// GraphEvaluator::evaluate at row idx (the value of the last calculation; zero for an empty graph, as upstream)
template <bool LOCAL>
H2B_DI Fe eval_graph(const EvalGates &g, uint32_t idx, const Fe &prev) {
    uint32_t rot_idx[kMaxRotations];
    for (uint32_t r = 0; r < g.num_rotations; r++) rot_idx[r] = rotation_idx(idx, g.rotations[r], g.rot_scale, g.size);
    Fe inter[LOCAL ? kLocalSlots : 1];
    Fe last = Fr::zero();
    const uint64_t *pc = g.calcs;
    for (uint32_t c = 0; c < g.num_calcs; c++) {
        const uint64_t hdr = __ldg(pc++);
        const uint32_t op = (uint32_t)(hdr & 0xff), target = (uint32_t)((hdr >> 8) & 0xffffffffu), nparts = (uint32_t)(hdr >> 40);
        Fe v;
        if (op == CALC_HORNER) {
            v = eval_source<LOCAL>(__ldg(pc), g, rot_idx, idx, prev, inter);
            const Fe factor = eval_source<LOCAL>(__ldg(pc + 1), g, rot_idx, idx, prev, inter);
            for (uint32_t k = 0; k < nparts; k++)
                v = Fr::add(Fr::mul(v, factor), eval_source<LOCAL>(__ldg(pc + 2 + k), g, rot_idx, idx, prev, inter));
            pc += 2 + nparts;
        } else if (op <= CALC_MUL) {
            const Fe x = eval_source<LOCAL>(__ldg(pc), g, rot_idx, idx, prev, inter);
            const Fe y = eval_source<LOCAL>(__ldg(pc + 1), g, rot_idx, idx, prev, inter);
            v = op == CALC_ADD ? Fr::add(x, y) : (op == CALC_SUB ? Fr::sub(x, y) : Fr::mul(x, y));
            pc += 2;
        } else {
            const Fe x = eval_source<LOCAL>(__ldg(pc), g, rot_idx, idx, prev, inter);
            v = op == CALC_SQUARE ? Fr::sqr(x) : (op == CALC_DOUBLE ? Fr::dbl(x) : (op == CALC_NEGATE ? Fr::neg(x) : x));
            pc += 1;
        }
        if (LOCAL) inter[target] = v;
        else store_fe(&g.scratch[(size_t)target * g.size + idx], v);
        last = v;
    }
    return last;
}

struct EvalPerm {
    const Fe *const *columns;   // the permutation columns' extended cosets, in cs.permutation order
    const Fe *const *sigma;     // pk.permutation.cosets
    const Fe *const *z;         // permutation_product_coset, one per chunk
    const Fe *l0, *l_last, *l_active;
    const Fe *ext_pows;         // extended_omega^i, i < 2^extended_k (the cached twiddle table of the extended domain)
    uint32_t num_columns, chunk_len, num_sets, size, rot_scale;
    int32_t last_rotation;
    Fe beta, gamma, y, zeta, delta;
};

// values[idx] folded with the permutation argument's constraints (evaluation.rs, "Permutation constraints")
H2B_DI Fe eval_permutation(const EvalPerm &p, uint32_t idx, Fe v) {
    const uint32_t r_next = rotation_idx(idx, 1, p.rot_scale, p.size);
    const uint32_t r_last = rotation_idx(idx, p.last_rotation, p.rot_scale, p.size);
    const Fe one = Fr::one();
    const Fe l0 = load_fe_ro(&p.l0[idx]), l_last = load_fe_ro(&p.l_last[idx]), l_active = load_fe_ro(&p.l_active[idx]);
    // l_0(X) * (1 - z_0(X))
    v = Fr::add(Fr::mul(v, p.y), Fr::mul(Fr::sub(one, load_fe_ro(&p.z[0][idx])), l0));
    // l_last(X) * (z_l(X)^2 - z_l(X))
    {
        const Fe zl = load_fe_ro(&p.z[p.num_sets - 1][idx]);
        v = Fr::add(Fr::mul(v, p.y), Fr::mul(Fr::sub(Fr::sqr(zl), zl), l_last));
    }
    // l_0(X) * (z_i(X) - z_{i-1}(omega^last X))
    for (uint32_t i = 1; i < p.num_sets; i++)
        v = Fr::add(Fr::mul(v, p.y), Fr::mul(Fr::sub(load_fe_ro(&p.z[i][idx]), load_fe_ro(&p.z[i - 1][r_last])), l0));
    // (1 - (l_last + l_blind)) * (z_i(omega X) prod (v + beta s + gamma) - z_i(X) prod (v + delta^j beta X + gamma))
    // X = zeta * extended_omega^idx on the coset: one table read instead of a 64-bit exponentiation per row
    Fe current_delta = Fr::mul(Fr::mul(p.beta, p.zeta), load_fe_ro(&p.ext_pows[idx]));
    for (uint32_t i = 0; i < p.num_sets; i++) {
        const uint32_t c0 = i * p.chunk_len, c1 = min(c0 + p.chunk_len, p.num_columns);
        Fe left = load_fe_ro(&p.z[i][r_next]), right = load_fe_ro(&p.z[i][idx]);
        for (uint32_t c = c0; c < c1; c++) {
            const Fe val = load_fe_ro(&p.columns[c][idx]);
            left = Fr::mul(left, Fr::add(Fr::add(val, Fr::mul(p.beta, load_fe_ro(&p.sigma[c][idx]))), p.gamma));
            right = Fr::mul(right, Fr::add(Fr::add(val, current_delta), p.gamma));
            current_delta = Fr::mul(current_delta, p.delta);
        }
        v = Fr::add(Fr::mul(v, p.y), Fr::mul(Fr::sub(left, right), l_active));
    }
    return v;
}

// One pass over the extended domain, one thread per row: values[idx] = custom_gates.evaluate(previous = values[idx] when
// `accumulate`, else zero -- upstream starts from domain.empty_extended() and threads the value through the circuits of
// a proof), then the permutation argument folded in with y.  The row's value never leaves registers in between.
template <bool LOCAL>
__global__ void __launch_bounds__(128)
evalh_fused_kernel(EvalGates g, EvalPerm p, Fe *__restrict__ values, uint32_t accumulate) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.size) return;
    const Fe prev = accumulate ? load_fe(&values[idx]) : Fr::zero();
    Fe v = eval_graph<LOCAL>(g, idx, prev);
    if (p.num_columns) v = eval_permutation(p, idx, v);
    store_fe(&values[idx], v);
}

// One lookup argument folded into values (evaluation.rs, "Lookup constraints").  g is that lookup's graph:
// (compressed input + beta) * (compressed table + gamma).
struct EvalLookup {
    const Fe *product, *permuted_input, *permuted_table, *l0, *l_last, *l_active;
};
template <bool LOCAL>
__global__ void __launch_bounds__(128)
evalh_lookup_kernel(EvalGates g, EvalLookup lk, Fe *__restrict__ values) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.size) return;
    const Fe table_value = eval_graph<LOCAL>(g, idx, Fr::zero());
    const uint32_t r_next = rotation_idx(idx, 1, g.rot_scale, g.size), r_prev = rotation_idx(idx, -1, g.rot_scale, g.size);
    const Fe one = Fr::one();
    const Fe l0 = load_fe_ro(&lk.l0[idx]), l_last = load_fe_ro(&lk.l_last[idx]), l_active = load_fe_ro(&lk.l_active[idx]);
    const Fe z = load_fe_ro(&lk.product[idx]), a = load_fe_ro(&lk.permuted_input[idx]), s = load_fe_ro(&lk.permuted_table[idx]);
    const Fe a_minus_s = Fr::sub(a, s);
    Fe v = load_fe(&values[idx]);
    v = Fr::add(Fr::mul(v, g.y), Fr::mul(Fr::sub(one, z), l0));
    v = Fr::add(Fr::mul(v, g.y), Fr::mul(Fr::sub(Fr::sqr(z), z), l_last));
    {
        const Fe left = Fr::mul(Fr::mul(load_fe_ro(&lk.product[r_next]), Fr::add(a, g.beta)), Fr::add(s, g.gamma));
        v = Fr::add(Fr::mul(v, g.y), Fr::mul(Fr::sub(left, Fr::mul(z, table_value)), l_active));
    }
    v = Fr::add(Fr::mul(v, g.y), Fr::mul(a_minus_s, l0));
    v = Fr::add(Fr::mul(v, g.y), Fr::mul(Fr::mul(a_minus_s, Fr::sub(a, load_fe_ro(&lk.permuted_input[r_prev]))), l_active));
    store_fe(&values[idx], v);
}

}  // namespace h2b
