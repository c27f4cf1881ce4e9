"""Host mirror of halo2_proofs::plonk::evaluation (src/plonk/evaluation.rs @6b43b6b) for the device-side
quotient numerator: ``GraphEvaluator`` (ValueSource / Calculation, same vocabulary and order as upstream) and
``evaluate_h`` over columns that already sit on the extended coset in HBM (h2b_dev_evaluate_h).  Lookups are
not covered (the reference's circuits have none)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from . import _ffi

# ValueSource kinds / Calculation kinds: upstream's enum order
CONSTANT, INTERMEDIATE, FIXED, ADVICE, INSTANCE, CHALLENGE, BETA, GAMMA, THETA, Y, PREVIOUS = range(11)
ADD, SUB, MUL, SQUARE, DOUBLE, NEGATE, HORNER, STORE = range(8)


def _src(s) -> int:
    kind, a, b = s
    return kind | (a << 8) | (b << 36)


@dataclass
class GraphEvaluator:
    """constants: Montgomery Fr limbs ((m,4) uint64); rotations: i32; calculations: tuples
    (ADD|SUB|MUL, target, a, b), (SQUARE|DOUBLE|NEGATE|STORE, target, a), (HORNER, target, start, parts, factor)."""
    constants: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), dtype=np.uint64))
    rotations: List[int] = field(default_factory=list)
    calculations: List[Tuple] = field(default_factory=list)
    num_intermediates: int = 0

    def serialize(self) -> np.ndarray:
        words: List[int] = []
        for c in self.calculations:
            kind, target = c[0], c[1]
            if kind == HORNER:
                start, parts, factor = c[2], c[3], c[4]
                words.append(kind | (target << 8) | (len(parts) << 40))
                words += [_src(start), _src(factor)] + [_src(p) for p in parts]
            else:
                words.append(kind | (target << 8))
                words += [_src(s) for s in c[2:]]
        return np.array(words, dtype=np.uint64)


class _EvalH(C.Structure):
    _fields_ = [
        ("num_fixed", C.c_uint32), ("num_advice", C.c_uint32), ("num_instance", C.c_uint32), ("num_challenges", C.c_uint32),
        ("fixed", C.POINTER(C.c_void_p)), ("advice", C.POINTER(C.c_void_p)), ("instance", C.POINTER(C.c_void_p)),
        ("challenges", C.POINTER(C.c_uint64)),
        ("beta", C.c_uint64 * 4), ("gamma", C.c_uint64 * 4), ("theta", C.c_uint64 * 4), ("y", C.c_uint64 * 4),
        ("num_constants", C.c_uint32), ("num_rotations", C.c_uint32), ("num_calcs", C.c_uint32), ("num_intermediates", C.c_uint32),
        ("constants", C.POINTER(C.c_uint64)), ("rotations", C.POINTER(C.c_int32)), ("calcs", C.POINTER(C.c_uint64)),
        ("calc_words", C.c_size_t),
        ("num_perm_columns", C.c_uint32), ("chunk_len", C.c_uint32), ("last_rotation", C.c_int32),
        ("perm_kind", C.POINTER(C.c_uint8)), ("perm_index", C.POINTER(C.c_uint32)),
        ("sigma_cosets", C.POINTER(C.c_void_p)), ("z_cosets", C.POINTER(C.c_void_p)),
        ("l0", C.c_void_p), ("l_last", C.c_void_p), ("l_active_row", C.c_void_p),
        ("flags", C.c_uint32),
    ]


EVALH_ACCUMULATE = 1  # start from the values already in values_t (PreviousValue threaded through several circuits)


@dataclass
class PermutationData:
    columns: Sequence[Tuple[int, int]]   # (ADVICE | FIXED | INSTANCE, index) in cs.permutation order
    sigma_cosets: Sequence               # cuda tensors
    z_cosets: Sequence
    chunk_len: int
    last_rotation: int
    l0: object
    l_last: object
    l_active_row: object


def _ptrs(tensors) -> C.Array:
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _make_args(graph: GraphEvaluator, fixed, advice, instance, challenges, beta, gamma, theta, y, perm):
    a = _EvalH()
    keep = []
    a.num_fixed, a.num_advice, a.num_instance = len(fixed), len(advice), len(instance)
    for name, ts in (("fixed", fixed), ("advice", advice), ("instance", instance)):
        arr = _ptrs(ts)
        keep.append(arr)
        setattr(a, name, C.cast(arr, C.POINTER(C.c_void_p)))
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    a.num_challenges = ch.shape[0]
    a.challenges = ch.ctypes.data_as(C.POINTER(C.c_uint64))
    for name, v in (("beta", beta), ("gamma", gamma), ("theta", theta), ("y", y)):
        v = np.ascontiguousarray(v, dtype=np.uint64).reshape(4)
        setattr(a, name, (C.c_uint64 * 4)(*[int(x) for x in v]))
    consts = np.ascontiguousarray(graph.constants, dtype=np.uint64).reshape(-1, 4)
    rots = np.array(graph.rotations, dtype=np.int32)
    calcs = graph.serialize()
    a.num_constants, a.num_rotations = consts.shape[0], rots.shape[0]
    a.num_calcs, a.num_intermediates = len(graph.calculations), graph.num_intermediates
    a.constants = consts.ctypes.data_as(C.POINTER(C.c_uint64))
    a.rotations = rots.ctypes.data_as(C.POINTER(C.c_int32))
    a.calcs = calcs.ctypes.data_as(C.POINTER(C.c_uint64))
    a.calc_words = calcs.shape[0]
    if perm is not None and len(perm.columns):
        kind_map = {ADVICE: 0, FIXED: 1, INSTANCE: 2}
        kinds = np.array([kind_map[k] for k, _ in perm.columns], dtype=np.uint8)
        idxs = np.array([i for _, i in perm.columns], dtype=np.uint32)
        sig, zc = _ptrs(perm.sigma_cosets), _ptrs(perm.z_cosets)
        keep += [kinds, idxs, sig, zc]
        a.num_perm_columns, a.chunk_len, a.last_rotation = len(perm.columns), perm.chunk_len, perm.last_rotation
        a.perm_kind = kinds.ctypes.data_as(C.POINTER(C.c_uint8))
        a.perm_index = idxs.ctypes.data_as(C.POINTER(C.c_uint32))
        a.sigma_cosets = C.cast(sig, C.POINTER(C.c_void_p))
        a.z_cosets = C.cast(zc, C.POINTER(C.c_void_p))
        a.l0, a.l_last, a.l_active_row = perm.l0.data_ptr(), perm.l_last.data_ptr(), perm.l_active_row.data_ptr()
    keep += [ch, consts, rots, calcs]
    return a, keep


def dev_evaluate_h(domain, graph: GraphEvaluator, fixed, advice, instance, challenges: np.ndarray, beta, gamma, theta, y,
                   perm: PermutationData | None, values_t, stream=None, accumulate: bool = False) -> None:
    """Evaluator::evaluate_h (custom gates + permutation argument) into ``values_t`` ((2^extended_k, 4) int64 cuda
    tensor).  Columns are cuda tensors holding extended cosets; scalars are (4,) uint64 Montgomery limbs."""
    from .arithmetic import _stream_ptr
    a, keep = _make_args(graph, fixed, advice, instance, challenges, beta, gamma, theta, y, perm)
    a.flags = EVALH_ACCUMULATE if accumulate else 0
    _ffi.check(_ffi.lib().h2b_dev_evaluate_h(C.byref(domain._d), C.byref(a), C.c_void_p(values_t.data_ptr()), _stream_ptr(stream)))


def dev_evaluate_h_lookup(domain, lookup_graph: GraphEvaluator, fixed, advice, instance, challenges: np.ndarray, beta, gamma,
                          theta, y, l0, l_last, l_active_row, product_t, permuted_input_t, permuted_table_t, values_t,
                          stream=None) -> None:
    """One lookup argument folded into ``values_t`` after dev_evaluate_h (call once per lookup, in order)."""
    from .arithmetic import _stream_ptr
    a, keep = _make_args(lookup_graph, fixed, advice, instance, challenges, beta, gamma, theta, y, None)
    a.l0, a.l_last, a.l_active_row = l0.data_ptr(), l_last.data_ptr(), l_active_row.data_ptr()
    _ffi.check(_ffi.lib().h2b_dev_evaluate_h_lookup(
        C.byref(domain._d), C.byref(a), C.c_void_p(product_t.data_ptr()), C.c_void_p(permuted_input_t.data_ptr()),
        C.c_void_p(permuted_table_t.data_ptr()), C.c_void_p(values_t.data_ptr()), _stream_ptr(stream)))
