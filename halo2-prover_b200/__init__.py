"""halo2-prover_b200 -- B200-native MSM / NTT hot path behind halo2_proofs' call sites.

Host-side mirror (Python) of the reference interface for the path, over the C ABI
in ``include/h2b200.h`` (``csrc/libh2b200.so``):

* ``arithmetic.best_multiexp`` / ``arithmetic.best_fft``
  (halo2_proofs @6b43b6b src/arithmetic.rs:147-180, :185-290)
* ``domain.EvaluationDomain`` -- ``new(j, k)``, ``lagrange_to_coeff``,
  ``coeff_to_extended``, ``extended_to_coeff``, ``divide_by_vanishing_poly``
  (src/poly/domain.rs:~40-140, :227, :244, :311)
* ``kzg.ParamsKZG`` -- ``commit`` / ``commit_lagrange`` against a device-resident SRS
  (src/poly/kzg/commitment.rs:319, :363)
* ``multi_gpu.sharded_multiexp`` -- point-range sharding across ranks, 96-byte gather.

Arrays are numpy ``uint64`` in the FFI layout (Fr/Fq: 4 limbs LE Montgomery;
G1Affine: 8 limbs; G1: 12 limbs homogeneous projective).  There is no CPU fallback: importing
works anywhere, computing requires the CUDA library and a GPU and raises otherwise.
"""
from . import _ffi  # noqa: F401
from .arithmetic import best_fft, best_multiexp, g1_fold, g1_to_bytes, g_to_lagrange  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .kzg import ParamsKZG  # noqa: F401

__all__ = ["best_multiexp", "best_fft", "g1_fold", "g1_to_bytes", "g_to_lagrange", "EvaluationDomain", "ParamsKZG"]
