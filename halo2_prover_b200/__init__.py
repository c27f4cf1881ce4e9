"""Import shim: the product package lives in ``halo2-prover_b200/`` (the directory name the
project layout prescribes); a hyphen is not importable, so this package points its
``__path__`` at that directory and re-exports it as ``halo2_prover_b200``."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "halo2-prover_b200")
__path__.insert(0, _impl)  # noqa: F821  (submodules resolve inside halo2-prover_b200/)
with open(_os.path.join(_impl, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _f, _os, _impl
