"""pgdrome_b200: the progressive PGD enrichment hot path of BAMresearch/PGDrome on B200 (sm_100a).

Drop-in surface (pgdrome/solver.py, pgdrome/model.py): ``PGDProblem`` / ``PGDProblem1``,
``FD_matrices``, ``PGD`` / ``PGDModel``, ``PGDMesh``, ``PGDAttribute``, ``PGDErrorComputation``;
``pgdrome_b200.dolfin`` stands in for the ``dolfin`` namespace the user callbacks use.
All dof arithmetic runs in libpgdb200.so (C ABI: include/pgd_b200.h); there is no CPU fallback.
"""
import sys

__version__ = "0.1.0"

_LAZY = {
    "PGDProblem": "solver", "PGDProblem1": "solver", "FD_matrices": "solver",
    "PGD": "model", "PGDModel": "model", "PGDMesh": "model", "PGDAttribute": "model", "PGDErrorComputation": "model",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module("." + _LAZY[name], __name__), name)
    raise AttributeError(name)


def install_as_reference():
    """Register this package under the reference's import names so that unmodified user scripts
    (``import dolfin``, ``from pgdrome.solver import PGDProblem``) run on the B200 path."""
    import types

    from . import dolfin as _dolfin
    from . import model as _model
    from . import solver as _solver

    sys.modules.setdefault("dolfin", _dolfin)
    sys.modules.setdefault("fenics", _dolfin)
    pkg = types.ModuleType("pgdrome")
    pkg.solver, pkg.model = _solver, _model
    sys.modules.setdefault("pgdrome", pkg)
    sys.modules.setdefault("pgdrome.solver", _solver)
    sys.modules.setdefault("pgdrome.model", _model)
