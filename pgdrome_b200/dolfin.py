"""The subset of the ``dolfin`` / ``fenics`` namespace that PGDrome user callbacks touch
(SURVEY.md 2.4, 8b "mini-UFL capture layer"), backed by pgdrome_b200.

User scripts written for the reference (``import dolfin`` + ``from pgdrome.solver import
PGDProblem, FD_matrices``; tests/integration/*.py) run unchanged after
``pgdrome_b200.install_as_reference()``, which registers this module as ``dolfin`` / ``fenics``
and ``pgdrome_b200.solver`` / ``.model`` as ``pgdrome.solver`` / ``pgdrome.model``.
Anything outside the registered separated forms raises NotImplementedError naming the construct.
"""
from .fem import (BoxMesh, FunctionSpace, IntervalMesh, Mesh, Point, RectangleMesh, UnitCubeMesh,  # noqa: F401
                  UnitIntervalMesh, UnitSquareMesh, VectorFunctionSpace)
from .forms import assemble, norm  # noqa: F401
from .functions import (DOLFIN_EPS, DOLFIN_PI, CompiledSubDomain, Constant, DirichletBC, Expression, Function,  # noqa: F401
                        MatrixOperator, MeshFunction, SubDomain, UserExpression, interpolate, near)
from .ufl import (Measure, TestFunction, TrialFunction, as_matrix, as_vector, derivative, dot, ds, dx, grad, inner,  # noqa: F401
                  lhs, rhs)

parameters = {"form_compiler": {"optimize": True, "cpp_optimize": True, "quadrature_degree": None}}


def set_log_level(level):
    return None


class LogLevel:
    DEBUG, INFO, WARNING, ERROR, CRITICAL, PROGRESS, TRACE = 10, 20, 30, 40, 50, 16, 13


def plot(*a, **k):
    raise NotImplementedError("dolfin.plot is outside the PGD hot path")


def errornorm(u, uh, norm_type="L2", degree_rise=0, mesh=None):
    """||u - uh||_L2 for two Functions of the same space."""
    from . import _lib
    from .forms import mass_product

    if u.function_space() is not uh.function_space() and u.function_space().n_dofs != uh.function_space().n_dofs:
        raise NotImplementedError("errornorm between different function spaces")
    e = Function(uh.function_space(), _lib.lincomb([u.tensor(), uh.tensor()], [1.0, -1.0]))
    return float(mass_product(e, e).sqrt())


def project(v, V=None, **kw):
    raise NotImplementedError("dolfin.project is outside the PGD hot path")
