"""Golden-vector generator infrastructure (NOT product code, NOT the oracle).

A NumPy-only stand-in for the handful of DOLFIN calls that the reference's *finite-difference*
code paths touch, so that the UNMODIFIED reference modules (pgdrome/solver.py, pgdrome/model.py)
and the reference's own test callbacks (tests/integration/test_laplace.py FD callbacks,
tests/unit/test_pgdclass.py) can be imported and executed in a container without FEniCS.

Only 1-D P1 interval meshes are modelled.  Everything the reference computes on these paths is
NumPy/SciPy arithmetic inside the reference itself (FD_matrices, spsolve, the enrichment loop,
the normalisation, the stopping tests, interp1d evaluation); the stub supplies containers only:

  IntervalMesh / FunctionSpace("CG", 1)   vertices a+i(b-a)/n; dof i <-> vertex n-i (DOLFIN's 1-D
                                           serial numbering, implied by tests/unit/test_FD.py:68-79)
  Function                                 dof vector with [:] access, axpy, compute_vertex_values
  Expression (C strings) + interpolate     nodal interpolation; `c ? a : b` -> numpy.where
  DirichletBC(V, value, inside).apply(vec) sets the boundary dofs
  norm(f)                                  sqrt(f^T M f), consistent P1 mass (exact integration)

Anything else raises, so a golden vector can never silently come from an un-modelled DOLFIN path.
"""
import re
import types

import numpy as np

DOLFIN_EPS = 3.0e-16


class _Topology:
    def __init__(self, d):
        self._d = d

    def dim(self):
        return self._d


class IntervalMesh:
    def __init__(self, n, a, b):
        self._x = (a + (b - a) * np.arange(n + 1) / n).reshape(-1, 1).astype(np.float64)
        self._cells = np.stack([np.arange(n), np.arange(1, n + 1)], 1)

    def coordinates(self):
        return self._x

    def cells(self):
        return self._cells

    def num_cells(self):
        return len(self._cells)

    def num_vertices(self):
        return len(self._x)

    def topology(self):
        return _Topology(1)

    def geometry(self):
        return types.SimpleNamespace(dim=lambda: 1)


class _Element:
    def __str__(self):
        return "<CG1 on a interval>"

    def degree(self):
        return 1


class FunctionSpace:
    def __init__(self, mesh, family, degree):
        if family not in ("CG", "Lagrange", "P") or degree != 1 or not isinstance(mesh, IntervalMesh):
            raise NotImplementedError("dolfin stub: only P1 on IntervalMesh")
        self._mesh = mesh
        n = mesh.num_vertices()
        self._v2d = (n - 1) - np.arange(n)  # vertex -> dof
        self._d2v = (n - 1) - np.arange(n)  # dof -> vertex

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._mesh.num_vertices()

    def tabulate_dof_coordinates(self):
        return self._mesh.coordinates()[self._d2v]

    def ufl_function_space(self):
        return self

    def ufl_element(self):
        return _Element()


class _Vector:
    def __init__(self, n):
        self.a = np.zeros(n)

    def __getitem__(self, k):
        return self.a[k]

    def __setitem__(self, k, v):
        self.a[k] = v

    def __len__(self):
        return len(self.a)

    def axpy(self, alpha, other):
        self.a += alpha * other.a

    def get_local(self):
        return self.a.copy()

    def size(self):
        return len(self.a)


class Function:
    def __init__(self, V):
        self._V = V
        self._vec = _Vector(V.dim())

    def vector(self):
        return self._vec

    def function_space(self):
        return self._V

    def compute_vertex_values(self, mesh=None):
        return self._vec.a[self._V._v2d].copy()

    def __call__(self, x):
        xs = self._V.mesh().coordinates()[:, 0]
        x = float(np.atleast_1d(x)[0])
        if x < xs[0] - 1e-12 or x > xs[-1] + 1e-12:
            raise RuntimeError("point outside mesh")
        return float(np.interp(x, xs, self.compute_vertex_values()))


_TERNARY = re.compile(r"^(?P<c>[^?]+)\?(?P<a>[^:]+):(?P<b>.+)$")


class Expression:
    def __init__(self, code, degree=None, element=None, **params):
        if not isinstance(code, str):
            raise NotImplementedError("dolfin stub: scalar C-string expressions only")
        self.code, self.params = code, params

    def eval_at(self, x):
        env = {"x": [x], "np": np, "exp": np.exp, "sqrt": np.sqrt, "pow": np.power, "pi": np.pi}
        env.update(self.params)
        m = _TERNARY.match(self.code)
        if m:
            c, a, b = (eval(m.group(k).strip(), env) for k in ("c", "a", "b"))
            return np.where(c, a, b) * np.ones_like(x)
        return eval(self.code, env) * np.ones_like(x)


class Constant:
    def __init__(self, v):
        self.v = float(v)

    def eval_at(self, x):
        return self.v * np.ones_like(x)


def interpolate(expr, V):
    f = Function(V)
    f.vector()[:] = expr.eval_at(V.tabulate_dof_coordinates()[:, 0])
    return f


def near(a, b, eps=DOLFIN_EPS):
    return abs(a - b) <= eps


class DirichletBC:
    def __init__(self, V, value, inside):
        self._V, self._value = V, value
        xs = V.tabulate_dof_coordinates()
        x0, x1 = V.mesh().coordinates()[0, 0], V.mesh().coordinates()[-1, 0]
        self._dofs = np.array([i for i in range(V.dim()) if inside(xs[i], bool(xs[i, 0] in (x0, x1)))], dtype=np.int64)

    def apply(self, vec):
        val = self._value.v if isinstance(self._value, Constant) else float(self._value)
        vec[self._dofs] = val

    def get_boundary_values(self):
        return {int(d): 0.0 for d in self._dofs}


def p1_mass(V):
    x = V.mesh().coordinates()[:, 0]
    n = len(x)
    M = np.zeros((n, n))
    for e in range(n - 1):
        h = x[e + 1] - x[e]
        d = V._v2d[[e, e + 1]]
        M[np.ix_(d, d)] += h / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]])
    return M


def norm(f, norm_type="L2"):
    if norm_type.lower() != "l2":
        raise NotImplementedError
    a = f.vector()[:]
    return float(np.sqrt(a @ (p1_mass(f.function_space()) @ a)))


class TestFunction:  # placeholder handed to FD callbacks, which ignore it
    __test__ = False

    def __init__(self, V):
        self._V = V


TrialFunction = TestFunction


def _unmodelled(name):
    def f(*a, **k):
        raise NotImplementedError("dolfin stub: %s is not modelled (FEM path needs real FEniCS)" % name)

    return f


for _n in ("assemble", "solve", "derivative", "inner", "dx", "LinearVariationalProblem", "LinearVariationalSolver",
           "NonlinearVariationalProblem", "NonlinearVariationalSolver", "XDMFFile", "HDF5File", "errornorm", "project"):
    globals()[_n] = _unmodelled(_n)
MPI = types.SimpleNamespace(comm_world=None)
