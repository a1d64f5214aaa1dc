"""dolfin-like value objects used by the user's callbacks: Function / Vector, Constant, Expression,
DirichletBC, MeshFunction, SubDomain, interpolate.

The reference hands these DOLFIN objects to ``lhs_fct`` / ``rhs_fct`` / ``bc_fct`` / ``dom_fct``
(pgdrome/solver.py:136-156, 547-569).  Here a Function owns a float64 CUDA tensor of dof values;
the host only touches it when user code asks for ``.vector()[:]``.  Expressions are evaluated on
the host ONCE, at set-up (nodal interpolation / per-cell coefficient sampling), exactly where
DOLFIN would JIT-compile and tabulate them.
"""
import math
import re

import numpy as np
import torch

from . import _lib, sharding
from .fem import FunctionSpace, tabulate_lagrange
from . import lazy as _lazy
from .lazy import LazyScalar
from .ufl import Expr as _Expr
from .ufl import Leaf

DOLFIN_EPS = 3.0e-16
DOLFIN_PI = math.pi


def near(a, b, eps=DOLFIN_EPS):
    """dolfin.near: |a-b| <= eps * max(1, |a|, |b|)-ish absolute/relative closeness (vectorised)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    r = np.abs(a - b) <= eps * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
    return bool(r) if r.ndim == 0 else r


def _device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------- Constant
class Constant(Leaf):
    kind = "constant"

    def __init__(self, value):
        if isinstance(value, Constant):
            value = value.value
        if isinstance(value, (tuple, list, np.ndarray)) and np.ndim(value) >= 1:
            self.value = tuple(float(v) for v in np.asarray(value, dtype=np.float64).ravel())
            self._init_leaf(len(self.value))
        else:
            self.value = value if isinstance(value, LazyScalar) else float(value)
            self._init_leaf(1)

    def values(self):
        if isinstance(self.value, tuple):
            return np.array(self.value)
        return np.array([float(self.value)])

    def __float__(self):
        if isinstance(self.value, tuple):
            raise TypeError("vector Constant has no float value")
        return float(self.value)

    def assign(self, v):
        self.value = v.value if isinstance(v, Constant) else float(v)


# ------------------------------------------------------------------------------- Expression
_FUNCS = {
    "pow": np.power, "exp": np.exp, "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos, "tan": np.tan,
    "log": np.log, "fabs": np.abs, "abs": np.abs, "tanh": np.tanh, "sinh": np.sinh, "cosh": np.cosh,
    "atan": np.arctan, "atan2": np.arctan2, "asin": np.arcsin, "acos": np.arccos, "erf": None,
    "fmin": np.minimum, "fmax": np.maximum, "min": np.minimum, "max": np.maximum, "floor": np.floor,
    "ceil": np.ceil, "pi": math.pi, "DOLFIN_PI": math.pi, "DOLFIN_EPS": DOLFIN_EPS, "M_PI": math.pi,
    "where": np.where, "logical_and": np.logical_and, "logical_or": np.logical_or, "logical_not": np.logical_not,
}


def _split_top(s, seps):
    """Split s at the top-level (parenthesis depth 0) occurrences of any separator in seps."""
    out, depth, start, i = [], 0, 0, 0
    while i < len(s):
        ch = s[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif depth == 0:
            for sep in seps:
                if s.startswith(sep, i):
                    out.append((s[start:i], sep))
                    i += len(sep) - 1
                    start = i + 1
                    break
        i += 1
    out.append((s[start:], None))
    return out


def _cpp_to_py(s):
    """Translate the C++ snippet of a dolfin.Expression into a NumPy expression string.
    Supported: arithmetic, x[i], math calls, comparisons, && || !, the ternary operator."""
    s = s.strip()
    # ternary (lowest precedence, right associative)
    depth = 0
    for i, ch in enumerate(s):
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif ch == "?" and depth == 0:
            nest, d2 = 0, 0
            for j in range(i + 1, len(s)):
                c2 = s[j]
                if c2 in "([":
                    d2 += 1
                elif c2 in ")]":
                    d2 -= 1
                elif d2 == 0 and c2 == "?":
                    nest += 1
                elif d2 == 0 and c2 == ":":
                    if nest == 0:
                        return "where(%s, %s, %s)" % (_cpp_to_py(s[:i]), _cpp_to_py(s[i + 1:j]), _cpp_to_py(s[j + 1:]))
                    nest -= 1
            raise ValueError("unbalanced ternary in Expression '%s'" % s)
    parts = _split_top(s, ("||",))
    if len(parts) > 1:
        out = _cpp_to_py(parts[0][0])
        for p, _ in parts[1:]:
            out = "logical_or(%s, %s)" % (out, _cpp_to_py(p))
        return out
    parts = _split_top(s, ("&&",))
    if len(parts) > 1:
        out = _cpp_to_py(parts[0][0])
        for p, _ in parts[1:]:
            out = "logical_and(%s, %s)" % (out, _cpp_to_py(p))
        return out
    # recurse into parenthesised groups / call arguments
    out, i = "", 0
    while i < len(s):
        ch = s[i]
        if ch == "(":
            depth, j = 1, i + 1
            while j < len(s) and depth:
                depth += s[j] == "("
                depth -= s[j] == ")"
                j += 1
            inner = s[i + 1:j - 1]
            args = [a for a, _ in _split_top(inner, (",",))]
            out += "(" + ", ".join(_cpp_to_py(a) for a in args) + ")"
            i = j
        else:
            out += ch
            i += 1
    out = re.sub(r"\bx\[(\d)\]", r"x[..., \1]", out)
    out = re.sub(r"!(?!=)", " ~", out)
    out = re.sub(r"(\d)\.(?![\d])", r"\1.0", out)  # "1." -> "1.0" (harmless, keeps ints distinct)
    return out


class Expression(Leaf):
    """dolfin.Expression(cpp_code, degree=p, **user_parameters); tuple of strings = vector valued."""

    kind = "expression"

    def __init__(self, cppcode=None, degree=None, element=None, **params):
        if cppcode is None:
            raise NotImplementedError("Expression subclasses with eval(); use a C++ string or UserExpression")
        self._codes = (cppcode,) if isinstance(cppcode, str) else tuple(cppcode)
        if degree is None and element is not None:
            degree = element.degree()
        if degree is None:
            raise ValueError("Expression needs degree= or element=")
        self.degree = int(degree)
        object.__setattr__(self, "_params", dict(params))
        self._py = [compile(_cpp_to_py(c), "<Expression %s>" % c, "eval") for c in self._codes]
        self._version = 0
        self._init_leaf(len(self._codes))

    def __getattr__(self, name):
        p = self.__dict__.get("_params")
        if p is not None and name in p:
            return p[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        p = self.__dict__.get("_params")
        if p is not None and name in p:
            p[name] = value
            self.__dict__["_version"] += 1
        else:
            object.__setattr__(self, name, value)

    def eval_np(self, X, comp=None):
        """X [..., gdim] -> [...] (scalar / selected component) or [..., ncomp]."""
        X = np.asarray(X, dtype=np.float64)
        ns = dict(_FUNCS)
        for k, v in self._params.items():
            ns[k] = float(v) if isinstance(v, (Constant, LazyScalar)) else v
        ns["x"] = X
        outs = []
        idx = range(len(self._py)) if comp is None else [comp]
        for i in idx:
            v = eval(self._py[i], {"__builtins__": {}}, ns)
            outs.append(np.broadcast_to(np.asarray(v, dtype=np.float64), X.shape[:-1]).copy())
        if comp is not None or len(self._py) == 1:
            return outs[0]
        return np.stack(outs, axis=-1)

    def __call__(self, *x):
        x = np.atleast_1d(np.asarray(x[0] if len(x) == 1 else x, dtype=np.float64).ravel())
        v = self.eval_np(x[None, :])
        return float(v[0]) if v.ndim == 1 else v[0]


class UserExpression(Expression):
    """Subclass with ``eval_np(self, X)`` or ``eval(self, value, x)`` (slow, per point)."""

    def __init__(self, degree=1, n_comp=1, **kw):
        self.degree = int(degree)
        object.__setattr__(self, "_params", dict(kw))
        self._version = 0
        self._n = n_comp
        self._init_leaf(n_comp)

    def eval_np(self, X, comp=None):
        X = np.asarray(X, dtype=np.float64)
        flat = X.reshape(-1, X.shape[-1])
        out = np.zeros((flat.shape[0], self._n))
        for i, p in enumerate(flat):
            self.eval(out[i], p)
        out = out.reshape(X.shape[:-1] + (self._n,))
        if comp is not None:
            return out[..., comp]
        return out[..., 0] if self._n == 1 else out


# ------------------------------------------------------------------------------- MatrixOperator
class MatrixOperator(Leaf):
    """A user-assembled sparse operator on the dofs of a space (the reference's FD matrices
    ``M``, ``D2``, ``D1_up`` of pgdrome/solver.py:947-988, which its callbacks apply with host
    SciPy products: tests/integration/test_heat1D.py:291-306) lifted into the form language so that
    the whole sweep stays on the device::

        D1 = MatrixOperator(D1_up_t, V_t)
        a  = Constant(c) * D1(u, v) * dx(mesh_t)          # operator atom (row = v, col = u)
        s  = assemble(D1(G, F) * dx(mesh_t))              # mode integral  F^T D1 G

    The nonzeros must lie inside the Lagrange sparsity pattern of ``V`` (true for the tridiagonal FD
    operators on a 1-D P1 space); they are embedded once, at set-up."""

    kind = "operator"

    def __init__(self, A, V):
        import scipy.sparse as sp

        self.V = V
        self.A = sp.csr_matrix(A).astype(np.float64)
        if self.A.shape != (V.n_dofs, V.n_dofs):
            raise ValueError("operator shape %s does not match the space (%d dofs)" % (self.A.shape, V.n_dofs))
        self._version = 0
        self._init_leaf(1)

    def __call__(self, u, v):
        """Form integrand with u in the trial (column) role and v in the test (row) role."""
        u, v = _Expr.wrap(u), _Expr.wrap(v)
        if u is None or v is None or not (u.is_scalar and v.is_scalar):
            raise NotImplementedError("MatrixOperator needs scalar operands")
        return self * v * u


# ------------------------------------------------------------------------------- Function / Vector
class Vector:
    """Host-facing proxy of a Function's device dof vector (``f.vector()``)."""

    __array_priority__ = 100

    def __init__(self, owner):
        self._o = owner

    # -- reads (device -> host)
    def get_local(self):
        return self._o.values_host().copy()

    def __array__(self, dtype=None, copy=None):
        a = self.get_local()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, key):
        a = self._o.values_host()
        if isinstance(key, slice) and key == slice(None):
            return a.copy()
        r = a[key]
        return float(r) if np.ndim(r) == 0 else np.array(r)

    def __len__(self):
        return self._o.n_dofs

    def size(self):
        return self._o.n_dofs

    def norm(self, kind="l2"):
        if kind == "l2":
            return math.sqrt(_owned_dot(self._o, self._o))
        if kind == "linf":
            return float(np.abs(self._o.values_host()).max())
        raise NotImplementedError("vector norm '%s'" % kind)

    def inner(self, other):
        return _owned_dot(self._o, other._o)

    def max(self):
        return float(self._o.values_host().max())

    def min(self):
        return float(self._o.values_host().min())

    def sum(self):
        return float(self._o.values_host().sum())

    # -- writes (host -> device)
    def set_local(self, a):
        self._o.set_values(np.asarray(a, dtype=np.float64))

    def apply(self, mode="insert"):
        return None

    def __setitem__(self, key, value):
        if isinstance(value, Vector):
            value = value.get_local()
        if isinstance(key, slice) and key == slice(None):
            if np.ndim(value) == 0:
                self._o.before_write()
                self._o.tensor().fill_(float(value))
                self._o.touch()
            else:
                self._o.set_values(np.asarray(value, dtype=np.float64))
            return
        a = self._o.values_host().copy()
        a[key] = value
        self._o.set_values(a)

    def __imul__(self, c):
        self._o.scale(float(c))
        return self

    def __itruediv__(self, c):
        self._o.scale(1.0 / float(c))
        return self

    def __iadd__(self, other):
        self.axpy(1.0, other)
        return self

    def __isub__(self, other):
        self.axpy(-1.0, other)
        return self

    def axpy(self, a, other):
        if isinstance(other, Vector):
            o = other._o.tensor()
        elif self._o._shard is not None:
            o = self._o._shard.scatter(other, self._o.tensor().device)
        else:
            o = _lib.to_device(np.asarray(other, dtype=np.float64))
        t = self._o.tensor()
        self._o.before_write()
        _lib.lincomb([t, o], [1.0, float(a)], out=t)
        self._o.touch()

    def zero(self):
        self._o.before_write()
        self._o.tensor().zero_()
        self._o.touch()

    def copy(self):
        return DeviceVector(self._o.tensor().clone(), self._o._shard)

    def vec(self):
        return self

    # numpy-style arithmetic on host copies (used by post-processing code in the reference tests)
    def __mul__(self, c):
        return self.get_local() * c

    __rmul__ = __mul__

    def __sub__(self, o):
        return self.get_local() - np.asarray(o)

    def __add__(self, o):
        return self.get_local() + np.asarray(o)

    def transpose(self):
        return self.get_local()


def _owned_dot(a, b):
    """x . y over the global dofs of two dof owners (sharded spaces: local owned part + all-reduce)"""
    sh = a._shard
    if sh is None:
        return float(_lib.dot(a.tensor(), b.tensor()).item())
    no = sh.n_owned
    return float(sh.allreduce(_lib.dot(a.tensor()[:no], b.tensor()[:no])).item())


class _DofOwner:
    """Shared storage logic of Function and DeviceVector.  ``n_dofs`` is always the GLOBAL number of dofs; on an
    element-partitioned space (sharding.SpaceShard) the device tensor holds this rank's ``[owned | ghost]`` entries
    and the host-side accessors gather / scatter (collective calls: every rank executes the same script)."""

    def _init_store(self, n_dofs, tensor=None, shard=None):
        self.n_dofs = int(n_dofs)
        self._shard = shard
        self._t = tensor
        self._version = 0
        self._host = None
        self._host_version = -1

    @property
    def n_vec(self):
        """length of the device tensor"""
        return self._shard.n_local if self._shard is not None else self.n_dofs

    def tensor(self):
        if self._t is None:
            self._t = torch.zeros(self.n_vec, dtype=torch.float64, device=_device())
        return self._t

    def before_write(self):
        """Functionals recorded against the current values must be evaluated before they change."""
        if _lazy.has_pending():
            _lazy.launch()  # enqueued before the write in stream order; the values are read back when somebody asks

    def touch(self):
        self._version += 1

    def values_host(self):
        if self._host_version != self._version or self._host is None:
            t = self.tensor()
            self._host = self._shard.gather_host(t) if self._shard is not None else _lib.to_host(t)
            self._host_version = self._version
        return self._host

    def set_values(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
        if a.size != self.n_dofs:
            raise ValueError("size mismatch: %d values for %d dofs" % (a.size, self.n_dofs))
        self.before_write()
        t = self.tensor()
        t.copy_(self._shard.scatter(a, t.device) if self._shard is not None else _lib.to_device(a))
        self.touch()

    def set_tensor(self, t):
        """Adopt a device tensor (no copy)."""
        self.before_write()
        self._t = t
        self.touch()

    def scale(self, c):
        self.before_write()
        self.tensor().mul_(c)
        self.touch()


class DeviceVector(_DofOwner):
    """Result of assembling a rank-1 form: supports ``ll[:]``, ``bc.apply(ll)``, ``ll.get_local()``."""

    def __init__(self, tensor, shard=None):
        self._init_store(shard.n_global if shard is not None else tensor.numel(), tensor, shard)

    def vector(self):
        return Vector(self)

    def __getitem__(self, key):
        return Vector(self)[key]

    def get_local(self):
        return Vector(self).get_local()

    def __len__(self):
        return self.n_dofs

    def __array__(self, dtype=None, copy=None):
        return Vector(self).__array__(dtype)


class Function(Leaf, _DofOwner):
    kind = "function"

    def __init__(self, V, values=None, name=None):
        if isinstance(V, Function):  # dolfin.Function(other) copies
            values, V = V.tensor().clone(), V.V
        if not isinstance(V, FunctionSpace):
            raise TypeError("Function needs a FunctionSpace")
        self.V = V
        self._name = name or "f"
        self.stable = False  # set for stored modes / interpolated data: K*f products are cached in panels
        self._init_store(V.n_dofs, None, sharding.shard_of(V))
        if values is not None:
            if isinstance(values, torch.Tensor):
                if values.numel() != self.n_vec:
                    raise ValueError("Function: tensor of %d entries for a vector of %d" % (values.numel(), self.n_vec))
                self._t = values
            else:
                self.set_values(values)
        self._init_leaf(V.bs)

    def function_space(self):
        return self.V

    def vector(self):
        return Vector(self)

    def name(self):
        return self._name

    def rename(self, name, label=None):
        self._name = name

    def copy(self, deepcopy=True):
        g = Function(self.V, self.tensor().clone() if deepcopy else self.tensor())
        return g

    def assign(self, other):
        if isinstance(other, Function):
            self.before_write()
            self.tensor().copy_(other.tensor())
            self.touch()
        else:
            raise NotImplementedError("Function.assign(%r)" % type(other))

    def geometric_dimension(self):
        return self.V.mesh().gdim

    def value_dimension(self, i=0):
        return self.V.bs

    def value_rank(self):
        return 0 if self.V.bs == 1 else 1

    def compute_vertex_values(self, mesh=None):
        """Values at the mesh vertices, vertex order; vector fields component-major (DOLFIN layout)."""
        a = self.values_host().reshape(self.V.n_nodes, self.V.bs)
        v = a[self.V.vertex_to_node]
        return v[:, 0].copy() if self.V.bs == 1 else v.T.reshape(-1).copy()

    def __call__(self, *x):
        """Point evaluation (host, post-processing only)."""
        x = np.atleast_1d(np.asarray(x[0] if len(x) == 1 else x, dtype=np.float64).ravel())
        cell, ref = locate_point(self.V.mesh(), x)
        phi, _ = tabulate_lagrange(self.V.mesh().tdim, self.V.degree, ref[None, :])
        vals = self.values_host()[self.V.cell_dofs[cell]].reshape(self.V.nd, self.V.bs)
        r = phi[0] @ vals
        return float(r[0]) if self.V.bs == 1 else r

    # dolfin lets ``f.dx(0)`` etc. through Expr; ``f.vector()`` is the numeric side
    def ufl_shape_(self):
        return self.ufl_shape


def locate_point(mesh, x):
    """(cell index, reference coordinates) of a point; raises like DOLFIN when outside the mesh."""
    X = mesh.coordinates()[mesh.cells()]
    t = mesh.tdim
    if t == 1:
        a, b = X[:, 0, 0], X[:, 1, 0]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        tol = 1e-12 * max(1.0, float(np.abs(mesh.coordinates()).max()))
        cand = np.nonzero((x[0] >= lo - tol) & (x[0] <= hi + tol))[0]
        if cand.size == 0:
            raise RuntimeError("point %s is outside the mesh (allow_extrapolation is not supported)" % (x,))
        e = int(cand[0])
        return e, np.array([(x[0] - a[e]) / (b[e] - a[e])])
    J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)
    xi = np.linalg.solve(J, (x[None, :] - X[:, 0, :])[:, :, None])[:, :, 0]
    lam = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)
    ok = np.nonzero(lam.min(axis=1) >= -1e-10)[0]
    if ok.size == 0:
        raise RuntimeError("point %s is outside the mesh (allow_extrapolation is not supported)" % (x,))
    e = int(ok[0])
    return e, xi[e]


def interpolate(v, V):
    """Nodal interpolation of an Expression / Constant / Function into V (host, set-up)."""
    X = V.node_coords
    if isinstance(v, Expression):
        a = v.eval_np(X)
        if V.bs > 1:
            a = np.asarray(a)
            if a.ndim == 1:
                raise ValueError("scalar Expression interpolated into a vector space")
        vals = np.asarray(a, dtype=np.float64).reshape(V.n_nodes * V.bs)
    elif isinstance(v, Constant):
        vals = np.tile(v.values(), V.n_nodes)
    elif isinstance(v, (int, float)):
        vals = np.full(V.n_dofs, float(v))
    elif isinstance(v, Function):
        if v.V is V or (v.V.n_dofs == V.n_dofs and np.array_equal(v.V.cell_dofs, V.cell_dofs)):
            vals = v.values_host().copy()
        else:
            vals = np.array([np.atleast_1d(v(p)) for p in X]).reshape(-1)
    elif callable(v):
        vals = np.asarray(v(X), dtype=np.float64).reshape(V.n_nodes * V.bs)
    else:
        raise NotImplementedError("interpolate(%r)" % type(v))
    f = Function(V, vals)
    f.stable = True
    return f


# ------------------------------------------------------------------------------- sub-domains / BCs
class SubDomain:
    def inside(self, x, on_boundary):
        raise NotImplementedError

    def mark(self, mf, value):
        mf._mark(self.inside, value)


class CompiledSubDomain(SubDomain):
    def __init__(self, cppcode, **params):
        self._expr = compile(_cpp_to_py(cppcode), "<CompiledSubDomain>", "eval")
        self._params = params

    def inside(self, x, on_boundary):
        ns = dict(_FUNCS)
        ns.update(self._params)
        ns["x"] = _PointView(x)
        ns["on_boundary"] = on_boundary
        return eval(self._expr, {"__builtins__": {}}, ns)


class _PointView:
    """x such that x[..., i] == x[i] (lets the translated C++ index a single point or a batch)."""

    def __init__(self, x):
        self.x = x

    def __getitem__(self, k):
        if isinstance(k, tuple):
            k = k[-1]
        return self.x[k]


def _try_vectorised_inside(fn, X, flag):
    """inside(x, on_boundary) for all rows of X with a Python-bool `on_boundary`: `x[k]` becomes the
    coordinate array, so `on_boundary and near(x[0], 0.0)` short-circuits or broadcasts.  The result is
    accepted only if it agrees with pointwise calls on a sample (a callback that reduces over the
    coordinates would otherwise be mis-read)."""
    m = X.shape[0]
    try:
        r = np.asarray(fn(X.T, flag))
    except Exception:
        return None
    if r.dtype != bool and r.dtype != np.bool_:
        return None
    if r.shape == ():
        r = np.full(m, bool(r))
    elif r.shape == (1, m):
        r = r[0]
    elif r.shape != (m,):
        return None
    for i in np.unique(np.linspace(0, m - 1, min(m, 24)).astype(np.int64)):
        try:
            if bool(np.all(fn(X[i], flag))) != bool(r[i]):
                return None
        except Exception:
            return None
    return r


def _eval_inside(fn, X, onb):
    """Evaluate inside(x, on_boundary) for all points: vectorised per on_boundary group where the
    callback allows it, point by point otherwise (4 M-node meshes: seconds instead of a minute)."""
    n = X.shape[0]
    onb = np.asarray(onb, dtype=bool)
    if onb.ndim == 0:
        onb = np.full(n, bool(onb))
    out = np.zeros(n, dtype=bool)
    for flag in (False, True):
        idx = np.nonzero(onb == flag)[0]
        if not len(idx):
            continue
        r = _try_vectorised_inside(fn, X[idx], flag)
        if r is None:
            r = np.array([bool(np.all(fn(X[i], flag))) for i in idx], dtype=bool)
        out[idx] = r
    return out


class MeshFunction:
    """dolfin.MeshFunction("size_t", mesh, dim): cell markers (dim = tdim) or boundary-facet markers
    (dim = tdim-1; interior facets are not represented -- only ``ds`` integrals and DirichletBC use them)."""

    def __init__(self, dtype, mesh, dim, value=0):
        self._mesh, self._dim = mesh, int(dim)
        t = mesh.tdim
        if self._dim == t:
            self._vals = np.full(mesh.num_cells(), value, dtype=np.int64)
        elif self._dim == t - 1:
            self._cell, self._loc = mesh.boundary_facets()
            self._vals = np.full(len(self._cell), value, dtype=np.int64)
        else:
            raise NotImplementedError("MeshFunction of dimension %d on a %d-D mesh" % (dim, t))
        self._version = 0

    def mesh(self):
        return self._mesh

    def dim(self):
        return self._dim

    def set_all(self, v):
        self._vals[:] = v
        self._version += 1

    def array(self):
        return self._vals

    def _mark(self, inside, value):
        m = self._mesh
        X = m.coordinates()
        if self._dim == m.tdim:
            C = m.cells()
            ok = np.ones(len(C), dtype=bool)
            for k in range(C.shape[1]):
                ok &= _eval_inside(inside, X[C[:, k]], np.zeros(len(C), dtype=bool))
            ok &= _eval_inside(inside, X[C].mean(axis=1), np.zeros(len(C), dtype=bool))
        else:
            fv = m.facet_vertices(self._cell, self._loc)
            onb = np.ones(len(fv), dtype=bool)
            ok = np.ones(len(fv), dtype=bool)
            for k in range(fv.shape[1]):
                ok &= _eval_inside(inside, X[fv[:, k]], onb)
            ok &= _eval_inside(inside, X[fv].mean(axis=1), onb)
        self._vals[ok] = value
        self._version += 1

    def facets(self, value):
        sel = self._vals == value
        return self._cell[sel], self._loc[sel]


class DirichletBC:
    """DirichletBC(V, value, where[, marker]).  ``where``: callable(x, on_boundary), SubDomain, or a
    facet MeshFunction with ``marker``.  The dof set is computed once on the host (pointwise on the
    Lagrange nodes); applying it to operators / vectors happens on the device (pgd_apply_dirichlet)."""

    def __init__(self, V, value, where, marker=None, method="topological"):
        self.V = V
        self._value = value
        if isinstance(where, MeshFunction):
            if marker is None:
                raise ValueError("DirichletBC(V, g, mesh_function, marker)")
            cell, loc = where.facets(marker)
            nodes = np.unique(V.facet_nodes(cell, loc).ravel()) if len(cell) else np.zeros(0, dtype=np.int64)
        else:
            fn = where.inside if isinstance(where, SubDomain) else where
            sel = _eval_inside(fn, V.node_coords, V.node_on_boundary())
            nodes = np.nonzero(sel)[0]
        self.nodes = nodes.astype(np.int64)
        self.dofs = (self.nodes[:, None] * V.bs + np.arange(V.bs)[None, :]).ravel().astype(np.int32)
        self.vals = self._values_at(nodes)
        self._dev = None

    def _values_at(self, nodes):
        V, g = self.V, self._value
        n = len(nodes)
        if isinstance(g, Constant):
            v = g.values()
            return np.tile(v if len(v) == V.bs else np.repeat(v, V.bs), n) if n else np.zeros(0)
        if isinstance(g, (int, float, np.integer, np.floating)):
            return np.full(n * V.bs, float(g))
        if isinstance(g, (tuple, list)):
            return np.tile(np.asarray(g, dtype=np.float64), n)
        if isinstance(g, Expression):
            return np.asarray(g.eval_np(V.node_coords[nodes]), dtype=np.float64).reshape(-1)
        if isinstance(g, Function):
            return g.values_host().reshape(V.n_nodes, V.bs)[nodes].reshape(-1)
        raise NotImplementedError("DirichletBC value of type %r" % type(g))

    def function_space(self):
        return self.V

    def get_boundary_values(self):
        return {int(d): float(v) for d, v in zip(self.dofs, self.vals)}

    def homogeneous(self):
        return not np.any(self.vals)

    def device(self):
        """(dofs, values) on the device; on an element-partitioned space the dofs this rank holds (owned or ghost),
        in its local numbering."""
        if self._dev is None:
            dofs, vals = local_bc(self.V, self.dofs, self.vals)
            self._dev = (_lib.to_device(dofs), _lib.to_device(vals))
        return self._dev

    def apply(self, *objs):
        """bc.apply(vector): set the constrained dofs (on the device)."""
        for o in objs:
            owner = o._o if isinstance(o, Vector) else o
            if not isinstance(owner, _DofOwner):
                raise NotImplementedError("DirichletBC.apply on %r (use the variational solver for matrices)" % type(o))
            if len(self.dofs):
                d, v = self.device()
                if d.numel() == 0:
                    continue
                owner.before_write()
                _lib.set_entries(owner.tensor(), d, None if self.homogeneous() else v)
                owner.touch()


def local_bc(V, dofs, vals):
    """Dirichlet dofs / values restricted to what this rank holds of V (identity on replicated spaces)."""
    sh = sharding.shard_of(V)
    if sh is None:
        return np.asarray(dofs, dtype=np.int32), np.asarray(vals, dtype=np.float64)
    loc = sh.g2l_dofs(dofs)
    keep = loc >= 0
    return loc[keep].astype(np.int32), np.asarray(vals, dtype=np.float64)[keep]


def bc_list(bc):
    """Normalise the per-dimension entry returned by bc_fct: 0 | DirichletBC | [DirichletBC]."""
    if bc is None or (not isinstance(bc, (list, tuple, DirichletBC)) and bc == 0):
        return []
    if isinstance(bc, DirichletBC):
        return [bc]
    return [b for b in bc if isinstance(b, DirichletBC)]


def merged_bc_dofs(bcs, V):
    """Union of the bcs' dofs (later bcs win on duplicates): int32 dofs, float64 values (host)."""
    if not bcs:
        return np.zeros(0, dtype=np.int32), np.zeros(0)
    vals = {}
    for b in bcs:
        for d, v in zip(b.dofs, b.vals):
            vals[int(d)] = float(v)
    dofs = np.array(sorted(vals), dtype=np.int32)
    return dofs, np.array([vals[int(d)] for d in dofs])
