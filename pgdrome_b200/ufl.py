"""Mini-UFL: the symbolic layer that turns the user's *unchanged* lhs_fct / rhs_fct callbacks
(pgdrome/solver.py:547-569; examples tests/integration/*.py) into separated-form descriptors.

Every expression is kept in expanded form: a (vector of) sum(s) of monomials
``coef * prod(factors)`` with factor = (leaf, component, derivative index).  Leaves are test/trial
Arguments, Functions, Expressions and Constants.  Multiplying by a Measure gives a Form; the
device layer (forms.py) classifies each monomial as a registered atom:

    test x trial  -> bilinear atom tensor T[iv, jv, iu, ju]  (mass / stiffness / advection / Voigt ...)
    test x Function -> atom applied to a (cached) vector;  test only -> load vector
    Function x Function -> mode integral  F^T K G;  Function only -> F . load

Anything outside that registry raises NotImplementedError naming the construct (no fallback).
"""
from collections import namedtuple

import numpy as np

Factor = namedtuple("Factor", "leaf comp deriv")


class Mono:
    __slots__ = ("coef", "factors")

    def __init__(self, coef, factors=()):
        self.coef = float(coef)
        self.factors = tuple(factors)

    def times(self, other):
        return _mono(self.coef * other.coef, self.factors + other.factors)

    def scaled(self, c):
        return _mono(self.coef * c, self.factors)


def _mono(coef, factors, _new=Mono.__new__, _cls=Mono):
    """Mono from an already-float coefficient and an already-tuple factor list (the capture layer builds tens of
    thousands of monomials per enrichment step: no conversions on this path)"""
    m = _new(_cls)
    m.coef = coef
    m.factors = factors
    return m


_LAZY_SCALAR = None

# While an enrichment step runs, the solver sets capture_cache[0] to a dict: `leaf[i]` of a vector leaf is then handed out
# as ONE shared expression object per (leaf, i), and that object remembers its derivatives -- the `eps(v)` / `grad(v)`
# algebra of the user's callbacks is evaluated hundreds of times per sweep on the same few Functions.  Expressions are
# immutable, and the table (which keeps its leaves alive) is dropped at the end of the step, so no reference cycle
# leaf -> cache -> expression -> leaf outlives it.
capture_cache = [None]


def _lazy_scalar_type():
    """LazyScalar, resolved once (a function-level import here costs ~1.5 us on every operator call)"""
    global _LAZY_SCALAR
    if _LAZY_SCALAR is None:
        from .lazy import LazyScalar

        _LAZY_SCALAR = LazyScalar
    return _LAZY_SCALAR


def _is_number(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


class Expr:
    """Expanded symbolic expression. shape () -> comps = [monos]; shape (n,) -> comps = [monos]*n."""

    __array_ufunc__ = None  # numpy scalars defer to our reflected operators
    __array_priority__ = 1000

    _dx = None  # {j: d/dx_j of this expression}, filled on demand for the shared component expressions (capture_cache)

    def __init__(self, shape, comps):
        self.ufl_shape = tuple(shape)
        self.comps = comps

    # ---- helpers
    @staticmethod
    def scalar(monos):
        return Expr((), [list(monos)])

    @staticmethod
    def wrap(x):
        if isinstance(x, Expr):
            return x
        if _is_number(x):
            return Expr.scalar([Mono(float(x))])
        if isinstance(x, _lazy_scalar_type()):
            return Expr.scalar([Mono(1.0, (Factor(_LazyLeaf(x), None, None),))])
        if isinstance(x, np.ndarray) and x.ndim == 0:
            return Expr.scalar([Mono(float(x))])
        return None

    @property
    def is_scalar(self):
        return self.ufl_shape == ()

    # ---- algebra
    def __add__(self, other):
        if isinstance(other, Form):
            return NotImplemented
        o = Expr.wrap(other)
        if o is None:
            return NotImplemented
        if _is_number(other) and other == 0:
            return self
        if o.ufl_shape != self.ufl_shape:
            raise ValueError("shape mismatch in sum: %s vs %s" % (self.ufl_shape, o.ufl_shape))
        return Expr(self.ufl_shape, [a + b for a, b in zip(self.comps, o.comps)])

    __radd__ = __add__

    def __neg__(self):
        return Expr(self.ufl_shape, [[m.scaled(-1.0) for m in c] for c in self.comps])

    def __sub__(self, other):
        o = Expr.wrap(other)
        if o is None:
            return NotImplemented
        return self + (-o)

    def __rsub__(self, other):
        o = Expr.wrap(other)
        if o is None:
            return NotImplemented
        return o + (-self)

    def __mul__(self, other):
        if isinstance(other, (Measure, Form)):
            return NotImplemented
        if isinstance(other, Leaf) and other._n_comp == 1 and isinstance(self, Leaf) and self._n_comp == 1:
            # scalar leaf * scalar leaf (F*G, Constant*F, ...): the product monomial directly
            return Expr((), [[_mono(1.0, (Factor(self, None, None), Factor(other, None, None)))]])
        o = Expr.wrap(other)
        if o is None:
            return NotImplemented
        if self.is_scalar:
            return Expr(o.ufl_shape, [[a.times(b) for a in self.comps[0] for b in c] for c in o.comps])
        if o.is_scalar:
            return Expr(self.ufl_shape, [[a.times(b) for a in c for b in o.comps[0]] for c in self.comps])
        raise NotImplementedError("product of two vector expressions: use inner() or dot()")

    def __rmul__(self, other):
        o = Expr.wrap(other)
        if o is None:
            return NotImplemented
        return o.__mul__(self)

    def __truediv__(self, other):
        if _is_number(other):
            return self * (1.0 / float(other))
        o = Expr.wrap(other)
        if o is not None and o.is_scalar and len(o.comps[0]) == 1 and not o.comps[0][0].factors:
            return self * (1.0 / o.comps[0][0].coef)
        raise NotImplementedError("division by a non-constant expression")

    def __pos__(self):
        return self

    def __getitem__(self, i):
        if self.is_scalar:
            raise IndexError("indexing a scalar expression")
        if isinstance(i, tuple):
            if len(i) != 1:
                raise NotImplementedError("rank-2 indexing")
            i = i[0]
        return Expr.scalar(self.comps[int(i)])

    def __len__(self):
        if self.is_scalar:
            raise TypeError("scalar expression has no len()")
        return self.ufl_shape[0]

    def __iter__(self):
        if self.is_scalar:
            raise TypeError("scalar expression is not iterable")
        return (self[i] for i in range(self.ufl_shape[0]))

    def dx(self, *idx):
        if len(idx) != 1:
            raise NotImplementedError("higher derivatives .dx(i, j)")
        j = int(idx[0])
        memo = self._dx
        if memo is not None:
            hit = memo.get(j)
            if hit is not None:
                return hit
        out = []
        for c in self.comps:
            monos = []
            for m in c:
                fs = m.factors
                if len(fs) == 1 and fs[0].deriv is None and fs[0].leaf.kind in ("function", "argument"):
                    monos.append(_mono(m.coef, (Factor(fs[0].leaf, fs[0].comp, j),)))  # d/dx_j of  c * leaf[comp]
                    continue
                for k, f in enumerate(fs):
                    kind = f.leaf.kind
                    if kind in ("constant", "lazy"):
                        continue
                    if kind == "operator":
                        raise NotImplementedError("derivative of a MatrixOperator term")
                    if kind == "expression":
                        raise NotImplementedError("derivative of an Expression (interpolate it into a Function first)")
                    if f.deriv is not None:
                        raise NotImplementedError("second derivatives")
                    nf = m.factors[:k] + (Factor(f.leaf, f.comp, j),) + m.factors[k + 1:]
                    monos.append(Mono(m.coef, nf))
            out.append(monos)
        res = Expr(self.ufl_shape, out)
        if memo is not None:
            memo[j] = res
        return res

    def __eq__(self, other):  # noqa: keep identity semantics for leaves used as dict keys
        return self is other

    def __hash__(self):
        return id(self)


class Leaf(Expr):
    """Terminal: subclasses set ``kind`` in {"argument","function","expression","constant","lazy"}."""

    kind = "leaf"

    def _init_leaf(self, n_comp):
        self._n_comp = 1 if n_comp in (None, 0, 1) else int(n_comp)
        self.ufl_shape = () if self._n_comp == 1 else (self._n_comp,)

    def __getitem__(self, i):
        if self._n_comp == 1:
            raise IndexError("indexing a scalar expression")
        if isinstance(i, tuple):
            if len(i) != 1:
                raise NotImplementedError("rank-2 indexing")
            i = i[0]
        i = int(i)
        if not 0 <= i < self._n_comp:
            raise IndexError("component %d of a %d-vector" % (i, self._n_comp))
        cache = capture_cache[0]
        if cache is None:
            return Expr((), [[_mono(1.0, (Factor(self, i, None),))]])
        ent = cache.get((id(self), i))
        if ent is None or ent[0] is not self:
            e = Expr((), [[_mono(1.0, (Factor(self, i, None),))]])
            e._dx = {}
            ent = cache[(id(self), i)] = (self, e)
        return ent[1]

    @property
    def comps(self):
        """The leaf as a polynomial in itself -- built per use, not stored: a stored copy would make every Function /
        Constant / Expression a reference cycle (leaf -> monomial -> factor -> leaf), i.e. device vectors that only
        the cycle collector can free."""
        if self._n_comp == 1:
            return [[_mono(1.0, (Factor(self, None, None),))]]
        return [[_mono(1.0, (Factor(self, i, None),))] for i in range(self._n_comp)]


class _LazyLeaf(Leaf):
    kind = "lazy"

    def __init__(self, lazy):
        self.lazy = lazy
        self._init_leaf(1)


class Argument(Leaf):
    kind = "argument"

    def __init__(self, V, number):
        self.V, self.number = V, number
        self._init_leaf(V.bs)

    def function_space(self):
        return self.V


def TestFunction(V):
    return Argument(V, 0)


def TrialFunction(V):
    return Argument(V, 1)


class ConstantMatrix:
    """as_matrix(ndarray of numbers): only the product with a vector expression is needed
    (C * eps(u), tests/integration/test_solver_problem.py:148)."""

    __array_ufunc__ = None

    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)
        if self.a.ndim != 2:
            raise ValueError("as_matrix needs a 2-D array")
        self._rows = [[(j, float(v)) for j, v in enumerate(row) if v != 0.0] for row in self.a]

    def __mul__(self, v):
        v = Expr.wrap(v)
        if v is None or v.is_scalar or v.ufl_shape[0] != self.a.shape[1]:
            raise NotImplementedError("matrix * non-conforming operand")
        vc = v.comps
        return Expr((self.a.shape[0],), [[m.scaled(c) for j, c in row for m in vc[j]] for row in self._rows])


def as_matrix(rows):
    try:
        return ConstantMatrix(np.asarray(rows, dtype=np.float64))
    except (TypeError, ValueError):
        raise NotImplementedError("as_matrix of non-numeric entries")


def as_vector(items):
    comps = []
    for it in items:
        e = Expr.wrap(it)
        if e is None or not e.is_scalar:
            raise NotImplementedError("as_vector of non-scalar entries")
        comps.append(list(e.comps[0]))
    return Expr((len(comps),), comps)


def inner(a, b):
    a, b = Expr.wrap(a), Expr.wrap(b)
    if a.ufl_shape != b.ufl_shape:
        raise ValueError("inner: shape mismatch %s vs %s" % (a.ufl_shape, b.ufl_shape))
    if a.is_scalar:
        return a * b
    return Expr((), [[x.times(y) for ca, cb in zip(a.comps, b.comps) for x in ca for y in cb]])


dot = inner


def grad(u):
    u = Expr.wrap(u)
    if not u.is_scalar:
        raise NotImplementedError("grad of a vector expression (use v[i].dx(j))")
    g = _gdim_of(u)
    return as_vector([u.dx(j) for j in range(g)])


def _gdim_of(e):
    for c in e.comps:
        for m in c:
            for f in m.factors:
                V = getattr(f.leaf, "V", None)
                if V is not None:
                    return V.mesh().gdim
    raise ValueError("cannot infer the geometric dimension of the expression")


# ------------------------------------------------------------------------------- measures / forms
class Measure:
    def __init__(self, kind="dx", domain=None, subdomain_data=None, subdomain_id=None):
        self.kind, self.domain, self.subdomain_data, self.subdomain_id = kind, domain, subdomain_data, subdomain_id

    def __call__(self, *args, **kw):
        domain, sid = self.domain, self.subdomain_id
        data = kw.get("subdomain_data", self.subdomain_data)
        if "domain" in kw:
            domain = kw["domain"]
        for a in args:
            if _is_number(a):
                sid = int(a)
            elif hasattr(a, "num_cells"):
                domain = a
            else:
                raise NotImplementedError("measure argument %r" % (a,))
        return Measure(self.kind, domain, data, sid)

    def __rmul__(self, integrand):
        e = Expr.wrap(integrand)
        if e is None or not e.is_scalar:
            raise ValueError("can only integrate scalar expressions")
        return Form([Integral(list(e.comps[0]), self)])


dx = Measure("dx")
ds = Measure("ds")


class Integral:
    def __init__(self, monos, measure):
        self.monos, self.measure = monos, measure


class Form:
    __array_ufunc__ = None

    def __init__(self, integrals):
        self.integrals = integrals

    def __add__(self, other):
        if _is_number(other) and other == 0:
            return self
        if not isinstance(other, Form):
            return NotImplemented
        return Form(self.integrals + other.integrals)

    __radd__ = __add__

    def __neg__(self):
        return Form([Integral([m.scaled(-1.0) for m in it.monos], it.measure) for it in self.integrals])

    def __sub__(self, other):
        if _is_number(other) and other == 0:
            return self
        if not isinstance(other, Form):
            return NotImplemented
        return self + (-other)

    def __rsub__(self, other):
        if _is_number(other) and other == 0:
            return -self
        return NotImplemented

    def __rmul__(self, c):
        e = Expr.wrap(c)
        if e is None or not e.is_scalar:
            return NotImplemented
        return Form([Integral([a.times(m) for a in e.comps[0] for m in it.monos], it.measure) for it in self.integrals])

    __mul__ = __rmul__

    def __eq__(self, other):
        return Equation(self, other)

    __hash__ = object.__hash__

    def arguments(self):
        seen = {}
        for it in self.integrals:
            for m in it.monos:
                for f in m.factors:
                    if f.leaf.kind == "argument":
                        seen[f.leaf.number] = f.leaf
        return [seen[k] for k in sorted(seen)]

    def rank(self):
        return len(self.arguments())


class Equation:
    def __init__(self, lhs, rhs):
        self.lhs, self.rhs = lhs, rhs


def _has_trial(m):
    return any(f.leaf.kind == "argument" and f.leaf.number == 1 for f in m.factors)


def lhs(F):
    return Form([Integral([m for m in it.monos if _has_trial(m)], it.measure) for it in F.integrals])


def rhs(F):
    return Form([Integral([m.scaled(-1.0) for m in it.monos if not _has_trial(m)], it.measure) for it in F.integrals])


def derivative(F, u, du=None):
    """Gateaux derivative of a form w.r.t. the Function u (direction = TrialFunction)."""
    du = du if du is not None else TrialFunction(u.function_space())
    out = []
    for it in F.integrals:
        monos = []
        for m in it.monos:
            for k, f in enumerate(m.factors):
                if f.leaf is u:
                    monos.append(Mono(m.coef, m.factors[:k] + (Factor(du, f.comp, f.deriv),) + m.factors[k + 1:]))
        out.append(Integral(monos, it.measure))
    return Form(out)


def replace_function(F, u, value_zero=True):
    """Monomials of F that do NOT contain the Function u (the part of a residual a(u)-l that is
    independent of the unknown)."""
    out = []
    for it in F.integrals:
        out.append(Integral([m for m in it.monos if not any(f.leaf is u for f in m.factors)], it.measure))
    return Form(out)
