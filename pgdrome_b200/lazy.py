"""Deferred scalars: ``assemble(<functional>)`` returns a LazyScalar instead of a Python float.

The reference executes ~2*(D-1)*(K+M+n*K) scalar ``dolfin.assemble`` calls per dimension per
fixed-point sweep, each a full mesh pass and a host round trip.  Here the callback
only *records* the functional; when a value is finally needed (operator / rhs assembly, float()),
all pending functionals are evaluated together -- one batched panel-dot launch per
(operator atom, vector) pair -- and read back with ONE device->host copy.  Products / sums of
LazyScalars are evaluated on the host in float64 in the order the callback wrote them, i.e. the
association order of the reference.
"""
import math

import numpy as np

_pending = []  # leaves recorded by a callback, nothing launched yet
_launched = []  # leaves whose kernels are enqueued (value in the device scalar pool, leaf._dev = slot), not read back
_launch_hook = [None]  # set by forms.py: callable(list_of_leaves) -> enqueues the kernels, sets leaf._dev
_fetch_hook = [None]  # set by forms.py: callable(list_of_leaves) -> ONE device->host copy, fills leaf._value
stats = {"flushes": 0, "leaves": 0, "fetches": 0}


def _num(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


class LazyScalar:
    __array_ufunc__ = None
    __array_priority__ = 2000
    __slots__ = ("op", "args", "_value", "_dev", "__weakref__")

    def __init__(self, op, args, value=None):
        self.op, self.args, self._value, self._dev = op, args, value, None
        if op == "leaf" and value is None:
            _pending.append(self)

    # ---- evaluation
    @property
    def value(self):
        if self._value is None:
            self._value = self._eval()
        return self._value

    def _eval(self):
        op, a = self.op, self.args
        if op == "leaf":
            flush()
            if self._value is None:
                raise RuntimeError("functional was not evaluated by flush()")
            return self._value
        if op == "const":
            return float(a[0])
        v = [x.value if isinstance(x, LazyScalar) else float(x) for x in a]
        if op == "mul":
            return v[0] * v[1]
        if op == "add":
            return v[0] + v[1]
        if op == "sub":
            return v[0] - v[1]
        if op == "div":
            return v[0] / v[1]
        if op == "neg":
            return -v[0]
        if op == "pow":
            return v[0] ** v[1]
        if op == "sqrt":
            return math.sqrt(v[0])
        if op == "abs":
            return abs(v[0])
        raise ValueError(op)

    def __float__(self):
        return float(self.value)

    def __repr__(self):
        return "LazyScalar(%s)" % ("%.17g" % self._value if self._value is not None else self.op)

    # ---- arithmetic (scalars stay lazy; arrays / sparse matrices force the value)
    def _bin(self, op, other, swap=False):
        if isinstance(other, LazyScalar) or _num(other):
            return LazyScalar(op, (other, self) if swap else (self, other))
        if isinstance(other, np.ndarray) and other.ndim == 0:
            return self._bin(op, float(other), swap)
        return None

    def __mul__(self, other):
        r = self._bin("mul", other)
        if r is not None:
            return r
        from .ufl import Expr, Form

        if isinstance(other, (Expr, Form)):
            return NotImplemented
        return float(self) * other  # ndarray, scipy.sparse matrix, ...

    def __rmul__(self, other):
        r = self._bin("mul", other, swap=True)
        if r is not None:
            return r
        return other * float(self)

    def __add__(self, other):
        r = self._bin("add", other)
        return r if r is not None else float(self) + other

    def __radd__(self, other):
        r = self._bin("add", other, swap=True)
        return r if r is not None else other + float(self)

    def __sub__(self, other):
        r = self._bin("sub", other)
        return r if r is not None else float(self) - other

    def __rsub__(self, other):
        r = self._bin("sub", other, swap=True)
        return r if r is not None else other - float(self)

    def __truediv__(self, other):
        r = self._bin("div", other)
        return r if r is not None else float(self) / other

    def __rtruediv__(self, other):
        r = self._bin("div", other, swap=True)
        return r if r is not None else other / float(self)

    def __pow__(self, p):
        return LazyScalar("pow", (self, p))

    def __neg__(self):
        return LazyScalar("neg", (self,))

    def __abs__(self):
        return LazyScalar("abs", (self,))

    def sqrt(self):
        return LazyScalar("sqrt", (self,))

    # comparisons force the value (used by user code such as ``if res < tol``)
    def __lt__(self, o):
        return float(self) < float(o)

    def __le__(self, o):
        return float(self) <= float(o)

    def __gt__(self, o):
        return float(self) > float(o)

    def __ge__(self, o):
        return float(self) >= float(o)


def constant(v):
    return LazyScalar("const", (float(v),), float(v))


def launch():
    """Enqueue the device evaluation of every recorded functional (no host synchronisation): afterwards the values
    exist in stream order in the device scalar pool, where device-side consumers (forms._coefs_dev) read them."""
    global _pending
    if not _pending:
        return
    leaves, _pending = _pending, []
    leaves = [l for l in leaves if l._value is None and l._dev is None]
    if not leaves:
        return
    if _launch_hook[0] is None:
        raise RuntimeError("no device evaluator registered (pgdrome_b200.forms not imported)")
    stats["flushes"] += 1
    stats["leaves"] += len(leaves)
    _launch_hook[0](leaves)
    _launched.extend(leaves)


def fetch():
    """Read the launched functionals back with ONE device->host copy (synchronises)."""
    global _launched
    if not _launched:
        return
    leaves, _launched = _launched, []
    leaves = [l for l in leaves if l._value is None]
    if leaves:
        stats["fetches"] += 1
        _fetch_hook[0](leaves)


def flush():
    """Evaluate every pending functional on the device in one batch and bring the values to the host."""
    launch()
    fetch()


def has_pending():
    return bool(_pending)
