"""Form compiler + device evaluator: turns the expanded mini-UFL forms produced by the user's
callbacks into calls of the libpgdb200 kernels.

Classification of every group of monomials (same scalar coefficients, weights, operands, measure):

  rank 2  test x trial [x weights]        -> operator atom  K[T, w]   (assembled once, cached)
  rank 1  test x Function [x weights]     -> K[T, w] @ f              (SpMV, cached for stable f)
          test [x weights]                -> load vector              (assembled once, cached)
  rank 0  Function x Function [x weights] -> f1^T K[T, w] f2          (mode integral: fused
                                             SpMV-reduction, or a panel dot with the cached K @ f)
          Function [x weights]            -> load . f
          [weights]                       -> sum(load)

T[iv, jv, iu, ju] is the constant form tensor (slot 0 = value, 1+m = d/dx_m; first pair = test /
first operand, second pair = trial / second operand).  This is the separated-form registry of
SURVEY.md 2.4: mass, weighted mass, stiffness, first-derivative (advection/time), Voigt elasticity,
volume / lifting / boundary-traction loads.  Anything else raises NotImplementedError.

The per-dimension system of the fixed-point sweep (pgdrome/solver.py:547-720) is then
    A = sum_g coef_g * K_g          (pgd_lincomb over CSR value arrays)
    b = sum_g coef_g * vec_g        (pgd_lincomb over cached vectors)
with coef_g = float constants x the mode integrals of the other dimensions (LazyScalars, evaluated
in one batch per system, one device->host copy).
"""
import math
import weakref

import numpy as np
import torch

from . import _lib, lazy
from .assembly import device_space
from .functions import DeviceVector, Expression, Function, _DofOwner, _device
from .lazy import LazyScalar
from .ufl import Form

MAX_PANEL_ROWS = 8
functional_memo = [None]  # dict while an enrichment step runs (see _fast_functional), else None
panel_rows_hint = [MAX_PANEL_ROWS]  # first allocation of a product panel (the solver sets it from PGD_nmax)


# ------------------------------------------------------------------------------- compile (host)
class Group:
    """Monomials of one integral that share scalar leaves, weights, operands and measure."""

    __slots__ = ("scalars", "weights", "operands", "measure", "entries", "space", "has_test", "has_trial", "op", "_sign")

    def __init__(self, scalars, weights, operands, measure, space, has_test, has_trial, op=None):
        self.scalars, self.weights, self.operands, self.measure = scalars, weights, operands, measure
        self.space, self.has_test, self.has_trial, self.op = space, has_test, has_trial, op
        self.entries = {}  # (iv, jv, iu, ju) | (iv, jv) | () -> float
        self._sign = None

    @property
    def rank(self):
        return int(self.has_test) + int(self.has_trial)

    def sign(self):
        """+-1: sign of the first non-zero tensor entry (C order).  The tensor is handed out with that sign divided out
        and the coefficient carries it, so that K and -K (the `-Constant(c) * op(G, v)` terms of every right-hand
        side) are ONE assembled atom with one panel of cached products instead of two."""
        if self._sign is None:
            sg = 1.0
            if int(self.has_test) + int(self.has_trial) + len(self.operands) > 0:
                nz = [k for k, v in self.entries.items() if v != 0.0]
                if nz and self.entries[min(nz)] < 0.0:
                    sg = -1.0
            self._sign = sg
        return self._sign

    def tensor(self):
        """Dense form tensor: T [bs,g+1,bs,g+1] (two slots), L [bs,g+1] (one slot) or a float; sign normalised (sign())."""
        bs, g = self.space.bs, self.space.mesh().gdim
        nslots = int(self.has_test) + int(self.has_trial) + len(self.operands)
        if nslots == 2:
            T = np.zeros((bs, g + 1, bs, g + 1))
        elif nslots == 1:
            T = np.zeros((bs, g + 1))
        else:
            return float(sum(self.entries.values()))
        sg = self.sign()
        for k, v in self.entries.items():
            T[k] += sg * v
        return T

    def coefficient(self):
        """float constant part is folded into the tensor; this is the product of the scalar leaves (times sign())."""
        c = self.sign()
        for s in self.scalars:
            c = c * _scalar_value(s)
        return c


def _scalar_value(leaf):
    if leaf.kind == "lazy":
        return leaf.lazy
    return leaf.value  # scalar Constant: float or LazyScalar


def _slot(deriv):
    return 0 if deriv is None else 1 + int(deriv)


def _same_space(A, B):
    return A is B or (A.mesh() is B.mesh() and A.degree == B.degree and A.bs == B.bs)


def compile_form(form):
    """Form -> list of Groups (pure host work, no device access)."""
    groups = {}
    order = []
    for it in form.integrals:
        meas = it.measure
        mkey = (meas.kind, id(meas.subdomain_data) if meas.subdomain_data is not None else None, meas.subdomain_id)
        for m in it.monos:
            if m.coef == 0.0:
                continue
            coef = m.coef
            test = trial = op = None
            scalars, weights, fns = [], [], []
            for f in m.factors:
                k = f.leaf.kind
                if k == "function":
                    fns.append(f)
                elif k == "argument":
                    if f.leaf.number == 0:
                        if test is not None:
                            raise NotImplementedError("form that is not linear in the test function")
                        test = f
                    else:
                        if trial is not None:
                            raise NotImplementedError("form that is not linear in the trial function")
                        trial = f
                elif k == "constant":
                    v = f.leaf.value
                    if isinstance(v, tuple):
                        coef *= v[f.comp or 0]
                    elif isinstance(v, LazyScalar):
                        scalars.append(f.leaf)
                    else:
                        coef *= v
                elif k == "lazy":
                    scalars.append(f.leaf)
                elif k == "expression":
                    if f.deriv is not None:
                        raise NotImplementedError("derivative of an Expression inside a form")
                    weights.append((f.leaf, f.comp))
                elif k == "operator":
                    if op is not None:
                        raise NotImplementedError("product of two MatrixOperators in one term")
                    op = f.leaf
                else:
                    raise NotImplementedError("leaf kind '%s' inside a form" % k)
            if coef == 0.0:
                continue
            # the space the integral lives on
            if test is not None:
                space = test.leaf.V
            elif trial is not None:
                space = trial.leaf.V
            elif fns:
                space = fns[0].leaf.V
            else:
                raise NotImplementedError("integral without any function or argument (pure coefficient integral)")
            if meas.domain is not None and meas.domain is not space.mesh():
                raise ValueError("integration domain does not match the mesh of the integrand")
            free = 2 - (test is not None) - (trial is not None)
            operands = []
            for f in fns:
                if len(operands) < free and _same_space(f.leaf.V, space):
                    operands.append(f)
                else:
                    if f.deriv is not None:
                        raise NotImplementedError("derivative of a coefficient Function used as a weight")
                    if f.leaf.V.mesh() is not space.mesh():
                        raise NotImplementedError("coefficient Function living on a different mesh")
                    weights.append((f.leaf, f.comp))
            if test is None and trial is not None:
                raise NotImplementedError("form with a trial but no test function")
            if op is not None:
                slots = ([test] if test is not None else []) + ([trial] if trial is not None else []) + operands
                if len(slots) != 2 or weights or any(f.deriv is not None for f in slots) or meas.kind != "dx":
                    raise NotImplementedError("MatrixOperator terms must be op(u, v)*dx with two plain operands")
                if not _same_space(op.V, space):
                    raise ValueError("MatrixOperator lives on a different space than its operands")
            gkey = (id(space), tuple(sorted([id(s) for s in scalars])) if scalars else (),
                    tuple(sorted([(id(w), c) for w, c in weights])) if weights else (),
                    tuple([id(o.leaf) for o in operands]) if operands else (), mkey,
                    test is not None, trial is not None, id(op) if op is not None else None)
            g = groups.get(gkey)
            if g is None:
                g = Group(tuple(scalars), tuple(sorted(weights, key=lambda wc: (id(wc[0]), wc[1] or 0))),
                          tuple([o.leaf for o in operands]), meas, space, test is not None, trial is not None, op)
                groups[gkey] = g
                order.append(gkey)
            idx = ()
            if test is not None:
                idx += (test.comp or 0, 0 if test.deriv is None else 1 + test.deriv)
            if trial is not None:
                idx += (trial.comp or 0, 0 if trial.deriv is None else 1 + trial.deriv)
            for f in operands:
                idx += (f.comp or 0, 0 if f.deriv is None else 1 + f.deriv)
            ent = g.entries
            ent[idx] = ent.get(idx, 0.0) + coef
    return [groups[k] for k in order]


# ------------------------------------------------------------------------------- device resources
def _weight_specs(weights):
    """[(Expression|Function, comp)] -> list of (sampler, degree, key) for DeviceSpace."""
    specs = []
    for w, comp in weights:
        if isinstance(w, Expression):
            specs.append(("expr", w, comp, w.degree, (id(w), comp, w._version)))
        else:
            specs.append(("fn", w, comp, w.V.degree, (id(w), comp, w._version)))
    return specs


def _measure_key(meas):
    if meas.kind == "dx":
        if meas.subdomain_id is not None and meas.subdomain_data is not None:
            return ("dx", id(meas.subdomain_data), meas.subdomain_data._version, meas.subdomain_id)
        return ("dx",)
    if meas.kind == "ds":
        if meas.subdomain_data is None or meas.subdomain_id is None:
            return ("ds", None, None, None)
        return ("ds", id(meas.subdomain_data), meas.subdomain_data._version, meas.subdomain_id)
    raise NotImplementedError("measure '%s'" % meas.kind)


class Atom:
    """One assembled operator K[T, weights] on a space + the panel of cached products K @ f."""

    def __init__(self, ds, values, symmetric):
        # the DeviceSpace owns its atoms and the Functions own their space: back-references are weak so that the
        # whole structure is released by reference counting when the user drops the problem and its spaces
        self._ds, self.values, self.symmetric = weakref.ref(ds), values, symmetric
        self.panel = None  # [cap, n_dofs]
        self.rows = {}  # id(fn) -> (row, version, weakref(fn))
        self.n_rows = 0

    @property
    def ds(self):
        return self._ds()

    def product(self, fn):
        """Cached K @ fn (a row of the panel); recomputed if fn changed."""
        ent = self.rows.get(id(fn))
        if ent is not None and ent[2]() is not fn:  # the id was recycled by another Function: reuse the row
            ent = (ent[0], -1, None)
        if ent is not None and ent[1] == fn._version:
            return ent[0]
        if ent is None:
            if self.panel is None or self.n_rows == self.panel.shape[0]:
                cap = panel_rows_hint[0] if self.panel is None else 2 * self.panel.shape[0]
                # (zeros: on a partitioned space the ghost tail of a row is never written by the SpMV below)
                new = torch.zeros((cap, self.ds.n_dofs), dtype=torch.float64, device=self.values.device)
                if self.panel is not None:
                    new[: self.n_rows].copy_(self.panel[: self.n_rows])
                self.panel = new
            row = self.n_rows
            self.n_rows += 1
        else:
            row = ent[0]
        ds = self.ds
        _lib.spmv(ds.rowptr_owned, ds.pattern[1], self.values, fn.tensor(), self.panel[row], lpr=ds.lpr)
        self.rows[id(fn)] = (row, fn._version, weakref.ref(fn))
        return row

    def has_fresh(self, fn):
        ent = self.rows.get(id(fn))
        return ent is not None and ent[2]() is fn and ent[1] == fn._version


def _embed_operator(ds, op, scale, transpose):
    """Values of the user matrix (times the scalar form-tensor entry) inside the space's CSR pattern."""
    import scipy.sparse as sp

    if ds.shard is not None:
        raise NotImplementedError("MatrixOperator on an element-partitioned space (user matrices live on the replicated "
                                  "1-D dimensions)")

    rowptr, colidx = ds.pattern[0], ds.pattern[1]
    rp, ci = _lib.to_host(rowptr).astype(np.int64), _lib.to_host(colidx).astype(np.int64)
    n = ds.n_dofs
    A = (op.A.T if transpose else op.A).tocoo()
    key_pat = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp)) * n + ci  # ascending
    key_a = A.row.astype(np.int64) * n + A.col.astype(np.int64)
    pos = np.searchsorted(key_pat, key_a)
    ok = (pos < len(key_pat))
    ok[ok] &= key_pat[pos[ok]] == key_a[ok]
    if np.any(~ok & (A.data != 0.0)):
        raise NotImplementedError("MatrixOperator has nonzeros outside the Lagrange sparsity pattern of its space")
    vals = np.zeros(len(key_pat))
    np.add.at(vals, pos[ok], scale * A.data[ok])
    sym = (abs(op.A - op.A.T)).max() == 0.0 if op.A.nnz else True
    return _lib.to_device(vals), bool(sym)


def get_atom(space, T, weights, meas, op=None, transpose=False):
    ds = device_space(space)
    if transpose and op is None:
        T = np.ascontiguousarray(T.transpose(2, 3, 0, 1))
    tb = _TBYTES.get(id(T)) if not transpose else None
    if tb is None:
        tb = T.tobytes()
    if op is not None:
        key = ("op", id(op), op._version, tb, bool(transpose))
        a = ds.atoms.get(key)
        if a is not None and a.keepalive() is not op:  # id recycled by another operator object
            a = None
        if a is None:
            nz = np.argwhere(T != 0.0)
            if len(nz) != 1 or tuple(nz[0]) != (0, 0, 0, 0):
                raise NotImplementedError("MatrixOperator applied to derivatives / vector components")
            vals, sym = _embed_operator(ds, op, float(T[0, 0, 0, 0]), bool(transpose))
            a = Atom(ds, vals, sym)
            a.keepalive = weakref.ref(op)  # weak: the operator owns its space, the space owns this atom
            ds.atoms[key] = a
        return a
    if meas.kind != "dx":
        raise NotImplementedError("bilinear forms over '%s' (only dx)" % meas.kind)
    mk = _measure_key(meas)
    if mk != ("dx",):
        raise NotImplementedError("bilinear forms restricted to a cell sub-domain")
    specs = _weight_specs(weights)
    key = ("atom", tb, tuple(s[4] for s in specs))
    a = ds.atoms.get(key)
    if a is None:
        vals = ds.assemble_bilinear(T, weights=specs)
        sym = bool(np.array_equal(T, T.transpose(2, 3, 0, 1)))
        a = Atom(ds, vals, sym)
        ds.atoms[key] = a
    return a


def get_load(space, L, weights, meas):
    """Cached load vector  int w sum L[i,j] D_j v_i  over dx or ds(id)."""
    ds = device_space(space)
    specs = _weight_specs(weights)
    mk = _measure_key(meas)
    key = ("load", L.tobytes(), tuple(s[4] for s in specs), mk)
    v = ds.atoms.get(key)
    if v is None:
        if meas.kind == "dx":
            if mk != ("dx",):
                raise NotImplementedError("linear forms restricted to a cell sub-domain")
            v = ds.assemble_linear(L, weights=specs)
        else:
            if np.any(L[:, 1:] != 0):
                raise NotImplementedError("derivatives of the test function in a boundary integral")
            if meas.subdomain_data is None:
                cell, loc = space.mesh().boundary_facets()
            else:
                cell, loc = meas.subdomain_data.facets(meas.subdomain_id)
            v = ds.assemble_facet_linear(mk, cell, loc, L[:, 0], weights=specs)
        ds.atoms[key] = v
    return v


# ------------------------------------------------------------------------------- functionals (rank 0)
class _Functional:
    """Payload of a LazyScalar leaf: one mode integral, evaluated in the next flush."""

    __slots__ = ("kind", "space", "T", "weights", "measure", "f1", "f2", "v1", "v2", "op")

    def __init__(self, kind, space, T, weights, measure, f1=None, f2=None, op=None):
        self.kind, self.space, self.T, self.weights, self.measure, self.f1, self.f2 = kind, space, T, weights, measure, f1, f2
        self.op = op
        self.v1 = f1._version if f1 is not None else None
        self.v2 = f2._version if f2 is not None else None


def _functional_value(g):
    """Group of rank 0 -> LazyScalar (leaf) for  sum_entries  operand integrals."""
    n = len(g.operands)
    T = g.tensor()
    if n == 2:
        p = _Functional("bil", g.space, T, g.weights, g.measure, g.operands[0], g.operands[1], g.op)
    elif n == 1:
        p = _Functional("lin", g.space, T, g.weights, g.measure, g.operands[0])
    else:
        raise NotImplementedError("functional without a Function operand")
    return LazyScalar("leaf", (p,))


POOL_DOUBLES = 1 << 16
_pool = {}  # device -> [tensor POOL_DOUBLES, cursor]: where the functionals of the current step deposit their values


def _pool_take(dev, n):
    """n consecutive slots of the device scalar pool (absolute index of the first one)."""
    ent = _pool.get(dev)
    if ent is None:
        ent = [torch.empty(POOL_DOUBLES, dtype=torch.float64, device=dev), 0]
        _pool[dev] = ent
    if n > POOL_DOUBLES:
        raise ValueError("more than %d functionals in one batch" % POOL_DOUBLES)
    if ent[1] + n > POOL_DOUBLES:
        lazy.fetch()  # everything still living in the pool goes to the host first, then the pool is reused
        ent[1] = 0
    base = ent[1]
    ent[1] += n
    return ent[0], base


def _launch(leaves):
    """Enqueue the evaluation of the given functionals: batched device launches writing into the scalar pool.
    Functionals on an element-partitioned space are local sums over the owned rows; they take the first slots of
    the batch and are completed by ONE all-reduce of that slice ("one batched scalar allreduce")."""
    dev = _device()
    plan = []  # (leaf, slot)
    n_slots = 0
    direct = []  # (slot, kind, args)
    all_jobs = []
    shards = {}
    for leaf in leaves:
        sh = device_space(leaf.args[0].space).shard
        shards.setdefault(sh.space_id if sh is not None else None, (sh, []))[1].append(leaf)  # creation order: same on all ranks
    n_sharded = 0
    reducers = []  # (shard, first slot, end slot)
    for key in sorted(shards, key=lambda k: (k is None, -1 if k is None else k)):  # partitioned spaces first
        sh, group = shards[key]
        first = n_slots
        panel_jobs = {}  # (id(atom), id(x)) -> [atom, x, base_slot, [(leaf, row)]]
        for leaf in group:
            p = leaf.args[0]
            if p.f1 is not None and p.f1._version != p.v1 or p.f2 is not None and p.f2._version != p.v2:
                raise RuntimeError("a Function was modified between assemble(<functional>) and its evaluation")
            if p.kind == "lin":
                vec = get_load(p.space, p.T, p.weights, p.measure)
                direct.append((n_slots, "dot", (vec, p.f1.tensor(), device_space(p.space).n_owned)))
                plan.append((leaf, n_slots))
                n_slots += 1
                continue
            if p.measure.kind != "dx":
                raise NotImplementedError("bilinear functionals over '%s'" % p.measure.kind)
            atom = get_atom(p.space, p.T, p.weights, p.measure, p.op)
            x = stable = None
            at = atom
            if p.f2.stable or atom.has_fresh(p.f2):
                stable, x = p.f2, p.f1
            elif p.f1.stable:
                stable, x = p.f1, p.f2
                at = atom if atom.symmetric else get_atom(p.space, p.T, p.weights, p.measure, p.op, transpose=True)
            if stable is None or stable is x:
                direct.append((n_slots, "bil", (atom, p.f1.tensor(), p.f2.tensor())))
                plan.append((leaf, n_slots))
                n_slots += 1
                continue
            row = at.product(stable)
            job = panel_jobs.get((id(at), id(x)))
            if job is None:
                job = [at, x, None, []]
                panel_jobs[(id(at), id(x))] = job
            job[3].append((leaf, row))
        for job in panel_jobs.values():
            at = job[0]
            job[2] = n_slots
            for leaf, row in job[3]:
                plan.append((leaf, n_slots + row))
            n_slots += at.n_rows
            all_jobs.append(job)
        if sh is not None:
            reducers.append((sh, first, n_slots))
    pool, base0 = _pool_take(dev, max(n_slots, 1))
    res = pool[base0:base0 + max(n_slots, 1)]
    for slot, kind, args in direct:
        if kind == "dot":
            no = args[2]
            _lib.dot(args[0][:no], args[1][:no], out=res[slot:slot + 1])
        else:
            atom = args[0]
            ds = atom.ds
            _lib.bilinear(ds.rowptr_owned, ds.pattern[1], atom.values, args[1], args[2], out=res[slot:slot + 1], lpr=ds.lpr)
    for at, x, base, _ in all_jobs:
        _lib.panel_dots(at.panel, at.n_rows, x.tensor()[: at.ds.n_owned], out=res[base:base + at.n_rows])
    for sh, lo, hi in reducers:
        if hi > lo:
            sh.allreduce(res[lo:hi])
    for leaf, slot in plan:
        leaf._dev = base0 + slot


def _fetch(leaves):
    """ONE device->host copy of the pool range that holds the given (launched) functionals."""
    lo = min(l._dev for l in leaves)
    hi = max(l._dev for l in leaves) + 1
    host = _lib.to_host(_pool[_device()][0][lo:hi])
    for leaf in leaves:
        leaf._value = float(host[leaf._dev - lo])
        leaf._dev = None


lazy._launch_hook[0] = _launch
lazy._fetch_hook[0] = _fetch


# ------------------------------------------------------------------------------- public assemble
_UNIT_T = {}  # (bs, g, comp_a, slot_a, comp_b, slot_b, coef) -> form tensor (shared, read-only)
_TBYTES = {}  # id(T) of the shared tensors above -> T.tobytes() (atom cache key)


def _parse_functional_mono(m):
    """(coef, f1, f2, op, weights) of one monomial with exactly two Function operands, or None."""
    coef = m.coef
    f1 = f2 = op = None
    weights = ()
    for f in m.factors:
        k = f.leaf.kind
        if k == "expression":  # weight, e.g. u*k(x)*v*dx (same bookkeeping as compile_form)
            if f.deriv is not None:
                return None
            weights += ((f.leaf, f.comp),)
        elif k == "function":
            if f1 is None:
                f1 = f
            elif f2 is None:
                f2 = f
            else:
                return None
        elif k == "constant":
            v = f.leaf.value
            if isinstance(v, float):
                coef *= v
            else:
                return None
        elif k == "operator":
            if op is not None:
                return None
            op = f.leaf
        else:
            return None
    if f2 is None:
        return None
    return coef, f1, f2, op, weights


def _fast_functional(meas, monos):
    """Mode integral whose monomials all pair the same two Functions of one space (F*G*dx, F.dx(i)*G.dx(j)*dx,
    inner(grad(F), grad(G))*dx, F*k(x)*G*dx, op(F, G)*dx, times float constants): the overwhelmingly common
    functional of a separated-form callback.  Builds the LazyScalar leaf directly -- same semantics as
    compile_form + Group.tensor, ~5x less host work.  Returns None for anything else (the general path then
    decides or raises)."""
    first = _parse_functional_mono(monos[0])
    if first is None:
        return None
    coef, f1, f2, op, weights = first
    space = f1.leaf.V
    if not _same_space(f2.leaf.V, space):
        return None
    if meas.domain is not None and meas.domain is not space.mesh():
        return None
    if meas.kind != "dx" or meas.subdomain_id is not None:
        return None
    if op is not None and (weights or f1.deriv is not None or f2.deriv is not None or not _same_space(op.V, space)):
        return None
    if len(weights) > 1:
        weights = tuple(sorted(weights, key=lambda wc: (id(wc[0]), wc[1] or 0)))
    key = (space.bs, space.mesh().gdim, f1.comp or 0, _slot(f1.deriv), f2.comp or 0, _slot(f2.deriv), coef)
    for m in monos[1:]:
        nxt = _parse_functional_mono(m)
        if nxt is None:
            return None
        c, g1, g2, op2, w2 = nxt
        if g1.leaf is not f1.leaf or g2.leaf is not f2.leaf or op2 is not None or op is not None:
            return None
        if len(w2) > 1:
            w2 = tuple(sorted(w2, key=lambda wc: (id(wc[0]), wc[1] or 0)))
        if w2 != weights:
            return None
        key += (g1.comp or 0, _slot(g1.deriv), g2.comp or 0, _slot(g2.deriv), c)
    ent = _UNIT_T.get(key)
    if ent is None:
        T = np.zeros((key[0], key[1] + 1, key[0], key[1] + 1))
        for i in range(2, len(key), 5):
            T[key[i], key[i + 1], key[i + 2], key[i + 3]] += key[i + 4]
        if not T.any():
            ent = (False, 1.0)  # the monomials cancel: let the general path produce the zero
        else:
            sg = -1.0 if T.ravel()[np.flatnonzero(T)[0]] < 0.0 else 1.0  # same normalisation as Group.sign()
            if sg < 0.0:
                T = 0.0 - T  # (not -T: the zeros must stay +0.0, the tensor's bytes are the atom's cache key)
            T.setflags(write=False)
            _TBYTES[id(T)] = T.tobytes()
            ent = (T, sg)
        _UNIT_T[key] = ent
    T, sg = ent
    if T is False:
        return None
    memo = functional_memo[0]
    if memo is None:
        leaf = LazyScalar("leaf", (_Functional("bil", space, T, weights, meas, f1.leaf, f2.leaf, op),))
        return leaf if sg > 0.0 else -leaf
    # inside an enrichment step: the same integral of the same (unchanged) functions is asked for again and again
    # (every other dimension's sub-problem needs it) -- hand out the one deferred scalar.  The entry keeps its operands
    # alive, so an id cannot be recycled while it is in the table; the solver clears the table at the end of the step.
    mkey = (id(T), id(f1.leaf), f1.leaf._version, id(f2.leaf), f2.leaf._version, id(op), weights and tuple((id(w), c, getattr(w, "_version", 0)) for w, c in weights))
    hit = memo.get(mkey)
    if hit is None:
        hit = LazyScalar("leaf", (_Functional("bil", space, T, weights, meas, f1.leaf, f2.leaf, op),))
        if sg < 0.0:
            hit = -hit
        memo[mkey] = hit
    return hit


def assemble(form):
    """dolfin.assemble: rank 0 -> LazyScalar, rank 1 -> DeviceVector, rank 2 -> AssembledMatrix."""
    if not isinstance(form, Form):
        raise TypeError("assemble expects a Form, got %r" % type(form))
    ints = form.integrals
    if len(ints) == 1 and 1 <= len(ints[0].monos) <= 64:
        fast = _fast_functional(ints[0].measure, ints[0].monos)
        if fast is not None:
            return fast
    groups = compile_form(form)
    if not groups:
        return lazy.constant(0.0)
    rank = groups[0].rank
    if any(g.rank != rank for g in groups):
        raise ValueError("all integrals of a form must have the same arity")
    if rank == 0:
        total = None
        for g in groups:
            v = _functional_value(g)
            c = g.coefficient()
            term = v if (isinstance(c, float) and c == 1.0) else c * v
            total = term if total is None else total + term
        return total
    if rank == 1:
        return DeviceVector(assemble_vector(groups), device_space(groups[0].space).shard)
    return assemble_matrix(groups)


def _coefs(groups):
    lazy.flush()
    return [float(g.coefficient()) for g in groups]


device_coefficients = [False]  # set by the solver for the duration of an enrichment step (settings["device_coefficients"])
_OPS = {"mul": _lib.OP_MUL, "add": _lib.OP_ADD, "sub": _lib.OP_SUB, "div": _lib.OP_DIV}


def _coefs_dev(groups):
    """The groups' scalar coefficients as a DEVICE tensor, computed by pgd_scalar_programs from the functionals'
    values in the scalar pool -- no device->host round trip between recording a sub-problem's mode integrals and
    assembling its operator / right-hand side.  Returns None (-> host path) for expressions the device evaluator
    does not cover (pow / sqrt / abs, very deep trees) or when nothing is pending on the device anyway."""
    lazy.launch()
    consts, cidx, programs = [], {}, []
    uses_pool = False

    def const(v):
        v = float(v)
        k = cidx.get(v)
        if k is None:
            k = cidx[v] = len(consts)
            consts.append(v)
        return (_lib.OP_CONST << 24) | k

    def emit(node, code, depth):
        """postfix code for node; returns the stack depth needed, or -1 if unsupported"""
        nonlocal uses_pool
        if not isinstance(node, LazyScalar):
            code.append(const(node))
            return depth + 1
        if node._value is not None:
            code.append(const(node._value))
            return depth + 1
        op = node.op
        if op == "leaf":
            if node._dev is None:
                return -1
            uses_pool = True
            code.append((_lib.OP_LOAD << 24) | node._dev)
            return depth + 1
        if op == "neg":
            d = emit(node.args[0], code, depth)
            code.append(_lib.OP_NEG << 24)
            return d
        bop = _OPS.get(op)
        if bop is None:
            return -1
        d0 = emit(node.args[0], code, depth)
        if d0 < 0:
            return -1
        d1 = emit(node.args[1], code, depth + 1)
        if d1 < 0:
            return -1
        code.append(bop << 24)
        return max(d0, d1)

    for g in groups:
        code = []
        need = emit(g.coefficient(), code, 0)
        if need < 0 or need > 16 or len(code) > _lib.SP_MAX_CODE or len(consts) > _lib.SP_MAX_CONST:
            return None
        programs.append(code)
    if not uses_pool:
        return None  # all values are already on the host: nothing to gain
    dev = _device()
    out = torch.empty(len(programs), dtype=torch.float64, device=dev)
    return _lib.scalar_programs(programs, consts, _pool[dev][0], out)


def _coefficients(groups):
    if device_coefficients[0]:
        c = _coefs_dev(groups)
        if c is not None:
            return c
    return _coefs(groups)


def assemble_vector(groups, out=None):
    """b = sum_g coef_g * (load_g | K_g @ f_g)  on the device."""
    space = groups[0].space
    vecs = []
    for g in groups:
        if not _same_space(g.space, space):
            raise ValueError("linear form mixes test spaces")
        if len(g.operands) == 0:
            vecs.append(get_load(g.space, g.tensor(), g.weights, g.measure))
        elif len(g.operands) == 1:
            atom = get_atom(g.space, g.tensor(), g.weights, g.measure, g.op)
            f = g.operands[0]
            if f.stable or atom.has_fresh(f):
                row = atom.product(f)
                vecs.append(atom.panel[row])
            else:
                ds = atom.ds
                y = torch.zeros(ds.n_dofs, dtype=torch.float64, device=atom.values.device) if ds.shard is not None else None
                vecs.append(_lib.spmv(ds.rowptr_owned, ds.pattern[1], atom.values, f.tensor(), y=y, lpr=ds.lpr))
        else:
            raise NotImplementedError("linear form with two Function operands in one term")
    coefs = _coefficients(groups)
    return _lib.lincomb(vecs, coefs, out=out)


class AssembledMatrix:
    def __init__(self, ds, values):
        self.ds, self.values = ds, values

    def scipy(self):
        import scipy.sparse as sp

        rowptr, colidx = self.ds.pattern[0], self.ds.pattern[1]
        n = self.ds.n_dofs
        return sp.csr_matrix((_lib.to_host(self.values), _lib.to_host(colidx), _lib.to_host(rowptr)), shape=(n, n))

    def array(self):
        return self.scipy().toarray()


def assemble_matrix(groups, out=None):
    space = groups[0].space
    ds = device_space(space)
    atoms = []
    for g in groups:
        if not _same_space(g.space, space) or g.operands:
            raise NotImplementedError("bilinear form with Function operands (non-linear form?)")
        atoms.append(get_atom(g.space, g.tensor(), g.weights, g.measure, g.op))
    coefs = _coefficients(groups)
    vals = _lib.lincomb([a.values for a in atoms], coefs, out=out)
    m = AssembledMatrix(ds, vals)
    m.symmetric = all(a.symmetric for a in atoms)
    return m


def norm(f, norm_type="L2", mesh=None):
    """dolfin.norm(Function): sqrt(int f.f dx)  (fused SpMV-reduction with the mass atom)."""
    if isinstance(f, _DofOwner) and not isinstance(f, Function):
        from .functions import _owned_dot

        return math.sqrt(_owned_dot(f, f))
    if norm_type.lower() != "l2":
        raise NotImplementedError("norm type '%s'" % norm_type)
    return mass_product(f, f).sqrt()


def mass_product(f, g):
    """LazyScalar  int f.g dx  for two Functions of the same space."""
    from .ufl import dx, inner

    return assemble(inner(f, g) * dx(f.V.mesh()))
