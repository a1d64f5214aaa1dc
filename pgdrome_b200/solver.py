"""PGDProblem: the progressive PGD enrichment of pgdrome/solver.py on the B200.

Same constructor, callbacks, public attributes and methods as the reference class
(pgdrome/solver.py:12-943); underneath, every operation on dof data is a libpgdb200 kernel:

  get_Fsinit     solver.py:158-304   ones -> Dirichlet -> (rand) -> / ||.||_M     (fill, set_entries, bilinear)
  residual check solver.py:347-395   ||BC(b_d(F_init))||                           (lincomb, set_entries, dot)
  FP_solve       solver.py:508-881   per dim: A = sum c_k K_k, b = sum c_m g_m - sum c_ik K_k U_i,
                                     Dirichlet, solve, norm; "norm" / "delta" stopping criterion
                                     (lincomb, apply_dirichlet, Jacobi-PCG | banded LU, bilinear, panel_dots)
  normalisation  solver.py:406-470   "no" | "stiff" | "l2"
  stopping       solver.py:476-504   prod ||F_d|| / first < PGD_tol

Deviations, all forced by "no CPU, no direct sparse solver on the device" (see DESIGN.md):
  * FEM systems on 2-D/3-D spaces are solved by (block-)Jacobi PCG to ``relative_tolerance``
    (default 1e-13) instead of MUMPS LU; 1-D systems (FEM or FD, symmetric or not) by a banded LU
    with partial pivoting, i.e. the same direct factorisation class as the reference.
  * with ``_problem="linear"`` the callbacks are invoked once per sub-problem (with a
    TrialFunction), not twice; ``bc_fct`` / ``dom_fct`` are evaluated once per ``solve_PGD``.
"""
import gc
import logging

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, forms, lazy, ufl
from .assembly import device_space
from .functions import Function, MatrixOperator, bc_list, local_bc, merged_bc_dofs
from .lazy import LazyScalar
from .ufl import Form, TestFunction, TrialFunction, derivative
from .ufl import dx as forms_dx

F64, I32 = torch.float64, torch.int32


def _f(x):
    return float(x)


class _CsrOnDevice:
    """A user-supplied SciPy matrix (FD operators, MM) mirrored as CSR on the device."""

    def __init__(self, A):
        A = sp.csr_matrix(A)
        A.sort_indices()
        dev = forms._device()
        self.n = A.shape[0]
        self.rowptr = _lib.to_device(A.indptr.astype(np.int32))
        self.colidx = _lib.to_device(A.indices.astype(np.int32))
        self.values = _lib.to_device(A.data.astype(np.float64))
        coo = A.tocoo()
        nz = coo.data != 0
        d = (coo.row[nz] - coo.col[nz]) if nz.any() else np.zeros(1, dtype=np.int64)
        self.kl, self.ku = int(max(0, d.max())), int(max(0, -d.min()))
        m = A.nnz / max(1, self.n)
        self.lpr = 2 if m <= 3 else (4 if m <= 6 else 8)

    def product(self, x, y):
        """device scalar tensor x^T A y"""
        return _lib.bilinear(self.rowptr, self.colidx, self.values, x, y, lpr=self.lpr)


MODE_ARENA_ROWS = 64  # rows per block of the mode arenas / first allocation of the product panels


class _ParkedHeap:
    """While an enrichment step runs, every object that existed before it sits in the collector's permanent
    generation (``gc.freeze``), and goes back afterwards (``gc.unfreeze``; both are O(1) list splices).

    A step creates ~10^5 short-lived Python objects (form nodes, deferred scalars).  They die by reference count,
    but their allocation still drives CPython's generational collector, and a full collection walks every tracked
    object of the process -- ~2*10^5 for torch + scipy + this package, measured 0.13 s, i.e. several enrichment
    steps of configs[1] -- whenever the long-lived heuristic fires.  With the old objects parked, collections
    inside the step only see what the step itself allocated.  ``settings={"gc_freeze": False}`` turns it off; a
    heap the caller froze is left as it is."""
    depth = 0
    base = gc.get_freeze_count()  # what the interpreter itself keeps frozen (a few hundred start-up objects here)

    def __init__(self, on):
        self.on = bool(on)
        self.mine = False

    def __enter__(self):
        if self.on:
            if _ParkedHeap.depth == 0 and gc.get_freeze_count() == _ParkedHeap.base:
                gc.freeze()
                self.mine = True
            _ParkedHeap.depth += 1
        return self

    def __exit__(self, *exc):
        if self.on:
            _ParkedHeap.depth -= 1
            if self.mine:
                gc.unfreeze()
                _ParkedHeap.base = gc.get_freeze_count()
        return False


class PGDProblem:
    def __init__(self, name=None, name_coord=[], modes_info=[], Vs=[], dom_fct=None, bc_fct=None, load=[], param=None,
                 rhs_fct=None, lhs_fct=None, probs=[], seq_fp=[], PGD_nmax=20, PGD_tol=1e-10, num_elem=[], order=[],
                 ranges=[], dims=[], *args, **kwargs):
        self.logger = logging.getLogger(__name__ + "." + self.__class__.__name__)
        if "Vs_out" in kwargs and not len(Vs):
            Vs = kwargs.pop("Vs_out")
        self.name = name
        self.name_coord = name_coord
        self.modes_info = modes_info
        self.num_pgd_var = len(self.name_coord)
        self.V = list(Vs) if len(Vs) else [0] * self.num_pgd_var
        self.meshes = [v.mesh() if v != 0 else 0 for v in self.V]
        self.dom_fct, self.bc_fct = dom_fct, bc_fct
        self.load, self.param = load, param
        self.rhs_fct, self.lhs_fct = rhs_fct, lhs_fct
        self.prob = probs
        self.seq_fp = list(range(self.num_pgd_var)) if len(seq_fp) == 0 else [int(s) for s in seq_fp]
        self.PGD_nmax, self.PGD_tol = PGD_nmax, PGD_tol
        self.num_elem, self.order, self.ranges, self.dims = num_elem, order, ranges, dims

        self.PGD_func = []
        self.alpha = []
        self.amplitude = []
        self.num_fp_it = []
        self.err_fp_it = []
        self.PGD_modes = None

        self.max_fp_it = 50
        self.tol_fp_it = 1e-5
        self.tol_abs = 1e-6
        self.stop_fp = "norm"
        self.fp_init = ""
        self.norm_modes = "stiff"
        # optional: number of fixed-point sweeps per enrichment step (overrides the stopping test of that step; steps
        # beyond the list use the test).  Lets two runs be compared at identical sweep counts when the "norm" test
        # sits on its round-off floor (DESIGN.md "Parity status").
        self.fp_schedule = None
        self.simulation_info = (
            "PGD solver option: PGD_nmax %s / PGD tolerance %s and max FP iterations %s and FP tolerance %s; \n"
            % (self.PGD_nmax, self.PGD_tol, self.max_fp_it, self.tol_fp_it))
        self.solve_mode = {"FEM": "FEM", "direct": "direct", "FD": "FD"}
        self.MM = []

        # device-side bookkeeping
        self._bc_cache = None
        self._dom_cache = None
        self._mm_dev = {}
        self._csr_cache = {}
        self.solver_stats = {"pcg_solves": 0, "pcg_iterations": 0, "banded_solves": 0, "flushes": 0, "pcg_log": []}

    # ------------------------------------------------------------------ domains / boundary conditions
    @property
    def dom(self):
        """dom_fct(Vs, param) (solver.py:136-145), evaluated once and cached."""
        if self._dom_cache is None:
            self._dom_cache = (self.dom_fct(self.V, self.param) if self.dom_fct else 0,)
        return self._dom_cache[0]

    @property
    def bc(self):
        """bc_fct(Vs, dom, param) (solver.py:147-156), evaluated once and cached: the Dirichlet dof
        sets live on the device afterwards."""
        if self._bc_cache is None:
            self._bc_cache = (self.bc_fct(self.V, self.dom, self.param) if self.bc_fct else [0] * self.num_pgd_var,)
        return self._bc_cache[0]

    def invalidate_caches(self):
        self._bc_cache = self._dom_cache = None
        self._bcd = {}
        self._mode_arena = {}
        forms.panel_rows_hint[0] = max(forms.MAX_PANEL_ROWS, min(int(self.PGD_nmax), MODE_ARENA_ROWS))

    def _bc_dev(self, dim):
        """(dofs int32 tensor, values tensor | None) of dimension dim, or None."""
        if not hasattr(self, "_bcd"):
            self._bcd = {}
        if dim not in self._bcd:
            bcs = bc_list(self.bc[dim])
            if not bcs:
                self._bcd[dim] = None
            else:
                dofs, vals = merged_bc_dofs(bcs, self.V[dim])
                nonzero = bool(np.any(vals))  # decided on the GLOBAL set: every rank takes the same code path
                dofs, vals = local_bc(self.V[dim], dofs, vals)  # partitioned space: the dofs this rank holds, local numbers
                self._bcd[dim] = (_lib.to_device(dofs), _lib.to_device(vals) if nonzero else None)
        return self._bcd[dim]

    def _mm(self, dim):
        if dim not in self._mm_dev or self._mm_dev[dim][0] is not self.MM[dim]:
            self._mm_dev[dim] = (self.MM[dim], _CsrOnDevice(self.MM[dim]))
        return self._mm_dev[dim][1]

    def _is_fd(self, solve_modes, dim):
        return solve_modes is not None and solve_modes[dim] == self.solve_mode["FD"]

    def _mm_operator(self, dim):
        """MM[dim] given as a device MatrixOperator (extension: keeps FD-type norms on the device)."""
        if len(self.MM) > dim and isinstance(self.MM[dim], MatrixOperator):
            return self.MM[dim]
        return None

    def _mass_product(self, f, g, dim, solve_modes):
        """<f, g> in the norm of dimension dim: MM[dim] for FD dims (solver.py:198-207, 814-835),
        the consistent L2 mass otherwise.  LazyScalar or float."""
        op = self._mm_operator(dim)
        if op is not None:
            return forms.assemble(op(g, f) * forms_dx(self.meshes[dim]))
        if self._is_fd(solve_modes, dim):
            return float(self._mm(dim).product(f.tensor(), g.tensor()).item())
        return forms.mass_product(f, g)

    def _norm(self, f, dim, solve_modes):
        """LazyScalar / float  ||f||: dolfin.norm, or sqrt(f^T MM f) for FD dims."""
        r = self._mass_product(f, f, dim, solve_modes)
        return r.sqrt() if isinstance(r, LazyScalar) else float(r) ** 0.5

    # ------------------------------------------------------------------ initial modes
    def get_Fsinit(self, V, bc=None, solve_modes=None):
        """ones -> Dirichlet values -> optional np.random.rand -> normalise (solver.py:158-304)."""
        if not bc:
            bc = [0] * len(V)
        Fs_init = [None] * len(V)
        dev = forms._device()
        pending = []
        for dim in range(len(V)):
            tdim = V[dim].mesh().topology().dim()
            if V[dim].bs > 1 and tdim not in (1, 2, 3):
                raise ValueError("ERROR DIMENSION NOT defined!!!!!!!!!!!")
            f = Function(V[dim])
            f.tensor().fill_(1.0)
            for b in bc_list(bc[dim]):
                b.apply(f.vector())
            if self.fp_init.lower() == "randomized":
                a = f.vector()[:]
                idx = np.where(a != 0)[0]
                a[idx] = np.random.rand(len(idx))
                f.vector()[:] = a
            Fs_init[dim] = f
            pending.append(self._norm(f, dim, solve_modes))
        for dim, n in enumerate(pending):
            Fs_init[dim].vector()[:] *= 1.0 / _f(n)
        return Fs_init

    # ------------------------------------------------------------------ enrichment loop
    def solve_PGD(self, _problem="nonlinear", solve_modes=None, settings={"linear_solver": "mumps"}):
        """Progressive enrichment (solver.py:306-506). Returns self."""
        _lib.require_cuda()
        self.invalidate_caches()
        D = self.num_pgd_var
        n_enr = -1
        normConv, relConv = [], []
        while n_enr < self.PGD_nmax - 1:
            n_enr += 1
            if n_enr == 0:
                self.PGD_func = [[] for _ in range(D)]
                normConv, relConv = [], []
            self.logger.info("enrichment step %s ", n_enr)
            done = self.enrichment_step(n_enr, normConv, relConv, _problem, solve_modes, settings)
            if done:
                break
        self.amplitude = relConv
        self.PGD_modes = len(self.PGD_func[0])
        return self

    def begin_PGD(self, _problem="nonlinear", solve_modes=None, settings={"linear_solver": "mumps"}):
        """Stepping form of solve_PGD (bench.py times single enrichment steps): returns the loop state."""
        _lib.require_cuda()
        self.invalidate_caches()
        self.PGD_func = [[] for _ in range(self.num_pgd_var)]
        return {"n_enr": -1, "normConv": [], "relConv": [], "args": (_problem, solve_modes, settings), "done": False}

    def step_PGD(self, state):
        """One enrichment step; returns True when the enrichment has stopped."""
        state["n_enr"] += 1
        state["done"] = self.enrichment_step(state["n_enr"], state["normConv"], state["relConv"], *state["args"])
        self.amplitude = state["relConv"]
        self.PGD_modes = len(self.PGD_func[0])
        return state["done"]

    def enrichment_step(self, n_enr, normConv, relConv, _problem="nonlinear", solve_modes=None,
                        settings={"linear_solver": "mumps"}):
        """One pass of the enrichment loop body (one new mode per dimension). True = stop."""
        with _ParkedHeap(settings.get("gc_freeze", True)):
            outer = forms.functional_memo[0]
            outer_dc = forms.device_coefficients[0]
            if outer is None and settings.get("memo_functionals", True):
                forms.functional_memo[0] = {}
            forms.device_coefficients[0] = bool(settings.get("device_coefficients", True))
            outer_cc = ufl.capture_cache[0]
            if outer_cc is None and settings.get("memo_functionals", True):
                ufl.capture_cache[0] = {}
            try:
                return self._enrichment_step(n_enr, normConv, relConv, _problem, solve_modes, settings)
            finally:
                forms.functional_memo[0] = outer
                forms.device_coefficients[0] = outer_dc
                ufl.capture_cache[0] = outer_cc
                self._collect_solve()
                lazy.fetch()  # functionals that were only consumed on the device: bring their values home, release them

    def _enrichment_step(self, n_enr, normConv, relConv, _problem, solve_modes, settings):
        D = self.num_pgd_var
        Fs_init = self.get_Fsinit(self.V, self.bc, solve_modes)
        norm_Fs = np.ones(D)
        nl = [forms.norm(Fs_init[i]) for i in range(D)]
        for i in range(D):
            norm_Fs[i] = _f(nl[i])
        delta = np.ones(D)

        # residual of the initial guess (solver.py:347-395)
        res_dev = torch.zeros(D, dtype=F64, device=forms._device())
        for dim in range(D):
            sh = None
            if solve_modes is None or solve_modes[dim] == self.solve_mode["FEM"]:
                var_F = TestFunction(self.V[dim])
                l = self.rhs_fct(Fs_init, var_F, Fs_init, self.meshes, self.dom, self.param, self.load, self.PGD_func,
                                 self.prob[dim], n_enr, dim)
                ll = forms.assemble_vector(forms.compile_form(l))
                bcd = self._bc_dev(dim)
                if bcd is not None:
                    _lib.set_entries(ll, bcd[0], bcd[1])
                ds = device_space(self.V[dim])
                sh = ds.shard
                if sh is not None:
                    ll = ll[: ds.n_owned]  # partitioned space: this rank's rows; summed over the ranks below
            else:
                v = self.rhs_fct(Fs_init, Fs_init, Fs_init, self.meshes, self.dom, self.param, self.load, self.PGD_func,
                                 self.prob[dim], n_enr, dim)
                ll = _lib.to_device(np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel())
            _lib.dot(ll, ll, out=res_dev[dim:dim + 1])
            if sh is not None:
                sh.allreduce(res_dev[dim:dim + 1])
        res = _lib.to_host(res_dev)
        res_error = np.sqrt(np.sum(res))
        self.simulation_info += f"-- residuum norm: {res_error} --\n"
        if res_error < 1e-10:
            self.logger.info("Residuum error %s smaller 1e-10 in enrichment step number %s\n STOPP" % (res_error, n_enr))
            self.simulation_info += f"<<<before enrichment step {n_enr} residuum norm smaller 1e-10: {res_error} STOP >>>\n"
            return True

        Fs, norm_Fs = self.FP_solve(Fs_init, norm_Fs, delta, n_enr, _problem, solve_modes, settings)

        # normalisation and storage of the new modes (solver.py:406-470)
        normU = np.prod(norm_Fs)
        mode = self.norm_modes.lower()
        if mode == "no":
            for dim in range(D):
                self._store(dim, Fs[dim])
            self.alpha.append(1.0)
        elif mode == "stiff":
            Fn = Fs
            for dim in range(D):
                Fn[dim].vector()[:] *= 1 / norm_Fs[dim]
            a = self.lhs_fct(Fn[-1], Fn[-1], Fn, self.meshes, self.dom, self.param, self.prob[-1], D)
            if solve_modes is not None and solve_modes[-1] == self.solve_mode["FD"]:
                t = Fn[-1].tensor()
                norm_aux = float(self._csr(a).product(t, t).item())
            elif solve_modes is not None and solve_modes[-1] == self.solve_mode["direct"]:
                norm_aux = _f(a)
            else:
                norm_aux = _f(forms.assemble(a))
            norm_fac = np.sqrt(np.absolute(norm_aux)) ** (1.0 / D)
            self.alpha.append(np.prod(norm_Fs) * norm_fac**D)
            for dim in range(D):
                Fn[dim].vector()[:] *= 1.0 / norm_fac
                Fn[dim].vector()[:] *= self.alpha[-1] ** (1.0 / D)
                self._store(dim, Fn[dim])
        elif mode == "l2":
            self.alpha.append(normU)
            norm_all = np.prod(norm_Fs) ** (1.0 / D)
            for dim in range(D):
                tmp = Function(self.V[dim])
                tmp.vector().axpy(norm_all / norm_Fs[dim], Fs[dim].vector())
                self._store(dim, tmp)
        else:
            raise ValueError('norm_modes must be "no", "stiff" or "l2"')

        normConv.append(normU)
        relConv.append(normU / normConv[0])
        self.logger.info("PGD modes updated: normU=%s; relNorm=%s; tol=%s; res_error=%s", normU, relConv[n_enr],
                         self.PGD_tol, res_error)
        if relConv[n_enr] < self.PGD_tol:
            self.logger.info("Convergence reached (normU = %s relative %s [res_error %s]), enriched basis number %s"
                             % (normU, relConv[n_enr], res_error, n_enr))
            return True
        return False

    def _store(self, dim, f):
        """Keep a converged mode.  Its values move into the dimension's mode arena -- [rows, n_dofs] blocks sized
        from PGD_nmax and allocated when the first mode arrives -- so that a run does not grow the device heap mode
        by mode (every new allocator segment is a blocking cudaMalloc), and all modes of a dimension are rows of one
        matrix (the X operand of evaluate)."""
        f.stable = True  # K @ U_i products of stored modes are cached in the atom panels
        arena = self.__dict__.setdefault("_mode_arena", {}).setdefault(dim, [None, 0])  # [current block, rows used]
        t = f.tensor()
        if arena[0] is None or arena[1] == arena[0].shape[0]:
            rows = max(1, min(int(self.PGD_nmax) - len(self.PGD_func[dim]), MODE_ARENA_ROWS))
            arena[0], arena[1] = torch.empty((rows, t.numel()), dtype=t.dtype, device=t.device), 0
        row = arena[0][arena[1]]
        arena[1] += 1
        row.copy_(t)
        f.set_tensor(row)
        self.PGD_func[dim].append(f)

    def _csr(self, A):
        """Device mirror of a host matrix returned by a callback (rebuilt when the object changes)."""
        return _CsrOnDevice(A)

    # ------------------------------------------------------------------ fixed point
    def FP_solve(self, Fs_init, norm_Fs, delta, n_enr, _problem, solve_modes, settings):
        """Alternating fixed point over the dimensions in seq_fp (solver.py:508-881)."""
        D = self.num_pgd_var
        Fs = list(Fs_init)
        Fs_init = list(Fs_init)
        norms = [None] * D
        stop = self.stop_fp.lower()
        if stop not in ("norm", "delta"):
            raise ValueError('stopping criterion not defined %s (self.stop_fp = "delta" or "norm")')
        forced = self.fp_schedule[n_enr] if (self.fp_schedule is not None and n_enr < len(self.fp_schedule)) else None
        for fpi in range(self.max_fp_it):
            for dim in self.seq_fp:
                # warm start of the iterative spatial solves: the previous sweep's mode of this dimension
                self._x0 = Fs[dim].tensor() if (fpi > 0 and (settings or {}).get("warm_start", True)) else None
                fct_F = self._solve_dimension(dim, Fs, n_enr, _problem, solve_modes, settings)
                self._x0 = None
                Fs[dim] = fct_F
                norms[dim] = self._norm(fct_F, dim, solve_modes)
            if stop == "delta":
                for dim in range(D):
                    if Fs[dim]._shard is not None:  # partitioned space: the (rarely used) max-norm test on gathered copies
                        a_new, a_old = Fs[dim].values_host(), Fs_init[dim].values_host()
                        dt = np.abs(a_new - a_old)
                        mi = int(np.argmax(dt))
                        delta[dim] = dt[mi] if abs(a_new[mi]) < 1e-8 else dt[mi] / abs(a_new[mi])
                        continue
                    d = (Fs[dim].tensor() - Fs_init[dim].tensor()).abs()
                    mx, mi = torch.max(d, dim=0)
                    at = abs(float(Fs[dim].tensor()[mi].item()))
                    mx = float(mx.item())
                    delta[dim] = mx if at < 1e-8 else mx / at
                notconv = len(np.where(delta > self.tol_fp_it)[0]) > 0
                if forced is not None:
                    notconv = fpi + 1 < forced
                if notconv and fpi < self.max_fp_it - 1:
                    Fs_init = list(Fs)
                    continue
                if notconv:
                    self.logger.error("ERROR: fix point iteration in maximum number of iterations NOT converged (enrichment loop %s)", n_enr)
                    self.simulation_info += f"<<<enrichment step {n_enr} fixed point iteration NOT converged in {fpi + 1} / delta: {delta} >>>\n"
                else:
                    self.simulation_info += f"enrichment step {n_enr} fixed point iteration converged in {fpi + 1} / delta: {delta} \n"
                self.num_fp_it.append(fpi + 1)
                self.err_fp_it.append(delta)
                break
            # "norm": err = sqrt|prod<new,new> + prod<old,old> - 2 prod<new,old>|  (solver.py:812-844)
            newnew, newold, oldold = 1, 1, 1
            terms = []
            for d in range(D):
                if self._is_fd(solve_modes, d) and self._mm_operator(d) is None:
                    M = self._mm(d)
                    tn, to = Fs[d].tensor(), Fs_init[d].tensor()
                    trip = _lib.to_host(torch.cat([M.product(tn, tn), M.product(tn, to), M.product(to, to)]))
                    terms.append((float(trip[0]), float(trip[1]), float(trip[2])))
                else:
                    terms.append((self._norm(Fs[d], d, solve_modes) ** 2, self._mass_product(Fs[d], Fs_init[d], d, solve_modes),
                                  self._norm(Fs_init[d], d, solve_modes) ** 2))
            for nn, no, oo in terms:
                newnew *= _f(nn)
                newold *= _f(no)
                oldold *= _f(oo)
            max_error = np.sqrt(np.absolute(newnew + oldold - 2 * newold))
            if (max_error < self.tol_fp_it) if forced is None else (fpi + 1 >= forced):
                self.logger.info(f"fix point iteration converged !!! in number of steps: {fpi + 1} (error {max_error:8.6e})")
                self.simulation_info += f"enrichment step {n_enr} fixed point iteration converged in {fpi + 1} / error: {max_error:8.6e} \n"
                self.num_fp_it.append(fpi + 1)
                self.err_fp_it.append(max_error)
                break
            elif fpi < self.max_fp_it - 1:
                Fs_init = list(Fs)
            else:
                self.logger.error(f"ERROR: fix point iteration in maximum number of iterations NOT converged (enrichment loop {n_enr}) (error {max_error:8.6e})")
                self.simulation_info += f"<<<enrichment step {n_enr} fixed point iteration NOT converged in {fpi + 1} / error: {max_error:8.6e} >>>\n"
                self.num_fp_it.append(fpi + 1)
                self.err_fp_it.append(max_error)
                break
        for dim in range(D):
            if norms[dim] is not None:
                norm_Fs[dim] = _f(norms[dim])
        return Fs, norm_Fs

    def _solve_dimension(self, dim, Fs, n_enr, _problem, solve_modes, settings):
        mode = "FEM" if solve_modes is None else solve_modes[dim]
        V = self.V[dim]
        if mode == self.solve_mode["FEM"]:
            var_F = TestFunction(V)
            if _problem.lower() == "linear":
                fct_F = TrialFunction(V)
                a = self.lhs_fct(fct_F, var_F, Fs, self.meshes, self.dom, self.param, self.prob[dim], dim)
                l = self.rhs_fct(fct_F, var_F, Fs, self.meshes, self.dom, self.param, self.load, self.PGD_func,
                                 self.prob[dim], n_enr, dim)
            elif _problem.lower() == "nonlinear":
                u = Function(V)
                a = self.lhs_fct(u, var_F, Fs, self.meshes, self.dom, self.param, self.prob[dim], dim)
                l = self.rhs_fct(u, var_F, Fs, self.meshes, self.dom, self.param, self.load, self.PGD_func,
                                 self.prob[dim], n_enr, dim)
                a, l = newton_linearise(a - l, u)
            else:
                raise ValueError("_problem must be 'linear' or 'nonlinear'")
            return self.variational_solve(a, l, dim, settings)
        u = Function(V)
        var_F = TestFunction(V)
        a = self.lhs_fct(u, var_F, Fs, self.meshes, self.dom, self.param, self.prob[dim], dim)
        l = self.rhs_fct(u, var_F, Fs, self.meshes, self.dom, self.param, self.load, self.PGD_func, self.prob[dim],
                         n_enr, dim)
        if mode == self.solve_mode["direct"]:
            return self.direct_solve(a, l, dim)
        if mode == self.solve_mode["FD"]:
            return self.FD_solve(a, l, dim)
        self.logger.error("ERROR: solver %s doesn't exist", mode)
        raise ValueError("solver %s doesn't exist" % mode)

    # ------------------------------------------------------------------ linear solves
    def variational_solve(self, a, l, dim, settings=None):
        """LinearVariationalSolver.solve() of a == l with the dimension's Dirichlet BCs
        (solver.py:598-636, 677-716): assemble into the fixed CSR pattern, symmetric elimination, solve."""
        V = self.V[dim]
        ds = device_space(V)
        settings = settings or {}
        ga, gl = forms.compile_form(a), forms.compile_form(l)
        if any(g.rank != 2 for g in ga) or not ga:
            raise ValueError("left-hand side is not a bilinear form")
        A = forms.assemble_matrix(ga)
        if gl:
            if any(g.rank != 1 for g in gl):
                raise ValueError("right-hand side is not a linear form")
            b = forms.assemble_vector(gl)
        else:
            b = torch.zeros(ds.n_dofs, dtype=F64, device=A.values.device)
        rowptr, colidx = ds.pattern[0], ds.pattern[1]
        bcd = self._bc_dev(dim)
        if bcd is not None:
            _lib.apply_dirichlet(rowptr, colidx, A.values, b, bcd[0], bcd[1])
        x = self._linear_solve(ds, A.values, b, A.symmetric, settings)
        return Function(V, x)

    def _linear_solve(self, ds, values, b, symmetric, settings):
        V = ds.space
        rowptr, colidx = ds.pattern[0], ds.pattern[1]
        if V.mesh().tdim == 1:
            perm, bw = V.band_permutation()
            if ds.band is None:
                ds.band = _lib.to_device(perm)
            x, info = _lib.banded_solve(rowptr, colidx, values, b, ds.band, bw, bw)
            self.solver_stats["banded_solves"] += 1
            self._last_info = info
            return x
        if not symmetric:
            raise NotImplementedError(
                "non-symmetric operator on a %d-D space: only (block-)Jacobi PCG is available for 2-D/3-D "
                "spaces (north_star); put first-derivative terms on 1-D dimensions" % V.mesh().tdim)
        rtol = float(settings.get("relative_tolerance", 1e-13))
        atol = float(settings.get("absolute_tolerance", 0.0))
        maxit = int(settings.get("maximum_iterations", max(10000, 4 * V.n_dofs // max(1, V.bs))))
        prec = str(settings.get("preconditioner", "default")).lower()
        block = V.bs if (V.bs <= 3 and prec not in ("jacobi", "none_block")) else 1
        x0 = getattr(self, "_x0", None)
        if x0 is not None and x0.numel() != b.numel():
            x0 = None
        self._collect_solve()
        if ds.shard is not None:
            return self._sharded_solve(ds, values, b, block, rtol, atol, maxit, settings, x0)
        if settings.get("async_solve", True):
            # SM-resident systems (known to fit from an earlier solve): enqueue and go on recording the next
            # sub-problem's forms while the GPU iterates; iterations / residual are collected before the next solve
            x = _lib.pcg_start(rowptr, colidx, values, b, rtol=rtol, atol=atol, maxit=maxit, block=block, x0=x0)
            if x is not None:
                self._pending_solve = (rtol, maxit)
                return x
        if block > 1 and V.n_dofs >= 32768 and settings.get("node_block_walk", True) and ds.bsr is not None:
            # vector operator in the HBM-bound regime: persistent kernel walking the CSR arrays by node blocks.  It is
            # only enqueued (async_solve): the sweep solves the spatial dimension first, so the forms of the parameter
            # dimensions are recorded while the GPU iterates; iterations / residual are collected before the next solve
            defer = bool(settings.get("async_solve", True))
            x, iters, relres = _lib.pcg_persist(rowptr, colidx, values, b, block=block, rtol=rtol, atol=atol, maxit=maxit,
                                                x0=x0, bsr=ds.bsr, defer=defer)
            if defer:
                self._pending_solve = (rtol, maxit)
            else:
                self._account_solve(iters, relres, rtol, maxit)
            return x
        x, iters, relres = _lib.pcg(rowptr, colidx, values, b, rtol=rtol, atol=atol, maxit=maxit,
                                    check_every=int(settings.get("check_every", 50)), block=block, lpr=ds.lpr, x0=x0)
        self._account_solve(iters, relres, rtol, maxit)
        return x

    def _account_solve(self, iters, relres, rtol, maxit):
        self.solver_stats["pcg_solves"] += 1
        self.solver_stats["pcg_iterations"] += iters
        self.solver_stats["pcg_log"].append(int(iters))
        if relres > max(rtol, 1e-15) * 10 and iters >= maxit:
            self.logger.warning("PCG stopped at relative residual %.3e after %d iterations", relres, iters)

    def _collect_solve(self):
        """Wait for a solve started with pgd_pcg_start and book its iteration count."""
        pend = self.__dict__.pop("_pending_solve", None)
        if pend is not None:
            iters, relres = _lib.pcg_finish()
            if iters >= 0:
                self._account_solve(iters, relres, *pend)

    # ---- multi-GPU: the spatial dimension is element-partitioned (sharding.py); this rank holds its rows only
    def _sharded_solve(self, ds, values, b, block, rtol, atol, maxit, settings, x0=None):
        """Sharded PCG on the owned rows of an element-partitioned space: halo exchange of the direction vector and
        dot-product all-reduces per iteration (persistent kernel over the NVLink peer window, NCCL otherwise); returns
        the local vector [owned | ghost] with up-to-date ghosts.  Nothing is replicated or gathered here."""
        from . import partition as pt

        sh = ds.shard
        no = ds.n_owned
        S = getattr(ds, "_shard_matrix", None)
        if S is None:
            halo = sh.halo(b.device)
            S = pt.ShardedMatrix(ds.rowptr_owned, ds.pattern[1][: ds.nnz_owned], None, halo, int(sh.part.bounds[sh.rank]),
                                 int(sh.part.bounds[sh.rank + 1]), block)
            ds._shard_matrix = S
        S.values = values[: ds.nnz_owned]
        x, iters, relres = pt.sharded_solve(S, b, x0=x0, rtol=rtol, atol=atol, maxit=maxit,
                                            check_every=int(settings.get("check_every", 50)), block=block,
                                            bsr=ds.bsr if settings.get("node_block_walk", True) else None)
        self.solver_stats["pcg_solves"] += 1
        self.solver_stats["pcg_iterations"] += iters
        self.solver_stats["pcg_log"].append(int(iters))
        self.solver_stats["sharded_solves"] = self.solver_stats.get("sharded_solves", 0) + 1
        if relres > max(rtol, 1e-15) * 10 and iters >= maxit:
            self.logger.warning("sharded PCG stopped at relative residual %.3e after %d iterations", relres, iters)
        return x

    def direct_solve(self, a, b, dim):
        """scalar problem: every dof = b / a (solver.py:909-925)."""
        f = Function(self.V[dim])
        vec = np.asarray(b, dtype=np.float64) / _f(a) if not np.isscalar(b) else _f(b) / _f(a)
        f.vector()[:] = vec
        return f

    def FD_solve(self, A, B, dim):
        """spsolve(A, B) of a user-assembled finite-difference system (solver.py:927-943) as a
        banded LU with partial pivoting on the device (k_banded_lu, bandwidths up to 64: FD_matrices operators have 1)."""
        M = _CsrOnDevice(A)
        b = _lib.to_device(np.ascontiguousarray(np.asarray(B, dtype=np.float64).ravel()))
        if max(M.kl, M.ku) <= 64:
            perm = torch.arange(M.n, dtype=I32, device=b.device)
            x, info = _lib.banded_solve(M.rowptr, M.colidx, M.values, b, perm, M.kl, M.ku)
            self._last_info = info
        else:
            raise NotImplementedError(
                "FD_solve: system with lower / upper bandwidth %d / %d; the device banded LU covers bandwidths up to 64 "
                "(there is no library or CPU fallback). Reorder the unknowns along the coordinate, or solve this dimension "
                'as "FEM".' % (M.kl, M.ku))
        self.solver_stats["banded_solves"] += 1
        return Function(self.V[dim], x)

    # ------------------------------------------------------------------ result
    def return_PGD(self):
        """Wrap meshes + modes into a PGD model (solver.py:883-907)."""
        from .model import PGD

        solution = PGD(name=self.name, n_modes=self.PGD_modes, fmeshes=self.meshes, pgd_modes=self.PGD_func,
                       name_coord=self.name_coord, modes_info=self.modes_info, verbose=False)
        solution.problem = self
        self.logger.info(solution._info_str())
        return solution


PGDProblem1 = PGDProblem


def newton_linearise(F, u):
    """One Newton step from u = 0 for a residual form F(u; v) that is affine in u
    (NonlinearVariationalSolver path, solver.py:579-595): returns (J, -F(0)) with
    J = derivative(F, u).  Raises for genuinely non-linear residuals."""
    J = derivative(F, u)
    for it in J.integrals:
        for m in it.monos:
            if any(f.leaf is u for f in m.factors):
                raise NotImplementedError("residual is not affine in the unknown (true Newton iterations are not implemented)")
    from .ufl import Integral

    rest = Form([Integral([m.scaled(-1.0) for m in it.monos if not any(f.leaf is u for f in m.factors)], it.measure)
                 for it in F.integrals])
    return J, rest


def FD_matrices(x):
    """1-D non-uniform finite-difference operators on the sorted coordinates x
    (pgdrome/solver.py:947-988): lumped mass M, second derivative D2, backward (upwind) first
    derivative D1_up scaled by the local mass.  Returned as SciPy ``lil_matrix`` like the reference.

    Built from diagonals instead of the reference's element loop; two reference quirks are kept on
    purpose because the restated tests pin them: the first D1_up row is the un-scaled (-1/2, 1/2)
    pair, and the last row re-uses the ``hp`` left over from the last interior node (solver.py:986-987)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    N = x.size
    if N < 2:
        raise ValueError("FD_matrices needs at least two coordinates")
    h = np.diff(x)
    hm, hp = h[:-1], h[1:]
    mass = np.empty(N)
    mass[0], mass[-1] = h[0] / 2, h[-1] / 2
    mass[1:-1] = (hp + hm) / 2
    d2 = np.empty(N)
    d2[0], d2[-1] = -1 / h[0], -1 / h[-1]
    d2[1:-1] = -(hp + hm) / (hp * hm)
    hp_last = h[-1] if N > 2 else h[0]  # stale hp of the reference loop
    w_last = (hp_last + h[-1]) / (2 * h[-1])
    d1 = np.empty(N)
    d1[0] = -0.5
    d1[1:-1] = (hp + hm) / (2 * hm)
    d1[-1] = w_last
    d1_lo = np.concatenate([-(hp + hm) / (2 * hm), [-w_last]])
    d1_up = np.zeros(N - 1)
    d1_up[0] = 0.5
    M = sp.diags([mass], [0], shape=(N, N)).tolil()
    D2 = sp.diags([np.concatenate([1 / hm, [1 / h[-1]]]), d2, np.concatenate([[1 / h[0]], 1 / hp])], [-1, 0, 1],
                  shape=(N, N)).tolil()
    D1 = sp.diags([d1_lo, d1, d1_up[:1]], [-1, 0, 1], shape=(N, N)).tolil() if N == 2 else None
    if D1 is None:
        D1 = sp.lil_matrix((N, N))
        D1.setdiag(d1)
        D1.setdiag(d1_lo, -1)
        D1[0, 1] = 0.5
    return M, D2, D1
