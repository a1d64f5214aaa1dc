"""Oracle (test infrastructure): the progressive PGD enrichment of pgdrome/solver.py, restated
on explicit matrices (NumPy/SciPy, direct SuperLU solves like the reference's MUMPS LU).

The reference builds every sub-problem through user UFL callbacks; algebraically each callback
set has the separated shape (e.g. tests/integration/test_elastic.py:71-219)

    LHS_d = sum_k c_k [prod_{j!=d} F_j^T K_{j,k} F_j] K_{d,k}
    RHS_d = sum_m c_m [prod_{j!=d} F_j^T g_{j,m}] g_{d,m}
          - sum_{i<n} sum_k c_k [prod_{j!=d} F_j^T K_{j,k} U_{j,i}] K_{d,k} U_{d,i}

with row = test function, col = trial function.  ``SeparatedProblem`` holds K, g, the mass
matrices used by ``dolfin.norm`` / ``MM`` and the Dirichlet dofs; ``solve_pgd`` follows
solver.py line by line:  get_Fsinit :158-304, residual check :347-395, FP_solve :508-881
(stop_fp "norm" :812-871 and "delta" :763-811), normalisation :406-470, stopping :476-504.

PARITY STATUS (see DESIGN.md "Parity status"):
  * PINNED to round-off against the unmodified reference: FD_matrices, and the whole enrichment loop
    (get_Fsinit, residual check, FP_solve "norm"/"delta", normalisation "no"/"stiff"/"l2", stopping) in
    FD mode -- tests/golden/laplace_fd.npz + fd_matrices.npz were produced by running
    pgdrome/solver.py itself (tests/golden/make_golden.py), checked in tests/test_oracle_golden.py.
  * parity unpinned for everything that goes through DOLFIN's FEM assembly (no FEniCS in this
    image, no golden vectors in the reference): pinned only to the tolerances of the reference's own
    tests, restated in tests/test_oracle_kat.py.
This module is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs import it.
"""
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .fem import apply_dirichlet_sym


@dataclass
class SeparatedProblem:
    n_dofs: list  # per dim
    mass: list  # per dim: matrix used for norms (consistent FEM mass, or the FD ``MM[d]``)
    bc_dofs: list  # per dim: int array (possibly empty); all reference BC values are 0
    lhs_terms: list  # [(c_k, [K_{0,k}, ..., K_{D-1,k}])]
    rhs_terms: list  # [(c_m, [g_{0,m}, ..., g_{D-1,m}])]
    seq_fp: list = None
    PGD_nmax: int = 20
    PGD_tol: float = 1e-10
    max_fp_it: int = 50
    tol_fp_it: float = 1e-5
    stop_fp: str = "norm"
    fp_init: str = ""
    norm_modes: str = "stiff"
    solver: str = "lu"  # "lu" (reference default: direct) | "cg" (settings={"linear_solver":"cg","preconditioner":"jacobi"},
    #                      SciPy) | "ccg" (the same algorithm in C + OpenMP with node-block Jacobi, oracle/cfem.c; dimensions
    #                      below `ccg_min_dofs` keep the direct solve)
    cg_rtol: float = 1e-13
    cg_block: list = None  # per dim: node-block size of the Jacobi preconditioner in "ccg" mode (default 1)
    ccg_min_dofs: int = 5000
    # outputs (same names as solver.py:105-126)
    PGD_func: list = field(default_factory=list)
    alpha: list = field(default_factory=list)
    amplitude: list = field(default_factory=list)
    num_fp_it: list = field(default_factory=list)
    err_fp_it: list = field(default_factory=list)
    res_errors: list = field(default_factory=list)
    fp_floor: list = field(default_factory=list)
    cg_iterations: list = field(default_factory=list)
    PGD_modes: int = None

    @property
    def D(self):
        return len(self.n_dofs)


def _mprod(M, f, g):
    """f^T M g in the reference's association order ``f.T @ MM @ g`` = (f^T M) g
    (solver.py:198-207, 748-752, 819-835): the "norm" stopping test cancels ~1e6-sized products
    down to round-off, so the order of the floating-point operations decides its branch."""
    return (f @ M) @ g


def _mnorm(M, f):
    return np.sqrt(_mprod(M, f, f))


def get_Fsinit(p, rng=None):
    """solver.py:158-304: ones -> Dirichlet (value 0) -> optional rand -> / sqrt(f^T M f)."""
    Fs = []
    for d in range(p.D):
        f = np.ones(p.n_dofs[d])
        f[p.bc_dofs[d]] = 0.0
        if p.fp_init.lower() == "randomized":
            idx = np.where(f != 0)[0]
            f[idx] = (rng or np.random).rand(len(idx))
        f *= 1.0 / _mnorm(p.mass[d], f)
        Fs.append(f)
    return Fs


def lhs_matrix(p, Fs, d):
    A = None
    for c, Ks in p.lhs_terms:
        s = c
        for j in range(p.D):
            if j != d:
                s = s * (Fs[j] @ (Ks[j] @ Fs[j]))
        A = s * Ks[d] if A is None else A + s * Ks[d]
    return A.tocsr()


def rhs_vector(p, Fs, d, n_enr):
    b = np.zeros(p.n_dofs[d])
    for c, gs in p.rhs_terms:
        s = c
        for j in range(p.D):
            if j != d:
                s = s * (Fs[j] @ gs[j])
        b += s * gs[d]
    for i in range(n_enr):
        for c, Ks in p.lhs_terms:
            s = c
            for j in range(p.D):
                if j != d:
                    s = s * (Fs[j] @ (Ks[j] @ p.PGD_func[j][i]))
            b -= s * (Ks[d] @ p.PGD_func[d][i])
    return b


def _solve(p, A, b, d):
    A, b = apply_dirichlet_sym(A, b, p.bc_dofs[d])
    if p.solver == "ccg" and A.shape[0] >= p.ccg_min_dofs:
        from . import cfem

        x, it, rr = cfem.pcg(A, b, block=(p.cg_block[d] if p.cg_block else 1), rtol=p.cg_rtol, max_iters=20 * A.shape[0])
        p.cg_iterations.append(it)
        return x
    if p.solver == "cg":
        dinv = 1.0 / A.diagonal()
        x, info = spla.cg(A, b, rtol=p.cg_rtol, atol=0.0, maxiter=20 * A.shape[0],
                          M=spla.LinearOperator(A.shape, lambda r: dinv * r))
        return x
    return spla.spsolve(A.tocsc(), b)


def FP_solve(p, Fs_init, norm_Fs, n_enr):
    """solver.py:508-881."""
    Fs = [f.copy() for f in Fs_init]
    seq = p.seq_fp if p.seq_fp is not None and len(p.seq_fp) else list(range(p.D))
    delta = np.ones(p.D)
    for fpi in range(p.max_fp_it):
        for d in seq:  # Gauss-Seidel: later dims see the updated Fs[d]
            A = lhs_matrix(p, Fs, d)
            b = rhs_vector(p, Fs, d, n_enr)
            Fs[d] = _solve(p, A, b, d)
            norm_Fs[d] = _mnorm(p.mass[d], Fs[d])
        if p.stop_fp.lower() == "delta":
            for d in range(p.D):
                dt = np.abs(Fs[d] - Fs_init[d])
                mi = np.argmax(dt)
                delta[d] = dt.max() if abs(Fs[d][mi]) < 1e-8 else dt.max() / abs(Fs[d][mi])
            notconv = np.any(delta > p.tol_fp_it)
            if notconv and fpi < p.max_fp_it - 1:
                Fs_init = [f.copy() for f in Fs]
            else:
                p.num_fp_it.append(fpi + 1)
                p.err_fp_it.append(delta.copy())
                break
        elif p.stop_fp.lower() == "norm":
            newnew = newold = oldold = 1.0
            for d in range(p.D):
                newnew *= _mprod(p.mass[d], Fs[d], Fs[d])
                newold *= _mprod(p.mass[d], Fs[d], Fs_init[d])
                oldold *= _mprod(p.mass[d], Fs_init[d], Fs_init[d])
            err = np.sqrt(np.abs(newnew + oldold - 2 * newold))
            if err < p.tol_fp_it or fpi == p.max_fp_it - 1:
                p.num_fp_it.append(fpi + 1)
                p.err_fp_it.append(err)
                # diagnostic (not in the reference): round-off floor of this cancellation
                p.fp_floor.append(np.sqrt(np.finfo(float).eps * (newnew + oldold + 2 * abs(newold))))
                break
            Fs_init = [f.copy() for f in Fs]
        else:
            raise ValueError('stopping criterion not defined (self.stop_fp = "delta" or "norm")')
    return Fs, norm_Fs


def solve_pgd(p, rng=None, step_hook=None):
    """solver.py:306-506.  step_hook(n_enr) is called after every completed enrichment step
    (bench.py's CPU legs time the steps with it)."""
    D = p.D
    p.PGD_func = [[] for _ in range(D)]
    p.alpha, p.num_fp_it, p.err_fp_it, p.res_errors, p.fp_floor = [], [], [], [], []
    normConv, relConv = [], []
    n_enr = -1
    while n_enr < p.PGD_nmax - 1:
        n_enr += 1
        Fs_init = get_Fsinit(p, rng)
        norm_Fs = np.array([_mnorm(p.mass[d], Fs_init[d]) for d in range(D)])
        res = []
        for d in range(D):
            ll = rhs_vector(p, Fs_init, d, n_enr)
            ll[p.bc_dofs[d]] = 0.0
            res.append(ll @ ll)
        res_error = np.sqrt(np.sum(res))
        p.res_errors.append(res_error)
        if res_error < 1e-10:
            break
        Fs, norm_Fs = FP_solve(p, Fs_init, norm_Fs, n_enr)
        normU = np.prod(norm_Fs)
        mode = p.norm_modes.lower()
        if mode == "no":
            for d in range(D):
                p.PGD_func[d].append(Fs[d])
            p.alpha.append(1.0)
        elif mode == "stiff":
            Fn = [Fs[d] * (1 / norm_Fs[d]) for d in range(D)]
            norm_aux = 0.0
            for c, Ks in p.lhs_terms:  # lhs_fct(Fn[-1], Fn[-1], Fn, ..., prob[-1], D) assembled as a scalar
                s = c
                for j in range(D):
                    s = s * (Fn[j] @ (Ks[j] @ Fn[j]))
                norm_aux += s
            norm_fac = np.sqrt(np.absolute(norm_aux)) ** (1.0 / D)
            p.alpha.append(np.prod(norm_Fs) * norm_fac**D)
            for d in range(D):
                Fn[d] = Fn[d] * (1.0 / norm_fac)
                Fn[d] = Fn[d] * p.alpha[-1] ** (1.0 / D)
                p.PGD_func[d].append(Fn[d])
        elif mode == "l2":
            p.alpha.append(normU)
            norm_all = np.prod(norm_Fs) ** (1.0 / D)
            for d in range(D):
                p.PGD_func[d].append((norm_all / norm_Fs[d]) * Fs[d])
        normConv.append(normU)
        relConv.append(normU / normConv[0])
        if step_hook is not None:
            step_hook(n_enr)
        if relConv[n_enr] < p.PGD_tol:
            break
    p.amplitude = relConv
    p.PGD_modes = len(p.PGD_func[0])
    return p


def FD_matrices(x):
    """pgdrome/solver.py:947-988 restated with array arithmetic (CSR output, same values).

    M: lumped mass (half weights at both ends); D2: 3-point second derivative with one-sided
    2-entry end rows; D1_up: backward difference scaled by (hp+hm)/2 -- i.e. M * (F_i-F_{i-1})/hm,
    first row the un-scaled (-1/2, 1/2) pair (:962-963) and last row re-using the *stale* hp of
    the final loop pass (:986-987), which makes it exactly (1, -1)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    N = len(x)
    h = np.diff(x)
    hp, hm = h[1:], h[:-1]  # interior i=1..N-2
    M = np.zeros(N)
    M[0], M[-1] = h[0] / 2, h[-1] / 2
    M[1:-1] = (hp + hm) / 2
    d2_main = np.zeros(N)
    d2_main[0], d2_main[-1] = -1 / h[0], -1 / h[-1]
    d2_main[1:-1] = -(hp + hm) / (hp * hm)
    d2_up = np.concatenate([[1 / h[0]], 1 / hp])
    d2_lo = np.concatenate([1 / hm, [1 / h[-1]]])
    d1_main = np.zeros(N)
    d1_main[0] = -1 / 2
    d1_main[1:-1] = (hp + hm) / (2 * hm)
    hp_stale = h[-1] if N > 2 else h[0]
    d1_main[-1] = (hp_stale + h[-1]) / (2 * h[-1])
    d1_lo = np.concatenate([-(hp + hm) / (2 * hm), [-(hp_stale + h[-1]) / (2 * h[-1])]])
    d1_up = np.zeros(N - 1)
    d1_up[0] = 1 / 2
    Mm = sp.diags(M).tocsr()
    D2 = sp.diags([d2_lo, d2_main, d2_up], [-1, 0, 1]).tocsr()
    D1 = sp.diags([d1_lo, d1_main, d1_up], [-1, 0, 1]).tocsr()
    D1.eliminate_zeros()
    return Mm, D2, D1
