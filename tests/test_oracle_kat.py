"""-m "not gpu": the reference's own known-answer tests, restated on the oracle (NumPy/SciPy).

Every test of the reference is a tolerance test against an analytic / independent solution
(SURVEY.md section 4); none needs DOLFIN once meshes, dofmaps and element matrices come from
oracle/.  These pin the FEM side of the oracle (assembly + enrichment loop) to the tolerances the
reference's CI enforces; the FD side is pinned to round-off by tests/test_oracle_golden.py.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import evaluate as oev
from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob
from oracle.meshes import interval_mesh


def test_FD_matrices_equal_backward_euler():
    """tests/unit/test_FD.py:147-170: rho c D1_up T = M q with row/col BC == explicit backward Euler
    (errornorm < 1e-8), and the FEM-in-time variant is worse."""
    rho, cp, Tamb, P = 71.0, 31.0, 25.0, 250.0
    S = ofem.Space(*interval_mesh(200, 0.0, 50.0))
    t = S.dof_coordinates().ravel()
    q = np.where(t < 5, 0.0, np.where(t > 20, 0.0, P))  # Expression degree=1 => nodal interpolation
    srt = np.argsort(t)
    ts = t[srt]
    # reference solution: explicit recursion (test_FD.py:24-40)
    Tr = np.zeros(len(ts))
    Tr[0] = Tamb
    qs = q[srt]
    for i in range(1, len(ts)):
        Tr[i] = Tr[i - 1] + (ts[i] - ts[i - 1]) / (rho * cp) * qs[i]
    Tref = np.empty_like(Tr)
    Tref[srt] = Tr
    # FD solution (test_FD.py:51-87): IC sits on the LAST dof (DOLFIN 1-D numbering is reversed)
    Mt, _, D1 = opgd.FD_matrices(ts)
    M1, D1u = Mt[srt, :][:, srt], D1[srt, :][:, srt]
    assert t[-1] == 0.0
    IC = np.zeros(len(t))
    IC[-1] = Tamb
    A = sp.lil_matrix(rho * cp * D1u)
    F = M1 @ q - rho * cp * D1u @ IC
    F[-1] = 0
    A[:, -1] = 0
    A[-1, :] = 0
    A[-1, -1] = 1
    TFD = spla.spsolve(A.tocsr(), F) + IC
    M = ofem.assemble_bilinear(S, ofem.T_mass(1, 1))
    errnorm = lambda a, b: np.sqrt((a - b) @ (M @ (a - b)))
    e1 = errnorm(TFD, Tref)
    # FEM in time (test_FD.py:90-124): rho c int T' v = int q v, T(0) = Tamb
    Aadv = rho * cp * ofem.assemble_bilinear(S, ofem.T_adv(1))
    b = M @ q
    bc = np.where(t < 1e-5)[0]
    A2, b2 = ofem.apply_dirichlet_sym(Aadv.tocsr(), b, bc, values=Tamb)
    TFEM = spla.spsolve(A2.tocsc(), b2)
    e2 = errnorm(TFEM, Tref)
    assert e1 < 1e-8
    assert e2 > e1


def test_truss_xpe_against_analytic():
    """tests/integration/test_elastic.py:325-380: P2, meshes [113,2,100], mean error over 10 LHS
    samples (seed 3452) < 1e-4, single-point error < 1e-5."""
    p, info = oprob.truss_xpe()
    opgd.solve_pgd(p)
    S = info["spaces"]
    x = S[0].dof_coordinates().ravel()
    smp = oev.sampling_LHS([-1.0, 0.2], [3.0, 2.0], 10)
    fom = lambda s: s[0] / (2 * s[1]) * (-x * x + x)
    pe = lambda s: oev.evaluate_dofs(p.PGD_func[0], S[1:], p.PGD_func[1:], s)
    _, mean, _ = oev.evaluate_error(fom, pe, smp)
    assert mean < 1e-4
    errs = []
    for s in ([2.0, 1.5], [1.0, 1.0]):
        u = pe(s)
        uh = oev.point_eval(S[0], u, [0.5])
        ref = s[0] / (2 * s[1]) * 0.25
        errs.append(abs(uh - ref) / abs(ref))
    assert np.mean(errs) < 1e-5


def _laplace_fem_reference(x_out, q, u0, nx=60, k=0.5, lx=3.0):
    """FEM_reference of test_laplace.py:866-930 restated: -k lap(u) = Q, u(0,y) = u0, u(lx,y) = 0 with
    P2 elements and Q = Expression("x[0]<L/2 ? q : 0", degree=1), i.e. the nodal P1 interpolant of the
    step (a ramp over one element).  Nothing depends on y, so the 2-D P2 solution restricted to a line
    y = const is the 1-D P2 solution computed here."""
    S = ofem.Space(*interval_mesh(nx, 0.0, lx), degree=2)
    xv = np.linspace(0.0, lx, nx + 1)
    Qh = lambda x: q * np.interp(x[..., 0], xv, np.where(xv < lx / 2, 1.0, 0.0))
    A = k * ofem.assemble_bilinear(S, ofem.T_stiff(1, 1))
    L = np.zeros((1, 2))
    L[0, 0] = 1.0
    b = ofem.assemble_linear(S, L, weight=Qh, weight_degree=1)
    xd = S.dof_coordinates().ravel()
    bc = np.where((np.abs(xd) < 1e-6) | (np.abs(xd - lx) < 1e-6))[0]
    A, b = ofem.apply_dirichlet_sym(A, b, bc, values=np.where(xd[bc] < 1e-6, u0, 0.0))
    u = spla.spsolve(A.tocsc(), b)
    return np.array([oev.point_eval(S, u, [xx]) for xx in x_out])


def test_laplace_one_mode_and_error():
    """tests/integration/test_laplace.py:957-1092: FEM and FD variants converge in exactly 1 mode;
    mean relative error vs the 2-D FEM solution < 1e-6 (FEM) / < 2e-4 (FD), BC lifting added back."""
    rng = np.random.default_rng(0)
    checks = [[rng.uniform(0, 3), rng.uniform(0, 50), rng.uniform(10, 50)] for _ in range(10)]
    for kind, tol in (("FEM", 1e-6), ("FD", 2e-4)):
        p, info = oprob.laplace_xyqu(kind=kind)
        opgd.solve_pgd(p)
        assert p.PGD_modes == 1
        S, BC = info["spaces"], info["BC"]
        x = S[0].dof_coordinates().ravel()
        errs = []
        for c in checks:
            u = oev.evaluate_dofs(p.PGD_func[0], S[1:], p.PGD_func[1:], c)
            lift = BC[0] * np.prod([oev.point_eval(S[i + 1], BC[i + 1], [c[i]]) for i in range(3)])
            ref = _laplace_fem_reference(x, c[1], c[2])
            errs.append(np.linalg.norm(u + lift - ref) / np.linalg.norm(ref))
        assert np.mean(errs) < tol, (kind, np.mean(errs))


def test_heat1d_cooling_fem_vs_fd():
    """tests/integration/test_heat1D.py:806-904 (cooling case): FEM-in-time and FD-in-time PGD
    solutions converge and agree with each other at the level the reference asserts vs FEM."""
    sols = {}
    for kind in ("FEM", "FD"):
        p, info = oprob.heat1d(kind=kind, case="cooling")
        opgd.solve_pgd(p)
        assert 1 <= p.PGD_modes <= 20 and p.amplitude[-1] < 1e-3
        S = info["spaces"]
        u = oev.evaluate_dofs(p.PGD_func[0], S[1:], p.PGD_func[1:], [0.5, 0.75])
        sols[kind] = u
    rel = np.linalg.norm(sols["FEM"] - sols["FD"]) / np.linalg.norm(sols["FEM"])
    assert rel < 5e-2  # first-order backward Euler vs P1-in-time on 10 time elements


def test_elasticity2d_reduced_linear_pcg_vs_lu():
    """tests/integration/test_solver_problem.py:748-752 pins linear-vs-Newton amplitudes to 1e-8;
    the analogue here: direct LU vs Jacobi-CG (rtol 1e-13) sub-solves give the same amplitudes
    (the B200 path replaces MUMPS by PCG).  Reduced mesh so the CPU suite stays short."""
    a, _ = oprob.elasticity2d(N=(20, 4), numElems=(2, 10, 10), PGD_nmax=3)
    opgd.solve_pgd(a)
    b, _ = oprob.elasticity2d(N=(20, 4), numElems=(2, 10, 10), PGD_nmax=3, solver="cg")
    opgd.solve_pgd(b)
    # the "norm" criterion sits on its round-off floor for the first mode (floor 3e-4 > tol_fp_it 1e-4:
    # the loop stops when the cancellation happens to return exactly 0.0), so one sweep more or less
    for n in range(3):
        assert a.num_fp_it[n] == b.num_fp_it[n] or min(a.err_fp_it[n], b.err_fp_it[n]) < 4 * a.fp_floor[n]
    assert np.allclose(a.amplitude, b.amplitude, rtol=1e-7, atol=0)
