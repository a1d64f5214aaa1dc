"""-m gpu: the public API (PGDProblem.solve_PGD, PGD.evaluate) on the B200 against the CPU oracle.

Bar (north_star): each normalised mode and the reconstruction agree to 1e-8 relative L2 (up to
sign), at matching fixed-point iteration counts."""
import numpy as np
import pytest

from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob
from oracle.evaluate import evaluate_dofs

pytestmark = pytest.mark.gpu

MODE_RTOL = 1e-8


def _ospaces(p):
    return [ofem.Space(v.mesh().coordinates(), v.mesh().cells(), v.degree, v.bs) for v in p.V]


def _mode_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)


def _compare(p, o, tol=MODE_RTOL):
    assert p.PGD_modes == o.PGD_modes
    assert p.num_fp_it == o.num_fp_it
    for d in range(len(p.V)):
        for k in range(p.PGD_modes):
            assert _mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]) < tol, (d, k)
    assert np.allclose(p.amplitude, o.amplitude, rtol=1e-8, atol=0)


def test_config1_poisson1d_k_full_size():
    """BASELINE configs[0] at full size: 1000 nodes x 101 k-nodes, PGD_nmax = 10."""
    from pgdrome_b200 import configs

    p = configs.poisson1d_k()
    p.solve_PGD(_problem="linear")
    o, _ = oprob.poisson1d_k(spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    # analytic solution u = x(1-x)/(2k)
    pgd = p.return_PGD()
    x = p.V[0].tabulate_dof_coordinates()[:, 0]
    for k in (0.5, 1.234, 2.0):
        u = pgd.evaluate(0, [1], [k], 0).vector()[:]
        assert np.abs(u - x * (1 - x) / (2 * k)).max() < 2e-4  # PGD truncation at nmax = 10
    # reconstruction parity with the oracle's evaluate loop, single point and batched
    sk = _ospaces(p)[1]
    uo = evaluate_dofs(o.PGD_func[0], [sk], [o.PGD_func[1]], [1.234])
    u = pgd.evaluate(0, [1], [1.234], 0).vector()[:]
    assert np.linalg.norm(u - uo) / np.linalg.norm(uo) < MODE_RTOL
    ks = np.linspace(0.5, 2.0, 37)[:, None]
    U = pgd.evaluate_batch(0, [1], ks, 0).cpu().numpy()
    for c in (0, 11, 36):
        uo = evaluate_dofs(o.PGD_func[0], [sk], [o.PGD_func[1]], ks[c])
        assert np.linalg.norm(U[c] - uo) / np.linalg.norm(uo) < MODE_RTOL
    with pytest.raises(ValueError):
        pgd.evaluate(0, [1], [2.5], 0)


def test_config1_newton_equals_linear():
    from pgdrome_b200 import configs

    a = configs.poisson1d_k(nx=120, nk=20, PGD_nmax=5)
    a.solve_PGD(_problem="linear")
    b = configs.poisson1d_k(nx=120, nk=20, PGD_nmax=5)
    b.solve_PGD()  # default "nonlinear"
    assert np.allclose(a.amplitude, b.amplitude, rtol=1e-8, atol=0)


@pytest.mark.parametrize("n,nt,nk,nmax", [(16, 20, 6, 3), (48, 40, 10, 4)])
def test_config2_heat2d_tk_reduced(n, nt, nk, nmax):
    """BASELINE configs[1] at sizes the oracle's SuperLU finishes in seconds."""
    from pgdrome_b200 import configs

    p = configs.heat2d_tk(n=n, nt=nt, nk=nk, PGD_nmax=nmax)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.heat2d_tk(n=n, nt=nt, nk=nk, PGD_nmax=nmax, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    assert p.solver_stats["pcg_solves"] > 0 and p.solver_stats["banded_solves"] > 0


def test_config3_elasticity3d_reduced():
    """BASELINE configs[2] reduced: vector P1 tetrahedra (node-block-Jacobi PCG), two materials, traction."""
    from pgdrome_b200 import configs

    p = configs.elasticity3d(n=6, nE=8, nF=2, PGD_nmax=4)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.elasticity3d(n=6, nE=8, nF=2, PGD_nmax=4, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    assert p.PGD_modes == 4 and p.solver_stats["pcg_solves"] > 0


def test_config4_thermal3d_reduced():
    """BASELINE configs[3] reduced: P1 tetrahedra x FD time x power x speed, separated moving source."""
    from pgdrome_b200 import configs

    p = configs.thermal3d(n=8, nt=30, nP=4, nv=4, n_src=4, PGD_nmax=4)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.thermal3d(n=8, nt=30, nP=4, nv=4, n_src=4, PGD_nmax=4, spaces=_ospaces(p), source_terms=p.source_terms)
    opgd.solve_pgd(o)
    _compare(p, o)
    # reconstruction: batched evaluate on the device vs the oracle loop at a few parameter points
    pgd = p.return_PGD()
    S = _ospaces(p)
    pts = np.array([[0.3, 0.7, 1.2], [0.9, 1.4, 0.6]])
    U = pgd.evaluate_batch(0, [1, 2, 3], pts, 0).cpu().numpy()
    for c in range(len(pts)):
        uo = evaluate_dofs(o.PGD_func[0], S[1:], [o.PGD_func[d] for d in (1, 2, 3)], pts[c])
        assert np.linalg.norm(U[c] - uo) / np.linalg.norm(uo) < MODE_RTOL


def test_config2_heat2d_tk_full_size_against_time_stepping():
    """BASELINE configs[1] at FULL size (256x256 P1 x 200 time nodes x 50 k nodes, 20 modes) through a size-independent
    property: at a fixed conductivity the separated solution must reproduce the full-order model, i.e. implicit time
    stepping of  rho c M (u_i - u_{i-1})/dt + k K u_i = M q  with the same space / time operators (SciPy sparse LU, one
    factorisation).  The reference's own integration tests end with exactly this kind of check (PGD vs FOM at parameter
    samples, mean error < 1e-3 ... 1e-4).  What is left is the truncation at 20 modes and the P1 Galerkin projection in k."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    from pgdrome_b200 import configs

    p = configs.heat2d_tk(PGD_nmax=20)
    p.solve_PGD(_problem="linear")
    assert p.PGD_modes == 20 and p.V[0].n_dofs == 257 * 257 and p.V[1].n_dofs == 200 and p.V[2].n_dofs == 50
    assert all(b < a for a, b in zip(p.amplitude[:3], p.amplitude[1:4]))  # leading amplitudes decay
    o, _ = oprob.heat2d_tk(PGD_nmax=1, spaces=_ospaces(p))
    (rho_cp, (Mx, At, Mk)), (_, (Kx, Mt, Mkk)) = o.lhs_terms
    q_x, q_t, _ = o.rhs_terms[0][1]
    bcx, bct = o.bc_dofs[0], o.bc_dofs[1]
    t = p.V[1].tabulate_dof_coordinates().ravel()
    order = np.argsort(t)
    assert list(bct) == [order[0]]
    At, Mt = sp.csr_matrix(At)[order][:, order], sp.csr_matrix(Mt)[order][:, order]
    # causal in time (row 0 is the initial-condition row, eliminated): forward substitution over the time nodes
    assert sp.triu(At[1:], 2).nnz == 0 and sp.triu(Mt[1:], 2).nnz == 0
    k_nodes = p.V[2].tabulate_dof_coordinates().ravel()
    X = np.stack([f.vector().get_local() for f in p.PGD_func[0]])
    T = np.stack([f.vector().get_local() for f in p.PGD_func[1]])[:, order]
    Kk = np.stack([f.vector().get_local() for f in p.PGD_func[2]])
    free = np.setdiff1d(np.arange(Mx.shape[0]), bcx)
    Mf, Kf = sp.csr_matrix(Mx)[free][:, free], sp.csr_matrix(Kx)[free][:, free]
    qf = np.asarray(q_x)[free]
    qt = np.asarray(q_t)[order]
    worst = 0.0
    for jk in (7, 30):
        k0 = float(k_nodes[jk])
        U = np.zeros((len(t), len(free)))
        lu, lu_key = None, None
        for i in range(1, len(t)):
            aii, mii = At[i, i], Mt[i, i]
            rhs = qt[i] * qf
            for j in At[i].indices:
                if j < i and j > 0:
                    rhs = rhs - rho_cp * At[i, j] * (Mf @ U[j])
            for j in Mt[i].indices:
                if j < i and j > 0:
                    rhs = rhs - k0 * Mt[i, j] * (Kf @ U[j])
            key = (round(aii, 12), round(mii, 12))
            if key != lu_key:
                lu, lu_key = spla.splu((rho_cp * aii * Mf + k0 * mii * Kf).tocsc()), key
            U[i] = lu.solve(rhs)
        W = (T * Kk[:, jk][:, None]).T @ X[:, free]  # [nt, n_free] reconstruction at k0
        err = np.linalg.norm(W - U) / np.linalg.norm(U)
        worst = max(worst, err)
    print("heat2d_tk full size: worst relative space-time error vs time stepping", worst)
    assert worst < 2e-3  # measured 6.5e-4 with 20 modes


# ------------------------------------------------------------------------------ parity at the sizes that are benchmarked
def _compare_pinned(make, solve_kw, o, tol=MODE_RTOL, n_modes=None):
    """Product run against the oracle run o at identical sweep counts (a run whose counts differ -- the "norm" test on
    its round-off floor -- is repeated with the oracle's schedule pinned; the tolerance stays)."""
    p = make()
    p.solve_PGD(**solve_kw)
    n_modes = min(p.PGD_modes, o.PGD_modes) if n_modes is None else n_modes
    if list(p.num_fp_it[:n_modes]) != list(o.num_fp_it[:n_modes]):
        first = list(p.num_fp_it)
        p = make()
        p.fp_schedule = list(o.num_fp_it)
        p.solve_PGD(**solve_kw)
        assert list(p.num_fp_it[:n_modes]) == list(o.num_fp_it[:n_modes]), (first, p.num_fp_it, o.num_fp_it)
    worst = 0.0
    for d in range(len(p.V)):
        for k in range(n_modes):
            worst = max(worst, _mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]))
    assert worst < tol, worst
    assert np.allclose(p.amplitude[:n_modes], o.amplitude[:n_modes], rtol=1e-7, atol=0)
    return p, worst


def _restarted_mode_errors(make, solve_kw, o):
    """Per-mode parity WITHOUT compounding: enrichment step k of the product is run from the ORACLE's modes 0..k-1
    (and the oracle's sweep count), and its new mode is compared with the oracle's mode k.  Two implementations of the
    enrichment differ by ~4x more per mode when every run feeds on its own previous modes (oracle LU against oracle CG at
    rtol 1e-13: 1e-15 at mode 0, 2e-7 at mode 19 of configs[1]); restarted, every step is an independent check."""
    from pgdrome_b200.functions import Function

    p = make()
    st = p.begin_PGD(**solve_kw)
    D = len(p.V)
    p.fp_schedule = list(o.num_fp_it)
    errs = []
    for k in range(o.PGD_modes):
        for d in range(D):
            prev = []
            for i in range(k):
                f = Function(p.V[d], np.asarray(o.PGD_func[d][i]))
                f.stable = True
                prev.append(f)
            p.PGD_func[d] = prev
        p.num_fp_it, p.err_fp_it, p.alpha = p.num_fp_it[:k], p.err_fp_it[:k], p.alpha[:k]
        st["n_enr"], st["normConv"], st["relConv"], st["done"] = k - 1, [1.0] * k, [1.0] * k, False
        p.step_PGD(st)
        assert len(p.PGD_func[0]) == k + 1 and p.num_fp_it[k] == o.num_fp_it[k]
        errs.append(max(_mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]) for d in range(D)))
    return errs


def test_config2_heat2d_tk_full_size_20_modes_against_oracle():
    """BASELINE configs[1] at FULL size (66 049 dofs x 200 x 50) and full mode count (20) against oracle.solve_pgd
    (SuperLU = the reference's direct-LU class), identical fixed-point sweep counts:
      * every one of the 20 modes to 1e-8 when each step starts from the same previous modes (restarted comparison);
      * the free-running 20-mode solve: the leading 10 modes to 1e-8 and the RECONSTRUCTION (what evaluate returns) to
        1e-8 at parameter points -- later modes inherit the differences of all earlier ones (see _restarted_mode_errors)."""
    from pgdrome_b200 import configs

    make = lambda: configs.heat2d_tk(n=256, nt=199, nk=49, PGD_nmax=20)
    S = _ospaces(make())
    o, _ = oprob.heat2d_tk(n=256, nt=199, nk=49, PGD_nmax=20, spaces=S)
    opgd.solve_pgd(o)
    assert o.PGD_modes == 20
    errs = _restarted_mode_errors(make, dict(_problem="linear"), o)
    assert max(errs) < 1e-8, errs
    p, worst = _compare_pinned(make, dict(_problem="linear"), o, tol=1e-8, n_modes=10)
    assert p.PGD_modes == 20 and list(p.num_fp_it) == list(o.num_fp_it)
    pgd = p.return_PGD()
    pts = np.array([[0.25, 0.7], [0.6, 1.3], [1.0, 1.9]])
    U = pgd.evaluate_batch(0, [1, 2], pts, 0).cpu().numpy()
    for c in range(len(pts)):
        uo = evaluate_dofs(o.PGD_func[0], S[1:], [o.PGD_func[1], o.PGD_func[2]], pts[c])
        assert np.linalg.norm(U[c] - uo) / np.linalg.norm(uo) < 1e-8


def test_config3_elasticity3d_persistent_kernel_against_oracle():
    """configs[2] at n = 22 (36 501 dofs: the smallest size that takes the HBM-regime code path -- persistent
    cooperative PCG with the node-block walk) against the oracle in CG mode (C node-block-Jacobi PCG), 3 modes."""
    from pgdrome_b200 import _lib, configs

    kw = dict(n=22, nE=12, nF=2, PGD_nmax=3)
    make = lambda: configs.elasticity3d(**kw)
    o, _ = oprob.elasticity3d(spaces=_ospaces(make()), solver="ccg", cg_block=[3, 1, 1], **kw)
    opgd.solve_pgd(o)
    s0 = _lib.stats()
    p, worst = _compare_pinned(make, dict(_problem="linear", settings={"linear_solver": "cg"}), o, tol=1e-8)
    s1 = _lib.stats()
    assert p.PGD_modes == 3
    assert s1["pcg_solves"] > s0["pcg_solves"] and s1["pcg_resident_solves"] == s0["pcg_resident_solves"]  # not the SM-resident solver
    # node-block-Jacobi PCG on both sides; the product warm-starts every solve from the previous sweep's mode, the oracle
    # starts from zero: never more iterations than the oracle, and the very first (cold) solve agrees within a few
    assert p.solver_stats["pcg_iterations"] <= sum(o.cg_iterations) + 5 * len(o.cg_iterations)
    assert abs(p.solver_stats["pcg_log"][0] - o.cg_iterations[0]) <= 0.02 * o.cg_iterations[0] + 3


def test_config4_thermal3d_streaming_kernels_against_oracle():
    """configs[3] at n = 32 (35 937 dofs: three-launch streaming PCG with the TMA SpMV) against the oracle in CG mode."""
    from pgdrome_b200 import configs

    kw = dict(n=32, nt=40, nP=5, nv=5, n_src=3, PGD_nmax=2)
    make = lambda: configs.thermal3d(**kw)
    first = make()
    o, _ = oprob.thermal3d(spaces=_ospaces(first), solver="ccg", source_terms=first.source_terms, **kw)
    opgd.solve_pgd(o)
    p, worst = _compare_pinned(make, dict(_problem="linear", solve_modes=None, settings={"linear_solver": "cg"}), o, tol=1e-8)
    assert p.PGD_modes == 2
