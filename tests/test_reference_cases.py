"""The reference's integration tests, restated end to end on the B200 path (tests/ref_cases.py are the
user scripts, oracle/problems.py the matrix-form restatements they are compared with), plus the
assertions the reference's own tests make.

Each case runs twice: `-m gpu` on the real kernels, and `-m "not gpu"` through the NumPy ABI stand-in
(host logic: UFL capture of P2 / vector P2 / Voigt / facet / degree-10 weights / lifting terms, Newton
path, mixed FEM-FD solve modes)."""
import numpy as np
import pytest

from oracle import evaluate as oev
from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob
from tests import cpu_abi, ref_cases

MODE_RTOL = 1e-8


@pytest.fixture
def cpu(monkeypatch):
    import pgdrome_b200.forms  # noqa: F401
    import pgdrome_b200.model  # noqa: F401
    import pgdrome_b200.solver  # noqa: F401

    cpu_abi.install(monkeypatch)
    from pgdrome_b200 import lazy

    lazy._pending.clear()
    yield


def _ospaces(p):
    return [ofem.Space(v.mesh().coordinates(), v.mesh().cells(), v.degree, v.bs) for v in p.V]


def _mode_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)


def _compare(p, o, n_modes=None, tol=MODE_RTOL, floor_ok=False, rerun=None):
    """Modes / amplitudes of the product run p against the oracle run o at IDENTICAL sweep counts.  The counts must
    agree by themselves, except (floor_ok) where the "norm" test sits on its round-off floor: there the
    product is run again with the oracle's sweep schedule pinned (``rerun(schedule) -> p``) and compared at the same
    tolerance -- the tolerance is never widened."""
    n_modes = min(p.PGD_modes, o.PGD_modes) if n_modes is None else n_modes
    for n in range(n_modes):
        same = p.num_fp_it[n] == o.num_fp_it[n]
        at_floor = floor_ok and min(float(np.max(p.err_fp_it[n])), float(np.max(o.err_fp_it[n]))) < 4 * o.fp_floor[n]
        assert same or at_floor, (n, p.num_fp_it, o.num_fp_it)
    if list(p.num_fp_it[:n_modes]) != list(o.num_fp_it[:n_modes]):
        assert rerun is not None, (p.num_fp_it, o.num_fp_it)
        p = rerun(list(o.num_fp_it))
        assert list(p.num_fp_it[:n_modes]) == list(o.num_fp_it[:n_modes])
    for d in range(len(p.V)):
        for k in range(n_modes):
            assert _mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]) < tol, (d, k)
    assert np.allclose(p.amplitude[:n_modes], o.amplitude[:n_modes], rtol=1e-7, atol=0)
    return p


def _pgd_point(p, coord):
    pgd = p.return_PGD()
    return pgd.evaluate(0, list(range(1, len(p.V))), list(coord), 0).vector()[:]


# ------------------------------------------------------------------------------ cases
def case_truss():
    """test_elastic.py:325-380: P2, default Newton path; mean error over 10 LHS samples < 1e-4."""
    p = ref_cases.truss_xpe()
    p.solve_PGD()  # _problem="nonlinear", settings mumps: the reference's defaults
    o, _ = oprob.truss_xpe(spaces=_ospaces(p))
    opgd.solve_pgd(o)
    assert p.PGD_modes == o.PGD_modes == 1
    _compare(p, o)
    x = p.V[0].tabulate_dof_coordinates()[:, 0]
    errs = []
    for s in oev.sampling_LHS([-1.0, 0.2], [3.0, 2.0], 10):
        ref = s[0] / (2 * s[1]) * (-x * x + x)
        errs.append(np.linalg.norm(_pgd_point(p, s) - ref) / np.linalg.norm(ref))
    assert np.mean(errs) < 1e-4


def case_heat1d(kind, case):
    """test_heat1D.py:672-904: FEM-in-time and FD-in-time PGD against the oracle restatement."""
    p = ref_cases.heat1d(kind, case=case)
    p.solve_PGD(_problem="linear", solve_modes=p._solve_modes)
    o, _ = oprob.heat1d(kind, case=case, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    if case == "heating":
        assert p.PGD_modes == o.PGD_modes
        _compare(p, o, tol=1e-7)  # 14-20 modes, several sweeps at max_fp_it: differences accumulate
    else:
        # cooling: amplitudes fall to 1e-5 within 3 modes and the later modes fit round-off-sized
        # residuals (ill-conditioned fixed points): compare the leading modes and the reconstruction
        _compare(p, o, n_modes=2)
        S = _ospaces(p)
        for coord in ([0.3, 0.75], [0.9, 0.55]):
            uo = oev.evaluate_dofs(o.PGD_func[0], S[1:], o.PGD_func[1:], coord)
            assert np.linalg.norm(_pgd_point(p, coord) - uo) <= 1e-4 * np.linalg.norm(uo)


def case_laplace_fem():
    """test_laplace.py:957-1092 (FEM variant): exactly one mode; error vs the 2-D FEM reference < 1e-6."""
    from tests.test_oracle_kat import _laplace_fem_reference

    p, BC = ref_cases.laplace_fem()
    p.solve_PGD(_problem="linear")
    o, info = oprob.laplace_xyqu("FEM", spaces=_ospaces(p))
    opgd.solve_pgd(o)
    assert p.PGD_modes == 1 == o.PGD_modes

    def pinned(schedule):
        q, _ = ref_cases.laplace_fem()
        q.fp_schedule = schedule
        return q.solve_PGD(_problem="linear")

    p = _compare(p, o, floor_ok=True, rerun=pinned)
    x = p.V[0].tabulate_dof_coordinates()[:, 0]
    rng = np.random.default_rng(0)
    errs = []
    for _ in range(5):
        c = [rng.uniform(0, 3), rng.uniform(0, 50), rng.uniform(10, 50)]
        lift = BC[0].vector()[:] * BC[1](c[0]) * BC[2](c[1]) * BC[3](c[2])
        ref = _laplace_fem_reference(x, c[1], c[2])
        errs.append(np.linalg.norm(_pgd_point(p, c) + lift - ref) / np.linalg.norm(ref))
    assert np.mean(errs) < 1e-6


def case_elasticity2d():
    """test_solver_problem.py:531-627 reduced: vector P2 on a crossed mesh, Voigt operators, degree-10
    coefficient weights, two facet tractions; linear vs Newton amplitudes agree to 1e-8 (:748-752)."""
    kw = dict(N=(20, 4), numElems=(2, 10, 10), PGD_nmax=3)
    p = ref_cases.elasticity2d(**kw)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.elasticity2d(spaces=_ospaces(p), **kw)
    opgd.solve_pgd(o)
    assert p.PGD_modes == o.PGD_modes == 3
    # the first mode stops on the round-off floor of the "norm" test (floor 3e-4 > tol_fp_it 1e-4): a sweep more or
    # less there shifts the following modes at the 1e-7 level, so a run whose counts differ is repeated with the
    # oracle's sweep schedule pinned and held to the same 1e-8

    def pinned(schedule):
        r = ref_cases.elasticity2d(**kw)
        r.fp_schedule = schedule
        return r.solve_PGD(_problem="linear")

    p = _compare(p, o, floor_ok=True, rerun=pinned)
    q = ref_cases.elasticity2d(**kw)
    q.fp_schedule = list(p.num_fp_it)  # Newton path at the linear path's sweep counts (test_solver_problem.py:748-752)
    q.solve_PGD()
    assert q.num_fp_it == p.num_fp_it
    assert np.allclose(p.amplitude, q.amplitude, rtol=1e-8, atol=0)


CASES = [("truss", case_truss, ()), ("heat1d_fem_heating", case_heat1d, ("FEM", "heating")),
         ("heat1d_fem_cooling", case_heat1d, ("FEM", "cooling")), ("heat1d_fd_heating", case_heat1d, ("FD", "heating")),
         ("heat1d_fd_cooling", case_heat1d, ("FD", "cooling")), ("laplace_fem", case_laplace_fem, ()),
         ("elasticity2d", case_elasticity2d, ())]


@pytest.mark.gpu
@pytest.mark.parametrize("name,fn,args", CASES, ids=[c[0] for c in CASES])
def test_reference_case_on_device(name, fn, args):
    fn(*args)


@pytest.mark.parametrize("name,fn,args", CASES, ids=[c[0] for c in CASES])
def test_reference_case_host_logic(cpu, name, fn, args):
    fn(*args)
