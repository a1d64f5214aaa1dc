"""Host logic on the CPU (-m "not gpu"): the user-callback path of PGDProblem / PGD.evaluate with
the C-ABI wrappers replaced by the NumPy stand-in of tests/cpu_abi.py, checked against the oracle.
The kernels themselves are checked by the -m gpu tests."""
import numpy as np
import pytest

from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob
from tests import cpu_abi


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


def _compare(p, o, tol=1e-8):
    assert p.PGD_modes == o.PGD_modes
    assert p.num_fp_it == o.num_fp_it
    for d in range(len(p.V)):
        for k in range(p.PGD_modes):
            assert _mode_err(p.PGD_func[d][k].vector()[:], o.PGD_func[d][k]) < tol, (d, k)
    assert np.allclose(p.amplitude, o.amplitude, rtol=1e-8, atol=0)


def test_poisson1d_k_matches_oracle(cpu):
    from pgdrome_b200 import configs

    p = configs.poisson1d_k(nx=60, nk=12, PGD_nmax=4)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.poisson1d_k(nx=60, nk=12, PGD_nmax=4, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
    # Newton path == linear path (tests/integration/test_solver_problem.py:748-752)
    q = configs.poisson1d_k(nx=60, nk=12, PGD_nmax=4)
    q.solve_PGD()
    assert np.allclose(q.amplitude, p.amplitude, rtol=1e-10, atol=0)


def test_heat2d_tk_matches_oracle(cpu):
    from pgdrome_b200 import configs

    p = configs.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3)
    p.solve_PGD(_problem="linear")
    o, _ = oprob.heat2d_tk(n=8, nt=12, nk=5, PGD_nmax=3, spaces=_ospaces(p))
    opgd.solve_pgd(o)
    _compare(p, o)
