"""-m "not gpu": the oracle (and the host-side FD_matrices) against golden vectors produced by the
UNMODIFIED reference code (tests/golden/make_golden.py ran pgdrome/solver.py and pgdrome/model.py
in the development container; see its header for what each fixture is).

This is what pins the oracle: FD_matrices, the whole enrichment loop in FD mode (get_Fsinit,
residual check, FP_solve with both stopping criteria, all three normalisations, stopping test),
PGD.evaluate (interp1d path and mode point-evaluation path), evaluate_min/max, LHS sampling and
the error loop agree with the reference to round-off.  The FEM assembly underneath DOLFIN stays
unpinned (no FEniCS here) -- see DESIGN.md "Parity status".
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import evaluate as oev
from oracle import fem as ofem
from oracle import pgd as opgd
from oracle import problems as oprob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _mode_err(a, b):
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)


# ------------------------------------------------------------------------------ FD_matrices
@pytest.mark.parametrize("grid", ["uniform200", "graded37", "three", "uniform11"])
def test_fd_matrices_oracle_and_host(grid):
    """pgdrome/solver.py:947-988 incl. the half-weight first row and the stale-hp last row."""
    g = _gold("fd_matrices")
    x = g["x_" + grid]
    from pgdrome_b200.solver import FD_matrices as host_fd

    for impl in (opgd.FD_matrices, host_fd):
        M, D2, D1 = impl(x)
        for got, key in ((M, "M_"), (D2, "D2_"), (D1, "D1_")):
            ref = g[key + grid]
            got = sp.csr_matrix(got).toarray()
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 4e-16 * np.abs(ref).max(), (impl.__module__, key)
    assert sp.issparse(host_fd(x)[0]) and host_fd(x)[0].format == "lil"  # same container as the reference


# ------------------------------------------------------------------------------ enrichment loop (FD mode)
def _check_against(g, key, p, spaces, tol=1e-9):
    assert p.PGD_modes == int(g[key + "_n_modes"])
    # Fixed-point iteration counts must match, except where the "norm" criterion has hit its
    # round-off floor: sqrt|nn + oo - 2 no| cancels ~|alpha|^2-sized products, so once converged it
    # evaluates to either 0.0 or ~sqrt(eps)*alpha (which can exceed tol_fp_it = 1e-5, here
    # alpha = 2.6e3 => 3.9e-5) depending on the last bit of the operands.  The reference test
    # verbatim is such a case: golden err_fp_it == 0.0 after 3 iterations, the oracle reaches 0.0 after 2.
    # p.fp_floor[n] = sqrt(eps (nn + oo + 2|no|)) is that floor; stop_fp = "delta" has no cancellation
    # and must match exactly.
    for n, (a, b) in enumerate(zip(p.num_fp_it, g[key + "_num_fp_it"])):
        at_floor = p.stop_fp == "norm" and min(float(p.err_fp_it[n]), float(g[key + "_err_fp_it"][n])) < 4 * p.fp_floor[n]
        assert a == b or at_floor, (key, n, a, b)
    # a different number of fixed-point sweeps at the floor leaves O(tol_fp_it^2)-level differences
    rt = 1e-12 if (p.stop_fp == "delta" or list(p.num_fp_it) == list(g[key + "_num_fp_it"])) else 1e-7
    assert np.allclose(p.amplitude, g[key + "_amplitude"], rtol=rt, atol=0)
    assert np.allclose(p.alpha, g[key + "_alpha"], rtol=rt, atol=0)
    if p.stop_fp == "delta":
        assert np.allclose([np.max(e) for e in p.err_fp_it], g[key + "_err_fp_it"], rtol=1e-5, atol=0)
    for d in range(4):
        assert np.array_equal(spaces[d].dof_coordinates().ravel(), g[key + "_dofx%d" % d])  # same dof numbering
        for k in range(p.PGD_modes):
            assert _mode_err(p.PGD_func[d][k], g[key + "_modes%d" % d][k]) < tol, (key, d, k)


def test_laplace_fd_reference_test_verbatim():
    """tests/integration/test_laplace.py create_PGD(_type="FD") on [60,40,200,80]: 1 mode."""
    g = _gold("laplace_fd")
    p, info = oprob.laplace_xyqu(kind="FD")
    opgd.solve_pgd(p)
    assert p.PGD_modes == 1 == int(g["ref_numModes"])  # pinned by test_laplace.py:970-971
    _check_against(g, "ref", p, info["spaces"])
    # PGD.evaluate, mode point-evaluation path (model.py:805-860)
    S = info["spaces"]
    scale = max(np.linalg.norm(r) for r in g["ref_eval"])  # one point has q = 0 => u ~ round-off
    for pt, ref in zip(g["ref_eval_points"], g["ref_eval"]):
        u = oev.evaluate_dofs(p.PGD_func[0], S[1:], p.PGD_func[1:], pt)
        assert np.linalg.norm(u - ref) <= 1e-9 * scale
    # PGDErrorComputation (model.py:1704-1825): LHS samples over the mesh ranges + error loop
    smp = oev.sampling_LHS([0.0, 0.0, 10.0], [3.0, 50.0, 50.0], 7)
    assert np.allclose(smp, g["ref_err_samples"], rtol=1e-14, atol=0)
    xv = np.sort(S[0].dof_coordinates().ravel())
    v2d = S[0].vertex_to_node
    fom = lambda s: (1.0 + 0.01 * s[1]) * np.sin(xv) * s[2] / 10.0 + 0.1 * s[0]
    pgd_eval = lambda s: oev.evaluate_dofs(p.PGD_func[0], S[1:], p.PGD_func[1:], s)[v2d]  # vertex order
    err, mean, mx = oev.evaluate_error(fom, pgd_eval, smp)
    assert np.allclose(err, g["ref_err"], rtol=1e-8, atol=0)
    assert np.allclose([mean, mx], g["ref_err_mean_max"], rtol=1e-8, atol=0)


@pytest.mark.parametrize("key,opts", [
    ("v_stiff", dict(norm_modes="stiff", PGD_nmax=6)),
    ("v_l2", dict(norm_modes="l2", PGD_nmax=6)),
    ("v_no", dict(norm_modes="no", PGD_nmax=4)),
    ("v_delta", dict(stop_fp="delta", PGD_nmax=4, tol_fp_it=1e-6)),
])
def test_laplace_fd_variants(key, opts):
    """Several modes, all normalisations, both fixed-point stopping criteria: reference
    PGDProblem + the reference test's FD callbacks vs the oracle's separated restatement."""
    g = _gold("laplace_fd")
    Qv = [g[key + "_" + n] for n in ("qx", "qy", "qq", "qu0")]
    p, info = oprob.laplace_xyqu(kind="FD", elems=tuple(int(e) for e in g[key + "_elem"]), Qv=Qv, **opts)
    opgd.solve_pgd(p)
    assert p.PGD_modes > 1
    # later modes inherit the fixed-point tolerance of the earlier ones: compare the first ones
    # tightly, all of them at the level the 1e-5 fixed-point tolerance allows
    _check_against(g, key, p, info["spaces"], tol=1e-11 if key == "v_delta" else 1e-7)


# ------------------------------------------------------------------------------ PGD.evaluate (interp1d path)
def test_pgdclass_evaluate_interp1d():
    g = _gold("pgdclass")
    for at in (0, 1):
        fixed = list(g["data_0_%d" % at])
        free_x = [g["x1"], g["x2"]]
        free = [list(g["data_1_%d" % at][:, :, 0]), list(g["data_2_%d" % at][:, :, 0])]
        for pt, ref in zip(g["points"], g["eval_%d" % at]):
            u = oev.evaluate_interp1d(fixed, free_x, free, pt)
            assert u.shape == ref.shape
            assert np.abs(u - ref).max() <= 1e-15 + 1e-14 * np.abs(ref).max()
            if at == 0:
                k = list(map(tuple, g["points"])).index(tuple(pt))
                assert np.isclose(u.min(), g["eval_min"][k], rtol=1e-13, atol=1e-16)
                assert np.isclose(u.max(), g["eval_max"][k], rtol=1e-13, atol=1e-16)
    assert int(g["out_of_range_raises"]) == 1
    with pytest.raises(ValueError):  # tests/unit/test_pgdclass.py:319-326
        oev.evaluate_interp1d(list(g["data_0_0"]), [g["x1"], g["x2"]],
                              [list(g["data_1_0"][:, :, 0]), list(g["data_2_0"][:, :, 0])], [0.2, 0.4])
    # analytic truss solution of the reference test: 5 decimals (test_pgdclass.py:298-317)
    E, L = 0.5, 0.4
    u = oev.evaluate_interp1d(list(g["data_0_0"]), [g["x1"], g["x2"]],
                              [list(g["data_1_0"][:, :, 0]), list(g["data_2_0"][:, :, 0])], [E, L]).ravel()
    np.testing.assert_almost_equal(u, 0.5 / E * (g["x0"] - g["x0"] ** 2) * L, 5)


def test_lhs_sampling_seed():
    g = _gold("pgdclass")
    lo, hi = g["lhs_bounds"]
    assert np.allclose(oev.sampling_LHS(list(lo), list(hi), 10), g["lhs_samples"], rtol=1e-15, atol=0)
