"""-m gpu: the PRODUCT path (PGDProblem.solve_PGD in FD mode -> device banded LU, device mass
products; PGD.evaluate interp1d path and batched evaluate -> weight kernel + GEMV / DMMA GEMM)
against golden vectors produced by the unmodified reference (tests/golden/make_golden.py).

The FD callbacks below restate tests/integration/test_laplace.py:372-767 (problem_assemble_lhs_FD /
problem_assemble_rhs_FD) in loop form; they are user code, not part of the package."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _mode_err(a, b):
    return min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b)


def _laplace_fd_problem(elem, Qv=None, **attrs):
    from pgdrome_b200 import dolfin as df
    from pgdrome_b200.solver import FD_matrices, PGDProblem

    ranges = [[0.0, 3.0], [0.0, 3.0], [0.0, 50.0], [10.0, 50.0]]
    vs = [df.FunctionSpace(df.IntervalMesh(int(elem[i]), ranges[i][0], ranges[i][1]), "CG", 1) for i in range(4)]
    k, lx = 0.5, 3.0
    xd = [np.array(v.tabulate_dof_coordinates()[:].flatten()) for v in vs]
    srt = [np.argsort(x) for x in xd]
    M, D2 = [], []
    for i in range(4):
        m, d2, _ = FD_matrices(xd[i][srt[i]])
        M.append(m[srt[i], :][:, srt[i]])
        D2.append(d2[srt[i], :][:, srt[i]])
    bc_idx = np.array([np.where(xd[0] == 0)[0], np.where(xd[0] == lx)[0]]).flatten()

    def fn(V, a):
        f = df.Function(V)
        f.vector()[:] = a
        return f

    BC = [fn(vs[0], 1.0 - xd[0] / 3.0), fn(vs[1], np.ones_like(xd[1])), fn(vs[2], np.ones_like(xd[2])), fn(vs[3], xd[3])]
    if Qv is None:
        Qv = [np.where(xd[0] < lx / 2, 1.0, 0.0), np.ones_like(xd[1]), xd[2], np.ones_like(xd[3])]
    Q = [[fn(vs[i], Qv[i])] for i in range(4)]
    # the two separated operators: (D2_x, M_y, M_q, M_u) and (M_x, D2_y, M_q, M_u)
    ops = [[D2[0], M[1], M[2], M[3]], [M[0], D2[1], M[2], M[3]]]
    which = {"r": 0, "s": 1, "t": 2, "u": 3}

    def v_(f):
        return f.vector()[:]

    def lhs_fct(fct_F, var_F, Fs, meshes, dom, param, typ, dim):
        d = which[typ]
        a = 0
        for op in ops:
            c = 1.0
            for j in range(4):
                if j != d:
                    c = c * (v_(Fs[j]).transpose() @ op[j] @ v_(Fs[j]))
            a = a - c * k * op[d]
        if d == 0:
            a = a.tolil()
            a[:, bc_idx] = 0.0
            a[bc_idx, :] = 0.0
            a[bc_idx, bc_idx] = 1.0
        return a

    def rhs_fct(fct_F, var_F, Fs, meshes, dom, param, Qs, PGD_func, typ, nE, dim):
        d = which[typ]
        c = 1.0
        for j in range(4):
            if j != d:
                c = c * (v_(Fs[j]).transpose() @ M[j] @ v_(Qs[j][0]))
        l = c * (M[d] @ v_(Qs[d][0]))
        for other in [BC] + [[PGD_func[j][old] for j in range(4)] for old in range(nE)]:
            for op in ops:
                c = 1.0
                for j in range(4):
                    if j != d:
                        c = c * (v_(Fs[j]).transpose() @ op[j] @ v_(other[j]))
                l = l + c * k * (op[d] @ v_(other[d]))
        if d == 0:
            l[bc_idx] = 0
        return l

    def bc_fct(Vs, dom, param):
        def leftright(x, on_boundary):
            return on_boundary and df.near(x[0], 0.0, 1e-6) or df.near(x[0], lx, 1e-6)

        return [df.DirichletBC(Vs[0], 0, leftright), 0, 0, 0]

    p = PGDProblem(name="test_x_y_q_u00", name_coord=["X", "Y", "q", "u0"], modes_info=["T", "Node", "Scalar"], Vs=vs,
                   dom=0, bc_fct=bc_fct, load=Q, param={}, rhs_fct=rhs_fct, lhs_fct=lhs_fct, probs=["r", "s", "t", "u"],
                   seq_fp=np.arange(4), PGD_nmax=7)
    p.MM = M
    p.stop_fp, p.max_fp_it, p.tol_fp_it, p.norm_modes = "norm", 50, 1e-5, "stiff"
    for a, v in attrs.items():
        setattr(p, a, v)
    return p, xd


def _check(g, key, p, xd, tol):
    assert p.PGD_modes == int(g[key + "_n_modes"])
    exact_counts = list(p.num_fp_it) == list(g[key + "_num_fp_it"])
    if p.stop_fp == "delta":
        assert exact_counts
    else:  # see tests/test_oracle_golden.py: the "norm" test sits on its round-off floor once converged
        assert max(abs(a - b) for a, b in zip(p.num_fp_it, g[key + "_num_fp_it"])) <= 8
    rt = 1e-10 if exact_counts else 1e-7
    assert np.allclose(p.amplitude, g[key + "_amplitude"], rtol=rt, atol=0)
    assert np.allclose(p.alpha, g[key + "_alpha"], rtol=rt, atol=0)
    for d in range(4):
        assert np.array_equal(xd[d], g[key + "_dofx%d" % d])  # DOLFIN's 1-D dof numbering
        for k in range(p.PGD_modes):
            assert _mode_err(p.PGD_func[d][k].vector()[:], g[key + "_modes%d" % d][k]) < tol, (key, d, k)


def test_laplace_fd_reference_test_on_device():
    g = _gold("laplace_fd")
    p, xd = _laplace_fd_problem([60, 40, 200, 80])
    p.solve_PGD(_problem="linear", solve_modes=["FD"] * 4)
    pgd = p.return_PGD()
    assert pgd.numModes == 1  # test_laplace.py:970-971
    _check(g, "ref", p, xd, 1e-8)
    assert p.solver_stats["banded_solves"] >= 8
    scale = max(np.linalg.norm(r) for r in g["ref_eval"])
    for pt, ref in zip(g["ref_eval_points"], g["ref_eval"]):
        u = pgd.evaluate(0, [1, 2, 3], list(pt), 0).vector()[:]
        assert np.linalg.norm(u - ref) <= 1e-8 * scale
    U = pgd.evaluate_batch(0, [1, 2, 3], g["ref_eval_points"], 0).cpu().numpy()
    assert np.linalg.norm(U - g["ref_eval"]) <= 1e-8 * scale
    # PGDErrorComputation (model.py:1666-1825) through the product classes: LHS samples over the mesh ranges
    # (seed 3452) and the relative-L2 error loop against the reference's own numbers for the same synthetic FOM
    from pgdrome_b200.model import PGDErrorComputation

    xv = p.meshes[0].coordinates()[:, 0]
    fom = lambda s: (1.0 + 0.01 * s[1]) * np.sin(xv) * s[2] / 10.0 + 0.1 * s[0]
    ec = PGDErrorComputation(fixed_dim=[0], n_samples=7, FOM_model=fom, PGD_model=pgd)
    err, mean_err, max_err = ec.evaluate_error()
    assert np.allclose(ec.data_test, g["ref_err_samples"], rtol=1e-14, atol=0)
    assert np.allclose(err, g["ref_err"], rtol=1e-7, atol=0)
    assert np.allclose([mean_err, max_err], g["ref_err_mean_max"], rtol=1e-7, atol=0)


@pytest.mark.parametrize("key,opts", [
    ("v_stiff", dict(norm_modes="stiff", PGD_nmax=6)),
    ("v_l2", dict(norm_modes="l2", PGD_nmax=6)),
    ("v_no", dict(norm_modes="no", PGD_nmax=4)),
    ("v_delta", dict(stop_fp="delta", PGD_nmax=4, tol_fp_it=1e-6)),
])
def test_laplace_fd_variants_on_device(key, opts):
    g = _gold("laplace_fd")
    Qv = [g[key + "_" + n] for n in ("qx", "qy", "qq", "qu0")]
    p, xd = _laplace_fd_problem(g[key + "_elem"], Qv=Qv, **opts)
    p.solve_PGD(_problem="linear", solve_modes=["FD"] * 4)
    _check(g, key, p, xd, 1e-8 if key == "v_delta" else 1e-7)


def test_pgdclass_evaluate_on_device():
    """tests/unit/test_pgdclass.py: NumPy-built PGD object, interp1d evaluation path."""
    from pgdrome_b200.model import PGD, PGDAttribute, PGDMesh

    g = _gold("pgdclass")
    pgd = PGD()
    pgd.name, pgd.numModes, pgd.used_numModes = "PGDsolution", 1, 1
    grids = []
    for d, nm in enumerate(["PGD1", "PGD2", "PGD3"]):
        m = PGDMesh(nm)
        m.dataX = g["x%d" % d]
        m.numNodes = len(m.dataX)
        m.numElements = m.numNodes - 1
        m.dataY, m.dataZ = np.zeros(m.numNodes), np.zeros(m.numNodes)
        m.typElements = "Polyline"
        m.topology = [[i, i + 1] for i in range(m.numElements)]
        attrs = []
        for at, an in enumerate(["U_x", "Sig_x"]):
            a = PGDAttribute()
            a.name, a._type, a.field = an, "Node", "Scalar"
            a.data = [np.array(x) for x in g["data_%d_%d" % (d, at)]]
            attrs.append(a)
        m.attributes = attrs
        grids.append(m)
    pgd.mesh = grids
    for at in (0, 1):
        for d in (1, 2):
            pgd.mesh[d].attributes[at].interpolationInfo = {"name": 0, "kind": "linear"}
        pgd.create_interpolation_fcts([1, 2], at)
        for pt, ref in zip(g["points"], g["eval_%d" % at]):
            u = np.asarray(pgd.evaluate(0, [1, 2], list(pt), at))
            assert u.shape == ref.shape
            assert np.abs(u - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
    for k, pt in enumerate(g["points"]):
        assert abs(pgd.evaluate_min(0, [1, 2], list(pt), 0) - g["eval_min"][k]) < 1e-13
        assert abs(pgd.evaluate_max(0, [1, 2], list(pt), 0) - g["eval_max"][k]) < 1e-13
    with pytest.raises(ValueError):
        pgd.evaluate_min(0, [1, 2], [0.2, 0.4], 0)
