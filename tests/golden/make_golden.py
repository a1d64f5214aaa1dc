#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run in the development container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

The reference's modules import DOLFIN at module scope and FEniCS is not installable here, but its
finite-difference paths are NumPy/SciPy only.  `_dolfin_stub.py` (a container-only stand-in, see
its header) is registered as `dolfin`, an empty module as `h5py`, and then the UNMODIFIED
reference code is imported and executed:

  fd_matrices.npz      pgdrome.solver.FD_matrices (solver.py:947-988) on uniform / graded /
                       two-point grids
  laplace_fd.npz       tests/integration/test_laplace.py create_PGD(_type="FD") verbatim: the
                       reference PGDProblem.solve_PGD / get_Fsinit / FP_solve / FD_solve
                       (solver.py:158-943) with the test's own FD callbacks (:372-767) on the
                       test's meshes [60,40,200,80]; plus variants (norm_modes "l2"/"no",
                       stop_fp "delta", a non-separable source => several modes)
  pgdclass.npz         tests/unit/test_pgdclass.py create_example_pgd_solution + the reference
                       PGD.evaluate interp1d path (model.py:780-803), evaluate_min/max
                       (:955-1010), PGDErrorComputation.sampling_LHS (:1704-1743, seed 3452)
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _install_stub():
    spec = importlib.util.spec_from_file_location("dolfin", os.path.join(HERE, "_dolfin_stub.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    sys.modules["dolfin"] = stub
    sys.modules["fenics"] = stub
    sys.modules["h5py"] = types.ModuleType("h5py")
    sys.path.insert(0, REF)
    return stub


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def fd_matrices(out):
    from pgdrome.solver import FD_matrices

    rng = np.random.default_rng(7)
    grids = {"uniform200": np.linspace(0.0, 50.0, 201), "graded37": np.cumsum(np.concatenate([[0.0], 0.05 + rng.random(36)])),
             "three": np.array([0.0, 0.4, 1.0]), "uniform11": np.linspace(0.5, 1.0, 11)}
    for k, x in grids.items():
        M, D2, D1 = FD_matrices(x)
        out["x_" + k] = x
        out["M_" + k] = M.toarray()
        out["D2_" + k] = D2.toarray()
        out["D1_" + k] = D1.toarray()


def _dump_problem(out, key, prob):
    D = len(prob.PGD_func)
    out[key + "_n_modes"] = np.array(prob.PGD_modes)
    out[key + "_num_fp_it"] = np.array(prob.num_fp_it, dtype=np.int64)
    out[key + "_err_fp_it"] = np.array([np.max(e) for e in prob.err_fp_it], dtype=np.float64)
    out[key + "_alpha"] = np.array(prob.alpha, dtype=np.float64)
    out[key + "_amplitude"] = np.array(prob.amplitude, dtype=np.float64)
    for d in range(D):
        out[key + "_modes%d" % d] = np.array([f.vector()[:].copy() for f in prob.PGD_func[d]])
        out[key + "_dofx%d" % d] = prob.V[d].tabulate_dof_coordinates()[:, 0].copy()


def laplace_fd(out):
    tl = _load(os.path.join(REF, "tests/integration/test_laplace.py"), "ref_test_laplace")
    ranges = [[0.0, 3.0], [0.0, 3.0], [0.0, 50.0], [10.0, 50.0]]
    elem = [60, 40, 200, 80]
    # (1) the reference test verbatim
    meshes, vs = tl.create_meshes(elem, [1, 1, 1, 1], ranges)
    pgd_s, param = tl.create_PGD(param={"k": 0.5, "lx": 3, "ly": 3}, vs=vs, _type="FD")
    prob = pgd_s.problem
    _dump_problem(out, "ref", prob)
    out["ref_numModes"] = np.array(pgd_s.numModes)
    # reference PGD.evaluate (DOLFIN path: Function.__call__ point evaluation of each free-dim mode)
    vals = [[1.5, 50.0, 10.0], [0.3, 12.5, 33.0], [2.9, 0.0, 49.5]]
    ev = []
    for v in vals:
        ev.append(pgd_s.evaluate(0, [1, 2, 3], v, 0).vector()[:].copy())
    out["ref_eval_points"] = np.array(vals)
    out["ref_eval"] = np.array(ev)
    # reference PGDErrorComputation (model.py:1666-1825): LHS sampling over the free-dim mesh ranges
    # and the relative L2 error loop, against a synthetic full-order model (vertex order)
    from pgdrome.model import PGDErrorComputation

    xv = meshes[0].coordinates()[:, 0]
    fom = lambda s: (1.0 + 0.01 * s[1]) * np.sin(xv) * s[2] / 10.0 + 0.1 * s[0]
    ec = PGDErrorComputation(fixed_dim=[0], n_samples=7, FOM_model=fom, PGD_model=pgd_s)
    err, mean_err, max_err = ec.evaluate_error()
    out["ref_err_samples"] = np.array(ec.data_test)
    out["ref_err"] = np.array(err)
    out["ref_err_mean_max"] = np.array([mean_err, max_err])

    # (2) variants through the same reference classes / callbacks (smaller meshes, richer sources)
    import dolfin as df
    from pgdrome.solver import FD_matrices, PGDProblem

    def variant(key, elem, src_x, src_q, norm_modes="stiff", stop_fp="norm", nmax=6, tol=1e-5, fp_init=""):
        meshes, vs = tl.create_meshes(elem, [1, 1, 1, 1], ranges)
        p = {"k": 0.5, "lx": 3, "ly": 3}
        p["BC_x"] = df.interpolate(df.Expression("1.0-1.0/3.0*x[0]", degree=1), vs[0])
        p["BC_y"] = df.interpolate(df.Expression("1.0", degree=1), vs[1])
        p["BC_q"] = df.interpolate(df.Expression("1.0", degree=1), vs[2])
        p["BC_u0"] = df.interpolate(df.Expression("x[0]", degree=1), vs[3])
        qx = [df.interpolate(df.Expression(src_x, degree=1, L=3.0), vs[0])]
        qy = [df.interpolate(df.Expression("1.0 + 0.5*x[0]*x[0]", degree=1), vs[1])]
        qq = [df.interpolate(df.Expression(src_q, degree=1), vs[2])]
        qu0 = [df.interpolate(df.Expression("1.0", degree=1), vs[3])]
        x_dofs, idx = tl.get_coordinates_and_sorts(vs)
        mats = [FD_matrices(x_dofs[i][idx[i]]) for i in range(4)]
        rs = lambda A, i: A[idx[i], :][:, idx[i]]
        p["M_x"], p["D2_x"] = rs(mats[0][0], 0), rs(mats[0][1], 0)
        p["M_y"], p["D2_y"] = rs(mats[1][0], 1), rs(mats[1][1], 1)
        p["M_q"], p["M_u"] = rs(mats[2][0], 2), rs(mats[3][0], 3)
        p["bc_idx"] = np.array([np.where(x_dofs[0] == 0)[0], np.where(x_dofs[0] == p["lx"])[0]]).flatten()
        prob = PGDProblem(name=key, name_coord=["X", "Y", "q", "u0"], modes_info=["T", "Node", "Scalar"], Vs=vs, dom=0,
                          bc_fct=tl.create_bc, load=[qx, qy, qq, qu0], param=p, rhs_fct=tl.problem_assemble_rhs_FD,
                          lhs_fct=tl.problem_assemble_lhs_FD, probs=["r", "s", "t", "u"], seq_fp=np.arange(4), PGD_nmax=nmax)
        prob.MM = [p["M_x"], p["M_y"], p["M_q"], p["M_u"]]
        prob.stop_fp, prob.max_fp_it, prob.tol_fp_it, prob.norm_modes, prob.fp_init = stop_fp, 50, tol, norm_modes, fp_init
        prob.solve_PGD(_problem="linear", solve_modes=["FD"] * 4)
        _dump_problem(out, key, prob)
        out[key + "_elem"] = np.array(elem)
        for nm, q in (("qx", qx), ("qy", qy), ("qq", qq), ("qu0", qu0)):
            out[key + "_" + nm] = q[0].vector()[:].copy()
        print(key, "modes", prob.PGD_modes, "fp", prob.num_fp_it, "amp", ["%.2e" % a for a in prob.amplitude])

    # a y-dependent source makes the (lifted) problem non-separable in (x, y) => several modes
    variant("v_stiff", [24, 16, 10, 8], "x[0]<L/2 ? 1.0 : 0", "x[0]")
    variant("v_l2", [24, 16, 10, 8], "x[0]<L/2 ? 1.0 : 0", "x[0]", norm_modes="l2")
    variant("v_no", [24, 16, 10, 8], "exp(-x[0])", "1.0+x[0]", norm_modes="no", nmax=4)
    variant("v_delta", [24, 16, 10, 8], "x[0]<L/2 ? 1.0 : 0", "x[0]", stop_fp="delta", nmax=4, tol=1e-6)


def pgdclass(out):
    tp = _load(os.path.join(REF, "tests/unit/test_pgdclass.py"), "ref_test_pgdclass")
    from pgdrome.model import PGDErrorComputation

    param = {"A": 1, "n": 1, "lae": 1}
    pgd = tp.create_example_pgd_solution(param)
    for d in range(3):
        out["x%d" % d] = np.asarray(pgd.mesh[d].dataX, dtype=np.float64)
        for at in (0, 1):
            out["data_%d_%d" % (d, at)] = np.array([np.asarray(m, dtype=np.float64) for m in pgd.mesh[d].attributes[at].data])
    pts = [[0.5, 0.4], [1.0, 0.0], [0.737, 0.93], [0.9999, 1.0], [0.5, 0.999], [0.6234, 0.05]]
    out["points"] = np.array(pts)
    for at in (0, 1):
        for d in (1, 2):
            pgd.mesh[d].attributes[at].interpolationInfo = {"name": 0, "kind": "linear"}
        pgd.create_interpolation_fcts([1, 2], at)
        out["eval_%d" % at] = np.array([np.asarray(pgd.evaluate(0, [1, 2], p, at)).copy() for p in pts])
    out["eval_min"] = np.array([pgd.evaluate_min(0, [1, 2], p, 0) for p in pts])
    out["eval_max"] = np.array([pgd.evaluate_max(0, [1, 2], p, 0) for p in pts])
    try:
        pgd.evaluate_min(0, [1, 2], [0.2, 0.4], 0)
        out["out_of_range_raises"] = np.array(0)
    except ValueError:
        out["out_of_range_raises"] = np.array(1)
    # the reference's Latin-hypercube sampler (model.py:1704-1743)
    lo, hi = [-1.0, 0.2], [3.0, 2.0]
    ec = PGDErrorComputation(fixed_dim=[0], n_samples=10, FOM_model=None, PGD_model=pgd,
                             lim_samples=[None, [lo[0], hi[0]], [lo[1], hi[1]]])
    out["lhs_bounds"] = np.array([lo, hi])
    out["lhs_samples"] = np.array(ec.sampling_LHS())


def pxdmf(out):
    """PXDMF round trip across the two implementations: the PRODUCT writer (pgdrome_b200 PGD.write_pxdmf,
    Format="XML") produces tests/golden/pxdmf/PGDsolution.pxdmf; the UNMODIFIED reference loader
    (model.py:406-575) reads it back, and what it stored (and what the reference then evaluates from it)
    is the golden data the product loader / device evaluate are compared with."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from pgdrome_b200 import model as pm

    from pgdrome.model import PGD as RefPGD

    rng = np.random.default_rng(11)
    R = 3
    nx, ny = 4, 3
    X, Y = np.meshgrid(np.linspace(0.0, 2.0, nx + 1), np.linspace(0.0, 1.0, ny + 1))
    tri = []
    for j in range(ny):
        for i in range(nx):
            v0 = j * (nx + 1) + i
            v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
            tri += [[v0, v1, v3], [v0, v2, v3]]
    x1 = np.array([0.0, 2.0, 0.35, 0.8, 1.1, 1.45, 1.9])  # deliberately not sorted
    x2 = np.linspace(1.0, 3.0, 5)
    grids = [("PGD1", 2, "Triangle", np.array(tri), X.ravel(), Y.ravel(), "X"),
             ("PGD2", 1, "Polyline", np.column_stack([np.argsort(x1)[:-1], np.argsort(x1)[1:]]), x1, 0 * x1, "E"),
             ("PGD3", 1, "Polyline", np.column_stack([np.arange(4), np.arange(1, 5)]), x2, 0 * x2, "F")]
    pgd = pm.PGD()
    pgd.name, pgd.numModes, pgd.used_numModes = "PGDsolution", R, R
    for name, dim, typ, topo, gx, gy, cname in grids:
        m = pm.PGDMesh(name)
        m.meshdim, m.info = dim, [dim, cname, "-?-"]
        m.topology, m.numElements, m.typElements = topo, len(topo), typ
        m.dataX, m.dataY, m.dataZ, m.numNodes = gx, gy, np.zeros(len(gx)), len(gx)
        m.attributes = []
        for an in ("U_x", "Sig_x"):
            a = pm.PGDAttribute()
            a.name, a._type, a.field = an, "Node", "Scalar"
            a.data = [rng.standard_normal((len(gx), 1)) * 10.0 ** rng.integers(-3, 3) for _ in range(R)]
            m.attributes.append(a)
        pgd.mesh.append(m)
    folder = os.path.join(HERE, "pxdmf")
    path = pgd.write_pxdmf(folder)
    ref = RefPGD().load_pxdmf(path)
    out["name"] = np.array(ref.name)
    out["num_modes"] = np.array(ref.numModes)
    for d, m in enumerate(ref.mesh):
        out["g%d_name" % d] = np.array(m.name)
        out["g%d_info" % d] = np.array(m.info)
        out["g%d_meshdim" % d] = np.array(m.meshdim)
        out["g%d_num_elements" % d] = np.array(m.numElements)
        out["g%d_typ_elements" % d] = np.array(m.typElements)
        out["g%d_topology" % d] = np.asarray(m.topology)
        out["g%d_num_nodes" % d] = np.array(m.numNodes)
        out["g%d_x" % d] = np.asarray(m.dataX)
        out["g%d_y" % d] = np.asarray(m.dataY)
        for a, att in enumerate(m.attributes):
            out["g%d_a%d_meta" % (d, a)] = np.array([att.name, att._type, att.field])
            out["g%d_a%d_data" % (d, a)] = np.array(att.data)
        # the reference read back exactly what the product wrote (17 significant digits)
        assert np.array_equal(np.asarray(m.dataX), pgd.mesh[d].dataX)
        for a, att in enumerate(m.attributes):
            assert all(np.array_equal(att.data[k], pgd.mesh[d].attributes[a].data[k]) for k in range(R))
    pts = [[0.0, 1.0], [2.0, 3.0], [0.35, 1.5], [1.0, 2.25], [1.777, 2.999], [0.1234, 1.001]]
    out["points"] = np.array(pts)
    for at in (0, 1):
        for d in (1, 2):
            ref.mesh[d].attributes[at].interpolationInfo = {"name": 0, "kind": "linear"}
        ref.create_interpolation_fcts([1, 2], at)
        out["eval_%d" % at] = np.array([np.asarray(ref.evaluate(0, [1, 2], p, at)).copy() for p in pts])


def sensor(out):
    """evaluate_sensor_response (model.py:862-953) of the UNMODIFIED reference with the Probes result injected into
    its own cache (model.py:115-117; fenicstools is absent): pins how the probed fixed modes are combined with the
    free-dimension factors and which shapes come back (scalar / vector field, one mode / several)."""
    from pgdrome.model import PGD as RefPGD

    rng = np.random.default_rng(5)
    ref = RefPGD().load_pxdmf(os.path.join(HERE, "pxdmf", "PGDsolution.pxdmf"))  # 3 modes, free dims 1-D
    for d in (1, 2):
        ref.mesh[d].attributes[0].interpolationInfo = {"name": 0, "kind": "linear"}
    ref.create_interpolation_fcts([1, 2], 0)
    pts = rng.random((5, 2))
    coords = [[0.35, 1.5], [1.777, 2.999], [0.0, 1.0]]
    out["points"], out["coords"] = pts, np.array(coords)
    key = (np.sum(pts.flatten()), 0, 0)
    for tag, E in (("scalar", rng.standard_normal((5, 3))), ("vector", rng.standard_normal((5, 2, 3)))):
        ref._eval_fixed_modes = {key: E}
        out["E_" + tag] = E
        out["resp_" + tag] = np.array([ref.evaluate_sensor_response(0, [1, 2], c, 0, pts) for c in coords])
        ref.used_numModes = 2  # truncated evaluation (model.py:911-921 loops over used_numModes)
        out["resp2_" + tag] = np.array([ref.evaluate_sensor_response(0, [1, 2], c, 0, pts) for c in coords])
        ref.used_numModes = 3
    # single mode: probes.array() drops the mode axis (model.py:930-935)
    tp = _load(os.path.join(REF, "tests/unit/test_pgdclass.py"), "ref_test_pgdclass_s")
    one = tp.create_example_pgd_solution({"A": 1, "n": 1, "lae": 1})
    for d in (1, 2):
        one.mesh[d].attributes[0].interpolationInfo = {"name": 0, "kind": "linear"}
    one.create_interpolation_fcts([1, 2], 0)
    pts1 = np.array([[0.25], [0.5], [0.9]])
    out["one_points"] = pts1
    out["one_coord"] = np.array([0.737, 0.93])
    for tag, E in (("scalar", rng.standard_normal(3)), ("vector", rng.standard_normal((3, 2)))):
        one._eval_fixed_modes = {(np.sum(pts1.flatten()), 0, 0): E}
        out["one_E_" + tag] = E
        out["one_resp_" + tag] = np.asarray(one.evaluate_sensor_response(0, [1, 2], [0.737, 0.93], 0, pts1))
    for d in range(3):
        out["one_x%d" % d] = np.asarray(one.mesh[d].dataX, dtype=np.float64)
        out["one_data%d" % d] = np.array([np.asarray(m, dtype=np.float64) for m in one.mesh[d].attributes[0].data])


def main():
    _install_stub()
    only = sys.argv[1:]
    for name, fn in (("fd_matrices", fd_matrices), ("laplace_fd", laplace_fd), ("pgdclass", pgdclass), ("pxdmf", pxdmf), ("sensor", sensor)):
        if only and name not in only:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "%d arrays, %.1f kB" % (len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
