"""The reference's own integration problems written as PGDrome user scripts for the B200 package
(callbacks in UFL through pgdrome_b200.dolfin), one builder per reference test:

  truss_xpe      tests/integration/test_elastic.py:71-266       1-D truss u(x,p,E), P2, Newton path
  heat1d         tests/integration/test_heat1D.py:55-559        FEM-in-time and FD-in-time (mixed solve modes)
  laplace_fem    tests/integration/test_laplace.py:73-330       2-D Laplace as x,y,q,u0 with BC lifting
  elasticity2d   tests/integration/test_solver_problem.py:127-627  plane strain, vector P2, facet tractions

They are test inputs (user code), not part of the package; oracle/problems.py holds the matrix-form
restatements the results are compared with."""
import numpy as np

from pgdrome_b200 import dolfin as df
from pgdrome_b200.configs import _separated_problem
from pgdrome_b200.solver import FD_matrices


def _spaces(num_elem, ords, ranges):
    return [df.FunctionSpace(df.IntervalMesh(n, r[0], r[1]), "CG", o) for n, o, r in zip(num_elem, ords, ranges)]


def _mass(mesh):
    return lambda u, v: u * v * df.dx(mesh)


def _stiff1(mesh):
    return lambda u, v: u.dx(0) * v.dx(0) * df.dx(mesh)


def _fn(V, values):
    f = df.Function(V)
    f.vector()[:] = values
    return f


def truss_xpe(num_elem=(113, 2, 100), ords=(2, 2, 2), A=1.0, p_0=1.0, E_0=1.0, PGD_nmax=10, **attrs):
    Vs = _spaces(num_elem, ords, ((0, 1), (-1.0, 3.0), (0.2, 2.0)))
    m = [v.mesh() for v in Vs]
    Efunc = df.Expression("x[0]", degree=4)
    ops = [[_stiff1(m[0]), _mass(m[1]), lambda u, v: u * Efunc * v * df.dx(m[2])]]
    g = [df.Expression("1.0", degree=4), df.Expression("p*A*x[0]", p=p_0, A=A, degree=4), df.Expression("1.0", degree=4)]
    loads = [[lambda w: w * g[0] * A * df.dx(m[0]), lambda w: w * g[1] * df.dx(m[1]), lambda w: w * g[2] * df.dx(m[2])]]

    def bc_fct(Vs, dom, param):
        def left_right(x, on_boundary):
            return x < 0.0 + 1e-5 or x > 1.0 - 1e-5

        return [df.DirichletBC(Vs[0], 0, left_right), 0, 0]

    return _separated_problem("Uniaxial1D-PGD-XPE", ["X", "P", "E"], Vs, ops, [E_0 * A], loads, bc_fct, ["r", "s", "t"],
                              PGD_nmax, **attrs)


def heat1d(kind="FEM", elems=(15, 10, 10), case="heating", rho=1.0, cp=1.0, k=0.5, Tamb=25.0, Q=1.0, af=0.2, ar=0.2,
           xc=0.5, **attrs):
    Vs = _spaces(elems, (1, 1, 1), ((0.0, 1.0), (0.0, 1.0), (0.5, 1.0)))
    m = [v.mesh() for v in Vs]
    xd = [v.tabulate_dof_coordinates()[:, 0] for v in Vs]
    if case == "heating":
        ff = 6 * np.sqrt(3) / ((af + ar) * af * af * np.pi ** 1.5)
        qx = ff * np.exp(-3 * ((xd[0] - xc) ** 2 / af**2))
        IC = [_fn(Vs[0], np.ones_like(xd[0])), _fn(Vs[1], Tamb * np.ones_like(xd[1])), _fn(Vs[2], np.ones_like(xd[2]))]
    else:
        vf_a = 6 * np.sqrt(3) / (2 * af**3 * np.pi ** 1.5)
        qx = 0.0 * xd[0]
        IC = [_fn(Vs[0], vf_a * np.exp(-3 * ((xd[0] - xc) ** 2 / af**2))), _fn(Vs[1], np.ones_like(xd[1])), _fn(Vs[2], xd[2])]
    Qf = [_fn(Vs[0], qx), _fn(Vs[1], np.ones_like(xd[1])), _fn(Vs[2], xd[2] * Q)]
    MM = None
    if kind == "FEM":
        adv_t = lambda u, v: u.dx(0) * v * df.dx(m[1])
        mass_t = _mass(m[1])
        load_t = lambda w: Qf[1] * w * df.dx(m[1])
        modes = None
    else:
        srt = np.argsort(xd[1])
        M_t, _, D1 = FD_matrices(xd[1][srt])
        Mt = df.MatrixOperator(M_t.tocsr()[srt, :][:, srt], Vs[1])
        Dt = df.MatrixOperator(D1.tocsr()[srt, :][:, srt], Vs[1])
        adv_t = lambda u, v: Dt(u, v) * df.dx(m[1])
        mass_t = lambda u, v: Mt(u, v) * df.dx(m[1])
        load_t = lambda w: Mt(Qf[1], w) * df.dx(m[1])
        MM = [0, Mt, 0]
        modes = ["FEM", "FEM", "FEM"]  # the time operators are device MatrixOperators inside FEM-style forms
    ops = [[_mass(m[0]), adv_t, _mass(m[2])], [_stiff1(m[0]), mass_t, _mass(m[2])]]
    loads = [[lambda w: Qf[0] * w * df.dx(m[0]), load_t, lambda w: Qf[2] * w * df.dx(m[2])]]

    def bc_fct(Vs, dom, param):
        def initial(x, on_boundary):
            return x < 0.0 + 1e-5

        return [0, df.DirichletBC(Vs[1], 0, initial), 0]

    opts = dict(PGD_tol=1e-5)
    opts.update(attrs)
    p = _separated_problem("heat1d", ["X", "T", "Q"], Vs, ops, [rho * cp, k], loads, bc_fct, ["r", "s", "w"], 20, MM=MM,
                           lifts=[IC], **opts)
    p._solve_modes = modes
    return p


def laplace_fem(elems=(60, 40, 200, 80), k=0.5, lx=3.0, **attrs):
    Vs = _spaces(elems, (1, 1, 1, 1), ((0.0, lx), (0.0, 3.0), (0.0, 50.0), (10.0, 50.0)))
    m = [v.mesh() for v in Vs]
    xd = [v.tabulate_dof_coordinates()[:, 0] for v in Vs]
    BC = [_fn(Vs[0], 1.0 - xd[0] / 3.0), _fn(Vs[1], np.ones_like(xd[1])), _fn(Vs[2], np.ones_like(xd[2])), _fn(Vs[3], xd[3])]
    Qf = [_fn(Vs[0], np.where(xd[0] < lx / 2, 1.0, 0.0)), _fn(Vs[1], np.ones_like(xd[1])), _fn(Vs[2], xd[2]),
          _fn(Vs[3], np.ones_like(xd[3]))]
    ops = [[_stiff1(m[0]), _mass(m[1]), _mass(m[2]), _mass(m[3])], [_mass(m[0]), _stiff1(m[1]), _mass(m[2]), _mass(m[3])]]
    loads = [[(lambda w, i=i: Qf[i] * w * df.dx(m[i])) for i in range(4)]]

    def bc_fct(Vs, dom, param):
        def leftright(x, on_boundary):
            return on_boundary and df.near(x[0], 0.0, 1e-6) or df.near(x[0], lx, 1e-6)

        return [df.DirichletBC(Vs[0], 0, leftright), 0, 0, 0]

    return _separated_problem("test_x_y_q_u00", ["X", "Y", "q", "u0"], Vs, ops, [k, k], loads, bc_fct, ["r", "s", "t", "u"], 7,
                              lifts=[BC], **attrs), BC


def elasticity2d(N=(200, 20), numElems=(2, 50, 50), E_0=30000.0, L=(1000.0, 100.0), g1=(0.0, -0.5), g2=(0.0, -1.5), PGD_nmax=7,
                 **attrs):
    mx = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(L[0], L[1]), N[0], N[1], "crossed")
    Vs = [df.VectorFunctionSpace(mx, "CG", 2)] + _spaces(numElems, (1, 1, 1), ((0.0, 2.0), (0.5, 1.5), (0.1, 0.4)))
    m = [v.mesh() for v in Vs]
    C1 = df.as_matrix([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
    C2 = df.as_matrix([[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])

    def eps(v):
        return df.as_vector([v[0].dx(0), v[1].dx(1), v[0].dx(1) + v[1].dx(0)])

    Efunc = df.Expression("x[0]", degree=4)
    nu1 = df.Expression("1.0/(2.0*(1.0+x[0])*(1.0-2.0*x[0]))", degree=10)
    nu2 = df.Expression("1.0/(2.0*(1.0+x[0]))", degree=10)
    ops = []
    for C, nu in ((C1, nu1), (C2, nu2)):
        ops.append([lambda u, v, C=C: df.inner(C * eps(u), eps(v)) * df.dx(m[0]), _mass(m[1]),
                    lambda u, v: u * Efunc * E_0 * v * df.dx(m[2]), lambda u, v, nu=nu: u * nu * v * df.dx(m[3])])
    facets = df.MeshFunction("size_t", mx, 1, 0)

    class TopLeft(df.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary and df.near(x[1], L[1]) and x[0] < 0.5 * L[0] + 1e-9

    class TopRight(df.SubDomain):
        def inside(self, x, on_boundary):
            return on_boundary and df.near(x[1], L[1]) and x[0] > 0.5 * L[0] - 1e-9

    TopLeft().mark(facets, 2)
    TopRight().mark(facets, 3)
    ds = df.Measure("ds", domain=mx, subdomain_data=facets)
    px = df.Expression("x[0]", degree=4)
    one = df.Expression("1.0", degree=4)
    loads = []
    for sid, gv in ((2, g1), (3, g2)):
        t = df.Constant(gv)
        loads.append([lambda w, t=t, sid=sid: df.dot(t, w) * ds(sid), lambda w: px * w * df.dx(m[1]),
                      lambda w: one * w * df.dx(m[2]), lambda w: one * w * df.dx(m[3])])

    def bc_fct(Vs, dom, param):
        def left(x, on_boundary):
            return on_boundary and df.near(x[0], 0.0)

        return [df.DirichletBC(Vs[0], df.Constant((0.0, 0.0)), left), 0, 0, 0]

    opts = dict(tol_fp_it=1e-4)
    opts.update(attrs)
    return _separated_problem("elasticity2d", ["X", "P", "E", "Nu"], Vs, ops, [1.0, 1.0], loads, bc_fct, ["r", "s", "t", "u"],
                              PGD_nmax, **opts)
