"""Oracle (test infrastructure): matrix-form restatements of the reference's callback sets.

Every builder returns ``(SeparatedProblem, info)``; ``info["spaces"]`` are the oracle ``Space``s.
Spaces can be injected (``spaces=``) so that a parity test hands the oracle exactly the dofmap
the product generated (the dofmap is an *input*, like DOLFIN's would be).

  truss_xpe      tests/integration/test_elastic.py:71-266       (1-D truss u(x,p,E), P2)
  heat1d         tests/integration/test_heat1D.py:55-559        (FEM-in-time and FD-in-time)
  laplace_xyqu   tests/integration/test_laplace.py:73-866       (FEM and full FD)
  elasticity2d   tests/integration/test_solver_problem.py:127-627 (plane strain, vector P2)
  poisson1d_k    BASELINE.json configs[0]  (no reference callback set: parity unpinned)
  heat2d_tk      BASELINE.json configs[1]  (no reference callback set: parity unpinned)
  elasticity3d   BASELINE.json configs[2]  (no reference callback set: parity unpinned)
  thermal3d      BASELINE.json configs[3]  (no reference callback set: parity unpinned)
"""
import numpy as np
import scipy.sparse as sp

from . import fem
from .meshes import interval_mesh, rectangle_mesh
from .pgd import FD_matrices, SeparatedProblem


def _interval_spaces(num_elem, ords, ranges):
    return [fem.Space(*interval_mesh(n, r[0], r[1]), degree=o) for n, o, r in zip(num_elem, ords, ranges)]


def _mass(s, w=None, wdeg=None):
    return fem.assemble_bilinear(s, fem.T_mass(s.bs, s.gdim), weight=w, weight_degree=wdeg)


def _stiff(s):
    return fem.assemble_bilinear(s, fem.T_stiff(s.bs, s.gdim))


def _load(s, w, wdeg):
    L = np.zeros((1, s.gdim + 1))
    L[0, 0] = 1.0
    return fem.assemble_linear(s, L, weight=w, weight_degree=wdeg)


def _interp(s, f):
    return fem._feval(f, s.dof_coordinates())


def _x(x):
    return x[..., 0]


def _one(x):
    return np.ones(x.shape[:-1])


# --------------------------------------------------------------------------- test_elastic.py
def truss_xpe(num_elem=(113, 2, 100), ords=(2, 2, 2), ranges=((0, 1), (-1.0, 3.0), (0.2, 2.0)),
              A=1.0, p_0=1.0, E_0=1.0, spaces=None, **kw):
    S = spaces or _interval_spaces(num_elem, ords, ranges)
    K = [_stiff(S[0]), _mass(S[1]), _mass(S[2], _x, 4)]
    g = [A * _load(S[0], _one, 4), _load(S[1], lambda x: p_0 * A * x[..., 0], 4), _load(S[2], _one, 4)]
    bc0 = fem.dirichlet_dofs(S[0], lambda x, ob: x[0] < 1e-5 or x[0] > 1.0 - 1e-5)
    none = np.zeros(0, dtype=np.int64)
    opts = dict(PGD_nmax=10, tol_fp_it=1e-5, max_fp_it=50, stop_fp="norm", norm_modes="stiff")
    opts.update(kw)
    p = SeparatedProblem(
        n_dofs=[s.n_dofs for s in S], mass=[_mass(s) for s in S], bc_dofs=[bc0, none, none],
        lhs_terms=[(E_0 * A, K)], rhs_terms=[(1.0, g)], seq_fp=[0, 1, 2], **opts)
    return p, {"spaces": S}


# --------------------------------------------------------------------------- test_heat1D.py
def heat1d(kind="FEM", elems=(15, 10, 10), ords=(1, 1, 1), case="heating", spaces=None,
           rho=1.0, cp=1.0, k=0.5, Tamb=25.0, Q=1.0, af=0.2, ar=0.2, xc=0.5, lx=1.0, lt=1.0,
           q_fixed=1.0, **kw):
    ranges = ((0.0, lx), (0.0, lt), (0.5, 1.0))
    S = spaces or _interval_spaces(elems, ords, ranges)
    if case == "heating":
        ff = 6 * np.sqrt(3) / ((af + ar) * af * af * np.pi ** 1.5)
        q = lambda x: ff * np.exp(-3 * ((x[..., 0] - xc) ** 2 / af**2))
        IC = [_interp(S[0], _one), _interp(S[1], lambda x: Tamb * _one(x)), _interp(S[2], _one)]
    else:
        vf_a = 6 * np.sqrt(3) / (2 * af**3 * np.pi ** 1.5)
        q = lambda x: 0.0 * _one(x)
        IC = [_interp(S[0], lambda x: vf_a * np.exp(-3 * ((x[..., 0] - xc) ** 2 / af**2))),
              _interp(S[1], _one), _interp(S[2], _x)]
    Qv = [_interp(S[0], q), _interp(S[1], _one), _interp(S[2], lambda x: x[..., 0] * Q)]
    M = [_mass(s) for s in S]
    Kx = _stiff(S[0])
    t_dofs = S[1].dof_coordinates().ravel()
    if kind == "FEM":
        At = fem.assemble_bilinear(S[1], fem.T_adv(1))
        Mt = M[1]
    else:
        srt = np.argsort(t_dofs)
        M_t, _, D1 = FD_matrices(t_dofs[srt])
        Mt, At = M_t[srt, :][:, srt].tocsr(), D1[srt, :][:, srt].tocsr()  # test_heat1D.py:512-515
    mass = [M[0], Mt if kind != "FEM" else M[1], M[2]]
    lhs = [(rho * cp, [M[0], At, M[2]]), (k, [Kx, Mt, M[2]])]
    rhs = [
        (1.0, [M[0] @ Qv[0], Mt @ Qv[1], M[2] @ Qv[2]]),
        (-rho * cp, [M[0] @ IC[0], At @ IC[1], M[2] @ IC[2]]),
        (-k, [Kx @ IC[0], Mt @ IC[1], M[2] @ IC[2]]),
    ]
    none = np.zeros(0, dtype=np.int64)
    bct = np.where(t_dofs < 1e-5)[0]
    opts = dict(PGD_nmax=20, PGD_tol=1e-5, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=mass, bc_dofs=[none, bct, none],
                         lhs_terms=lhs, rhs_terms=rhs, seq_fp=[0, 1, 2], **opts)
    return p, {"spaces": S, "IC": IC, "q": q, "Q": Q, "params": dict(rho=rho, cp=cp, k=k)}


# --------------------------------------------------------------------------- test_laplace.py
def laplace_xyqu(kind="FEM", elems=(60, 40, 200, 80), ords=(1, 1, 1, 1), k=0.5, lx=3.0, ly=3.0,
                 spaces=None, Qv=None, **kw):
    ranges = ((0.0, lx), (0.0, ly), (0.0, 50.0), (10.0, 50.0))
    S = spaces or _interval_spaces(elems, ords, ranges)
    BC = [_interp(S[0], lambda x: 1.0 - x[..., 0] / 3.0), _interp(S[1], _one),
          _interp(S[2], _one), _interp(S[3], _x)]
    if Qv is None:  # source of the reference test; tests/golden variants pass their own dof vectors
        Qv = [_interp(S[0], lambda x: np.where(x[..., 0] < lx / 2, 1.0, 0.0)), _interp(S[1], _one),
              _interp(S[2], _x), _interp(S[3], _one)]
    xd = [s.dof_coordinates().ravel() for s in S]
    if kind == "FEM":
        M = [_mass(s) for s in S]
        Kx, Ky = _stiff(S[0]), _stiff(S[1])
    else:  # full FD: -k*D2 replaces the stiffness (test_laplace.py:372-401)
        M, D2 = [], []
        for x in xd:
            srt = np.argsort(x)
            m, d2, _ = FD_matrices(x[srt])
            M.append(m[srt, :][:, srt].tocsr())
            D2.append(d2[srt, :][:, srt].tocsr())
        Kx, Ky = -D2[0], -D2[1]
    lhs = [(k, [Kx, M[1], M[2], M[3]]), (k, [M[0], Ky, M[2], M[3]])]
    rhs = [
        (1.0, [M[0] @ Qv[0], M[1] @ Qv[1], M[2] @ Qv[2], M[3] @ Qv[3]]),
        (-k, [Kx @ BC[0], M[1] @ BC[1], M[2] @ BC[2], M[3] @ BC[3]]),
        (-k, [M[0] @ BC[0], Ky @ BC[1], M[2] @ BC[2], M[3] @ BC[3]]),
    ]
    bcx = np.where((np.abs(xd[0]) < 1e-6) | (np.abs(xd[0] - lx) < 1e-6))[0]
    none = np.zeros(0, dtype=np.int64)
    opts = dict(PGD_nmax=7, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=M, bc_dofs=[bcx, none, none, none],
                         lhs_terms=lhs, rhs_terms=rhs, seq_fp=[0, 1, 2, 3], **opts)
    return p, {"spaces": S, "BC": BC, "k": k, "lx": lx, "ly": ly}


# --------------------------------------------------------------------------- test_solver_problem.py
C1_NP = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
C2_NP = np.array([[1.0, -1.0, 0.0], [-1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])


def elasticity2d(N=(200, 20), order_x=2, numElems=(2, 50, 50), E_0=30000.0, L=(1000.0, 100.0),
                 g1=(0.0, -0.5), g2=(0.0, -1.5), spaces=None, **kw):
    ranges = ((0.0, 2.0), (0.5, 1.5), (0.1, 0.4))
    if spaces is None:
        Sx = fem.Space(*rectangle_mesh(0.0, 0.0, L[0], L[1], N[0], N[1], "crossed"), degree=order_x, bs=2)
        S = [Sx] + _interval_spaces(numElems, (1, 1, 1), ranges)
    else:
        S = spaces
    nu1 = lambda x: 1.0 / (2.0 * (1.0 + x[..., 0]) * (1.0 - 2.0 * x[..., 0]))
    nu2 = lambda x: 1.0 / (2.0 * (1.0 + x[..., 0]))
    Kc = [fem.assemble_bilinear(S[0], fem.T_voigt(C, 2)) for C in (C1_NP, C2_NP)]
    Mp = _mass(S[1])
    ME = E_0 * _mass(S[2], _x, 4)
    Mnu = [_mass(S[3], nu1, 10), _mass(S[3], nu2, 10)]
    lhs = [(1.0, [Kc[0], Mp, ME, Mnu[0]]), (1.0, [Kc[1], Mp, ME, Mnu[1]])]
    near = lambda a, b: abs(a - b) < 3e-16 * max(1.0, abs(a), abs(b)) + 3e-16
    top_l = fem.facet_space(S[0], lambda x: near(x[1], L[1]) and x[0] < 0.5 * L[0] + 1e-9)
    top_r = fem.facet_space(S[0], lambda x: near(x[1], L[1]) and x[0] > 0.5 * L[0] - 1e-9)
    rhs = []
    for fs, gvec in ((top_l, g1), (top_r, g2)):
        Lt = np.zeros((2, 3))
        Lt[:, 0] = gvec
        gx = fem.assemble_linear(fs, Lt)
        rhs.append((1.0, [gx, _load(S[1], _x, 4), _load(S[2], _one, 4), _load(S[3], _one, 4)]))
    bcx = fem.dirichlet_dofs(S[0], lambda x, ob: near(x[0], 0.0))
    none = np.zeros(0, dtype=np.int64)
    opts = dict(PGD_nmax=7, tol_fp_it=1e-4, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=[_mass(s) for s in S],
                         bc_dofs=[bcx, none, none, none], lhs_terms=lhs, rhs_terms=rhs,
                         seq_fp=[0, 1, 2, 3], **opts)
    return p, {"spaces": S, "E_0": E_0}


# --------------------------------------------------------------------------- BASELINE configs
def poisson1d_k(nx=999, nk=100, krange=(0.5, 2.0), ords=(1, 1), spaces=None, **kw):
    """configs[0]: -(k u')' = 1 on (0,1), u(0)=u(1)=0, k in krange. Analytic u = x(1-x)/(2k)."""
    S = spaces or _interval_spaces((nx, nk), ords, ((0.0, 1.0), krange))
    lhs = [(1.0, [_stiff(S[0]), _mass(S[1], _x, 1)])]
    rhs = [(1.0, [_load(S[0], _one, 1), _load(S[1], _one, 1)])]
    bc0 = fem.dirichlet_dofs(S[0], lambda x, ob: x[0] < 1e-8 or x[0] > 1.0 - 1e-8)
    opts = dict(PGD_nmax=10, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=[_mass(s) for s in S],
                         bc_dofs=[bc0, np.zeros(0, dtype=np.int64)], lhs_terms=lhs, rhs_terms=rhs,
                         seq_fp=[0, 1], **opts)
    return p, {"spaces": S}


def heat2d_tk(n=256, nt=199, nk=49, krange=(0.5, 2.0), rho_cp=1.0, a=0.2, xc=(0.5, 0.5), spaces=None, **kw):
    """configs[1]: rho cp u_t - k lap(u) = Q(x) on the unit square, u=0 on the boundary and at
    t=0; P1 triangles x FD-in-time (M_t, D1_up as test_heat1D.py:507-519) x P1 conductivity k."""
    if spaces is None:
        S = [fem.Space(*rectangle_mesh(0.0, 0.0, 1.0, 1.0, n, n, "right"))] + _interval_spaces(
            (nt, nk), (1, 1), ((0.0, 1.0), krange))
    else:
        S = spaces
    src = lambda x: np.exp(-3.0 * ((x[..., 0] - xc[0]) ** 2 + (x[..., 1] - xc[1]) ** 2) / a**2)
    t_dofs = S[1].dof_coordinates().ravel()
    srt = np.argsort(t_dofs)
    M_t, _, D1 = FD_matrices(t_dofs[srt])
    Mt, At = M_t[srt, :][:, srt].tocsr(), D1[srt, :][:, srt].tocsr()
    Mx, Kx = _mass(S[0]), _stiff(S[0])
    Mk, Mkk = _mass(S[2]), _mass(S[2], _x, 1)
    lhs = [(rho_cp, [Mx, At, Mk]), (1.0, [Kx, Mt, Mkk])]
    ones_t = np.ones(S[1].n_dofs)
    rhs = [(1.0, [Mx @ _interp(S[0], src), Mt @ ones_t, _load(S[2], _one, 1)])]
    bcx = np.nonzero(fem.node_on_boundary(S[0]))[0]
    bct = np.where(t_dofs < 1e-12)[0]
    opts = dict(PGD_nmax=20, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=[Mx, Mt, Mk],
                         bc_dofs=[bcx, bct, np.zeros(0, dtype=np.int64)], lhs_terms=lhs, rhs_terms=rhs,
                         seq_fp=[0, 1, 2], **opts)
    return p, {"spaces": S, "src": src}


def elasticity3d(n=68, nE=49, nF=2, nu=0.3, Erange=(0.5, 1.5), Frange=(0.0, 2.0), spaces=None, zones=3, **kw):
    """configs[2] (pgdrome_b200/configs.py:elasticity3d) in matrix form: material zones along x with moduli 1 | E | E^2
    (zones=3; zones=2: 1 | E), clamped at x=0, traction on x=1.  No reference callback set exists for it: parity unpinned
    (oracle only)."""
    from .meshes import box_mesh

    if spaces is None:
        S = [fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, n, n, n), degree=1, bs=3)] + _interval_spaces(
            (nE, nF), (1, 1), (Erange, Frange))
    else:
        S = spaces
    lam, mu = nu / ((1 + nu) * (1 - 2 * nu)), 1.0 / (2 * (1 + nu))
    T = fem.T_voigt(fem.isotropic_C(lam, mu, 3), 3)
    if zones == 2:
        chis = [lambda x: np.where(x[..., 0] < 0.5, 1.0, 0.0), lambda x: np.where(x[..., 0] < 0.5, 0.0, 1.0)]
    else:
        a, b = 1.0 / 3.0, 2.0 / 3.0
        chis = [lambda x: np.where(x[..., 0] < a, 1.0, 0.0), lambda x: np.where((x[..., 0] >= a) & (x[..., 0] < b), 1.0, 0.0),
                lambda x: np.where(x[..., 0] >= b, 1.0, 0.0)]
    Ks = [fem.assemble_bilinear(S[0], T, weight=c, weight_degree=0) for c in chis]
    MEs = [_mass(S[1]), _mass(S[1], _x, 1), _mass(S[1], lambda x: x[..., 0] * x[..., 0], 2)]
    MF = _mass(S[2])
    near = lambda a, b: abs(a - b) < 3e-16 * max(1.0, abs(a), abs(b)) + 3e-16
    face = fem.facet_space(S[0], lambda x: near(x[0], 1.0))
    Lt = np.zeros((3, 4))
    Lt[:, 0] = (0.0, 0.0, -1.0)
    gx = fem.assemble_linear(face, Lt)
    rhs = [(1.0, [gx, _load(S[1], _one, 1), _load(S[2], _x, 1)])]
    bcx = fem.dirichlet_dofs(S[0], lambda x, ob: near(x[0], 0.0))
    none = np.zeros(0, dtype=np.int64)
    opts = dict(PGD_nmax=30, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=[_mass(s) for s in S], bc_dofs=[bcx, none, none],
                         lhs_terms=[(1.0, [Ks[k], MEs[k], MF]) for k in range(len(chis))], rhs_terms=rhs, seq_fp=[0, 1, 2], **opts)
    return p, {"spaces": S}


def thermal3d(n=158, nt=199, nP=19, nv=19, kappa=0.05, rho_cp=1.0, a=0.12, n_src=6, spaces=None, source_terms=None, **kw):
    """configs[3] (pgdrome_b200/configs.py:thermal3d) in matrix form: moving source pre-separated into
    n_src terms g_m(x) h_m(t) P w_m(v); FD in time.  Parity unpinned (oracle only).
    source_terms: the separated tables of the moving Gaussian (dict with G, H, W, x, t, v, a as produced by the
    problem's set-up tool); None = the way-point surrogate (static Gaussians switched on in turn)."""
    from .meshes import box_mesh

    if spaces is None:
        S = [fem.Space(*box_mesh(0.0, 0.0, 0.0, 1.0, 1.0, 1.0, n, n, n))] + _interval_spaces(
            (nt, nP, nv), (1, 1, 1), ((0.0, 1.0), (0.5, 1.5), (0.5, 1.5)))
    else:
        S = spaces
    t_dofs = S[1].dof_coordinates().ravel()
    srt = np.argsort(t_dofs)
    M_t, _, D1 = FD_matrices(t_dofs[srt])
    Mt, At = M_t[srt, :][:, srt].tocsr(), D1[srt, :][:, srt].tocsr()
    Mx, Kx = _mass(S[0]), _stiff(S[0])
    MP, Mv = _mass(S[2]), _mass(S[3])
    lhs = [(rho_cp, [Mx, At, MP, Mv]), (kappa, [Kx, Mt, MP, Mv])]
    rhs = []
    if source_terms is not None:
        T = source_terms
        X0, v_dofs = S[0].node_coords, S[3].dof_coordinates().ravel()
        lateral = np.exp(-3.0 * ((X0[:, 1] - 0.5) ** 2 + (X0[:, 2] - 1.0) ** 2) / (T["a"] * T["a"]))
        for m in range(len(T["G"])):
            g = np.interp(X0[:, 0], T["x"], T["G"][m]) * lateral
            h = np.interp(t_dofs, T["t"], T["H"][m])
            w = np.interp(v_dofs, T["v"], T["W"][m])
            rhs.append((1.0, [Mx @ g, Mt @ h, _load(S[2], _x, 1), Mv @ w]))
        n_src = 0
    for m in range(n_src):
        xm = 0.2 + 0.6 * m / max(n_src - 1, 1)
        tm = 0.1 + 0.8 * m / max(n_src - 1, 1)
        vm = 0.5 + m / max(n_src - 1, 1)
        g = _interp(S[0], lambda x: np.exp(-3.0 * ((x[..., 0] - xm) ** 2 + (x[..., 1] - 0.5) ** 2 + (x[..., 2] - 1.0) ** 2) / (a * a)))
        h = _interp(S[1], lambda x: np.exp(-(((x[..., 0] - tm) / 0.08) ** 2)))
        w = lambda x, vm=vm: np.exp(-(((x[..., 0] - vm) / 0.6) ** 2))
        rhs.append((1.0, [Mx @ g, Mt @ h, _load(S[2], _x, 1), _load(S[3], w, 2)]))
    bcx = fem.dirichlet_dofs(S[0], lambda x, ob: abs(x[2]) < 1e-14)
    bct = np.where(t_dofs < 1e-12)[0]
    none = np.zeros(0, dtype=np.int64)
    opts = dict(PGD_nmax=50, tol_fp_it=1e-5, max_fp_it=50)
    opts.update(kw)
    p = SeparatedProblem(n_dofs=[s.n_dofs for s in S], mass=[Mx, Mt, MP, Mv], bc_dofs=[bcx, bct, none, none],
                         lhs_terms=lhs, rhs_terms=rhs, seq_fp=[0, 1, 2, 3], **opts)
    return p, {"spaces": S}
