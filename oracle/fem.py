"""Oracle (test infrastructure): Lagrange P1/P2 finite elements on simplices, NumPy/SciPy.

Restates what DOLFIN 2019.1 does underneath the reference's calls [DOLFIN-knowledge]:
``dolfin.assemble`` of bilinear / linear / scalar forms (pgdrome/solver.py:365-367,443,
839-841 and every callback in tests/integration/*.py), ``dolfin.norm`` (solver.py:207,754),
``DirichletBC`` + variational solve with symmetric elimination (solver.py:627-636,704-716).

Conventions (shared with the product through *inputs*, never through imports):
  * a space is (coords, cells, degree, bs); nodes = Lagrange points, dof = bs*node + comp
  * 1-D meshes number their nodes by DEscending coordinate (tests/unit/test_FD.py:68-79
    implies DOLFIN's serial 1-D P1 numbering is reversed); 2-D/3-D: vertex nodes = vertex
    index, then (P2) edge nodes in order of first appearance (cell-major, FIAT local edge order)
  * sparsity = union of per-cell dof cliques, columns ascending, explicit zeros kept
  * a bilinear atom is a constant tensor T[iv, jv, iu, ju]: coefficient of
    D_jv v_iv * D_ju u_iu  with slot 0 = value, slot 1+m = d/dx_m; row = test, col = trial
"""
import numpy as np
import scipy.sparse as sp

LOCAL_EDGES = {
    1: [],
    2: [(1, 2), (0, 2), (0, 1)],
    3: [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)],
}


class Space:
    def __init__(self, coords, cells, degree=1, bs=1):
        coords = np.asarray(coords, dtype=np.float64)
        if coords.ndim == 1:
            coords = coords.reshape(-1, 1)
        cells = np.asarray(cells, dtype=np.int64)
        self.coords, self.cells = coords, cells
        self.gdim = coords.shape[1]
        self.tdim = cells.shape[1] - 1
        self.degree, self.bs = int(degree), int(bs)
        nv, nc = coords.shape[0], cells.shape[0]
        if self.degree not in (1, 2):
            raise ValueError("oracle supports P1/P2")
        if self.tdim == 1:
            # nodes by descending coordinate
            order = np.argsort(coords[:, 0], kind="stable")
            rank = np.empty(nv, dtype=np.int64)
            rank[order] = np.arange(nv)
            if self.degree == 1:
                v2n = nv - 1 - rank
                self.cell_nodes = v2n[cells]
                self.node_coords = np.empty((nv, 1))
                self.node_coords[v2n] = coords
            else:
                ntot = 2 * nc + 1
                v2n = ntot - 1 - 2 * rank
                lo = np.minimum(rank[cells[:, 0]], rank[cells[:, 1]])
                mid = ntot - 1 - (2 * lo + 1)
                self.cell_nodes = np.concatenate([v2n[cells], mid[:, None]], axis=1)
                self.node_coords = np.empty((ntot, 1))
                self.node_coords[v2n] = coords
                self.node_coords[mid] = 0.5 * (coords[cells[:, 0]] + coords[cells[:, 1]])
            self.vertex_to_node = v2n
        else:
            self.vertex_to_node = np.arange(nv)
            if self.degree == 1:
                self.cell_nodes = cells.copy()
                self.node_coords = coords.copy()
            else:
                le = np.array(LOCAL_EDGES[self.tdim])
                ev = np.sort(cells[:, le], axis=2).reshape(-1, 2)  # cell-major, local-edge order
                key = ev[:, 0] * nv + ev[:, 1]
                uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
                # number edges by first appearance
                appearance = np.argsort(first, kind="stable")
                renum = np.empty(len(uniq), dtype=np.int64)
                renum[appearance] = np.arange(len(uniq))
                edge_id = renum[inv].reshape(nc, len(le))
                self.cell_nodes = np.concatenate([cells, nv + edge_id], axis=1)
                ecoord = np.empty((len(uniq), self.gdim))
                ecoord[renum[inv]] = 0.5 * (coords[ev[:, 0]] + coords[ev[:, 1]])
                self.node_coords = np.vstack([coords, ecoord])
        self.n_nodes = self.node_coords.shape[0]
        self.n_dofs = self.n_nodes * self.bs
        self.nd = self.cell_nodes.shape[1]
        self.cell_dofs = (
            self.cell_nodes[:, :, None] * self.bs + np.arange(self.bs)[None, None, :]
        ).reshape(nc, self.nd * self.bs)

    def dof_coordinates(self):
        return np.repeat(self.node_coords, self.bs, axis=0)


# ----------------------------------------------------------------------------- tabulation
def tabulate(tdim, degree, pts):
    """phi[q, a], dphi[q, a, tdim] of Lagrange P1/P2 on the reference simplex."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, tdim)
    nq = pts.shape[0]
    lam = np.concatenate([1.0 - pts.sum(1, keepdims=True), pts], axis=1)  # [q, tdim+1]
    dlam = np.zeros((tdim + 1, tdim))
    dlam[0, :] = -1.0
    dlam[1:, :] = np.eye(tdim)
    if degree == 1:
        return lam.copy(), np.broadcast_to(dlam, (nq, tdim + 1, tdim)).copy()
    edges = LOCAL_EDGES[tdim] if tdim > 1 else [(0, 1)]
    nd = tdim + 1 + len(edges)
    phi = np.zeros((nq, nd))
    dphi = np.zeros((nq, nd, tdim))
    for a in range(tdim + 1):
        phi[:, a] = lam[:, a] * (2 * lam[:, a] - 1)
        dphi[:, a, :] = (4 * lam[:, a] - 1)[:, None] * dlam[a][None, :]
    for e, (i, j) in enumerate(edges):
        a = tdim + 1 + e
        phi[:, a] = 4 * lam[:, i] * lam[:, j]
        dphi[:, a, :] = 4 * (lam[:, i, None] * dlam[j][None, :] + lam[:, j, None] * dlam[i][None, :])
    return phi, dphi


def quadrature(tdim, deg):
    """Gauss-Legendre x Duffy collapse; exact for total degree <= deg. Weights sum to 1/tdim!."""
    n = max(1, (deg + tdim + 1) // 2 + 1)
    x, w = np.polynomial.legendre.leggauss(n)
    x, w = 0.5 * (x + 1), 0.5 * w
    if tdim == 1:
        return x.reshape(-1, 1), w
    if tdim == 2:
        U, V = np.meshgrid(x, x, indexing="ij")
        W = np.outer(w, w) * (1 - U)
        return np.stack([U.ravel(), (V * (1 - U)).ravel()], 1), W.ravel()
    U, V, Wc = np.meshgrid(x, x, x, indexing="ij")
    W = np.einsum("i,j,k->ijk", w, w, w) * (1 - U) ** 2 * (1 - V)
    pts = np.stack([U.ravel(), (V * (1 - U)).ravel(), (Wc * (1 - U) * (1 - V)).ravel()], 1)
    return pts, W.ravel()


def lagrange_1d(p, xi):
    """Equispaced degree-p Lagrange basis on [0,1] at xi -> [len(xi), p+1] (product form)."""
    nodes = np.arange(p + 1) / p if p > 0 else np.array([0.5])
    xi = np.asarray(xi, dtype=np.float64)
    L = np.ones((xi.size, p + 1))
    for m in range(p + 1):
        for k in range(p + 1):
            if k != m:
                L[:, m] *= (xi - nodes[k]) / (nodes[m] - nodes[k])
    return nodes, L


def weight_at_quad(space, f, degree, pts):
    """Values at reference points ``pts`` of the per-cell P_degree interpolant of f(x)
    (DOLFIN semantics of Expression(..., degree=p) inside a form [DOLFIN-knowledge])."""
    X = space.coords[space.cells]  # [e, tdim+1, gdim]
    if degree == 0:
        lam = np.full((1, space.tdim + 1), 1.0 / (space.tdim + 1))
        xc = np.einsum("qa,eag->eqg", lam, X)
        return np.repeat(_feval(f, xc)[:, :1], len(pts), axis=1)
    if space.tdim == 1:
        nodes, L = lagrange_1d(degree, np.asarray(pts).ravel())
        xn = X[:, 0, None, :] + nodes[None, :, None] * (X[:, 1, None, :] - X[:, 0, None, :])
        fn = _feval(f, xn)  # [e, p+1]
        return fn @ L.T
    if degree > 2:
        raise NotImplementedError("oracle: weight degree > 2 only on 1-D meshes")
    # nodal points of P_degree in reference coordinates
    ref_nodes = np.vstack([np.zeros((1, space.tdim)), np.eye(space.tdim)])
    if degree == 2:
        ref_nodes = np.vstack(
            [ref_nodes] + [0.5 * (ref_nodes[i] + ref_nodes[j])[None] for (i, j) in LOCAL_EDGES[space.tdim]]
        )
    lam_n = np.concatenate([1 - ref_nodes.sum(1, keepdims=True), ref_nodes], 1)
    xn = np.einsum("na,eag->eng", lam_n, X)
    fn = _feval(f, xn)
    phi, _ = tabulate(space.tdim, degree, pts)
    return fn @ phi.T


def _feval(f, x):
    """f takes x[..., gdim] and returns [...]; constants allowed."""
    if callable(f):
        out = np.asarray(f(x), dtype=np.float64)
        return np.broadcast_to(out, x.shape[:-1]).copy()
    return np.full(x.shape[:-1], float(f))


def _geometry(space, dphi):
    """Physical slot table D[e, q, a, 1+gdim] (slot 0 filled by caller) and detJ[e]."""
    X = space.coords[space.cells]
    J = np.swapaxes(X[:, 1:, :] - X[:, :1, :], 1, 2)  # [e, gdim, tdim]
    if space.tdim == space.gdim:
        detJ = np.abs(np.linalg.det(J))
        Jinv = np.linalg.inv(J)  # [e, tdim, gdim]
        grad = np.einsum("qat,etg->eqag", dphi, Jinv)
    else:
        G = np.einsum("egt,egs->ets", J, J)
        detJ = np.sqrt(np.abs(np.linalg.det(G)))
        grad = None
    return grad, detJ


C_FAST_MIN_CELLS = 20000  # assemble_bilinear uses oracle/cfem.c from this many P1 cells on (None: never)


def sparsity(cell_dofs, n_dofs):
    """CSR pattern (rowptr, colidx): union of per-cell cliques, columns ascending."""
    nd = cell_dofs.shape[1]
    r = np.repeat(cell_dofs, nd, axis=1).ravel()
    c = np.tile(cell_dofs, (1, nd)).ravel()
    key = np.unique(r.astype(np.int64) * n_dofs + c)
    rows, cols = key // n_dofs, key % n_dofs
    rowptr = np.zeros(n_dofs + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32)


def assemble_bilinear(space, T, weight=None, weight_degree=None, qdeg=None, chunk=200000):
    """A[row=test dof, col=trial dof] = int w * sum T[iv,jv,iu,ju] D_jv v_iv D_ju u_iu dx."""
    T = np.asarray(T, dtype=np.float64)
    bs, g = space.bs, space.gdim
    assert T.shape == (bs, g + 1, bs, g + 1), T.shape
    wdeg = 0 if weight is None else (weight_degree if weight_degree is not None else 2)
    if (C_FAST_MIN_CELLS is not None and space.degree == 1 and space.tdim == g and qdeg is None and wdeg == 0
            and space.cells.shape[0] >= C_FAST_MIN_CELLS):
        # large P1 meshes: the same integrals in closed form by the C restatement (oracle/cfem.c, validated against
        # the quadrature path below in tests/test_oracle_c.py); a degree-0 coefficient is its value at the cell centroid
        from . import cfem

        wc = None if weight is None else np.ascontiguousarray(weight_at_quad(space, weight, 0, np.zeros((1, g)))[:, 0])
        return cfem.assemble_bilinear_p1(space, T, wc)
    if qdeg is None:
        qdeg = 2 * space.degree + wdeg
    pts, wts = quadrature(space.tdim, qdeg)
    phi, dphi = tabulate(space.tdim, space.degree, pts)
    ne, nd = space.cells.shape[0], space.nd
    rows, cols, vals = [], [], []
    for s in range(0, ne, chunk):
        sub = _SubSpace(space, s, min(ne, s + chunk))
        grad, detJ = _geometry(sub, dphi)
        D = np.zeros((sub.cells.shape[0], len(wts), nd, g + 1))
        D[:, :, :, 0] = phi[None]
        if grad is not None:
            D[:, :, :, 1:] = grad
        W = detJ[:, None] * wts[None, :]
        if weight is not None:
            W = W * weight_at_quad(sub, weight, wdeg, pts)
        Ae = np.einsum("eq,ijkl,eqaj,eqbl->eaibk", W, T, D, D, optimize=True)
        Ae = Ae.reshape(sub.cells.shape[0], nd * bs, nd * bs)
        cd = space.cell_dofs[s : s + sub.cells.shape[0]]
        rows.append(np.repeat(cd, nd * bs, axis=1).ravel())
        cols.append(np.tile(cd, (1, nd * bs)).ravel())
        vals.append(Ae.ravel())
    A = sp.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
        shape=(space.n_dofs, space.n_dofs),
    ).tocsr()
    A.sort_indices()
    return A


def assemble_linear(space, L, weight=None, weight_degree=None, qdeg=None, cells=None):
    """b[test dof] = int w * sum L[iv, jv] D_jv v_iv dx   (L: [bs, gdim+1])."""
    L = np.asarray(L, dtype=np.float64)
    bs, g = space.bs, space.gdim
    assert L.shape == (bs, g + 1)
    wdeg = 0 if weight is None else (weight_degree if weight_degree is not None else 2)
    if qdeg is None:
        qdeg = space.degree + wdeg
    pts, wts = quadrature(space.tdim, qdeg)
    phi, dphi = tabulate(space.tdim, space.degree, pts)
    grad, detJ = _geometry(space, dphi)
    ne, nd = space.cells.shape[0], space.nd
    D = np.zeros((ne, len(wts), nd, g + 1))
    D[:, :, :, 0] = phi[None]
    if grad is not None:
        D[:, :, :, 1:] = grad
    W = detJ[:, None] * wts[None, :]
    if weight is not None:
        W = W * weight_at_quad(space, weight, wdeg, pts)
    be = np.einsum("eq,ij,eqaj->eai", W, L, D, optimize=True).reshape(ne, nd * bs)
    b = np.zeros(space.n_dofs)
    np.add.at(b, space.cell_dofs.ravel(), be.ravel())
    return b


class _SubSpace:
    """Cell slice of a Space (geometry helpers only need coords/cells/tdim/gdim)."""

    def __init__(self, space, s, e):
        self.coords, self.cells = space.coords, space.cells[s:e]
        self.tdim, self.gdim = space.tdim, space.gdim


def facet_space(space, marker_fn):
    """Boundary-facet sub-"space" (tdim = gdim-1 elements embedded in gdim) holding the
    facets whose vertices all satisfy marker_fn(x); shares node numbering with ``space``.
    Used for ``ds(id)`` integrals (tests/integration/test_solver_problem.py:256-282)."""
    t = space.tdim
    nv = space.coords.shape[0]
    loc = [tuple(k for k in range(t + 1) if k != i) for i in range(t + 1)]  # facet i opposite vertex i
    fv = np.stack([space.cells[:, list(l)] for l in loc], axis=1)  # [e, t+1, t]
    key = np.zeros(fv.shape[:2], dtype=np.int64)
    for k in range(t):
        key = key * nv + fv[:, :, k]
    uniq, counts = np.unique(key.ravel(), return_counts=True)
    on_bnd = counts[np.searchsorted(uniq, key)] == 1
    e_idx, f_idx = np.nonzero(on_bnd)
    sel = []
    for e, f in zip(e_idx, f_idx):
        if all(marker_fn(space.coords[v]) for v in fv[e, f]):
            sel.append((e, f))
    fs = Space.__new__(Space)
    fs.coords, fs.gdim, fs.tdim = space.coords, space.gdim, t - 1
    fs.degree, fs.bs = space.degree, space.bs
    fs.cells = np.array([fv[e, f] for e, f in sel], dtype=np.int64).reshape(-1, t)
    nodes = []
    for e, f in sel:
        vloc = list(loc[f])
        nn = [space.cell_nodes[e, k] for k in vloc]
        if space.degree == 2:
            if t == 2:  # facet = edge (vloc[0], vloc[1]) -> its midpoint node
                le = LOCAL_EDGES[2].index(tuple(vloc))
                nn.append(space.cell_nodes[e, 3 + le])
            else:
                sub_edges = [(vloc[1], vloc[2]), (vloc[0], vloc[2]), (vloc[0], vloc[1])]
                for se in sub_edges:
                    nn.append(space.cell_nodes[e, 4 + LOCAL_EDGES[3].index(se)])
        nodes.append(nn)
    fs.cell_nodes = np.array(nodes, dtype=np.int64).reshape(len(sel), -1)
    fs.nd = fs.cell_nodes.shape[1]
    fs.n_nodes, fs.n_dofs, fs.node_coords = space.n_nodes, space.n_dofs, space.node_coords
    fs.cell_dofs = (fs.cell_nodes[:, :, None] * fs.bs + np.arange(fs.bs)[None, None, :]).reshape(len(sel), -1)
    return fs


# ----------------------------------------------------------------------------- atoms
def T_mass(bs, g):
    T = np.zeros((bs, g + 1, bs, g + 1))
    for i in range(bs):
        T[i, 0, i, 0] = 1.0
    return T


def T_stiff(bs, g):
    T = np.zeros((bs, g + 1, bs, g + 1))
    for i in range(bs):
        for m in range(g):
            T[i, 1 + m, i, 1 + m] = 1.0
    return T


def T_adv(g, m=0):
    """int (d u / d x_m) v : derivative on the TRIAL function (test_heat1D.py:80)."""
    T = np.zeros((1, g + 1, 1, g + 1))
    T[0, 0, 0, 1 + m] = 1.0
    return T


def T_voigt(C, g):
    """inner(C * eps(u), eps(v)) with Voigt strain (test_solver_problem.py:131-148)."""
    if g == 2:
        voigt = [[(0, 0, 1.0)], [(1, 1, 1.0)], [(0, 1, 1.0), (1, 0, 1.0)]]
    else:
        voigt = [
            [(0, 0, 1.0)], [(1, 1, 1.0)], [(2, 2, 1.0)],
            [(1, 2, 1.0), (2, 1, 1.0)], [(0, 2, 1.0), (2, 0, 1.0)], [(0, 1, 1.0), (1, 0, 1.0)],
        ]
    C = np.asarray(C, dtype=np.float64)
    T = np.zeros((g, g + 1, g, g + 1))
    for a, ea in enumerate(voigt):  # test strain component a
        for b, eb in enumerate(voigt):  # trial strain component b ; (C eps(u))_a = C[a,b] eps_b(u)
            for (iv, mv, cv) in ea:
                for (iu, mu, cu) in eb:
                    T[iv, 1 + mv, iu, 1 + mu] += C[a, b] * cv * cu
    return T


def isotropic_C(lmbda, mu, g=3):
    n = 3 if g == 2 else 6
    C = np.zeros((n, n))
    C[:g, :g] = lmbda
    C[np.arange(g), np.arange(g)] += 2 * mu
    C[np.arange(g, n), np.arange(g, n)] = mu
    return C


# ----------------------------------------------------------------------------- BC / solve
def node_on_boundary(space):
    t, nv = space.tdim, space.coords.shape[0]
    flag = np.zeros(space.n_nodes, dtype=bool)
    if t == 1:
        cnt = np.bincount(space.cells.ravel(), minlength=nv)
        flag[space.vertex_to_node[np.nonzero(cnt == 1)[0]]] = True
        return flag
    fs = facet_space(space, lambda x: True)
    flag[np.unique(fs.cell_nodes)] = True
    return flag


def dirichlet_dofs(space, inside, comps=None):
    """Dofs whose node satisfies inside(x, on_boundary) (pointwise DirichletBC)."""
    onb = node_on_boundary(space)
    nodes = [n for n in range(space.n_nodes) if inside(space.node_coords[n], bool(onb[n]))]
    comps = range(space.bs) if comps is None else comps
    return np.array(sorted(n * space.bs + c for n in nodes for c in comps), dtype=np.int64)


def apply_dirichlet_sym(A, b, bc_dofs, values=0.0):
    """Symmetric elimination keeping the pattern: zero row+col, diag 1, rhs lifted."""
    A = A.tocsr(copy=True)
    b = None if b is None else np.array(b, dtype=np.float64, copy=True)
    bc_dofs = np.asarray(bc_dofs, dtype=np.int64)
    if bc_dofs.size == 0:
        return A, b
    vals = np.broadcast_to(np.asarray(values, dtype=np.float64), bc_dofs.shape)
    if b is not None:
        g = np.zeros(A.shape[0])
        g[bc_dofs] = vals
        b -= A @ g
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[bc_dofs] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    kill = mask[rows] | mask[A.indices]
    A.data[kill] = 0.0
    A.data[kill & (rows == A.indices)] = 1.0
    if b is not None:
        b[bc_dofs] = vals
    return A, b
