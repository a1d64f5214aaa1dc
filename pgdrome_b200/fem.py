"""Host-side mesh / dofmap / element-table layer of pgdrome_b200 (NumPy, set-up only).

This plays the role DOLFIN 2019.1 plays for the reference *once, on the host*: it produces the
mesh, the dofmap, dof coordinates and boundary markers (north_star: "DOLFIN is used only once, on
the host").  All arithmetic on dofs (assembly, BC, solves, integrals) runs in libpgdb200.so.
Nothing here imports ``oracle``.

Conventions ([DOLFIN-knowledge]):
  * built-in meshes use DOLFIN's vertex numbering and cell splitting (cells' vertices ascending);
  * 1-D spaces number dofs by DEscending coordinate (implied by tests/unit/test_FD.py:68-79);
  * 2-D/3-D: P1 dof = vertex; P2 appends edge dofs in order of first appearance; vector spaces are
    node-blocked (dof = bs*node + comp).  A dofmap read from a real DOLFIN can be injected through
    ``FunctionSpace.from_arrays`` instead.
"""
import numpy as np
from scipy.special import roots_jacobi

_TRI_EDGES = ((1, 2), (0, 2), (0, 1))
_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


# ------------------------------------------------------------------------------- meshes
class _Topology:
    def __init__(self, d):
        self._d = d

    def dim(self):
        return self._d


class Mesh:
    def __init__(self, coords, cells):
        self._x = np.ascontiguousarray(np.asarray(coords, dtype=np.float64).reshape(len(coords), -1))
        self._c = np.ascontiguousarray(np.sort(np.asarray(cells, dtype=np.int32), axis=1))
        self.tdim = self._c.shape[1] - 1
        self.gdim = self._x.shape[1]
        self._cache = {}

    def coordinates(self):
        return self._x

    def cells(self):
        return self._c

    def num_cells(self):
        return self._c.shape[0]

    def num_vertices(self):
        return self._x.shape[0]

    def topology(self):
        return _Topology(self.tdim)

    def geometry(self):
        return _Topology(self.gdim)

    def hmin(self):
        return float(self._edge_lengths().min())

    def hmax(self):
        return float(self._edge_lengths().max())

    def _edge_lengths(self):
        X = self._x[self._c]
        out = []
        for i in range(self.tdim + 1):
            for j in range(i + 1, self.tdim + 1):
                out.append(np.linalg.norm(X[:, i] - X[:, j], axis=1))
        return np.concatenate(out)

    # boundary facets: (cell, local facet) pairs, facet i is opposite local vertex i
    def boundary_facets(self):
        if "bf" not in self._cache:
            t, nv = self.tdim, self.num_vertices()
            if t == 1:
                cnt = np.bincount(self._c.ravel(), minlength=nv)
                verts = np.nonzero(cnt == 1)[0]
                cell, loc = [], []
                for v in verts:
                    e, l = np.argwhere(self._c == v)[0]
                    cell.append(e)
                    loc.append(1 - l)  # facet opposite the other vertex == this vertex
                self._cache["bf"] = (np.array(cell, dtype=np.int64), np.array(loc, dtype=np.int64))
            else:
                keys = []
                for i in range(t + 1):
                    idx = [k for k in range(t + 1) if k != i]
                    f = self._c[:, idx].astype(np.int64)
                    k = f[:, 0]
                    for m in range(1, t):
                        k = k * nv + f[:, m]
                    keys.append(k)
                keys = np.stack(keys, axis=1)
                single = _appears_once(keys.ravel()).reshape(keys.shape)
                cell, loc = np.nonzero(single)
                self._cache["bf"] = (cell.astype(np.int64), loc.astype(np.int64))
        return self._cache["bf"]

    def facet_vertices(self, cell, loc):
        idx = np.array([[k for k in range(self.tdim + 1) if k != i] for i in range(self.tdim + 1)])
        return self._c[cell[:, None], idx[loc]]


def _appears_once(keys):
    """mask of the entries of an int64 key array that occur exactly once.  Sorting 24 M facet keys of a
    6 M-tet mesh takes seconds in NumPy and ~10 ms on the device, so the sort runs there when a GPU is
    present (set-up plumbing only; cells stay host arrays like DOLFIN's mesh)."""
    if len(keys) > 200000:
        try:
            import torch

            if torch.cuda.is_available():
                k = torch.as_tensor(keys, device="cuda")
                srt, order = torch.sort(k)
                once = torch.ones_like(srt, dtype=torch.bool)
                eq = srt[1:] == srt[:-1]
                once[1:] &= ~eq
                once[:-1] &= ~eq
                out = torch.empty_like(once)
                out[order] = once
                return out.cpu().numpy()
        except Exception:
            pass
    _, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    return cnt[inv] == 1


class Point:
    def __init__(self, *xyz):
        self._v = np.array(xyz, dtype=np.float64)

    def __getitem__(self, i):
        return self._v[i]

    def array(self):
        return self._v


def _pt(p):
    return p.array() if isinstance(p, Point) else np.atleast_1d(np.asarray(p, dtype=np.float64))


def IntervalMesh(n, a, b):
    x = a + (b - a) * np.arange(n + 1) / n
    c = np.column_stack([np.arange(n), np.arange(n) + 1])
    return Mesh(x[:, None], c)


def UnitIntervalMesh(n):
    return IntervalMesh(n, 0.0, 1.0)


def RectangleMesh(p0, p1, nx, ny, diagonal="right"):
    p0, p1 = _pt(p0), _pt(p1)
    gx = p0[0] + (p1[0] - p0[0]) * np.arange(nx + 1) / nx
    gy = p0[1] + (p1[1] - p0[1]) * np.arange(ny + 1) / ny
    pts = np.empty(((nx + 1) * (ny + 1), 2))
    pts[:, 0] = np.tile(gx, ny + 1)
    pts[:, 1] = np.repeat(gy, nx + 1)
    base = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).ravel()
    a, b, c, d = base, base + 1, base + nx + 1, base + nx + 2
    if diagonal == "right":
        tri = np.stack([a, b, d, a, c, d], axis=1).reshape(-1, 3)
    elif diagonal == "left":
        tri = np.stack([a, b, c, b, c, d], axis=1).reshape(-1, 3)
    elif diagonal == "crossed":
        mx, my = 0.5 * (gx[1:] + gx[:-1]), 0.5 * (gy[1:] + gy[:-1])
        mid = np.empty((nx * ny, 2))
        mid[:, 0] = np.tile(mx, ny)
        mid[:, 1] = np.repeat(my, nx)
        m = (nx + 1) * (ny + 1) + np.arange(nx * ny)
        pts = np.vstack([pts, mid])
        tri = np.stack([a, b, m, a, c, m, b, d, m, c, d, m], axis=1).reshape(-1, 3)
    else:
        raise ValueError("unknown diagonal '%s'" % diagonal)
    return Mesh(pts, tri)


def UnitSquareMesh(nx, ny, diagonal="right"):
    return RectangleMesh((0.0, 0.0), (1.0, 1.0), nx, ny, diagonal)


def BoxMesh(p0, p1, nx, ny, nz):
    p0, p1 = _pt(p0), _pt(p1)
    gx = p0[0] + (p1[0] - p0[0]) * np.arange(nx + 1) / nx
    gy = p0[1] + (p1[1] - p0[1]) * np.arange(ny + 1) / ny
    gz = p0[2] + (p1[2] - p0[2]) * np.arange(nz + 1) / nz
    npl = (nx + 1) * (ny + 1)
    pts = np.empty((npl * (nz + 1), 3))
    pts[:, 0] = np.tile(gx, (ny + 1) * (nz + 1))
    pts[:, 1] = np.tile(np.repeat(gy, nx + 1), nz + 1)
    pts[:, 2] = np.repeat(gz, npl)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (k * npl + j * (nx + 1) + i).ravel()
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    v4, v5, v6, v7 = v0 + npl, v1 + npl, v2 + npl, v3 + npl
    tet = np.stack(
        [v0, v1, v3, v7, v0, v1, v7, v5, v0, v5, v7, v4, v0, v3, v2, v7, v0, v6, v4, v7, v0, v2, v6, v7], axis=1
    ).reshape(-1, 4)
    return Mesh(pts, tet)


def UnitCubeMesh(nx, ny, nz):
    return BoxMesh((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), nx, ny, nz)


# ------------------------------------------------------------------------------- elements
def reference_nodes(tdim, degree):
    """Lagrange node coordinates on the reference simplex, local order = vertices then edges."""
    verts = np.vstack([np.zeros((1, tdim)), np.eye(tdim)])
    if degree == 1:
        return verts
    if tdim == 1:
        edges = ((0, 1),)
    else:
        edges = _TRI_EDGES if tdim == 2 else _TET_EDGES
    mids = np.array([(verts[i] + verts[j]) / 2 for i, j in edges])
    return np.vstack([verts, mids])


def tabulate_lagrange(tdim, degree, pts):
    """(phi [nq, nd], dphi [nq, nd, tdim]) on the reference simplex; degree 1 or 2."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, tdim)
    nq = len(pts)
    L = np.empty((nq, tdim + 1))
    L[:, 0] = 1.0 - pts.sum(axis=1)
    L[:, 1:] = pts
    dL = np.vstack([-np.ones((1, tdim)), np.eye(tdim)])  # [tdim+1, tdim]
    if degree == 1:
        return L, np.repeat(dL[None], nq, axis=0)
    if degree != 2:
        raise NotImplementedError("Lagrange degree %d (P1 and P2 are supported)" % degree)
    edges = ((0, 1),) if tdim == 1 else (_TRI_EDGES if tdim == 2 else _TET_EDGES)
    nd = tdim + 1 + len(edges)
    phi = np.empty((nq, nd))
    dphi = np.empty((nq, nd, tdim))
    phi[:, : tdim + 1] = L * (2.0 * L - 1.0)
    dphi[:, : tdim + 1, :] = (4.0 * L - 1.0)[:, :, None] * dL[None, :, :]
    for e, (i, j) in enumerate(edges):
        phi[:, tdim + 1 + e] = 4.0 * L[:, i] * L[:, j]
        dphi[:, tdim + 1 + e, :] = 4.0 * (L[:, [i]] * dL[j][None, :] + L[:, [j]] * dL[i][None, :])
    return phi, dphi


def simplex_quadrature(tdim, degree):
    """Collapsed Gauss-Jacobi rule exact for polynomials of total degree <= degree.
    Weights sum to the reference volume 1/tdim!."""
    m = max(1, (int(degree) + 2) // 2)
    x0, w0 = roots_jacobi(m, 0.0, 0.0)
    x0, w0 = (x0 + 1.0) / 2.0, w0 / 2.0
    if tdim == 1:
        return x0[:, None], w0
    x1, w1 = roots_jacobi(m, 1.0, 0.0)
    x1, w1 = (x1 + 1.0) / 2.0, w1 / 4.0
    if tdim == 2:
        # x = x1, y = x0 * (1 - x1)
        X = np.repeat(x1, m)
        Y = np.tile(x0, m) * (1.0 - X)
        return np.column_stack([X, Y]), np.repeat(w1, m) * np.tile(w0, m)
    x2, w2 = roots_jacobi(m, 2.0, 0.0)
    x2, w2 = (x2 + 1.0) / 2.0, w2 / 8.0
    A, B, C = np.meshgrid(x2, x1, x0, indexing="ij")
    WA, WB, WC = np.meshgrid(w2, w1, w0, indexing="ij")
    X = A
    Y = B * (1.0 - A)
    Z = C * (1.0 - A) * (1.0 - B)
    return np.column_stack([X.ravel(), Y.ravel(), Z.ravel()]), (WA * WB * WC).ravel()


def lagrange_interval_basis(p, xi):
    """Equispaced degree-p Lagrange basis on [0, 1] (p = 0: the constant 1 at the midpoint)."""
    xi = np.asarray(xi, dtype=np.float64).ravel()
    if p == 0:
        return np.array([0.5]), np.ones((xi.size, 1))
    t = np.linspace(0.0, 1.0, p + 1)
    B = np.ones((xi.size, p + 1))
    for a in range(p + 1):
        others = np.delete(t, a)
        B[:, a] = np.prod((xi[:, None] - others[None, :]) / (t[a] - others)[None, :], axis=1)
    return t, B


# ------------------------------------------------------------------------------- spaces
class _Str:
    def __init__(self, s):
        self._s = s

    def __str__(self):
        return self._s

    __repr__ = __str__


class _UFLElement:
    def __init__(self, space):
        self._space = space

    def __str__(self):
        s = self._space
        cell = {1: "interval", 2: "triangle", 3: "tetrahedron"}[s.mesh().tdim]
        base = "<CG%d on a %s>" % (s.degree, cell)
        if s.bs > 1:
            return "<vector element with %d components of %s>" % (s.bs, base)
        return base

    def degree(self):
        return self._space.degree

    def family(self):
        return "Lagrange"

    def value_shape(self):
        return () if self._space.bs == 1 else (self._space.bs,)


class _UFLSpace:
    def __init__(self, space):
        self._space = space

    def ufl_element(self):
        return _UFLElement(self._space)


class FunctionSpace:
    """Lagrange P1/P2 space (scalar or node-blocked vector) on a simplicial Mesh."""

    _next_id = 0

    def __init__(self, mesh, family="P", degree=1, bs=1):
        if str(family) not in ("P", "CG", "Lagrange"):
            raise NotImplementedError("element family '%s' (only Lagrange P/CG)" % family)
        if degree not in (1, 2):
            raise NotImplementedError("Lagrange degree %s (P1 and P2 are supported)" % degree)
        self._mesh, self.degree, self.bs = mesh, int(degree), int(bs)
        self.id = FunctionSpace._next_id
        FunctionSpace._next_id += 1
        cells, X = mesh.cells().astype(np.int64), mesh.coordinates()
        nv, nc, t = mesh.num_vertices(), mesh.num_cells(), mesh.tdim
        if t == 1:
            pos = np.empty(nv, dtype=np.int64)
            pos[np.argsort(X[:, 0], kind="stable")] = np.arange(nv)
            if self.degree == 1:
                self.vertex_to_node = (nv - 1) - pos
                self.cell_nodes = self.vertex_to_node[cells]
                nn = nv
            else:
                nn = 2 * nc + 1
                self.vertex_to_node = (nn - 1) - 2 * pos
                left = np.minimum(pos[cells[:, 0]], pos[cells[:, 1]])
                self.cell_nodes = np.column_stack([self.vertex_to_node[cells], (nn - 2) - 2 * left])
            ref = reference_nodes(1, self.degree)
            lam = np.column_stack([1 - ref[:, 0], ref[:, 0]])
            self.node_coords = np.empty((nn, 1))
            self.node_coords[self.cell_nodes.ravel()] = np.einsum("la,cag->clg", lam, X[cells]).reshape(-1, 1)
        else:
            self.vertex_to_node = np.arange(nv, dtype=np.int64)
            if self.degree == 1:
                self.cell_nodes = cells.copy()
                self.node_coords = X.copy()
            else:
                edges = _TRI_EDGES if t == 2 else _TET_EDGES
                # vectorised "first appearance" numbering
                pairs = np.stack([np.stack([cells[:, i], cells[:, j]], axis=1) for i, j in edges], axis=1)
                pairs = np.sort(pairs, axis=2).reshape(-1, 2)
                code = pairs[:, 0] * nv + pairs[:, 1]
                uniq, first, inverse = np.unique(code, return_index=True, return_inverse=True)
                rank = np.empty(len(uniq), dtype=np.int64)
                rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
                enodes = rank[inverse].reshape(nc, len(edges))
                ecoords = np.empty((len(uniq), mesh.gdim))
                ecoords[rank] = 0.5 * (X[uniq // nv] + X[uniq % nv])
                self.cell_nodes = np.column_stack([cells, nv + enodes])
                self.node_coords = np.vstack([X, ecoords])
        self.n_nodes = self.node_coords.shape[0]
        self.nd = self.cell_nodes.shape[1]
        self._finish()

    @classmethod
    def from_arrays(cls, mesh, cell_nodes, node_coords, degree=1, bs=1, vertex_to_node=None):
        """Build a space from an externally supplied dofmap (e.g. exported from real DOLFIN)."""
        self = cls.__new__(cls)
        self._mesh, self.degree, self.bs = mesh, int(degree), int(bs)
        self.id = FunctionSpace._next_id
        FunctionSpace._next_id += 1
        self.cell_nodes = np.asarray(cell_nodes, dtype=np.int64)
        self.node_coords = np.asarray(node_coords, dtype=np.float64).reshape(len(node_coords), -1)
        self.n_nodes, self.nd = self.node_coords.shape[0], self.cell_nodes.shape[1]
        if vertex_to_node is None:
            v2n = np.empty(mesh.num_vertices(), dtype=np.int64)
            v2n[mesh.cells().ravel()] = self.cell_nodes[:, : mesh.tdim + 1].ravel()
            vertex_to_node = v2n
        self.vertex_to_node = np.asarray(vertex_to_node, dtype=np.int64)
        self._finish()
        return self

    def _finish(self):
        bs = self.bs
        self.n_dofs = self.n_nodes * bs
        self.ndl = self.nd * bs
        self.cell_dofs = np.ascontiguousarray(
            (self.cell_nodes[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(len(self.cell_nodes), self.ndl),
            dtype=np.int32,
        )
        self._dev = {}

    # --- dolfin-like surface
    def mesh(self):
        return self._mesh

    def dim(self):
        return self.n_dofs

    def ufl_element(self):
        return _UFLElement(self)

    def ufl_function_space(self):
        return _UFLSpace(self)

    def tabulate_dof_coordinates(self):
        return np.repeat(self.node_coords, self.bs, axis=0)

    def num_sub_spaces(self):
        return 0 if self.bs == 1 else self.bs

    # --- band ordering of the dofs (1-D: by coordinate) for the banded direct solver
    def band_permutation(self):
        if "perm" not in self._dev:
            if self._mesh.tdim != 1:
                raise NotImplementedError("band ordering only for 1-D spaces")
            order = np.argsort(self.node_coords[:, 0], kind="stable")
            perm = (order[:, None] * self.bs + np.arange(self.bs)[None, :]).ravel()
            bw = self.degree * self.bs + (self.bs - 1)
            self._dev["perm"] = (perm.astype(np.int32), int(bw))
        return self._dev["perm"]

    def node_on_boundary(self):
        if "onb" not in self._dev:
            m = self._mesh
            cell, loc = m.boundary_facets()
            flag = np.zeros(self.n_nodes, dtype=bool)
            fn = self.facet_nodes(cell, loc)
            flag[fn.ravel()] = True
            self._dev["onb"] = flag
        return self._dev["onb"]

    def facet_nodes(self, cell, loc):
        """Nodes of facets (cell, local facet) in facet-local Lagrange order (vertices, then edges)."""
        t = self._mesh.tdim
        vloc = np.array([[k for k in range(t + 1) if k != i] for i in range(t + 1)])[loc]  # [nf, t]
        out = [np.take_along_axis(self.cell_nodes[cell], vloc, axis=1)]
        if self.degree == 2 and t >= 2:
            edges = _TRI_EDGES if t == 2 else _TET_EDGES
            sub = ((0, 1),) if t == 2 else _TRI_EDGES
            cols = []
            for (i, j) in sub:
                a, b = vloc[:, i], vloc[:, j]
                lo, hi = np.minimum(a, b), np.maximum(a, b)
                eidx = np.array([edges.index((int(l), int(h))) for l, h in zip(lo, hi)])
                cols.append(self.cell_nodes[cell, t + 1 + eidx])
            out.append(np.column_stack(cols))
        return np.concatenate(out, axis=1)


def VectorFunctionSpace(mesh, family="P", degree=1, dim=None):
    return FunctionSpace(mesh, family, degree, bs=dim or mesh.gdim)
