"""Device-resident data of a FunctionSpace and the separated-form *atoms* assembled on it.

An atom is one constant-coefficient bilinear form  int w(x) sum T[iv,jv,iu,ju] D_jv v_iv D_ju u_iu
(slot 0 = value, 1+m = d/dx_m) assembled ONCE into the space's fixed CSR pattern; the per-mode
operator of the fixed-point sweep is then  A_d = sum_k c_k K_{d,k}  (``_lib.lincomb`` over the value
arrays, or the fused P1 kernel ``_lib.assemble_p1``), cf. SURVEY.md 7.1 and
pgdrome/solver.py:547-556.  All arithmetic happens in libpgdb200.so; this module only stages the
host tables (basis tabulation, quadrature, coefficient samples) that FFC would have generated.
"""
import weakref

import numpy as np
import torch

from . import _lib
from .fem import lagrange_interval_basis, reference_nodes, simplex_quadrature, tabulate_lagrange


def _dev():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _up(a, dtype):
    return _lib.to_device(a, dtype)


class _Pattern:
    """(rowptr, colidx, gptr, gidx) of a space.  The gather lists of the generic element-matrix route (gptr[nnz + 1],
    gidx[n_cells * ndl^2]: 1.4 GB at configs[2]) are built on first use: the P1 routes never ask for them."""

    __slots__ = ("rowptr", "colidx", "_gather", "_make")

    def __init__(self, rowptr, colidx, gather=None, make=None):
        self.rowptr, self.colidx, self._gather, self._make = rowptr, colidx, gather, make

    def __len__(self):
        return 4

    def __getitem__(self, i):
        if i == 0:
            return self.rowptr
        if i == 1:
            return self.colidx
        if i in (2, 3):
            if self._gather is None:
                self._gather = self._make()
                self._make = None
            return self._gather[i - 2]
        raise IndexError(i)

    def __iter__(self):
        return (self[i] for i in range(4))


class DeviceSpace:
    """GPU mirror of a FunctionSpace: mesh arrays, CSR pattern, gather lists, cached atoms."""

    def __init__(self, space):
        self._space = weakref.ref(space)  # the space owns this object (space._dev): no strong back-reference, so
        m = space.mesh()                  # dropping the space frees the device arrays by reference count alone
        self.coords = _up(m.coordinates(), torch.float64)
        self.cell_verts = _up(m.cells(), torch.int32)
        self.cell_dofs = _up(space.cell_dofs, torch.int32)
        self.n_dofs, self.ndl = space.n_dofs, space.ndl
        self._pattern = None
        self._vecmap = None
        self._rowplan = None
        self._tabs = {}
        self._facet = {}
        self.atoms = {}  # key -> values tensor
        self.band = None
        self.shard = None        # sharding.SpaceShard on an element-partitioned space (ShardedDeviceSpace)
        self.n_owned = self.n_dofs
        self._bsr = 0

    @property
    def space(self):
        return self._space()

    # ---- structure
    @property
    def pattern(self):
        if self._pattern is None:
            if self.space.bs > 1:
                self._pattern = self._blocked_pattern()
            else:
                rowptr, colidx, gptr, gidx = _lib.pattern_build(self.cell_dofs, self.n_dofs)
                self._pattern = _Pattern(rowptr, colidx, gather=(gptr, gidx))
        return self._pattern

    def _blocked_pattern(self):
        """Pattern of a vector space from the pattern of its NODES: the dofs of a space are node-blocked (dof = node * bs +
        component, fem.FunctionSpace._finish), so row (I, a) holds the columns (J, b) for every neighbour J of I and every
        component b -- bs^2 times fewer keys to sort than the dof-level build (30 M instead of 271 M at configs[2]); the
        node-level arrays are at the same time the block-column list of the node-block SpMV walk and the row-owner plan
        of the fused P1 assembly.  Index arithmetic on device tensors only (set-up plumbing); identical to the dof-level
        pattern bit for bit."""
        s = self.space
        bs, n = s.bs, self.n_dofs
        cn = _up(np.ascontiguousarray(s.cell_nodes, dtype=np.int32), torch.int32)
        brp, bci, bgptr, bgidx = _lib.pattern_build(cn, s.n_nodes)
        self._node_pattern = (brp, bci, cn)
        dev = brp.device
        blocks = (brp[1:] - brp[:-1]).to(torch.int64)                    # node blocks per node row
        rowlen = (blocks * bs).repeat_interleave(bs)                     # entries per dof row
        rowptr64 = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(rowlen, 0, out=rowptr64[1:])
        nnz = int(rowptr64[-1].item())
        if nnz >= 2**31:
            raise ValueError("pattern with more than 2^31 entries")
        row_of = torch.repeat_interleave(torch.arange(n, device=dev), rowlen)
        off = torch.arange(nnz, device=dev) - rowptr64[row_of]
        t = brp.to(torch.int64)[torch.div(row_of, bs, rounding_mode="floor")] + torch.div(off, bs, rounding_mode="floor")
        colidx = (bci.to(torch.int64)[t] * bs + off % bs).to(torch.int32)
        rowptr = rowptr64.to(torch.int32)
        del row_of, off, t

        bs_, nd_, nc_ = s.bs, s.nd, int(s.cell_nodes.shape[0])  # (no reference to self in the closure: no cycle)

        def make_gather():
            return DeviceSpace._blocked_gather(bs_, nd_, nc_, brp, bgptr, bgidx, rowptr64, nnz)

        return _Pattern(rowptr, colidx, make=make_gather)

    @staticmethod
    def _blocked_gather(bs, nd, n_cells, brp, bgptr, bgidx, rowptr64, nnz, chunk=1 << 22):
        """gather lists of the generic assembler on a vector space, from the node-level lists: the contribution
        (cell e, local nodes a, b) of node entry (I, J) is the contribution (e, a*bs+i, b*bs+j) of the dof entry
        ((I, i), (J, j)); ascending inside every group like the node-level list.  Built in chunks of entries."""
        ndl = nd * bs
        dev = brp.device
        brp64, bgptr64 = brp.to(torch.int64), bgptr.to(torch.int64)
        gptr = torch.zeros(nnz + 1, dtype=torch.int64, device=dev)
        total = n_cells * ndl * ndl
        gidx = torch.empty(total, dtype=torch.int32, device=dev)
        base = 0
        for k0 in range(0, nnz, chunk):
            k1 = min(nnz, k0 + chunk)
            k = torch.arange(k0, k1, device=dev)
            r = torch.searchsorted(rowptr64, k, right=True) - 1
            off = k - rowptr64[r]
            ci, cj = r % bs, off % bs
            t = brp64[torch.div(r, bs, rounding_mode="floor")] + torch.div(off, bs, rounding_mode="floor")
            cnt = bgptr64[t + 1] - bgptr64[t]
            csum = torch.cumsum(cnt, 0)
            gptr[k0 + 1:k1 + 1] = base + csum
            m = int(csum[-1].item())
            kk = torch.repeat_interleave(torch.arange(k1 - k0, device=dev), cnt)
            pos = torch.arange(m, device=dev) - (csum - cnt)[kk]
            c = bgidx[bgptr64[t[kk]] + pos].to(torch.int64)
            e = torch.div(c, nd * nd, rounding_mode="floor")
            ab = c % (nd * nd)
            a_, b_ = torch.div(ab, nd, rounding_mode="floor"), ab % nd
            gidx[base:base + m] = (e * (ndl * ndl) + (a_ * bs + ci[kk]) * ndl + (b_ * bs + cj[kk])).to(torch.int32)
            base += m
        assert base == total, (base, total)
        return gptr, gidx

    @property
    def nnz(self):
        return self.pattern[1].numel()

    @property
    def rowptr_owned(self):
        """row pointers of the rows this rank owns (all rows on a replicated space): what SpMV-type kernels iterate"""
        rp = self.pattern[0]
        return rp if self.n_owned == self.n_dofs else rp[: self.n_owned + 1]

    @property
    def bsr(self):
        """block-column plan of the node-block walk (vector spaces; None when the pattern is not block structured)"""
        if self._bsr == 0:
            s = self.space
            self._bsr = None
            if 1 < s.bs <= 3:
                self.pattern  # noqa: B018  (builds the node pattern)
                brp, bci, _ = self._node_pattern
                n_on = self.n_owned // s.bs  # owned node rows come first
                nb = brp[1:n_on + 1] - brp[:n_on]
                self._bsr = (bci[: int(brp[n_on].item())], int(nb.max().item()))
        return self._bsr

    @property
    def nnz_owned(self):
        if self.n_owned == self.n_dofs:
            return self.nnz
        if getattr(self, "_nnz_owned", None) is None:
            self._nnz_owned = int(self.pattern[0][self.n_owned].item())
        return self._nnz_owned

    @property
    def lpr(self):
        """lanes per CSR row for the SpMV-type kernels, from the mean row length."""
        m = self.nnz / max(1, self.n_dofs)
        for l in (2, 4, 8, 16):
            if m <= 1.5 * l:
                return l
        return 32

    @property
    def vecmap(self):
        if self._vecmap is None:
            self._vecmap = _lib.vecmap_build(self.cell_dofs, self.n_dofs)
        return self._vecmap

    @property
    def rowplan(self):
        """Plan of the row-owner fused P1 kernel (scalar P1 spaces only; False if a row is too long)."""
        if self._rowplan is None:
            rowptr, colidx = self.pattern[0], self.pattern[1]
            vptr, vidx = self.vecmap
            try:
                self._rowplan = _lib.p1_rowplan_build(rowptr, colidx, self.cell_dofs, vptr, vidx, self.n_dofs)
            except _lib.PGDB200Error:
                self._rowplan = False
        return self._rowplan

    @property
    def coords_soa(self):
        """component-major copy of the vertex coordinates for the row-owner assembly kernel"""
        if getattr(self, "_coords_soa", None) is None:
            self._coords_soa = self.coords.reshape(-1, self.space.mesh().gdim).t().contiguous()
        return self._coords_soa

    @property
    def node_plan(self):
        """(node_vptr, node_vent) of the row-owner plan on the NODE-level pattern (block rows / block columns) for the
        general-tensor fused P1 kernel, or False (not an affine P1 space, or a node row longer than 255 blocks)."""
        if getattr(self, "_node_plan", None) is None:
            s = self.space
            self._node_plan = False
            if s.degree == 1 and s.mesh().tdim == s.mesh().gdim:
                try:
                    if s.bs == 1:
                        if self.rowplan is not False:
                            self._node_plan = (self.vecmap[0], self.rowplan)
                    else:
                        self.pattern  # noqa: B018  (builds the node pattern)
                        brp, bci, cn = self._node_pattern
                        vptr, vidx = _lib.vecmap_build(cn, s.n_nodes)
                        vent = _lib.p1_rowplan_build(brp, bci, cn, vptr, vidx, s.n_nodes)
                        self._node_plan = (vptr, vent)
                except _lib.PGDB200Error:
                    self._node_plan = False
        return self._node_plan

    def _cell_weight(self, weights, weight, wdeg):
        """per-cell coefficient (device, [n_cells]) when every weight is a degree-0 Expression / callable; None if there
        is no weight; False if some weight varies inside the cells"""
        specs = list(weights or [])
        if weight is not None:
            specs.append(("callable", weight, None, wdeg, None))
        if not specs:
            return None
        if any(kind == "fn" or d != 0 for kind, _, _, d, _ in specs):
            return False
        # (id, component, version) keys of the Expressions: the same zone indicator is asked for by every atom that
        # carries it; entries keep their Expression alive so that an id cannot be recycled while it is in the table
        ckey = tuple(k for _, _, _, _, k in specs) if all(k is not None for _, _, _, _, k in specs) else None
        cache = self.__dict__.setdefault("_cell_w", {})
        if ckey is not None and ckey in cache:
            return cache[ckey][0]
        if getattr(self, "_centroids", None) is None:
            m = self.space.mesh()
            self._centroids = m.coordinates()[m.cells()].mean(axis=1)
        w = None
        for kind, obj, comp, d, _ in specs:
            fn = obj if kind == "callable" else (lambda X, o=obj, c=comp: o.eval_np(X, c))
            v = np.asarray(fn(self._centroids), dtype=np.float64).reshape(-1)
            w = v if w is None else w * v
        out = _up(np.ascontiguousarray(w), torch.float64)
        if ckey is not None:
            cache[ckey] = (out, [obj for _, obj, _, _, _ in specs])
        return out

    def _p1_closed_form(self, T):
        """(c_mass, c_stiff, c_adv) if T is  c_m u v + c_k grad u.grad v + sum_m c_adv[m] (d_m u) v, else None."""
        s = self.space
        g = s.mesh().gdim
        if s.bs != 1 or s.degree != 1 or s.mesh().tdim != g:
            return None
        t = T[0, :, 0, :]
        if np.any(t[1:, 0] != 0):  # (d_m v) u terms are not in the closed form
            return None
        ck = t[1, 1]
        if np.any(t[1:, 1:] != ck * np.eye(g)):
            return None
        return float(t[0, 0]), float(ck), [float(v) for v in t[0, 1:]]

    def tables(self, qdeg, tdim=None, degree=None):
        tdim = self.space.mesh().tdim if tdim is None else tdim
        degree = self.space.degree if degree is None else degree
        key = (tdim, degree, int(qdeg))
        if key not in self._tabs:
            pts, w = simplex_quadrature(tdim, qdeg)
            phi, dphi = tabulate_lagrange(tdim, degree, pts)
            self._tabs[key] = dict(pts=pts, w=w, phi=_up(phi, torch.float64), dphi=_up(dphi, torch.float64),
                                   qw=_up(w, torch.float64), nq=len(w))
        return self._tabs[key]

    # ---- coefficient sampling (host, set-up): per-cell P_p interpolant at the quadrature points
    def sample_weight(self, fn, wdeg, pts, cells=None, tdim=None):
        """fn(x[..., gdim]) -> values; returns float64 [n_cells, nq] (host)."""
        m = self.space.mesh()
        cells = m.cells() if cells is None else cells
        tdim = m.tdim if tdim is None else tdim
        X = m.coordinates()[cells]  # [e, tdim+1, g]
        pts = np.asarray(pts).reshape(-1, tdim)
        if wdeg == 0:
            xc = X.mean(axis=1)
            return np.repeat(np.asarray(fn(xc), dtype=np.float64).reshape(-1, 1), len(pts), axis=1)
        if tdim == 1:
            t, B = lagrange_interval_basis(wdeg, pts[:, 0])
            xn = X[:, 0, None, :] * (1 - t)[None, :, None] + X[:, 1, None, :] * t[None, :, None]
            return np.asarray(fn(xn), dtype=np.float64) @ B.T
        if wdeg > 2:
            raise NotImplementedError("coefficient degree > 2 on a 2-D/3-D mesh")
        ref = reference_nodes(tdim, wdeg)
        lam = np.column_stack([1 - ref.sum(axis=1), ref])
        xn = np.einsum("na,eag->eng", lam, X)
        phi, _ = tabulate_lagrange(tdim, wdeg, pts)
        return np.asarray(fn(xn), dtype=np.float64) @ phi.T

    # ---- weights: product of per-cell P_p interpolants at the quadrature points (host, set-up)
    def _weights_at(self, weights, weight, wdeg, pts, cells=None, tdim=None):
        """weights: list of specs ("expr"|"fn"|"callable", obj, comp, degree, key); returns
        (float64 [n_cells, nq] on the device or None, total polynomial degree)."""
        specs = list(weights or [])
        if weight is not None:
            specs.append(("callable", weight, None, wdeg, None))
        if not specs:
            return None, 0
        wq, deg = None, 0
        for kind, obj, comp, d, _ in specs:
            if kind == "fn":
                if cells is not None:
                    raise NotImplementedError("coefficient Function inside a boundary integral")
                w = self._sample_function(obj, comp, pts)
            else:
                fn = obj if kind == "callable" else (lambda X, o=obj, c=comp: o.eval_np(X, c))
                w = self.sample_weight(fn, d, pts, cells=cells, tdim=tdim)
            wq = w if wq is None else wq * w
            deg += d
        return _up(wq, torch.float64), deg

    def _sample_function(self, f, comp, pts):
        V = f.V
        phi, _ = tabulate_lagrange(V.mesh().tdim, V.degree, pts)
        vals = f.values_host().reshape(V.n_nodes, V.bs)[:, comp or 0]
        return vals[V.cell_nodes] @ phi.T

    def _wdeg(self, weights, weight, wdeg):
        return sum(s[3] for s in (weights or [])) + (wdeg if weight is not None else 0)

    # ---- atoms
    def assemble_bilinear(self, T, weight=None, wdeg=0, weights=None):
        """values [nnz] of the atom with form tensor T and optional coefficient weight(s)."""
        s = self.space
        m = s.mesh()
        g = m.gdim
        T = np.asarray(T, dtype=np.float64).reshape(s.bs, g + 1, s.bs, g + 1)
        if weight is None and not weights:
            cf = self._p1_closed_form(T)
            if cf is not None and self.rowplan is not False:
                # constant-coefficient P1 operator: one fused kernel straight into the CSR pattern
                rowptr = self.pattern[0]
                return _lib.assemble_p1_rows(self.coords, self.cell_verts, g, cf[0], cf[1], cf[2] if any(cf[2]) else None,
                                             rowptr, self.vecmap[0], self.rowplan, self.n_dofs, coords_soa=self.coords_soa,
                                             nnz=self.nnz)
        if s.degree == 1 and m.tdim == g:
            # any other constant-coefficient P1 atom (vector spaces, Voigt elasticity, mixed derivative terms, material
            # zones): the general-tensor row-owner kernel, still straight into the CSR pattern
            wc = self._cell_weight(weights, weight, wdeg)
            if wc is not False and self.node_plan is not False:
                return _lib.assemble_p1_tensor(g, s.bs, T, self.coords_soa, self.cell_verts, self.pattern[0], self.node_plan[0],
                                               self.node_plan[1], self.n_dofs, self.nnz, w_cell=wc)
        # polynomial degree of the integrand on an affine simplex
        dv = s.degree if np.any(T[:, 0, :, :] != 0) else s.degree - 1
        du = s.degree if np.any(T[:, :, :, 0] != 0) else s.degree - 1
        qdeg = max(dv + du + self._wdeg(weights, weight, wdeg), 1)
        tab = self.tables(qdeg)
        wq, _ = self._weights_at(weights, weight, wdeg, tab["pts"])
        Ae = _lib.elem_bilinear(self.coords, self.cell_verts, m.tdim, g, s.bs, s.nd, tab["phi"], tab["dphi"], tab["qw"], wq,
                                _up(T, torch.float64))
        rowptr, colidx, gptr, gidx = self.pattern
        return _lib.gather_values(Ae, gptr, gidx, colidx.numel())

    def assemble_linear(self, L, weight=None, wdeg=0, weights=None):
        s = self.space
        m = s.mesh()
        g = m.gdim
        L = np.asarray(L, dtype=np.float64).reshape(s.bs, g + 1)
        dv = s.degree if np.any(L[:, 0] != 0) else s.degree - 1
        qdeg = max(dv + self._wdeg(weights, weight, wdeg), 1)
        tab = self.tables(qdeg)
        wq, _ = self._weights_at(weights, weight, wdeg, tab["pts"])
        be = _lib.elem_linear(self.coords, self.cell_verts, m.tdim, g, s.bs, s.nd, tab["phi"], tab["dphi"], tab["qw"], wq,
                              _up(L, torch.float64))
        vptr, vidx = self.vecmap
        return _lib.gather_values(be, vptr, vidx, self.n_dofs)

    # ---- boundary-facet integrals (ds(id)): facets as embedded (tdim-1)-simplices
    def facet_set(self, key, cell, loc):
        if key not in self._facet:
            s = self.space
            m = s.mesh()
            fverts = m.facet_vertices(cell, loc).astype(np.int32)
            fnodes = s.facet_nodes(cell, loc)
            fdofs = (fnodes[:, :, None] * s.bs + np.arange(s.bs)[None, None, :]).reshape(len(fnodes), -1).astype(np.int32)
            d = dict(verts=fverts, verts_d=_up(fverts, torch.int32), n=len(fverts), nd=fnodes.shape[1])
            if len(fverts):
                d["dofs_d"] = _up(fdofs, torch.int32)
                d["vecmap"] = _lib.vecmap_build(d["dofs_d"], self.n_dofs)
            self._facet[key] = d
        return self._facet[key]

    def assemble_facet_linear(self, key, cell, loc, Lvec, weight=None, wdeg=0, weights=None):
        """b[dof] = int_facets w * sum_i Lvec[i] v_i ds  (value slots only)."""
        s = self.space
        m = s.mesh()
        fs = self.facet_set(key, cell, loc)
        out = torch.zeros(self.n_dofs, dtype=torch.float64, device=self.coords.device)
        if fs["n"] == 0:
            return out
        ft = m.tdim - 1
        if ft == 0:
            raise NotImplementedError("point 'facet' integrals on 1-D meshes")
        qdeg = max(s.degree + self._wdeg(weights, weight, wdeg), 1)
        tab = self.tables(qdeg, tdim=ft)
        wq, _ = self._weights_at(weights, weight, wdeg, tab["pts"], cells=fs["verts"], tdim=ft)
        L = np.zeros((s.bs, m.gdim + 1))
        L[:, 0] = np.asarray(Lvec, dtype=np.float64).reshape(s.bs)
        be = _lib.elem_linear(self.coords, fs["verts_d"], ft, m.gdim, s.bs, fs["nd"], tab["phi"], tab["dphi"], tab["qw"], wq,
                              _up(L, torch.float64))
        vptr, vidx = fs["vecmap"]
        return _lib.gather_values(be, vptr, vidx, self.n_dofs, out=out)


class ShardedDeviceSpace(DeviceSpace):
    """DeviceSpace of an element-partitioned space: the device arrays are those of this rank's sub-mesh (cells that
    touch an owned node, nodes numbered [owned | ghost], sharding.SpaceShard.local), so pattern, atoms and loads are
    built by the unchanged code above on 1/world of the cells.  Rows of owned dofs are complete; rows of ghost dofs
    are partial and never enter a product (``rowptr_owned`` ends at n_owned)."""

    def __init__(self, gspace, shard):
        super().__init__(shard.local)
        self._gspace = weakref.ref(gspace)
        self.shard = shard
        self.n_owned = shard.n_owned

    def facet_set(self, key, cell, loc):
        """(cell, local facet) pairs arrive in GLOBAL cell numbers (MeshFunction / boundary_facets of the user's mesh)"""
        if key not in self._facet:
            lc = self.shard.local_cells(cell)
            keep = lc >= 0
            return super().facet_set(key, lc[keep], np.asarray(loc)[keep])
        return self._facet[key]

    def _sample_function(self, f, comp, pts):
        from . import sharding

        V = f.V
        sh = sharding.shard_of(V)
        if sh is not self.shard:
            raise NotImplementedError("coefficient Function from a different space than the (partitioned) form space")
        phi, _ = tabulate_lagrange(V.mesh().tdim, V.degree, pts)
        loc = sh.local
        vals = _lib.to_host(f.tensor()).reshape(loc.n_nodes, loc.bs)[:, comp or 0]
        return vals[loc.cell_nodes] @ phi.T


def device_space(space):
    """Lazily attach (and cache) the DeviceSpace of a FunctionSpace."""
    ds = space._dev.get("device_space")
    if ds is None:
        from . import sharding

        sh = sharding.shard_of(space)
        ds = DeviceSpace(space) if sh is None else ShardedDeviceSpace(space, sh)
        space._dev["device_space"] = ds
    return ds
