"""Element partition of a spatial FunctionSpace over the GPUs of one box (north_star: "the spatial mesh is
partitioned by element across the 8 GPUs of one box; NCCL over NVLink is used only for PCG halo exchange and
dot-product allreduces; the small 1-D parameter / time dimensions stay replicated").

The reference is serial, so this layer has no counterpart there; it sits entirely below the
reference API: a user script builds its global mesh, spaces, boundary conditions and callbacks exactly as on one
GPU, under ``torchrun`` with one process per GPU.  What changes is where the dof data of a LARGE 2-D/3-D space lives:

* nodes are split into ``world`` contiguous ranges (with the mesh-ordered numbering of the built-in meshes these
  are slabs); rank r keeps the cells that touch one of its nodes (its slab plus one halo layer of elements) and
  numbers their nodes ``[owned | ghost]`` (ghosts ascending by global id, i.e. grouped by owner rank);
* the local sub-mesh with that numbering is an ordinary ``FunctionSpace`` (``SpaceShard.local``), so pattern,
  atoms, loads, Dirichlet elimination and panels are built by the unchanged single-GPU code on 1/world of the
  cells; rows of owned dofs are complete (all their cells are local), rows of ghost dofs are partial and never used;
* every dof vector of the space is a device tensor of ``n_local`` entries whose ghost entries are kept equal to the
  owners' values (element-wise operations preserve that; a solve ends with one halo exchange);
* mode integrals are local sums over the owned rows followed by ONE batched all-reduce of the step's scalar pool
  slice (``forms._launch``); the spatial solve is the sharded PCG (``partition.sharded_pcg`` /
  ``pgd_pcg_persist_sync``: halo of the direction vector + dot products per iteration);
* host-side accessors (``f.vector()[:]``, ``compute_vertex_values``, ``interpolate``, Dirichlet dof sets) stay
  GLOBAL: reads gather the owned slices, writes scatter, so results look exactly as on one GPU.

Nothing here launches kernels; index plumbing only (NumPy on the host at set-up, torch for the collectives).
"""
import numpy as np
import torch

from .fem import FunctionSpace, Mesh
from .partition import HaloPlan, RowPartition

_policy = {"mode": "auto", "min_dofs": 200000, "group": None}


def configure(mode=None, min_dofs=None, group=None):
    """mode: "auto" (shard 2-D/3-D spaces with >= min_dofs dofs when more than one rank runs), True (shard every
    2-D/3-D space), False (never).  Must be identical on all ranks and set before the first space is used."""
    if mode is not None:
        _policy["mode"] = mode
    if min_dofs is not None:
        _policy["min_dofs"] = int(min_dofs)
    _policy["group"] = group
    return dict(_policy)


def _world():
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return None
    g = _policy["group"]
    w = dist.get_world_size(g)
    return (dist.get_rank(g), w, g) if w > 1 else None


def shard_of(V):
    """The SpaceShard of a FunctionSpace, or None when the space is replicated (decided once per space)."""
    ent = V._dev.get("shard", 0)
    if ent != 0:
        return ent
    sh = None
    w = _world()
    mode = _policy["mode"]
    if w is not None and mode is not False and V.mesh().tdim >= 2:
        if mode is True or V.n_dofs >= _policy["min_dofs"]:
            if V.n_nodes >= 4 * w[1]:
                sh = SpaceShard(V, *w)
    V._dev["shard"] = sh
    return sh


class SpaceShard:
    def __init__(self, V, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.space_id = V.id  # creation counter of the space: identical on every rank (same script)
        bs = V.bs
        self.bs = bs
        self.part = RowPartition(V.n_dofs, world, bs)  # contiguous node ranges, expressed in dofs
        nb = self.part.bounds // bs  # node bounds
        g0, g1 = int(nb[rank]), int(nb[rank + 1])
        cn = V.cell_nodes
        mask = ((cn >= g0) & (cn < g1)).any(axis=1)
        self.cell_ids = np.nonzero(mask)[0]  # global ids of the local cells (ascending)
        lc = cn[self.cell_ids]
        uniq = np.unique(lc)
        ghost = uniq[(uniq < g0) | (uniq >= g1)]  # ascending => grouped by owner rank
        n_on, n_gn = g1 - g0, len(ghost)
        self.node_range = (g0, g1)
        self.ghost_nodes = ghost
        self.l2g_nodes = np.concatenate([np.arange(g0, g1, dtype=np.int64), ghost.astype(np.int64)])
        g2l = np.full(V.n_nodes, -1, dtype=np.int64)
        g2l[g0:g1] = np.arange(n_on)
        g2l[ghost] = n_on + np.arange(n_gn)
        self._g2l_nodes = g2l
        self.n_owned, self.n_ghost = n_on * bs, n_gn * bs
        self.n_local = self.n_owned + self.n_ghost
        self.n_global = V.n_dofs
        self.l2g_dofs = (self.l2g_nodes[:, None] * bs + np.arange(bs)[None, :]).ravel()
        # ---- the local sub-mesh: vertices ascending by global id (cell orientation as in the global mesh), nodes
        # in [owned | ghost] order
        m = V.mesh()
        gcells = m.cells()[self.cell_ids].astype(np.int64)
        gv = np.unique(gcells)
        vmap = np.full(m.num_vertices(), -1, dtype=np.int64)
        vmap[gv] = np.arange(len(gv))
        lmesh = Mesh(m.coordinates()[gv], vmap[gcells])
        v2n = g2l[V.vertex_to_node[gv]]
        self.local = FunctionSpace.from_arrays(lmesh, g2l[lc], V.node_coords[self.l2g_nodes], degree=V.degree, bs=bs,
                                               vertex_to_node=v2n)
        self._cell_g2l = None
        self._n_cells_global = len(cn)
        # ---- halo plan, computed from the replicated global mesh (no communication): the nodes this rank owns that
        # are ghosts elsewhere, per destination in ascending global order (= the destination's ghost order)
        bounds_hi = nb[1:]
        own = np.searchsorted(bounds_hi, lc, side="right")  # owner rank of every node of every local cell
        mixed = (own != own[:, :1]).any(axis=1)
        lcm, ownm = lc[mixed], own[mixed]
        pairs = []
        nd = lcm.shape[1]
        for a in range(nd):
            mine = ownm[:, a] == rank
            if not mine.any():
                continue
            for b in range(nd):
                if a == b:
                    continue
                sel = mine & (ownm[:, b] != rank)
                if sel.any():
                    pairs.append(ownm[sel, b].astype(np.int64) * V.n_nodes + lcm[sel, a])
        key = np.unique(np.concatenate(pairs)) if pairs else np.zeros(0, dtype=np.int64)
        dest, snode = key // V.n_nodes, key % V.n_nodes
        send_counts = np.bincount(dest, minlength=world) * bs
        recv_counts = np.bincount(np.searchsorted(bounds_hi, ghost, side="right"), minlength=world) * bs
        self.send_idx_host = ((snode - g0)[:, None] * bs + np.arange(bs)[None, :]).ravel().astype(np.int64)
        self.send_counts = [int(c) for c in send_counts]
        self.recv_counts = [int(c) for c in recv_counts]
        self._halo = None
        self._l2g_dev = None

    # ------------------------------------------------------------------ index maps
    def g2l_dofs(self, dofs):
        """local index of global dofs (-1 where the dof is neither owned nor a ghost here)"""
        dofs = np.asarray(dofs, dtype=np.int64)
        ln = self._g2l_nodes[dofs // self.bs]
        return np.where(ln >= 0, ln * self.bs + dofs % self.bs, -1)

    def local_cells(self, cells):
        """global cell ids -> local cell ids (-1 for cells that are not kept on this rank)"""
        if self._cell_g2l is None:
            m = np.full(self._n_cells_global, -1, dtype=np.int64)
            m[self.cell_ids] = np.arange(len(self.cell_ids))
            self._cell_g2l = m
        return self._cell_g2l[np.asarray(cells, dtype=np.int64)]

    # ------------------------------------------------------------------ device side
    def halo(self, device):
        if self._halo is None:
            gg = torch.as_tensor(self.l2g_dofs[self.n_owned:], dtype=torch.int64, device=device)
            self._halo = HaloPlan(self.n_owned, gg, torch.as_tensor(self.send_idx_host, dtype=torch.int64, device=device),
                                  self.send_counts, self.recv_counts, self.group)
        return self._halo

    def update_ghosts(self, t):
        """ghost entries of the local vector t <- the owners' values (in place; collective)"""
        return self.halo(t.device).exchange(t)

    def allreduce(self, t):
        """in-place sum over the ranks (collective); identical result on every rank"""
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def scatter(self, a, device):
        """global host array -> local device tensor [n_local] (owned + ghost entries)"""
        from . import _lib

        return _lib.to_device(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel()[self.l2g_dofs]))

    def gather_host(self, t):
        """local device tensor -> global host array (collective: every rank receives the full vector)"""
        import torch.distributed as dist

        from . import _lib

        sizes = [int(self.part.bounds[r + 1] - self.part.bounds[r]) for r in range(self.world)]
        m = max(sizes)
        mine = torch.zeros(m, dtype=t.dtype, device=t.device)
        mine[: self.n_owned] = t[: self.n_owned]
        outs = [torch.empty(m, dtype=t.dtype, device=t.device) for _ in sizes]
        dist.all_gather(outs, mine, group=self.group)
        return _lib.to_host(torch.cat([o[:s] for o, s in zip(outs, sizes)]))
