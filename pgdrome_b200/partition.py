"""Row partition of a spatial operator over the GPUs of one box, halo plan and the sharded PCG.

The reference is serial (no MPI/NCCL anywhere); this module is the multi-GPU side of
the B200 path for the large spatial dimensions of BASELINE configs[2]-[4] (north_star: "the spatial
mesh is partitioned by element across the GPUs of one box; NCCL over NVLink is used only for PCG halo
exchange and dot-product allreduces").  One process per GPU (``torch.distributed``, backend nccl;
gloo on CPU for the host-logic tests).

* ``RowPartition``   contiguous, block-aligned dof ranges per rank.  With a mesh-ordered numbering
  (DOLFIN's built-in meshes, or any bandwidth-reducing ordering) this is a partition of the mesh into
  slabs of elements: rank r owns the rows of its nodes and needs the nodes of one element layer
  around its slab as ghosts.
* ``ShardedMatrix``  the owned rows of a CSR matrix with columns renumbered to [owned | ghost].
* ``HaloPlan``       which owned entries every neighbour needs (send lists) and where the received
  ghosts land (the tail of the local vector, grouped by source rank => no unpack kernel).
  ``exchange`` is one grouped NCCL send/recv (``all_to_all_single`` with split sizes).
* ``sharded_pcg``    Jacobi / node-block-Jacobi PCG over the partition: the per-rank kernels of
  ``pgd_spcg_*`` (libpgdb200) with one halo exchange and two scalar all-reduces per iteration; all
  scalars and the convergence flag stay on the device and are identical on every rank.

Only index plumbing lives here (torch ops on the device); all floating-point work of an iteration is
done by the CUDA kernels behind the C ABI.  The kernel set (``ops``) is injectable so that the
world_size-2 gloo tests can drive the same plan and driver logic with a NumPy stand-in.
"""
import numpy as np
import torch

F64, I32, I64 = torch.float64, torch.int32, torch.int64


class RowPartition:
    def __init__(self, n_rows, world, block=1):
        n_nodes = n_rows // block
        base, rem = divmod(n_nodes, world)
        counts = [base + (1 if r < rem else 0) for r in range(world)]
        self.bounds = np.concatenate([[0], np.cumsum(counts)]) * block
        self.world, self.n_rows, self.block = world, n_rows, block

    def range(self, rank):
        return int(self.bounds[rank]), int(self.bounds[rank + 1])

    def owner(self, rows):
        """rank owning each global row (torch int64 tensor in, tensor out)."""
        b = torch.as_tensor(self.bounds[1:], dtype=I64, device=rows.device)
        return torch.searchsorted(b, rows, right=True)


class HaloPlan:
    def __init__(self, n_owned, ghost_global, send_idx, send_counts, recv_counts, group=None):
        self.n_owned = n_owned
        self.ghost_global = ghost_global  # [n_ghost] global ids, ascending (=> grouped by owner rank)
        self.send_idx = send_idx  # [sum(send_counts)] local owned indices, grouped by destination rank
        self.send_counts, self.recv_counts = list(send_counts), list(recv_counts)
        self.n_ghost = int(ghost_global.numel())
        self.n_local = n_owned + self.n_ghost
        self.group = group
        self._buf = torch.empty(int(send_idx.numel()), dtype=F64, device=send_idx.device)

    def ensure_peer_layout(self):
        """Collective: where this rank's block of ghosts starts inside every neighbour's p
        (= n_owned_r + the ghosts r receives from lower ranks); needed by the NVLink peer-window path."""
        import torch.distributed as dist

        if getattr(self, "peer_ghost_base", None) is not None:
            return
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = self.send_idx.device
        mine = torch.tensor([self.n_owned] + [int(c) for c in self.recv_counts], dtype=I64, device=dev)
        allv = [torch.empty(world + 1, dtype=I64, device=dev) for _ in range(world)]
        dist.all_gather(allv, mine, group=self.group)
        base = []
        for r in range(world):
            v = allv[r].tolist()
            base.append(int(v[0] + sum(v[1:1 + rank])))
        self.peer_ghost_base = base

    @property
    def bytes_per_exchange(self):
        return 8 * (int(self.send_idx.numel()) + self.n_ghost)

    def exchange(self, v):
        """fill v[n_owned:] with the owners' current values (v: [n_local] float64, in place)."""
        import torch.distributed as dist

        if self.n_local == self.n_owned and self.send_idx.numel() == 0 and not dist.is_initialized():
            return v
        torch.index_select(v[: self.n_owned], 0, self.send_idx, out=self._buf)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_to_all_single(v[self.n_owned:], self._buf, output_split_sizes=self.recv_counts,
                                   input_split_sizes=self.send_counts, group=self.group)
        return v


class ShardedMatrix:
    """Owned rows [r0, r1) of a global CSR matrix, columns in [owned | ghost] numbering."""

    def __init__(self, rowptr, colidx, values, halo, r0, r1, block=1):
        self.rowptr, self.colidx, self.values = rowptr, colidx, values
        self.halo, self.r0, self.r1, self.block = halo, r0, r1, block
        self.n_owned, self.n_local = r1 - r0, halo.n_local

    @property
    def nnz(self):
        return int(self.colidx.numel())

    def take_values(self, global_values):
        """local value array of another operator on the same global pattern (atoms share one pattern)"""
        return global_values[self.k0:self.k1][self.perm].contiguous()


def shard_csr(rowptr, colidx, values, part, rank, group=None):
    """Cut the owned rows of a replicated global CSR (device tensors) and build the halo plan.
    Collective: every rank of `group` must call it (the send lists are agreed with one all-to-all)."""
    import torch.distributed as dist

    dev = rowptr.device
    r0, r1 = part.range(rank)
    k0, k1 = int(rowptr[r0].item()), int(rowptr[r1].item())
    rp = (rowptr[r0:r1 + 1] - k0).to(I32).contiguous()
    cg = colidx[k0:k1].to(I64)
    owned = (cg >= r0) & (cg < r1)
    ghost_global = torch.unique(cg[~owned])  # sorted ascending
    n_owned = r1 - r0
    cl = torch.where(owned, cg - r0, n_owned + torch.searchsorted(ghost_global, cg))
    world = part.world
    owners = part.owner(ghost_global)
    recv_counts = torch.bincount(owners, minlength=world).tolist() if ghost_global.numel() else [0] * world
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        # tell every owner how many and which of its rows this rank needs
        rc = torch.tensor(recv_counts, dtype=I64, device=dev)
        sc = torch.empty(world, dtype=I64, device=dev)
        dist.all_to_all_single(sc, rc, group=group)
        send_counts = sc.tolist()
        want = torch.empty(int(sum(send_counts)), dtype=I64, device=dev)
        dist.all_to_all_single(want, ghost_global.contiguous(), output_split_sizes=send_counts,
                               input_split_sizes=recv_counts, group=group)
        send_idx = (want - r0).contiguous()
    else:
        send_counts = [0] * world
        send_idx = torch.empty(0, dtype=I64, device=dev)
    halo = HaloPlan(n_owned, ghost_global, send_idx, send_counts, recv_counts, group)
    # keep the columns of every row ascending in the LOCAL numbering (ghosts owned by lower ranks moved
    # behind the owned block): the kernels look entries up by binary search
    rows = torch.repeat_interleave(torch.arange(n_owned, dtype=I64, device=dev), (rp[1:] - rp[:-1]).to(I64))
    perm = torch.argsort(rows * (n_owned + int(ghost_global.numel()) + 1) + cl, stable=True)
    S = ShardedMatrix(rp, cl[perm].to(I32).contiguous(), None, halo, r0, r1, part.block)
    S.k0, S.k1, S.perm = k0, k1, perm
    if values is not None:
        S.values = S.take_values(values)
    return S


class _DeviceOps:
    """The per-rank kernels of the sharded PCG (libpgdb200 pgd_spcg_*)."""

    def __init__(self, A, block):
        from . import _lib

        self.lib, self.h = _lib.load_library(), _lib.handle(A.rowptr.device)
        self._lib, self.A, self.block = _lib, A, block
        dev = A.rowptr.device
        no, nl = A.n_owned, A.n_local
        ns = (no + 1) & ~1  # the library keeps every work array on an even offset (pgd_b200.h)
        p_off = 3 * ns + ((no * block + 1) & ~1)
        self.work = torch.empty(p_off + nl + 8, dtype=F64, device=dev)
        self.p = self.work[p_off:p_off + nl]
        self.sc = torch.zeros(16, dtype=F64, device=dev)
        self.fl = torch.zeros(4, dtype=I32, device=dev)

    def _ck(self, rc, what):
        self._lib._check(rc, self.h, what)

    def init(self, b, x):
        L, A = self._lib, self.A
        self._ck(self.lib.pgd_spcg_init(self.h, L._p(A.rowptr, I32), L._p(A.colidx, I32), L._p(A.values, F64), L._p(b, F64),
                                        L._p(x, F64), A.n_owned, A.n_local, self.block, L._p(self.work), L._p(self.sc),
                                        L._p(self.fl), L._stream()), "pgd_spcg_init")

    def init_fin(self, rtol, atol):
        L = self._lib
        self._ck(self.lib.pgd_spcg_init_fin(self.h, L._p(self.sc), L._p(self.fl), float(rtol), float(atol), L._stream()),
                 "pgd_spcg_init_fin")

    def direction(self):
        L = self._lib
        self._ck(self.lib.pgd_spcg_direction(self.h, L._p(self.work), self.A.n_owned, self.block, L._p(self.sc), L._p(self.fl),
                                             L._stream()), "pgd_spcg_direction")

    def matvec(self):
        L, A = self._lib, self.A
        self._ck(self.lib.pgd_spcg_matvec(self.h, L._p(A.rowptr, I32), L._p(A.colidx, I32), L._p(A.values, F64),
                                          L._p(self.work), A.n_owned, self.block, L._p(self.sc), L._stream()), "pgd_spcg_matvec")

    def update(self, x):
        L = self._lib
        self._ck(self.lib.pgd_spcg_update(self.h, L._p(x, F64), L._p(self.work), self.A.n_owned, self.block, L._p(self.sc),
                                          L._p(self.fl), L._stream()), "pgd_spcg_update")

    def rotate(self):
        L = self._lib
        self._ck(self.lib.pgd_spcg_rotate(self.h, L._p(self.sc), L._p(self.fl), L._stream()), "pgd_spcg_rotate")


def _allreduce(t, group):
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def sharded_pcg(A, b, rtol=1e-13, atol=0.0, maxit=10000, check_every=50, block=None, ops=None, use_graph=False,
                use_p2p=True):
    """Solve A x = b over the row partition (A: ShardedMatrix, b: owned slice).  Returns
    (x_owned, iterations, relative residual); identical iteration count on every rank.

    Per iteration: 4 kernels (+ halo pack), one halo exchange of p, two scalar all-reduces (p.q; r.z and
    r.r).  The host only reads the done flag every `check_every` iterations.  Default on CUDA: the loop
    runs inside libpgdb200 (pgd_spcg_solve_sync, its own NCCL communicator).  ``ops`` / ``use_graph``
    select the host-driven variant built from the pgd_spcg_* blocks and torch.distributed collectives
    (with ``use_graph`` captured once in a CUDA graph and replayed; measured slower than the C loop)."""
    block = A.block if block is None else block
    group = A.halo.group
    if ops is None and not use_graph and b.is_cuda:
        # production path: the whole loop (kernels + NCCL send/recv + allreduce) inside libpgdb200
        import torch.distributed as dist

        from . import _lib

        if dist.is_initialized() and dist.get_world_size(group) > 1:
            _lib.comm_init(group)
            if use_p2p and _lib.peer_window(6 * A.n_local + 16, group) is not None:
                A.halo.ensure_peer_layout()
            else:
                A.halo.peer_ghost_base = None
        return _lib.spcg_solve(A, b, rtol=rtol, atol=atol, maxit=maxit, check_every=check_every, block=block)
    ops = ops or _DeviceOps(A, block)
    x = torch.zeros(A.n_owned, dtype=F64, device=b.device)
    ops.init(b, x)
    _allreduce(ops.sc[8:11], group)  # r.z, b.b, r.r
    ops.init_fin(rtol, atol)

    def iteration():
        ops.direction()
        A.halo.exchange(ops.p)
        ops.matvec()
        _allreduce(ops.sc[2:3], group)
        ops.update(x)
        _allreduce(ops.sc[8:10], group)
        ops.rotate()

    graph = None
    if use_graph and b.is_cuda:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            iteration()  # warm-up outside capture (NCCL communicator / buffers)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            iteration()
    done_it = 1 if graph is not None else 0
    launched = done_it
    while True:
        fl = ops.fl.cpu()
        if int(fl[0]) or launched >= maxit:
            break
        todo = min(check_every, maxit - launched)
        for _ in range(todo):
            if graph is not None:
                graph.replay()
            else:
                iteration()
        launched += todo
    sc = ops.sc.cpu()
    fl = ops.fl.cpu()
    iters = int(fl[1])
    relres = float(np.sqrt(float(sc[3]) / float(sc[4]))) if float(sc[4]) > 0 else 0.0
    if int(fl[2]):
        raise RuntimeError("sharded_pcg: NaN encountered (matrix not SPD?)")
    return x, iters, relres


ops_factory = [None]  # tests: callable(A, block) -> kernel stand-in (see module docstring); None = the CUDA kernels


def sharded_solve(A, b, x0=None, rtol=1e-13, atol=0.0, maxit=10000, check_every=50, block=None, bsr=None):
    """Solve A x = b on an element-partitioned space.  b, x0: local vectors [n_local] (only the owned part of b is
    read; x0 with valid ghosts is a warm start).  Returns (x [n_local] with up-to-date ghosts, iterations, relative
    residual), identical iteration count on every rank.

    CUDA: ONE persistent cooperative kernel per rank (pgd_pcg_persist_sync) -- neighbour values of the direction
    vector are stored into the peers' ghost slots over the NVLink peer window, the dot products go through per-rank
    mailboxes, no host round trip and no NCCL call inside the solve.  When the peers cannot be mapped the multi-launch
    loop over NCCL (pgd_spcg_solve_sync) is used instead."""
    import torch.distributed as dist

    block = A.block if block is None else block
    group = A.halo.group
    no = A.n_owned
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if ops_factory[0] is None and b.is_cuda:
        from . import _lib

        if multi:
            if _lib.peer_window(6 * A.n_local + 16, group) is not None:
                A.halo.ensure_peer_layout()
                return _lib.pcg_persist(A.rowptr, A.colidx, A.values, b, n_owned=no, block=block, rtol=rtol, atol=atol,
                                        maxit=maxit, x0=x0, halo=A.halo, bsr=bsr)
            # no peer mapping (the decision above is collective): the NCCL path, whose communicator inside the library
            # is only created now -- it costs ~2 s on 8 GPUs and the persistent kernel never uses it
            _lib.comm_init(group)
            A.halo.peer_ghost_base = None
            x_owned, iters, relres = _lib.spcg_solve(A, b[:no].contiguous(), rtol=rtol, atol=atol, maxit=maxit,
                                                     check_every=check_every, block=block)
        else:
            return _lib.pcg_persist(A.rowptr, A.colidx, A.values, b, n_owned=no, block=block, rtol=rtol, atol=atol,
                                    maxit=maxit, x0=x0, halo=None, bsr=bsr)
    else:
        ops = ops_factory[0](A, block) if ops_factory[0] is not None else None
        x_owned, iters, relres = sharded_pcg(A, b[:no].contiguous(), rtol=rtol, atol=atol, maxit=maxit, check_every=check_every,
                                             block=block, ops=ops)
    x = torch.zeros(A.n_local, dtype=b.dtype, device=b.device)
    x[:no] = x_owned
    A.halo.exchange(x)
    return x, iters, relres


def gather_owned(x_owned, part, group=None):
    """All ranks: the full vector assembled from the owned slices (set-up / test helper)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_owned.clone()
    sizes = [part.range(r)[1] - part.range(r)[0] for r in range(part.world)]
    m = max(sizes)
    mine = torch.zeros(m, dtype=x_owned.dtype, device=x_owned.device)
    mine[: x_owned.numel()] = x_owned
    outs = [torch.empty(m, dtype=x_owned.dtype, device=x_owned.device) for _ in sizes]
    dist.all_gather(outs, mine, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])
