"""-m "not gpu": the multi-GPU host logic (row partition, halo plan, sharded PCG driver) with
world_size 2 and 3 on the gloo backend.  The per-rank kernels are replaced by a NumPy stand-in with the
same interface as pgdrome_b200.partition._DeviceOps; the -m gpu test drives the real kernels."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

F64, I32 = torch.float64, torch.int32


class CpuOps:
    """NumPy stand-in for the pgd_spcg_* kernels (same state layout: sc[16], fl[4], p with ghost tail)."""

    def __init__(self, A, block):
        self.A, self.block = A, block
        no, nl = A.n_owned, A.n_local
        self.M = sp.csr_matrix((A.values.numpy(), A.colidx.numpy(), A.rowptr.numpy()), shape=(no, nl))
        self.p = torch.zeros(nl, dtype=F64)
        self.sc = torch.zeros(16, dtype=F64)
        self.fl = torch.zeros(4, dtype=I32)

    def init(self, b, x):
        no, bs = self.A.n_owned, self.block
        D = self.M[:, :no].tocsr()
        self.minv = []
        for nd in range(no // bs):
            self.minv.append(np.linalg.inv(D[nd * bs:(nd + 1) * bs, nd * bs:(nd + 1) * bs].toarray()))
        self.r = b.numpy().copy()
        self.z = self._prec(self.r)
        self.q = np.zeros(no)
        self.x = x
        x.zero_()
        self.p.zero_()
        self.sc.zero_()
        self.fl.zero_()
        self.sc[8], self.sc[9], self.sc[10] = float(self.r @ self.z), float(self.r @ self.r), float(self.r @ self.r)

    def _prec(self, r):
        bs = self.block
        if not len(self.minv):
            return r.copy()
        return np.concatenate([self.minv[nd] @ r[nd * bs:(nd + 1) * bs] for nd in range(len(r) // bs)])

    def init_fin(self, rtol, atol):
        rz, bb, rr = float(self.sc[8]), float(self.sc[9]), float(self.sc[10])
        self.sc[0], self.sc[1], self.sc[2], self.sc[3], self.sc[4] = 1.0, rz, 1.0, rr, bb
        self.sc[5] = max(rtol * rtol * bb, atol * atol)
        self.fl[0] = 1 if (bb <= float(self.sc[5]) or bb == 0.0) else 0

    def direction(self):
        if int(self.fl[0]):
            return
        beta = 0.0 if int(self.fl[1]) == 0 else float(self.sc[1] / self.sc[0])
        no = self.A.n_owned
        self.p[:no] = torch.as_tensor(self.z) + beta * self.p[:no]

    def matvec(self):
        self.q = self.M @ self.p.numpy()
        self.sc[2] = float(self.p.numpy()[: self.A.n_owned] @ self.q)

    def update(self, x):
        if int(self.fl[0]):
            return
        alpha = float(self.sc[1] / self.sc[2])
        no = self.A.n_owned
        x += alpha * self.p[:no]
        self.r = self.r - alpha * self.q
        self.z = self._prec(self.r)
        self.sc[8], self.sc[9] = float(self.r @ self.z), float(self.r @ self.r)

    def rotate(self):
        if int(self.fl[0]):
            return
        self.sc[0] = self.sc[1].clone()
        self.sc[1] = self.sc[8].clone()
        self.sc[3] = self.sc[9].clone()
        self.fl[1] += 1
        if not float(self.sc[9]) > float(self.sc[5]):
            self.fl[0] = 1


def _system(n_nodes, block, seed=0):
    rng = np.random.default_rng(seed)
    n = n_nodes * block
    rows = np.repeat(np.arange(n), 6)
    cols = np.clip(rows + rng.integers(-9 * block, 9 * block + 1, size=rows.size), 0, n - 1)
    A = sp.coo_matrix((rng.uniform(-1, 1, rows.size), (rows, cols)), shape=(n, n)).tocsr()
    A = A + A.T
    A = (A + sp.diags(np.abs(A).sum(axis=1).A1 + 1.0)).tocsr()
    A.sort_indices()
    return A, rng.uniform(-1, 1, n)


def _worker(rank, world, port, block, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pgdrome_b200 import partition as pt

        A, xs = _system(401, block)
        n = A.shape[0]
        part = pt.RowPartition(n, world, block)
        assert part.bounds[0] == 0 and part.bounds[-1] == n and all(b % block == 0 for b in part.bounds)
        rp, ci, va = (torch.as_tensor(A.indptr.astype(np.int32)), torch.as_tensor(A.indices.astype(np.int32)),
                      torch.as_tensor(A.data))
        S = pt.shard_csr(rp, ci, va, part, rank)
        r0, r1 = part.range(rank)
        assert S.n_owned == r1 - r0 and S.nnz == A.indptr[r1] - A.indptr[r0]
        # ghosts are exactly the off-range columns of the owned rows, ascending => grouped by owner
        cols = np.unique(A.indices[A.indptr[r0]:A.indptr[r1]])
        ghosts = cols[(cols < r0) | (cols >= r1)]
        assert np.array_equal(S.halo.ghost_global.numpy(), ghosts)
        assert sum(S.halo.recv_counts) == len(ghosts) and S.halo.recv_counts[rank] == 0
        # halo exchange: every ghost slot receives the owner's value
        v = torch.zeros(S.n_local, dtype=F64)
        v[: S.n_owned] = torch.as_tensor(xs[r0:r1])
        S.halo.exchange(v)
        assert np.array_equal(v[S.n_owned:].numpy(), xs[ghosts])
        M = sp.csr_matrix((S.values.numpy(), S.colidx.numpy(), S.rowptr.numpy()), shape=(S.n_owned, S.n_local))
        assert M.has_sorted_indices or all(np.all(np.diff(S.colidx.numpy()[a:b]) > 0)
                                           for a, b in zip(S.rowptr.numpy()[:-1], S.rowptr.numpy()[1:]))  # kernels bisect rows
        assert all(np.all(np.diff(S.colidx.numpy()[a:b]) > 0) for a, b in zip(S.rowptr.numpy()[:-1], S.rowptr.numpy()[1:]))
        assert np.allclose(M @ v.numpy(), (A @ xs)[r0:r1], rtol=1e-14, atol=1e-14)
        # sharded PCG (driver logic + collectives) against a direct solve
        b = torch.as_tensor((A @ xs)[r0:r1])
        x, iters, relres = pt.sharded_pcg(S, b, rtol=1e-13, maxit=500, check_every=7, block=block, ops=CpuOps(S, block))
        full = pt.gather_owned(x, part).numpy()
        ref = spla.spsolve(A.tocsc(), A @ xs)
        assert relres <= 1e-13 and 0 < iters < 500
        assert np.linalg.norm(full - ref) / np.linalg.norm(ref) < 1e-11
        its = torch.tensor([iters])
        dist.all_reduce(its, op=dist.ReduceOp.MAX)
        assert int(its) == iters  # same count on every rank
        # zero right-hand side: converged at once, no iteration
        x0, it0, _ = pt.sharded_pcg(S, torch.zeros_like(b), maxit=5, block=block, ops=CpuOps(S, block))
        assert it0 == 0 and float(x0.abs().max()) == 0.0
        out[rank] = iters
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,block", [(2, 1), (2, 3), (3, 1)])
def test_partition_halo_and_sharded_pcg_gloo(world, block):
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), block, out), nprocs=world, join=True)
    assert len(out) == world and len(set(out.values())) == 1


def test_row_partition_single_rank():
    from pgdrome_b200 import partition as pt

    A, xs = _system(50, 2)
    part = pt.RowPartition(A.shape[0], 1, 2)
    S = pt.shard_csr(torch.as_tensor(A.indptr.astype(np.int32)), torch.as_tensor(A.indices.astype(np.int32)),
                     torch.as_tensor(A.data), part, 0)
    assert S.halo.n_ghost == 0 and S.n_local == A.shape[0]
    x, iters, relres = pt.sharded_pcg(S, torch.as_tensor(A @ xs), rtol=1e-13, maxit=300, block=2, ops=CpuOps(S, 2))
    assert np.linalg.norm(x.numpy() - xs) / np.linalg.norm(xs) < 1e-11
