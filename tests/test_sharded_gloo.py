"""-m "not gpu": the element-partitioned enrichment loop (pgdrome_b200/sharding.py) with world_size 2 and 3 on the
gloo backend.  Every rank runs the unchanged user script (configs.heat2d_tk / elasticity3d); the spatial space is
partitioned, the kernels are the NumPy stand-ins of tests/cpu_abi.py and the sharded PCG runs the stand-in ops of
tests/test_partition_gloo.py.  Checked against the same problem solved without partitioning in the same process:
modes to 1e-9, identical fixed-point iteration counts, identical host-visible vectors on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Patch:
    """minimal monkeypatch (the worker processes have no pytest fixture)"""

    def __init__(self):
        self.undo = []

    def setattr(self, obj, name, value):
        self.undo.append((obj, name, getattr(obj, name)))
        setattr(obj, name, value)


def _space_checks(p, rank, world):
    from pgdrome_b200 import sharding
    from pgdrome_b200.assembly import device_space

    V = p.V[0]
    sh = sharding.shard_of(V)
    assert sh is not None and sh.world == world
    # owned ranges tile the global dofs; ghosts are exactly the foreign nodes of the local cells
    cn = V.cell_nodes
    g0, g1 = sh.node_range
    mine = ((cn >= g0) & (cn < g1)).any(axis=1)
    nodes = np.unique(cn[mine])
    assert np.array_equal(sh.ghost_nodes, nodes[(nodes < g0) | (nodes >= g1)])
    assert sh.n_local == len(nodes) * V.bs
    # halo exchange delivers the owners' values
    ds = device_space(V)
    t = torch.zeros(sh.n_local, dtype=torch.float64)
    glob = np.arange(V.n_dofs, dtype=np.float64) * 0.5 + 1.0
    t[: sh.n_owned] = torch.as_tensor(glob[sh.l2g_dofs[: sh.n_owned]])
    sh.update_ghosts(t)
    assert np.array_equal(t.numpy(), glob[sh.l2g_dofs])
    assert np.array_equal(sh.gather_host(t), glob)
    # local pattern rows of owned dofs = the global pattern rows (columns mapped back to global numbers)
    return ds


def _worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pgdrome_b200 import configs, partition, sharding
        from tests import cpu_abi
        from tests.test_partition_gloo import CpuOps

        cpu_abi.install(_Patch())
        partition.ops_factory[0] = CpuOps
        kw = dict(heat2d_tk=dict(n=10, nt=12, nk=5, PGD_nmax=3), elasticity3d=dict(n=3, nE=5, nF=2, PGD_nmax=2))[name]
        sharding.configure(mode=False)
        ref = getattr(configs, name)(**kw)
        ref.solve_PGD(_problem="linear")
        ref_modes = [[f.vector().get_local() for f in ref.PGD_func[d]] for d in range(ref.num_pgd_var)]
        sharding.configure(mode=True)
        p = getattr(configs, name)(**kw)
        p.solve_PGD(_problem="linear")
        ds = _space_checks(p, rank, world)
        assert ds.shard is not None and ds.n_owned < ds.n_dofs
        assert p.solver_stats.get("sharded_solves", 0) > 0
        assert p.PGD_modes == ref.PGD_modes and p.num_fp_it == ref.num_fp_it, (p.num_fp_it, ref.num_fp_it)
        worst = 0.0
        for d in range(p.num_pgd_var):
            for k in range(p.PGD_modes):
                a, b = p.PGD_func[d][k].vector().get_local(), ref_modes[d][k]
                assert a.shape == b.shape  # host-visible vectors are global on every rank
                worst = max(worst, min(np.linalg.norm(a - b), np.linalg.norm(a + b)) / np.linalg.norm(b))
        assert worst < 1e-9, worst
        assert np.allclose(p.alpha, ref.alpha, rtol=1e-9)
        # a spatial mode evaluated at a point / at the vertices: post-processing on the gathered vector
        f = p.PGD_func[0][0]
        assert np.allclose(f.compute_vertex_values(), ref.PGD_func[0][0].compute_vertex_values(), rtol=1e-8, atol=1e-12) or \
            np.allclose(f.compute_vertex_values(), -ref.PGD_func[0][0].compute_vertex_values(), rtol=1e-8, atol=1e-12)
        out[rank] = (tuple(p.num_fp_it), float(worst))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,name", [(2, "heat2d_tk"), (3, "heat2d_tk"), (2, "elasticity3d")])
def test_sharded_enrichment_matches_serial_gloo(world, name):
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    assert len(out) == world
    assert len({v[0] for v in out.values()}) == 1  # identical control flow on every rank
