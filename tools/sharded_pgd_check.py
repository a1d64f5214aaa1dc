"""2-GPU check of the sharded spatial solve inside the enrichment loop:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P -m tools.sharded_pgd_check

Every rank runs reduced configs[1]-[3] twice -- replicated on one GPU, then with the spatial space element-partitioned
over all ranks (sharding.py: local pattern / atoms / panels, all-reduced mode integrals, persistent sharded PCG over the
NVLink peer window) -- and compares modes, amplitudes and iteration counts; also reports device memory per rank."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist

    from pgdrome_b200 import _lib, configs

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.set_option("spin_ms", 5000)
    cases = [("heat2d_tk", lambda: configs.heat2d_tk(n=96, nt=40, nk=10, PGD_nmax=3)),
             ("elasticity3d (vector P1, node-block Jacobi)", lambda: configs.elasticity3d(n=12, nE=8, nF=2, PGD_nmax=2)),
             ("thermal3d", lambda: configs.thermal3d(n=20, nt=30, nP=4, nv=4, n_src=3, PGD_nmax=2))]
    for name, make in cases:
        check(name, make)
    dist.destroy_process_group()


def check(name, make):
    import torch.distributed as dist

    rank = dist.get_rank()
    n = 0
    nmax = 0
    from pgdrome_b200 import sharding

    sharding.configure(mode=False)
    a = make()
    a.solve_PGD(_problem="linear")
    torch.cuda.synchronize()
    mem_a = torch.cuda.max_memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    sharding.configure(mode=True)
    b = make()
    b.solve_PGD(_problem="linear")
    torch.cuda.synchronize()
    mem_b = torch.cuda.max_memory_allocated()
    sharding.configure(mode="auto")
    assert b.solver_stats.get("sharded_solves", 0) > 0 and a.solver_stats.get("sharded_solves", 0) == 0
    assert a.PGD_modes == b.PGD_modes and a.num_fp_it == b.num_fp_it, (a.num_fp_it, b.num_fp_it)
    worst = 0.0
    for d in range(len(a.V)):
        for k in range(a.PGD_modes):
            u, v = a.PGD_func[d][k].vector()[:], b.PGD_func[d][k].vector()[:]
            worst = max(worst, min(np.linalg.norm(u - v), np.linalg.norm(u + v)) / np.linalg.norm(u))
    assert worst < 1e-8, worst
    assert np.allclose(a.amplitude, b.amplitude, rtol=1e-8, atol=0)
    # every rank must hold bitwise the same modes (the replicated dimensions rely on it)
    chk = torch.tensor([float(np.sum(b.PGD_func[0][-1].vector()[:]))], dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(lo) == float(hi)
    if rank == 0:
        print(json.dumps({"ok": True, "case": name, "world": dist.get_world_size(), "spatial_dofs": a.V[0].n_dofs, "modes": a.PGD_modes,
                          "fp_iterations": a.num_fp_it, "worst_mode_diff": worst, "sharded_solves": b.solver_stats["sharded_solves"],
                          "pcg_iterations": [a.solver_stats["pcg_iterations"], b.solver_stats["pcg_iterations"]],
                          "peak_device_bytes": {"replicated": mem_a, "partitioned_rank0": mem_b}}))


if __name__ == "__main__":
    main()
