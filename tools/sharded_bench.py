"""Sharded spatial PCG over the GPUs of one box (one process per GPU, NCCL).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      -m tools.sharded_bench --mesh 128 [--check] [--graph]

Builds the P1 operator 0.3 M + 1.7 K on a mesh^3 box (replicated set-up: every rank assembles the
global pattern and values with the same kernels, then keeps its row slab), shards it by rows, runs a
fixed number of Jacobi-PCG iterations and prints one JSON line from rank 0: per-iteration device time
(CUDA events, max over ranks), the HBM roofline of the local slab and the halo volume.  --check also
solves to 1e-12 and compares with the single-GPU solver (rank 0 solves the full system)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(mesh_n=128, iters=200, check=False, graph=False, bs=1, hbm_peak=6451.2, host_loop=False, p2p=True):
    import torch.distributed as dist

    from pgdrome_b200 import _lib, fem, partition as pt
    from pgdrome_b200.assembly import device_space

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    own_pg = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        own_pg = True
    m = fem.UnitCubeMesh(mesh_n, mesh_n, mesh_n)
    V = fem.FunctionSpace(m, "P", 1) if bs == 1 else fem.VectorFunctionSpace(m, "P", 1)
    ds = device_space(V)
    rowptr, colidx, _, _ = ds.pattern
    g = 3
    if bs == 1:
        T = np.zeros((1, g + 1, 1, g + 1))
        T[0, 0, 0, 0] = 0.3
        for k in range(1, g + 1):
            T[0, k, 0, k] = 1.7
    else:  # isotropic elasticity + mass shift (SPD without boundary conditions)
        lam, mu = 1.3, 0.7
        T = np.zeros((bs, g + 1, bs, g + 1))
        for i in range(bs):
            T[i, 0, i, 0] = 0.3
            for j in range(bs):
                T[i, 1 + i, j, 1 + j] += lam
                T[i, 1 + j, j, 1 + i] += mu
                T[i, 1 + j, i, 1 + j] += mu
    vals = ds.assemble_bilinear(T)
    n, nnz = ds.n_dofs, ds.nnz
    dev = rowptr.device
    gen = torch.Generator(device=dev).manual_seed(0)
    xs = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    b_full = _lib.spmv(rowptr, colidx, vals, xs, lpr=ds.lpr)
    part = pt.RowPartition(n, world, bs)
    A = pt.shard_csr(rowptr, colidx, vals, part, rank)
    hops = (lambda: pt._DeviceOps(A, bs)) if host_loop else (lambda: None)
    r0, r1 = part.range(rank)
    b = b_full[r0:r1].contiguous()
    out = {"mesh": "BoxMesh %d^3 cells, P1%s" % (mesh_n, "" if bs == 1 else " vector"), "n_dofs": n, "nnz": nnz,
           "world": world, "rows_per_rank": r1 - r0, "ghosts_rank0": A.halo.n_ghost,
           "halo_bytes_per_exchange_rank0": A.halo.bytes_per_exchange, "cuda_graph": bool(graph), "loop": "host (torch.distributed)" if (host_loop or graph) else ("libpgdb200 (NVLink peer window)" if p2p and world > 1 else "libpgdb200 (NCCL)")}
    if check:
        x, it, rr = pt.sharded_pcg(A, b, rtol=1e-12, maxit=20000, check_every=50, block=bs, use_graph=graph, ops=hops(), use_p2p=p2p)
        full = pt.gather_owned(x, part)
        err = float((full - xs).norm() / xs.norm())
        _lib.set_option("pcg_resident", 0)
        x1, it1, _ = _lib.pcg(rowptr, colidx, vals, b_full, rtol=1e-12, maxit=20000, check_every=50, block=bs, lpr=ds.lpr)
        d = float((full - x1).norm() / x1.norm())
        out["check"] = {"iters_sharded": it, "iters_single": it1, "relres": rr, "err_vs_exact": err, "diff_vs_single_gpu": d}
        assert rr <= 1e-12 and err < 1e-8 and d < 1e-8 and abs(it - it1) <= max(3, it1 // 20), out["check"]
    del b_full
    # timing: fixed iteration count (rtol 0 never converges), events on the launching stream
    pt.sharded_pcg(A, b, rtol=0.0, maxit=20, check_every=20, block=bs, use_graph=graph, ops=hops(), use_p2p=p2p)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, it, _ = pt.sharded_pcg(A, b, rtol=0.0, maxit=iters, check_every=iters, block=bs, use_graph=graph, ops=hops(), use_p2p=p2p)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per_it = float(ms.item()) / max(it, 1)
    loc_bytes = 12 * A.nnz + 4 * (A.n_owned + 1) + 56 * A.n_owned
    out.update(iters=it, ms_per_iteration=per_it, local_bytes_per_iteration=loc_bytes,
               local_gbs=loc_bytes / (per_it * 1e-3) / 1e9, frac_hbm=loc_bytes / (per_it * 1e-3) / 1e9 / hbm_peak,
               global_gbs=(12 * nnz + 4 * n + 56 * n) / (per_it * 1e-3) / 1e9)
    # configs[4]: the vademecum sweep U[C, N] = W^T X sharded by rows of the spatial dimension -- every rank
    # reconstructs its own N/world slab, no communication (model.evaluate_batch(rows=...))
    R, N, C = 50, 100000, 10000
    n0, n1 = (N * rank) // world, (N * (rank + 1)) // world
    n0, n1 = n0 - n0 % 2, (n1 - n1 % 2) if rank < world - 1 else n1
    X = torch.randn((R, N), dtype=torch.float64, device=dev, generator=gen)
    Wt = torch.randn((R, C), dtype=torch.float64, device=dev, generator=gen)
    U = torch.empty((C, n1 - n0), dtype=torch.float64, device=dev)
    Xs = X[:, n0:n1]
    for _ in range(3):
        _lib.eval_gemm(Wt, Xs, R, out=U)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(5):
        _lib.eval_gemm(Wt, Xs, R, out=U)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out["evaluate_sweep"] = {"N": N, "C": C, "R": R, "rows_per_rank": n1 - n0, "ms": float(ms.item()),
                             "tflops_total": 2.0 * N * C * R / (float(ms.item()) * 1e-3) / 1e12, "collectives": 0}
    if own_pg:
        dist.destroy_process_group()
    return out if rank == 0 else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=128)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--bs", type=int, default=1)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--fused", action="store_true", help="peer-window path with the collectives fused into the SpMV / update kernels")
    ap.add_argument("--no-cgraph", action="store_true", help="peer-window path without CUDA-graph replay")
    ap.add_argument("--no-p2p", action="store_true", help="NCCL send/recv + allreduce instead of the NVLink peer window")
    ap.add_argument("--host-loop", action="store_true", help="torch.distributed-driven iteration instead of the C loop")
    a = ap.parse_args()
    if a.no_cgraph or a.fused:
        from pgdrome_b200 import _lib as _l

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if a.no_cgraph:
            _l.set_option("graph", 0)
        if a.fused:
            _l.set_option("fused", 1)
    r = run(a.mesh, a.iters, a.check, a.graph, a.bs, host_loop=a.host_loop, p2p=not a.no_p2p)
    if r is not None:
        print(json.dumps(r))
