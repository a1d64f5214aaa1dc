"""Streaming (HBM-bound) PCG iteration on the operators of BASELINE configs[2] / configs[3]: per-kernel and
per-iteration device times against the byte model of DESIGN.md.

  python -m tools.pcg_bench --mesh 68 --bs 3            configs[2] pattern (vector P1, 985 527 dofs, 45 nnz/row)
  python -m tools.pcg_bench --mesh 158 --bs 1           configs[3] pattern (scalar P1, 4 019 679 dofs, 15 nnz/row)
  ... --profile                                          few launches only (for ncu captures)

Prints one JSON line.  CUDA events on the launching stream, >= 3 warm-ups, CSR arrays >> L2."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def operator(mesh_n, bs):
    from pgdrome_b200 import fem
    from pgdrome_b200.assembly import device_space

    m = fem.UnitCubeMesh(mesh_n, mesh_n, mesh_n)
    V = fem.FunctionSpace(m, "P", 1) if bs == 1 else fem.VectorFunctionSpace(m, "P", 1)
    ds = device_space(V)
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
    operator.T = T
    return V, ds, vals


def run_sharded(mesh_n=68, bs=3, iters=200, hbm_peak=6451.2):
    """torchrun: the same operator element-partitioned over the ranks; per-iteration time and phase profile of the
    persistent sharded PCG (max over ranks)."""
    import torch.distributed as dist

    from pgdrome_b200 import _lib, partition as pt, sharding

    rank, world = dist.get_rank(), dist.get_world_size()
    sharding.configure(mode=True)
    V, ds, vals = operator(mesh_n, bs)
    sh = ds.shard
    dev = vals.device
    halo = sh.halo(dev)
    S = pt.ShardedMatrix(ds.rowptr_owned, ds.pattern[1][: ds.nnz_owned], vals[: ds.nnz_owned], halo,
                         int(sh.part.bounds[rank]), int(sh.part.bounds[rank + 1]), bs)
    xg = np.random.default_rng(0).uniform(-1, 1, V.n_dofs)
    x = sh.scatter(xg, dev)
    b = torch.zeros(sh.n_local, dtype=torch.float64, device=dev)
    _lib.spmv(ds.rowptr_owned, ds.pattern[1], vals, x, y=b)
    _lib.set_option("spin_ms", 5000)
    out = {"mesh": "BoxMesh %d^3, P1 bs=%d" % (mesh_n, bs), "world": world, "n_dofs": V.n_dofs, "rows_rank0": sh.n_owned,
           "ghosts_rank0": sh.n_ghost, "nnz_rank0": ds.nnz_owned, "hbm_peak_gbs": hbm_peak}
    for name, bsr, sr, ll in (("persist_bsr", ds.bsr, 0, 1), ("persist_bsr_flag_protocol", ds.bsr, 0, 0),
                              ("persist_bsr_repeat", ds.bsr, 0, 1), ("persist_bsr_single_reduction", ds.bsr, 1, 0),
                              ("persist_csr", None, 0, 1), ("persist_csr_single_reduction", None, 1, 0)):
        if name.startswith("persist_bsr") and bsr is None:
            continue
        if name.startswith("persist_csr") and ds.bsr is not None and sr == 1:
            continue
        _lib.set_option("single_reduction", sr)
        _lib.set_option("ll", ll)
        pt.sharded_solve(S, b, rtol=1e-30, maxit=10, block=bs, bsr=bsr)
        _lib.set_option("prof", 1)
        _lib.phase_ns(reset=True)
        pt.sharded_solve(S, b, rtol=1e-30, maxit=iters, block=bs, bsr=bsr)
        ph = _lib.phase_ns(reset=True)
        _lib.set_option("prof", 0)
        torch.cuda.synchronize()
        dist.barrier()
        _lib.stats(reset=True)
        pt.sharded_solve(S, b, rtol=1e-30, maxit=iters, block=bs, bsr=bsr)
        s = _lib.stats()
        ms = torch.tensor([s["pcg_ms"] / max(s["pcg_iters"], 1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        loc = 12 * ds.nnz_owned + 4 * (sh.n_owned + 1) + 56 * sh.n_owned
        xs, its, rr = pt.sharded_solve(S, b, rtol=1e-13, maxit=20000, block=bs, bsr=bsr)
        err = torch.tensor([float((xs[: sh.n_owned] - x[: sh.n_owned]).pow(2).sum()), float(x[: sh.n_owned].pow(2).sum())],
                           dtype=torch.float64, device=dev)
        dist.all_reduce(err)
        ghost_ok = bool(torch.allclose(xs[sh.n_owned:], x[sh.n_owned:], rtol=0, atol=1e-8))
        ax = torch.zeros(sh.n_local, dtype=torch.float64, device=dev)
        _lib.spmv(ds.rowptr_owned, ds.pattern[1], vals, xs, y=ax)
        tr = torch.tensor([float((b[: sh.n_owned] - ax[: sh.n_owned]).pow(2).sum()), float(b[: sh.n_owned].pow(2).sum())],
                          dtype=torch.float64, device=dev)
        dist.all_reduce(tr)
        out[name] = {"ms_per_iteration": ms, "local_bytes": loc, "local_gbs": loc / (ms * 1e-3) / 1e9,
                     "frac_hbm_local": loc / (ms * 1e-3) / 1e9 / hbm_peak, "phase_us_per_iteration_rank0": {k: v / 1e3 / iters for k, v in ph.items()},
                     "solve": {"iters": its, "relres": rr, "true_relres": float((tr[0] / tr[1]).sqrt()),
                               "err": float((err[0] / err[1]).sqrt()), "ghosts_of_solution_ok": ghost_ok}}
    _lib.set_option("single_reduction", 0)
    _lib.set_option("ll", 1)
    return out if rank == 0 else None


def run(mesh_n=68, bs=3, iters=200, profile=False, hbm_peak=6451.2):
    from pgdrome_b200 import _lib

    V, ds, vals = operator(mesh_n, bs)
    rowptr, colidx, _, _ = ds.pattern
    n, nnz = ds.n_dofs, ds.nnz
    dev = rowptr.device
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    b = _lib.spmv(rowptr, colidx, vals, x, lpr=ds.lpr)
    out = {"mesh": "BoxMesh %d^3, P1 bs=%d" % (mesh_n, bs), "n_dofs": n, "nnz": nnz, "hbm_peak_gbs": hbm_peak}

    def timed(fn, reps=20, warm=3):
        if profile:
            reps, warm = 1, 1
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, c in evs:
            a.record()
            fn()
            c.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(c) for a, c in evs)
        return ms[len(ms) // 2]

    def entry(name, ms, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = dict(ms=ms, bytes=nbytes, gbs=gbs, frac_hbm=gbs / hbm_peak, **kw)

    # the atom itself: general-tensor row-owner kernel against the element-matrix + gather route it replaces for P1
    nc, nv = V.mesh().num_cells(), V.mesh().num_vertices()
    asm_bytes = 16 * nc + 24 * nv + 8 * nnz
    entry("assemble_p1_tensor", timed(lambda: ds.assemble_bilinear(operator.T), reps=5, warm=2), asm_bytes,
          kernel="k_assemble_p1_tensor<3,%d>" % bs)
    if not profile:
        plan_keep, ds._node_plan = ds.node_plan, False  # without the node plan the same atom takes the element-matrix route
        entry("assemble_elem_gather", timed(lambda: ds.assemble_bilinear(operator.T), reps=3, warm=1), asm_bytes,
              kernel="k_elem_bilinear + k_gather_sum")
        ds._node_plan = plan_keep
    y = torch.empty(n, dtype=torch.float64, device=dev)
    sc = torch.empty(1, dtype=torch.float64, device=dev)
    spmv_bytes = 12 * nnz + 4 * (n + 1) + 16 * n
    entry("spmv", timed(lambda: _lib.spmv(rowptr, colidx, vals, x, y=y)), spmv_bytes)
    entry("spmv_dot", timed(lambda: _lib.spmv_dot(rowptr, colidx, vals, x, x, y=y, out=sc)), spmv_bytes)
    it_bytes = 12 * nnz + 4 * (n + 1) + 56 * n
    work = torch.empty((5 + bs) * n + 16, dtype=torch.float64, device=dev)
    _lib.set_option("pcg_resident", 0)
    _lib.set_option("spin_ms", 3000)
    k = 3 if profile else iters
    plan = _lib.bsr_plan(rowptr, colidx, bs) if bs > 1 else None
    out["bsr_plan"] = None if plan is None else {"blocks": int(plan[0].numel()), "max_blocks_per_row": plan[1]}
    variants = [("pcg_3launch", dict(persist=1), None), ("pcg_persist_csr", dict(persist=2, bsr=0), None)]
    if plan is not None:
        variants.append(("pcg_persist_bsr", dict(persist=1, bsr=1), plan))
        variants.append(("pcg_persist_bsr_direct", dict(persist=1, bsr=2), plan))
        variants.append(("pcg_persist_bsr_single_reduction", dict(persist=1, bsr=1, single_reduction=2), plan))
    else:
        # scalar operator: the direct walk with bcol = colidx (lanes per row from the longest row)
        direct = (colidx, int((rowptr[1:] - rowptr[:-1]).max().item()))
        variants.append(("pcg_persist_csr_direct", dict(persist=2, bsr=2), direct))
        variants.append(("pcg_persist_csr_single_reduction", dict(persist=2, bsr=0, single_reduction=2), None))
    sols = {}
    for name, opts, pl in variants:
        _lib.set_option("single_reduction", 0)
        for o, v in opts.items():
            _lib.set_option(o, v)

        def solve(rtol, maxit):
            if pl is not None:
                return _lib.pcg_persist(rowptr, colidx, vals, b, block=bs, rtol=rtol, maxit=maxit, bsr=pl, work=work)
            return _lib.pcg(rowptr, colidx, vals, b, rtol=rtol, maxit=maxit, check_every=maxit, block=bs, work=work)

        solve(1e-30, 10)
        _lib.set_option("prof", 1)
        _lib.phase_ns(reset=True)
        solve(1e-30, k)
        ph = _lib.phase_ns(reset=True)
        _lib.set_option("prof", 0)
        _lib.stats(reset=True)
        solve(1e-30, k)
        s = _lib.stats()
        bytes_moved = it_bytes if (pl is None or bs == 1) else (8 * nnz + 4 * (nnz // (bs * bs)) + 4 * (n + 1) + 56 * n)
        entry(name, s["pcg_ms"] / max(s["pcg_iters"], 1), it_bytes, iters=s["pcg_iters"], launches=s["launches"],
              phase_us_per_iteration={kk: v / 1e3 / k for kk, v in ph.items()} if (opts.get("persist") == 2 or pl is not None) else None,
              bytes_of_format=bytes_moved, frac_hbm_of_format=bytes_moved / (s["pcg_ms"] / max(s["pcg_iters"], 1) * 1e-3) / 1e9 / hbm_peak)
        if not profile:
            _lib.stats(reset=True)
            xs, its, rr = solve(1e-13, 20000)
            s = _lib.stats()
            sols[name] = xs[:n].clone()
            true_res = float((b - _lib.spmv(rowptr, colidx, vals, xs[:n].contiguous())).norm() / b.norm())
            out[name]["solve"] = {"iters": its, "relres": rr, "true_relres": true_res, "err": float((xs[:n] - x).norm() / x.norm()),
                                  "ms": s["pcg_ms"]}
            # warm start from the converged solution: must stop at once
            if pl is not None:
                _, its2, rr2 = _lib.pcg_persist(rowptr, colidx, vals, b, block=bs, rtol=1e-12, maxit=100, x0=xs, bsr=pl, work=work)
            else:
                _, its2, rr2 = _lib.pcg(rowptr, colidx, vals, b, rtol=1e-12, maxit=100, check_every=50, block=bs, work=work, x0=xs)
            out[name]["warm_restart"] = {"iters": its2, "relres": rr2}
    if len(sols) > 1:
        ref = sols["pcg_3launch"]
        out["max_rel_diff_vs_3launch"] = {kk: float((v - ref).norm() / ref.norm()) for kk, v in sols.items()}
    _lib.set_option("persist", 1)
    _lib.set_option("bsr", 1)
    _lib.set_option("single_reduction", 1)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=68)
    ap.add_argument("--bs", type=int, default=3)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    peak = 6451.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
    except Exception:
        pass
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        r = run_sharded(a.mesh, a.bs, a.iters, peak)
        if r is not None:
            print(json.dumps(r))
        dist.destroy_process_group()
    else:
        print(json.dumps(run(a.mesh, a.bs, a.iters, a.profile, peak)))
