"""Run one BASELINE config at a chosen size on the B200 and print a timing JSON line.

  python -m tools.run_config --config thermal3d --n 158 --modes 3

Reports set-up (mesh, pattern, atoms), per-enrichment-step device time, PCG iterations and the
byte-based PCG roofline of the spatial solves.  Full sizes: elasticity3d --n 68 (985 527 dofs),
thermal3d --n 158 (4 019 679 dofs)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="thermal3d", choices=["heat2d_tk", "elasticity3d", "thermal3d"])
    ap.add_argument("--size", "--n", dest="n", type=int, default=None, help="cells per edge of the spatial mesh")
    ap.add_argument("--modes", type=int, default=3)
    ap.add_argument("--rtol", type=float, default=1e-13)
    ap.add_argument("--setup-profile", default=None, help="write a cProfile of problem construction + step 0 (rank 0) here")
    ap.add_argument("--step-profile", default=None, help="write a cProfile of the steps --step-profile-range (rank 0) here")
    ap.add_argument("--step-profile-range", default="20:22")
    ap.add_argument("--counts", default=None, help="write the per-step / per-solve PCG iteration counts to this JSON file "
                                                   "(bench.py scales its bounded CPU sample with them)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:  # torchrun: the spatial solves are sharded over all ranks (settings["sharded"] = "auto")
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pgdrome_b200 import _lib, configs

    _lib.set_option("spin_ms", 10000)
    peak = 6451.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
    except Exception:
        pass
    prof = None
    if a.setup_profile and rank == 0:  # where does the first (set-up) step go?  cProfile of build + step 0 on rank 0
        import cProfile

        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    kw = {} if a.n is None else {"n": a.n}
    p = getattr(configs, a.config)(PGD_nmax=a.modes, PGD_tol=0.0, **kw)
    t_build = time.perf_counter() - t0
    st = p.begin_PGD(_problem="linear", settings={"linear_solver": "cg", "relative_tolerance": a.rtol})
    steps = []
    _lib.stats(reset=True)
    sprof = None
    for i in range(a.modes):
        if a.step_profile and rank == 0:  # cProfile of the enrichment steps [first, last] on rank 0 (host share of a warm step)
            first, last = (int(v) for v in a.step_profile_range.split(":"))
            if i == first:
                import cProfile

                sprof = cProfile.Profile()
                sprof.enable()
            if i == last + 1 and sprof is not None:
                import io
                import pstats

                sprof.disable()
                sio = io.StringIO()
                pstats.Stats(sprof, stream=sio).sort_stats("cumulative").print_stats(80)
                pstats.Stats(sprof, stream=sio).sort_stats("tottime").print_stats(40)
                open(a.step_profile, "w").write(sio.getvalue())
                sprof = None
        if i == 1 and prof is not None:
            import io
            import pstats

            prof.disable()
            sio = io.StringIO()
            pstats.Stats(prof, stream=sio).sort_stats("cumulative").print_stats(70)
            pstats.Stats(prof, stream=sio).sort_stats("tottime").print_stats(30)
            open(a.setup_profile, "w").write(sio.getvalue())
            prof = None
        torch.cuda.synchronize()
        s0 = _lib.stats()
        log0 = len(p.solver_stats["pcg_log"])
        t1 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = p.step_PGD(st)
        e1.record()
        torch.cuda.synchronize()
        s1 = _lib.stats()
        steps.append({"ms": e0.elapsed_time(e1), "wall_s": time.perf_counter() - t1, "fp_iterations": p.num_fp_it[-1] if p.num_fp_it else None,
                      "pcg_solves": s1["pcg_solves"] - s0["pcg_solves"], "pcg_iters": s1["pcg_iters"] - s0["pcg_iters"],
                      "pcg_ms": s1["pcg_ms"] - s0["pcg_ms"], "launches": s1["launches"] - s0["launches"],
                      "pcg_iterations": list(p.solver_stats["pcg_log"][log0:])})
        if done:
            break
    ds = p.V[0]._dev["device_space"]
    n, nnz = ds.n_owned, ds.nnz_owned  # partitioned space: this rank's rows
    s = _lib.stats()
    it_bytes = 12 * nnz + 4 * (n + 1) + 56 * n
    it_ms = s["pcg_ms"] / max(s["pcg_iters"], 1)
    out = {"config": a.config, "world": world, "global_spatial_dofs": p.V[0].n_dofs, "partitioned": ds.shard is not None, "sharded_solves": p.solver_stats.get("sharded_solves", 0), "spatial_dofs": n, "nnz": nnz, "dims": [v.n_dofs for v in p.V], "modes": p.PGD_modes,
           "build_problem_s": t_build, "steps": steps, "amplitude": [float(x) for x in p.amplitude],
           "pcg": {"iterations": s["pcg_iters"], "ms_per_iteration": it_ms, "bytes_per_iteration": it_bytes,
                   "gbs": it_bytes / (it_ms * 1e-3) / 1e9 if s["pcg_iters"] else None,
                   "frac_hbm": it_bytes / (it_ms * 1e-3) / 1e9 / peak if s["pcg_iters"] else None,
                   "share_of_step_time": s["pcg_ms"] / max(sum(x["ms"] for x in steps), 1e-9)},
           "enrichment_steps_per_s": len(steps) / (sum(x["ms"] for x in steps) * 1e-3),
           "device_mem_gb": torch.cuda.max_memory_allocated() / 1e9,  # peak, including set-up transients (facet sort, pattern sort)
           "device_mem_resident_gb": torch.cuda.memory_allocated() / 1e9}
    if rank == 0:
        print(json.dumps(out))
        if a.counts:
            with open(a.counts, "w") as f:
                json.dump({"config": a.config, "n": a.n, "modes": p.PGD_modes, "world": world, "rtol": a.rtol,
                           "amplitude": [float(x) for x in p.amplitude],
                           "steps": [{"fp_iterations": st_["fp_iterations"], "pcg_iterations": st_["pcg_iterations"],
                                      "ms": st_["ms"]} for st_ in steps]}, f)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
