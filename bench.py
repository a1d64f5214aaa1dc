#!/usr/bin/env python
"""bench.py -- PGD enrichment throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's B200 path
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle port)

A *step* is one enrichment step of the progressive PGD (pgdrome/solver.py:325-504: initial modes,
residual check, alternating fixed-point solve to convergence, normalisation) on BASELINE
configs[1]: 2-D transient heat, P1 on a 256x256 unit square (66 049 dofs) x 200 time nodes x 50
conductivity nodes.  W warm-up steps (they also build the sparsity pattern and the separated-form
atoms), then exactly K timed steps between barrier + synchronize, CUDA events, max over ranks.
With N > 1 every rank enriches its own replica of the problem (configs[1] is the reference's
single-GPU case: "replicas only") and `value` is the job total.

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PGD enrichment iters/sec at N dofs"
UNIT = "enrichment_steps/s"
WORKLOAD = dict(n=256, nt=199, nk=49)


def _config(args, extra=None):
    n = args.n
    c = {"workload": "configs[1]: heat2d_tk P1 %dx%d unit square (%d dofs) x %d time nodes (FD) x %d k nodes" % (
        n, n, (n + 1) ** 2, args.nt + 1, args.nk + 1), "spatial_dofs": (n + 1) ** 2, "tol_fp_it": 1e-5, "max_fp_it": 50,
        "stop_fp": "norm", "norm_modes": "stiff", "linear_solver": "Jacobi-PCG rtol 1e-13 (space) + banded LU (t, k)",
        "l2": "flushed between steps (256 MiB write); per-solve working set 8 MB is L2-resident by construction of the config",
        "parallelism": "replicas" if args.gpus > 1 else "single"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        # under load = upper half of the samples (the sampler also sees the idle gaps of host-side work)
        med = sm[(3 * len(sm)) // 4] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- oracle (CPU) legs
def _oracle_problem(args, nmax):
    from oracle import problems as oprob

    o, info = oprob.heat2d_tk(n=args.n, nt=args.nt, nk=args.nk, PGD_nmax=nmax, PGD_tol=0.0)
    return o


def _oracle_steps(args, n_warm, n_steps):
    """Wall-clock seconds of enrichment steps n_warm .. n_warm+n_steps-1 of the oracle port
    (NumPy/SciPy: COO->CSR assembly done before, SuperLU solves = the reference's default LU)."""
    from oracle import pgd as opgd

    o = _oracle_problem(args, n_warm + n_steps)
    marks = [time.perf_counter()]
    opgd.solve_pgd(o, step_hook=lambda n: marks.append(time.perf_counter()))
    done = len(marks) - 1
    if done <= n_warm:
        return 0.0, 0
    return marks[done] - marks[n_warm], done - n_warm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scipy

    nmax = 20  # the workload's mode budget (BASELINE configs[1]); a longer request is timed on this bounded sample
    w = min(args.warmup, nmax - 1)
    t, k = _oracle_steps(args, w, min(args.steps, nmax - w))
    val = k / t if t > 0 else 0.0
    cores = 1
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
           "warmup": args.warmup, "ms_per_step": 1e3 * t / max(k, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(args),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": "enrichment steps %d..%d of the same workload, oracle port (SciPy %s SuperLU), 1 thread"
                                      % (w, w + k - 1, scipy.__version__)},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "the real reference (FEniCS 2019.1 + PETSc/MUMPS) cannot be installed in this image; this is its CPU port"}
    print(json.dumps(out))


# ------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback. Use --impl reference for the CPU port.")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from pgdrome_b200 import _lib, configs, lazy

    W, K = args.warmup, args.steps
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def make(nmax):
        p = configs.heat2d_tk(n=args.n, nt=args.nt, nk=args.nk, PGD_nmax=nmax, PGD_tol=0.0)
        return p

    # ---------------- device-resident arm: W warm-up + K timed enrichment steps
    # The workload enriches up to NMAX = 20 modes (BASELINE configs[1]).  Runs with W + K > NMAX continue on a fresh
    # problem instance (its W warm-up steps untimed again), so that every timed step is a real enrichment step and
    # never the cheap "residual below 1e-10, stop" exit of an exhausted enrichment.
    NMAX = 20
    W = min(W, NMAX - 1)

    def fresh():
        q = make(NMAX)
        s_ = q.begin_PGD(_problem="linear")
        for _ in range(W):
            q.step_PGD(s_)
        return q, s_

    p, st = fresh()
    ds0 = p.V[0]._dev["device_space"]
    n, nnz = ds0.n_dofs, ds0.nnz
    barrier()
    _lib.stats(reset=True)
    flushes0 = lazy.stats["flushes"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
    barrier()
    t0 = time.perf_counter()
    step_ms, fp_its = [], []
    done_steps, in_problem, warm_stats = 0, W, None
    while done_steps < K:
        if in_problem >= NMAX or st["done"]:
            s_before = _lib.stats()
            p, st = fresh()  # untimed: set-up and warm-up of the next problem instance
            s_after = _lib.stats()
            warm_stats = {k: (warm_stats[k] if warm_stats else 0) + s_after[k] - s_before[k] for k in s_after}
            in_problem = W
        i = done_steps
        flush_buf.fill_(i & 0xFF)  # L2 flush between timed steps (untimed)
        ev[2 * i].record()
        p.step_PGD(st)
        ev[2 * i + 1].record()
        if st["done"] and len(p.num_fp_it) < in_problem + 1:
            continue  # the step only detected an exhausted residual: not an enrichment step, not counted
        fp_its.append(p.num_fp_it[-1])
        done_steps += 1
        in_problem += 1
    barrier()
    wall = time.perf_counter() - t0
    step_ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)]
    clocks = sampler.stop() if rank == 0 else None
    s = _lib.stats()
    if warm_stats:  # library counters of the untimed re-warm-ups do not belong to the timed steps
        s = {k: s[k] - warm_stats[k] for k in s}
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * K / (total_ms * 1e-3)

    # live roofline of the dominant kernel group: one Jacobi-PCG iteration (k_pcg_spmv + k_pcg_update)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    it_bytes = 12 * nnz + 4 * (n + 1) + 56 * n
    it_ms = s["pcg_ms"] / max(s["pcg_iters"], 1)
    achieved = it_bytes / (it_ms * 1e-3) / 1e9 if s["pcg_iters"] else 0.0
    resident = s.get("pcg_resident_solves", 0) >= s["pcg_solves"] > 0
    traffic = None
    try:  # DRAM bytes of the dominant kernel from the committed ncu --set full capture (per launch = per solve)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_pcg_resident_traffic.json")))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "hbm",
                "kernel": ("Jacobi-PCG iteration inside k_pcg_resident (one cooperative launch per solve, matrix slice "
                           "resident in shared memory; figures are per ITERATION)") if resident else
                          "Jacobi-PCG iteration (k_pcg_spmv_bulk + k_pcg_update)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if resident else None,
                "traffic_note": ("ncu dram__bytes_read+write of ONE k_pcg_resident launch = one whole solve (~660 iterations): the "
                                 "matrix is read from HBM once per solve; a streaming PCG would move bytes_per_launch per "
                                 "iteration") if resident else None,
                "bytes_per_launch": it_bytes, "us_per_launch": 1e3 * it_ms, "launches": s["pcg_iters"],
                "share_of_step": s["pcg_ms"] / total_ms if total_ms else None,
                "note": "matrix (5.5 MB) + vectors are on-chip at this config (shared memory / L2): the iteration is bound by "
                        "two grid barriers (~2 us each), not by HBM; `kernels` carries the HBM-bound sizes (128^3 mesh: "
                        "SpMV 80 %, PCG iteration 63 % of the measured HBM peak)"}

    # ---------------- end-to-end arm: host arrays in, modes out, everything inside the timed region
    e2e = None
    if rank == 0 or world > 1:
        Ke = min(K, args.e2e_steps) if args.e2e_steps else min(K, NMAX)
        p = st = None  # the device-resident problem is done: its memory goes back to the allocator before the next one
        q = make(Ke)
        for V in q.V:
            V._dev.pop("device_space", None)
        _lib.traffic["h2d"] = _lib.traffic["d2h"] = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.perf_counter()
        e0.record()
        q.solve_PGD(_problem="linear")
        modes = [[f.vector().get_local() for f in q.PGD_func[d]] for d in range(3)]
        e1.record()
        barrier()
        tw = time.perf_counter() - tw
        ms = e0.elapsed_time(e1)
        te = torch.tensor([ms], dtype=torch.float64, device="cuda")
        per_rank = [ms]
        if dist is not None:
            allms = [torch.zeros_like(te) for _ in range(world)]
            dist.all_gather(allms, te)
            per_rank = [float(t.item()) for t in allms]
        ms = max(per_rank)
        e2e = {"value": world * q.PGD_modes / (ms * 1e-3), "unit": UNIT, "steps": q.PGD_modes,
               "h2d_bytes_per_step": _lib.traffic["h2d"] // max(q.PGD_modes, 1),
               "d2h_bytes_per_step": _lib.traffic["d2h"] // max(q.PGD_modes, 1), "ms_total": ms,
               "ms_per_rank": per_rank, "host_wall_ms_rank0": tw * 1e3,
               "what": "fresh PGDProblem from host (NumPy) mesh/dofmap arrays -> solve_PGD(%d modes) -> all modes read back "
                       "to host; includes mesh upload, pattern build, atom assembly" % Ke}
        del modes

    sharded = None
    if world > 1 and not args.no_kernels:
        # the path's real exchange step (configs[2]-[3]): a spatial Jacobi-PCG sharded by rows over all ranks,
        # NCCL send/recv halo of p + allreduce of the dot products inside libpgdb200's loop
        try:
            from tools import sharded_bench

            sharded = sharded_bench.run(args.kernel_mesh, iters=200, hbm_peak=peak)
        except Exception as e:
            sharded = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline (bounded sample of the same workload) and large-mesh kernel rooflines
    cpu = None
    if world == 1 and not args.no_cpu:
        import scipy

        tc, kc = _oracle_steps(args, 0, args.cpu_steps)
        cpu = {"value": kc / tc if tc > 0 else 0.0, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "first %d enrichment steps of the same workload, oracle port (SciPy %s, SuperLU direct solves)"
                         % (kc, scipy.__version__), "seconds": tc}
    kernels = None
    if world == 1 and not args.no_kernels:
        try:
            from tools import kernel_bench

            kernels = kernel_bench.run(args.kernel_mesh, peak)
        except Exception as e:  # the headline number must not depend on the micro-benchmarks
            kernels = {"error": repr(e)}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": _config(args), "clocks": clocks, "e2e": e2e, "gpu_launches": s["launches"],
           "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels, "sharded_pcg": sharded,
           "detail": {"step_ms": step_ms, "pcg_solves": s["pcg_solves"], "pcg_iters": s["pcg_iters"], "pcg_ms": s["pcg_ms"],
                      "fp_iterations": fp_its, "functional_flushes": lazy.stats["flushes"] - flushes0,
                      "wall_s_timed_region": wall, "nnz": nnz}}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=WORKLOAD["n"])
    ap.add_argument("--nt", type=int, default=WORKLOAD["nt"])
    ap.add_argument("--nk", type=int, default=WORKLOAD["nk"])
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=0)
    ap.add_argument("--kernel-mesh", type=int, default=128, help="cells per edge of the box mesh of the kernel rooflines")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        sys.stderr.write("note: timing rules ask for >= 3 warm-up steps\n")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
