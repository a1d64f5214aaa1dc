"""Kernel rooflines on a mesh larger than L2 (the HBM-bound regime of BASELINE configs[2]-[4]).

Every kernel: >= 3 warm-ups, 10 timed launches with CUDA events on the launching stream, inputs
(CSR arrays 12 B/nnz) far larger than the 126 MB L2.  ``achieved`` uses the ALGORITHMIC bytes of
SURVEY.md 8(d) / DESIGN.md, not measured traffic.  Called by bench.py (key ``kernels``) and usable
stand-alone:  python -m tools.kernel_bench --mesh 158
"""
import json
import sys
import time

import numpy as np
import torch


PROFILE = False  # --profile: one warm-up + one timed launch per kernel (for ncu captures)


def _time(fn, reps=10, warm=3):
    if PROFILE:
        reps, warm = 1, 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    return ms[len(ms) // 2]


def run(mesh_n=128, hbm_peak=6451.2, evaluate=True, fp64_peak=None):
    from pgdrome_b200 import _lib, fem
    from pgdrome_b200.assembly import device_space

    out = {"mesh": "BoxMesh %d^3 cells, P1" % mesh_n, "hbm_peak_gbs": hbm_peak}
    t0 = time.perf_counter()
    m = fem.UnitCubeMesh(mesh_n, mesh_n, mesh_n)
    V = fem.FunctionSpace(m, "P", 1)
    out["host_mesh_s"] = time.perf_counter() - t0
    ds = device_space(V)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rowptr, colidx, gptr, gidx = ds.pattern
    torch.cuda.synchronize()
    out["pattern_build_s"] = time.perf_counter() - t0
    n, nnz, nc = ds.n_dofs, ds.nnz, m.num_cells()
    out.update(n_dofs=n, nnz=nnz, n_cells=nc)
    dev = rowptr.device
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1
    y = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1

    def entry(name, ms, nbytes, **kw):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = dict(ms=ms, bytes=nbytes, gbs=gbs, frac_hbm=gbs / hbm_peak, **kw)

    # fused P1 operator  A = cm M + ck K  straight into the CSR pattern
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    ms = _time(lambda: _lib.assemble_p1(ds.coords, ds.cell_verts, 3, 0.3, 1.7, None, gptr, gidx, nnz, out=vals))
    entry("assemble_p1_fused", ms, 4 * 4 * nc + 8 * 3 * m.num_vertices() + 8 * nnz,
          note="algorithmic bytes exclude the 8 B/nnz + 4 B/contribution gather list the kernel also reads")
    t0 = time.perf_counter()
    plan = ds.rowplan
    torch.cuda.synchronize()
    out["rowplan_build_s"] = time.perf_counter() - t0
    vptr = ds.vecmap[0]
    soa = ds.coords_soa
    ms = _time(lambda: _lib.assemble_p1_rows(ds.coords, ds.cell_verts, 3, 0.3, 1.7, None, rowptr, vptr, plan, n, out=vals,
                                             coords_soa=soa))
    entry("assemble_p1_rows", ms, 4 * 4 * nc + 8 * 3 * m.num_vertices() + 8 * nnz,
          note="row-owner kernel; also reads the 8 B/(cell,vertex) plan (%d MB) and re-reads cell vertices 4x through L1/L2"
               % (8 * 4 * nc // 1000000))
    # generic element kernel + gather (used once per atom at set-up)
    T = np.zeros((1, 4, 1, 4))
    for k in range(1, 4):
        T[0, k, 0, k] = 1.0
    Tg = T.copy()
    Tg[0, 1, 0, 0] = 1e-300  # a (d_x v) u term is outside the closed form => general-tensor row-owner kernel
    ms = _time(lambda: ds.assemble_bilinear(Tg), reps=5, warm=2)
    entry("assemble_p1_tensor", ms, 4 * 4 * nc + 8 * 3 * m.num_vertices() + 8 * nnz, kernel="k_assemble_p1_tensor<3,1>")
    plan_keep, ds._node_plan = ds.node_plan, False  # without the node plan the same atom takes the element-matrix route
    ms = _time(lambda: ds.assemble_bilinear(Tg), reps=3, warm=1)
    ds._node_plan = plan_keep
    entry("assemble_atom_generic", ms, 4 * 4 * nc + 8 * 3 * m.num_vertices() + 8 * nnz, kernel="k_elem_bilinear + k_gather_sum")
    K = ds.assemble_bilinear(T)
    Tm = np.zeros((1, 4, 1, 4))
    Tm[0, 0, 0, 0] = 1.0
    M = ds.assemble_bilinear(Tm)
    # A = sum_k c_k K_k over value arrays
    ms = _time(lambda: _lib.lincomb([K, M], [1.7, 0.3], out=vals))
    entry("lincomb_2atoms", ms, 3 * 8 * nnz)
    lpr = ds.lpr
    yv = torch.empty(n, dtype=torch.float64, device=dev)
    spmv_bytes = 12 * nnz + 4 * (n + 1) + 16 * n
    for opt, name in ((0, "spmv_subwarp_per_row"), (1, "spmv_regstaged_rowblocks"), (2, "spmv")):
        _lib.set_option("spmv_stream", opt)
        ms = _time(lambda: _lib.spmv(rowptr, colidx, vals, x, y=yv, lpr=lpr))
        entry(name, ms, spmv_bytes, lanes_per_row=lpr, kernel=["k_spmv", "k_spmv_stream", "k_spmv_bulk (TMA ring)"][opt])
    sc = torch.empty(1, dtype=torch.float64, device=dev)
    ms = _time(lambda: _lib.bilinear(rowptr, colidx, vals, x, y, out=sc, lpr=lpr))
    entry("bilinear_functional", ms, spmv_bytes, lanes_per_row=lpr)
    # PCG iteration (device time from the library's own events)
    b = _lib.spmv(rowptr, colidx, vals, x, lpr=lpr)
    work = torch.empty(6 * n + 8, dtype=torch.float64, device=dev)
    _lib.pcg(rowptr, colidx, vals, b, rtol=1e-30, maxit=20, check_every=20, lpr=lpr, work=work)
    _lib.stats(reset=True)
    _lib.pcg(rowptr, colidx, vals, b, rtol=1e-30, maxit=100, check_every=100, lpr=lpr, work=work)
    s = _lib.stats()
    entry("pcg_iteration", s["pcg_ms"] / max(s["pcg_iters"], 1), 12 * nnz + 4 * (n + 1) + 56 * n, iters=s["pcg_iters"])
    # panel dots: 16 cached K U_i products against one vector
    P = torch.rand((16, n), dtype=torch.float64, device=dev, generator=g)
    res = torch.empty(16, dtype=torch.float64, device=dev)
    ms = _time(lambda: _lib.panel_dots(P, 16, x, out=res))
    entry("panel_dots_16", ms, 8 * n * 16 + 8 * n)
    del P, K, M, vals, work
    # sensor point location: one pass over the cells for 64 sensors
    pts = torch.rand((64, 3), dtype=torch.float64, device=dev, generator=g)
    ms = _time(lambda: _lib.locate_points(ds.coords, ds.cell_verts, pts))
    entry("locate_points_64", ms, 4 * 4 * nc + 8 * 3 * m.num_vertices(), kernel="k_locate<3> (+ fill, finalise)")
    if evaluate:
        R, N, C = 50, 100000, 10000
        X = torch.randn((R, N), dtype=torch.float64, device=dev, generator=g)
        Wt = torch.randn((R, C), dtype=torch.float64, device=dev, generator=g)
        U = torch.empty((C, N), dtype=torch.float64, device=dev)
        ms = _time(lambda: _lib.eval_gemm(Wt, X, R, out=U), reps=5, warm=2)
        flops = 2.0 * N * C * R
        if fp64_peak is None:  # ONE definition of the FP64 peak: cuBLAS DGEMM 8192^3 measured on this pool
            import os

            try:
                fp64_peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                                        "r01_fp64_peak.json")))["fp64_gemm_tflops"]
            except Exception:
                a = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
                fp64_peak = 2.0 * 8192**3 / (_time(lambda: torch.matmul(a, a), reps=3, warm=1) * 1e-3) / 1e12
        peak = fp64_peak
        tf = flops / (ms * 1e-3) / 1e12
        out["evaluate_gemm_f64"] = dict(ms=ms, N=N, C=C, R=R, tflops=tf, fp64_peak_tflops=peak, frac_fp64=tf / peak,
                                        out_write_gbs=8.0 * N * C / (ms * 1e-3) / 1e9,
                                        frac_hbm_outwrite=8.0 * N * C / (ms * 1e-3) / 1e9 / hbm_peak,
                                        peak_source="cuBLAS DGEMM 8192^3 (profiles/r01_fp64_peak.json)")
        del U, X, Wt
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import argparse
    import os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=128)
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    PROFILE = a.profile
    print(json.dumps(run(a.mesh)))
