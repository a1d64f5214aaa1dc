"""cProfile of the set-up part of configs[2] at full size: problem construction, enrichment steps 0 and 1 (pattern build,
atom assembly, first panels) with synchronous launches so that device time is attributed to the calling Python function."""
import cProfile, pstats, sys, io, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import torch
from pgdrome_b200 import configs, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 68
torch.zeros(1, device="cuda")
for label, k in (("build+step0", 0), ("step1", 1), ("step2", 2)):
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    if k == 0:
        p = configs.elasticity3d(n=n, PGD_nmax=30, PGD_tol=0.0)
        st = p.begin_PGD(_problem="linear", settings={"linear_solver": "cg"})
    _lib.stats(reset=True)
    p.step_PGD(st)
    torch.cuda.synchronize()
    pr.disable()
    print("=====", label, "wall", time.perf_counter() - t, "fp", p.num_fp_it[-1:], _lib.stats())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print(s.getvalue()[:4500])
