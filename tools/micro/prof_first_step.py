"""cProfile of the FIRST enrichment step of a fresh configs[1] problem (the set-up share of the end-to-end leg)."""
import cProfile, pstats, sys, io, time
sys.path.insert(0, "/root/repo")
import torch
from pgdrome_b200 import configs, _lib
w = configs.heat2d_tk(PGD_nmax=2, PGD_tol=0.0); w.solve_PGD(_problem="linear"); w = None
torch.cuda.synchronize()
for rep in range(2):
    p = configs.heat2d_tk(PGD_nmax=3, PGD_tol=0.0)
    st = p.begin_PGD(_problem="linear")
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    p.step_PGD(st)
    torch.cuda.synchronize()
    pr.disable()
    t1 = time.perf_counter() - t
    t = time.perf_counter()
    p.step_PGD(st)
    torch.cuda.synchronize()
    print("rep", rep, "first step %.1f ms, second step %.1f ms" % (1e3 * t1, 1e3 * (time.perf_counter() - t)), p.num_fp_it)
    p = st = None
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
