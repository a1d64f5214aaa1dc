import cProfile, pstats, sys, io, time
sys.path.insert(0, "/root/repo")
import torch
from pgdrome_b200 import configs, _lib
w = configs.heat2d_tk(PGD_nmax=2, PGD_tol=0.0); w.solve_PGD(_problem="linear")   # warm the process (module init, kernels)
torch.cuda.synchronize()
for rep in range(2):
    q = configs.heat2d_tk(PGD_nmax=5, PGD_tol=0.0)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    q.solve_PGD(_problem="linear")
    modes = [[f.vector().get_local() for f in q.PGD_func[d]] for d in range(3)]
    torch.cuda.synchronize()
    pr.disable()
    print("rep", rep, "wall", time.perf_counter() - t)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(40)
print(s.getvalue()[:7000])
