import cProfile, pstats, sys, io, time
sys.path.insert(0, "/root/repo")
import torch
from pgdrome_b200 import configs, _lib
t = time.perf_counter()
p = configs.thermal3d(n=100, PGD_nmax=1, PGD_tol=0.0)
print("build", time.perf_counter() - t)
st = p.begin_PGD(_problem="linear")
pr = cProfile.Profile()
t = time.perf_counter()
pr.enable()
p.step_PGD(st)
torch.cuda.synchronize()
pr.disable()
print("first step wall", time.perf_counter() - t, _lib.stats())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:8000])
