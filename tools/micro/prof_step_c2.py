"""cProfile of enrichment steps of configs[2] at full size (where does the non-PCG time of a step go?)."""
import cProfile, pstats, sys, io, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pgdrome_b200 import configs, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 68
p = configs.elasticity3d(n=n, PGD_nmax=6, PGD_tol=0.0)
st = p.begin_PGD(_problem="linear", settings={"linear_solver": "cg"})
for _ in range(2): p.step_PGD(st)
torch.cuda.synchronize()
# synchronous launches: the host time of every call includes its device time (attribution by Python function)
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
pr = cProfile.Profile()
_lib.stats(reset=True)
t=time.perf_counter()
pr.enable()
for _ in range(2): p.step_PGD(st)
torch.cuda.synchronize()
pr.disable()
print("wall", time.perf_counter()-t, "fp", p.num_fp_it[-2:], _lib.stats())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
print(s.getvalue()[:14000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
