"""cProfile of the end-to-end leg under torchrun (rank 0 prints): what is slower when NCCL is up?"""
import cProfile, pstats, sys, io, time, os
sys.path.insert(0, "/root/repo")
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(1, device="cuda"); dist.all_reduce(t)
from pgdrome_b200 import configs, _lib
w = configs.heat2d_tk(PGD_nmax=2, PGD_tol=0.0); w.solve_PGD(_problem="linear")
torch.cuda.synchronize()
for rep in range(2):
    q = configs.heat2d_tk(PGD_nmax=5, PGD_tol=0.0)
    torch.cuda.synchronize(); dist.barrier()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    q.solve_PGD(_problem="linear")
    modes = [[f.vector().get_local() for f in q.PGD_func[d]] for d in range(3)]
    torch.cuda.synchronize()
    pr.disable()
    if dist.get_rank() == 0:
        print("rep", rep, "wall", time.perf_counter() - t0, flush=True)
if dist.get_rank() == 0:
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25)
    print(s.getvalue()[:6000])
dist.destroy_process_group()
