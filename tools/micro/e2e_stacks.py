"""Where does the host thread sit during slow end-to-end repeats?  A sampler thread records the main thread's
innermost Python frames every 2 ms; per repeat we print the frames that gained the most samples."""
import collections, sys, threading, time
sys.path.insert(0, "/root/repo")
import torch
from pgdrome_b200 import configs, _lib
main_id = threading.main_thread().ident
samples = collections.Counter()
run = [True]
def sampler():
    while run[0]:
        f = sys._current_frames().get(main_id)
        if f is not None:
            k = []
            g = f
            for _ in range(3):
                if g is None: break
                k.append("%s:%d" % (g.f_code.co_filename.split("/")[-1], g.f_lineno)); g = g.f_back
            samples[" < ".join(k)] += 1
        time.sleep(0.002)
th = threading.Thread(target=sampler, daemon=True); th.start()
w = configs.heat2d_tk(PGD_nmax=3, PGD_tol=0.0); w.solve_PGD(_problem="linear")
torch.cuda.synchronize()
base = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    q = configs.heat2d_tk(PGD_nmax=10, PGD_tol=0.0)
    torch.cuda.synchronize()
    samples.clear()
    t = time.perf_counter()
    q.solve_PGD(_problem="linear")
    modes = [[f.vector().get_local() for f in q.PGD_func[d]] for d in range(3)]
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t)
    snap = collections.Counter(samples)
    if base is None or wall < base[0]:
        base = (wall, snap)
    print("rep %d wall %.1f ms, samples %d" % (rep, wall, sum(snap.values())), flush=True)
    if wall > 1.4 * base[0]:
        diff = collections.Counter({k: v - base[1].get(k, 0) for k, v in snap.items()})
        for k, v in diff.most_common(6):
            print("      +%d samples (x2 ms)  %s" % (v, k))
run[0] = False
