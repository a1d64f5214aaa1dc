"""Repeat the end-to-end leg of bench.py and print per-repeat device time, wall time and collector activity."""
import gc, sys, time
sys.path.insert(0, "/root/repo")
import torch
from pgdrome_b200 import configs, _lib
acc = {"n": 0, "t": 0.0, "t0": 0.0}
tim = {}
def wrap(name):
    f = getattr(_lib, name)
    def g(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize()
        tim[name] = tim.get(name, 0.0) + time.perf_counter() - t
        return r
    setattr(_lib, name, g)
for nm in ("pattern_build", "vecmap_build", "p1_rowplan_build", "to_device", "banded_solve", "pcg"):
    wrap(nm)
def cb(phase, info):
    if info["generation"] == 2:
        if phase == "start": acc["t0"] = time.perf_counter()
        else: acc["n"] += 1; acc["t"] += time.perf_counter() - acc["t0"]
gc.callbacks.append(cb)
w = configs.heat2d_tk(PGD_nmax=3, PGD_tol=0.0); w.solve_PGD(_problem="linear")
torch.cuda.synchronize()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    tm = time.perf_counter()
    q = configs.heat2d_tk(PGD_nmax=10, PGD_tol=0.0)
    tm = time.perf_counter() - tm
    torch.cuda.synchronize()
    acc["n"], acc["t"] = 0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.process_time(); t = time.perf_counter(); e0.record()
    q.solve_PGD(_problem="linear")
    modes = [[f.vector().get_local() for f in q.PGD_func[d]] for d in range(3)]
    e1.record(); torch.cuda.synchronize()
    ms_ = torch.cuda.memory_stats()
    print("    device allocs %d frees %d (cumulative), reserved %.0f MB, host pinned stage %s" % (
        ms_.get("num_device_alloc", 0), ms_.get("num_device_free", 0), ms_.get("reserved_bytes.all.current", 0) / 1e6, list(_lib._stage)))
    st = _lib.stats(reset=True)
    print("   ", {k: round(1e3 * v, 1) for k, v in tim.items()}); tim.clear()
    print("rep %d: device %.1f ms wall %.1f ms (make %.1f ms) full collections %d (%.1f ms) pcg_iters %d pcg_ms %.1f solves %d cpu %.1f ms" % (
        rep, e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t), 1e3 * tm, acc["n"], 1e3 * acc["t"], st["pcg_iters"], st["pcg_ms"],
        st["pcg_solves"], 1e3 * (time.process_time() - c0)), flush=True)
