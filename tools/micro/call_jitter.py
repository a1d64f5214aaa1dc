"""Host-side jitter of the blocking C-ABI calls: wall time minus device time per call, percentiles."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pgdrome_b200 import _lib, fem
from pgdrome_b200.assembly import device_space

m = fem.UnitSquareMesh(256, 256)
V = fem.FunctionSpace(m, "P", 1)
ds = device_space(V)
rowptr, colidx, gptr, gidx = ds.pattern
T = np.zeros((1, 3, 1, 3)); T[0, 0, 0, 0] = 1.0; T[0, 1, 0, 1] = T[0, 2, 0, 2] = 1.0
vals = ds.assemble_bilinear(T)
n = ds.n_dofs
b = torch.ones(n, dtype=torch.float64, device=vals.device)
work = torch.empty(6 * n + 8, dtype=torch.float64, device=vals.device)
x = torch.zeros_like(b)
def pct(a):
    a = np.sort(np.array(a)); return "p50 %.3f p90 %.3f p99 %.3f max %.3f ms" % (a[len(a)//2], a[int(.9*len(a))], a[int(.99*len(a))], a[-1])
for maxit in (20, 600):
    over, dev = [], []
    for i in range(300):
        torch.cuda.synchronize()
        _lib.stats(reset=True)
        t = time.perf_counter()
        _lib.pcg(rowptr, colidx, vals, b, rtol=1e-30, maxit=maxit, check_every=maxit, lpr=ds.lpr, work=work, x=x)
        w = 1e3 * (time.perf_counter() - t)
        d = _lib.stats()["pcg_ms"]
        over.append(w - d); dev.append(d)
    print("pcg maxit=%d: device %s | wall-device %s" % (maxit, pct(dev), pct(over)), flush=True)
a = np.random.rand(66049)
o = []
for i in range(300):
    torch.cuda.synchronize(); t = time.perf_counter(); _lib.to_device(a); torch.cuda.synchronize(); o.append(1e3 * (time.perf_counter() - t))
print("to_device 528 KB:", pct(o))
o = []
r = torch.zeros(16, dtype=torch.float64, device=vals.device)
for i in range(300):
    torch.cuda.synchronize(); t = time.perf_counter(); _lib.to_host(r); o.append(1e3 * (time.perf_counter() - t))
print("to_host 128 B:", pct(o))
o = []
for i in range(300):
    torch.cuda.synchronize(); t = time.perf_counter(); torch.cuda.synchronize(); o.append(1e3 * (time.perf_counter() - t))
print("synchronize:", pct(o))
