"""What the HBM delivers for read-only vs copy traffic on this box (context for the SpMV roofline)."""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
dev = "cuda"
n = 1 << 27  # 1 GiB of fp64
x = torch.rand(n, dtype=torch.float64, device=dev)
y = torch.empty_like(x)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: torch.sum(x)); print("torch.sum   read  %.0f GB/s" % (8 * n / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy_       r+w   %.0f GB/s" % (16 * n / ms / 1e6))
ms = t(lambda: torch.dot(x, y)); print("torch.dot   read  %.0f GB/s" % (16 * n / ms / 1e6))
ms = t(lambda: y.zero_()); print("zero_       write %.0f GB/s" % (8 * n / ms / 1e6))
from pgdrome_b200 import _lib
ms = t(lambda: _lib.lincomb([x, y], [1.0, 2.0], out=y)); print("pgd_lincomb 2r+1w %.0f GB/s" % (24 * n / ms / 1e6))
ms = t(lambda: _lib.dot(x, y)); print("pgd_dot     read  %.0f GB/s" % (16 * n / ms / 1e6))
