"""Measure the FP64 GEMM throughput (cuBLAS DGEMM via torch.matmul) used as the roofline
denominator of the evaluate kernel; MEASURED_PEAKS.json records no FP64 figure."""
import json
import sys

import torch


def measure(n=8192, reps=5):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n**3 / (best * 1e-3) / 1e12


if __name__ == "__main__":
    out = {"fp64_gemm_tflops": measure(), "how": "torch.matmul fp64 8192^3, best of 5, CUDA events",
           "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"))
