"""What does this B200 sustain for simple streaming kernels at the bench's tensor sizes?
(context for the roofline fractions: copy / triad via torch, and the hop kernel on an identity graph)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
from deepsphere import _ops, utils  # noqa: E402
from scipy import sparse  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for gb in (1.6, 6.4):
    n = int(gb * 1e9 / 4)
    x = torch.randn(n, device="cuda")
    y = torch.randn(n, device="cuda")
    z = torch.empty_like(x)
    t = timeit(lambda: z.copy_(x))
    print(f"{gb} GB copy   : {t:.3f} ms  {2 * n * 4 / t / 1e6:.0f} GB/s")
    t = timeit(lambda: torch.add(x, y, out=z))
    print(f"{gb} GB add    : {t:.3f} ms  {3 * n * 4 / t / 1e6:.0f} GB/s")
    t = timeit(lambda: torch.sum(x))
    print(f"{gb} GB sum    : {t:.3f} ms  {n * 4 / t / 1e6:.0f} GB/s (read only)")
    t = timeit(lambda: z.fill_(1.0))
    print(f"{gb} GB fill   : {t:.3f} ms  {n * 4 / t / 1e6:.0f} GB/s (write only)")
    del x, y, z

M, F, B = 786432, 64, 32
plan = utils.plan_from_sparse(sparse.identity(M, format="csr") * 0.5)
x = torch.randn(B, M, F, device="cuda")
p = torch.randn(B, M, F, device="cuda")
t = timeit(lambda: _ops.spmm(plan, x), 5)
print(f"hop identity graph, no prev : {t:.3f} ms  {2 * x.numel() * 4 / t / 1e6:.0f} GB/s")
t = timeit(lambda: _ops.spmm(plan, x, 2.0, p, -1.0), 5)
print(f"hop identity graph, prev    : {t:.3f} ms  {3 * x.numel() * 4 / t / 1e6:.0f} GB/s")
