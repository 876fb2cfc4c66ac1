"""Does the fused forward's duration depend on where y lands relative to x?  Times layer(x) with y at shifted addresses."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
from deepsphere import gnn_layers
from deepsphere.graph import SphereHealpix
import deepsphere.gnn_layers as gl
g = SphereHealpix(256, k=8)
gl.eigsh = lambda L, **kw: [1.85]
layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=64, mode="tf32")
B, M, F = 32, g.L.shape[0], 64
x = torch.randn(B, M, F, device="cuda")
with torch.no_grad():
    for _ in range(2): y = layer(x)
    torch.cuda.synchronize()
    keep = []
    for i, pad_mb in enumerate([0, 0, 1, 2, 3, 7, 64, 100, 257, 513, 1025, 0, 2049, 33, 0]):
        if pad_mb:
            keep.append(torch.empty(pad_mb * 1024 * 1024 // 4, device="cuda"))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y = layer(x); b.record(); torch.cuda.synchronize()
        d = y.data_ptr() - x.data_ptr()
        print(f"pad {pad_mb:5d} MB  y-x = {d / 2**20:10.2f} MiB  (mod 2MiB {d % 2**21:8d}, mod 201MB {d % (M*F*4):10d})  {a.elapsed_time(b):7.2f} ms")
        keep.append(y)  # keep y alive so the next one lands elsewhere
