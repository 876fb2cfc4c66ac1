"""Small driver for ncu captures: runs the HealpyChebyshev layer forward (+ backward) a few times
at a reduced batch so that kernel replay stays cheap.  Usage: profile_layer.py [mode] [batch] [bwd]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
from deepsphere import gnn_layers  # noqa: E402
from deepsphere.graph import SphereHealpix  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
nside = int(os.environ.get("NSIDE", "256"))
F = int(os.environ.get("FEATURES", "64"))
g = SphereHealpix(nside, k=8)
# lmax of the normalised 8-neighbour Laplacian only scales L~; skip ARPACK for profiling runs
import deepsphere.gnn_layers as gl  # noqa: E402
gl.eigsh = lambda L, **kw: [1.85]
layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=F, mode=mode)
x = torch.randn(B, g.L.shape[0], F, device="cuda", requires_grad=bwd)
for _ in range(2):
    y = layer(x)
    if bwd:
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done", mode, B, bwd)
