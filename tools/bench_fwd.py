"""Times the fused forward (and optionally fwd+bwd) of the HealpyChebyshev bench layer; prints ms."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
from deepsphere import gnn_layers
from deepsphere.graph import SphereHealpix
import deepsphere.gnn_layers as gl
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
bwd = len(sys.argv) > 3 and sys.argv[3] == "bwd"
F = int(os.environ.get("FEATURES", "64"))
g = SphereHealpix(256, k=8)
gl.eigsh = lambda L, **kw: [1.85]
layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=F, mode=mode)
x = torch.randn(B, g.L.shape[0], F, device="cuda", requires_grad=bwd)
dy = torch.randn(B, g.L.shape[0], F, device="cuda")
def step():
    y = layer(x)
    if bwd:
        x.grad = None; layer.kernel.grad = None
        y.backward(dy)
for _ in range(2): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
n = 5
for _ in range(n): step()
b.record(); torch.cuda.synchronize()
print("RESULT", mode, B, "bwd" if bwd else "fwd", "ms", a.elapsed_time(b) / n, "env", {k: v for k, v in os.environ.items() if k.startswith("DEEPSPHERE")})
