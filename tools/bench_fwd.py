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
n = int(os.environ.get("ITERS", "8"))
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
evs[0].record()
for i in range(n):
    step()
    evs[i + 1].record()
torch.cuda.synchronize()
ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
import subprocess
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print("RESULT", mode, B, "bwd" if bwd else "fwd", "ms", " ".join(f"{t:.2f}" for t in ts), "| clk", clk, "env", {k: v for k, v in os.environ.items() if k.startswith("DEEPSPHERE")})
