import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200")); sys.path.insert(0, ROOT)
from deepsphere import gnn_layers, healpix as hpx, utils
from deepsphere.graph import SphereHealpix
from oracle import deepsphere_oracle as orc
ext = utils.extend_indices(hpx.query_disc(64, [1, 0, 0], 1.2), 64, 8)
g = SphereHealpix(64, indexes=ext, k=8)
M = len(ext)
Lt, _ = orc.prepare_laplacian(g.L, 0.75)
for (K, Fin, Fout, act, bias) in [(4, 16, 32, "relu", True), (4, 16, 32, None, False), (5, 16, 16, None, False), (4, 32, 16, None, False)]:
    torch.manual_seed(1)
    layer = gnn_layers.Chebyshev(L=g.L, K=K, Fout=Fout, healpix=(64, ext), use_bias=bias, activation=act, mode="tf32")
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, M, Fin)); dy = rng.standard_normal((2, M, Fout))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    y = layer(xt); y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    xr = torch.tensor(x, requires_grad=True)
    wr = layer.kernel.detach().double().cpu().requires_grad_(True)
    z = orc.torch_cpu_graph_conv(xr, Lt, wr, K, "chebyshev")
    if bias: z = z + layer.bias.detach().double().cpu()
    yr = torch.relu(z) if act == "relu" else z
    yr.backward(torch.tensor(dy))
    ey = np.abs(y.detach().cpu().numpy() - yr.detach().numpy()).max(axis=(0, 2))
    ex = np.abs(xt.grad.cpu().numpy() - xr.grad.numpy()).max(axis=(0, 2))
    ek = np.abs(layer.kernel.grad.cpu().numpy() - wr.grad.numpy()).max() / np.abs(wr.grad.numpy()).max()
    sx = np.abs(xr.grad.numpy()).max()
    bad = np.flatnonzero(ex > 2e-3 * sx)
    print(f"K={K} Fin={Fin} Fout={Fout} act={act}: y err {ey.max()/np.abs(yr.detach().numpy()).max():.2e}  dx err {ex.max()/sx:.2e} bad rows {len(bad)}/{M}  dk err {ek:.2e}")
    if len(bad):
        tiles = np.unique(ext[bad] // 256)
        print("   bad tiles (nested/256):", len(tiles), tiles[:12], " rows in first:", (ext[bad][ext[bad] // 256 == tiles[0]] % 256)[:20])
