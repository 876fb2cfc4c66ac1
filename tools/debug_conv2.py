"""Debug driver for ds_lattice_conv2.cu: isolates each T_k through a selector kernel and reports errors."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
sys.path.insert(0, ROOT)
from deepsphere import gnn_layers  # noqa: E402
from deepsphere.graph import SphereHealpix  # noqa: E402
from oracle import deepsphere_oracle as orc  # noqa: E402

nside = int(os.environ.get("NSIDE", "32"))
K = int(os.environ.get("K", "5"))
F = int(os.environ.get("F", "16"))
B = int(os.environ.get("B", "2"))
cls = os.environ.get("CLS", "Chebyshev")
g = SphereHealpix(nside, k=8)
M = g.L.shape[0]
rng = np.random.default_rng(0)
x = rng.standard_normal((B, M, F))
layer = getattr(gnn_layers, cls)(L=g.L, K=K, Fout=F, mode="tf32")
layer.build_from_shape((B, M, F))
Lt, _ = orc.prepare_laplacian(g.L, 0.75 if cls == "Chebyshev" else 1.0)
xt = torch.tensor(x, dtype=torch.float32, device="cuda")
for k in range(K):
    w = np.zeros((F * K, F))
    for f in range(F):
        w[f * K + k, f] = 1.0
    with torch.no_grad():
        layer.kernel.copy_(torch.tensor(w, dtype=torch.float32))
        y = layer(xt).cpu().numpy()
    ref = orc.graph_conv_forward(x, Lt, w, K, cls.lower(), dtype=np.float64)
    err = np.abs(y - ref).max(axis=(0, 2))  # per pixel
    scale = np.abs(ref).max()
    bad = np.flatnonzero(err > 2e-3 * scale)
    print(f"k={k}: max rel err {err.max() / scale:.3e}; bad pixels {len(bad)} / {M}; first {bad[:12]}")
    if len(bad):
        tiles = np.unique(bad // 256)
        print("   bad tiles", len(tiles), tiles[:20], " in-tile offsets of first tile:", (bad[bad // 256 == tiles[0]] % 256)[:40])
        p = bad[0]
        print("   y", y[0, p, :4], "ref", ref[0, p, :4])
print("lattice", layer._plan.info(0)["lattice"])
