"""SURVEY 8f rows on the GPU, through the layer API -> ctypes -> C-ABI: Bernstein (= the Chebyshev kernels with
basis-changed weights) and HealpySmoothing (= ds_spmm), against the oracle's literal restatements of the
reference loops and the committed golden vectors.  fp32 bar: rel <= 1e-5 (max-abs error / max-abs reference)."""
import os

import numpy as np
import pytest
import torch

from deepsphere import _native as nat
from deepsphere import gnn_layers, healpy_layers as hl
from deepsphere.graph import SphereHealpix
from helpers import GOLDEN, load_golden, orc, rel_err

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-5


def dev(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


@pytest.mark.parametrize("name", ["bern_nside4_k8", "bern_masked16_k20"])
def test_bernstein_forward_backward_match_golden(name):
    g = load_golden(name)
    layer = gnn_layers.Bernstein(L=g["L"], K=g["K"], Fout=g["kernel"].shape[1], mode="fp32")
    layer.build_from_shape(g["x"].shape)
    with torch.no_grad():
        layer.kernel.copy_(dev(g["kernel"]))
    x = dev(g["x"]).requires_grad_(True)
    before = nat.launch_count()
    y = layer(x)
    assert nat.launch_count() > before
    assert rel_err(y.detach().cpu().numpy(), g["y64"]) <= TOL_FP32
    y.backward(dev(g["dy"]))
    assert rel_err(x.grad.cpu().numpy(), g["dx64"]) <= TOL_FP32
    assert rel_err(layer.kernel.grad.cpu().numpy(), g["dkernel64"]) <= TOL_FP32


def test_bernstein_bias_activation_and_textbook_last_term():
    g = load_golden("bern_nside4_k8")
    rng = np.random.default_rng(5)
    K, Fout = g["K"], g["kernel"].shape[1]
    bias = rng.standard_normal((1, 1, Fout)).astype(np.float32)
    layer = gnn_layers.Bernstein(L=g["L"], K=K, Fout=Fout, use_bias=True, activation="elu", mode="fp32")
    layer.build_from_shape(g["x"].shape)
    with torch.no_grad():
        layer.kernel.copy_(dev(g["kernel"]))
        layer.bias.copy_(dev(bias))
    ref = orc.bernstein_forward(g["x"], g["Lt"], g["kernel"], K, bias=bias, activation="elu", dtype=np.float64)
    assert rel_err(layer(dev(g["x"])).detach().cpu().numpy(), ref) <= TOL_FP32
    # stale_last_term=False: last column theta_K L~^K x; spectral reference
    tb = gnn_layers.Bernstein(L=g["L"], K=K, Fout=Fout, stale_last_term=False, mode="fp32")
    tb.build_from_shape(g["x"].shape)
    with torch.no_grad():
        tb.kernel.copy_(dev(g["kernel"]))
    lam, V = np.linalg.eigh(g["Lt"].toarray())
    P = gnn_layers.bernstein_polynomials(K, stale_last_term=False)(lam)
    Fin = g["x"].shape[2]
    W = g["kernel"].reshape(Fin, K + 1, Fout)
    spec = sum(np.einsum("mj,bjf,fo->bmo", (V * P[i]) @ V.T, g["x"], W[:, i, :]) for i in range(K + 1))
    assert rel_err(tb(dev(g["x"])).detach().cpu().numpy(), spec) <= TOL_FP32


def test_bernstein_on_the_fused_tensor_core_kernel():
    """nside 32 full sphere (the smallest with regular 16 x 16 tiles), order 4 = 5 Chebyshev terms: tf32 mode runs the
    register-resident fused kernel.  It
    must equal the Chebyshev layer fed with the basis-changed weights (same launch), and the literal Bernstein
    loops within the TF32 bar scaled by the basis change: the TF32 rounding (1e-3) acts on W' = C W and on the
    Chebyshev basis, and sum_j |C[i, j]| reaches 6 at K = 4 -> 4e-3 forward (a CPU emulation of the truncation
    gives 1.6e-3), 1e-2 for the gradients, which pass through C a second time."""
    sphere = SphereHealpix(32, k=8)
    rng = np.random.default_rng(9)
    K, Fin, Fout = 4, 8, 16
    x = rng.standard_normal((2, 12 * 32 * 32, Fin)).astype(np.float32)
    W = (rng.standard_normal(((K + 1) * Fin, Fout)) * 0.2).astype(np.float32)
    bern = gnn_layers.Bernstein(L=sphere.L, K=K, Fout=Fout, mode="tf32")
    bern.build_from_shape(x.shape)
    cheb = gnn_layers.Chebyshev(L=sphere.L, K=K + 1, Fout=Fout, mode="tf32")
    cheb.build_from_shape(x.shape)
    with torch.no_grad():
        bern.kernel.copy_(dev(W))
        cheb.kernel.copy_(bern._device_kernel())
    xb = dev(x).requires_grad_(True)
    yb = bern(xb)
    assert bern._plan.info(0)["lattice"] == 1
    assert rel_err(yb.detach().cpu().numpy(), cheb(dev(x)).detach().cpu().numpy()) <= 1e-6
    Lt, _ = orc.prepare_laplacian(sphere.L, 0.75)
    ref = orc.bernstein_forward(x, Lt, W, K, dtype=np.float64)
    assert rel_err(yb.detach().cpu().numpy(), ref) <= 4e-3
    dy = rng.standard_normal(ref.shape).astype(np.float32)
    yb.backward(dev(dy))
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wr = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    orc.torch_cpu_bernstein(xr, Lt, wr, K).backward(torch.tensor(dy, dtype=torch.float64))
    assert rel_err(xb.grad.cpu().numpy(), xr.grad.numpy()) <= 1e-2
    assert rel_err(bern.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= 1e-2


def _smoothing_layer(tmp_path, g, **kw):
    """A HealpySmoothing layer carrying exactly the golden (BallTree-built) kernel, through the layer's own
    cache-file mechanism (healpy_layers.py:650-660)."""
    sigma = float(g["sigma_arcmin"])
    label = f"-nside{int(g['nside'])}-sigma{sigma:4.2f}-n_sigma3"
    np.save(os.path.join(tmp_path, f"ind_coo{label}.npy"), g["ind_coo"])
    np.save(os.path.join(tmp_path, f"val_coo{label}.npy"), g["val_coo"])
    return hl.HealpySmoothing(int(g["nside"]), g["indices"], sigma=sigma, data_path=str(tmp_path), **kw)


def test_smoothing_matches_golden(tmp_path):
    d = np.load(os.path.join(GOLDEN, "smooth_masked16.npz"))
    g = {k: d[k] for k in d.files}
    n = int(g["M"])
    once = _smoothing_layer(tmp_path, g)
    x = dev(g["x"]).requires_grad_(True)
    before = nat.launch_count()
    y = once(x)
    assert before < nat.launch_count() <= before + 2  # all channels in ONE SpMM (ELL kernel [+ CSR-tail kernel])
    assert rel_err(y.detach().cpu().numpy(), g["y_once64"]) <= TOL_FP32
    # gradient = transposed (unsymmetric) kernel
    dy = np.random.default_rng(3).standard_normal(g["x"].shape)
    y.backward(dev(dy))
    Ks = orc.smoothing_kernel(g["ind_coo"], g["val_coo"], n).astype(np.float64)
    ref_dx = np.stack([Ks.T @ dy[b] for b in range(dy.shape[0])])
    assert rel_err(x.grad.cpu().numpy(), ref_dx) <= TOL_FP32
    # per-channel repetitions + mask
    reps = _smoothing_layer(tmp_path, g, per_channel_repetitions=[int(r) for r in g["reps"]], mask=g["mask"])
    x2 = dev(g["x"]).requires_grad_(True)
    y2 = reps(x2)
    assert rel_err(y2.detach().cpu().numpy(), g["y_reps_mask64"]) <= TOL_FP32
    y2.backward(dev(dy))
    m = g["mask"].astype(np.float64)
    KT = Ks.T.toarray()
    ref2 = np.stack([np.stack([np.linalg.matrix_power(KT, int(r)) @ (m[:, 0] * dy[b, :, c])
                               for c, r in enumerate(g["reps"])], axis=1) for b in range(dy.shape[0])])
    assert rel_err(x2.grad.cpu().numpy(), ref2) <= TOL_FP32


def test_smoothing_built_in_place_preserves_constants(tmp_path):
    """Own builder (cKDTree) end to end on the full sphere at nside 8: rows of the normalised kernel sum to ~1, so a
    constant map stays constant (to the asymmetry of the reference's normalisation), and the result matches the
    oracle applied to the layer's own kernel."""
    nside = 8
    idx = np.arange(12 * nside * nside)
    layer = hl.HealpySmoothing(nside, idx, fwhm=1200.0, data_path=str(tmp_path))
    label = layer.file_label
    ind = np.load(os.path.join(tmp_path, f"ind_coo{label}.npy"))
    val = np.load(os.path.join(tmp_path, f"val_coo{label}.npy"))
    Ks = orc.smoothing_kernel(ind, val, len(idx))
    x = np.random.default_rng(1).standard_normal((3, len(idx), 2)).astype(np.float32)
    y = layer(dev(x)).detach().cpu().numpy()
    assert rel_err(y, orc.smoothing_forward(x, Ks, dtype=np.float64)) <= TOL_FP32
    ones = layer(torch.ones(1, len(idx), 1, device="cuda")).cpu().numpy()
    assert np.abs(ones - 1).max() < 0.05


_SKINNY_CHILD = r"""
import sys
import numpy as np, torch
sys.path[:0] = [%(pkg)r, %(root)r]
from deepsphere import healpy_layers as hl, _native as nat
from oracle import deepsphere_oracle as orc
rng = np.random.default_rng(0)
worst = 0.0
for (p, Fin, Fout, act, need_dx) in [(1, 1, 16, "relu", False), (1, 2, 8, "elu", True), (2, 1, 64, None, True),
                                     (1, 3, 32, "tanh", True), (1, 4, 4, "sigmoid", False)]:
    M, B = 12 * 8 * 8, 3
    x = rng.standard_normal((B, M, Fin)).astype(np.float32)
    layer = hl.HealpyPseudoConv(p=p, Fout=Fout, activation=act)
    layer.build_from_shape(x.shape)
    with torch.no_grad():
        layer.bias.normal_()
    xt = torch.tensor(x, device="cuda", requires_grad=need_dx)
    y = layer(xt)
    w = layer.kernel.detach().cpu().numpy().astype(np.float64)
    b = layer.bias.detach().cpu().numpy().astype(np.float64)
    ref = orc.pseudo_conv(x.astype(np.float64), w, b, act)
    e = np.abs(y.detach().cpu().numpy() - ref).max() / np.abs(ref).max()
    dy = rng.standard_normal(ref.shape).astype(np.float32)
    y.backward(torch.tensor(dy, device="cuda"))
    # float64 autograd reference of the same Conv1D
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wr = torch.tensor(w, requires_grad=True); br = torch.tensor(b, requires_grad=True)
    r = 4 ** p
    z = xr.reshape(B, M // r, r * Fin) @ wr.reshape(r * Fin, Fout) + br
    yr = {None: lambda v: v, "relu": torch.relu, "elu": torch.nn.functional.elu, "tanh": torch.tanh,
          "sigmoid": torch.sigmoid}[act](z)
    yr.backward(torch.tensor(dy, dtype=torch.float64))
    rel = lambda a, c: float(np.abs(a - c).max() / max(np.abs(c).max(), 1e-300))
    e = max(e, rel(layer.kernel.grad.cpu().numpy().reshape(r * Fin, Fout), wr.grad.numpy().reshape(r * Fin, Fout)),
            rel(layer.bias.grad.cpu().numpy().ravel(), br.grad.numpy().ravel()))
    if need_dx:
        e = max(e, rel(xt.grad.cpu().numpy(), xr.grad.numpy()))
    worst = max(worst, e)
# graph convolution with bias + activation: dz and dbias come from the fused act-backward/column-sum sweep
from deepsphere import gnn_layers
from deepsphere.graph import SphereHealpix
sphere = SphereHealpix(4, k=8)
Lt, _ = orc.prepare_laplacian(sphere.L, 0.75)
for (Fin, Fout, act) in [(3, 32, "elu"), (4, 64, "tanh"), (2, 4, "sigmoid")]:
    x = rng.standard_normal((2, 192, Fin)).astype(np.float32)
    layer = gnn_layers.Chebyshev(L=sphere.L, K=3, Fout=Fout, use_bias=True, activation=act, mode="fp32")
    layer.build_from_shape(x.shape)
    with torch.no_grad():
        layer.bias.normal_()
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    y = layer(xt)
    dy = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    y.backward(torch.tensor(dy, device="cuda"))
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wr = layer.kernel.detach().cpu().double().requires_grad_(True)
    br = layer.bias.detach().cpu().double().requires_grad_(True)
    z = orc.torch_cpu_graph_conv(xr, Lt, wr, 3, "chebyshev") + br
    yr = {"elu": torch.nn.functional.elu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[act](z)
    yr.backward(torch.tensor(dy, dtype=torch.float64))
    rel = lambda a, c: float(np.abs(a - c).max() / max(np.abs(c).max(), 1e-300))
    worst = max(worst, rel(y.detach().cpu().numpy(), yr.detach().numpy()), rel(xt.grad.cpu().numpy(), xr.grad.numpy()),
                rel(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()),
                rel(layer.bias.grad.cpu().numpy().ravel(), br.grad.numpy().ravel()))
print("SKINNY_WORST", worst)
"""


def test_streaming_pseudo_conv_kernels_opt_in(tmp_path):
    """csrc/ds_skinny.cu (DEEPSPHERE_SKINNY=1, read once per process -> child process): forward, weight, bias and
    input gradients of HealpyPseudoConv on the shapes the streaming kernels serve, and a Chebyshev layer with bias +
    activation (fused act-backward / column-sum sweep), against the oracle / float64 autograd at the fp32 bar."""
    import subprocess
    import sys

    from conftest import PKG, ROOT

    script = tmp_path / "skinny_child.py"
    script.write_text(_SKINNY_CHILD % {"pkg": PKG, "root": ROOT})
    env = dict(os.environ, DEEPSPHERE_SKINNY="1")
    res = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    worst = float([ln for ln in res.stdout.splitlines() if ln.startswith("SKINNY_WORST")][-1].split()[1])
    assert worst <= TOL_FP32
