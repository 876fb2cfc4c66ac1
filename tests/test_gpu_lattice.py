"""Fused lattice kernel (ds_lattice.cu) against the oracle and against the generic per-hop kernels."""
import numpy as np
import pytest
import torch

from deepsphere import _native as nat
from deepsphere import gnn_layers, healpix as hpx, utils
from deepsphere.graph import SphereHealpix
from helpers import orc, rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _run(layer, x, dy):
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    y = layer(xt)
    y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    return y.detach().cpu().numpy(), xt.grad.cpu().numpy(), layer.kernel.grad.cpu().numpy()


@pytest.mark.parametrize("cls", ["Chebyshev", "Monomial"])
@pytest.mark.parametrize("nside,B,Fin,Fout,K", [(32, 3, 4, 4, 5), (32, 2, 16, 8, 3), (64, 2, 64, 64, 5), (32, 5, 8, 4, 2),
                                                (32, 2, 12, 4, 8), (32, 2, 8, 8, 10), (32, 1, 4, 4, 13)])
def test_lattice_forward_backward_matches_oracle(cls, nside, B, Fin, Fout, K):
    g = SphereHealpix(nside, k=8)
    M = g.L.shape[0]
    torch.manual_seed(0)
    layer = getattr(gnn_layers, cls)(L=g.L, K=K, Fout=Fout)
    rng = np.random.default_rng(K)
    x = rng.standard_normal((B, M, Fin))
    dy = rng.standard_normal((B, M, Fout))
    y, dx, dk = _run(layer, x, dy)
    assert layer._plan.info(0)["lattice"] == 1, "the fused lattice path was expected to be active"
    rec = cls.lower()
    Lt, _ = orc.prepare_laplacian(g.L, 0.75 if rec == "chebyshev" else 1.0)
    w = layer.kernel.detach().double().cpu().numpy()
    assert rel_err(y, orc.graph_conv_forward(x, Lt, w, K, rec, dtype=np.float64)) <= 1e-5
    rdx, rdk, _ = orc.graph_conv_backward(x, Lt, w, K, dy, rec)
    assert rel_err(dx, rdx) <= 1e-5 and rel_err(dk, rdk) <= 1e-5


def test_lattice_equals_generic_path(monkeypatch):
    """Same layer with the lattice attachment disabled: results agree to fp32 round-off."""
    g = SphereHealpix(64, k=8)
    M = g.L.shape[0]
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, M, 16))
    dy = rng.standard_normal((2, M, 16))
    torch.manual_seed(0)
    fused = gnn_layers.Chebyshev(L=g.L, K=5)
    a = _run(fused, x, dy)
    monkeypatch.setenv("DEEPSPHERE_LATTICE", "0")
    plain = gnn_layers.Chebyshev(L=g.L, K=5)
    plain.build_from_shape(x.shape)
    with torch.no_grad():
        plain.kernel.copy_(fused.kernel)
    b = _run(plain, x, dy)
    assert plain._plan.info(0)["lattice"] == 0 and fused._plan.info(0)["lattice"] == 1
    for u, v in zip(a, b):
        assert rel_err(u, v) <= 2e-6


def test_lattice_masked_sky():
    ext = utils.extend_indices(hpx.query_disc(64, [1, 0, 0], 1.2), 64, 8)
    g = SphereHealpix(64, indexes=ext, k=8)
    M = len(ext)
    layer = gnn_layers.Chebyshev(L=g.L, K=4, Fout=8, healpix=(64, ext))
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, M, 8))
    dy = rng.standard_normal((3, M, 8))
    y, dx, dk = _run(layer, x, dy)
    assert layer._plan.info(0)["lattice"] == 1
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    w = layer.kernel.detach().double().cpu().numpy()
    assert rel_err(y, orc.graph_conv_forward(x, Lt, w, 4, "chebyshev", dtype=np.float64)) <= 1e-5
    rdx, rdk, _ = orc.graph_conv_backward(x, Lt, w, 4, dy, "chebyshev")
    assert rel_err(dx, rdx) <= 1e-5 and rel_err(dk, rdk) <= 1e-5


@pytest.mark.parametrize("mode,tol", [("tf32", 1e-3), ("tf32x3", 2e-5)])
@pytest.mark.parametrize("cls", ["Chebyshev", "Monomial"])
@pytest.mark.parametrize("nside,B,Fin,Fout,K", [(32, 2, 16, 16, 5), (64, 2, 64, 64, 5), (32, 3, 32, 16, 3),
                                                (32, 2, 16, 64, 2), (32, 1, 48, 32, 4), (32, 2, 24, 80, 5)])
def test_fused_lattice_conv_matches_oracle(mode, tol, cls, nside, B, Fin, Fout, K):
    """Fused path (tf32: ds_lattice_conv2.cu, recursion + tcgen05 contraction in one kernel; tf32x3: lattice recursion +
    tensor-core GEMM kernels) - forward, and the same kernels on dz
    plus the transposed weight-gradient contraction (backward), against the float64 oracle."""
    g = SphereHealpix(nside, k=8)
    M = g.L.shape[0]
    torch.manual_seed(0)
    layer = getattr(gnn_layers, cls)(L=g.L, K=K, Fout=Fout, use_bias=True, activation="elu", mode=mode)
    rng = np.random.default_rng(K + Fin)
    x = rng.standard_normal((B, M, Fin))
    dy = rng.standard_normal((B, M, Fout))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    before = nat.launch_count()
    y = layer(xt)
    fwd_launches = nat.launch_count() - before
    y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    assert layer._plan.info(0)["lattice"] == 1
    # fused forward = weight-image prep + fused kernel + the irregular-tile sub-problem; no per-hop launches
    assert fwd_launches <= 4 + 2 * K, fwd_launches
    if mode == "tf32":  # weight image + fused kernel + ONE launch for the irregular rows (ds_patch.cu)
        assert fwd_launches == 3, fwd_launches
    rec = cls.lower()
    Lt, _ = orc.prepare_laplacian(g.L, 0.75 if rec == "chebyshev" else 1.0)
    xr = torch.tensor(x, requires_grad=True)
    wr = layer.kernel.detach().double().cpu().requires_grad_(True)
    br = layer.bias.detach().double().cpu().requires_grad_(True)
    yr = torch.nn.functional.elu(orc.torch_cpu_graph_conv(xr, Lt, wr, K, rec) + br)
    yr.backward(torch.tensor(dy))
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= tol
    assert rel_err(xt.grad.cpu().numpy(), xr.grad.numpy()) <= tol
    assert rel_err(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= tol
    assert rel_err(layer.bias.grad.cpu().numpy(), br.grad.numpy()) <= tol
    if mode == "tf32":  # the rows around the valence-3 vertices come from the fp32 patch kernel (forward: exact inputs)
        pay = layer._plan._lattice_payload
        irr = pay["closure_rows"][pay["own_sub"]]
        assert len(irr) == 360 and pay["patches"]["n_patches"] == 8
        assert rel_err(y.detach().cpu().numpy()[:, irr], yr.detach().numpy()[:, irr]) <= 2e-5
    # the scale-aware metric as well (errors relative to the tensor's norm, not to its largest element)
    assert rel_l2(y.detach().cpu().numpy(), yr.detach().numpy()) <= tol
    assert rel_l2(xt.grad.cpu().numpy(), xr.grad.numpy()) <= tol
    assert rel_l2(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= tol


def test_fused_conv2_masked_sky_tf32():
    """ds_lattice_conv2.cu on a partial sky: tiles with holes (zero-filled lattice positions, zero weights), padded
    index set, K = 4 on the 24 x 24 lattice plan (3 of the 4 halo rings used), bias + elu epilogue (a smooth
    activation: with relu the TF32 rounding of y flips the gradient mask of the elements next to zero)."""
    ext = utils.extend_indices(hpx.query_disc(64, [1, 0, 0], 1.2), 64, 8)
    g = SphereHealpix(64, indexes=ext, k=8)
    M = len(ext)
    torch.manual_seed(1)
    layer = gnn_layers.Chebyshev(L=g.L, K=4, Fout=32, healpix=(64, ext), use_bias=True, activation="elu", mode="tf32")
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, M, 16))
    dy = rng.standard_normal((3, M, 32))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    y = layer(xt)
    y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    assert layer._plan.info(0)["lattice"] == 1
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    xr = torch.tensor(x, requires_grad=True)
    wr = layer.kernel.detach().double().cpu().requires_grad_(True)
    br = layer.bias.detach().double().cpu().requires_grad_(True)
    yr = torch.nn.functional.elu(orc.torch_cpu_graph_conv(xr, Lt, wr, 4, "chebyshev") + br)
    yr.backward(torch.tensor(dy))
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= 1e-3
    assert rel_err(xt.grad.cpu().numpy(), xr.grad.numpy()) <= 1e-3
    assert rel_err(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= 1e-3


def test_fused_conv2_linearity_and_batch_independence():
    """Size-independent properties of the fused tf32 kernel at a larger shape: the layer is linear in x, and a
    sample's output does not depend on what else is in the batch (work items are (tile, sample, chunk))."""
    g = SphereHealpix(64, k=8)
    M = g.L.shape[0]
    torch.manual_seed(2)
    layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=32, mode="tf32")
    gen = torch.Generator(device="cuda").manual_seed(3)
    x1 = torch.randn(4, M, 32, device="cuda", generator=gen)
    x2 = torch.randn(4, M, 32, device="cuda", generator=gen)
    with torch.no_grad():
        y1, y2, y12 = layer(x1), layer(x2), layer(x1 + 2.0 * x2)
        scale = float(y12.abs().max())
        assert float((y12 - (y1 + 2.0 * y2)).abs().max()) <= 2e-3 * scale  # TF32 operand truncation is not linear
        ya = layer(x1[1:2])
        assert float((ya[0] - y1[1]).abs().max()) <= 1e-6 * scale  # same items, same arithmetic


def test_fused_forward_without_input_gradient():
    """First layer of a network: x does not require a gradient, so the backward takes the non-fused weight-gradient
    path and must recompute the basis the fused forward never wrote (ds_graph_conv_forward_writes_basis == 0)."""
    g = SphereHealpix(32, k=8)
    M = g.L.shape[0]
    torch.manual_seed(4)
    layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=32, use_bias=True, mode="tf32")
    rng = np.random.default_rng(9)
    x = rng.standard_normal((2, M, 16))
    dy = rng.standard_normal((2, M, 32))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda")  # requires_grad = False
    y = layer(xt)
    y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    h = layer._plan.handle(0)
    assert nat.lib().ds_graph_conv_forward_writes_basis(h, 5, 2, 16, 32, nat.MODES["tf32"]) == 0
    assert nat.lib().ds_graph_conv_forward_writes_basis(h, 5, 2, 16, 32, nat.MODES["fp32"]) == 1
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    wr = layer.kernel.detach().double().cpu().requires_grad_(True)
    br = layer.bias.detach().double().cpu().requires_grad_(True)
    yr = orc.torch_cpu_graph_conv(torch.tensor(x), Lt, wr, 5, "chebyshev") + br
    yr.backward(torch.tensor(dy))
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= 1e-3
    assert rel_err(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= 1e-3
    assert rel_err(layer.bias.grad.cpu().numpy(), br.grad.numpy()) <= 1e-3
