"""HealpyGCNN end to end on the GPU: forward, a training step, linearity / adjoint properties
at a larger size, save/load.  Run with -m gpu."""
import numpy as np
import pytest
import torch

import deepsphere
from deepsphere import gnn_layers, healpy_layers as hl, keras_compat
from deepsphere.graph import SphereHealpix
from helpers import orc, rel_err

pytestmark = pytest.mark.gpu


def _reference_forward(model, x):
    """The same network restated on the CPU in float64 with the oracle, layer by layer."""
    h = np.asarray(x, dtype=np.float64)
    for layer in model.layers:
        if isinstance(layer, (gnn_layers.Chebyshev, gnn_layers.Monomial)):
            rec = "chebyshev" if isinstance(layer, gnn_layers.Chebyshev) else "monomial"
            Lt, _ = orc.prepare_laplacian(layer.L, 0.75 if rec == "chebyshev" else 1.0)
            bias = layer.bias.detach().double().detach().cpu().numpy() if layer.use_bias else None
            bn = None
            if layer.use_bn:
                bn = (layer.bn.moving_mean.double().detach().cpu().numpy(), layer.bn.moving_variance.double().detach().cpu().numpy())
            act = [k for k, v in keras_compat.ACTIVATIONS.items() if v[1] is layer.activation]
            h = orc.graph_conv_forward(h, Lt, layer.kernel.detach().double().detach().cpu().numpy(), layer.K, rec, bias=bias,
                                       activation=act[0] if act else None, use_bn=layer.use_bn, training=False,
                                       bn_state=bn, dtype=np.float64)
        elif isinstance(layer, hl.HealpyPool):
            h = orc.healpy_pool(h, layer.p, layer.pool_type)
        elif isinstance(layer, hl.HealpyPseudoConv):
            h = orc.pseudo_conv(h, layer.kernel.detach().double().detach().cpu().numpy(), layer.bias.detach().double().detach().cpu().numpy())
        elif isinstance(layer, hl.HealpyPseudoConv_Transpose):
            h = orc.pseudo_conv_transpose(h, layer.kernel.detach().double().detach().cpu().numpy(),
                                          layer.bias.detach().double().detach().cpu().numpy())
        else:
            h = layer(torch.tensor(h)).numpy()
    return h


def test_healpy_gcnn_forward_matches_oracle_pipeline():
    nside = 16
    layers = [hl.HealpyPseudoConv(p=1, Fout=4), hl.HealpyChebyshev(K=5, Fout=8, use_bias=True, use_bn=True,
                                                                   activation="relu"),
              hl.HealpyPool(p=1), hl.HealpyMonomial(K=4, Fout=8, activation="elu"), hl.HealpyPool(p=1, pool_type="AVG"),
              hl.HealpyChebyshev(K=3, Fout=16), hl.HealpyPseudoConv_Transpose(p=1, Fout=3),
              keras_compat.Lambda(lambda t: t.mean(dim=1))]
    torch.manual_seed(0)
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(12 * nside**2), layers=layers, n_neighbors=20)
    x = np.random.default_rng(0).standard_normal((3, 12 * nside**2, 2)).astype(np.float32)
    y = model(x, training=False).detach().cpu().numpy()
    assert y.shape == (3, 3)
    assert rel_err(y, _reference_forward(model, x)) <= 2e-5


def test_training_step_decreases_loss_and_matches_cpu_gradients():
    nside = 8
    npix = 12 * nside**2
    layers = [hl.HealpyChebyshev(K=4, Fout=6, use_bias=True, activation="relu"), hl.HealpyPool(p=1),
              hl.HealpyChebyshev(K=3, Fout=2), keras_compat.Lambda(lambda t: t.mean(dim=1))]
    torch.manual_seed(1)
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(npix), layers=layers)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((8, npix, 1)).astype(np.float32)
    t = rng.standard_normal((8, 2)).astype(np.float32)
    model.build(input_shape=(None, npix, 1))
    opt = torch.optim.Adam(model.trainable_variables, lr=1e-2)
    losses = []
    for _ in range(15):
        opt.zero_grad()
        loss = ((model(x, training=True) - torch.tensor(t).cuda()) ** 2).mean()
        loss.backward()
        if not losses:
            # gradient of the first conv kernel vs a float64 torch-CPU restatement of the network
            l0, l2 = model.layers[0], model.layers[2]
            w0 = l0.kernel.detach().double().cpu().requires_grad_(True)
            b0 = l0.bias.detach().double().cpu()
            w2 = l2.kernel.detach().double().cpu()
            Lt0, _ = orc.prepare_laplacian(l0.L, 0.75)
            Lt2, _ = orc.prepare_laplacian(l2.L, 0.75)
            h = torch.relu(orc.torch_cpu_graph_conv(torch.tensor(x, dtype=torch.float64), Lt0, w0, 4) + b0)
            h = h.reshape(8, npix // 4, 4, 6).max(dim=2).values
            out = orc.torch_cpu_graph_conv(h, Lt2, w2, 3).mean(dim=1)
            ((out - torch.tensor(t, dtype=torch.float64)) ** 2).mean().backward()
            assert rel_err(l0.kernel.grad.detach().cpu().numpy(), w0.grad.numpy()) <= 5e-5
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("nside,B,F", [(64, 4, 16), (128, 2, 64)])
def test_linearity_and_adjoint_identity_at_scale(nside, B, F):
    """Properties that hold at any size: the layer is linear in x (no bias/activation), and
    <conv(x), dy> == <x, dx> == <kernel, dkernel> for its gradients."""
    g = SphereHealpix(nside, k=8)
    layer = gnn_layers.Chebyshev(L=g.L, K=5, Fout=F)
    gen = torch.Generator(device="cuda").manual_seed(0)
    M = g.L.shape[0]
    x1 = torch.randn(B, M, F, device="cuda", generator=gen)
    x2 = torch.randn(B, M, F, device="cuda", generator=gen)
    y1, y2 = layer(x1), layer(x2)
    y12 = layer(2.0 * x1 - 0.5 * x2)
    scale = float(y12.abs().max())
    assert float((y12 - (2.0 * y1 - 0.5 * y2)).abs().max()) <= 2e-5 * scale
    x = x1.clone().requires_grad_(True)
    y = layer(x)
    dy = torch.randn(y.shape, device="cuda", generator=gen)
    y.backward(dy)
    lhs = float((y.detach().double() * dy.double()).sum())
    rhs_x = float((x.grad.double() * x.detach().double()).sum())
    rhs_w = float((layer.kernel.grad.double() * layer.kernel.detach().double()).sum())
    norm = float(y.detach().double().norm() * dy.double().norm())
    assert abs(lhs - rhs_x) <= 1e-5 * norm and abs(lhs - rhs_w) <= 1e-5 * norm
