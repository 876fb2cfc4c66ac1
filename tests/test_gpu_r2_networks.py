"""Round 2: the networks of the reference's example notebooks (deepsphere/example_networks.py) end to end on the GPU
against the oracle's torch-CPU float64 restatement (oracle.torch_cpu_network via oracle/bridge.py): output and EVERY weight
gradient, training mode (batch statistics in the BatchNormalization layers)."""
import numpy as np
import pytest
import torch

import deepsphere
from deepsphere import example_networks as nets, healpix as hpx, utils
from helpers import orc, rel_err
from oracle import bridge

pytestmark = pytest.mark.gpu


def _check(model, x, tol_y, tol_g):
    xt = torch.tensor(x, dtype=torch.float32, device="cuda")
    y = model(xt, training=True)
    c = np.random.default_rng(1).standard_normal(tuple(y.shape))
    (y * torch.tensor(c, dtype=torch.float32, device="cuda")).sum().backward()
    specs, pairs = bridge.specs_from_layers(model.layers)
    yr = orc.torch_cpu_network(torch.tensor(x, dtype=torch.float64), specs, training=True)
    (yr * torch.tensor(c)).sum().backward()
    assert tuple(yr.shape) == tuple(y.shape)
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= tol_y
    assert len(pairs) == len(list(model.parameters()))
    worst = 0.0
    for p, r in pairs:
        assert p.grad is not None and r.grad is not None
        worst = max(worst, rel_err(p.grad.cpu().numpy(), r.grad.numpy()))
    assert worst <= tol_g, worst


def test_quick_start_network_matches_oracle():
    """examples/quick_start.ipynb:118-127,197: 4 x HealpyChebyshev K = 10 (bias + BatchNorm + relu) with MAX pools,
    n_neighbors = 20, mean + softmax head."""
    nside = 16
    torch.manual_seed(0)
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(12 * nside**2), layers=nets.quick_start_layers(),
                                  n_neighbors=20)
    x = np.random.default_rng(0).standard_normal((4, 12 * nside**2, 1))
    _check(model, x, 2e-5, 2e-3)


def test_advanced_tutorial_network_matches_oracle():
    """examples/advanced_tutorial.ipynb:137,211,309-325 at nside 32: masked sky padded with extend_indices, Chebyshev +
    Monomial K = 10 with BatchNorm, AVG / MAX pools, a residual layer whose sub-layers carry their own BatchNorm, a
    pseudo-convolution, n_neighbors = 20."""
    nside = 32
    idx = utils.extend_indices(hpx.query_disc(nside, [1, 0, 0], 1.5), nside, 4)
    torch.manual_seed(1)
    model = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=nets.advanced_tutorial_layers(), n_neighbors=20)
    x = np.random.default_rng(2).standard_normal((3, len(idx), 1))
    _check(model, x, 5e-5, 5e-3)


def test_autoencoder_networks_match_oracle():
    """examples/generative_models.ipynb:185-213: pseudo-convolutions, Chebyshev K = 5 F = 16, LayerNormalization(axis=1),
    elu, transposed pseudo-convolutions back to the input resolution (encoder and decoder as two HealpyGCNN)."""
    nside = 16
    enc_layers, dec_layers = nets.autoencoder_layers()
    torch.manual_seed(2)
    enc = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(12 * nside**2), layers=enc_layers, n_neighbors=20)
    x = np.random.default_rng(3).standard_normal((2, 12 * nside**2, 1))
    _check(enc, x, 2e-5, 5e-4)
    nb = nside // 8
    dec = deepsphere.HealpyGCNN(nside=nb, indices=np.arange(12 * nb**2), layers=dec_layers, n_neighbors=20)
    z = np.random.default_rng(4).standard_normal((2, 12 * nb**2, 16))
    _check(dec, z, 2e-5, 5e-4)
