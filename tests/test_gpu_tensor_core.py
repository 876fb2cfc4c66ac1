"""tcgen05 contraction modes against the fp64 oracle: DS_MODE_TF32 within the stated 1e-3,
DS_MODE_TF32X3 (error-compensated split) within the fp32 bar of 1e-5 x small factor."""
import numpy as np
import pytest
import torch

from deepsphere import gnn_layers
from deepsphere.graph import SphereHealpix
from helpers import orc, rel_err

pytestmark = pytest.mark.gpu
TOL = {"tf32": 1e-3, "tf32x3": 2e-5}


@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
@pytest.mark.parametrize("nside,B,Fin,Fout,K", [(4, 3, 64, 64, 5), (8, 2, 32, 16, 3), (4, 5, 16, 64, 4),
                                                (8, 1, 8, 32, 1), (16, 2, 64, 128, 2)])
def test_tensor_core_forward_backward(mode, nside, B, Fin, Fout, K):
    g = SphereHealpix(nside, k=8)
    M = g.L.shape[0]
    torch.manual_seed(0)
    layer = gnn_layers.Chebyshev(L=g.L, K=K, Fout=Fout, use_bias=True, activation="elu", mode=mode)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, M, Fin))
    dy = rng.standard_normal((B, M, Fout))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    y = layer(xt)
    y.backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    xr = torch.tensor(x, requires_grad=True)
    wr = layer.kernel.detach().double().cpu().requires_grad_(True)
    br = layer.bias.detach().double().cpu().requires_grad_(True)
    yr = torch.nn.functional.elu(orc.torch_cpu_graph_conv(xr, Lt, wr, K) + br)
    yr.backward(torch.tensor(dy))
    tol = TOL[mode]
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= tol
    assert rel_err(xt.grad.cpu().numpy(), xr.grad.numpy()) <= tol
    assert rel_err(layer.kernel.grad.cpu().numpy(), wr.grad.numpy()) <= tol
    assert rel_err(layer.bias.grad.cpu().numpy(), br.grad.numpy()) <= tol


def test_tensor_core_mode_rejects_unsupported_shapes_loudly():
    from deepsphere import _native as nat

    layer = gnn_layers.Chebyshev(L=np.eye(48), K=2, Fout=5, mode="tf32")
    with pytest.raises(nat.NativeError, match="tensor-core mode"):
        layer(np.zeros((1, 48, 3), np.float32))


@pytest.mark.parametrize("Fin,Fout,K", [(16, 32, 5), (32, 32, 5), (64, 64, 5), (16, 16, 2), (48, 64, 3), (64, 16, 5),
                                        (32, 16, 1)])
def test_weight_gradient_kernel_ragged_rows_and_block_counts(Fin, Fout, K):
    """umma_gemm_tn_kernel's cp.async producers are instantiated per number of 32-channel operand blocks (2 .. 12 here)
    and zero-fill the rows past the end of a ragged problem: a partial sky whose row count is not a multiple of the
    16-row stage, enough rows for several stages per CTA, weight and bias gradients against the oracle."""
    from deepsphere import healpix as hpx

    idx = hpx.query_disc(64, [0.2, -0.4, 0.9], 0.6)
    idx = idx[: len(idx) - (len(idx) % 16) - 5]  # M % 16 == 11
    g = SphereHealpix(64, indexes=idx, k=8)
    M, B = g.L.shape[0], 7
    assert M % 16 != 0 and (B * M) % 16 != 0 and B * M // 16 > 4 * 148
    torch.manual_seed(0)
    layer = gnn_layers.Chebyshev(L=g.L, K=K, Fout=Fout, use_bias=True, mode="tf32")
    rng = np.random.default_rng(Fin + K)
    x = rng.standard_normal((B, M, Fin))
    dy = rng.standard_normal((B, M, Fout))
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    layer(xt).backward(torch.tensor(dy, dtype=torch.float32, device="cuda"))
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    w = layer.kernel.detach().double().cpu().numpy()
    rdx, rdk, _ = orc.graph_conv_backward(x, Lt, w, K, dy, "chebyshev")
    assert rel_err(layer.kernel.grad.cpu().numpy(), rdk) <= 1e-3
    assert rel_err(xt.grad.cpu().numpy(), rdx) <= 1e-3
    assert rel_err(layer.bias.grad.cpu().numpy(), dy.sum(axis=(0, 1))) <= 1e-5
