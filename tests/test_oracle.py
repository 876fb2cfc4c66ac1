"""The oracle against (a) its committed golden vectors, (b) independent formulations, (c) the
reference's own known-answer tests.  CPU only."""
import numpy as np
import pytest
import torch
from scipy import sparse

from helpers import CONV_CASES, load_golden, orc, rel_err


@pytest.mark.parametrize("name", CONV_CASES)
def test_oracle_reproduces_golden(name):
    g = load_golden(name)
    Lt, lmax = orc.prepare_laplacian(g["L"], 0.75 if g["recursion"] == "chebyshev" else 1.0)
    assert abs(lmax - float(g["lmax"])) <= 1e-9 * abs(lmax)
    assert abs(Lt - g["Lt"]).max() < 1e-12
    y = orc.graph_conv_forward(g["x"], g["Lt"], g["kernel"], g["K"], g["recursion"], bias=g["bias"],
                               activation=g["activation"], dtype=np.float64)
    assert rel_err(y, g["y64"]) < 1e-13
    y32 = orc.graph_conv_forward(g["x"].astype(np.float32), g["Lt"], g["kernel"].astype(np.float32), g["K"],
                                 g["recursion"], bias=None if g["bias"] is None else g["bias"].astype(np.float32),
                                 activation=g["activation"], dtype=np.float32)
    assert rel_err(y32, g["y64"]) < 1e-5  # the fp32 restatement itself sits inside the parity tolerance
    dx, dk, db = orc.graph_conv_backward(g["x"], g["Lt"], g["kernel"], g["K"], g["dy"], g["recursion"])
    assert rel_err(dx, g["dx64"]) < 1e-13 and rel_err(dk, g["dkernel64"]) < 1e-13


@pytest.mark.parametrize("name", ["cheb_nside4_k8", "cheb_ref3x3", "cheb_masked16_k8"])
def test_chebyshev_oracle_vs_spectral_definition(name):
    """Independent formulation: T_k(L~) = V diag(cos(k arccos(lambda))) V^T for symmetric L~."""
    g = load_golden(name)
    Ld = g["Lt"].toarray()
    lam, V = np.linalg.eigh(Ld)
    assert lam.min() >= -1 - 1e-9 and lam.max() <= 1 + 1e-9  # rescaled spectrum
    x, K = g["x"], g["K"]
    B, M, Fin = x.shape
    kernel = g["kernel"].reshape(Fin, K, -1)  # rows f*K + k
    z = np.zeros((B, M, kernel.shape[-1]))
    for k in range(K):
        Tk = (V * np.cos(k * np.arccos(np.clip(lam, -1, 1)))) @ V.T
        z += np.einsum("mj,bjf,fo->bmo", Tk, x, kernel[:, k, :])
    y_lin = orc.graph_conv_forward(x, g["Lt"], g["kernel"], K, "chebyshev", dtype=np.float64)
    assert rel_err(y_lin, z) < 1e-10


@pytest.mark.parametrize("name", ["cheb_nside4_k8", "mono_nside4_k8", "cheb_masked16_k20"])
def test_oracle_gradients_vs_torch_autograd(name):
    g = load_golden(name)
    x = torch.tensor(g["x"], dtype=torch.float64, requires_grad=True)
    w = torch.tensor(g["kernel"], dtype=torch.float64, requires_grad=True)
    y = orc.torch_cpu_graph_conv(x, g["Lt"], w, g["K"], g["recursion"])
    y_lin = orc.graph_conv_forward(g["x"], g["Lt"], g["kernel"], g["K"], g["recursion"], dtype=np.float64)
    assert rel_err(y.detach().numpy(), y_lin) < 1e-12
    y.backward(torch.tensor(g["dy"]))
    assert rel_err(x.grad.numpy(), g["dx64"]) < 1e-12
    assert rel_err(w.grad.numpy(), g["dkernel64"]) < 1e-12
    assert rel_err(g["dy"].sum(axis=(0, 1)), g["dbias64"].ravel()) < 1e-12


def test_pool_reference_known_answer():
    """reference tests/test_healpy_layers.py:9-37: np.random.seed(11), nside 4 -> 2; AVG equals
    hp.ud_grade (mean of the four nested children), MAX equals the reshape-max; tol 1e-5."""
    g = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "pool_nside4.npz"))
    np.random.seed(11)
    m_in = np.random.normal(size=12 * 4 * 4)
    assert np.array_equal(m_in, g["m_in"])
    avg = orc.healpy_pool(m_in[None, :, None], 1, "AVG").ravel()
    mx = orc.healpy_pool(m_in[None, :, None], 1, "MAX").ravel()
    assert np.all(np.abs(avg - g["avg"]) < 1e-5) and np.all(np.abs(mx - g["max"]) < 1e-5)
    with pytest.raises(IOError):
        orc.healpy_pool(m_in[None, :, None], 0, "MAX")
    with pytest.raises(IOError):
        orc.healpy_pool(m_in[None, :, None], 2, "HUHU")


def test_pseudo_conv_oracle_vs_torch_conv():
    """Keras Conv1D(k = s = 4^p) and Conv2DTranspose((1,4^p)) restated with torch's own conv ops."""
    g = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "pconv_nside8.npz"))
    x, w, b, wt = (torch.tensor(g[k]) for k in ("x", "w", "b", "wt"))
    # Conv1D: torch weight [Fout, Fin, r] = keras [r, Fin, Fout] permuted
    y = torch.nn.functional.conv1d(x.permute(0, 2, 1), w.permute(2, 1, 0), b, stride=16).permute(0, 2, 1)
    assert rel_err(torch.nn.functional.elu(y).numpy(), g["y"]) < 1e-12
    # Conv2DTranspose: torch weight [Fin, Fout, 1, r] = keras [1, r, Fout, Fin] permuted
    xt = x[:, :48].permute(0, 2, 1)[:, :, None, :]
    yt = torch.nn.functional.conv_transpose2d(xt, wt.permute(3, 2, 0, 1), b, stride=(1, 16))[:, :, 0].permute(0, 2, 1)
    assert rel_err(torch.relu(yt).numpy(), g["yt"]) < 1e-12


def test_batch_norm_oracle_vs_torch():
    rng = np.random.default_rng(3)
    z = rng.standard_normal((4, 50, 6))
    out, mm, mv = orc.batch_norm(z, training=True)
    ref = torch.nn.functional.batch_norm(torch.tensor(z).permute(0, 2, 1), None, None, training=True, eps=1e-5)
    assert rel_err(out, ref.permute(0, 2, 1).numpy()) < 1e-12
    assert np.allclose(mm, 0.1 * z.mean(axis=(0, 1))) and np.allclose(mv, 0.9 + 0.1 * z.var(axis=(0, 1)))


def test_rescale_does_not_mutate_input():
    L = sparse.csr_matrix(np.diag([1.0, 2.0, 3.0]))
    before = L.copy()
    orc.prepare_laplacian(L, 0.75)
    assert abs(L - before).max() == 0
