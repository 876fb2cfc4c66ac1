"""Parity of the CUDA path (through the C-ABI) against the oracle and the golden vectors.
Bars: bit-exact for pooling / indexing, rel <= 1e-5 for the fp32 path (max-abs error over
max-abs reference), <= 1e-3 for TF32 contraction modes.  Run with -m gpu on a B200."""
import numpy as np
import pytest
import torch
from scipy import sparse

from deepsphere import _native as nat
from deepsphere import _ops, gnn_layers, healpy_layers, keras_compat, utils
from deepsphere.graph import SphereHealpix
from helpers import CONV_CASES, GOLDEN, load_golden, orc, rel_err

pytestmark = pytest.mark.gpu
TOL_FP32 = 1e-5


def dev(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


def make_layer(g, cls=None, **kw):
    """A reference-API layer carrying the golden case's Laplacian and weights."""
    cls = cls or (gnn_layers.Chebyshev if g["recursion"] == "chebyshev" else gnn_layers.Monomial)
    layer = cls(L=g["L"], K=g["K"], Fout=g["kernel"].shape[1], use_bias=g["bias"] is not None,
                activation=g["activation"], **kw)
    layer.build_from_shape(g["x"].shape)
    with torch.no_grad():
        layer.kernel.copy_(dev(g["kernel"]))
        if g["bias"] is not None:
            layer.bias.copy_(dev(g["bias"]))
    return layer


def test_native_library_is_the_one_running():
    assert nat.lib().ds_device_count() >= 1
    before = nat.launch_count()
    healpy_layers.HealpyPool(1)(np.zeros((1, 48, 1), np.float32))
    assert nat.launch_count() > before


@pytest.mark.parametrize("name", CONV_CASES)
def test_graph_conv_forward_matches_golden(name):
    g = load_golden(name)
    layer = make_layer(g)
    # the layer's own Laplacian prep must land on the golden L~ (fp64 -> fp32 rounding point, gnn_layers.py:71)
    dense = sparse.csr_matrix((layer._L_values, (layer._L_indices[:, 0], layer._L_indices[:, 1])), shape=g["Lt"].shape)
    assert abs(dense - g["Lt"]).max() < 1e-6
    y = layer(g["x"].astype(np.float32)).detach().cpu().numpy()
    assert y.shape == g["y64"].shape
    assert rel_err(y, g["y64"]) <= TOL_FP32
    assert rel_err(y, g["y32"]) <= TOL_FP32


@pytest.mark.parametrize("name", CONV_CASES)
@pytest.mark.parametrize("save_basis", [True, False])
def test_graph_conv_backward_matches_golden(name, save_basis, monkeypatch):
    g = load_golden(name)
    if not save_basis:
        monkeypatch.setattr(_ops, "_SAVE_BASIS_MAX_BYTES", 0)  # force the recompute path
    layer = gnn_layers.Chebyshev if g["recursion"] == "chebyshev" else gnn_layers.Monomial
    layer = layer(L=g["L"], K=g["K"], Fout=g["kernel"].shape[1], use_bias=True)
    layer.build_from_shape(g["x"].shape)
    with torch.no_grad():
        layer.kernel.copy_(dev(g["kernel"]))
    x = dev(g["x"]).requires_grad_(True)
    y = layer(x)
    y.backward(dev(g["dy"]))
    assert rel_err(x.grad.detach().cpu().numpy(), g["dx64"]) <= TOL_FP32
    assert rel_err(layer.kernel.grad.detach().cpu().numpy(), g["dkernel64"]) <= TOL_FP32
    assert rel_err(layer.bias.grad.detach().cpu().numpy(), g["dbias64"]) <= TOL_FP32


@pytest.mark.parametrize("act", ["relu", "elu", "sigmoid", "tanh", "softplus", "gelu"])
@pytest.mark.parametrize("use_bn", [False, True])
def test_bias_bn_activation_forward_backward(act, use_bn):
    """BN sits before the bias (gnn_layers.py:152-159); checked against a torch-CPU float64
    autograd restatement of the whole layer."""
    g = load_golden("cheb_nside4_k8")
    torch.manual_seed(3)
    layer = gnn_layers.Chebyshev(L=g["L"], K=g["K"], Fout=6, use_bias=True, use_bn=use_bn, activation=act)
    layer.build_from_shape(g["x"].shape)
    with torch.no_grad():
        layer.kernel.copy_(dev(g["kernel"]))
    x = dev(g["x"]).requires_grad_(True)
    y = layer(x, training=True)
    dy = dev(g["dy"])
    y.backward(dy)
    # float64 CPU restatement
    xr = torch.tensor(g["x"], requires_grad=True)
    wr = torch.tensor(g["kernel"], requires_grad=True)
    br = layer.bias.detach().double().cpu().requires_grad_(True)
    z = orc.torch_cpu_graph_conv(xr, g["Lt"], wr, g["K"], "chebyshev")
    if use_bn:
        z = (z - z.mean(dim=(0, 1))) / torch.sqrt(z.var(dim=(0, 1), unbiased=False) + 1e-5)
    fn = keras_compat.ACTIVATIONS[act][1]
    yr = fn(z + br)
    yr.backward(torch.tensor(g["dy"]))
    tol = 2e-5 if use_bn else TOL_FP32
    assert rel_err(y.detach().cpu().numpy(), yr.detach().numpy()) <= tol
    assert rel_err(x.grad.detach().cpu().numpy(), xr.grad.numpy()) <= 5 * tol
    assert rel_err(layer.kernel.grad.detach().cpu().numpy(), wr.grad.numpy()) <= 5 * tol
    assert rel_err(layer.bias.grad.detach().cpu().numpy(), br.grad.numpy()) <= 5 * tol
    if use_bn:  # moving statistics, momentum 0.9 (gnn_layers.py:53)
        zz = orc.torch_cpu_graph_conv(xr, g["Lt"], wr, g["K"], "chebyshev").detach()
        assert rel_err(layer.bn.moving_mean.detach().cpu().numpy(), 0.1 * zz.mean(dim=(0, 1)).numpy()) < 1e-4
        assert rel_err(layer.bn.moving_variance.detach().cpu().numpy(),
                       0.9 + 0.1 * zz.var(dim=(0, 1), unbiased=False).numpy()) < 1e-4


@pytest.mark.parametrize("F", [1, 2, 3, 5, 8, 16, 20, 64, 132])
@pytest.mark.parametrize("transpose", [False, True])
def test_spmm_generic_shapes_tail_and_transpose(F, transpose):
    """ds_spmm = utils.split_sparse_dense_matmul (+ fused axpy terms) on an unsymmetric random
    matrix with a few long rows (CSR tail) and empty rows."""
    rng = np.random.default_rng(F)
    M = 517
    A = sparse.random(M, M, density=0.01, random_state=7, format="lil")
    A[5, :] = rng.standard_normal(M)  # long rows -> tail
    A[300, ::3] = 1.5
    A[17, :] = 0  # empty row
    A = sparse.csr_matrix(A)
    plan = utils.plan_from_sparse(A, ell_width=0)
    info = plan.info(0)
    assert info["tail_rows"] >= 2 and info["symmetric"] == 0 and info["nnz"] == A.nnz
    B = 3
    x, prev, add = (rng.standard_normal((B, M, F)).astype(np.float32) for _ in range(3))
    out = _ops.spmm(plan, dev(x), 2.0, dev(prev), -1.0, dev(add), 0.5, transpose=transpose).detach().cpu().numpy()
    Aop = A.T.tocsr() if transpose else A
    ref = np.stack([2.0 * (Aop @ x[b].astype(np.float64)) - prev[b] + 0.5 * add[b] for b in range(B)])
    assert rel_err(out, ref) <= TOL_FP32
    # the 2-D reference signature
    out2 = utils.split_sparse_dense_matmul(plan, dev(x[0]), n_splits=4).detach().cpu().numpy()
    assert rel_err(out2, A @ x[0].astype(np.float64)) <= TOL_FP32


@pytest.mark.parametrize("K", [1, 2, 3])
@pytest.mark.parametrize("cls", ["Chebyshev", "Monomial"])
def test_small_K_and_default_fout(K, cls):
    g = load_golden("cheb_nside4_k8")
    rng = np.random.default_rng(K)
    x = rng.standard_normal((2, 192, 3))
    layer = getattr(gnn_layers, cls)(L=g["L"], K=K)  # Fout=None -> Fout = Fin (gnn_layers.py:85-88)
    xt = dev(x).requires_grad_(True)
    y = layer(xt)
    assert tuple(y.shape) == (2, 192, 3)
    w = layer.kernel.detach().double().detach().cpu().numpy()
    Lt, _ = orc.prepare_laplacian(g["L"], 0.75 if cls == "Chebyshev" else 1.0)
    rec = cls.lower()
    assert rel_err(y.detach().cpu().numpy(), orc.graph_conv_forward(x, Lt, w, K, rec, dtype=np.float64)) <= TOL_FP32
    dy = rng.standard_normal((2, 192, 3))
    y.backward(dev(dy))
    dx, dk, _ = orc.graph_conv_backward(x, Lt, w, K, dy, rec)
    assert rel_err(xt.grad.detach().cpu().numpy(), dx) <= TOL_FP32
    assert rel_err(layer.kernel.grad.detach().cpu().numpy(), dk) <= TOL_FP32


def test_unsymmetric_laplacian_gradient():
    """The backward must use L~^T, not L~ (only visible for an unsymmetric operator)."""
    rng = np.random.default_rng(5)
    M = 40
    Lraw = sparse.random(M, M, density=0.15, random_state=3).toarray() + np.eye(M)
    layer = gnn_layers.Monomial(L=Lraw, K=4, Fout=3)
    x = rng.standard_normal((2, M, 5))
    xt = dev(x).requires_grad_(True)
    y = layer(xt)
    dy = rng.standard_normal((2, M, 3))
    y.backward(dev(dy))
    Lt = sparse.csr_matrix((layer._L_values.astype(np.float64), (layer._L_indices[:, 0], layer._L_indices[:, 1])),
                           shape=(M, M))
    w = layer.kernel.detach().double().detach().cpu().numpy()
    dx, dk, _ = orc.graph_conv_backward(x, Lt, w, 4, dy, "monomial")
    assert rel_err(y.detach().cpu().numpy(), orc.graph_conv_forward(x, Lt, w, 4, "monomial", dtype=np.float64)) <= 2e-5
    assert rel_err(xt.grad.detach().cpu().numpy(), dx) <= 2e-5 and rel_err(layer.kernel.grad.detach().cpu().numpy(), dk) <= 2e-5


def test_pool_bit_exact_and_reference_known_answer():
    g = np.load(f"{GOLDEN}/pool_nside4.npz")
    m_in = g["m_in"]  # float64 in the reference test; the layer computes in floatx = float32
    avg = healpy_layers.HealpyPool(1, pool_type="AVG")(m_in[None, :, None]).detach().cpu().numpy().ravel()
    mx = healpy_layers.HealpyPool(1, pool_type="MAX")(m_in[None, :, None]).detach().cpu().numpy().ravel()
    assert np.all(np.abs(g["avg"] - avg) < 1e-5) and np.all(np.abs(g["max"] - mx) < 1e-5)  # the reference's bar
    rng = np.random.default_rng(0)
    for (B, M, F, p) in [(2, 192, 1, 1), (3, 768, 5, 2), (2, 3072, 16, 3), (1, 768, 64, 1), (2, 48, 3, 1)]:
        x = rng.standard_normal((B, M, F)).astype(np.float32)
        x[0, :4, 0] = 1.0  # ties: first maximum takes the gradient
        for typ in ("MAX", "AVG"):
            xt = torch.tensor(x).cuda().requires_grad_(True)
            y = healpy_layers.HealpyPool(p, pool_type=typ)(xt)
            ref = orc.healpy_pool(x, p, typ)
            assert np.array_equal(y.detach().cpu().numpy(), ref), (B, M, F, p, typ)  # bit-exact
            dy = rng.standard_normal(ref.shape).astype(np.float32)
            y.backward(torch.tensor(dy).cuda())
            assert np.array_equal(xt.grad.detach().cpu().numpy(), orc.healpy_pool_backward(x, dy, p, typ)), (typ, p)
    with pytest.raises(IOError):
        healpy_layers.HealpyPool(1)(np.zeros((1, 10, 1), np.float32))


@pytest.mark.parametrize("act", [None, "elu"])
def test_pseudo_conv_and_transpose(act):
    g = np.load(f"{GOLDEN}/pconv_nside8.npz")
    x = g["x"]
    pc = healpy_layers.HealpyPseudoConv(p=2, Fout=5, activation=act)
    pc.build_from_shape(x.shape)
    with torch.no_grad():
        pc.kernel.copy_(dev(g["w"]))
        pc.bias.copy_(dev(g["b"]))
    xt = dev(x).requires_grad_(True)
    y = pc(xt)
    ref = orc.pseudo_conv(x, g["w"], g["b"], act)
    assert tuple(y.shape) == (2, 48, 5) and rel_err(y.detach().cpu().numpy(), ref) <= TOL_FP32
    if act == "elu":
        assert rel_err(y.detach().cpu().numpy(), g["y"]) <= TOL_FP32
    # gradients against torch's own conv (float64 CPU)
    xr = torch.tensor(x, requires_grad=True)
    wr = torch.tensor(g["w"], requires_grad=True)
    br = torch.tensor(g["b"], requires_grad=True)
    yr = torch.nn.functional.conv1d(xr.permute(0, 2, 1), wr.permute(2, 1, 0), br, stride=16).permute(0, 2, 1)
    yr = torch.nn.functional.elu(yr) if act else yr
    dy = np.random.default_rng(1).standard_normal(ref.shape)
    y.backward(dev(dy))
    yr.backward(torch.tensor(dy))
    assert rel_err(xt.grad.detach().cpu().numpy(), xr.grad.numpy()) <= TOL_FP32
    assert rel_err(pc.kernel.grad.detach().cpu().numpy(), wr.grad.numpy()) <= TOL_FP32
    assert rel_err(pc.bias.grad.detach().cpu().numpy(), br.grad.numpy()) <= TOL_FP32

    xs = x[:, :48]
    pt = healpy_layers.HealpyPseudoConv_Transpose(p=2, Fout=5, activation=act)
    pt.build_from_shape(xs.shape)
    with torch.no_grad():
        pt.kernel.copy_(dev(g["wt"]))
        pt.bias.copy_(dev(g["b"]))
    xt = dev(xs).requires_grad_(True)
    y = pt(xt)
    ref = orc.pseudo_conv_transpose(xs, g["wt"], g["b"], act)
    assert tuple(y.shape) == (2, 768, 5) and rel_err(y.detach().cpu().numpy(), ref) <= TOL_FP32
    xr = torch.tensor(xs, requires_grad=True)
    wr = torch.tensor(g["wt"], requires_grad=True)
    br = torch.tensor(g["b"], requires_grad=True)
    yr = torch.nn.functional.conv_transpose2d(xr.permute(0, 2, 1)[:, :, None, :], wr.permute(3, 2, 0, 1), br,
                                              stride=(1, 16))[:, :, 0].permute(0, 2, 1)
    yr = torch.nn.functional.elu(yr) if act else yr
    dy = np.random.default_rng(2).standard_normal(ref.shape)
    y.backward(dev(dy))
    yr.backward(torch.tensor(dy))
    assert rel_err(xt.grad.detach().cpu().numpy(), xr.grad.numpy()) <= TOL_FP32
    assert rel_err(pt.kernel.grad.detach().cpu().numpy(), wr.grad.numpy()) <= TOL_FP32
    assert rel_err(pt.bias.grad.detach().cpu().numpy(), br.grad.numpy()) <= TOL_FP32


def test_reference_shape_tests():
    """tests/test_healpy_layers.py:40-63 (nside 8, p = 3, Fout = 5) and tests/test_gnn_layers.py:104-137."""
    np.random.seed(11)
    m_in = np.random.normal(size=768)
    assert tuple(healpy_layers.HealpyPseudoConv(3, 5)(m_in[None, :, None]).shape) == (1, 768 // 64, 5)
    assert tuple(healpy_layers.HealpyPseudoConv_Transpose(3, 5)(m_in[None, :, None]).shape) == (1, 768 * 64, 5)
    x = np.random.normal(size=(3, 192, 7)).astype(np.float32)
    kw = {"L": np.eye(192), "K": 5, "activation": "relu", "regularizer": lambda w: (w**2).sum()}
    assert tuple(gnn_layers.GCNN_ResidualLayer("CHEBY", kw)(x).shape) == (3, 192, 7)
    res = gnn_layers.GCNN_ResidualLayer("MONO", kw, activation="relu", use_bn=True, norm_type="layer_norm",
                                        bn_kwargs={"axis": (1, 2)})
    assert tuple(res(x, training=True).shape) == (3, 192, 7)


def test_property_constant_eigenvector_closed_form():
    """Size-independent property: for a normalised Laplacian, v = D^(1/2) 1 has L v = 0, hence
    L~ v = -v and T_k(L~) v = (-1)^k v (Chebyshev) — a closed form for the whole recursion."""
    for nside in (16, 64):
        g = SphereHealpix(nside, k=8)
        v = np.sqrt(np.asarray(g.W.sum(axis=1)).ravel())
        K, Fin, Fout = 6, 4, 3
        layer = gnn_layers.Chebyshev(L=g.L, K=K, Fout=Fout)
        x = np.repeat(v[None, :, None], Fin, axis=2).astype(np.float32) * np.arange(1, Fin + 1, dtype=np.float32)
        y = layer(x).detach().cpu().numpy()
        w = layer.kernel.detach().double().detach().cpu().numpy().reshape(Fin, K, Fout)
        coef = np.einsum("f,fko,k->o", np.arange(1, Fin + 1.0), w, (-1.0) ** np.arange(K))
        ref = v[None, :, None] * coef[None, None, :]
        assert rel_err(y, ref) <= 5e-5  # K hops of fp32 round-off on top of the fp32 L~ values
