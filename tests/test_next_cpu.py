"""SURVEY 8f rows on the CPU: Bernstein (basis change to the Chebyshev recursion), HealpySmoothing (kernel builder,
constructor semantics, cache files) and the get_gsp_filters weight view — oracle, golden vectors and host logic.
No compute calls into the CUDA library."""
import os

import numpy as np
import pytest
import torch
from scipy import sparse

import deepsphere
from deepsphere import gnn_layers, healpix as hpx, healpy_layers as hl, keras_compat as kc
from deepsphere.graph import SphereHealpix
from helpers import GOLDEN, load_golden, orc, rel_err


# ----------------------------------------------------------------------------------------------- Bernstein
@pytest.mark.parametrize("name", ["bern_nside4_k8", "bern_masked16_k20"])
def test_bernstein_oracle_reproduces_golden_and_autograd(name):
    g = load_golden(name)
    Lt, lmax = orc.prepare_laplacian(g["L"], 0.75)  # gnn_layers.py:473: same prep as Chebyshev
    assert abs(lmax - float(g["lmax"])) <= 1e-9 * abs(lmax) and abs(Lt - g["Lt"]).max() < 1e-12
    y = orc.bernstein_forward(g["x"], g["Lt"], g["kernel"], g["K"], dtype=np.float64)
    assert rel_err(y, g["y64"]) < 1e-13
    y32 = orc.bernstein_forward(g["x"].astype(np.float32), g["Lt"], g["kernel"].astype(np.float32), g["K"])
    assert rel_err(y32, g["y64"]) < 1e-5
    x = torch.tensor(g["x"], requires_grad=True)
    w = torch.tensor(g["kernel"], requires_grad=True)
    orc.torch_cpu_bernstein(x, g["Lt"], w, g["K"]).backward(torch.tensor(g["dy"]))
    assert rel_err(x.grad.numpy(), g["dx64"]) < 1e-12 and rel_err(w.grad.numpy(), g["dkernel64"]) < 1e-12


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 8, 12])
def test_bernstein_is_a_chebyshev_layer_with_transformed_weights(K):
    """The design of gnn_layers.Bernstein: the reference's K(K+1)/2 + K sparse products, stale last term included,
    equal the (K+1)-term Chebyshev recursion with W'_j = sum_i C[i, j] W_i."""
    g = load_golden("cheb_nside4_k8")
    rng = np.random.default_rng(K)
    Fin, Fout = 3, 4
    x = rng.standard_normal((2, g["Lt"].shape[0], Fin))
    W = rng.standard_normal(((K + 1) * Fin, Fout))
    ref = orc.bernstein_forward(x, g["Lt"], W, K, dtype=np.float64)
    C = gnn_layers.bernstein_to_chebyshev(K)
    Wc = np.einsum("ij,fio->fjo", C, W.reshape(Fin, K + 1, Fout)).reshape(-1, Fout)
    got = orc.graph_conv_forward(x, g["Lt"], Wc, K + 1, "chebyshev", dtype=np.float64)
    assert rel_err(got, ref) < 1e-11
    # and in the fp32 arithmetic the device uses, inside the fp32 parity bar
    got32 = orc.graph_conv_forward(x.astype(np.float32), g["Lt"], Wc.astype(np.float32), K + 1, "chebyshev")
    assert rel_err(got32, ref) < 1e-5


@pytest.mark.parametrize("K", [1, 3, 6])
def test_bernstein_polynomials_spectral_definition(K):
    """Independent formulation through the eigen-decomposition: column i of the reference's stack is
    theta_i (2 - L~)^(K-i) L~^i x for i < K, and 2^-K times column K-1 for i = K (the stale x3)."""
    from math import comb

    g = load_golden("cheb_nside4_k8")
    lam, V = np.linalg.eigh(g["Lt"].toarray())
    P = gnn_layers.bernstein_polynomials(K, stale_last_term=True)(lam)
    for i in range(K):
        assert np.allclose(P[i], comb(K, i) / 2**K * (2 - lam) ** (K - i) * lam**i, rtol=1e-13, atol=0)
    assert np.allclose(P[K], P[K - 1] / 2**K, rtol=1e-13, atol=0)
    textbook = gnn_layers.bernstein_polynomials(K, stale_last_term=False)(lam)
    assert np.allclose(textbook[K], lam**K / 2**K, rtol=1e-13, atol=1e-300)
    assert np.allclose(textbook.sum(axis=0), 1.0)  # a Bernstein basis on [0, 2] sums to one
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, len(lam), 1))
    W = rng.standard_normal((K + 1, 2))
    ref = orc.bernstein_forward(x, g["Lt"], W, K, dtype=np.float64)
    spec = sum(((V * P[i]) @ V.T @ x[0]) @ W[i:i + 1] for i in range(K + 1))
    assert rel_err(ref[0], spec) < 1e-11


def test_bernstein_layer_weights_follow_the_reference():
    kc.reset_name_counts()
    L = SphereHealpix(4, k=8).L
    torch.manual_seed(0)
    layer = gnn_layers.Bernstein(L=L, K=4, Fout=200, use_bias=True)
    layer.build_from_shape((None, 192, 100))
    assert tuple(layer.kernel.shape) == (5 * 100, 200) and tuple(layer.bias.shape) == (1, 1, 200)  # :500-508
    std = np.sqrt(6 / (100 + 200))  # gnn_layers.py:498
    w = layer.kernel.detach().numpy()
    assert np.abs(w).max() <= 2 * std + 1e-6 and abs(w.std() / (0.87962566 * std) - 1) < 0.02  # truncated normal
    assert layer.K == 4 and layer._n_terms == 5
    # Fout=None -> Fout = Fin (:490-493)
    l2 = gnn_layers.Bernstein(L=L, K=2)
    l2.build_from_shape((None, 192, 3))
    assert tuple(l2.kernel.shape) == (9, 3)
    with pytest.raises(ValueError):
        gnn_layers.Bernstein(L=L, K=2, activation="no_such_activation")  # :460-465
    # the device weights are the Chebyshev-basis image of the trainable Bernstein weights, inside autograd
    Wd = l2._device_kernel()
    C = gnn_layers.bernstein_to_chebyshev(2)
    ref = np.einsum("ij,fio->fjo", C, l2.kernel.detach().numpy().reshape(3, 3, 3)).reshape(9, 3)
    assert rel_err(Wd.detach().numpy(), ref) < 1e-6
    gW = torch.randn_like(Wd)
    (Wd * gW).sum().backward()
    gref = np.einsum("ij,fjo->fio", C, gW.numpy().reshape(3, 3, 3)).reshape(9, 3)
    assert rel_err(l2.kernel.grad.numpy(), gref) < 1e-6


def test_healpy_gcnn_accepts_bernstein_layers():
    kc.reset_name_counts()
    layers = [hl.HealpyBernstein(K=3, Fout=4, use_bias=True, activation="relu"), hl.HealpyPool(p=1),
              hl.HealpyBernstein(K=2, Fout=2)]
    model = deepsphere.HealpyGCNN(nside=4, indices=np.arange(192), layers=layers)
    model.build(input_shape=(None, 192, 1))
    assert isinstance(model.get_layer(index=0), gnn_layers.Bernstein)
    assert model.get_layer(index=0)._plan.M == 192 and model.get_layer(index=2)._plan.M == 48
    assert model.count_params() == (4 * 1 * 4 + 4) + 3 * 4 * 2


# ------------------------------------------------------------------------------------------ HealpySmoothing
def _golden_smoothing():
    d = np.load(os.path.join(GOLDEN, "smooth_masked16.npz"))
    return {k: d[k] for k in d.files}


def test_smoothing_oracle_reproduces_golden():
    g = _golden_smoothing()
    n = int(g["M"])
    theta, phi = hpx.pix2ang(int(g["nside"]), g["indices"])
    ind, val = orc.smoothing_neighbours(0.5 * np.pi - theta, phi, float(g["sigma_arcmin"]) * np.pi / (60 * 180), 3)
    assert np.array_equal(ind, g["ind_coo"]) and np.array_equal(val, g["val_coo"])
    Ks = orc.smoothing_kernel(ind, val, n)
    assert rel_err(orc.smoothing_forward(g["x"], Ks, dtype=np.float64), g["y_once64"]) < 1e-13
    assert rel_err(orc.smoothing_forward(g["x"], Ks, g["reps"], g["mask"][None], dtype=np.float64),
                   g["y_reps_mask64"]) < 1e-13
    # the normalisation as written in the reference: entry (i, j) / rowsum(j) -- columns of K^T sum to one
    raw = sparse.csr_matrix((val, (ind[:, 0], ind[:, 1])), shape=(n, n)).toarray().astype(np.float64)
    assert np.allclose(Ks.toarray(), raw / raw.sum(axis=1)[None, :], rtol=1e-6)
    # independent formulation of repeated smoothing: matrix powers
    Kd = Ks.toarray().astype(np.float64)
    y = np.stack([np.linalg.matrix_power(Kd, int(r)) @ g["x"][:, :, c].T for c, r in enumerate(g["reps"])], axis=2)
    assert rel_err(np.transpose(y, (1, 0, 2)) * g["mask"][None], g["y_reps_mask64"]) < 1e-12


class _KeepCoo(hl.HealpySmoothing):
    def _build_sparse_tensor(self):
        self.kept = (self.ind_coo.copy(), self.val_coo.copy())
        super()._build_sparse_tensor()


def test_smoothing_builder_matches_the_balltree_recipe():
    """cKDTree on unit vectors (product) vs BallTree/haversine on (lat, lon) (reference recipe): the same
    neighbour count, the same values on the common pattern; only exact distance ties at the cut may pick a
    different (equidistant) pixel."""
    g = _golden_smoothing()
    n = int(g["M"])
    layer = _KeepCoo(int(g["nside"]), g["indices"], sigma=float(g["sigma_arcmin"]))
    ind, val = layer.kept
    assert ind.dtype == np.int64 and val.dtype == np.float32 and ind.shape == g["ind_coo"].shape
    assert layer.max_neighbors == len(g["val_coo"]) // n
    A = sparse.csr_matrix((val, (ind[:, 0], ind[:, 1])), shape=(n, n))
    R = sparse.csr_matrix((g["val_coo"], (g["ind_coo"][:, 0], g["ind_coo"][:, 1])), shape=(n, n))
    common = (A != 0).multiply(R != 0)
    assert common.nnz >= 0.995 * R.nnz
    assert abs(A.multiply(common) - R.multiply(common)).max() <= 2e-7
    # rows are complete: every pixel is its own nearest neighbour with weight exactly 1
    assert np.all(A.diagonal() == 1.0)
    # the ties differ by value-preserving swaps: per-row value multisets agree
    va = np.sort(val.reshape(n, -1), axis=1)
    vr = np.sort(g["val_coo"].reshape(n, -1), axis=1)
    assert np.abs(va - vr).max() <= 2e-7
    assert layer._nnz == A.nnz and layer.sparse_kernel.M == n


def test_smoothing_constructor_semantics(tmp_path):
    idx = np.arange(192)
    with pytest.raises(AssertionError):
        hl.HealpySmoothing(4, idx)
    with pytest.raises(AssertionError):
        hl.HealpySmoothing(4, idx, fwhm=10.0, sigma=10.0)
    ident = hl.HealpySmoothing(4, idx, fwhm=0.0)
    assert ident.do_smoothing is False
    x = torch.randn(2, 192, 3)
    assert ident(x) is x  # identity layer, no device needed (healpy_layers.py:763-764)
    # list of scales -> smallest one builds the kernel, the others repeat ceil((s/s_min)^2) times (:595-622)
    lay = hl.HealpySmoothing(4, idx, fwhm=[900.0, 1800.0, 1000.0], data_path=str(tmp_path))
    assert lay.fwhm == 900.0 and list(lay.per_channel_repetitions) == [1, 4, 2]
    assert abs(lay.sigma - 900.0 / np.sqrt(8 * np.log(2))) < 1e-12
    assert abs(lay.sigma_rad - lay.sigma * np.pi / (60 * 180)) < 1e-15 and abs(lay.fwhm_arcmin - 900.0) < 1e-9
    label = f"-nside4-sigma{lay.sigma_arcmin:4.2f}-n_sigma3"
    assert lay.file_label == label
    f_ind, f_val = tmp_path / f"ind_coo{label}.npy", tmp_path / f"val_coo{label}.npy"
    assert f_ind.exists() and f_val.exists()  # :801-829
    with pytest.raises(AssertionError):
        hl.HealpySmoothing(4, idx, sigma=[300.0, 600.0], per_channel_repetitions=[1, 2])
    rad = hl.HealpySmoothing(4, idx, sigma=0.2, arcmin=False)
    assert abs(rad.sigma_arcmin - 0.2 / np.pi * 180 * 60) < 1e-9
    # a second layer loads the stored kernel instead of rebuilding it (:650-660): poison the file to prove it
    val = np.load(f_val)
    val[:] = 1.0
    np.save(f_val, val)
    again = _KeepCoo(4, idx, fwhm=900.0, data_path=str(tmp_path))
    assert np.all(again.kept[1] == 1.0) and not hasattr(again, "max_neighbors")
    # build(): shape checks and the mask ranks (:675-723)
    lay.build((None, 192, 3))
    assert lay.n_matmul_splits == 1 and lay.n_channels == 3
    with pytest.raises(AssertionError):
        hl.HealpySmoothing(4, idx, fwhm=900.0).build((None, 191, 3))
    with pytest.raises(AssertionError):
        hl.HealpySmoothing(4, idx, fwhm=900.0, per_channel_repetitions=[1, 2]).build((None, 192, 3))
    m1 = hl.HealpySmoothing(4, idx, fwhm=900.0, mask=np.ones(192, dtype=bool))
    m1.build((2, 192, 3))
    assert tuple(m1.mask.shape) == (1, 192, 1) and m1.mask.dtype == torch.float32
    m2 = hl.HealpySmoothing(4, idx, fwhm=900.0, mask=torch.ones(192, 3, dtype=torch.bool))
    m2.build((2, 192, 3))
    assert tuple(m2.mask.shape) == (1, 192, 3)
    with pytest.raises(Exception):  # no CPU fallback for the product
        lay(torch.randn(1, 192, 3))


# ------------------------------------------------------------------------------------------ get_gsp_filters
def test_get_gsp_filters_weight_view():
    """healpy_networks.py:190-289: [K, Fout, Fin] views of the f*K + k ordered kernel."""
    kc.reset_name_counts()
    layers = [hl.HealpyChebyshev(K=4, Fout=3), hl.HealpyMonomial(K=2, Fout=3),
              hl.Healpy_ResidualLayer("CHEBY", {"K": 3})]
    model = deepsphere.HealpyGCNN(nside=4, indices=np.arange(192), layers=layers)
    model.build(input_shape=(None, 192, 2))
    cheb = model.get_layer(index=0)
    (w,) = model.get_gsp_filters(0, return_weights=True)
    assert w.shape == (4, 3, 2)
    kern = cheb.kernel.detach().numpy()
    for f in range(2):
        for k in range(4):
            assert np.array_equal(w[k, :, f], kern[f * 4 + k])
    assert model.get_gsp_filters(cheb.name, return_weights=True)[0].shape == (4, 3, 2)
    assert model.get_gsp_filters(0, ind_in=[1], ind_out=[0, 2], return_weights=True)[0].shape == (4, 2, 1)
    w1, w2 = model.get_gsp_filters(2, return_weights=True)  # residual layer: Fout is None -> inferred (:203-205)
    assert w1.shape == (3, 3, 3) and w2.shape == (3, 3, 3)
    with pytest.raises(ValueError):
        model.get_gsp_filters(1)  # Monomial
    with pytest.raises(ValueError):
        model.get_gsp_filters(1.5)
    # the filter objects evaluate sum_k c_k T_k(1.5 lam / lmax - 1): check against the layer's own operator
    (flt,) = model.get_gsp_filters(0)
    L = SphereHealpix(4, k=8).L.toarray()
    lam, V = np.linalg.eigh(L)
    resp = flt.evaluate(lam)  # [Fout, Fin, M]
    assert resp.shape == (3, 2, 192)
    Lt, _ = orc.prepare_laplacian(L, 0.75)
    x = np.random.default_rng(0).standard_normal((1, 192, 2))
    ref = orc.graph_conv_forward(x, Lt, kern.astype(np.float64), 4, "chebyshev", dtype=np.float64)
    spec = np.einsum("mj,ofj,jf->mo", V, resp, V.T @ x[0])
    assert rel_err(spec, ref[0]) < 1e-9


# ------------------------------------------------------------------------------------- Laplacian prep: lmax
def test_largest_eigenvalue_matches_the_reference_arpack_call():
    """utils.largest_eigenvalue replaces `eigsh(L, k=1, which="LM")` (gnn_layers.py:66) for large symmetric L by an
    un-restarted Lanczos: same value to fp64 round-off on full-sphere, k-NN and masked graphs, on a matrix whose
    largest-magnitude eigenvalue is negative, and on the identity (Krylov breakdown at step 0)."""
    from scipy.sparse.linalg import eigsh

    from deepsphere import utils

    disc = hpx.query_disc(64, [1, 0, 0], 1.0)
    ext = orc.extend_indices(disc, 64, 8)
    cases = [SphereHealpix(32, k=8).L, SphereHealpix(32, k=20).L, SphereHealpix(64, indexes=ext, k=8).L,
             sparse.identity(5000, format="csr"), sparse.diags(np.linspace(-3, 2, 6000)).tocsr()]
    for L in cases:
        L = sparse.csr_matrix(L, dtype=np.float64)
        ref = float(eigsh(L, k=1, which="LM", return_eigenvectors=False)[0])
        assert abs(utils.largest_eigenvalue(L) - ref) <= 1e-11 * abs(ref)
    # small or unsymmetric matrices take the reference call itself
    A = sparse.random(6000, 6000, density=0.001, random_state=1, format="csr")
    assert utils.largest_eigenvalue(sparse.identity(10, format="csr")) == 1.0
    assert np.isfinite(utils.largest_eigenvalue(A + sparse.identity(6000)))
    # and the layer's lmax is the reference recipe's 1.02 * lambda_max
    L = SphereHealpix(32, k=8).L
    layer = gnn_layers.Chebyshev(L=L, K=3)
    _, lmax = orc.prepare_laplacian(L, 0.75)
    assert abs(layer.lmax - lmax) <= 1e-11 * lmax


# ------------------------------------------------------------------------------------- bench.py host logic
def test_bench_experimental_child_validation(monkeypatch):
    """bench.py runs the opt-in kernels in a child process and accepts its number only if the first training step
    (loss, per-parameter gradient norms) reproduces the default path's; a crash or a mismatch is reported, never
    raised into the parent's measurement."""
    import argparse
    import json
    import subprocess
    import types

    import bench

    args = argparse.Namespace(mode="tf32", model_nside=256, model_batch=16)
    base = {"ms_per_step": 8.0, "final_loss": 0.7, "first_step": {"loss": 1.25, "grad_norms": [0.5, 2.0, 1e-3]}}

    monkeypatch.setenv("RANK", "0")

    def fake_run(out_obj=None, rc=0, stderr=""):
        def run(cmd, env=None, **kw):
            assert env["DEEPSPHERE_SKINNY"] == "1" and "--model-only" in cmd and "RANK" not in env
            stdout = "log line\n" + (json.dumps(out_obj) + "\n" if out_obj is not None else "")
            return types.SimpleNamespace(returncode=rc, stdout=stdout, stderr=stderr)
        return run

    good = {"ms_per_step": 6.4, "final_loss": 0.7, "first_step": {"loss": 1.25 * (1 + 2e-6), "grad_norms": [0.5, 2.0, 1e-3]}}
    monkeypatch.setattr(subprocess, "run", fake_run(good))
    sw = {"DEEPSPHERE_SKINNY": "1"}
    r = bench.experimental_model_run(args, base, sw)
    assert r["validated"] is True and abs(r["speedup_vs_default"] - 1.25) < 1e-12 and r["switch"] == "DEEPSPHERE_SKINNY=1"
    # a graph replay must also reproduce the final loss
    r = bench.experimental_model_run(args, base, sw, ("--model-graph",))
    assert r["validated"] is True and r["switch"] == "DEEPSPHERE_SKINNY=1 --model-graph"
    drift = dict(good, final_loss=0.71)
    monkeypatch.setattr(subprocess, "run", fake_run(drift))
    assert bench.experimental_model_run(args, base, sw, ("--model-graph",))["validated"] is False
    bad = {"ms_per_step": 6.4, "final_loss": 0.7, "first_step": {"loss": 1.25, "grad_norms": [0.5, 2.1, 1e-3]}}
    monkeypatch.setattr(subprocess, "run", fake_run(bad))
    assert bench.experimental_model_run(args, base, sw)["validated"] is False
    monkeypatch.setattr(subprocess, "run", fake_run(None, rc=-11, stderr="CUDA error: an illegal memory access"))
    r = bench.experimental_model_run(args, base, sw)
    assert "error" in r and "illegal" in r["error"]

    def boom(*a, **k):
        raise subprocess.TimeoutExpired("bench", 600)

    monkeypatch.setattr(subprocess, "run", boom)
    assert "error" in bench.experimental_model_run(args, base, sw)


def test_lmax_table_entries_are_what_the_solver_returns(monkeypatch):
    """deepsphere/_lmax_table.json (tools/make_lmax_table.py) is keyed by a fingerprint of the matrix CONTENT: the entries of
    the small graphs are recomputed here with the cache bypassed; an unknown or perturbed matrix must miss."""
    import json

    from scipy import sparse

    from deepsphere import utils
    from deepsphere.graph import SphereHealpix

    path = os.path.join(os.path.dirname(utils.__file__), "_lmax_table.json")
    table = json.load(open(path))
    assert len(table) >= 8
    for nside, k in ((32, 8), (64, 8), (32, 20)):
        L = sparse.csr_matrix(SphereHealpix(nside, k=k).L, dtype=np.float64)
        key = utils.matrix_fingerprint(L)
        assert key in table
        utils._LMAX_MEMO.clear()
        cached = utils.largest_eigenvalue(L)
        monkeypatch.setenv("DEEPSPHERE_LMAX", "nocache")
        fresh = utils.largest_eigenvalue(L)
        monkeypatch.delenv("DEEPSPHERE_LMAX")
        assert cached == table[key] and abs(fresh - cached) <= 1e-12 * abs(fresh)
        L2 = L.copy()
        L2.data[7] *= 1.001
        assert utils.matrix_fingerprint(L2) != key
