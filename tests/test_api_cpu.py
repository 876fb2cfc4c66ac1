"""Host-side contract of the drop-in (no GPU): the C-ABI library loads and exports every symbol
the header declares, the layer / model API mirrors the reference's names, errors, weights and
parameter counts, and the hot path fails loudly without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import deepsphere
from deepsphere import _native as nat
from deepsphere import gnn_layers, healpy_layers, keras_compat
from deepsphere.graph import SphereHealpix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "deepsphere_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ds_[a-zA-Z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(nat.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert declared == set(nat.SIGNATURES), "ctypes binding and header disagree"
    assert nat.lib().ds_abi_version() == 1
    assert nat.lib().ds_device_count() >= 0


def test_entry_points_reject_bad_arguments_without_a_gpu():
    L = nat.lib()
    assert L.ds_pool_forward(1, 10, 1, 1, 0, None, None, None) != 0  # 10 % 4 != 0
    assert b"not compatible" in L.ds_last_error()
    assert L.ds_pool_forward(1, 16, 1, 0, 0, None, None, None) != 0  # p = 0
    assert L.ds_pool_forward(1, 16, 1, 1, 7, None, None, None) != 0  # unknown pooling type
    out = ctypes.c_void_p()
    idx = np.zeros((1, 2), dtype=np.int64)
    val = np.ones(1, dtype=np.float32)
    rc = L.ds_plan_create_coo(0, 1, idx.ctypes.data_as(ctypes.c_void_p), val.ctypes.data_as(ctypes.c_void_p), 0,
                              ctypes.byref(out))
    assert rc != 0 and out.value is None


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU failure mode")
def test_hot_path_fails_loudly_without_cuda():
    L = np.eye(12)
    layer = gnn_layers.Chebyshev(L=L, K=3, Fout=2)
    with pytest.raises(nat.NativeError, match="no CPU fallback"):
        layer(np.zeros((1, 12, 4), np.float32))
    with pytest.raises(nat.NativeError):
        healpy_layers.HealpyPool(1)(np.zeros((1, 48, 1), np.float32))
    with pytest.raises(nat.NativeError):
        nat.GraphPlan(np.zeros((1, 2), np.int64), np.ones(1, np.float32), (3, 3)).handle(0)


def test_reference_exceptions():
    # healpy_layers.py:39-40,64-65 / tests/test_healpy_layers.py:16-19
    with pytest.raises(IOError):
        healpy_layers.HealpyPool(0, pool_type="MAX")
    with pytest.raises(IOError):
        healpy_layers.HealpyPool(2, pool_type="HUHU")
    with pytest.raises(IOError):
        healpy_layers.HealpyPseudoConv(0, 5)
    with pytest.raises(IOError):
        healpy_layers.HealpyPseudoConv_Transpose(0, 5)
    # gnn_layers.py:55-60
    with pytest.raises(ValueError):
        gnn_layers.Chebyshev(L=np.eye(8), K=3, activation="not_an_activation")
    # tests/test_gnn_layers.py:99-100,138-145
    with pytest.raises(IOError):
        gnn_layers.GCNN_ResidualLayer("HUHU", {"L": np.eye(8), "K": 3})
    with pytest.raises(ValueError):
        gnn_layers.GCNN_ResidualLayer("CHEBY", {"L": np.eye(8), "K": 3}, use_bn=True, norm_type="moving_norm")
    # healpy_networks.py:39-42 / tests/test_healpy_networks.py:155-156
    with pytest.raises(NotImplementedError):
        deepsphere.HealpyGCNN(4, np.arange(192), [healpy_layers.HealpyChebyshev(K=3, Fout=2)], n_neighbors=12)
    # healpy_networks.py:60-65: too many reductions / inconsistent index set (:80-86)
    with pytest.raises(ValueError):
        deepsphere.HealpyGCNN(2, np.arange(48), [healpy_layers.HealpyPool(p=2)])
    with pytest.raises(ValueError):
        deepsphere.HealpyGCNN(4, np.arange(192)[::4], [healpy_layers.HealpyPool(p=1)])


def test_laplacian_prep_matches_reference_recipe():
    """gnn_layers.py:64-72: lmax = 1.02*eigsh, scale 0.75 (Chebyshev) / 1 (Monomial), COO int64 +
    float32, and the caller's matrix is left untouched."""
    g = SphereHealpix(4, k=8)
    before = g.L.copy()
    cheb = gnn_layers.Chebyshev(L=g.L, K=3)
    mono = gnn_layers.Monomial(L=g.L, K=3)
    assert abs(g.L - before).max() == 0
    lam = np.linalg.eigvalsh(g.L.toarray())
    assert abs(cheb.lmax - 1.02 * lam.max()) < 1e-9
    assert cheb._L_indices.dtype == np.int64 and cheb._L_indices.shape == (g.L.nnz, 2)
    assert cheb._L_values.dtype == np.float32
    dense = np.zeros((192, 192))
    dense[cheb._L_indices[:, 0], cheb._L_indices[:, 1]] = cheb._L_values
    assert np.abs(dense - (g.L.toarray() * (1.5 / cheb.lmax) - np.eye(192))).max() < 1e-7
    densem = np.zeros((192, 192))
    densem[mono._L_indices[:, 0], mono._L_indices[:, 1]] = mono._L_values
    assert np.abs(densem - (g.L.toarray() * (2.0 / mono.lmax) - np.eye(192))).max() < 1e-7
    # every diagonal entry of L~ equals a - 1 for a normalised Laplacian
    assert np.allclose(np.diag(dense), 1.5 / cheb.lmax - 1, atol=1e-7)
    # accepted L types: numpy, scipy sparse, torch tensor (tests/test_healpy_layers.py:79-80)
    gnn_layers.Chebyshev(L=g.L.toarray(), K=2)
    gnn_layers.Chebyshev(L=torch.tensor(g.L.toarray()), K=2)


def test_weights_follow_reference_shapes_and_initialisers():
    keras_compat.reset_name_counts()
    layer = gnn_layers.Chebyshev(L=np.eye(192), K=5, Fout=7, use_bias=True, use_bn=True)
    layer.build_from_shape((3, 192, 4))
    assert tuple(layer.kernel.shape) == (20, 7) and tuple(layer.bias.shape) == (1, 1, 7)
    std = 1 / np.sqrt(4 * (5 + 0.5) / 2)
    assert layer.kernel.abs().max() <= 2 * std + 1e-6  # TruncatedNormal, gnn_layers.py:92-93
    assert layer.bias.abs().max() <= np.sqrt(6 / (1 + 7)) + 1e-6  # add_weight default glorot_uniform
    assert [n for n, _ in layer.named_buffers()] == ["bn.moving_mean", "bn.moving_variance"]
    mono = gnn_layers.Monomial(L=np.eye(192), K=5)
    mono.build_from_shape((3, 192, 4))
    assert tuple(mono.kernel.shape) == (20, 4) and mono.kernel.abs().max() <= 0.2 + 1e-6
    pc = healpy_layers.HealpyPseudoConv(p=1, Fout=5)
    pc.build_from_shape((1, 192, 3))
    assert tuple(pc.kernel.shape) == (4, 3, 5) and tuple(pc.bias.shape) == (5,)
    pt = healpy_layers.HealpyPseudoConv_Transpose(p=1, Fout=5)
    pt.build_from_shape((1, 48, 3))
    assert tuple(pt.kernel.shape) == (1, 4, 5, 3)
    # kwargs flow through to add_weight (tests/test_gnn_layers.py:104-109 passes `regularizer`)
    reg = gnn_layers.Chebyshev(L=np.eye(192), K=5, regularizer=lambda w: (w**2).sum())
    reg.build_from_shape((3, 192, 2))
    assert len(reg.losses) == 1 and float(reg.losses[0]) > 0
    # custom initializer object (tests/test_gnn_layers.py:19-21)
    ini = gnn_layers.Chebyshev(L=np.eye(8), K=4, Fout=3, initializer=keras_compat.RandomNormal(stddev=0.3, seed=13))
    ini.build_from_shape((5, 8, 7))
    assert tuple(ini.kernel.shape) == (28, 3)


def _quick_start_layers():
    hl = healpy_layers
    return [
        hl.HealpyChebyshev(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu"),
        hl.HealpyPool(p=1),
        hl.HealpyChebyshev(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu"),
        hl.HealpyPool(p=1),
        hl.HealpyChebyshev(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu"),
        hl.HealpyPool(p=1),
        hl.HealpyChebyshev(K=10, Fout=2),
        keras_compat.Lambda(lambda x: torch.softmax(x.mean(dim=1), dim=-1)),
    ]


def test_quick_start_model_param_counts_and_names():
    """examples/quick_start.ipynb:118-127,188: per-layer params 65 / 265 / 265 / 100, total 695."""
    nside = 16
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(12 * nside**2), layers=_quick_start_layers(),
                                  n_neighbors=20)
    model.build(input_shape=(None, 12 * nside**2, 1))
    counts = [layer.count_params() for layer in model.layers]
    assert counts == [65, 0, 265, 0, 265, 0, 100, 0]
    assert model.summary(print_fn=lambda s: None) == 695
    assert [layer.name for layer in model.layers[:4]] == ["chebyshev", "healpy_pool", "chebyshev_1", "healpy_pool_1"]
    assert model.get_layer("chebyshev_1") is model.layers[2] and model.get_layer(index=6).K == 10
    assert model._summary_shapes[-1] == (1, 2) and model._summary_shapes[5] == (1, 12 * 2**2, 5)
    with pytest.raises(ValueError):
        model.get_layer("nope")


def test_generative_model_param_counts():
    """examples/generative_models.ipynb:245-292: PseudoConv p=1 1->4: 20; 4->8: 136; 8->16: 528;
    Chebyshev K=5 16->16: 1296 (bias); PseudoConv_Transpose 16->16: 1040; last transpose to 1: 65."""
    hl = healpy_layers
    nside = 8
    layers = [hl.HealpyPseudoConv(p=1, Fout=4), hl.HealpyPseudoConv(p=1, Fout=8), hl.HealpyPseudoConv(p=1, Fout=16),
              hl.HealpyChebyshev(K=5, Fout=16, use_bias=True, activation="elu"),
              hl.HealpyPseudoConv_Transpose(p=1, Fout=16), hl.HealpyPseudoConv_Transpose(p=1, Fout=16),
              hl.HealpyPseudoConv_Transpose(p=1, Fout=1)]
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(12 * nside**2), layers=layers, n_neighbors=8)
    model.build(input_shape=(None, 12 * nside**2, 1))
    assert [l.count_params() for l in model.layers] == [20, 136, 528, 1296, 1040, 1040, 65]
    assert model.nside_out == 8 and model._summary_shapes[-1] == (1, 768, 1)
    assert "gcnn__residual_layer" == keras_compat.to_snake_case("GCNN_ResidualLayer")


def test_healpy_gcnn_index_bookkeeping_masked():
    """healpy_networks.py:98-188 on the advanced_tutorial index set (24 832 -> ... -> 388)."""
    from deepsphere import healpix as hpx, utils

    ext = utils.extend_indices(hpx.query_disc(64, [1, 0, 0], 1.5), 64, 8)
    hl = healpy_layers
    layers = [hl.HealpyChebyshev(K=3, Fout=5), hl.HealpyPool(p=1), hl.HealpyMonomial(K=3, Fout=5),
              hl.HealpyPool(p=1, pool_type="AVG"), hl.HealpyPseudoConv(p=1, Fout=2)]
    model = deepsphere.HealpyGCNN(nside=64, indices=ext, layers=layers, n_neighbors=8, max_batch_size=16,
                                  initial_Fin=1)
    model.build(input_shape=(None, len(ext), 1))
    assert [s[1] for s in model._summary_shapes] == [24832, 6208, 6208, 1552, 388]
    assert model.layers[0]._plan.M == 24832 and model.layers[2]._plan.M == 6208
    assert model.layers[0].n_matmul_splits == 1
    with pytest.raises(ValueError):  # not closed under the reduction -> must call extend_indices first
        deepsphere.HealpyGCNN(nside=64, indices=hpx.query_disc(64, [1, 0, 0], 1.5), layers=layers)


def test_save_load_weights_roundtrip(tmp_path):
    """tests/test_healpy_networks.py:132-152: save -> new model differs -> load -> equal."""
    def make(seed):
        torch.manual_seed(seed)
        m = deepsphere.HealpyGCNN(4, np.arange(192), [healpy_layers.HealpyChebyshev(K=3, Fout=4, use_bias=True,
                                                                                      use_bn=True),
                                                      healpy_layers.HealpyPseudoConv(p=1, Fout=2)])
        m.build(input_shape=(None, 192, 3))
        return m
    a, b = make(1), make(2)
    path = str(tmp_path / "w.npz")
    a.save_weights(path)
    assert not np.allclose(a.get_weights()[0], b.get_weights()[0])
    b.load_weights(path)
    for wa, wb in zip(a.get_weights(), b.get_weights()):
        assert np.allclose(wa, wb, atol=1e-6)


def test_losses_aggregate_over_nested_layers():
    """Keras reports the regularisation losses of nested layers on the outer model: HealpyGCNN and the
    residual layer must not drop the `regularizer=` kwargs of their graph layers."""
    import deepsphere
    from deepsphere import gnn_layers, healpy_layers as hl

    reg = lambda w: (w ** 2).sum()  # noqa: E731
    layers = [hl.HealpyChebyshev(K=3, Fout=4, regularizer=reg), hl.HealpyPool(p=1), hl.HealpyMonomial(K=2, Fout=2, regularizer=reg)]
    model = deepsphere.HealpyGCNN(nside=4, indices=np.arange(192), layers=layers)
    model.build(input_shape=(None, 192, 1))
    assert len(model.losses) == 2
    res = gnn_layers.GCNN_ResidualLayer("CHEBY", {"L": np.eye(12), "K": 2, "regularizer": reg})
    res.build_from_shape((1, 12, 3))
    assert len(res.losses) == 2 and all(float(v) > 0 for v in res.losses)


def test_partitioned_model_rejects_incomplete_sibling_groups():
    """A masked index set whose COUNT is compatible with the pooling depth but whose 4^p groups are incomplete must raise
    the same ValueError as HealpyGCNN."""
    from deepsphere import healpy_layers as hl, partition

    idx = np.arange(0, 192 * 4, 1)[::3][:64 * 3]  # 192 pixels of nside 8, not closed under 4-groups
    assert len(idx) % 4 == 0
    with pytest.raises(ValueError):
        partition.PartitionedHealpyGCNN(8, idx, [hl.HealpyPool(p=1)], rank=0, world=1)


def test_healpy_gcnn_takes_the_graphs_of_a_user_supplied_builder():
    """the built-in graph builder is not pinned against the PyGSP fork the reference calls
    (healpy_networks.py:110-118), so a checkpoint trained there needs its own Laplacians: `graph_builder` is called
    once per (nside, index set) with the reference's arguments and its `.L` is what the layers get."""
    import deepsphere
    from deepsphere import healpy_layers as hl
    from deepsphere.graph import SphereHealpix
    from helpers import orc

    calls = []

    class Scaled:  # a builder whose L is recognisably not the built-in one
        def __init__(self, nside, idx, k):
            calls.append((nside, len(idx), k))
            self.L = SphereHealpix(nside, indexes=idx, k=k).L * 0.5

    nside = 8
    idx = np.arange(12 * nside**2)
    layers = [hl.HealpyChebyshev(K=3, Fout=4), hl.HealpyChebyshev(K=3, Fout=4), hl.HealpyPool(p=1),
              hl.HealpyMonomial(K=2, Fout=2)]
    model = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=layers, n_neighbors=20, graph_builder=Scaled)
    assert calls == [(8, 768, 20), (4, 192, 20)]  # one graph per resolution, k passed through
    ref = SphereHealpix(nside, k=20).L.tocoo()
    got = model.layers_use[0]
    dense = np.zeros((768, 768))
    dense[got._L_indices[:, 0], got._L_indices[:, 1]] = got._L_values
    expect, _ = orc.prepare_laplacian(ref.tocsr() * 0.5, 0.75)
    assert np.abs(dense - expect.toarray()).max() <= 1e-6
    with pytest.raises(ValueError, match="graph_builder must return"):
        deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=[hl.HealpyChebyshev(K=2, Fout=2)],
                              graph_builder=lambda n, i, k: object())
