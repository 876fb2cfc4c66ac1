"""Sphere partition with halo exchange on 2 gloo ranks, CPU only: the halo plan, the differentiable
exchange and the partitioned graph convolution reproduce the single-process result (forward, input gradient and
weight gradient).  The convolution itself is the float64 torch-CPU oracle here; the GPU test swaps in the CUDA layer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepsphere import distributed as dsd
from deepsphere import partition
from deepsphere.graph import SphereHealpix
from helpers import orc

K, F_IN, F_OUT, NSIDE, B = 4, 3, 2, 8, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _CpuCheb(torch.nn.Module):
    def __init__(self, Lt, w):
        super().__init__()
        self.Lt = Lt
        self.w = torch.nn.Parameter(w.clone())

    def forward(self, x):
        return orc.torch_cpu_graph_conv(x, self.Lt, self.w, K, "chebyshev")


def _problem():
    g = SphereHealpix(NSIDE, k=8)
    Lt, _ = orc.prepare_laplacian(g.L, 0.75)
    gen = torch.Generator().manual_seed(3)
    M = Lt.shape[0]
    x = torch.randn(B, M, F_IN, generator=gen, dtype=torch.float64)
    w = torch.randn(K * F_IN, F_OUT, generator=gen, dtype=torch.float64)
    dy = torch.randn(B, M, F_OUT, generator=gen, dtype=torch.float64)
    return Lt, x, w, dy


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dsd.init_from_env(backend="gloo")
    Lt, x, w, dy = _problem()
    conv = partition.PartitionedGraphConv(Lt, K - 1, lambda L_ext, rows: _CpuCheb(L_ext, w), align=4)
    b, e = conv.plan.own[rank]
    assert (b, e) == tuple(4 * v for v in dsd.shard_range(Lt.shape[0] // 4, rank, world))
    assert conv.plan.n_ext > conv.plan.n_own and conv.plan.n_ext < Lt.shape[0]  # a real halo, not the whole sphere
    x_own = x[:, b:e].clone().requires_grad_(True)
    y_own = conv(x_own)
    y_own.backward(dy[:, b:e])
    dsd.allreduce_gradients([conv.layer.w], average=False)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), y_own.detach().numpy())
    np.save(os.path.join(out_dir, f"dx{rank}.npy"), x_own.grad.numpy())
    np.save(os.path.join(out_dir, f"dw{rank}.npy"), conv.layer.w.grad.numpy())
    np.save(os.path.join(out_dir, f"halo{rank}.npy"), np.array([conv.plan.n_own, conv.plan.n_ext]))
    dist.destroy_process_group()


def test_partitioned_graph_conv_equals_single_process(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    Lt, x, w, dy = _problem()
    xg = x.clone().requires_grad_(True)
    wg = w.clone().requires_grad_(True)
    y = orc.torch_cpu_graph_conv(xg, Lt, wg, K, "chebyshev")
    y.backward(dy)
    y2 = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")], axis=1)
    dx2 = np.concatenate([np.load(tmp_path / "dx0.npy"), np.load(tmp_path / "dx1.npy")], axis=1)
    assert np.allclose(y2, y.detach().numpy(), atol=1e-11)
    assert np.allclose(dx2, xg.grad.numpy(), atol=1e-11)
    for r in (0, 1):  # summed over the group: every rank holds the full weight gradient
        assert np.allclose(np.load(tmp_path / f"dw{r}.npy"), wg.grad.numpy(), atol=1e-10)


import pytest  # noqa: E402


@pytest.mark.parametrize("k", [8, 20])
def test_halo_plan_is_the_hop_closure(k):
    """The halo is derived from the sparsity of L (k = 20 reaches about two pixel rings per hop): own + halo ==
    everything within n_hops, and the send / receive lists of the ranks mirror each other."""
    g = SphereHealpix(NSIDE, k=k)
    M = g.L.shape[0]
    plans = [partition.HaloPlan(g.L, 2, r, 3, align=4) for r in range(3)]
    A = (abs(g.L) > 0).astype(np.float64)
    A = ((A + A.T) > 0).astype(np.float64)
    for r, p in enumerate(plans):
        b, e = p.own[r]
        reach = np.zeros(M)
        reach[b:e] = 1
        for _ in range(2):
            reach = ((A @ reach) + reach > 0).astype(np.float64)
        assert np.array_equal(np.flatnonzero(reach), p.ext)
        assert np.array_equal(p.ext[p.own_pos], np.arange(b, e))
        for q, pq in enumerate(plans):
            if q != r:
                # what r sends to q (global rows) is what q expects from r
                assert np.array_equal(p.send_rows[q] + b, pq.ext[pq.recv_pos[r]])
    assert sum(p.n_own for p in plans) == M


# ---- the partitioned NETWORK through the real layer classes, with CPU stand-ins for the CUDA ops -------------------
def _install_cpu_ops():
    """Replace the CUDA entry points the layers call by float64 torch-CPU restatements (oracle arithmetic)."""
    from scipy import sparse as sp

    from deepsphere import _native as nat
    from deepsphere import _ops

    def graph_conv(x, kernel, bias, plan, recursion, K, act=nat.ACT_LINEAR, mode=nat.MODE_FP32):
        Lt = sp.csr_matrix((plan.values.astype(np.float64), (plan.indices[:, 0], plan.indices[:, 1])), shape=plan.shape)
        rec = "chebyshev" if recursion == nat.RECURSION_CHEBYSHEV else "monomial"
        y = orc.torch_cpu_graph_conv(x.double(), Lt, kernel.double(), K, rec)
        if bias is not None:
            y = y + bias.double().reshape(1, 1, -1)
        if act == nat.ACT_RELU:
            y = torch.relu(y)
        elif act != nat.ACT_LINEAR:
            raise NotImplementedError
        return y

    def pool(x, p, pool_type):
        B, M, F = x.shape
        v = x.reshape(B, M // 4**p, 4**p, F)
        return v.max(dim=2).values if pool_type == nat.POOL_MAX else v.mean(dim=2)

    _ops.graph_conv, _ops.pool = graph_conv, pool


def _layers(head):
    from deepsphere import healpy_layers as hl
    from deepsphere import keras_compat as kc

    return [hl.HealpyChebyshev(K=3, Fout=4, use_bias=True, activation="relu"), hl.HealpyPool(p=1, pool_type="MAX"),
            hl.HealpyChebyshev(K=4, Fout=2), head, kc.Dense(3)]


def _net_worker(rank, world, port, out_dir, masked):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dsd.init_from_env(backend="gloo")
    import deepsphere
    from deepsphere import healpix as hpx, utils
    from deepsphere import keras_compat as kc

    _install_cpu_ops()
    nside = 8
    if masked:  # partial sky, padded so that it survives the pooling level (advanced_tutorial.ipynb:137,211)
        idx = utils.extend_indices(hpx.query_disc(nside, [1, 0, 0], 1.3), nside, nside // 2)
    else:
        idx = np.arange(12 * nside * nside)
    npix = len(idx)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(3, npix, 2, generator=gen, dtype=torch.float64)
    t = torch.randn(3, 3, generator=gen, dtype=torch.float64)
    torch.manual_seed(0)  # the whole-sphere reference (and hence the copied weights) must be the same on every rank
    whole = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=_layers(kc.Lambda(lambda v: v.mean(dim=1))))
    part = partition.PartitionedHealpyGCNN(nside, idx, _layers(partition.PartitionedMean()))
    b, e = part.own_range
    assert (e - b) % 4 == 0 and e > b
    if not masked and world == 2:
        assert (b, e) == ((0, npix // 2) if rank == 0 else (npix // 2, npix))
    xw = x.clone().requires_grad_(True)
    xo = x[:, b:e].clone().requires_grad_(True)
    yw = whole(xw, training=True)
    part(xo.detach(), training=True)  # builds the weights
    pw, pp = list(whole.parameters()), list(part.parameters())
    assert len(pw) == len(pp) and all(a.shape == c.shape for a, c in zip(pw, pp))
    with torch.no_grad():
        for a, c in zip(pw, pp):
            c.copy_(a)
    dsd.broadcast_parameters(part)  # what a training script does: every rank holds rank 0's weights
    yp = part(xo, training=True)
    ((yw - t) ** 2).sum().backward()
    ((yp - t) ** 2).sum().backward()
    # graph-layer weights hold partial sums over own rows; the Dense head sees replicated activations on every rank:
    # PartitionedHealpyGCNN.allreduce_gradients() sums exactly the row-local ones
    graph_params = [p_ for m in part.layers_use if isinstance(m, partition.PartitionedGraphConv) for p_ in m.parameters()]
    assert [id(q) for q in part.row_local_parameters()] == [id(q) for q in graph_params]
    assert len(part.replicated_parameters()) == 2 and len(part.row_local_parameters()) + 2 == len(pp)
    part.allreduce_gradients()
    errs = [float((yp - yw).abs().max()), float((xo.grad - xw.grad[:, b:e]).abs().max())]
    errs += [float((c.grad.double() - a.grad.double()).abs().max()) for a, c in zip(pw, pp)]
    halo = [m.plan.halo_rows for m in part.layers_use if isinstance(m, partition.PartitionedGraphConv)]
    np.save(os.path.join(out_dir, f"net{rank}.npy"), np.array(errs + [float(min(halo))]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,masked", [(2, False), (3, True)])
def test_partitioned_healpy_gcnn_equals_single_process(tmp_path, world, masked):
    """Chebyshev -> MAX pool -> Chebyshev -> mean over the sphere -> Dense on 2 ranks (full sphere) and on 3 ranks
    (masked, padded sky): output, input gradient and every weight gradient equal the whole HealpyGCNN (the pooling
    level keeps 4^p siblings on one rank)."""
    mp.spawn(_net_worker, args=(world, _free_port(), str(tmp_path), masked), nprocs=world, join=True)
    for r in range(world):
        errs = np.load(tmp_path / f"net{r}.npy")
        assert errs[:-1].max() <= 1e-5, errs
        assert errs[-1] > 0  # every rank really has a halo


def test_native_halo_lists_reproduce_the_exchange_and_its_transpose():
    """The concatenated row lists and the row -> slot CSR that the C-ABI halo kernels take (ds_halo_pack / _assemble /
    _reduce, include/deepsphere_b200.h), evaluated here with numpy on every rank of a 3-way partition with the
    all-to-all simulated in-process: forward == the global tensor restricted to ext, backward == the transposed
    exchange (every rank's g_ext scattered to the global rows and summed, restricted to the own rows)."""
    from deepsphere.graph import SphereHealpix
    from deepsphere.partition import HaloPlan

    g = SphereHealpix(8, k=8)
    M, world, B, F = g.L.shape[0], 3, 2, 3
    plans = [HaloPlan(g.L, 3, r, world, align=M // 48) for r in range(world)]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, M, F))
    sends = []
    for p in plans:
        b0, e0 = p.own[p.rank]
        x_own = x[:, b0:e0]
        sends.append(np.transpose(x_own[:, p.send_cat], (1, 0, 2)))  # ds_halo_pack: [n, B, F], peer blocks contiguous
    offs = [np.concatenate(([0], np.cumsum([len(v) for v in p.send_rows]))) for p in plans]
    g_ext_all, backs = [], []
    for r, p in enumerate(plans):
        recv = np.concatenate([sends[q][offs[q][r]: offs[q][r + 1]] for q in range(world)])  # all_to_all
        assert len(recv) == len(p.recv_cat)
        b0, e0 = p.own[r]
        x_ext = np.full((B, p.n_ext, F), np.nan)
        x_ext[:, p.own_start: p.own_start + p.n_own] = x[:, b0:e0]          # ds_halo_assemble
        x_ext[:, p.recv_cat] = np.transpose(recv, (1, 0, 2))
        assert np.array_equal(x_ext, x[:, p.ext])
        ge = rng.standard_normal((B, p.n_ext, F))
        g_ext_all.append(ge)
        backs.append(np.transpose(ge[:, p.recv_cat], (1, 0, 2)))            # ds_halo_pack on g_ext with recv_cat
    g_global = np.zeros((B, M, F))
    for p, ge in zip(plans, g_ext_all):
        np.add.at(g_global, (slice(None), p.ext), ge)
    roffs = [np.concatenate(([0], np.cumsum([len(v) for v in p.recv_pos]))) for p in plans]
    for r, p in enumerate(plans):
        got = np.concatenate([backs[q][roffs[q][r]: roffs[q][r + 1]] for q in range(world)])  # transposed all_to_all
        assert len(got) == len(p.send_cat)
        b0, e0 = p.own[r]
        g_own = g_ext_all[r][:, p.own_start: p.own_start + p.n_own].copy()   # ds_halo_reduce
        for row in range(p.n_own):
            for s in p.reduce_slots[p.reduce_ptr[row]: p.reduce_ptr[row + 1]]:
                g_own[:, row] += got[s]
        assert np.allclose(g_own, g_global[:, b0:e0], rtol=0, atol=1e-12)


@pytest.mark.parametrize("world", [2, 8])
def test_lattice_payload_of_every_rank_of_a_partition(world):
    """The fused-kernel tables of a rank's EXTENDED row set (own rows + 4-hop halo): every tile is launched, the rows the
    lattice cannot reproduce (within reach of a valence-3 vertex; their neighbourhoods are cut by the partition into patches
    of different sizes) are listed exactly once in the patch tables ds_plan_attach_patches checks, and lattice rows +
    irregular rows cover the local graph."""
    from scipy import sparse

    from deepsphere import lattice, partition, utils
    from deepsphere.graph import SphereHealpix

    nside = 32
    M = 12 * nside**2
    L = sparse.csr_matrix(SphereHealpix(nside, k=8).L)
    Lt = sparse.csr_matrix(utils.rescale_L(L, lmax=1.9, scale=0.75).astype(np.float32))
    sizes = set()
    for rank in range(world):
        rows = np.asarray(partition.HaloPlan(L, 4, rank, world, align=M // 48).ext)
        pay = lattice.make_payload(sparse.csr_matrix(Lt[rows][:, rows]), nside, np.arange(M)[rows], 4)
        assert pay is not None and pay["H"] == 4
        wanted = pay["closure_rows"][pay["own_sub"]]
        assert np.array_equal(np.sort(np.concatenate([pay["lattice_rows"], wanted])), np.arange(len(rows)))
        pt = pay["patches"]
        if pt is None:
            assert len(wanted) == 0
            continue
        nr, no = np.diff(pt["row_ptr"]), np.diff(pt["own_ptr"])
        sizes.update(nr.tolist())
        assert pt["row_ptr"][-1] == len(pay["closure_rows"]) and pt["own_ptr"][-1] == len(wanted)
        assert np.array_equal(np.sort(pt["rows"]), pay["closure_rows"]) and nr.max() <= 1024 and no.max() <= 45
        got = np.concatenate([pt["rows"][pt["row_ptr"][p] + pt["own_local"][pt["own_ptr"][p]:pt["own_ptr"][p + 1]]]
                              for p in range(pt["n_patches"])])
        assert np.array_equal(np.sort(got), wanted)
        for p in range(pt["n_patches"]):  # columns stay inside their patch; the ELL is L~ restricted to it
            sl = slice(pt["row_ptr"][p], pt["row_ptr"][p + 1])
            assert pt["ell_col"][sl].max() < nr[p] and pt["ell_col"][sl].min() >= -1
            r = pt["rows"][sl]
            dense = np.zeros((nr[p], nr[p]), np.float32)
            rr, jj = np.nonzero(pt["ell_col"][sl] >= 0)
            dense[rr, pt["ell_col"][sl][rr, jj]] = pt["ell_val"][sl][rr, jj]
            assert np.array_equal(dense, Lt[rows[r]][:, rows[r]].toarray())
    assert sizes and max(sizes) <= 189
    if world == 2:
        assert len(sizes) > 1  # whole (189 rows) and cut neighbourhoods both occur
