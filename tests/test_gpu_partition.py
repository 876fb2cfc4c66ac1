"""Sphere partition with halo exchange on 2 GPUs (NCCL): the CUDA Chebyshev layer on each rank's own + halo rows
reproduces the single-GPU layer on the whole sphere - forward, input gradient, weight gradient.  Needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NSIDE, K, F_IN, F_OUT, B = 32, 5, 16, 16, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, tol, out_dir):
    import torch.distributed as dist
    from scipy.sparse.linalg import eigsh

    from deepsphere import distributed as dsd
    from deepsphere import gnn_layers, partition
    from deepsphere.graph import SphereHealpix

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dsd.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    g = SphereHealpix(NSIDE, k=8)
    M = g.L.shape[0]
    lmax = 1.02 * eigsh(g.L.astype(np.float64), k=1, which="LM", return_eigenvectors=False)[0]
    pix = np.arange(M)

    def make(L_ext, rows):
        torch.manual_seed(0)
        layer = gnn_layers.Chebyshev(L=L_ext, K=K, Fout=F_OUT, lmax=lmax, healpix=(NSIDE, pix[rows]), mode=mode)
        layer.build_from_shape((B, len(rows), F_IN))
        return layer

    conv = partition.PartitionedGraphConv(g.L, K - 1, make, align=256)
    dsd.broadcast_parameters(conv.layer)
    torch.manual_seed(0)
    whole = gnn_layers.Chebyshev(L=g.L, K=K, Fout=F_OUT, lmax=lmax, mode=mode)
    whole.build_from_shape((B, M, F_IN))
    with torch.no_grad():
        whole.kernel.copy_(conv.layer.kernel)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, M, F_IN, generator=gen).to(dev)
    dy = torch.randn(B, M, F_OUT, generator=gen).to(dev)
    b, e = conv.plan.own[rank]
    x_own = x[:, b:e].clone().requires_grad_(True)
    y_own = conv(x_own)
    y_own.backward(dy[:, b:e].contiguous())
    dsd.allreduce_gradients([conv.layer.kernel], average=False)
    xg = x.clone().requires_grad_(True)
    y = whole(xg)
    y.backward(dy)
    torch.cuda.synchronize()

    def rel(a, ref):
        return float((a - ref).abs().max() / ref.abs().max())

    errs = (rel(y_own.detach(), y.detach()[:, b:e]), rel(x_own.grad, xg.grad[:, b:e]),
            rel(conv.layer.kernel.grad, whole.kernel.grad))
    np.save(os.path.join(out_dir, f"err{rank}.npy"), np.array(errs + (conv.plan.n_own, conv.plan.n_ext,
                                                                    conv.layer._plan.info(rank)["lattice"])))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("tf32", 2e-3)])
def test_partitioned_chebyshev_two_gpus(tmp_path, mode, tol):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), mode, tol, str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        ey, edx, edw, n_own, n_ext, lattice = np.load(tmp_path / f"err{r}.npy")
        assert n_ext > n_own
        assert ey <= tol and edx <= tol and edw <= tol, (mode, r, ey, edx, edw)


def _net_worker(rank, world, port, out_dir, use_bn=False):
    import torch.distributed as dist

    import deepsphere
    from deepsphere import distributed as dsd
    from deepsphere import healpy_layers as hl
    from deepsphere import keras_compat as kc
    from deepsphere import partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dsd.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    nside, npix = 32, 12 * 32 * 32

    def layers(head):
        return [hl.HealpyPseudoConv(p=1, Fout=8, activation="elu"),
                hl.HealpyChebyshev(K=5, Fout=16, use_bias=True, use_bn=use_bn, activation="elu", mode="tf32"),
                hl.HealpyPool(p=1, pool_type="AVG"),
                hl.HealpyChebyshev(K=3, Fout=8, use_bias=use_bn, use_bn=use_bn, activation="tanh" if use_bn else None), head,
                kc.Dense(2)]

    torch.manual_seed(0)
    whole = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(npix), layers=layers(kc.Lambda(lambda v: v.mean(dim=1))))
    part = partition.PartitionedHealpyGCNN(nside, np.arange(npix), layers(partition.PartitionedMean()))
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, npix, 1, generator=gen).to(dev)
    t = torch.randn(2, 2, generator=gen).to(dev)
    b, e = part.own_range
    xw = x.clone().requires_grad_(True)
    xo = x[:, b:e].clone().requires_grad_(True)
    yw = whole(xw, training=True)
    part(xo.detach(), training=True)
    pw, pp = list(whole.parameters()), list(part.parameters())
    with torch.no_grad():
        for a, c in zip(pw, pp):
            c.copy_(a)
    dsd.broadcast_parameters(part)
    yp = part(xo, training=True)
    ((yw - t) ** 2).sum().backward()
    ((yp - t) ** 2).sum().backward()
    graph_params = [p_ for m in part.layers_use if isinstance(m, partition.PartitionedGraphConv) for p_ in m.parameters()]
    # the pseudo-convolution is row-local: its weight gradient is a partial sum over own rows too
    local_params = [p_ for m in part.layers_use if isinstance(m, hl.HealpyPseudoConv) for p_ in m.parameters()]
    dsd.allreduce_gradients(graph_params + local_params, average=False)
    torch.cuda.synchronize()

    def rel(a, ref):
        return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))

    errs = [rel(yp.detach(), yw.detach()), rel(xo.grad, xw.grad[:, b:e])] + [rel(c.grad, a.grad) for a, c in zip(pw, pp)]
    np.save(os.path.join(out_dir, f"net{rank}.npy"), np.array(errs))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_healpy_gcnn_two_gpus(tmp_path):
    """PseudoConv -> Chebyshev (fused tf32 kernel) -> AVG pool -> Chebyshev -> mean over the sphere -> Dense on a sphere
    split over 2 GPUs equals the whole-sphere HealpyGCNN: output, input gradient, every weight gradient."""
    import torch.multiprocessing as mp

    mp.spawn(_net_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        errs = np.load(tmp_path / f"net{r}.npy")
        assert errs.max() <= 5e-3, errs


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_healpy_gcnn_two_gpus_batchnorm(tmp_path):
    """The same with use_bn=True in the graph layers (round 2): the BatchNormalization statistics of a partitioned layer are
    those of the whole sphere (own rows of every rank, ds_bn_* row range + one all-reduce of the 2F + 1 sums per direction),
    so output, input gradient and every weight gradient still equal the whole-sphere network."""
    import torch.multiprocessing as mp

    mp.spawn(_net_worker, args=(2, _free_port(), str(tmp_path), True), nprocs=2, join=True)
    for r in (0, 1):
        errs = np.load(tmp_path / f"net{r}.npy")
        assert errs.max() <= 5e-3, errs


@pytest.mark.parametrize("F", [1, 5, 16])
def test_halo_kernels_single_gpu_simulated_ranks(F):
    """ds_halo_pack / ds_halo_assemble / ds_halo_reduce (include/deepsphere_b200.h) on ONE GPU: the 3 ranks of a
    partition are simulated in-process (the all-to-all is a slice copy), so the C-ABI kernels are checked on every box:
    forward == the global tensor restricted to ext (bit-exact: pure data movement), backward == the transposed exchange."""
    from deepsphere import _native as nat
    from deepsphere.graph import SphereHealpix
    from deepsphere.partition import HaloPlan

    g = SphereHealpix(16, k=8)
    M, world, B = g.L.shape[0], 3, 2
    plans = [HaloPlan(g.L, 4, r, world, align=M // 48) for r in range(world)]
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, M, F, device=dev, generator=gen)
    st = nat.current_stream()
    lib = nat.lib()
    sends = []
    for p in plans:
        b0, e0 = p.own[p.rank]
        x_own = x[:, b0:e0].contiguous()
        send_cat = p.native(dev)[0]
        buf = torch.empty(len(p.send_cat), B, F, device=dev)
        nat.check(lib.ds_halo_pack(B, e0 - b0, F, len(p.send_cat), nat.ptr(send_cat), nat.ptr(x_own), nat.ptr(buf), st))
        sends.append(buf)
    offs = [np.concatenate(([0], np.cumsum([len(v) for v in p.send_rows]))) for p in plans]
    roffs = [np.concatenate(([0], np.cumsum([len(v) for v in p.recv_pos]))) for p in plans]
    g_exts, backs = [], []
    for r, p in enumerate(plans):
        recv = torch.cat([sends[q][offs[q][r]: offs[q][r + 1]] for q in range(world)]).contiguous()
        b0, e0 = p.own[r]
        x_own = x[:, b0:e0].contiguous()
        x_ext = torch.full((B, p.n_ext, F), float("nan"), device=dev)
        nat.check(lib.ds_halo_assemble(B, p.n_own, p.n_ext, p.own_start, F, len(p.recv_cat), nat.ptr(p.native(dev)[1]),
                                       nat.ptr(x_own), nat.ptr(recv), nat.ptr(x_ext), st))
        assert torch.equal(x_ext, x[:, torch.as_tensor(p.ext, device=dev)])
        ge = torch.randn(B, p.n_ext, F, device=dev, generator=gen)
        g_exts.append(ge)
        back = torch.empty(len(p.recv_cat), B, F, device=dev)
        nat.check(lib.ds_halo_pack(B, p.n_ext, F, len(p.recv_cat), nat.ptr(p.native(dev)[1]), nat.ptr(ge), nat.ptr(back), st))
        backs.append(back)
    g_global = torch.zeros(B, M, F, device=dev, dtype=torch.float64)
    for p, ge in zip(plans, g_exts):
        g_global.index_add_(1, torch.as_tensor(p.ext, device=dev), ge.double())
    for r, p in enumerate(plans):
        got = torch.cat([backs[q][roffs[q][r]: roffs[q][r + 1]] for q in range(world)]).contiguous()
        b0, e0 = p.own[r]
        g_own = torch.empty(B, p.n_own, F, device=dev)
        _, _, red_ptr, red_slots = p.native(dev)
        nat.check(lib.ds_halo_reduce(B, p.n_own, p.n_ext, p.own_start, F, nat.ptr(red_ptr), nat.ptr(red_slots),
                                     nat.ptr(g_exts[r]), nat.ptr(got), nat.ptr(g_own), st))
        torch.cuda.synchronize()
        assert float((g_own.double() - g_global[:, b0:e0]).abs().max()) <= 1e-5 * float(g_global.abs().max())


def _comm_worker(rank, world, id_path, out_dir):
    """ds_comm_* + ds_halo_exchange(_backward) through ctypes alone (no torch.distributed): what a TF binder would call."""
    import ctypes
    import glob
    import time

    from deepsphere import _native as nat
    from deepsphere.graph import SphereHealpix
    from deepsphere.partition import HaloPlan

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    lib = nat.lib()
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so*"))
    nat.check(lib.ds_comm_load_nccl(cands[0].encode() if cands else None), "ds_comm_load_nccl")
    idbuf = ctypes.create_string_buffer(128)
    if rank == 0:
        nat.check(lib.ds_comm_unique_id(idbuf), "ds_comm_unique_id")
        with open(id_path + ".tmp", "wb") as f:
            f.write(idbuf.raw)
        os.replace(id_path + ".tmp", id_path)
    else:
        t0 = time.time()
        while not os.path.exists(id_path):
            assert time.time() - t0 < 120
            time.sleep(0.05)
        idbuf = ctypes.create_string_buffer(open(id_path, "rb").read(), 128)
    comm = ctypes.c_void_p()
    nat.check(lib.ds_comm_create(world, rank, idbuf, ctypes.byref(comm)), "ds_comm_create")
    g = SphereHealpix(16, k=8)
    M, B, F = g.L.shape[0], 2, 8
    plans = [HaloPlan(g.L, 4, r, world, align=M // 48) for r in range(world)]
    p = plans[rank]
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, M, F, generator=gen)
    g_exts = [torch.randn(B, q.n_ext, F, generator=torch.Generator().manual_seed(10 + r)) for r, q in enumerate(plans)]
    b0, e0 = p.own[rank]
    st = nat.current_stream()
    send_cat, recv_cat, red_ptr, red_slots = p.native(dev)
    spp = (ctypes.c_int64 * world)(*[len(v) for v in p.send_rows])
    rpp = (ctypes.c_int64 * world)(*[len(v) for v in p.recv_pos])
    ws = torch.empty((len(p.send_cat) + len(p.recv_cat)) * B * F + 4, device=dev)
    x_own = x[:, b0:e0].contiguous().to(dev)
    x_ext = torch.full((B, p.n_ext, F), float("nan"), device=dev)
    nat.check(lib.ds_halo_exchange(comm, B, p.n_own, p.n_ext, p.own_start, F, nat.ptr(send_cat), spp, nat.ptr(recv_cat), rpp,
                                   nat.ptr(x_own), nat.ptr(x_ext), nat.ptr(ws), st), "ds_halo_exchange")
    torch.cuda.synchronize()
    err_f = float((x_ext.cpu() - x[:, torch.as_tensor(p.ext)]).abs().max())
    g_ext = g_exts[rank].to(dev)
    g_own = torch.empty(B, p.n_own, F, device=dev)
    nat.check(lib.ds_halo_exchange_backward(comm, B, p.n_own, p.n_ext, p.own_start, F, spp, nat.ptr(recv_cat), rpp,
                                            nat.ptr(red_ptr), nat.ptr(red_slots), nat.ptr(g_ext), nat.ptr(g_own), nat.ptr(ws), st),
              "ds_halo_exchange_backward")
    g_global = torch.zeros(B, M, F, dtype=torch.float64)
    for q, ge in zip(plans, g_exts):
        g_global.index_add_(1, torch.as_tensor(q.ext), ge.double())
    torch.cuda.synchronize()
    err_b = float((g_own.cpu().double() - g_global[:, b0:e0]).abs().max())
    v = torch.full((5,), float(rank + 1), device=dev, dtype=torch.float64)
    nat.check(lib.ds_comm_allreduce_sum(comm, nat.ptr(v), 5, 1, st), "ds_comm_allreduce_sum")
    torch.cuda.synchronize()
    err_a = float((v.cpu() - sum(range(1, world + 1))).abs().max())
    lib.ds_comm_destroy(comm)
    np.save(os.path.join(out_dir, f"comm{rank}.npy"), np.array([err_f, err_b, err_a]))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ds_comm_and_halo_exchange_through_the_c_abi_alone(tmp_path):
    """ds_comm_* (NCCL loaded at run time) + ds_halo_exchange / _backward: the halo exchange of a partitioned layer and the
    all-reduce as ONE C-ABI call each, driven through ctypes without torch.distributed — forward bit-exact, transposed
    exchange to fp32 round-off."""
    import torch.multiprocessing as mp

    mp.spawn(_comm_worker, args=(2, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=2, join=True)
    for r in (0, 1):
        errs = np.load(tmp_path / f"comm{r}.npy")
        assert errs[0] == 0.0 and errs[1] <= 1e-5 and errs[2] == 0.0, errs
