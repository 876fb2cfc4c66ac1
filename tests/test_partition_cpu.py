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


def test_halo_plan_is_the_hop_closure():
    """The halo is derived from the sparsity of L: own + halo == everything within n_hops, and the send / receive
    lists of the two ranks mirror each other."""
    g = SphereHealpix(NSIDE, k=8)
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
