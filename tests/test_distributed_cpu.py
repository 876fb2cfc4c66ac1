"""world_size-2 gloo test of the batch-sharding exchange step (the flat gradient all-reduce):
N-rank result == 1-rank result.  CPU only."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepsphere import distributed as dsd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = dsd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    extra = torch.nn.Parameter(torch.ones(4))  # never receives a gradient on rank 1
    if rank == 1:
        with torch.no_grad():
            lin.weight.add_(1.0)  # ranks start different; broadcast must fix that
    dsd.broadcast_parameters(lin, src=0)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    b, e = dsd.shard_range(8, rank, world)
    loss = ((lin(x[b:e]) - y[b:e]) ** 2).sum() / 8
    if rank == 0:
        loss = loss + extra.sum() * 0.0
    loss.backward()
    n = dsd.allreduce_gradients(list(lin.parameters()) + [extra], average=False)
    assert n == 6 * 3 + 3 + 4
    assert abs(dsd.allreduce_max(float(rank), torch.device("cpu")) - (world - 1)) < 1e-12
    np.save(os.path.join(out_dir, f"g{rank}.npy"), lin.weight.grad.numpy())
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_rank(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 6, generator=g)
    y = torch.randn(8, 3, generator=g)
    (((lin(x) - y) ** 2).sum() / 8).backward()
    g0 = np.load(tmp_path / "g0.npy")
    g1 = np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1)
    assert np.allclose(g0, lin.weight.grad.numpy(), atol=1e-6)


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [dsd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _bn_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dsd.init_from_env(backend="gloo")
    from deepsphere.keras_compat import BatchNormalization

    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, 10, 4, generator=g, dtype=torch.float64)
    tgt = torch.randn(6, 10, 4, generator=g, dtype=torch.float64)
    b, e = dsd.shard_range(6, rank, world)
    xs = x[b:e].clone().requires_grad_(True)
    bn = BatchNormalization(axis=-1, momentum=0.9, epsilon=1e-5, center=False, scale=False)  # gnn_layers.py:53
    y = bn(xs, training=True)
    ((y - tgt[b:e]) ** 2).sum().backward()
    np.save(os.path.join(out_dir, f"y{rank}.npy"), y.detach().numpy())
    np.save(os.path.join(out_dir, f"dx{rank}.npy"), xs.grad.numpy())
    np.save(os.path.join(out_dir, f"mm{rank}.npy"), bn.moving_mean.numpy())
    dist.destroy_process_group()


def test_two_rank_batchnorm_uses_global_batch_statistics(tmp_path):
    """with the batch sharded, BatchNormalization output, input gradient and moving statistics equal
    the single-process result on the whole batch."""
    from deepsphere.keras_compat import BatchNormalization

    port = _free_port()
    mp.spawn(_bn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, 10, 4, generator=g, dtype=torch.float64).requires_grad_(True)
    tgt = torch.randn(6, 10, 4, generator=g, dtype=torch.float64)
    bn = BatchNormalization(axis=-1, momentum=0.9, epsilon=1e-5, center=False, scale=False)
    y = bn(x, training=True)
    ((y - tgt) ** 2).sum().backward()
    y2 = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    dx2 = np.concatenate([np.load(tmp_path / "dx0.npy"), np.load(tmp_path / "dx1.npy")])
    assert np.allclose(y2, y.detach().numpy(), atol=1e-10)
    assert np.allclose(dx2, x.grad.numpy(), rtol=1e-5, atol=1e-5)
    assert np.allclose(np.load(tmp_path / "mm0.npy"), bn.moving_mean.numpy(), atol=1e-6)
    assert np.allclose(np.load(tmp_path / "mm1.npy"), bn.moving_mean.numpy(), atol=1e-6)
