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
