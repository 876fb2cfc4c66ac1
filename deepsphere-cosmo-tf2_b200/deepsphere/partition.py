"""Sphere partition with halo exchange: one sphere spread over the ranks of a process group.

The reference is single-device.  Here rank r owns a contiguous range of the rows of L (for a NESTED full sphere:
whole base-pixel blocks, so pooling / pseudo-convolutions with 4^p | block stay local) and, per graph-convolution
layer, needs the (K-1)-hop neighbourhood of its rows.  Schedule (i) of SURVEY 8e.2 is implemented: ONE exchange of
the (K-1)-ring halo per layer, then the unchanged single-GPU layer (fused lattice kernel included: the extended row
set is just a masked sky) on own + halo rows; results are exact on the own rows and the redundant halo rows are
dropped.  The halo set is derived from the sparsity of L by K-1 rounds of neighbour expansion on the host - never
from a ring count - so k = 20/40/60 graphs and masked skies work the same way.

Backward: the exchange is a torch.autograd.Function whose backward is the transposed exchange (gradients of halo
rows travel back to their owners and are added).  Weight gradients are partial sums over the own rows (the loss
only sees own rows, so dy is zero on the halo): sum them over the group (`distributed.allreduce_gradients(...,
average=False)`).

Collectives: all_to_all_single with per-peer row counts (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""

import numpy as np
import torch
import torch.distributed as dist
from scipy import sparse

from .distributed import shard_range


def _closure(pattern, rows, n_hops):
    """Rows within n_hops of `rows` in the graph of the (structurally symmetric) sparse `pattern`."""
    M = pattern.shape[0]
    mask = np.zeros(M, dtype=bool)
    mask[rows] = True
    frontier = mask.copy()
    for _ in range(int(n_hops)):
        reach = (pattern @ frontier.astype(np.float32)) > 0
        frontier = reach & ~mask
        if not frontier.any():
            break
        mask |= frontier
    return np.flatnonzero(mask)


class HaloPlan:
    """Who owns what and who sends which rows to whom, for `n_hops` hops over the sparsity of L.

    Every rank computes the full plan from the (replicated, host-side) Laplacian, so no negotiation is needed.
    own[r]     = [begin, end) of rank r (multiples of `align`)
    ext        = sorted global rows this rank computes on (own + halo)
    own_pos    = positions of the own rows inside ext
    send_rows[q] = LOCAL (own-relative) rows sent to rank q;  recv_pos[q] = positions in ext filled by rank q
    """

    def __init__(self, L, n_hops, rank, world, align=1):
        L = sparse.csr_matrix(L)
        M = L.shape[0]
        if M % align:
            raise ValueError(f"{M} rows are not a multiple of align={align}")
        pattern = sparse.csr_matrix((np.ones(L.nnz, dtype=np.float32), L.indices, L.indptr), shape=L.shape)
        pattern = pattern + pattern.T  # expansion must be symmetric: row i needs j  <=>  L[i, j] != 0
        units = M // align
        self.rank, self.world, self.M, self.n_hops = int(rank), int(world), M, int(n_hops)
        self.own = [tuple(align * v for v in shard_range(units, r, world)) for r in range(world)]
        owner_of = np.empty(M, dtype=np.int32)
        for r, (b, e) in enumerate(self.own):
            owner_of[b:e] = r
        exts = [_closure(pattern, np.arange(b, e), n_hops) for (b, e) in self.own]
        b0, e0 = self.own[rank]
        self.ext = exts[rank]
        self.own_pos = np.searchsorted(self.ext, np.arange(b0, e0))
        self.send_rows, self.recv_pos = [], []
        for q in range(world):
            if q == rank:
                self.send_rows.append(np.zeros(0, dtype=np.int64))
                self.recv_pos.append(np.zeros(0, dtype=np.int64))
                continue
            need_q = exts[q][owner_of[exts[q]] == rank]          # rows of mine in q's extended set
            self.send_rows.append((need_q - b0).astype(np.int64))
            mine_from_q = self.ext[owner_of[self.ext] == q]      # rows of q in my extended set
            self.recv_pos.append(np.searchsorted(self.ext, mine_from_q).astype(np.int64))
        self.n_own, self.n_ext = e0 - b0, len(self.ext)
        self.halo_rows = self.n_ext - self.n_own
        # the own rows are one interval of the sorted extended set: the exchange copies them as a slice
        self.own_start = int(self.own_pos[0]) if self.n_own else 0
        assert self.n_own == 0 or np.array_equal(self.own_pos, self.own_start + np.arange(self.n_own))
        self._dev = {}

    def on(self, device):
        """Index tensors of the exchange, resident on `device` (uploaded once)."""
        key = str(device)
        if key not in self._dev:
            self._dev[key] = ([torch.as_tensor(v, device=device) for v in self.send_rows],
                              [torch.as_tensor(v, device=device) for v in self.recv_pos])
        return self._dev[key]

    def restrict(self, L):
        """L restricted to the extended row set (rows and columns), CSR."""
        L = sparse.csr_matrix(L)
        return L[self.ext][:, self.ext].tocsr()


class _HaloExchange(torch.autograd.Function):
    """x_own [B, n_own, F] -> x_ext [B, n_ext, F]; backward = transposed exchange (halo gradients are returned to
    their owners and added)."""

    @staticmethod
    def forward(ctx, x_own, plan, group):
        ctx.plan, ctx.group = plan, group
        return _exchange(x_own, plan, group)

    @staticmethod
    def backward(ctx, g_ext):
        plan, group = ctx.plan, ctx.group
        g_ext = g_ext.contiguous()
        B, _, F = g_ext.shape
        send_idx, recv_idx = plan.on(g_ext.device)
        g_own = g_ext[:, plan.own_start: plan.own_start + plan.n_own, :].clone()
        # send back what I received (halo positions), receive what I sent (own rows) and accumulate
        send = [g_ext.index_select(1, recv_idx[q]).permute(1, 0, 2).reshape(-1) for q in range(plan.world)]
        counts_out = [len(plan.recv_pos[q]) * B * F for q in range(plan.world)]
        counts_in = [len(plan.send_rows[q]) * B * F for q in range(plan.world)]
        recv = _all_to_all(torch.cat(send) if send else g_ext.new_zeros(0), counts_out, counts_in, group)
        off = 0
        for q in range(plan.world):
            n = len(plan.send_rows[q])
            if n:
                blk = recv[off: off + n * B * F].reshape(n, B, F).permute(1, 0, 2)
                g_own.index_add_(1, send_idx[q], blk)
            off += n * B * F
        return g_own, None, None


def _all_to_all(flat, counts_out, counts_in, group):
    out = flat.new_empty(int(sum(counts_in)))
    if not dist.is_initialized():  # single process: nothing to exchange
        return out
    dist.all_to_all_single(out, flat.contiguous(), output_split_sizes=[int(c) for c in counts_in],
                           input_split_sizes=[int(c) for c in counts_out], group=group)
    return out


def _exchange(x_own, plan, group):
    x_own = x_own.contiguous()
    B, n_own, F = x_own.shape
    if n_own != plan.n_own:
        raise ValueError(f"rank {plan.rank} owns {plan.n_own} rows, got a tensor with {n_own}")
    send_idx, recv_idx = plan.on(x_own.device)
    x_ext = x_own.new_empty((B, plan.n_ext, F))
    x_ext[:, plan.own_start: plan.own_start + n_own, :] = x_own
    # rows go out row-major ([row, b, f]) so that a peer's block is contiguous
    send = [x_own.index_select(1, send_idx[q]).permute(1, 0, 2).reshape(-1) for q in range(plan.world)]
    counts_out = [len(plan.send_rows[q]) * B * F for q in range(plan.world)]
    counts_in = [len(plan.recv_pos[q]) * B * F for q in range(plan.world)]
    recv = _all_to_all(torch.cat(send), counts_out, counts_in, group)
    off = 0
    for q in range(plan.world):
        n = len(plan.recv_pos[q])
        if n:
            x_ext.index_copy_(1, recv_idx[q], recv[off: off + n * B * F].reshape(n, B, F).permute(1, 0, 2))
        off += n * B * F
    return x_ext


def halo_exchange(x_own, plan, group=None):
    """Differentiable gather of the halo rows: [B, n_own, F] -> [B, n_ext, F] (rows ordered like plan.ext)."""
    return _HaloExchange.apply(x_own, plan, group)


class PartitionedGraphConv(torch.nn.Module):
    """A graph convolution on a row-partitioned sphere: y_own = conv_ext(halo_exchange(x_own))[own rows].

    `make_layer(L_ext, ext_rows)` builds the single-device layer on the restricted Laplacian (e.g.
    `lambda L, rows: Chebyshev(L=L, K=5, Fout=64, lmax=lmax_global, healpix=(nside, pix[rows]))`); its weights must
    be identical on all ranks (broadcast them) and their gradients summed over the group after backward."""

    def __init__(self, L, n_hops, make_layer, rank=None, world=None, group=None, align=1):
        super().__init__()
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.plan = HaloPlan(L, n_hops, rank, world, align)
        self.layer = make_layer(self.plan.restrict(L), self.plan.ext)

    def forward(self, x_own, *args, **kwargs):
        x_ext = halo_exchange(x_own, self.plan, self.group)
        y_ext = self.layer(x_ext, *args, **kwargs)
        return y_ext[:, self.plan.own_start: self.plan.own_start + self.plan.n_own, :]


class PartitionedMean(torch.nn.Module):
    """Mean over ALL pixels of the partitioned sphere (the `reduce_mean(axis=1)` head of the reference's networks):
    local sum, one all-reduce of [B*F + 1] floats, replicated result [B, F].  The layers after it run replicated on
    every rank (same loss everywhere), so its backward hands each rank the gradient of ITS partial sum only."""

    def __init__(self, group=None):
        super().__init__()
        self.group = group

    def forward(self, x_own):
        from .distributed import allreduce_sum_replicated

        B, n, F = x_own.shape
        stats = torch.cat([x_own.sum(dim=1).reshape(-1), x_own.new_full((1,), float(n))])
        # everything after the head is replicated (every rank evaluates the same loss): identity backward
        stats = allreduce_sum_replicated(stats, self.group)
        return stats[: B * F].reshape(B, F) / stats[B * F]


class _Callable(torch.nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


class PartitionedHealpyGCNN(torch.nn.Module):
    """healpy_networks.HealpyGCNN on a row-partitioned sphere.

    Same layer list as HealpyGCNN.  Every rank owns a contiguous range of the (sorted, NESTED) input pixels, aligned
    so that all pooling / pseudo-convolution levels of the network keep their 4^p siblings on one rank: those layers
    run locally and unchanged.  Every HealpyChebyshev / HealpyMonomial becomes a PartitionedGraphConv on the graph of
    its level (global lmax, one (K-1)-ring halo exchange per layer).  Use PartitionedMean for the mean-over-pixels
    head; whatever follows it (Dense, ...) sees replicated tensors.  All ranks must construct the model with the same
    torch seed (or broadcast the parameters), feed the SAME batch restricted to their rows, and sum the weight
    gradients over the group after backward (`allreduce_gradients(params, average=False)`).
    Not supported: use_bn inside graph layers (statistics would need the own-row mask), residual layers."""

    def __init__(self, nside, indices, layers, n_neighbors=8, rank=None, world=None, group=None):
        super().__init__()
        import copy

        from . import healpix as hpx
        from . import utils
        from . import healpy_layers as hp_nn
        from .graph import SphereHealpix

        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        idx = np.sort(np.asarray(indices, dtype=np.int64))
        M = len(idx)
        # deepest pooling level reached anywhere in the network -> alignment of the row ranges
        depth, max_depth = 0, 0
        for layer in layers:
            if isinstance(layer, (hp_nn.HealpyPool, hp_nn.HealpyPseudoConv)):
                depth += int(layer.p)
            elif isinstance(layer, hp_nn.HealpyPseudoConv_Transpose):
                depth -= int(layer.p)
            max_depth = max(max_depth, depth)
        align = 4 ** max_depth
        if M % align:
            raise ValueError(f"{M} pixels cannot be pooled {max_depth} times: use utils.extend_indices first")
        if M % 48 == 0 and (M // 48) % align == 0 and world <= 48:
            align = M // 48  # quarter-face blocks of a full sphere
        if M // align < world:
            raise ValueError(f"{M // align} partition units for {world} ranks")
        self.rank, self.world, self.group = rank, world, group
        b, e = shard_range(M // align, rank, world)
        self.own_range = (align * b, align * e)
        self.layers_use = torch.nn.ModuleList()
        cur_nside, cur_idx, cur_align = int(nside), idx, align
        for layer in layers:
            if isinstance(layer, (hp_nn.HealpyChebyshev, hp_nn.HealpyMonomial, hp_nn.HealpyBernstein)):
                if layer.use_bn:
                    raise NotImplementedError("use_bn inside a partitioned graph layer")
                sphere = SphereHealpix(subdivisions=cur_nside, indexes=cur_idx, nest=True, k=n_neighbors,
                                       lap_type="normalized")
                L = sphere.L
                lmax = 1.02 * utils.largest_eigenvalue(sparse.csr_matrix(L, dtype=np.float64))

                def make(L_ext, rows, layer=layer, lmax=lmax, nside_l=cur_nside, idx_l=cur_idx):
                    f = copy.copy(layer)
                    f.kwargs = dict(layer.kwargs, lmax=lmax, healpix=(nside_l, idx_l[rows]))
                    return f._get_layer(L_ext)

                # hops = polynomial degree: K - 1 for K-term Chebyshev / Monomial, K for an order-K Bernstein layer
                hops = int(layer.K) if isinstance(layer, hp_nn.HealpyBernstein) else int(layer.K) - 1
                self.layers_use.append(PartitionedGraphConv(L, hops, make, rank, world, group, cur_align))
            elif isinstance(layer, (hp_nn.HealpyPool, hp_nn.HealpyPseudoConv)):
                cur_idx = hpx.coarsen_indices(cur_idx, int(layer.p))
                cur_nside //= 2 ** int(layer.p)
                cur_align //= 4 ** int(layer.p)
                self.layers_use.append(layer)
            elif isinstance(layer, hp_nn.HealpyPseudoConv_Transpose):
                cur_idx = hpx.refine_indices(cur_idx, int(layer.p))
                cur_nside *= 2 ** int(layer.p)
                cur_align *= 4 ** int(layer.p)
                self.layers_use.append(layer)
            elif isinstance(layer, hp_nn.Healpy_ResidualLayer):
                raise NotImplementedError("residual layers on a partitioned sphere")
            else:
                self.layers_use.append(layer if isinstance(layer, torch.nn.Module) else _Callable(layer))

    def forward(self, x_own, training=False):
        from .keras_compat import _accepts_training

        h = x_own
        for layer in self.layers_use:
            inner = layer.layer if isinstance(layer, PartitionedGraphConv) else layer
            h = layer(h, training=training) if _accepts_training(inner) else layer(h)
        return h
