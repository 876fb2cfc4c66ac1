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


class _Timing:
    """Optional device-side timing of the exchanges (bench.py: time split of a partitioned training step).  When
    `enabled`, every halo exchange (forward and transposed) is bracketed by CUDA events on the current stream."""

    enabled = False
    events = []

    @classmethod
    def start(cls):
        cls.enabled, cls.events = True, []

    @classmethod
    def stop(cls):
        """Total milliseconds spent between the recorded event pairs (call after a device synchronisation)."""
        cls.enabled = False
        ms = sum(a.elapsed_time(b) for a, b in cls.events)
        n = len(cls.events)
        cls.events = []
        return ms, n


class _timed:
    def __enter__(self):
        if _Timing.enabled and torch.cuda.is_available():
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        else:
            self.a = None

    def __exit__(self, *exc):
        if self.a is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _Timing.events.append((self.a, b))


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
        # the C-ABI kernels (ds_halo_pack / _assemble / _reduce) take the row lists of ALL peers concatenated, so that
        # one launch serves every peer: the send buffer is [sum_q n_q, B, F] with peer q's block contiguous
        self.send_cat = np.concatenate(self.send_rows).astype(np.int32)
        self.recv_cat = np.concatenate(self.recv_pos).astype(np.int32)
        # transposed exchange: own row r receives the slots of the returned buffer listed in slots[ptr[r]:ptr[r+1]]
        # (ascending slot = ascending peer: a fixed summation order)
        order = np.argsort(self.send_cat, kind="stable").astype(np.int32)
        counts = np.bincount(self.send_cat, minlength=self.n_own) if len(self.send_cat) else np.zeros(self.n_own, np.int64)
        self.reduce_ptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int32)
        self.reduce_slots = order

    def on(self, device):
        """Index tensors of the exchange, resident on `device` (uploaded once)."""
        key = str(device)
        if key not in self._dev:
            self._dev[key] = ([torch.as_tensor(v, device=device) for v in self.send_rows],
                              [torch.as_tensor(v, device=device) for v in self.recv_pos])
        return self._dev[key]

    def native(self, device):
        """int32 device tensors of the C-ABI halo kernels: (send_cat, recv_cat, reduce_ptr, reduce_slots)."""
        key = "native:" + str(device)
        if key not in self._dev:
            self._dev[key] = tuple(torch.as_tensor(v, dtype=torch.int32, device=device)
                                   for v in (self.send_cat, self.recv_cat, self.reduce_ptr, self.reduce_slots))
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
        with _timed():
            return _HaloExchange._backward(g_ext, plan, group)

    @staticmethod
    def _backward(g_ext, plan, group):
        g_ext = g_ext.contiguous()
        B, _, F = g_ext.shape
        if g_ext.is_cuda:  # C-ABI kernels: one pack launch, the all-to-all, one reduce launch
            return _native_exchange_backward(g_ext, plan, group), None, None
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
    with _timed():
        return _exchange_impl(x_own, plan, group)


def _native_exchange(x_own, plan, group):
    """The exchange on CUDA tensors through libdeepsphere_b200.so: ds_halo_pack -> all_to_all -> ds_halo_assemble."""
    from . import _native as nat

    B, n_own, F = x_own.shape
    send_cat, recv_cat, _, _ = plan.native(x_own.device)
    n_send, n_recv = len(plan.send_cat), len(plan.recv_cat)
    st = nat.current_stream()
    send = x_own.new_empty(n_send * B * F)
    nat.check(nat.lib().ds_halo_pack(B, n_own, F, n_send, nat.ptr(send_cat), nat.ptr(x_own), nat.ptr(send), st), "ds_halo_pack")
    counts_out = [len(plan.send_rows[q]) * B * F for q in range(plan.world)]
    counts_in = [len(plan.recv_pos[q]) * B * F for q in range(plan.world)]
    recv = _all_to_all(send, counts_out, counts_in, group)
    x_ext = x_own.new_empty((B, plan.n_ext, F))
    nat.check(nat.lib().ds_halo_assemble(B, n_own, plan.n_ext, plan.own_start, F, n_recv, nat.ptr(recv_cat), nat.ptr(x_own),
                                         nat.ptr(recv), nat.ptr(x_ext), st), "ds_halo_assemble")
    return x_ext


def _native_exchange_backward(g_ext, plan, group):
    from . import _native as nat

    B, n_ext, F = g_ext.shape
    _, recv_cat, red_ptr, red_slots = plan.native(g_ext.device)
    n_send, n_recv = len(plan.send_cat), len(plan.recv_cat)
    st = nat.current_stream()
    back = g_ext.new_empty(n_recv * B * F)  # what I received in the forward goes back to its owners
    nat.check(nat.lib().ds_halo_pack(B, n_ext, F, n_recv, nat.ptr(recv_cat), nat.ptr(g_ext), nat.ptr(back), st), "ds_halo_pack")
    counts_out = [len(plan.recv_pos[q]) * B * F for q in range(plan.world)]
    counts_in = [len(plan.send_rows[q]) * B * F for q in range(plan.world)]
    got = _all_to_all(back, counts_out, counts_in, group)
    g_own = g_ext.new_empty((B, plan.n_own, F))
    nat.check(nat.lib().ds_halo_reduce(B, plan.n_own, n_ext, plan.own_start, F, nat.ptr(red_ptr) if n_send else None,
                                       nat.ptr(red_slots) if n_send else None, nat.ptr(g_ext), nat.ptr(got),
                                       nat.ptr(g_own), st), "ds_halo_reduce")
    return g_own


def _exchange_impl(x_own, plan, group):
    x_own = x_own.contiguous()
    B, n_own, F = x_own.shape
    if n_own != plan.n_own:
        raise ValueError(f"rank {plan.rank} owns {plan.n_own} rows, got a tensor with {n_own}")
    if x_own.is_cuda:
        return _native_exchange(x_own, plan, group)
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
        if getattr(self.layer, "use_bn", False):
            # BatchNormalization inside the layer (gnn_layers.py:53,152-153): statistics over the OWN rows of every rank,
            # summed over the group = the statistics of the whole sphere; halo rows carry no gradient (ds_bn_* row range)
            self.layer._bn_rows = (self.plan.own_start, self.plan.own_start + self.plan.n_own)
            self.layer._bn_sync = True if group is None else group

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
    use_bn inside graph layers: the statistics are those of the whole sphere (own rows of every rank, one all-reduce of
    2F + 1 doubles per direction).  Not supported: residual layers."""

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
        # the index set must be closed under the 4^depth sibling groups (the check of HealpyGCNN, healpy_networks.py:
        # 66-86): a compatible COUNT with incomplete groups would let local pooling mix unrelated pixels
        if max_depth > 0 and not np.array_equal(hpx.refine_indices(hpx.coarsen_indices(idx, max_depth), max_depth), idx):
            raise ValueError(
                "With the given indices it would not be possible to properly reduce the input maps "
                "with the reduction factor determined by the layers. Use the function "
                "<extend_indices> from utils with the determined minimal nside to make your set of "
                "indices compatible...")
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

    def row_local_parameters(self):
        """Weights of the layers that run on a rank's OWN rows (everything before the PartitionedMean head): their
        gradients are partial sums over the rows and have to be SUMMED over the group."""
        out = []
        for layer in self.layers_use:
            if isinstance(layer, PartitionedMean):
                break
            out += [p for p in layer.parameters() if p.requires_grad]
        return out

    def replicated_parameters(self):
        """Weights of the layers behind the PartitionedMean head: they see replicated tensors, every rank computes the
        same gradient (average them, or leave them alone)."""
        out, behind = [], False
        for layer in self.layers_use:
            if behind:
                out += [p for p in layer.parameters() if p.requires_grad]
            behind = behind or isinstance(layer, PartitionedMean)
        return out

    def allreduce_gradients(self):
        """The gradient exchange of one training step: sum of the row-local partial sums over the group (one flat
        all-reduce); the replicated head is identical on every rank by construction and needs no exchange."""
        from .distributed import allreduce_gradients

        return allreduce_gradients(self.row_local_parameters(), group=self.group, average=False)

    def forward(self, x_own, training=False):
        from .keras_compat import _accepts_training

        h = x_own
        for layer in self.layers_use:
            inner = layer.layer if isinstance(layer, PartitionedGraphConv) else layer
            h = layer(h, training=training) if _accepts_training(inner) else layer(h)
        return h
