"""torch.autograd glue over the C-ABI: the only place the host framework touches the
hot path.  Forward and backward of every op are single C-ABI calls on the current stream.
"""

import os

import torch

from . import _native as nat

# Keep T_1..T_{K-1} from the forward for the backward (what TF autodiff does)
# while the basis stays below this many bytes; beyond it the backward recomputes them.
_SAVE_BASIS_MAX_BYTES = int(os.environ.get("DEEPSPHERE_SAVE_BASIS_MAX_BYTES", 48 * 2**30))


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _need_cuda(t, who):
    if not t.is_cuda:
        raise nat.NativeError(
            f"{who}: input lives on '{t.device}'. The deepsphere B200 hot path runs on CUDA only (no CPU fallback)."
        )


class _GraphConv(torch.autograd.Function):
    """y = act( sum_{f,k} T_k(L~) x [.,.,f] * kernel[f*K+k, :] + bias )  (gnn_layers.py:130-159)."""

    @staticmethod
    def forward(ctx, x, kernel, bias, plan, recursion, K, act, mode):
        _need_cuda(x, "graph convolution")
        x, kernel = _f32c(x), _f32c(kernel)
        bias_c = None if bias is None else _f32c(bias).reshape(-1)
        B, M, Fin = x.shape
        Fout = kernel.shape[1]
        dev = x.device.index
        h = plan.handle(dev)
        y = torch.empty((B, M, Fout), device=x.device, dtype=torch.float32)
        # the fused lattice kernel keeps the basis on chip: nothing to allocate, and the backward gets basis = NULL
        writes = nat.lib().ds_graph_conv_forward_writes_basis(h, K, B, Fin, Fout, mode)
        n_basis = nat.lib().ds_graph_conv_basis_elems(M, B, Fin, K) if writes else 0
        basis = torch.empty(n_basis, device=x.device, dtype=torch.float32) if n_basis > 0 else None
        with torch.cuda.device(dev):
            nat.check(
                nat.lib().ds_graph_conv_forward(
                    h, recursion, K, B, Fin, Fout, nat.ptr(x), nat.ptr(kernel), nat.ptr(bias_c), act, nat.ptr(y),
                    nat.ptr(basis), mode, nat.current_stream(),
                ),
                "ds_graph_conv_forward",
            )
        needs_grad = any(ctx.needs_input_grad[:3])
        keep_basis = needs_grad and n_basis > 0 and n_basis * 4 <= _SAVE_BASIS_MAX_BYTES
        ctx.save_for_backward(x, kernel, y if act != nat.ACT_LINEAR else None, basis if keep_basis else None)
        ctx.meta = (plan, recursion, K, act, mode, bias is not None, None if bias is None else bias.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, kernel, y, basis = ctx.saved_tensors
        plan, recursion, K, act, mode, has_bias, bias_shape = ctx.meta
        dy = _f32c(dy)
        B, M, Fin = x.shape
        Fout = kernel.shape[1]
        dev = x.device.index
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_like(x) if need_dx else None
        dk = torch.empty_like(kernel)
        db = torch.empty(Fout, device=x.device, dtype=torch.float32) if has_bias else None
        n_ws = nat.lib().ds_graph_conv_backward_workspace_elems(M, B, Fin, Fout, K, 1 if basis is not None else 0, act)
        ws = torch.empty(n_ws, device=x.device, dtype=torch.float32)
        with torch.cuda.device(dev):
            nat.check(
                nat.lib().ds_graph_conv_backward(
                    plan.handle(dev), recursion, K, B, Fin, Fout, nat.ptr(x), nat.ptr(kernel), nat.ptr(y),
                    nat.ptr(dy), act, nat.ptr(basis), nat.ptr(dx), nat.ptr(dk), nat.ptr(db), nat.ptr(ws), mode,
                    nat.current_stream(),
                ),
                "ds_graph_conv_backward",
            )
        if db is not None:
            db = db.reshape(bias_shape)
        return dx, dk, db, None, None, None, None, None


def graph_conv(x, kernel, bias, plan, recursion, K, act=nat.ACT_LINEAR, mode=nat.MODE_FP32):
    return _GraphConv.apply(x, kernel, bias, plan, recursion, K, act, mode)


class _BiasAct(torch.autograd.Function):
    """y = act(z + bias) — the tail of Chebyshev.call when BatchNorm sits in between
    (gnn_layers.py:155-159)."""

    @staticmethod
    def forward(ctx, z, bias, act):
        _need_cuda(z, "bias/activation")
        z = _f32c(z)
        F = z.shape[-1]
        R = z.numel() // F
        bias_c = None if bias is None else _f32c(bias).reshape(-1)
        y = torch.empty_like(z)
        with torch.cuda.device(z.device.index):
            nat.check(
                nat.lib().ds_bias_act_forward(R, F, nat.ptr(z), nat.ptr(bias_c), act, nat.ptr(y), nat.current_stream()),
                "ds_bias_act_forward",
            )
        ctx.save_for_backward(y if act != nat.ACT_LINEAR else None)
        ctx.meta = (act, None if bias is None else bias.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        act, bias_shape = ctx.meta
        dy = _f32c(dy)
        F = dy.shape[-1]
        R = dy.numel() // F
        dz = torch.empty_like(dy)
        db = ws = None
        if bias_shape is not None:
            db = torch.empty(F, device=dy.device, dtype=torch.float32)
            ws = torch.empty(512 * F, device=dy.device, dtype=torch.float32)
        with torch.cuda.device(dy.device.index):
            nat.check(
                nat.lib().ds_bias_act_backward(
                    R, F, nat.ptr(y), nat.ptr(dy), act, nat.ptr(dz), nat.ptr(db), nat.ptr(ws), nat.current_stream()
                ),
                "ds_bias_act_backward",
            )
        return dz, None if db is None else db.reshape(bias_shape), None


class _BnBiasAct(torch.autograd.Function):
    """y = act(BatchNormalization(center=False, scale=False)(z) + bias)  (gnn_layers.py:152-159) through ds_bn_*.

    `rows` = (r0, r1): the rows of every sample that contribute to the statistics (and carry gradient); `group` /
    `sync`: with an initialised process group the 2F + 1 sums are all-reduced so that the statistics are those of the
    global batch - one collective per direction, issued between the two C-ABI calls."""

    @staticmethod
    def forward(ctx, z, bias, moving_mean, moving_var, act, training, eps, momentum, rows, sync_group):
        import torch.distributed as dist

        _need_cuda(z, "BatchNormalization")
        z = _f32c(z)
        B, M, F = z.shape
        r0, r1 = (0, M) if rows is None else (int(rows[0]), int(rows[1]))
        bias_c = None if bias is None else _f32c(bias).reshape(-1)
        dev = z.device
        lib = nat.lib()
        y = torch.empty_like(z)
        mean_rstd = torch.empty(2 * F, device=dev, dtype=torch.float32)
        scratch = torch.empty(2 * F, device=dev, dtype=torch.float32)
        sums = count_dev = None
        count = float(B * (r1 - r0))
        do_sync = bool(training) and sync_group is not False and dist.is_initialized() and \
            dist.get_world_size(None if sync_group is True else sync_group) > 1
        group = None if sync_group in (True, False) else sync_group
        with torch.cuda.device(dev.index):
            st = nat.current_stream()
            if training:
                ws = torch.empty(int(lib.ds_bn_workspace_doubles(B, M, F)), device=dev, dtype=torch.float64)
                sums = torch.empty(2 * F + 1, device=dev, dtype=torch.float64)
                nat.check(lib.ds_bn_stats(B, M, F, r0, r1, nat.ptr(z), nat.ptr(sums), nat.ptr(ws), st), "ds_bn_stats")
                if do_sync:  # the global row count travels with the sums and stays on the device (no host sync)
                    sums[2 * F:].fill_(count)  # (a fill kernel: an indexed assignment of a Python float is a host copy,
                    dist.all_reduce(sums, group=group)  # which a CUDA-graph capture refuses)
                    count_dev = sums[2 * F:]
            nat.check(
                lib.ds_bn_bias_act_forward(B, M, F, nat.ptr(z), nat.ptr(sums), count, nat.ptr(count_dev), float(eps),
                                           float(momentum),
                                           int(bool(training)), nat.ptr(moving_mean), nat.ptr(moving_var), nat.ptr(bias_c),
                                           act, nat.ptr(mean_rstd), nat.ptr(scratch), nat.ptr(y), st),
                "ds_bn_bias_act_forward",
            )
        ctx.save_for_backward(z, y if act != nat.ACT_LINEAR else None, mean_rstd, count_dev)
        ctx.meta = (act, bool(training), (r0, r1), count, do_sync, group, None if bias is None else bias.shape)
        ctx.mark_non_differentiable(moving_mean, moving_var)
        return y

    @staticmethod
    def backward(ctx, dy):
        import torch.distributed as dist

        z, y, mean_rstd, count_dev = ctx.saved_tensors
        act, training, (r0, r1), count, do_sync, group, bias_shape = ctx.meta
        dy = _f32c(dy)
        B, M, F = z.shape
        dev = z.device
        lib = nat.lib()
        dz = torch.empty_like(z)
        db = torch.empty(F, device=dev, dtype=torch.float32) if bias_shape is not None else None
        with torch.cuda.device(dev.index):
            st = nat.current_stream()
            ws = torch.empty(int(lib.ds_bn_workspace_doubles(B, M, F)), device=dev, dtype=torch.float64)
            sums = torch.empty(2 * F, device=dev, dtype=torch.float64)
            nat.check(lib.ds_bn_backward_stats(B, M, F, r0, r1, nat.ptr(z), nat.ptr(y), nat.ptr(dy), nat.ptr(mean_rstd), act,
                                               nat.ptr(sums), nat.ptr(ws), st), "ds_bn_backward_stats")
            local_sums = sums
            if do_sync:
                local_sums = sums.clone()  # dbias stays this rank's partial sum (the gradient all-reduce adds the ranks)
                dist.all_reduce(sums, group=group)
            nat.check(lib.ds_bn_backward_apply(B, M, F, r0, r1, nat.ptr(z), nat.ptr(y), nat.ptr(dy), nat.ptr(mean_rstd),
                                               nat.ptr(sums), count, nat.ptr(count_dev), act, int(training), nat.ptr(dz),
                                               None, st),
                      "ds_bn_backward_apply")
            if db is not None:
                db.copy_(local_sums[:F])
        return dz, None if db is None else db.reshape(bias_shape), None, None, None, None, None, None, None, None


def bn_bias_act(z, bias, bn, act, training, rows=None, sync_group=True):
    """BatchNormalization(center=False, scale=False) + bias + activation in the C-ABI kernels; `bn` is the
    keras_compat.BatchNormalization that owns the moving statistics."""
    return _BnBiasAct.apply(z, bias, bn.moving_mean, bn.moving_variance, act, training, bn.epsilon, bn.momentum, rows,
                            sync_group)


def bias_act(z, bias, act):
    if bias is None and act == nat.ACT_LINEAR:
        return z
    return _BiasAct.apply(z, bias, act)


class _Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, pool_type):
        _need_cuda(x, "HealpyPool")
        x = _f32c(x)
        B, M, F = x.shape
        r = 4**p
        y = torch.empty((B, M // r, F), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device.index):
            nat.check(
                nat.lib().ds_pool_forward(B, M, F, p, pool_type, nat.ptr(x), nat.ptr(y), nat.current_stream()),
                "ds_pool_forward",
            )
        ctx.save_for_backward(x if pool_type == nat.POOL_MAX else None)
        ctx.meta = (p, pool_type, (B, M, F))
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        p, pool_type, (B, M, F) = ctx.meta
        dy = _f32c(dy)
        dx = torch.empty((B, M, F), device=dy.device, dtype=torch.float32)
        with torch.cuda.device(dy.device.index):
            nat.check(
                nat.lib().ds_pool_backward(
                    B, M, F, p, pool_type, nat.ptr(x), nat.ptr(dy), nat.ptr(dx), nat.current_stream()
                ),
                "ds_pool_backward",
            )
        return dx, None, None


def pool(x, p, pool_type):
    return _Pool.apply(x, p, pool_type)


class _PseudoConv(torch.autograd.Function):
    """Conv1D(kernel = stride = 4^p) (transpose=False) / Conv2DTranspose((1,4^p)) (True)."""

    @staticmethod
    def forward(ctx, x, w, bias, p, Fout, act, mode, transpose):
        _need_cuda(x, "HealpyPseudoConv")
        x, w = _f32c(x), _f32c(w)
        bias_c = None if bias is None else _f32c(bias).reshape(-1)
        B, M, Fin = x.shape
        r = 4**p
        Mo = M * r if transpose else M // r
        y = torch.empty((B, Mo, Fout), device=x.device, dtype=torch.float32)
        fn = nat.lib().ds_pconvT_forward if transpose else nat.lib().ds_pconv_forward
        with torch.cuda.device(x.device.index):
            nat.check(
                fn(B, M, Fin, Fout, p, nat.ptr(x), nat.ptr(w), nat.ptr(bias_c), act, nat.ptr(y), mode,
                   nat.current_stream()),
                "ds_pconv_forward",
            )
        ctx.save_for_backward(x, w, y if act != nat.ACT_LINEAR else None)
        ctx.meta = (p, Fout, act, mode, transpose, None if bias is None else bias.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        p, Fout, act, mode, transpose, bias_shape = ctx.meta
        dy = _f32c(dy)
        B, M, Fin = x.shape
        L = nat.lib()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        db = torch.empty(Fout, device=x.device, dtype=torch.float32) if bias_shape is not None else None
        n_ws = (L.ds_pconvT_backward_workspace_elems if transpose else L.ds_pconv_backward_workspace_elems)(
            B, M, Fin, Fout, p, act
        )
        ws = torch.empty(n_ws, device=x.device, dtype=torch.float32)
        fn = L.ds_pconvT_backward if transpose else L.ds_pconv_backward
        with torch.cuda.device(x.device.index):
            nat.check(
                fn(B, M, Fin, Fout, p, nat.ptr(x), nat.ptr(w), nat.ptr(y), nat.ptr(dy), act, nat.ptr(dx), nat.ptr(dw),
                   nat.ptr(db), nat.ptr(ws), mode, nat.current_stream()),
                "ds_pconv_backward",
            )
        return dx, dw, None if db is None else db.reshape(bias_shape), None, None, None, None, None


def pseudo_conv(x, w, bias, p, Fout, act=nat.ACT_LINEAR, mode=nat.MODE_FP32, transpose=False):
    return _PseudoConv.apply(x, w, bias, p, Fout, act, mode, transpose)


def spmm(plan, x, alpha=1.0, prev=None, beta=0.0, add=None, gamma=0.0, transpose=False):
    """out = alpha * L~ x + beta * prev + gamma * add on [B, M, F] tensors (no autograd);
    the operator behind utils.split_sparse_dense_matmul."""
    _need_cuda(x, "spmm")
    x = _f32c(x)
    B, M, F = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device.index):
        nat.check(
            nat.lib().ds_spmm(
                plan.handle(x.device.index), 1 if transpose else 0, B, F, nat.ptr(x), alpha,
                nat.ptr(None if prev is None else _f32c(prev)), beta, nat.ptr(None if add is None else _f32c(add)),
                gamma, nat.ptr(out), nat.current_stream(),
            ),
            "ds_spmm",
        )
    return out


class _SparseMatmul(torch.autograd.Function):
    """y[b] = S x[b] for a fixed sparse S given as a GraphPlan; dx[b] = S^T dy[b]."""

    @staticmethod
    def forward(ctx, x, plan):
        ctx.plan = plan
        return spmm(plan, x)

    @staticmethod
    def backward(ctx, dy):
        return spmm(ctx.plan, dy, transpose=True), None


def sparse_matmul(plan, x):
    """Differentiable ``S @ x`` on [B, M, F] tensors (one ds_spmm launch each way)."""
    return _SparseMatmul.apply(x, plan)


def basis(plan, x, K, recursion=nat.RECURSION_CHEBYSHEV, transpose=False):
    """T_1..T_{K-1} of the recursion (gnn_layers.py:135-143) as a [K-1, B, M, F] tensor (no autograd)."""
    _need_cuda(x, "basis")
    x = _f32c(x)
    B, M, F = x.shape
    out = torch.empty((max(K - 1, 1), B, M, F), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device.index):
        nat.check(
            nat.lib().ds_graph_conv_basis(
                plan.handle(x.device.index), recursion, K, B, F, nat.ptr(x), nat.ptr(out), 1 if transpose else 0,
                nat.current_stream(),
            ),
            "ds_graph_conv_basis",
        )
    return out[: K - 1]
