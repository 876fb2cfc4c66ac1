"""ORACLE — CPU restatement of the reference's Chebyshev graph-convolution hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product (``deepsphere-cosmo-tf2_b200/``) never does, and has no
CPU fallback.

It restates, op for op and in the reference's own order, what the reference computes
through TensorFlow ops (file:line into /root/reference/src/deepsphere/):

* Laplacian preparation            gnn_layers.py:64-72, utils.py:40-46
* Chebyshev.call                   gnn_layers.py:130-161
* Monomial.call                    gnn_layers.py:281-309
* Bernstein.call                   gnn_layers.py:518-572 (incl. its stale last term)
* HealpySmoothing kernel + call    healpy_layers.py:725-764, 831-846
* BatchNormalization config        gnn_layers.py:53
* HealpyPool                       healpy_layers.py:48-63
* HealpyPseudoConv (Conv1D)        healpy_layers.py:118-126
* HealpyPseudoConv_Transpose       healpy_layers.py:180-188,215-216
* extend_indices                   utils.py:9-37
* HealpyGCNN._transform_indices    healpy_networks.py:169-188

PARITY PINNING STATUS
---------------------
The arithmetic of this path lives in un-vendored third-party code: TensorFlow
(``tensorflow>=2.14.0``, setup.cfg:19 — ``tf.sparse.sparse_dense_matmul``, ``tf.matmul``,
Keras pooling/conv/BN), healpy (unpinned, setup.cfg:18) and PyGSP (git branch,
setup.cfg:21).  None of them is installable in this image, so the reference cannot be
executed here and no reference-generated golden vectors exist.

* HealpyPool: PINNED by the reference's own known-answer test
  (tests/test_healpy_layers.py:22-37: ``np.random.seed(11)``, nside 4 -> 2, AVG equals
  ``hp.ud_grade`` (mean of the 4 nested children), MAX equals a reshape-max, tol 1e-5).
* extend_indices / index bookkeeping: PINNED by tests/test_utils.py:7-31 (every-4th-pixel
  set extends to all pixels) and by the notebook known answer
  examples/advanced_tutorial.ipynb:137,211,356 (24 832 -> 6 208 -> 1 552 -> 388 pixels).
* Parameter counts / shapes: PINNED by the Keras summaries in the notebooks.
* Chebyshev / Monomial / PseudoConv / PseudoConv_Transpose forward VALUES and every
  GRADIENT: **parity unpinned** — the reference tests assert no values for them
  (tests/test_gnn_layers.py:9-62 only call the layers).  The restatement below is the
  published algorithm of those TF ops (SpMM as sum over the row's non-zeros in column
  order, dense matmul, Keras BN/Conv semantics) applied in the reference's op order.

Two precisions: ``dtype=np.float32`` reproduces the reference's floatx arithmetic (the
parity target, rel <= 1e-5); ``dtype=np.float64`` is the error yardstick.
"""

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import eigsh

# --------------------------------------------------------------------------------------
# Laplacian preparation
# --------------------------------------------------------------------------------------


def rescale_L(L, lmax=2, scale=1):
    """utils.py:40-46 — ``L *= 2*scale/lmax; L -= I`` (here on a copy)."""
    L = sparse.csr_matrix(L, dtype=np.float64, copy=True)
    M = L.shape[0]
    L = L * (2 * scale / lmax)
    L = L - sparse.identity(M, format="csr", dtype=L.dtype)
    return sparse.csr_matrix(L)


def prepare_laplacian(L, scale):
    """gnn_layers.py:64-72 — csr, lmax = 1.02*eigsh(L,k=1,'LM'), rescale, row-major COO.

    scale = 0.75 for Chebyshev (gnn_layers.py:67), 1.0 for Monomial (gnn_layers.py:219).
    Returns (L_tilde CSR float64 with sorted indices, lmax)."""
    L = sparse.csr_matrix(L, dtype=np.float64, copy=True)
    lmax = 1.02 * eigsh(L, k=1, which="LM", return_eigenvectors=False)[0]
    Lt = rescale_L(L, lmax=lmax, scale=scale)
    Lt.sort_indices()  # tf.sparse.reorder, gnn_layers.py:115
    return Lt, float(lmax)


# --------------------------------------------------------------------------------------
# activations / batch norm
# --------------------------------------------------------------------------------------


def _act(name):
    if name is None or name == "linear":
        return lambda v: v
    table = {
        "relu": lambda v: np.maximum(v, 0),
        "elu": lambda v: np.where(v > 0, v, np.expm1(np.minimum(v, 0))),
        "sigmoid": lambda v: 1 / (1 + np.exp(-v)),
        "tanh": np.tanh,
        "softplus": lambda v: np.logaddexp(v, 0),
    }
    if name not in table:
        raise ValueError(f"oracle has no activation {name}")
    return table[name]


def batch_norm(z, training, moving_mean=None, moving_var=None, momentum=0.9, eps=1e-5):
    """Keras BatchNormalization(axis=-1, momentum=0.9, epsilon=1e-5, center=False,
    scale=False) — gnn_layers.py:53.  Returns (out, new_moving_mean, new_moving_var)."""
    F = z.shape[-1]
    mm = np.zeros(F, z.dtype) if moving_mean is None else moving_mean
    mv = np.ones(F, z.dtype) if moving_var is None else moving_var
    if training:
        mean = z.mean(axis=(0, 1))
        var = z.var(axis=(0, 1))  # biased, as Keras
        out = (z - mean) / np.sqrt(var + z.dtype.type(eps))
        mm = mm * momentum + mean * (1 - momentum)
        mv = mv * momentum + var * (1 - momentum)
        return out.astype(z.dtype), mm.astype(z.dtype), mv.astype(z.dtype)
    out = (z - mm) / np.sqrt(mv + z.dtype.type(eps))
    return out.astype(z.dtype), mm, mv


# --------------------------------------------------------------------------------------
# Chebyshev / Monomial forward in the reference's op order
# --------------------------------------------------------------------------------------


def _basis_stack(x, Lt, K, recursion, dtype):
    """gnn_layers.py:131-147 — returns X [N*M, Fin*K] with column order f*K + k."""
    N, M, Fin = x.shape
    Ls = sparse.csr_matrix(Lt, dtype=dtype)  # values rounded to floatx, gnn_layers.py:71
    x0 = np.transpose(x.astype(dtype), (1, 2, 0)).reshape(M, Fin * N)  # :131-132
    stack = [x0]
    if recursion == "chebyshev":
        if K > 1:
            x1 = Ls @ x0  # :138
            stack.append(x1)
        for _ in range(2, K):
            x2 = dtype(2) * (Ls @ x1) - x0  # :141
            stack.append(x2)
            x0, x1 = x1, x2
    elif recursion == "monomial":
        for _ in range(1, K):  # gnn_layers.py:287-290
            x1 = Ls @ x0
            stack.append(x1)
            x0 = x1
    else:
        raise ValueError(recursion)
    X = np.stack(stack, axis=0)  # K x M x Fin*N           :144
    X = X.reshape(K, M, Fin, N)  #                          :145
    X = np.transpose(X, (3, 1, 2, 0))  # N x M x Fin x K    :146
    return np.ascontiguousarray(X).reshape(N * M, Fin * K)  # :147


def graph_conv_forward(
    x, Lt, kernel, K, recursion="chebyshev", bias=None, activation=None, use_bn=False, training=False,
    bn_state=None, dtype=np.float32,
):
    """Chebyshev.call (gnn_layers.py:106-161) / Monomial.call (:255-309).

    x [N, M, Fin]; Lt = prepared L_tilde (M x M sparse); kernel [K*Fin, Fout] with row
    order f*K + k; bias [1,1,Fout] or None.  BN comes before the bias (:152-156)."""
    dtype = np.dtype(dtype).type
    N, M, Fin = x.shape
    X = _basis_stack(x, Lt, K, recursion, dtype)
    z = X @ kernel.astype(dtype)  # :149
    z = z.reshape(N, M, -1)  # :150
    if use_bn:
        mm, mv = bn_state if bn_state is not None else (None, None)
        z, _, _ = batch_norm(z, training, mm, mv)
    if bias is not None:
        z = z + bias.astype(dtype).reshape(1, 1, -1)
    return _act(activation)(z).astype(dtype)


def graph_conv_backward(x, Lt, kernel, K, dy, recursion="chebyshev", dtype=np.float64):
    """Gradients of z = graph_conv_forward(x) (no BN / bias / activation) for an upstream
    gradient dy [N, M, Fout]: returns (dx, dkernel, dbias).  This is what TF autodiff
    yields for gnn_layers.py:131-150: dkernel = X^T dY; dX = dY kernel^T
    pushed back through the transposed recursion."""
    dtype = np.dtype(dtype).type
    N, M, Fin = x.shape
    Fout = kernel.shape[1]
    X = _basis_stack(x, Lt, K, recursion, dtype)
    dY = dy.astype(dtype).reshape(N * M, Fout)
    dkernel = X.T @ dY
    dbias = dY.sum(axis=0).reshape(1, 1, Fout)
    G = dY @ kernel.astype(dtype).T  # [N*M, Fin*K], column f*K + k
    G = G.reshape(N, M, Fin, K)
    G = np.transpose(G, (3, 1, 2, 0)).reshape(K, M, Fin * N)  # per-k gradient of the stack
    LT = sparse.csr_matrix(Lt, dtype=dtype).T.tocsr()
    g = [G[k].copy() for k in range(K)]
    if recursion == "chebyshev":
        # reverse of x_k = 2 L x_{k-1} - x_{k-2}  (k >= 2),  x_1 = L x_0
        for k in range(K - 1, 1, -1):
            g[k - 1] += dtype(2) * (LT @ g[k])
            g[k - 2] -= g[k]
        if K > 1:
            g[0] += LT @ g[1]
    else:
        for k in range(K - 1, 0, -1):
            g[k - 1] += LT @ g[k]
    dx0 = g[0].reshape(M, Fin, N)
    dx = np.transpose(dx0, (2, 0, 1))
    return np.ascontiguousarray(dx), dkernel, dbias


def bernstein_forward(x, Lt, kernel, K, bias=None, activation=None, dtype=np.float32):
    """Bernstein.call (gnn_layers.py:518-572), loop for loop: K = polynomial ORDER, kernel
    [(K+1)*Fin, Fout] with row order f*(K+1) + i, Lt prepared with scale 0.75 (:473).

    The stale ``x3`` of the last term is reproduced as written (:543-554): for i = K the loop
    ``range(K - i)`` is empty, so ``x3`` is still the (already theta-scaled) tensor of i = K-1."""
    from math import comb

    dtype = np.dtype(dtype).type
    N, M, Fin = x.shape
    Ls = sparse.csr_matrix(Lt, dtype=dtype)
    x0 = np.transpose(x.astype(dtype), (1, 2, 0)).reshape(M, Fin * N)  # :535-536
    stack = []
    x3 = None
    for i in range(0, K + 1):  # :540
        x1 = x0
        theta = dtype(comb(K, i) / (2**K))
        for _ in range(i):
            x1 = Ls @ x1
        x2 = x1
        for _ in range(K - i):
            x3 = dtype(2) * x2 - Ls @ x2
            x2 = x3
        x3 = theta * x3  # NameError in the reference when K = 0
        stack.append(x3)
    X = np.stack(stack, axis=0).reshape(K + 1, M, Fin, N)  # :555-556
    X = np.ascontiguousarray(np.transpose(X, (3, 1, 2, 0))).reshape(N * M, Fin * (K + 1))  # :557-558
    z = (X @ kernel.astype(dtype)).reshape(N, M, -1)  # :560-561
    if bias is not None:
        z = z + bias.astype(dtype).reshape(1, 1, -1)
    return _act(activation)(z).astype(dtype)


def smoothing_neighbours(lat, lon, sigma_rad, n_sigma_support=3):
    """HealpySmoothing._build_tree + _build_kernel (healpy_layers.py:766-829) with the reference's own tool,
    a scikit-learn BallTree under the haversine metric: every pixel gets its ``max_neighbors`` nearest pixels,
    max_neighbors = the largest population of any radius-(n_sigma * sigma) ball.  Returns (ind_coo [nnz, 2]
    int64, val_coo [nnz] float32)."""
    from sklearn.neighbors import BallTree

    theta = np.stack([lat, lon], axis=1)
    tree = BallTree(theta, metric="haversine")
    inds_r = tree.query_radius(theta, r=sigma_rad * n_sigma_support)
    max_neighbors = int(np.max([len(i) for i in inds_r]))
    dist_k, inds_k = tree.query(theta, k=max_neighbors, return_distance=True, sort_results=True)
    kernel_k = np.exp(-0.5 / sigma_rad**2 * dist_k**2).astype(np.float32)
    n = len(lat)
    rows = np.repeat(np.arange(n, dtype=np.int64)[:, None], max_neighbors, axis=1)
    ind_coo = np.concatenate([rows.reshape(-1, 1), inds_k.astype(np.int64).reshape(-1, 1)], axis=1)
    return ind_coo, kernel_k.reshape(-1)


def smoothing_kernel(ind_coo, val_coo, n):
    """HealpySmoothing._build_sparse_tensor (healpy_layers.py:831-846): COO -> row-major sparse,
    then ``sparse_kernel / expand_dims(reduce_sum(sparse_kernel, axis=1), axis=0)``.  The row sums
    get shape [1, n] and broadcast along the LAST axis: entry (i, j) is divided by the row sum of row
    j, not of row i (the comment there says "rows have to sum to one"; they only do when the kernel
    is symmetric).  Restated as written."""
    Ks = sparse.csr_matrix((np.asarray(val_coo, dtype=np.float32), (ind_coo[:, 0], ind_coo[:, 1])), shape=(n, n))
    row_sum = np.asarray(Ks.sum(axis=1)).ravel().astype(np.float32)
    coo = Ks.tocoo()
    return sparse.csr_matrix((coo.data / row_sum[coo.col], (coo.row, coo.col)), shape=(n, n))


def smoothing_forward(x, Ks, per_channel_repetitions=None, mask=None, dtype=np.float32):
    """HealpySmoothing.call (healpy_layers.py:725-764): every channel [n_indices, n_batch] is multiplied
    by the sparse kernel once, or per_channel_repetitions[i] times; then the optional mask."""
    dtype = np.dtype(dtype).type
    Kd = sparse.csr_matrix(Ks, dtype=dtype)
    xt = np.transpose(x.astype(dtype), (1, 0, 2))  # :732
    out = []
    for i in range(xt.shape[2]):  # :738-748
        c = xt[:, :, i]
        reps = 1 if per_channel_repetitions is None else int(per_channel_repetitions[i])
        for _ in range(reps):
            c = Kd @ c
        out.append(c)
    y = np.transpose(np.stack(out, axis=2), (1, 0, 2))  # :751-754
    if mask is not None:
        y = y * mask.astype(dtype)
    return y


# --------------------------------------------------------------------------------------
# Pool / pseudo-convolutions
# --------------------------------------------------------------------------------------


def healpy_pool(x, p, pool_type="MAX"):
    """healpy_layers.py:48-63 — MaxPool1D / AveragePooling1D(pool=stride=4^p, 'valid')."""
    if not p >= 1:
        raise IOError("The reduction factors has to be at least 2!")
    r = int(4**p)
    N, M, F = x.shape
    if M % r != 0:
        raise IOError("Input shape not compatible with the filter size")
    xr = x.reshape(N, M // r, r, F)
    if pool_type == "MAX":
        return xr.max(axis=2)
    if pool_type == "AVG":
        return xr.mean(axis=2).astype(x.dtype)
    raise IOError(f"Pooling type not understood: {pool_type}")


def healpy_pool_backward(x, dy, p, pool_type="MAX"):
    r = int(4**p)
    N, M, F = x.shape
    xr = x.reshape(N, M // r, r, F)
    if pool_type == "AVG":
        return np.repeat(dy / r, r, axis=1).astype(x.dtype)
    arg = xr.argmax(axis=2)  # first maximum wins, as TF's MaxPoolGrad
    dx = np.zeros_like(xr)
    n, j, f = np.meshgrid(np.arange(N), np.arange(M // r), np.arange(F), indexing="ij")
    dx[n, j, arg, f] = dy
    return dx.reshape(N, M, F)


def pseudo_conv(x, kernel, bias=None, activation=None):
    """healpy_layers.py:118-126 — Conv1D(Fout, 4^p, strides=4^p, 'valid').
    kernel [4^p, Fin, Fout] (Keras layout), bias [Fout]."""
    r, Fin, Fout = kernel.shape
    N, M, _ = x.shape
    z = x.reshape(N * (M // r), r * Fin) @ kernel.reshape(r * Fin, Fout).astype(x.dtype)
    z = z.reshape(N, M // r, Fout)
    if bias is not None:
        z = z + bias.astype(x.dtype)
    return _act(activation)(z).astype(x.dtype)


def pseudo_conv_transpose(x, kernel, bias=None, activation=None):
    """healpy_layers.py:180-188,215-216 — Conv2DTranspose(Fout, (1,4^p), strides (1,4^p)).
    kernel [1, 4^p, Fout, Fin] (Keras layout), bias [Fout].
    y[b, 4^p*j + c, o] = sum_f x[b,j,f] * kernel[0,c,o,f] + bias[o]."""
    _, r, Fout, Fin = kernel.shape
    N, M, _ = x.shape
    Wt = kernel[0].reshape(r * Fout, Fin).astype(x.dtype)
    z = x.reshape(N * M, Fin) @ Wt.T  # [N*M, r*Fout]
    z = z.reshape(N, M * r, Fout)
    if bias is not None:
        z = z + bias.astype(x.dtype)
    return _act(activation)(z).astype(x.dtype)


# --------------------------------------------------------------------------------------
# Index bookkeeping (pure NESTED integer arithmetic; the reference goes through hp.ud_grade
# of a 0/1 mask and a > 1e-12 threshold, which in NEST is exactly this)
# --------------------------------------------------------------------------------------


def extend_indices(indices, nside_in, nside_out):
    """utils.py:9-37 (nest=True)."""
    r = (nside_in // nside_out) ** 2
    parents = np.unique(np.asarray(indices, dtype=np.int64) // r)
    return (parents[:, None] * r + np.arange(r)[None, :]).ravel()


def transform_indices(indices, nside_in, nside_out):
    """healpy_networks.py:169-188."""
    indices = np.asarray(indices, dtype=np.int64)
    if nside_in == nside_out:
        return indices
    if nside_out < nside_in:
        return np.unique(indices // (nside_in // nside_out) ** 2)
    r = (nside_out // nside_in) ** 2
    return (np.sort(indices)[:, None] * r + np.arange(r)[None, :]).ravel()


# --------------------------------------------------------------------------------------
# torch-CPU restatement (all host threads) — used for the timed CPU baseline and as an
# independent autograd gradient oracle.
# --------------------------------------------------------------------------------------


def torch_cpu_graph_conv(x, Lt, kernel, K, recursion="chebyshev"):
    """Same op sequence as gnn_layers.py:131-150 with torch CPU ops (torch.sparse.mm,
    matmul).  x, kernel: torch CPU tensors (may require grad); Lt scipy sparse."""
    import torch

    N, M, Fin = x.shape
    coo = sparse.coo_matrix(Lt)
    Ls = torch.sparse_coo_tensor(
        np.vstack([coo.row, coo.col]), torch.as_tensor(coo.data, dtype=x.dtype), size=coo.shape
    ).coalesce()
    x0 = x.permute(1, 2, 0).reshape(M, Fin * N)
    stack = [x0]
    if recursion == "chebyshev":
        if K > 1:
            x1 = torch.sparse.mm(Ls, x0)
            stack.append(x1)
        for _ in range(2, K):
            x2 = 2 * torch.sparse.mm(Ls, x1) - x0
            stack.append(x2)
            x0, x1 = x1, x2
    else:
        for _ in range(1, K):
            x1 = torch.sparse.mm(Ls, x0)
            stack.append(x1)
            x0 = x1
    X = torch.stack(stack, dim=0).reshape(K, M, Fin, N).permute(3, 1, 2, 0).reshape(N * M, Fin * K)
    return (X @ kernel).reshape(N, M, -1)


def torch_cpu_bernstein(x, Lt, kernel, K):
    """Bernstein.call (gnn_layers.py:518-561) with differentiable torch CPU ops in x's dtype, loop for loop
    (incl. the stale last term) — the autograd yardstick for the Bernstein gradients."""
    import torch
    from math import comb

    N, M, Fin = x.shape
    coo = sparse.coo_matrix(Lt)
    Ls = torch.sparse_coo_tensor(np.vstack((coo.row, coo.col)), torch.tensor(coo.data, dtype=x.dtype), coo.shape).coalesce()
    x0 = x.permute(1, 2, 0).reshape(M, Fin * N)
    stack = []
    x3 = None
    for i in range(K + 1):
        x1 = x0
        for _ in range(i):
            x1 = torch.sparse.mm(Ls, x1)
        x2 = x1
        for _ in range(K - i):
            x3 = 2 * x2 - torch.sparse.mm(Ls, x2)
            x2 = x3
        x3 = (comb(K, i) / 2**K) * x3
        stack.append(x3)
    X = torch.stack(stack, dim=0).reshape(K + 1, M, Fin, N).permute(3, 1, 2, 0).reshape(N * M, Fin * (K + 1))
    return (X @ kernel).reshape(N, M, -1)


def torch_cpu_layer(x, Lt, kernel, K, recursion="chebyshev", bias=None, activation=None, use_bn=False, training=False,
                    moving_mean=None, moving_var=None, eps=1e-5):
    """A whole Chebyshev / Monomial layer (gnn_layers.py:130-161: contraction -> BatchNormalization(center=False,
    scale=False) -> + bias -> activation) with differentiable torch CPU ops in x's dtype."""
    import torch

    z = torch_cpu_graph_conv(x, Lt, kernel, K, recursion)
    if use_bn:
        if training:
            mean, var = z.mean(dim=(0, 1)), z.var(dim=(0, 1), unbiased=False)
        else:
            F = z.shape[-1]
            mean = torch.zeros(F, dtype=z.dtype) if moving_mean is None else moving_mean
            var = torch.ones(F, dtype=z.dtype) if moving_var is None else moving_var
        z = (z - mean) / torch.sqrt(var + eps)
    if bias is not None:
        z = z + bias
    return _torch_act(activation)(z)


def _torch_act(name):
    import torch

    if name is None or name == "linear":
        return lambda v: v
    if callable(name):
        return name
    return {"relu": torch.relu, "elu": torch.nn.functional.elu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
            "softplus": torch.nn.functional.softplus}[name]


def torch_cpu_residual(x, Lt, kernels, K, recursion="chebyshev", layer_activation=None, layer_biases=(None, None),
                       layer_use_bn=False, activation=None, act_before=False, use_bn=False, norm_type="batch_norm",
                       alpha=1.0, training=False, eps_bn=1e-3, eps_ln=1e-3, ln_axis=-1,
                       layer_moving=((None, None), (None, None))):
    """GCNN_ResidualLayer.call (gnn_layers.py:384-413): in -> layer1 -> [norm] -> layer2 -> [norm] -> skip.

    Both sub-layers are built from the same ``layer_kwargs`` (gnn_layers.py:365-370) and are called WITHOUT an explicit
    ``training`` argument (:391, :398); inside a Keras model call the training flag of the enclosing call context
    propagates to them, so a sub-layer's own BatchNormalization (``use_bn`` in layer_kwargs) follows ``training`` too.
    The norms between the layers are Keras defaults (``BatchNormalization(axis=-1)``: momentum 0.99, epsilon 1e-3, learned
    gamma = 1 / beta = 0 at initialisation; ``LayerNormalization(axis=...)``: epsilon 1e-3) — evaluated here at their
    initial affine parameters.  Without an activation the skip is ``x + in`` (alpha ignored, :407-408)."""
    import torch

    def norm(z):
        if norm_type == "batch_norm":
            if training:
                mean, var = z.mean(dim=(0, 1)), z.var(dim=(0, 1), unbiased=False)
            else:
                mean, var = torch.zeros(z.shape[-1], dtype=z.dtype), torch.ones(z.shape[-1], dtype=z.dtype)
            return (z - mean) / torch.sqrt(var + eps_bn)
        axes = ln_axis if isinstance(ln_axis, (tuple, list)) else (ln_axis,)
        axes = tuple(a % z.dim() for a in axes)
        mean = z.mean(dim=axes, keepdim=True)
        var = z.var(dim=axes, unbiased=False, keepdim=True)
        return (z - mean) / torch.sqrt(var + eps_ln)

    h = torch_cpu_layer(x, Lt, kernels[0], K, recursion, bias=layer_biases[0], activation=layer_activation,
                        use_bn=layer_use_bn, training=training, moving_mean=layer_moving[0][0], moving_var=layer_moving[0][1])
    if use_bn:
        h = norm(h)
    h = torch_cpu_layer(h, Lt, kernels[1], K, recursion, bias=layer_biases[1], activation=layer_activation,
                        use_bn=layer_use_bn, training=training, moving_mean=layer_moving[1][0], moving_var=layer_moving[1][1])
    if use_bn:
        h = norm(h)
    if activation is None:
        return h + x
    act = _torch_act(activation)
    if act_before:
        return act(h) + alpha * x
    return act(h + alpha * x)


# --------------------------------------------------------------------------------------
# Whole networks (healpy_networks.HealpyGCNN as a Sequential of the layers above) with differentiable torch CPU ops:
# the yardstick for network-level parity (forward AND every weight gradient) and the timed CPU baseline of the named
# configurations.  A network is a list of specs (kind, params) whose weights are torch tensors:
#   ("conv", dict(Lt, K, recursion, kernel, bias|None, activation, use_bn))      gnn_layers.py:130-161 / 281-309
#   ("pool", dict(p, pool_type))                                                  healpy_layers.py:48-63
#   ("pconv", dict(kernel [4^p, Fin, Fout], bias, activation))                    healpy_layers.py:118-126
#   ("pconvT", dict(kernel [1, 4^p, Fout, Fin], bias, activation))                healpy_layers.py:180-188
#   ("residual", dict(Lt, K, recursion, kernels, biases, layer_activation, layer_use_bn, activation, act_before, use_bn,
#                     norm_type, alpha))                                          gnn_layers.py:384-413
#   ("layernorm", dict(axis, gamma, beta, eps))                                   tf.keras.layers.LayerNormalization
#   ("mean", {}) / ("mean_softmax", {})     reduce_mean over pixels (+ softmax), the Lambda heads of the notebooks
#   ("dense", dict(kernel [Fin, Fout], bias))
# --------------------------------------------------------------------------------------


def torch_cpu_pool(x, p, pool_type="MAX"):
    r = int(4**p)
    N, M, F = x.shape
    xr = x.reshape(N, M // r, r, F)
    return xr.max(dim=2).values if pool_type == "MAX" else xr.mean(dim=2)


def torch_cpu_pconv(x, kernel, bias=None, activation=None):
    r, Fin, Fout = kernel.shape
    N, M, _ = x.shape
    z = (x.reshape(N * (M // r), r * Fin) @ kernel.reshape(r * Fin, Fout)).reshape(N, M // r, Fout)
    if bias is not None:
        z = z + bias
    return _torch_act(activation)(z)


def torch_cpu_pconv_transpose(x, kernel, bias=None, activation=None):
    _, r, Fout, Fin = kernel.shape
    N, M, _ = x.shape
    z = (x.reshape(N * M, Fin) @ kernel[0].reshape(r * Fout, Fin).T).reshape(N, M * r, Fout)
    if bias is not None:
        z = z + bias
    return _torch_act(activation)(z)


def torch_cpu_network(x, specs, training=False):
    import torch

    h = x
    for kind, p in specs:
        if kind == "conv":
            h = torch_cpu_layer(h, p["Lt"], p["kernel"], p["K"], p["recursion"], bias=p.get("bias"),
                                activation=p.get("activation"), use_bn=p.get("use_bn", False), training=training,
                                moving_mean=p.get("moving_mean"), moving_var=p.get("moving_var"))
        elif kind == "pool":
            h = torch_cpu_pool(h, p["p"], p["pool_type"])
        elif kind == "pconv":
            h = torch_cpu_pconv(h, p["kernel"], p.get("bias"), p.get("activation"))
        elif kind == "pconvT":
            h = torch_cpu_pconv_transpose(h, p["kernel"], p.get("bias"), p.get("activation"))
        elif kind == "residual":
            h = torch_cpu_residual(h, p["Lt"], p["kernels"], p["K"], p["recursion"], layer_activation=p.get("layer_activation"),
                                   layer_biases=p.get("biases", (None, None)), layer_use_bn=p.get("layer_use_bn", False),
                                   activation=p.get("activation"), act_before=p.get("act_before", False),
                                   use_bn=p.get("use_bn", False), norm_type=p.get("norm_type", "batch_norm"),
                                   alpha=p.get("alpha", 1.0), training=training,
                                   layer_moving=p.get("layer_moving", ((None, None), (None, None))))
        elif kind == "layernorm":
            axes = p["axis"] if isinstance(p["axis"], (tuple, list)) else (p["axis"],)
            axes = tuple(a % h.dim() for a in axes)
            mean = h.mean(dim=axes, keepdim=True)
            var = h.var(dim=axes, unbiased=False, keepdim=True)
            h = (h - mean) / torch.sqrt(var + p.get("eps", 1e-3))
            shape = [h.shape[a] if a in axes else 1 for a in range(h.dim())]
            if p.get("gamma") is not None:
                h = h * p["gamma"].reshape(shape)
            if p.get("beta") is not None:
                h = h + p["beta"].reshape(shape)
        elif kind == "mean":
            h = h.mean(dim=1)
        elif kind == "mean_softmax":
            h = torch.softmax(h.mean(dim=1), dim=-1)
        elif kind == "dense":
            h = h @ p["kernel"]
            if p.get("bias") is not None:
                h = h + p["bias"]
        else:
            raise ValueError(f"oracle network: unknown layer kind {kind}")
    return h
