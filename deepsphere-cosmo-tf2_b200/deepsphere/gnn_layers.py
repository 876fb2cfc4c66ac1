"""Graph convolution layers — same names, constructor arguments, weights and errors as the
reference's src/deepsphere/gnn_layers.py, executed by the sm_100a C-ABI library.

What happens where:
  host, once per layer   Laplacian prep exactly as gnn_layers.py:64-72 (csr -> lmax by
                         ARPACK -> rescale_L -> COO int64 indices + float32 values) and the
                         device plan (fixed-width ELL + CSR tail of L~ and L~^T).
  device, per call       everything between gnn_layers.py:113 and :159 as C-ABI calls: the
                         recursion in the native [B, M, F] layout (no transposes), the
                         K*Fin -> Fout contraction reading the basis tensors in place, bias +
                         activation fused into the contraction epilogue when no BatchNorm sits
                         in between.
"""

import os

import numpy as np
import torch
from scipy import sparse

from . import _native as nat
from . import _ops
from . import logger
from . import utils
from .keras_compat import BatchNormalization, LayerNormalization, Model, TruncatedNormal, resolve_activation


def _default_mode():
    """Arithmetic of the contraction: env DEEPSPHERE_MODE in {fp32, tf32, tf32x3} (default fp32)."""
    name = os.environ.get("DEEPSPHERE_MODE", "fp32").lower()
    if name not in nat.MODES:
        raise ValueError(f"DEEPSPHERE_MODE must be one of {sorted(nat.MODES)}, got {name}")
    return nat.MODES[name]


def _to_csr(L):
    """The reference accepts numpy arrays, scipy sparse matrices and tf tensors (tests/
    test_gnn_layers.py:22, test_healpy_layers.py:80, healpy_networks.py:117); here: numpy,
    scipy sparse, torch tensors.  Always a copy (the reference scales a CSR input in place,
    SURVEY A.3 — benign there, avoided here)."""
    if isinstance(L, torch.Tensor):
        L = L.detach().cpu().numpy()
    return sparse.csr_matrix(L, dtype=np.float64, copy=True)


class _GraphConvBase(Model):
    """Shared shell of Chebyshev and Monomial (the reference duplicates it verbatim,
    gnn_layers.py:17-104 and :169-253)."""

    _recursion = None
    _scale = None

    @property
    def _n_terms(self):
        """Number of basis tensors T_0..T_{n-1} of the recursion = rows of the kernel per input channel."""
        return self.K

    def _device_kernel(self):
        """The [n_terms*Fin, Fout] weights the C-ABI contraction receives (row order f*n_terms + k)."""
        return self.kernel

    def __init__(self, L, K, Fout=None, initializer=None, activation=None, use_bias=False, use_bn=False,
                 n_matmul_splits=1, **kwargs):
        super().__init__()
        self.L = L
        self.K = int(K)
        if self.K < 1:
            raise ValueError(f"K must be at least 1, got {K}")
        self.Fout = Fout
        self.use_bias = use_bias
        self.use_bn = use_bn
        if self.use_bn:
            # gnn_layers.py:53
            self.bn = BatchNormalization(axis=-1, momentum=0.9, epsilon=1e-5, center=False, scale=False)
        self.initializer = initializer
        # gnn_layers.py:55-60 — unknown name -> ValueError
        self._act_id, self.activation = resolve_activation(activation)
        # accepted for API compatibility; the kernel has no 2^31 size limit (utils.py:59)
        self.n_matmul_splits = n_matmul_splits
        self.mode = kwargs.pop("mode", None)
        # optional HEALPix geometry (nside, NESTED pixel ids of the rows of L): enables the fused lattice
        # kernel; HealpyGCNN provides it, and a full-sphere Laplacian is recognised automatically
        self._healpix = kwargs.pop("healpix", None)
        # BatchNormalization statistics: rows of every sample that contribute (None = all; a sphere-partitioned layer
        # sets its own rows, partition.PartitionedGraphConv) and the process group they are summed over (True = the
        # default group when one is initialised, False = never, or a group)
        self._bn_rows = None
        self._bn_sync = True
        # largest eigenvalue of L supplied by the caller (skips ARPACK): the sphere-partitioned layer passes the
        # GLOBAL value so that every rank's restricted Laplacian is the restriction of the same L~
        lmax_given = kwargs.pop("lmax", None)
        self.kwargs = kwargs

        # gnn_layers.py:64-72: rescale the Laplacian and keep it as COO (indices, values, shape)
        Lc = _to_csr(L)
        # 1.02 * eigsh(L, k=1, which="LM")[0] of the reference; same value, ~7x fewer products with L at nside 256
        lmax = float(lmax_given) if lmax_given is not None else 1.02 * utils.largest_eigenvalue(Lc)
        self.lmax = float(lmax)
        Lc = utils.rescale_L(Lc, lmax=lmax, scale=self._scale)
        L_coo = Lc.tocoo()
        self._L_indices = np.column_stack((L_coo.row, L_coo.col)).astype(np.int64)
        self._L_values = L_coo.data.astype(np.float32)  # floatx
        self._L_shape = np.asarray(L_coo.shape, dtype=np.int64)
        self._plan = nat.GraphPlan(self._L_indices, self._L_values, self._L_shape, lattice_builder=self._lattice_payload)

    def _attach_healpix(self, nside, indices):
        """Tell the layer which HEALPix pixels its graph lives on (called by HealpyGCNN)."""
        self._healpix = (int(nside), np.asarray(indices, dtype=np.int64))

    def _lattice_payload(self):
        """Tile plan of the fused K-hop kernel, or None when it does not apply (K = 1, no HEALPix
        geometry, not an 8-neighbour graph, resolution below the 16 x 16 tile)."""
        from . import healpix as hpx
        from . import lattice

        if self._n_terms < 2:
            return None
        geo = self._healpix
        M = int(self._L_shape[0])
        if geo is None:
            if M % 12 != 0:
                return None
            nside = int(round(np.sqrt(M // 12)))
            if 12 * nside * nside != M or not hpx.isnsideok(nside, nest=True):
                return None
            geo = (nside, np.arange(M, dtype=np.int64))
        nside, indices = geo
        if len(indices) != M or nside < 16:
            return None
        Lt = sparse.csr_matrix((self._L_values, (self._L_indices[:, 0], self._L_indices[:, 1])), shape=(M, M))
        try:
            # always the 4-ring halo (24 x 24 lattice): the geometry of the register-resident fused kernel
            # (ds_lattice_conv2.cu) and of the fp32 lattice recursion; K - 1 <= 4 hops run in one pass on it, longer
            # recursions are chained in passes of <= 4 hops by the C-ABI (ds_graph_conv_*), so the tables never grow
            # with K (a (16 + 2 (K-1))^2 lattice would cost 40 B x positions per tile: 2.3 GB at nside 1024, K = 10)
            return lattice.make_payload(Lt, nside, indices, 4)
        except Exception as exc:  # never let the optional fast path break the layer
            logger.warning(f"lattice plan construction failed ({exc!r}); using the generic kernels")
            return None

    def _default_initializer(self, Fin, Fout):
        raise NotImplementedError

    def build(self, input_shape):
        """gnn_layers.py:74-104: kernel [K*Fin, Fout] (row order f*K + k), bias [1, 1, Fout]."""
        Fin = int(input_shape[-1])
        Fout = Fin if self.Fout is None else int(self.Fout)
        if self.initializer is None:
            initializer = self._default_initializer(Fin, Fout)
        else:
            logger.debug(self.kwargs)
            initializer = self.initializer
        self.kernel = self.add_weight(name="kernel", shape=[self._n_terms * Fin, Fout], initializer=initializer,
                                      **self.kwargs)
        if self.use_bias:
            # Keras' add_weight default initialiser (glorot_uniform) — gnn_layers.py:104
            self.bias = self.add_weight(name="bias", shape=[1, 1, Fout])
        if self.use_bn:
            self.bn.build_from_shape((input_shape[0], input_shape[1], Fout))

    def compute_output_shape(self, input_shape):
        return (input_shape[0], input_shape[1], int(input_shape[-1]) if self.Fout is None else int(self.Fout))

    def call(self, input_tensor, training=False):
        """gnn_layers.py:106-161."""
        if input_tensor.dim() != 3:
            raise ValueError(f"expected input of shape (batch, nodes, channels), got {tuple(input_tensor.shape)}")
        N, M, Fin = input_tensor.shape
        if M != self._plan.M:
            raise ValueError(f"input has {M} nodes but the graph Laplacian is {self._plan.M} x {self._plan.M}")
        if self.kernel.shape[0] != self._n_terms * Fin:
            raise ValueError(f"layer was built for {self.kernel.shape[0] // self._n_terms} input channels, got {Fin}")
        mode = self.mode if self.mode is not None else _default_mode()
        if isinstance(mode, str):
            mode = nat.MODES[mode]
        bias = self.bias if self.use_bias else None
        # bias + activation fuse into the contraction epilogue unless BatchNorm sits between
        # them (gnn_layers.py:152-159) or the activation is an arbitrary callable
        fuse = (not self.use_bn) and self._act_id is not None
        x = _ops.graph_conv(
            input_tensor, self._device_kernel(), bias if fuse else None, self._plan, self._recursion, self._n_terms,
            self._act_id if fuse else nat.ACT_LINEAR, mode,
        )
        if fuse:
            return x
        if self.use_bn and self._act_id is not None and x.is_cuda:
            # BatchNormalization(center=False, scale=False) -> + bias -> activation as C-ABI kernels (ds_bn_*): one
            # statistics pass and one normalise + bias + activation pass, no framework kernels in between; with a process
            # group the 2F + 1 sums are all-reduced (statistics of the global batch / of the whole partitioned sphere)
            return _ops.bn_bias_act(x, bias, self.bn, self._act_id, training, rows=self._bn_rows, sync_group=self._bn_sync)
        if self.use_bn:
            if self._bn_rows is not None:
                raise nat.NativeError("BatchNormalization over a row range (sphere-partitioned layer) runs in the CUDA "
                                      "kernels only (ds_bn_*): CUDA tensors and a named activation are required")
            self.bn.sync_group = self._bn_sync
            x = self.bn(x, training=training)
        if self._act_id is not None:
            return _ops.bias_act(x, bias, self._act_id)
        if bias is not None:
            x = _ops.bias_act(x, bias, nat.ACT_LINEAR)
        return self.activation(x)


class Chebyshev(_GraphConvBase):
    """A graph convolutional layer using the Chebyshev approximation (gnn_layers.py:12-161):
    T_0 = x, T_1 = L~ x, T_k = 2 L~ T_{k-1} - T_{k-2}; L~ = 1.5 L / lmax - I."""

    _recursion = nat.RECURSION_CHEBYSHEV
    _scale = 0.75  # gnn_layers.py:67

    def _default_initializer(self, Fin, Fout):
        stddev = 1 / np.sqrt(Fin * (self.K + 0.5) / 2)  # gnn_layers.py:92
        return TruncatedNormal(stddev=stddev)


class Monomial(_GraphConvBase):
    """A graph convolutional layer using Monomials (gnn_layers.py:164-309):
    T_k = L~ T_{k-1}; L~ = 2 L / lmax - I."""

    _recursion = nat.RECURSION_MONOMIAL
    _scale = 1.0  # gnn_layers.py:219 (rescale_L default scale)

    def _default_initializer(self, Fin, Fout):
        return TruncatedNormal(stddev=0.1)  # gnn_layers.py:243


def bernstein_polynomials(K, stale_last_term=True):
    """The K+1 basis polynomials of the reference's Bernstein layer as callables' values: returns
    ``p(lam)[i]`` evaluated on an array ``lam`` (gnn_layers.py:543-554).

    ``p_i(lam) = theta_i (2 - lam)^(K-i) lam^i`` with ``theta_i = C(K, i) / 2^K``.  In the reference
    the inner loop of the last term (i = K) is empty, so its ``x3`` is the already scaled tensor
    of i = K-1 and the last basis column is ``theta_K * theta_{K-1} (2 - lam) lam^(K-1)`` rather
    than ``theta_K lam^K``.  ``stale_last_term=True`` reproduces that (weights trained
    with the reference give the same outputs); False evaluates the textbook polynomial."""
    from math import comb

    def evaluate(lam):
        lam = np.asarray(lam, dtype=np.float64)
        out = np.empty((K + 1,) + lam.shape, dtype=np.float64)
        for i in range(K + 1):
            theta = comb(K, i) / 2.0**K
            out[i] = theta * (2.0 - lam) ** (K - i) * lam**i
        if stale_last_term:
            out[K] = (1.0 / 2.0**K) * out[K - 1]
        return out

    return evaluate


def bernstein_to_chebyshev(K, stale_last_term=True):
    """``C[i, j]`` with ``p_i(lam) = sum_j C[i, j] T_j(lam)`` (float64, [K+1, K+1]).

    Computed by interpolation at the K+1 Chebyshev nodes (exact for degree-K polynomials and free
    of the cancellation a monomial expansion of ``(2 - lam)^(K-i)`` would bring): the entries are
    bounded by ``max |p_i|`` on [-1, 1] <= 1.5^K."""
    n = K + 1
    m = np.arange(n)
    nodes = np.cos(np.pi * (m + 0.5) / n)
    vals = bernstein_polynomials(K, stale_last_term)(nodes)                    # [K+1, n]
    dct = np.cos(np.pi * np.outer(np.arange(n), m + 0.5) / n)                  # [j, m]
    C = (2.0 / n) * vals @ dct.T
    C[:, 0] *= 0.5
    return C


class Bernstein(_GraphConvBase):
    """A graph convolutional layer using the Bernstein approximation (gnn_layers.py:416-572,
    https://arxiv.org/abs/2106.10994): ``y = sum_i p_i(L~) x W_i``, ``L~ = 1.5 L / lmax - I``, K = ORDER of
    the polynomial, kernel ``[(K+1)*Fin, Fout]`` with row order ``f*(K+1) + i``.

    The reference evaluates every ``p_i(L~) x`` separately, K(K+1)/2 + K sparse products per call.  All
    p_i are polynomials of degree <= K in the same L~ that the Chebyshev layer uses (scale 0.75), so here the
    layer IS the (K+1)-term Chebyshev graph convolution (K hops, the fused sm_100a kernel for K <= 4) with the
    weights moved to the Chebyshev basis on the fly, ``W'_j = sum_i C[i, j] W_i`` (a [(K+1) x (K+1)] matrix
    from `bernstein_to_chebyshev`, applied to the small weight tensor inside autograd): same polynomial of L~,
    same result up to fp32 rounding, same trainable variable as the reference."""

    _recursion = nat.RECURSION_CHEBYSHEV
    _scale = 0.75  # gnn_layers.py:473

    def __init__(self, L, K, Fout=None, initializer=None, activation=None, use_bias=False, use_bn=False,
                 n_matmul_splits=1, **kwargs):
        self.stale_last_term = bool(kwargs.pop("stale_last_term", True))
        super().__init__(L, K, Fout=Fout, initializer=initializer, activation=activation, use_bias=use_bias,
                         use_bn=use_bn, n_matmul_splits=n_matmul_splits, **kwargs)
        # (K = 0 raises NameError in the reference: its x3 is never assigned; the base class rejects K < 1)
        # a constant of the layer, not a Keras weight: kept out of parameters()/buffers()
        self._basis_change = bernstein_to_chebyshev(self.K, self.stale_last_term)
        self._basis_change_dev = {}

    @property
    def _n_terms(self):
        return self.K + 1

    def _default_initializer(self, Fin, Fout):
        return TruncatedNormal(stddev=np.sqrt(6 / (Fin + Fout)))  # gnn_layers.py:498-499

    def _device_kernel(self):
        n = self.K + 1
        Fout = self.kernel.shape[1]
        dev = self.kernel.device
        C = self._basis_change_dev.get(dev)
        if C is None:
            C = self._basis_change_dev[dev] = torch.tensor(self._basis_change, dtype=torch.float32, device=dev)
        return torch.einsum("ij,fio->fjo", C, self.kernel.reshape(-1, n, Fout)).reshape(-1, Fout)


class GCNN_ResidualLayer(Model):
    """in -> layer -> [norm] -> layer -> [norm] -> out + alpha * in, with the activation either
    before or after the skip connection (reference gnn_layers.py:312-413)."""

    def __init__(self, layer_type, layer_kwargs, activation=None, act_before=False, use_bn=False,
                 norm_type="batch_norm", bn_kwargs=None, alpha=1.0):
        super().__init__()
        self.layer_type = layer_type
        self.layer_kwargs = layer_kwargs
        _, self.activation = resolve_activation(activation)  # ValueError for unknown names (:353)
        self.act_before = act_before
        self.use_bn = use_bn
        self.norm_type = norm_type
        # default normalisation axis, gnn_layers.py:357-363
        if bn_kwargs is None:
            self.bn_kwargs = {"axis": -1}
        else:
            self.bn_kwargs = bn_kwargs
            if "axis" not in self.bn_kwargs and norm_type != "moving_norm":
                self.bn_kwargs.update({"axis": -1})
        # both sub-layers receive the same kwargs incl. the same L (gnn_layers.py:365-370)
        if layer_type == "CHEBY":
            self.layer1, self.layer2 = Chebyshev(**self.layer_kwargs), Chebyshev(**self.layer_kwargs)
        elif layer_type == "MONO":
            self.layer1, self.layer2 = Monomial(**self.layer_kwargs), Monomial(**self.layer_kwargs)
        else:
            raise IOError(f"Layertype not understood: {self.layer_type}")
        if use_bn:
            if norm_type == "layer_norm":
                self.bn1, self.bn2 = LayerNormalization(**self.bn_kwargs), LayerNormalization(**self.bn_kwargs)
            elif norm_type == "batch_norm":
                self.bn1, self.bn2 = BatchNormalization(**self.bn_kwargs), BatchNormalization(**self.bn_kwargs)
            else:
                raise ValueError(f"norm_type <{norm_type}> not understood!")
        self.alpha = alpha

    def _norm(self, bn, x, training):
        return bn(x, training=training) if isinstance(bn, BatchNormalization) else bn(x)

    def build(self, input_shape):
        shape = self.layer1.build_from_shape(input_shape)
        if self.use_bn:
            self.bn1.build_from_shape(shape)
        shape = self.layer2.build_from_shape(shape)
        if self.use_bn:
            self.bn2.build_from_shape(shape)

    def call(self, input_tensor, training=False):
        """gnn_layers.py:384-413.  The reference calls its sub-layers without an explicit `training` (:391, :398); in
        tf.keras a layer called inside another layer's `call` inherits the training flag of the enclosing call context,
        so a sub-layer built with use_bn=True trains its BatchNormalization when the residual layer is trained — the
        flag is passed on explicitly here."""
        x = self.layer1(input_tensor, training=training)
        if self.use_bn:
            x = self._norm(self.bn1, x, training)
        x = self.layer2(x, training=training)
        if self.use_bn:
            x = self._norm(self.bn2, x, training)
        if self.activation is None:
            return x + input_tensor  # alpha is ignored here in the reference too (:407-408)
        if self.act_before:
            return self.activation(x) + self.alpha * input_tensor
        return self.activation(x + self.alpha * input_tensor)
