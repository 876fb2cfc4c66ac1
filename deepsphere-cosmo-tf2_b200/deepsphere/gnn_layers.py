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
from scipy.sparse.linalg import eigsh

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
        # largest eigenvalue of L supplied by the caller (skips ARPACK): the sphere-partitioned layer passes the
        # GLOBAL value so that every rank's restricted Laplacian is the restriction of the same L~
        lmax_given = kwargs.pop("lmax", None)
        self.kwargs = kwargs

        # gnn_layers.py:64-72: rescale the Laplacian and keep it as COO (indices, values, shape)
        Lc = _to_csr(L)
        lmax = float(lmax_given) if lmax_given is not None else \
            1.02 * eigsh(Lc, k=1, which="LM", return_eigenvectors=False)[0]
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

        if self.K < 2:
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
            # K <= 5: always a 4-ring halo (24 x 24 lattice) - the geometry of the register-resident fused kernel
            # (ds_lattice_conv2.cu); the kernels run K - 1 <= 4 hops on it
            return lattice.make_payload(Lt, nside, indices, max(self.K - 1, 4))
        except Exception as exc:  # never let the optional fast path break the layer
            logger.warning(f"lattice plan construction failed ({exc!r}); using the generic kernels")
            return None

    def _default_initializer(self, Fin):
        raise NotImplementedError

    def build(self, input_shape):
        """gnn_layers.py:74-104: kernel [K*Fin, Fout] (row order f*K + k), bias [1, 1, Fout]."""
        Fin = int(input_shape[-1])
        Fout = Fin if self.Fout is None else int(self.Fout)
        if self.initializer is None:
            initializer = self._default_initializer(Fin)
        else:
            logger.debug(self.kwargs)
            initializer = self.initializer
        self.kernel = self.add_weight(name="kernel", shape=[self.K * Fin, Fout], initializer=initializer,
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
        if self.kernel.shape[0] != self.K * Fin:
            raise ValueError(f"layer was built for {self.kernel.shape[0] // self.K} input channels, got {Fin}")
        mode = self.mode if self.mode is not None else _default_mode()
        if isinstance(mode, str):
            mode = nat.MODES[mode]
        bias = self.bias if self.use_bias else None
        # bias + activation fuse into the contraction epilogue unless BatchNorm sits between
        # them (gnn_layers.py:152-159) or the activation is an arbitrary callable
        fuse = (not self.use_bn) and self._act_id is not None
        x = _ops.graph_conv(
            input_tensor, self.kernel, bias if fuse else None, self._plan, self._recursion, self.K,
            self._act_id if fuse else nat.ACT_LINEAR, mode,
        )
        if fuse:
            return x
        if self.use_bn:
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

    def _default_initializer(self, Fin):
        stddev = 1 / np.sqrt(Fin * (self.K + 0.5) / 2)  # gnn_layers.py:92
        return TruncatedNormal(stddev=stddev)


class Monomial(_GraphConvBase):
    """A graph convolutional layer using Monomials (gnn_layers.py:164-309):
    T_k = L~ T_{k-1}; L~ = 2 L / lmax - I."""

    _recursion = nat.RECURSION_MONOMIAL
    _scale = 1.0  # gnn_layers.py:219 (rescale_L default scale)

    def _default_initializer(self, Fin):
        return TruncatedNormal(stddev=0.1)  # gnn_layers.py:243


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
        """gnn_layers.py:384-413 (the sub-layers are called without `training`, as there)."""
        x = self.layer1(input_tensor)
        if self.use_bn:
            x = self._norm(self.bn1, x, training)
        x = self.layer2(x)
        if self.use_bn:
            x = self._norm(self.bn2, x, training)
        if self.activation is None:
            return x + input_tensor  # alpha is ignored here in the reference too (:407-408)
        if self.act_before:
            return self.activation(x) + self.alpha * input_tensor
        return self.activation(x + self.alpha * input_tensor)
