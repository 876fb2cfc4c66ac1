"""HEALPix-aware layers — same names, arguments, weights and errors as the reference's
src/deepsphere/healpy_layers.py:20-378, executed by the sm_100a C-ABI library.

In NESTED ordering the 4^p children of a pixel p levels up are contiguous
(healpy_layers.py:22-23,89-90), so pooling is a reshape-reduce and the pseudo convolutions
are plain GEMMs over [B*M/4^p, 4^p*Fin]; the layers only check divisibility, never the
ordering (healpy_layers.py:29-31), exactly like the reference.
"""

import numpy as np

from . import _native as nat
from . import _ops
from .gnn_layers import Chebyshev, GCNN_ResidualLayer, Monomial, _default_mode
from .keras_compat import Model, get_initializer, resolve_activation

# the reference sets this at import time (healpy_layers.py:17)
np.set_printoptions(precision=1)


class HealpyPool(Model):
    """Pooling over the 4^p nested children (healpy_layers.py:20-84)."""

    def __init__(self, p, pool_type="MAX", **kwargs):
        super().__init__()
        if not p >= 1:
            raise IOError("The reduction factors has to be at least 2!")  # healpy_layers.py:39-40
        self.p = p
        self.filter_size = int(4**p)
        self.pool_type = pool_type
        self.kwargs = kwargs
        if pool_type == "MAX":
            self._type = nat.POOL_MAX
        elif pool_type == "AVG":
            self._type = nat.POOL_AVG
        else:
            raise IOError(f"Pooling type not understood: {self.pool_type}")  # healpy_layers.py:64-65

    def build(self, input_shape):
        n_nodes = int(input_shape[1])
        if n_nodes % self.filter_size != 0:
            raise IOError(f"Input shape {input_shape} not compatible with the filter size {self.filter_size}")

    def compute_output_shape(self, input_shape):
        return (input_shape[0], int(input_shape[1]) // self.filter_size, input_shape[2])

    def call(self, input_tensor):
        self.build(input_tensor.shape)
        return _ops.pool(input_tensor, int(self.p), self._type)


class _PseudoConvBase(Model):
    _transpose = False

    def __init__(self, p, Fout, kernel_initializer=None, **kwargs):
        super().__init__()
        if not p >= 1:
            raise IOError("The reduction factors has to be at least 1!" if not self._transpose
                          else "The boost factors has to be at least 1!")
        self.p = p
        self.filter_size = int(4**p)
        self.Fout = int(Fout)
        self.kernel_initializer = kernel_initializer
        self.kwargs = dict(kwargs)
        # the Keras conv kwargs the reference's users pass on (healpy_layers.py:125,187)
        kw = dict(kwargs)
        self._act_id, self.activation = resolve_activation(kw.pop("activation", None))
        self.use_bias = kw.pop("use_bias", True)
        self.bias_initializer = kw.pop("bias_initializer", "zeros")
        self._weight_kwargs = {}
        if "kernel_regularizer" in kw:
            self._weight_kwargs["regularizer"] = kw.pop("kernel_regularizer")
        self.mode = kw.pop("mode", None)
        kw.pop("name", None)
        if kw:
            raise TypeError(f"unsupported keyword arguments for {type(self).__name__}: {sorted(kw)}")

    def _kernel_shape(self, Fin):
        raise NotImplementedError

    def build(self, input_shape):
        n_nodes = int(input_shape[1])
        if n_nodes % self.filter_size != 0:
            raise IOError(f"Input shape {input_shape} not compatible with the filter size {self.filter_size}")
        Fin = int(input_shape[-1])
        # Keras conv defaults: glorot_uniform kernel, zeros bias
        self.kernel = self.add_weight("kernel", self._kernel_shape(Fin), get_initializer(self.kernel_initializer),
                                      **self._weight_kwargs)
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.Fout], self.bias_initializer)

    def compute_output_shape(self, input_shape):
        M = int(input_shape[1])
        return (input_shape[0], M * self.filter_size if self._transpose else M // self.filter_size, self.Fout)

    def call(self, input_tensor):
        if input_tensor.shape[1] % self.filter_size != 0 and not self._transpose:
            raise IOError(f"Input shape {tuple(input_tensor.shape)} not compatible with the filter size "
                          f"{self.filter_size}")
        mode = self.mode if self.mode is not None else _default_mode()
        if isinstance(mode, str):
            mode = nat.MODES[mode]
        fused = self._act_id is not None
        y = _ops.pseudo_conv(
            input_tensor, self.kernel, self.bias if self.use_bias else None, int(self.p), self.Fout,
            self._act_id if fused else nat.ACT_LINEAR, mode, self._transpose,
        )
        return y if fused else self.activation(y)


class HealpyPseudoConv(_PseudoConvBase):
    """Conv1D(Fout, kernel = stride = 4^p, 'valid') over the nested children
    (healpy_layers.py:87-146).  Weights: kernel [4^p, Fin, Fout], bias [Fout]."""

    _transpose = False

    def _kernel_shape(self, Fin):
        return [self.filter_size, Fin, self.Fout]


class HealpyPseudoConv_Transpose(_PseudoConvBase):
    """Conv2DTranspose(Fout, (1, 4^p), strides (1, 4^p)) (healpy_layers.py:149-216).
    Weights: kernel [1, 4^p, Fout, Fin], bias [Fout]."""

    _transpose = True

    def _kernel_shape(self, Fin):
        return [1, self.filter_size, self.Fout, Fin]


class HealpyChebyshev:
    """Deferred factory for a Chebyshev layer (healpy_layers.py:219-264): stores the
    arguments until HealpyGCNN knows the graph Laplacian."""

    def __init__(self, K, Fout=None, initializer=None, activation=None, use_bias=False, use_bn=False, **kwargs):
        self.K = K
        self.Fout = Fout
        self.initializer = initializer
        self.activation = activation
        self.use_bias = use_bias
        self.use_bn = use_bn
        self.kwargs = kwargs

    def _get_layer(self, L, n_matmul_splits=1):
        return Chebyshev(L=L, K=self.K, Fout=self.Fout, initializer=self.initializer, activation=self.activation,
                         use_bias=self.use_bias, use_bn=self.use_bn, n_matmul_splits=n_matmul_splits, **self.kwargs)


class HealpyMonomial:
    """Deferred factory for a Monomial layer (healpy_layers.py:267-313)."""

    def __init__(self, K, Fout=None, initializer=None, activation=None, use_bias=False, use_bn=False, **kwargs):
        self.K = K
        self.Fout = Fout
        self.initializer = initializer
        self.activation = activation
        self.use_bias = use_bias
        self.use_bn = use_bn
        self.kwargs = kwargs

    def _get_layer(self, L, n_matmul_splits=1):
        return Monomial(L=L, K=self.K, Fout=self.Fout, initializer=self.initializer, activation=self.activation,
                        use_bias=self.use_bias, use_bn=self.use_bn, n_matmul_splits=n_matmul_splits, **self.kwargs)


class Healpy_ResidualLayer:
    """Deferred factory for GCNN_ResidualLayer (healpy_layers.py:316-378).  The user's
    layer_kwargs dict is copied, not mutated (the reference updates it in place, :366-367)."""

    def __init__(self, layer_type, layer_kwargs, activation=None, act_before=False, use_bn=False,
                 norm_type="batch_norm", bn_kwargs=None, alpha=1.0):
        self.layer_type = layer_type
        self.layer_kwargs = layer_kwargs
        self.activation = activation
        self.act_before = act_before
        self.use_bn = use_bn
        self.norm_type = norm_type
        self.bn_kwargs = bn_kwargs
        self.alpha = alpha

    def _get_layer(self, L, n_matmul_splits=1):
        layer_kwargs = dict(self.layer_kwargs)
        layer_kwargs.update({"L": L, "n_matmul_splits": n_matmul_splits})
        return GCNN_ResidualLayer(layer_type=self.layer_type, layer_kwargs=layer_kwargs, activation=self.activation,
                                  act_before=self.act_before, use_bn=self.use_bn, norm_type=self.norm_type,
                                  bn_kwargs=self.bn_kwargs, alpha=self.alpha)
