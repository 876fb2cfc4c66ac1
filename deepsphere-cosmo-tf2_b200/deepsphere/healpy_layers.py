"""HEALPix-aware layers — same names, arguments, weights and errors as the reference's
src/deepsphere/healpy_layers.py:20-378 and :462-853, executed by the sm_100a C-ABI library.

In NESTED ordering the 4^p children of a pixel p levels up are contiguous
(healpy_layers.py:22-23,89-90), so pooling is a reshape-reduce and the pseudo convolutions
are plain GEMMs over [B*M/4^p, 4^p*Fin]; the layers only check divisibility, never the
ordering (healpy_layers.py:29-31), exactly like the reference.
"""

import os

import numpy as np
import torch
from scipy import sparse
from scipy.spatial import cKDTree

from . import _native as nat
from . import _ops
from . import healpix as hpx
from . import logger
from .gnn_layers import Bernstein, Chebyshev, GCNN_ResidualLayer, Monomial, _default_mode
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


class HealpyBernstein:
    """Deferred factory for a Bernstein layer (healpy_layers.py:462-507); K is the polynomial order."""

    def __init__(self, K, Fout=None, initializer=None, activation=None, use_bias=False, use_bn=False, **kwargs):
        self.K = K
        self.Fout = Fout
        self.initializer = initializer
        self.activation = activation
        self.use_bias = use_bias
        self.use_bn = use_bn
        self.kwargs = kwargs

    def _get_layer(self, L, n_matmul_splits=1):
        return Bernstein(L=L, K=self.K, Fout=self.Fout, initializer=self.initializer, activation=self.activation,
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


class HealpySmoothing(Model):
    """A layer that smoothes a HEALPix map with a Gaussian kernel (healpy_layers.py:510-853): same
    constructor, attributes, on-disk cache files (``ind_coo-nside..-sigma..-n_sigma...npy`` / ``val_coo...``)
    and arithmetic.  The sparse kernel is applied by the sm_100a SpMM in the layer's native
    [n_batch, n_indices, n_channels] layout: ONE launch smoothes all channels (the reference unstacks the channels
    and runs one TF SpMM each, :732-754); with per-channel repetitions, launch r only keeps its result for the
    channels that need at least r passes.

    Neighbour search: the reference uses a scikit-learn BallTree with the haversine metric on (lat, lon)
    (:766-799); here a scipy cKDTree on the pixels' unit vectors — chord length is monotone in the great-circle
    distance, so the k nearest pixels are the same (up to exact ties at the cut) and the distances are converted
    back with ``2 asin(chord / 2)``."""

    def __init__(self, nside, indices, nest=True, mask=None, fwhm=None, sigma=None, n_sigma_support=3, arcmin=True,
                 per_channel_repetitions=None, data_path=None, max_batch_size=None):
        super().__init__()
        self.nside = nside
        self.indices = indices
        self.nest = nest
        self.mask = mask

        assert fwhm is not None or sigma is not None, "One of fwhm and sigma has to be specified"
        assert fwhm is None or sigma is None, "Only one of fwhm and sigma can be specified"
        self.fwhm = fwhm
        self.sigma = sigma
        self.n_sigma_support = n_sigma_support
        self.arcmin = arcmin
        self.per_channel_repetitions = per_channel_repetitions
        self.data_path = data_path
        self.max_batch_size = max_batch_size

        if np.ndim(self.fwhm) == 0 and self.fwhm == 0.0 or np.ndim(self.sigma) == 0 and self.sigma == 0.0:
            self.do_smoothing = False
            logger.info("The layer implements the identity, smoothing is disabled")
            return
        self.do_smoothing = True

        # a list of scales: the kernel is built for the smallest one, wider channels repeat it
        # ceil((s / s_min)^2) times since Gaussian variances add (healpy_layers.py:595-622)
        if isinstance(self.fwhm, (list, np.ndarray)):
            assert self.per_channel_repetitions is None, \
                "per_channel_repetitions can't be specified when fwhm is a list, since it is then inferred"
            self.fwhm = np.array(self.fwhm)
            fwhm_min = np.min(self.fwhm)
            self.per_channel_repetitions = np.ceil((self.fwhm / fwhm_min) ** 2).astype(int)
            self.fwhm = fwhm_min
        elif isinstance(self.sigma, (list, np.ndarray)):
            assert self.per_channel_repetitions is None, \
                "per_channel_repetitions can't be specified when sigma is a list, since it is then inferred"
            self.sigma = np.array(self.sigma)
            sigma_min = np.min(self.sigma)
            self.per_channel_repetitions = np.ceil((self.sigma / sigma_min) ** 2).astype(int)
            self.sigma = sigma_min
        elif isinstance(self.per_channel_repetitions, list):
            self.per_channel_repetitions = np.array(self.per_channel_repetitions)

        if self.sigma is None:
            self.sigma = self.fwhm / np.sqrt(8 * np.log(2))
        if self.arcmin:
            self.sigma_arcmin = self.sigma
            self.sigma_rad = self._arcmin_to_rad(self.sigma_arcmin)
        else:
            self.sigma_rad = self.sigma
            self.sigma_arcmin = self._rad_to_arcmin(self.sigma_rad)
        self.fwhm_arcmin = self.sigma_arcmin * np.sqrt(8 * np.log(2))

        self.n_indices = len(indices)
        self.kernel_func = lambda r: np.exp(-0.5 / self.sigma_rad**2 * r**2)
        self.file_label = f"-nside{self.nside}-sigma{self.sigma_arcmin:4.2f}-n_sigma{n_sigma_support}"

        if self.per_channel_repetitions is not None:
            per_channel_factor = np.sqrt(self.per_channel_repetitions)
            logger.info(f"Using the per channel smoothing repetitions {self.per_channel_repetitions}")
            logger.info("Using the per channel smoothing scales "
                        f"sigma = {per_channel_factor * self.sigma_arcmin} arcmin, "
                        f"fwhm = {per_channel_factor * self.fwhm_arcmin} arcmin")
        else:
            logger.info(f"Using the per channel smoothing scale sigma = {self.sigma_arcmin:4.2f} arcmin, "
                        f" fwhm = {self.fwhm_arcmin:4.2f} arcmin")

        loaded = False
        if self.data_path is not None:
            try:
                self.ind_coo = np.load(os.path.join(self.data_path, f"ind_coo{self.file_label}.npy"))
                self.val_coo = np.load(os.path.join(self.data_path, f"val_coo{self.file_label}.npy"))
                logger.info(f"Successfully loaded sparse kernel indices and values from {self.data_path}")
                loaded = True
            except FileNotFoundError:
                pass
        if not loaded:
            self._build_tree()
            self._build_kernel()
        self._build_sparse_tensor()
        logger.info("Successfully created the sparse kernel tensor")

    def build(self, input_shape):
        """Shape checks of healpy_layers.py:675-723.  ``n_matmul_splits`` is computed like there for API
        compatibility; the SpMM kernel has no 2^31 limit and never splits."""
        if not self.do_smoothing:
            return
        if self.max_batch_size is not None:
            self.n_batch = self.max_batch_size
        elif input_shape[0] is not None:
            self.n_batch = input_shape[0]
        else:
            self.n_batch = None
        assert self.n_indices == input_shape[1]
        self.n_channels = input_shape[2]
        if self.per_channel_repetitions is not None:
            assert len(self.per_channel_repetitions) == self.n_channels, \
                f"The list per_channel_repetitions has to have length {self.n_channels}"
            assert np.issubdtype(self.per_channel_repetitions.dtype, np.integer), \
                "The list per_channel_repetitions has to contain integers only"
        if self.mask is not None:
            mask = torch.as_tensor(np.asarray(self.mask.detach().cpu() if isinstance(self.mask, torch.Tensor)
                                              else self.mask)).to(torch.float32)
            if mask.dim() == 1:
                mask = mask[None, :, None]
            elif mask.dim() == 2:
                mask = mask[None]
            assert mask.shape[1] == self.n_indices, \
                "The mask has to have shape (1, n_indices, 1) or (1, n_indices, n_channels)"
            self.mask = mask
        self.n_matmul_splits = 1
        if self.n_batch is not None:
            while not ((self.n_batch % self.n_matmul_splits == 0)
                       and (self.n_matmul_splits >= self.n_batch * self._nnz / 2**31)):
                self.n_matmul_splits += 1
        logger.info("Successfully built the smoothing layer")

    def compute_output_shape(self, input_shape):
        return tuple(input_shape)

    def call(self, inputs):
        """healpy_layers.py:725-764."""
        if not self.do_smoothing:
            return inputs
        x = inputs
        if self.per_channel_repetitions is None:
            x = _ops.sparse_matmul(self.sparse_kernel, x)
        else:
            cache = self.__dict__.setdefault("_dev_cache", {})  # constants of the layer, uploaded once per device
            reps = cache.get(("reps", x.device))
            if reps is None:
                reps = cache[("reps", x.device)] = torch.as_tensor(self.per_channel_repetitions, device=x.device)
            for r in range(1, int(self.per_channel_repetitions.max()) + 1):
                x = torch.where((reps >= r)[None, None, :], _ops.sparse_matmul(self.sparse_kernel, x), x)
        if self.mask is not None:
            cache = self.__dict__.setdefault("_dev_cache", {})
            mask = cache.get(("mask", x.device))
            if mask is None:
                mask = cache[("mask", x.device)] = self.mask.to(x.device)
            x = x * mask
        return x

    def _build_tree(self):
        """Per pixel, the ``max_neighbors`` nearest pixels (itself included) and the Gaussian of their great-circle
        distances, where max_neighbors is the largest number of pixels any pixel has within
        n_sigma_support * sigma (healpy_layers.py:766-799)."""
        logger.info(f"Creating tree for {self.n_indices} pixels and radius n_sigma_support * sigma = "
                    f"{self.sigma_arcmin * self.n_sigma_support:4.2f} arcmin")
        vec = np.ascontiguousarray(hpx.pix2vec(self.nside, np.asarray(self.indices, dtype=np.int64), nest=self.nest))
        tree = cKDTree(vec)
        radius = self.sigma_rad * self.n_sigma_support
        chord = 2.0 * np.sin(0.5 * min(radius, np.pi))
        counts = tree.query_ball_point(vec, r=chord * (1 + 1e-12), return_length=True)
        self.max_neighbors = int(np.max(counts))
        logger.info(f"The maximal number of neighbors within that radius is {self.max_neighbors}")
        k = min(self.max_neighbors, self.n_indices)
        dist_c, inds_k = tree.query(vec, k=k)
        dist_c, inds_k = dist_c.reshape(self.n_indices, k), inds_k.reshape(self.n_indices, k)
        dist_k = 2.0 * np.arcsin(np.clip(0.5 * dist_c, 0.0, 1.0))
        self.inds_k = inds_k.astype(np.int64)
        self.kernel_k = self.kernel_func(dist_k).astype(np.float32)

    def _build_kernel(self):
        """COO indices [nnz, 2] int64 and values [nnz] float32, optionally stored (healpy_layers.py:801-829)."""
        k = self.inds_k.shape[1]
        inds_r = np.repeat(np.arange(self.n_indices, dtype=np.int64)[:, None], k, axis=1)
        self.ind_coo = np.concatenate([inds_r.reshape(-1, 1), self.inds_k.reshape(-1, 1)], axis=1)
        self.val_coo = self.kernel_k.reshape(-1)
        if self.data_path is not None:
            logger.info(f"Storing sparse kernel indices ({self.ind_coo.nbytes/1e9:4.2f} GB, dtype {self.ind_coo.dtype}) "
                        f"and values ({self.val_coo.nbytes/1e9:4.2f} GB, dtype {self.val_coo.dtype})")
            os.makedirs(self.data_path, exist_ok=True)
            np.save(os.path.join(self.data_path, f"ind_coo{self.file_label}.npy"), self.ind_coo)
            np.save(os.path.join(self.data_path, f"val_coo{self.file_label}.npy"), self.val_coo)

    def _build_sparse_tensor(self):
        """healpy_layers.py:831-846.  The reference divides by ``expand_dims(row_sums, axis=0)``, shape [1, n], which
        broadcasts along the column axis: entry (i, j) is divided by the row sum of row j (SURVEY A.3-style quirk;
        identical to a row normalisation only for a symmetric kernel).  Reproduced as written."""
        n = self.n_indices
        Ks = sparse.csr_matrix((np.asarray(self.val_coo, dtype=np.float32),
                                (self.ind_coo[:, 0], self.ind_coo[:, 1])), shape=(n, n))
        row_sum = np.asarray(Ks.sum(axis=1)).ravel().astype(np.float32)
        coo = Ks.tocoo()
        vals = (coo.data / row_sum[coo.col]).astype(np.float32)
        self._nnz = int(coo.nnz)
        self.sparse_kernel = nat.GraphPlan(np.column_stack((coo.row, coo.col)).astype(np.int64), vals, (n, n))
        del self.ind_coo
        del self.val_coo

    @staticmethod
    def _rad_to_arcmin(theta):
        return theta / np.pi * (180 * 60)

    @staticmethod
    def _arcmin_to_rad(theta):
        return theta * np.pi / (60 * 180)
