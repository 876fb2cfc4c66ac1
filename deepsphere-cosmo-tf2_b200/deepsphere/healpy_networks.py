"""HealpyGCNN — the reference's Sequential model shell (src/deepsphere/healpy_networks.py:14-188)
over the B200 layers: same constructor, same validation and error types, same nside / index
bookkeeping through pooling layers; the graph comes from the in-repo builder
(deepsphere.graph.SphereHealpix) instead of PyGSP.
"""

import numpy as np

from . import gnn_layers as gnn
from . import healpix as hpx
from . import healpy_layers as hp_nn
from . import logger
from .graph import SphereHealpix
from .keras_compat import Sequential


class HealpyGCNN(Sequential):
    """A graph convolutional network using the Keras-style model API and the layers of this package."""

    def __init__(self, nside, indices, layers, n_neighbors=8, max_batch_size=None, initial_Fin=None,
                 graph_builder=None):
        """
        :param nside: integer, the nside of the input
        :param indices: 1d array of pixel ids (NESTED) of the input of the network
        :param layers: list of layers that make up the network
        :param n_neighbors: neighbours per pixel in the graph: 8 (default), 20, 40 or 60
        :param max_batch_size: kept for API compatibility: in the reference it sizes the column
                               splits of tf.sparse.sparse_dense_matmul (healpy_networks.py:125-134);
                               the splits are computed the same way and ignored by the kernels
        :param initial_Fin: initial number of input features (same remark)
        :param graph_builder: optional callable ``(nside, indices, n_neighbors) -> graph`` whose result has a sparse
                              Laplacian ``.L`` over ``indices`` (NESTED, same order).  NOT in the reference: it builds
                              every graph with the PyGSP fork's ``SphereHealpix`` (healpy_networks.py:110-118), which is
                              not a dependency here; ``deepsphere.graph.SphereHealpix`` is an independent builder whose
                              edge set and weights are not pinned against PyGSP, so weights trained with the
                              reference see a different ``L`` unless the original graphs are supplied, e.g.
                              ``graph_builder=lambda n, idx, k: pygsp.graphs.SphereHealpix(subdivisions=n, indexes=idx,
                              nest=True, k=k, lap_type="normalized")``.  The kernels are chosen from the matrix itself
                              (8-neighbour HEALPix pattern: fused lattice kernels; anything else: generic SpMM).
        """
        super().__init__(name="")
        logger.info("WARNING: This network assumes that everything concerning healpy is in NEST ordering...")

        if n_neighbors not in [8, 20, 40, 60]:
            raise NotImplementedError(
                f"The requested number of neighbors {n_neighbors} is nor supported. Choose either 8, 20, 40 or 60."
            )

        self.nside_in = nside
        self.indices_in = np.asarray(indices)
        self.layers_in = layers
        self.n_neighbors = n_neighbors

        # total reduction factor of the nside (healpy_networks.py:51-58)
        self.reduction_fac = 1.0
        for layer in self.layers_in:
            if isinstance(layer, (hp_nn.HealpyPool, hp_nn.HealpyPseudoConv)):
                self.reduction_fac *= 2 ** (layer.p)
            if isinstance(layer, hp_nn.HealpyPseudoConv_Transpose):
                self.reduction_fac /= 2 ** (layer.p)

        self.nside_out = int(self.nside_in // self.reduction_fac)
        if self.nside_out < 1:
            raise ValueError(
                "With the given input, the layers would reduce the nside below zero!"
                "Use less layers that reduce the nside, e.g. HealpyPool or HealpyPseudoConv..."
            )
        if not hpx.isnsideok(self.nside_out, nest=True):
            raise ValueError(f"The ouput of the network does not have a valid nside {self.nside_out}...")

        logger.info(
            f"Detected a reduction factor of {self.reduction_fac}, the input with nside {self.nside_in} will be "
            f"transformed to {self.nside_out} during a forward pass. Checking for consistency with indices..."
        )

        # the index set must survive going down to nside_out and back (healpy_networks.py:73-88)
        idx = self.indices_in.astype(np.int64)
        npix_in = hpx.nside2npix(self.nside_in)
        if idx.size == 0 or idx.min() < 0 or idx.max() >= npix_in:
            raise ValueError(f"indices must be pixel ids of a nside {self.nside_in} map")
        if self.nside_out <= self.nside_in:
            p = hpx.nside2order(self.nside_in) - hpx.nside2order(self.nside_out)
            transformed = hpx.refine_indices(hpx.coarsen_indices(idx, p), p)
        else:
            transformed = np.unique(idx)
        if not (len(transformed) == len(idx) and np.all(np.sort(transformed) == np.sort(idx))):
            raise ValueError(
                "With the given indices it would not be possible to properly reduce the input maps "
                "with the reduction factor determined by the layers. Use the function "
                "<extend_indices> from utils with the determined minimal nside to make your set of "
                "indices compatible..."
            )
        logger.info("indices seem consistent...")

        # build the actual layers (healpy_networks.py:90-167)
        self.layers_use = []
        current_nside = self.nside_in
        current_indices = idx
        current_Fin = initial_Fin
        graph_cache = {}

        for layer in self.layers_in:
            if isinstance(layer, (hp_nn.HealpyChebyshev, hp_nn.HealpyMonomial, hp_nn.Healpy_ResidualLayer,
                                  hp_nn.HealpyBernstein)):
                # the reference builds one SphereHealpix per graph layer even at equal nside
                # (healpy_networks.py:110); the result only depends on (nside, indices, k), so cache it
                key = (current_nside, len(current_indices), int(current_indices[0]), int(current_indices[-1]),
                       int(np.sum(current_indices, dtype=np.int64)))
                if key not in graph_cache:
                    if graph_builder is not None:
                        graph_cache[key] = graph_builder(current_nside, current_indices, self.n_neighbors)
                        L_user = getattr(graph_cache[key], "L", None)
                        if L_user is None or tuple(L_user.shape) != (len(current_indices),) * 2:
                            raise ValueError(
                                f"graph_builder must return an object with a Laplacian .L of shape "
                                f"({len(current_indices)}, {len(current_indices)}) for nside {current_nside}"
                            )
                    else:
                        graph_cache[key] = SphereHealpix(
                            subdivisions=current_nside, indexes=current_indices, nest=True, k=self.n_neighbors,
                            lap_type="normalized",
                        )
                sphere = graph_cache[key]
                current_L = sphere.L
                if (max_batch_size is not None) and (current_Fin is not None):
                    n_matmul_splits = 1
                    while not (
                        (max_batch_size * current_Fin % n_matmul_splits == 0)
                        and (n_matmul_splits >= max_batch_size * current_Fin * len(current_L.indices) / 2**31)
                    ):
                        n_matmul_splits += 1
                    actual_layer = layer._get_layer(current_L, n_matmul_splits)
                else:
                    actual_layer = layer._get_layer(current_L)
                # HEALPix geometry of this layer's graph (enables the fused lattice kernel)
                for sub in (actual_layer, getattr(actual_layer, "layer1", None), getattr(actual_layer, "layer2", None)):
                    if hasattr(sub, "_attach_healpix"):
                        sub._attach_healpix(current_nside, current_indices)
                self.layers_use.append(actual_layer)
            elif isinstance(layer, (hp_nn.HealpyPool, hp_nn.HealpyPseudoConv)):
                new_nside = int(current_nside // 2**layer.p)
                current_indices = self._transform_indices(current_nside, new_nside, current_indices)
                current_nside = new_nside
                self.layers_use.append(layer)
            elif isinstance(layer, hp_nn.HealpyPseudoConv_Transpose):
                new_nside = int(current_nside * 2**layer.p)
                current_indices = self._transform_indices(current_nside, new_nside, current_indices)
                current_nside = new_nside
                self.layers_use.append(layer)
            else:
                self.layers_use.append(layer)

            try:
                current_Fin = layer.Fout
            except AttributeError:
                # e.g. residual or pooling layers, which have Fin = Fout
                pass

        for layer in self.layers_use:
            self.add(layer)

    def _transform_indices(self, nside_in, nside_out, indices):
        """Index set at a new nside (healpy_networks.py:169-188; there via hp.ud_grade of a 0/1
        mask and a > 1e-12 threshold, which in NESTED is integer parent/child arithmetic)."""
        if nside_in == nside_out:
            return indices
        if nside_out < nside_in:
            return hpx.coarsen_indices(indices, hpx.nside2order(nside_in) - hpx.nside2order(nside_out))
        return hpx.refine_indices(indices, hpx.nside2order(nside_out) - hpx.nside2order(nside_in))

    def _get_filter_coeffs(self, layer, ind_in=None, ind_out=None):
        """Chebyshev filter coefficients of a layer as [K, Fout, Fin] (healpy_networks.py:190-212): the kernel
        is read as Fin x K x Fout — which is what pins the f*K + k row order — and transposed to K x Fout x Fin.
        ``ind_in`` / ``ind_out`` select input / output filters (tested for truthiness like there, so index 0 alone
        selects everything)."""
        K, Fout = layer.K, layer.Fout
        trained_weights = layer.kernel.detach().cpu().numpy()  # Fin*K x Fout
        if Fout is None:  # possible in res layers: Fin == Fout
            Fout = int(np.sqrt(np.prod(trained_weights.shape) // K))
        trained_weights = trained_weights.reshape((-1, K, Fout)).transpose([1, 2, 0])
        if ind_in:
            trained_weights = trained_weights[:, :, ind_in]
        if ind_out:
            trained_weights = trained_weights[:, ind_out, :]
        return trained_weights

    def get_gsp_filters(self, layer, ind_in=None, ind_out=None, return_weights=False):
        """The Chebyshev filters of a layer (healpy_networks.py:214-289).  ``layer`` is an index or a name;
        only Chebyshev layers and residual layers with Chebyshev sub-layers qualify (ValueError otherwise).
        ``return_weights=True`` gives the list of [K, Fout, Fin] coefficient arrays exactly like the reference.
        Otherwise the reference wraps them into ``pygsp.filters.Chebyshev`` objects on a full-sphere graph of
        the layer's nside; PyGSP is not a dependency here, so a `ChebyshevFilter` with the same ``evaluate(x)``
        contract (frequency response on [0, lmax]) is returned instead."""
        if isinstance(layer, (int, np.integer)):
            tf_layer = self.get_layer(index=int(layer))
        elif isinstance(layer, str):
            tf_layer = self.get_layer(name=layer)
        else:
            raise ValueError("layer should be either string or int.")
        msg = (f"The requested layer ({layer}) is of type {type(tf_layer)}, but only Chebyshev5 or "
               f"GCNN_ResidualLayer layers (with Chebyshev5 sublayers) are supported...")
        if isinstance(tf_layer, gnn.GCNN_ResidualLayer):
            if not (isinstance(tf_layer.layer1, gnn.Chebyshev) and isinstance(tf_layer.layer2, gnn.Chebyshev)):
                raise ValueError(msg)
            subs = [tf_layer.layer1, tf_layer.layer2]
        elif isinstance(tf_layer, gnn.Chebyshev):
            subs = [tf_layer]
        else:
            raise ValueError(msg)
        weights = [self._get_filter_coeffs(sub, ind_in=ind_in, ind_out=ind_out) for sub in subs]
        if return_weights:
            return weights
        return [ChebyshevFilter(w, sub.lmax) for w, sub in zip(weights, subs)]


class ChebyshevFilter:
    """Frequency response of a bank of Chebyshev filters, the part of ``pygsp.filters.Chebyshev`` that
    HealpyGCNN's plotting helpers use (healpy_networks.py:286, 312-329).  ``coefficients`` [K, Fout, Fin];
    ``lmax`` = the 1.02 * lambda_max the layer rescaled its Laplacian with.  The layer maps an eigenvalue lam
    of L to ``1.5 * lam / lmax - 1`` (gnn_layers.py:66-67) and its response is ``sum_k c_k T_k`` of that."""

    def __init__(self, coefficients, lmax):
        self.coefficients = np.asarray(coefficients, dtype=np.float64)
        self.lmax = float(lmax)
        self.n_filters = int(np.prod(self.coefficients.shape[1:]))

    def evaluate(self, x):
        """Response at graph frequencies ``x`` (eigenvalues of L): array [Fout, Fin, len(x)]."""
        lam = 1.5 * np.asarray(x, dtype=np.float64) / self.lmax - 1.0
        T = np.polynomial.chebyshev.chebvander(lam, self.coefficients.shape[0] - 1)  # [n, K]
        return np.einsum("nk,kof->ofn", T, self.coefficients)
