"""HEALPix graph / Laplacian builder (host side, float64, scipy sparse).

Replaces the reference's call into PyGSP,
``SphereHealpix(subdivisions=nside, indexes=indices, nest=True, k=n_neighbors,
lap_type="normalized")`` followed by ``.L`` / ``.A`` (healpy_networks.py:110-118).
PyGSP (git branch ``jafluri/pygsp@sphere-graphs``, setup.cfg:21) is not vendored in the
reference and not installable here, so this is new code in the DeepSphere style, NOT a
bit-identical PyGSP clone (see DESIGN.md "Graph builder").  The hot-path contract is
"given the same L the outputs match": L is an *input* of ``Chebyshev(L=...)``
(gnn_layers.py:19,31).

Two neighbourhood rules:

* ``k == 8`` (default of HealpyGCNN, healpy_networks.py:19): the true HEALPix
  8-neighbourhood (7 at the 24 pixels around the valence-3 vertices).  Every row of L
  then has at most 9 entries -> fixed-width ELL with no tail on the full sphere.
* ``k in {20, 40, 60}``: k nearest neighbours in 3-D among the *selected* pixels
  (boundary pixels of a masked sky pick farther neighbours), symmetrised by union.

Weights are a Gaussian of the Euclidean distance between pixel centres,
``w = exp(-(d / kernel_width)^2)``; the default widths follow the per-(k, nside)
table DeepSphere uses (values proportional to 1/nside), extrapolated as 1/nside.
"""

import numpy as np
from scipy import sparse
from scipy.spatial import cKDTree

from . import healpix as hpx

# kernel widths at nside 32 for each supported k; width(nside) = width32 * 32 / nside
_KERNEL_WIDTH_32 = {8: 0.02500, 20: 0.03185, 40: 0.042432, 60: 0.051720}


def default_kernel_width(nside, k):
    if k not in _KERNEL_WIDTH_32:
        raise ValueError(f"no default kernel width for k={k}")
    return _KERNEL_WIDTH_32[k] * 32.0 / float(nside)


class SphereHealpix:
    """Graph over (a subset of) the HEALPix pixel centres.

    Attributes mirror what the reference reads from PyGSP: ``L`` (scipy CSR Laplacian),
    ``A`` (adjacency pattern with weights), ``W``, ``coords``, ``n_vertices``.
    """

    def __init__(self, subdivisions=2, indexes=None, nest=True, k=8, lap_type="normalized", kernel_width=None):
        nside = int(subdivisions)
        if not hpx.isnsideok(nside, nest=True):
            raise ValueError(f"nside {nside} is not valid")
        npix = hpx.nside2npix(nside)
        if indexes is None:
            indexes = np.arange(npix, dtype=np.int64)
        indexes = np.asarray(indexes, dtype=np.int64)
        if not nest:
            indexes = hpx.ring2nest(nside, indexes)
        self.subdivisions = nside
        self.indexes = indexes
        self.k = int(k)
        self.lap_type = lap_type
        self.kernel_width = default_kernel_width(nside, self.k) if kernel_width is None else float(kernel_width)
        self.n_vertices = M = len(indexes)
        self.coords = hpx.pix2vec(nside, indexes, nest=True)

        if self.k == 8:
            rows, cols = self._healpix_edges(nside, npix, indexes)
        else:
            rows, cols = self._knn_edges(self.coords, self.k)
        d2 = np.sum((self.coords[rows] - self.coords[cols]) ** 2, axis=1)
        w = np.exp(-d2 / self.kernel_width**2)
        W = sparse.csr_matrix((w, (rows, cols)), shape=(M, M))
        # symmetrise by union (weights depend on the distance only, so both directions agree)
        W = W.maximum(W.T).tocsr()
        W.sort_indices()
        self.W = W
        self.A = W
        self.L = self._laplacian(W, lap_type)

    @staticmethod
    def _healpix_edges(nside, npix, indexes):
        nb = hpx.neighbours(nside, indexes)  # [M, 8] global pixel ids
        M = len(indexes)
        full = len(indexes) == npix and np.array_equal(indexes, np.arange(npix))
        if full:
            local = nb
        else:
            lut = np.full(npix, -1, dtype=np.int64)
            lut[indexes] = np.arange(M, dtype=np.int64)
            local = np.where(nb >= 0, lut[np.where(nb >= 0, nb, 0)], -1)
        rows = np.repeat(np.arange(M, dtype=np.int64), 8)
        cols = local.ravel()
        keep = cols >= 0
        return rows[keep], cols[keep]

    @staticmethod
    def _knn_edges(coords, k):
        M = len(coords)
        kk = min(k + 1, M)
        tree = cKDTree(coords)
        _, idx = tree.query(coords, k=kk)
        rows = np.repeat(np.arange(M, dtype=np.int64), kk)
        cols = idx.ravel().astype(np.int64)
        keep = rows != cols
        return rows[keep], cols[keep]

    @staticmethod
    def _laplacian(W, lap_type):
        M = W.shape[0]
        d = np.asarray(W.sum(axis=1)).ravel()
        if lap_type == "combinatorial":
            L = sparse.diags(d) - W
        elif lap_type == "normalized":
            with np.errstate(divide="ignore"):
                dis = np.where(d > 0, 1.0 / np.sqrt(d), 0.0)
            Dm = sparse.diags(dis)
            L = sparse.identity(M, format="csr") - Dm @ W @ Dm
        else:
            raise ValueError(f"unknown lap_type {lap_type}")
        L = sparse.csr_matrix(L)
        L.sort_indices()
        return L
