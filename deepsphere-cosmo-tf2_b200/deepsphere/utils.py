"""Utilities module — mirrors reference src/deepsphere/utils.py (names, arguments, results)."""

import numpy as np
from scipy import sparse

from . import _ops
from . import healpix as hpx
from ._native import GraphPlan


def extend_indices(indices, nside_in, nside_out, nest=True):
    """Minimally extends a set of pixel ids so that it can be reduced to nside_out the healpy
    way (four pixels always merge into their parent).  Reference utils.py:9-37, where this
    is done by hp.ud_grade-ing a 0/1 mask down and up and thresholding at 1e-12; in NESTED
    order that is exactly "take every child of every touched parent", pure integer work.
    Returns the ids in the same ordering scheme as the input, sorted."""
    indices = np.asarray(indices, dtype=np.int64)
    if not (hpx.isnsideok(nside_in) and hpx.isnsideok(nside_out)) or nside_out > nside_in:
        raise ValueError(f"cannot extend indices from nside {nside_in} to nside {nside_out}")
    if not nest:
        indices = hpx.ring2nest(nside_in, indices)
    p = hpx.nside2order(nside_in) - hpx.nside2order(nside_out)
    out = hpx.refine_indices(hpx.coarsen_indices(indices, p), p)
    if not nest:
        out = np.sort(hpx.nest2ring(nside_in, out))
    return out


def rescale_L(L, lmax=2, scale=1):
    """Rescale the Laplacian eigenvalues in [-scale,scale] (reference utils.py:40-46):
    ``L * (2*scale/lmax) - I``.  Unlike the reference the caller's matrix is
    never modified."""
    L = sparse.csr_matrix(L, copy=True)
    M, _ = L.shape
    identity = sparse.identity(M, format="csr", dtype=L.dtype)
    L = L * (2 * scale / lmax)
    L = L - identity
    return sparse.csr_matrix(L)


def plan_from_sparse(L_tilde, ell_width=0):
    """Device plan (ELL + CSR tail of L~ and L~^T) from a scipy sparse matrix."""
    coo = sparse.coo_matrix(L_tilde)
    indices = np.column_stack((coo.row, coo.col)).astype(np.int64)
    return GraphPlan(indices, coo.data.astype(np.float32), coo.shape, ell_width)


def split_sparse_dense_matmul(sparse_tensor, dense_tensor, n_splits=1):
    """``sparse_tensor @ dense_tensor`` (reference utils.py:49-78).

    sparse_tensor: a GraphPlan (or a scipy sparse matrix, converted on the fly);
    dense_tensor: CUDA tensor [M, C].  ``n_splits`` exists in the reference only to dodge
    TF-GPU's ``output.shape[1] * nnz <= 2^31`` limit (utils.py:59); the sm_100a kernel uses
    64-bit addressing, so the argument is accepted and ignored."""
    plan = sparse_tensor if isinstance(sparse_tensor, GraphPlan) else plan_from_sparse(sparse_tensor)
    x = dense_tensor
    if x.dim() != 2 or x.shape[0] != plan.M:
        raise ValueError(f"dense_tensor must be [M={plan.M}, C], got {tuple(x.shape)}")
    return _ops.spmm(plan, x.unsqueeze(0)).squeeze(0)
