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


def largest_eigenvalue(L, tol=1e-9, maxit=20000, min_size=4096):
    """``eigsh(L, k=1, which="LM", return_eigenvectors=False)[0]`` — the eigenvalue of largest magnitude that the
    reference takes from ARPACK at machine tolerance (gnn_layers.py:66) — for a symmetric sparse L.

    ARPACK's restarted Lanczos with its default 20 basis vectors needs ~1 400 products with L at nside 256 and
    ~3 000 at nside 512 (33 s and 217 s on 8 cores, per layer, before the first batch can run); an un-restarted
    three-term Lanczos reaches the same Ritz value in ~230 steps (k = 20 graphs have a clustered top of the spectrum:
    ARPACK 8 700 products / 98 s at nside 128, here 7 s).  The stopping rule is the residual bound of the extreme Ritz
    pair, ``|beta_j s_j| <= tol |theta|``, which bounds the eigenvalue error by ``tol |theta|`` and in practice (the
    error is quadratic in the residual for a symmetric matrix) reproduces ARPACK's value to 1e-11 .. 1e-15; the
    rescaled Laplacian needs it to ~1e-8 for fp32 parity.  Falls back to ARPACK for small or unsymmetric
    matrices (where `which="LM"` semantics of the reference are whatever ARPACK does) and on non-convergence;
    ``DEEPSPHERE_LMAX=arpack`` forces the reference call."""
    import os

    from scipy.sparse.linalg import eigsh

    key = None
    if L.shape[0] >= min_size and os.environ.get("DEEPSPHERE_LMAX", "").lower() not in ("arpack", "nocache"):
        key = matrix_fingerprint(L)
        hit = _lmax_lookup(key)
        if hit is not None:
            return hit

    def arpack():
        return _lmax_store(key, float(eigsh(L, k=1, which="LM", return_eigenvectors=False)[0]))

    M = L.shape[0]
    if M < min_size or os.environ.get("DEEPSPHERE_LMAX", "").lower() == "arpack":
        return arpack()
    L = sparse.csr_matrix(L)
    if L.dtype != np.float64:
        L = L.astype(np.float64)
    asym = abs(L - L.T)
    if asym.nnz and asym.max() > 1e-12 * max(abs(L).max(), 1e-300):
        return arpack()
    theta = _lanczos_extreme(lambda v: L @ v, M, tol, min(maxit, M))
    return arpack() if theta is None else _lmax_store(key, theta)


# ---- memo of largest eigenvalues -----------------------------------------------------------------------------------
# The same Laplacians come back again and again: every graph layer of a network level, every rank of a sphere-partitioned
# model, every process of a benchmark.  Values are keyed by a fingerprint of the matrix CONTENT (never by nside / k), so
# a changed graph builder simply misses.  Three tiers: in-process dict, the table shipped with the package
# (_lmax_table.json: the full-sphere HEALPix graphs of graph.SphereHealpix, written by tools/make_lmax_table.py with
# this very function), and an optional directory DEEPSPHERE_CACHE_DIR shared by the processes of one box.
_LMAX_MEMO = {}
_LMAX_TABLE = None


def matrix_fingerprint(L):
    """Content fingerprint of a sparse matrix: shape, nnz, sum of |values| and of squares (both well conditioned; 10
    significant digits, so that a different summation order on another CPU gives the same key) and strided checksums
    of the column indices and row pointers."""
    L = sparse.csr_matrix(L)
    if not L.has_canonical_format:
        L = L.copy()
        L.sum_duplicates()
    d = np.asarray(L.data, dtype=np.float64)
    step = max(1, L.nnz // 4096)
    return (f"{L.shape[0]}x{L.shape[1]}:{L.nnz}:{float(np.abs(d).sum()):.9e}:{float((d * d).sum()):.9e}:"
            f"{int(np.asarray(L.indices[::step], dtype=np.int64).sum())}:{int(L.indptr[:: max(1, L.shape[0] // 1024)].sum())}")


def _lmax_lookup(key):
    global _LMAX_TABLE
    import json
    import os

    if key in _LMAX_MEMO:
        return _LMAX_MEMO[key]
    if _LMAX_TABLE is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lmax_table.json")
        try:
            with open(path) as f:
                _LMAX_TABLE = json.load(f)
        except (OSError, ValueError):
            _LMAX_TABLE = {}
    if key in _LMAX_TABLE:
        _LMAX_MEMO[key] = float(_LMAX_TABLE[key])
        return _LMAX_MEMO[key]
    cdir = os.environ.get("DEEPSPHERE_CACHE_DIR")
    if cdir:
        import hashlib

        path = os.path.join(cdir, "lmax_" + hashlib.sha1(key.encode()).hexdigest() + ".txt")
        try:
            with open(path) as f:
                k2, v = f.read().split("\n")[:2]
            if k2 == key:
                _LMAX_MEMO[key] = float(v)
                return _LMAX_MEMO[key]
        except (OSError, ValueError):
            pass
    return None


def _lmax_store(key, value):
    import os

    if key is not None:
        _LMAX_MEMO[key] = float(value)
        cdir = os.environ.get("DEEPSPHERE_CACHE_DIR")
        if cdir:
            import hashlib

            try:
                os.makedirs(cdir, exist_ok=True)
                path = os.path.join(cdir, "lmax_" + hashlib.sha1(key.encode()).hexdigest() + ".txt")
                tmp = f"{path}.{os.getpid()}"
                with open(tmp, "w") as f:
                    f.write(f"{key}\n{float(value)!r}\n")
                os.replace(tmp, path)
            except OSError:
                pass
    return float(value)


def _lanczos_extreme(matvec, M, tol, maxit):
    """Ritz value of largest magnitude of the symmetric operator `matvec`, or None if `maxit` steps do not bring
    its residual bound below ``tol * |theta|``."""
    from scipy.linalg import eigh_tridiagonal

    rng = np.random.default_rng(0)
    q = rng.standard_normal(M)
    q /= np.linalg.norm(q)
    q_prev = np.zeros(M)
    beta = 0.0
    alphas, betas = [], []
    for j in range(maxit):
        w = matvec(q)
        a = float(q @ w)
        w -= a * q
        w -= beta * q_prev
        b = float(np.linalg.norm(w))
        alphas.append(a)
        betas.append(b)
        breakdown = b <= 1e-14 * max(abs(a), 1.0)  # Krylov space exhausted: T holds exact eigenvalues
        if breakdown or (j >= 20 and j % 10 == 0):
            d, e = np.asarray(alphas), np.asarray(betas[:-1])
            lo, vlo = eigh_tridiagonal(d, e, select="i", select_range=(0, 0))
            hi, vhi = eigh_tridiagonal(d, e, select="i", select_range=(j, j))
            theta, s_last = (hi[0], vhi[-1, 0]) if abs(hi[0]) >= abs(lo[0]) else (lo[0], vlo[-1, 0])
            if breakdown or abs(b * s_last) <= tol * abs(theta):
                return float(theta)
        q_prev, q = q, w / b
        beta = b
    return None


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
