"""Generates tests/golden/*.npz — small seeded input/output vectors of the hot path.

The reference (TensorFlow + healpy + PyGSP) cannot be imported in this image, so these are
produced by the fp64 restatement in oracle/deepsphere_oracle.py (see its header for the
pinning status); they freeze the oracle so that later edits cannot silently move it, and
give the CUDA tests fixed targets.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy import sparse

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))

from oracle import deepsphere_oracle as orc  # noqa: E402
from deepsphere.graph import SphereHealpix  # noqa: E402
from deepsphere import healpix as hpx  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def conv_case(name, L, recursion, K, B, Fin, Fout, seed, bias=False, activation=None):
    rng = np.random.default_rng(seed)
    scale = 0.75 if recursion == "chebyshev" else 1.0
    Lt, lmax = orc.prepare_laplacian(L, scale)
    M = Lt.shape[0]
    x = rng.standard_normal((B, M, Fin))
    kernel = rng.standard_normal((K * Fin, Fout)) / np.sqrt(Fin * (K + 0.5) / 2)
    b = rng.standard_normal((1, 1, Fout)) * 0.1 if bias else None
    dy = rng.standard_normal((B, M, Fout))
    y64 = orc.graph_conv_forward(x, Lt, kernel, K, recursion, bias=b, activation=activation, dtype=np.float64)
    y32 = orc.graph_conv_forward(x.astype(np.float32), Lt, kernel.astype(np.float32), K, recursion,
                                 bias=None if b is None else b.astype(np.float32), activation=activation,
                                 dtype=np.float32)
    # gradients of the linear part (no bias / activation): dx, dkernel, dbias
    dx, dk, db = orc.graph_conv_backward(x, Lt, kernel, K, dy, recursion, dtype=np.float64)
    coo = Lt.tocoo()
    Lin = sparse.coo_matrix(L)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        L_row=Lin.row.astype(np.int64), L_col=Lin.col.astype(np.int64), L_val=Lin.data.astype(np.float64),
        Lt_row=coo.row.astype(np.int64), Lt_col=coo.col.astype(np.int64), Lt_val=coo.data.astype(np.float64),
        M=M, lmax=lmax, K=K, recursion=recursion, x=x, kernel=kernel,
        bias=np.zeros(0) if b is None else b, activation="" if activation is None else activation,
        dy=dy, y64=y64, y32=y32, dx64=dx, dkernel64=dk, dbias64=db,
    )
    print(name, "M", M, "nnz", Lt.nnz, "max|y|", np.abs(y64).max(), "fp32 rel err",
          np.abs(y32 - y64).max() / np.abs(y64).max())


def bernstein_case(name, L, K, B, Fin, Fout, seed, stale=True):
    """Bernstein layer (gnn_layers.py:416-572): y from the literal loop restatement; gradients of the linear part
    from float64 torch autograd over the same loops (torch.sparse.mm)."""
    import torch

    rng = np.random.default_rng(seed)
    Lt, lmax = orc.prepare_laplacian(L, 0.75)
    M = Lt.shape[0]
    x = rng.standard_normal((B, M, Fin))
    kernel = rng.standard_normal(((K + 1) * Fin, Fout)) * np.sqrt(6 / (Fin + Fout))
    dy = rng.standard_normal((B, M, Fout))
    y64 = orc.bernstein_forward(x, Lt, kernel, K, dtype=np.float64)
    xt = torch.tensor(x, requires_grad=True)
    wt = torch.tensor(kernel, requires_grad=True)
    yt = orc.torch_cpu_bernstein(xt, Lt, wt, K)
    assert np.abs(yt.detach().numpy() - y64).max() <= 1e-12 * np.abs(y64).max()
    yt.backward(torch.tensor(dy))
    coo, Lin = Lt.tocoo(), sparse.coo_matrix(L)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        L_row=Lin.row.astype(np.int64), L_col=Lin.col.astype(np.int64), L_val=Lin.data.astype(np.float64),
        Lt_row=coo.row.astype(np.int64), Lt_col=coo.col.astype(np.int64), Lt_val=coo.data.astype(np.float64),
        M=M, lmax=lmax, K=K, recursion="bernstein", x=x, kernel=kernel, bias=np.zeros(0), activation="",
        dy=dy, y64=y64, dx64=xt.grad.numpy(), dkernel64=wt.grad.numpy(),
    )
    print(name, "M", M, "K", K, "max|y|", np.abs(y64).max())


def smoothing_case(name, nside, indices, sigma_arcmin, B, C, reps, seed):
    """HealpySmoothing (healpy_layers.py:510-853): kernel from the reference's BallTree recipe on this repo's pixel
    centres, normalised as written there, applied once and with per-channel repetitions + mask."""
    rng = np.random.default_rng(seed)
    theta, phi = hpx.pix2ang(nside, indices)
    sigma_rad = sigma_arcmin * np.pi / (60 * 180)
    ind_coo, val_coo = orc.smoothing_neighbours(0.5 * np.pi - theta, phi, sigma_rad, 3)
    n = len(indices)
    Ks = orc.smoothing_kernel(ind_coo, val_coo, n)
    x = rng.standard_normal((B, n, C))
    mask = (rng.random((n, 1)) > 0.2)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), M=n, nside=nside, indices=np.asarray(indices, dtype=np.int64),
        sigma_arcmin=sigma_arcmin, ind_coo=ind_coo, val_coo=val_coo, x=x, reps=np.asarray(reps, dtype=np.int64),
        mask=mask, y_once64=orc.smoothing_forward(x, Ks, dtype=np.float64),
        y_reps_mask64=orc.smoothing_forward(x, Ks, reps, mask[None], dtype=np.float64),
    )
    print(name, "n", n, "nnz", len(val_coo), "neighbours per row", len(val_coo) // n)


def next_cases():
    """Bernstein and HealpySmoothing."""
    bernstein_case("bern_nside4_k8", SphereHealpix(4, k=8).L, 4, 2, 3, 5, seed=21)
    disc = hpx.query_disc(16, [1, 0, 0], 0.9)
    ext = orc.extend_indices(disc, 16, 4)
    bernstein_case("bern_masked16_k20", SphereHealpix(16, indexes=ext, k=20).L, 7, 2, 2, 3, seed=22)
    smoothing_case("smooth_masked16", 16, ext, 300.0, 2, 3, [1, 3, 2], seed=23)


def main():
    # 1. the reference's own test input shape: L = A A^T (3x3), x [5,3,7], K = 4, Fout = 3
    #    (tests/test_gnn_layers.py:9-33; TF's RNG is not reproducible here -> numpy seed 11)
    rng = np.random.default_rng(11)
    A = rng.standard_normal((3, 3))
    L3 = A @ A.T
    conv_case("cheb_ref3x3", L3, "chebyshev", 4, 5, 7, 3, seed=12)
    conv_case("mono_ref3x3", L3, "monomial", 4, 5, 7, 3, seed=12, bias=True, activation="elu")
    # 2. identity Laplacian (tests/test_gnn_layers.py:104-109 uses np.eye(192), K = 5)
    conv_case("cheb_eye192", np.eye(192), "chebyshev", 5, 3, 2, 7, seed=13)
    # 3. full-sphere HEALPix graphs
    conv_case("cheb_nside4_k8", SphereHealpix(4, k=8).L, "chebyshev", 5, 2, 4, 6, seed=14, bias=True,
              activation="relu")
    conv_case("mono_nside4_k8", SphereHealpix(4, k=8).L, "monomial", 3, 2, 5, 5, seed=15)
    conv_case("cheb_nside8_k20", SphereHealpix(8, k=20).L, "chebyshev", 10, 2, 1, 5, seed=16)
    # 4. masked sky (irregular degrees -> ELL + CSR tail)
    disc = hpx.query_disc(16, [1, 0, 0], 0.9)
    ext = orc.extend_indices(disc, 16, 4)
    conv_case("cheb_masked16_k20", SphereHealpix(16, indexes=ext, k=20).L, "chebyshev", 4, 3, 3, 2, seed=17)
    conv_case("cheb_masked16_k8", SphereHealpix(16, indexes=ext, k=8).L, "chebyshev", 6, 2, 8, 16, seed=18)

    # 5. HealpyPool — the reference's known-answer recipe (tests/test_healpy_layers.py:9-37)
    np.random.seed(11)
    m_in = np.random.normal(size=12 * 4 * 4)
    np.savez_compressed(
        os.path.join(OUT, "pool_nside4.npz"), m_in=m_in,
        avg=m_in.reshape(-1, 4).mean(axis=1),  # == hp.ud_grade(nside 4 -> 2, NEST, power=None)
        max=np.max(m_in.reshape((-1, 4)), axis=1),
    )
    # 6. pseudo convolutions
    rng = np.random.default_rng(19)
    x = rng.standard_normal((2, 12 * 8 * 8, 3))
    w = rng.standard_normal((16, 3, 5)) * 0.2
    b = rng.standard_normal(5) * 0.1
    wt = rng.standard_normal((1, 16, 5, 3)) * 0.2
    np.savez_compressed(
        os.path.join(OUT, "pconv_nside8.npz"), x=x, w=w, b=b, wt=wt,
        y=orc.pseudo_conv(x, w, b, "elu"), yt=orc.pseudo_conv_transpose(x[:, :48], wt, b, "relu"),
    )
    next_cases()
    print("golden written to", OUT)


if __name__ == "__main__":
    if "--next-only" in sys.argv:  # only the section-8f cases (leaves the other files untouched)
        next_cases()
    else:
        main()
