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
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
