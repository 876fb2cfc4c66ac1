"""Shared test helpers (imports the ORACLE — tests are one of the few places allowed to)."""
import os

import numpy as np
from scipy import sparse

from oracle import deepsphere_oracle as orc  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: d[k] for k in d.files}
    M = int(g["M"])
    if "Lt_row" in g:
        g["Lt"] = sparse.csr_matrix((g["Lt_val"], (g["Lt_row"], g["Lt_col"])), shape=(M, M))
        g["L"] = sparse.csr_matrix((g["L_val"], (g["L_row"], g["L_col"])), shape=(M, M))
        g["K"] = int(g["K"])
        g["recursion"] = str(g["recursion"])
        g["activation"] = str(g["activation"]) or None
        g["bias"] = g["bias"] if g["bias"].size else None
    return g


def rel_err(a, ref):
    """max |a - ref| / max |ref|  — the parity metric (fp32 SpMM path: <= 1e-5)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


def rel_l2(a, ref):
    """||a - ref||_2 / ||ref||_2 — a second, scale-aware parity metric next to the max-norm one (the max-norm
    ratio is lenient where most of a tensor is small next to its largest element)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.linalg.norm((a - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


CONV_CASES = ["cheb_ref3x3", "mono_ref3x3", "cheb_eye192", "cheb_nside4_k8", "mono_nside4_k8", "cheb_nside8_k20",
              "cheb_masked16_k20", "cheb_masked16_k8"]
