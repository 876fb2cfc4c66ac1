"""Diagnostic for the tcgen05 contraction (run on a B200): structured inputs that expose layout
mistakes (swizzle, descriptor strides, TMEM lane/column mapping), then random inputs per mode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
from deepsphere import _native as nat, _ops, gnn_layers  # noqa: E402


def run(M, B, Fin, Fout, K, mode, pattern):
    L = np.eye(M)
    layer = gnn_layers.Chebyshev(L=L, K=K, Fout=Fout, mode=mode)
    layer.build_from_shape((B, M, Fin))
    rng = np.random.default_rng(0)
    if pattern == "identity":
        # K = 1: y = x @ W with W = [I | 0]: y must equal the first Fout columns of x
        x = rng.integers(-8, 8, size=(B, M, Fin)).astype(np.float32)
        W = np.zeros((K * Fin, Fout), np.float32)
        for o in range(min(Fin, Fout)):
            W[o * K, o] = 1.0
    else:
        x = rng.standard_normal((B, M, Fin)).astype(np.float32)
        W = (rng.standard_normal((K * Fin, Fout)) * 0.1).astype(np.float32)
    with torch.no_grad():
        layer.kernel.copy_(torch.tensor(W).cuda())
        y = layer(torch.tensor(x).cuda()).cpu().numpy()
    # reference in float64 using the layer's own L~ (identity graph: L~ = a*I - I)
    a = 1.5 / layer.lmax - 1.0 if K > 1 else 0.0
    T = [x.astype(np.float64)]
    if K > 1:
        T.append(a * T[0])
    for k in range(2, K):
        T.append(2 * a * T[-1] - T[-2])
    Wr = W.astype(np.float64).reshape(Fin, K, Fout)
    ref = sum(np.einsum("bmf,fo->bmo", T[k], Wr[:, k, :]) for k in range(K))
    err = np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30)
    bad = np.argwhere(np.abs(y - ref) > 1e-2 * max(np.abs(ref).max(), 1e-30))
    print(f"M={M} B={B} Fin={Fin} Fout={Fout} K={K} mode={mode} {pattern}: rel err {err:.3e}, bad {len(bad)}")
    if len(bad):
        print("  first bad (b, m, o):", bad[:8].tolist())
        b0, m0, o0 = bad[0]
        print("  got", y[b0, m0, :8], "\n  ref", ref[b0, m0, :8])
    return err


def tf32_conversion_probe():
    """x = 1 + 2^-11 + 2^-12 sits between two TF32 values (1 and 1 + 2^-10): truncation gives 1,
    round-to-nearest gives 1 + 2^-10.  K = 1, W = identity-ish so y = tf32(x) * 1."""
    M, Fin, Fout = 128, 32, 32
    layer = gnn_layers.Chebyshev(L=np.eye(M), K=1, Fout=Fout, mode="tf32")
    layer.build_from_shape((1, M, Fin))
    W = np.zeros((Fin, Fout), np.float32)
    W[np.arange(Fout), np.arange(Fout)] = 1.0
    x = np.full((1, M, Fin), 1.0 + 2.0**-11 + 2.0**-12, np.float32)
    with torch.no_grad():
        layer.kernel.copy_(torch.tensor(W).cuda())
        y = layer(torch.tensor(x).cuda()).cpu().numpy()
    v = float(y[0, 0, 0])
    kind = "truncation" if v == 1.0 else ("round-to-nearest" if v == 1.0 + 2.0**-10 else f"other ({v!r})")
    print(f"hardware fp32->tf32 conversion of the A operand: {kind} (y = 1 + {v - 1.0:.3e})")


if __name__ == "__main__":
    torch.cuda.set_device(0)
    tf32_conversion_probe()
    worst = {}
    for mode in ("tf32", "tf32x3"):
        errs = []
        errs.append(run(256, 1, 32, 32, 1, mode, "identity"))
        errs.append(run(256, 1, 64, 64, 1, mode, "identity"))
        errs.append(run(192, 3, 64, 64, 1, mode, "random"))
        errs.append(run(192, 3, 64, 64, 5, mode, "random"))
        errs.append(run(1000, 2, 16, 48, 3, mode, "random"))
        errs.append(run(4096, 4, 8, 256, 2, mode, "random"))
        worst[mode] = max(errs)
    print("worst:", worst)
    print("launches:", nat.launch_count())
