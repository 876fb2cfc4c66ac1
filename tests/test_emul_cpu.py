"""Host emulation of simple CUDA kernels (tests/emul): the UNCHANGED kernel source of csrc/ds_skinny.cu is compiled by
g++ against a stand-in <cuda_runtime.h> (every CUDA thread a std::thread, __syncthreads a barrier, warp shuffles through
an exchange buffer) and checked against naive loops — index arithmetic, the shuffle/shared-memory reduction tree and the
partial-sum layout are verified without a GPU; ThreadSanitizer looks for unsynchronised shared-memory accesses."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "emul_skinny.cpp")


def _build_and_run(tmp_path, extra):
    exe = os.path.join(tmp_path, "emul_skinny")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-DDS_EMULATE", "-I", os.path.join(HERE, "emul"), *extra, SRC,
           "-o", exe]
    subprocess.run(cmd, check=True, cwd=ROOT, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_skinny_kernels_on_the_host_emulator(tmp_path):
    res = _build_and_run(str(tmp_path), [])
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok ") == 11 and "FAIL" not in res.stdout


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_skinny_kernels_have_no_shared_memory_race(tmp_path):
    res = _build_and_run(str(tmp_path), ["-g", "-fsanitize=thread"])
    if "FATAL: ThreadSanitizer" in res.stderr and "unexpected memory mapping" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (ASLR settings)")
    assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr, res.stdout + res.stderr[-2000:]


# ---------------------------------------------------------------------------------------------------------------
# The fused lattice convolution kernel (csrc/ds_lattice_conv2.cu) on the host emulator: warp-specialised roles, mbarrier
# protocol, cp.async gathers, UMMA shared-memory descriptors, tensor-memory epilogue — the unchanged source, default build
# and the experiment variants (tools/build_variants.sh), against the oracle.
def _build_conv2(tmp_path, name, defines):
    exe = os.path.join(tmp_path, name)
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-DDS_EMULATE", *defines, "-I", os.path.join(HERE, "emul"),
           os.path.join(HERE, "emul", "emul_conv2.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, cwd=ROOT, capture_output=True, text=True)
    return exe


def _conv2_problem():
    import numpy as np
    from deepsphere import gnn_layers
    from deepsphere.graph import SphereHealpix

    g = SphereHealpix(32, k=8)
    out = {}
    for cls, K in ((gnn_layers.Chebyshev, 5), (gnn_layers.Monomial, 4)):
        layer = cls(L=g.L, K=K, Fout=16)
        out[cls.__name__] = (layer, layer._lattice_payload())
    # a masked sky with holes inside the tiles' lattices (zero-filled gathers, rows without an output)
    from deepsphere import healpix as hpx
    from helpers import orc

    disc = hpx.query_disc(64, [0.3, 0.5, 0.8], 0.55)
    ext = orc.extend_indices(disc, 64, 16)
    gm = SphereHealpix(64, indexes=ext, k=8)
    layer = gnn_layers.Chebyshev(L=gm.L, K=5, Fout=16, healpix=(64, ext))
    out["Masked"] = (layer, layer._lattice_payload())
    return g, out


@pytest.fixture(scope="module")
def conv2_problem():
    return _conv2_problem()


# (the C2_FENCE_BY_ISSUER variants are not listed: proxy fences are no-ops on the emulator, they run like their base)
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("variant,defines", [
    ("default", []),
    ("split", ["-DC2_SPLIT_BAR=1"]),
    ("symw", ["-DC2_SYMW=1"]),
    ("br3all", ["-DC2_SYMW=1", "-DC2_EPI_PIPE=1", "-DC2_PROBE=0", "-DC2_SPLIT_BAR=1"]),
    # round 2: two channels per thread (8 compute warps), and the 3 x 6 pixel block with the scalar diagonal
    ("cpt2", ["-DC2_CPT=2", "-DC2_SYMW=1", "-DC2_PROBE=0"]),
    ("bc6", ["-DC2_CPT=2", "-DC2_BC=6", "-DC2_SYMW=1", "-DC2_PROBE=0"]),
    # the hops as a loop over one code body per buffer parity (instruction-cache footprint)
    ("loop", ["-DC2_LOOP=1", "-DC2_PROBE=0"]),
    ("bc6loop", ["-DC2_LOOP=1", "-DC2_CPT=2", "-DC2_BC=6", "-DC2_SYMW=1", "-DC2_PROBE=0"]),
    # the accumulator drain on the gather warps + one extra warp (288 threads)
    ("iodrain", ["-DC2_IO_DRAIN=1", "-DC2_SYMW=1"]),
    # the per-hop basis stores of a launch with `out` done by the gather warps from the exchange buffer
    ("ioout", ["-DC2_IO_OUT=1"]),
])
def test_fused_lattice_kernel_on_the_host_emulator(tmp_path, conv2_problem, variant, defines):
    import numpy as np
    from scipy import sparse

    from helpers import orc

    exe = _build_conv2(str(tmp_path), "emul_conv2_" + variant, defines)
    g, layers = conv2_problem
    M = g.L.shape[0]
    rng = np.random.default_rng(7)
    cases = [  # (layer class, K, B, F, N, activation id, bias, b_split, grid, basis wanted)
        ("Chebyshev", 5, 1, 8, 16, 0, True, 1, 2, True),
        ("Monomial", 4, 2, 16, 32, 1, False, 2, 3, False),
        ("Chebyshev", 3, 1, 16, 16, 0, False, 1, 2, True),   # 2 hops
        ("Chebyshev", 2, 1, 8, 64, 1, True, 1, 1, False),    # 1 hop, widest accumulator (3 x 64 TMEM columns)
        ("Chebyshev-bwd", 5, 1, 16, 32, 0, False, 1, 2, True),  # backward-data launch: dx from dz, basis U_k out
        ("Masked", 5, 2, 8, 16, 0, True, 1, 2, True),           # partial sky: holes in the lattices
        ("Chebyshev", 5, 1, 64, 64, 1, True, 1, 3, False),      # the bench layer's shape: 8 chunks, 192 TMEM columns
    ]
    for ci, (name, K, B, F, N, act, has_bias, b_split, grid, want_basis) in enumerate(cases):
        if F == 64 and variant not in ("default", "bc6"):
            continue  # the large case only for the measured kernel and the most changed variant (CPU suite budget)
        if K <= 3 and variant not in ("default", "split", "bc6", "loop", "bc6loop", "iodrain", "ioout"):
            continue  # 1- and 2-hop launches: only where the hop sequence itself differs
        bwd = name.endswith("-bwd")
        name = name.split("-")[0]
        layer, pay = layers[name]  # the tile tables do not depend on K (4-ring halo for every K <= 5)
        M = int(layer._L_shape[0])
        recursion = "monomial" if name == "Monomial" else "chebyshev"
        assert pay is not None and pay["n_tiles"] >= 8 and pay["LW"] == 24 and pay["H"] == 4
        d = os.path.join(str(tmp_path), f"{variant}_{ci}_{name}")
        os.makedirs(d)
        x = rng.standard_normal((B, M, F)).astype(np.float32)
        W = (rng.standard_normal((N * K, F) if bwd else (F * K, N)) * 0.2).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32)
        for arr, fn in ((pay["pix"], "pix"), (pay["w"], "w"), (x, "x"), (W, "W"), (bias, "bias")):
            arr.tofile(os.path.join(d, fn + ".bin"))
        cheb = int(recursion == "chebyshev")
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"{pay['n_tiles']} {B} {M} {F} {N} {K - 1} {cheb} {act} {int(has_bias)} {grid} {b_split} "
                    f"{int(want_basis)} {int(bwd)}\n")
        res = subprocess.run([exe, d], capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stdout + res.stderr[-2000:]
        y = np.fromfile(os.path.join(d, "y.bin"), dtype=np.float32).reshape(B, M, N)
        Lt = sparse.csr_matrix((layer._L_values.astype(np.float64), (layer._L_indices[:, 0], layer._L_indices[:, 1])),
                               shape=(M, M))
        if bwd:  # x plays dz [B, M, Fout = F]; the layer kernel is [(N*K), F]; the launch returns dx [B, M, Fin = N]
            ref, _, _ = orc.graph_conv_backward(np.zeros((B, M, N)), Lt, W.astype(np.float64), K, x.astype(np.float64),
                                                recursion)
        else:
            ref = orc.graph_conv_forward(x.astype(np.float64), Lt, W.astype(np.float64), K, recursion,
                                         bias=bias.reshape(1, 1, -1).astype(np.float64) if has_bias else None,
                                         activation="relu" if act == 1 else None, dtype=np.float64)
        launched = pay["pix"].reshape(pay["n_tiles"], 24, 24)[:, 4:20, 4:20].ravel()
        launched = np.sort(launched[launched >= 0])
        assert len(launched) == len(np.unique(launched)) and (name == "Masked" or len(launched) == M)
        other = np.setdiff1d(np.arange(M), launched)
        assert np.isnan(y[:, other]).all()  # rows of tiles that are not launched: untouched
        # the rows the launch answers for: all own pixels but the 15 per tile within reach of a valence-3 vertex (those
        # are overwritten by the generic sub-problem, lattice.make_payload)
        own = pay["lattice_rows"]
        assert name == "Masked" or len(own) == M - 24 * 15
        err = np.abs(y[:, own] - ref[:, own]).max() / np.abs(ref).max()
        assert err <= 1e-3, (variant, name, err)  # TF32 contraction (the GPU measures 5e-4)
        if want_basis:  # fp32 recursion: T_1..T_{K-1} on the own pixels
            t_prev, t_cur = x.astype(np.float64), np.stack([Lt @ x[b].astype(np.float64) for b in range(B)])
            for s in range(1, K):
                u = np.fromfile(os.path.join(d, f"u{s}.bin"), dtype=np.float32).reshape(B, M, F)
                assert np.abs(u[:, own] - t_cur[:, own]).max() <= 1e-5 * np.abs(t_cur).max(), (variant, s)
                t_prev, t_cur = t_cur, 2 * np.stack([Lt @ t_cur[b] for b in range(B)]) - t_prev


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("variant,defines", [
    ("default", []),
    ("bc6", ["-DC2_CPT=2", "-DC2_BC=6", "-DC2_SYMW=1", "-DC2_PROBE=0"]),
    ("loop", ["-DC2_LOOP=1", "-DC2_PROBE=0"]),
    ("iodrain", ["-DC2_IO_DRAIN=1", "-DC2_SYMW=1"]),
    ("ioout", ["-DC2_IO_OUT=1"]),
])
def test_fused_lattice_kernel_protocol_under_thread_sanitizer(tmp_path, conv2_problem, variant, defines):
    """The barrier protocol of the fused kernel under ThreadSanitizer: every emulated mbarrier is its own lock, so the
    only happens-before edges are the ones the protocol creates; an exchange buffer written while a neighbour still reads
    it, or read before it was published, is reported as a data race (removing the wait for the previous item's last
    phase from the kernel is — checked by hand — caught this way)."""
    import numpy as np

    exe = _build_conv2(str(tmp_path), "emul_conv2_tsan_" + variant, ["-g", "-fsanitize=thread", *defines])
    g, layers = conv2_problem
    layer, pay = layers["Chebyshev"]
    M, K, B, F, N = g.L.shape[0], 5, 1, 16, 16
    d = os.path.join(str(tmp_path), "case")
    os.makedirs(d)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, M, F)).astype(np.float32)
    W = (rng.standard_normal((F * K, N)) * 0.2).astype(np.float32)
    for arr, fn in ((pay["pix"], "pix"), (pay["w"], "w"), (x, "x"), (W, "W")):
        arr.tofile(os.path.join(d, fn + ".bin"))
    with open(os.path.join(d, "meta.txt"), "w") as f:
        f.write(f"{pay['n_tiles']} {B} {M} {F} {N} {K - 1} 1 0 0 2 1 1 0\n")
    res = subprocess.run([exe, d], capture_output=True, text=True, timeout=900)
    if "FATAL: ThreadSanitizer" in res.stderr and "unexpected memory mapping" in res.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (ASLR settings)")
    assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr, res.stderr[-3000:]


# ---------------------------------------------------------------------------------------------------------------
# patch_conv_kernel (csrc/ds_patch.cu): the irregular rows of a lattice plan - the own pixels within reach of the valence-3
# vertices - in one launch, on the host emulator against the oracle; with the ThreadSanitizer build the rotation of the
# three shared-memory buffers and the per-hop weight image are checked for missing barriers.
def _run_patch_cases(tmp_path, exe, conv2_problem, cases):
    import numpy as np
    from scipy import sparse

    from helpers import orc

    g, layers = conv2_problem
    rng = np.random.default_rng(3)
    for ci, (name, K, B, F, N, act, has_bias, want_basis) in enumerate(cases):
        bwd = name.endswith("-bwd")
        name = name.split("-")[0]
        layer, pay = layers[name]
        pt = pay["patches"]
        assert pt is not None and pt["n_patches"] == 8
        M = int(layer._L_shape[0])
        recursion = "monomial" if name == "Monomial" else "chebyshev"
        d = os.path.join(str(tmp_path), f"patch_{ci}")
        os.makedirs(d)
        x = rng.standard_normal((B, M, F)).astype(np.float32)
        # forward: kernel [(F*K), N], B_k(f, n) = W[(f*K + k)*N + n]; backward-data: kernel [(N*K), F], x plays dz,
        # B_k(f, n) = W[(n*K + k)*F + f] (the transposed read of the same layer kernel)
        W = (rng.standard_normal((N * K, F) if bwd else (F * K, N)) * 0.2).astype(np.float32)
        s_f, s_k, s_n = (1, F, K * F) if bwd else (K * N, N, 1)
        bias = rng.standard_normal(N).astype(np.float32)
        for key in ("row_ptr", "rows", "ell_col", "ell_val", "own_ptr", "own_local"):
            pt[key].tofile(os.path.join(d, key + ".bin"))
        for arr, fn in ((x, "x"), (W, "W"), (bias, "bias")):
            arr.tofile(os.path.join(d, fn + ".bin"))
        max_rows = int(np.diff(pt["row_ptr"]).max())
        with open(os.path.join(d, "meta.txt"), "w") as f:
            f.write(f"{pt['n_patches']} {B} {M} {F} {N} {K - 1} {int(recursion == 'chebyshev')} {act} {int(has_bias)} "
                    f"{int(want_basis)} {s_f} {s_k} {s_n} {max_rows}\n")
        res = subprocess.run([exe, d], capture_output=True, text=True, timeout=900)
        assert res.returncode == 0 and "WARNING: ThreadSanitizer" not in res.stderr, res.stdout + res.stderr[-3000:]
        y = np.fromfile(os.path.join(d, "y.bin"), dtype=np.float32).reshape(B, M, N)
        Lt = sparse.csr_matrix((layer._L_values.astype(np.float64), (layer._L_indices[:, 0], layer._L_indices[:, 1])),
                               shape=(M, M))
        if bwd:
            ref, _, _ = orc.graph_conv_backward(np.zeros((B, M, N)), Lt, W.astype(np.float64), K, x.astype(np.float64),
                                                recursion)
        else:
            ref = orc.graph_conv_forward(x.astype(np.float64), Lt, W.astype(np.float64), K, recursion,
                                         bias=bias.reshape(1, 1, -1).astype(np.float64) if has_bias else None,
                                         activation="relu" if act == 1 else None, dtype=np.float64)
        want = pay["closure_rows"][pay["own_sub"]]
        assert len(want) == 360
        other = np.setdiff1d(np.arange(M), want)
        assert np.isnan(y[:, other]).all() and np.isfinite(y[:, want]).all()  # exactly the irregular rows are written
        err = np.abs(y[:, want] - ref[:, want]).max() / np.abs(ref).max()
        assert err <= 2e-6, (name, K, F, N, err)  # fp32 FMA throughout
        if want_basis:
            t_prev, t_cur = x.astype(np.float64), np.stack([Lt @ x[b].astype(np.float64) for b in range(B)])
            for s in range(1, K):
                u = np.fromfile(os.path.join(d, f"u{s}.bin"), dtype=np.float32).reshape(B, M, F)
                assert np.isnan(u[:, other]).all()
                assert np.abs(u[:, want] - t_cur[:, want]).max() <= 2e-6 * np.abs(t_cur).max(), s
                nxt = np.stack([Lt @ t_cur[b] for b in range(B)])
                t_prev, t_cur = t_cur, (2 * nxt - t_prev if recursion == "chebyshev" else nxt)


def _build_patch(tmp_path, extra):
    exe = os.path.join(tmp_path, "emul_patch")
    cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-DDS_EMULATE", *extra, "-I", os.path.join(HERE, "emul"),
           os.path.join(HERE, "emul", "emul_patch.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, cwd=ROOT, capture_output=True, text=True)
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_patch_kernel_on_the_host_emulator(tmp_path, conv2_problem):
    exe = _build_patch(str(tmp_path), [])
    _run_patch_cases(tmp_path, exe, conv2_problem, [
        # (layer, K, B, F, N, activation id, bias, basis wanted)
        ("Chebyshev", 5, 2, 8, 16, 0, True, True),
        ("Monomial", 4, 1, 16, 32, 1, False, True),
        ("Chebyshev", 2, 1, 8, 64, 1, True, False),       # one hop
        ("Chebyshev-bwd", 5, 1, 32, 16, 0, False, True),  # backward-data launch: transposed weight strides
        ("Chebyshev", 5, 1, 64, 80, 1, True, False),      # widest: 3 row groups x 16 accumulators >= 45 wanted rows
        ("Chebyshev", 3, 1, 4, 5, 0, True, False),        # odd N: 51 row groups
    ])


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_patch_kernel_has_no_shared_memory_race(tmp_path, conv2_problem):
    exe = _build_patch(str(tmp_path), ["-g", "-fsanitize=thread"])
    probe = subprocess.run([exe], capture_output=True, text=True)
    if "FATAL: ThreadSanitizer" in probe.stderr and "unexpected memory mapping" in probe.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (ASLR settings)")
    _run_patch_cases(tmp_path, exe, conv2_problem, [("Chebyshev", 5, 1, 8, 16, 1, True, True),
                                                    ("Chebyshev-bwd", 4, 1, 16, 8, 0, False, True)])
