"""Host-side tile plan of the fused lattice kernel: embedding, verification, shrinking-region recursion
(emulated in numpy exactly as the kernel runs it) against the sparse-matrix recursion.  CPU only."""
import numpy as np
import pytest
from scipy import sparse

from deepsphere import healpix as hpx
from deepsphere import lattice, utils
from deepsphere.graph import SphereHealpix


def _prepared(nside, indices=None, k=8):
    g = SphereHealpix(nside, indexes=indices, k=k)
    Lt = utils.rescale_L(g.L, lmax=1.9, scale=0.75)
    return sparse.csr_matrix(Lt.astype(np.float32))


@pytest.mark.parametrize("nside,H,order", [(32, 4, 4), (64, 4, 4), (64, 9, 4), (16, 2, 3), (32, 1, 4)])
def test_full_sphere_plan(nside, H, order):
    Lt = _prepared(nside)
    plan = lattice.build_lattice_plan(Lt, nside, np.arange(12 * nside**2), H, order)
    assert plan is not None and plan.LW == (1 << order) + 2 * H
    # the three tiles around each of the 8 valence-3 vertices are irregular (with a single ring the missing
    # diagonal neighbour is just a hole, so H = 1 keeps them regular)
    assert (~plan.regular).sum() == (24 if H > 1 else 0)
    if H > 1:
        assert (plan.pix[plan.regular] >= 0).all()  # no holes on the full sphere
    assert lattice.check_plan(plan, Lt) < 1e-12
    own = plan.pix[:, plan.own_mask()]
    assert np.array_equal(np.sort(own.ravel()), np.arange(12 * nside**2))  # every row owned exactly once


def test_masked_plan_and_payload():
    ext = utils.extend_indices(hpx.query_disc(64, [1, 0, 0], 1.5), 64, 8)
    Lt = _prepared(64, ext)
    plan = lattice.build_lattice_plan(Lt, 64, ext, 3, 4)
    assert plan is not None and (plan.pix < 0).any()  # holes outside the survey footprint
    assert lattice.check_plan(plan, Lt) < 1e-12
    p = lattice.make_payload(Lt, 64, ext, 3)
    assert p["n_tiles"] == int(plan.regular.sum()) and p["pix"].dtype == np.int32 and p["w"].shape[2] == 9
    # the sub-problem covers every row owned by an irregular tile
    assert np.array_equal(p["closure_rows"][p["own_sub"]], plan.irregular_rows())


def test_not_applicable_cases():
    g20 = SphereHealpix(16, k=20)
    assert lattice.build_lattice_plan(g20.L, 16, np.arange(12 * 256), 2, 3) is None  # not an 8-neighbour graph
    Lt = _prepared(8)
    assert lattice.build_lattice_plan(Lt, 8, np.arange(768), 2, 4) is None  # tile larger than a face
    assert lattice.make_payload(_prepared(16), 16, np.arange(12 * 256), 4) is None  # every tile touches a vertex


@pytest.mark.parametrize("recursion", ["chebyshev", "monomial"])
def test_shrinking_region_recursion_matches_sparse(recursion):
    """Emulates the kernel: T_0 on the whole lattice, step s computed on ring <= H - s only, in place
    over the buffer holding T_{s-2}; the own pixels must equal the global sparse recursion for every s."""
    nside, H = 32, 4
    Lt = _prepared(nside)
    M = Lt.shape[0]
    plan = lattice.build_lattice_plan(Lt, nside, np.arange(M), H, 4)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(M)
    L64 = sparse.csr_matrix(Lt, dtype=np.float64)
    T = [x, L64 @ x]
    for k in range(2, H + 1):
        T.append(2 * (L64 @ T[-1]) - T[-2] if recursion == "chebyshev" else L64 @ T[-1])
    if recursion == "monomial":
        T = [x]
        for k in range(1, H + 1):
            T.append(L64 @ T[-1])
    LW = plan.LW
    own = plan.own_mask().reshape(LW, LW)
    for t in np.flatnonzero(plan.regular)[:6]:
        pix = plan.pix[t].reshape(LW, LW)
        w = plan.w[t].reshape(LW, LW, 9).astype(np.float64)
        cur = np.where(pix >= 0, x[np.maximum(pix, 0)], 0.0)
        oth = np.zeros_like(cur)
        for s in range(1, H + 1):
            lo, hi = s, LW - 1 - s
            new = oth.copy()
            for j in range(lo, hi + 1):
                for i in range(lo, hi + 1):
                    acc = w[j, i, 8] * cur[j, i]
                    for d in range(8):
                        acc += w[j, i, d] * cur[j + lattice.DJ[d], i + lattice.DI[d]]
                    if recursion == "chebyshev" and s >= 2:
                        new[j, i] = 2 * acc - oth[j, i]
                    else:
                        new[j, i] = acc
            cur, oth = new, cur
            assert np.abs(cur[own] - T[s][pix[own]]).max() < 1e-12, (t, s)
