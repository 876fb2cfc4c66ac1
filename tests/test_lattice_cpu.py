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
    assert p["n_tiles"] == int(plan.usable.sum()) and p["pix"].dtype == np.int32 and p["w"].shape[2] == 9
    # the sub-problem covers every row the lattice launch does not get right, the launch all the others
    assert np.array_equal(p["closure_rows"][p["own_sub"]], plan.irregular_rows())
    assert np.array_equal(np.sort(np.concatenate([p["lattice_rows"], plan.irregular_rows()])), np.arange(len(ext)))


def test_not_applicable_cases():
    g20 = SphereHealpix(16, k=20)
    assert lattice.build_lattice_plan(g20.L, 16, np.arange(12 * 256), 2, 3) is None  # not an 8-neighbour graph
    Lt = _prepared(8)
    assert lattice.build_lattice_plan(Lt, 8, np.arange(768), 2, 4) is None  # tile larger than a face
    # nside 16: every 16 x 16 tile (= base face) touches valence-3 vertices, yet all but 30 pixels of each are exact
    p = lattice.make_payload(_prepared(16), 16, np.arange(12 * 256), 4)
    assert p["n_tiles"] == 12 and p["n_irregular_tiles"] == 12 and len(p["own_sub"]) == 360


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
    DJ, DI = np.asarray(lattice.DJ), np.asarray(lattice.DI)
    irregular = np.flatnonzero(~plan.regular)
    assert len(irregular) == 24
    n_wrong = 0
    for t in list(np.flatnonzero(plan.regular)[:6]) + list(irregular):
        pix = plan.pix[t].reshape(LW, LW)
        w = plan.w[t].reshape(LW, LW, 9).astype(np.float64)
        good = plan.exact[t].reshape(LW, LW) & own & (pix >= 0)
        assert good.sum() == (256 if plan.regular[t] else 256 - 15)
        cur = np.where(pix >= 0, x[np.maximum(pix, 0)], 0.0)
        oth = np.zeros_like(cur)
        for s in range(1, H + 1):
            lo, hi = s, LW - s  # computed region [lo, hi)
            acc = w[lo:hi, lo:hi, 8] * cur[lo:hi, lo:hi]
            for d in range(8):
                acc = acc + w[lo:hi, lo:hi, d] * cur[lo + DJ[d] : hi + DJ[d], lo + DI[d] : hi + DI[d]]
            new = oth.copy()
            new[lo:hi, lo:hi] = 2 * acc - oth[lo:hi, lo:hi] if recursion == "chebyshev" and s >= 2 else acc
            cur, oth = new, cur
            assert np.abs(cur[good] - T[s][pix[good]]).max() < 1e-12, (t, s)
        # ... and the pixels the plan hands to the generic path really are wrong on the lattice (the rule is tight)
        rest = own & (pix >= 0) & ~good
        n_wrong += int((np.abs(cur[rest] - T[H][pix[rest]]) > 1e-9).sum())
    assert n_wrong == 24 * 15
