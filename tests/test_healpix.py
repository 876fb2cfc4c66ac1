"""HEALPix NESTED geometry (host, exact) — self-consistency and the reference's known answers."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from deepsphere import healpix as hpx
from deepsphere import utils


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 64])
def test_nest_xyf_roundtrip_and_ring_bijection(nside):
    npix = hpx.nside2npix(nside)
    ids = np.arange(npix)
    x, y, f = hpx.nest2xyf(nside, ids)
    assert np.array_equal(hpx.xyf2nest(nside, x, y, f), ids)
    r = hpx.nest2ring(nside, ids)
    assert np.array_equal(np.sort(r), ids)
    assert np.array_equal(hpx.ring2nest(nside, r), ids)
    # RING order runs north -> south: z must not increase with the ring index
    z = np.empty(npix)
    z[r] = hpx.pix2vec(nside, ids)[:, 2]
    assert np.all(np.diff(z) <= 1e-12)


@pytest.mark.parametrize("nside", [2, 8, 32])
def test_pixel_centres(nside):
    npix = hpx.nside2npix(nside)
    v = hpx.pix2vec(nside, np.arange(npix))
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-14)
    assert np.abs(v.sum(axis=0)).max() < 1e-10  # symmetric tessellation
    d, _ = cKDTree(v).query(v, k=2)
    assert d[:, 1].min() > 0.5 * np.sqrt(4 * np.pi / npix)  # distinct, roughly equal-area


@pytest.mark.parametrize("nside", [2, 4, 16, 64])
def test_neighbours(nside):
    npix = hpx.nside2npix(nside)
    ids = np.arange(npix)
    nb = hpx.neighbours(nside, ids)
    assert (nb < 0).sum() == 24  # 3 pixels at each of the 8 valence-3 vertices have 7 neighbours
    for row in nb[:: max(1, npix // 500)]:
        assert len(set(row[row >= 0])) == (row >= 0).sum()
    for d in range(8):  # symmetry: i in nb(j) <=> j in nb(i), exactly once
        j = nb[:, d]
        m = j >= 0
        assert np.all((nb[j[m]] == ids[m][:, None]).sum(axis=1) == 1)
    # geometric: every neighbour is among the 12 nearest pixel centres
    v = hpx.pix2vec(nside, ids)
    _, idx = cKDTree(v).query(v, k=13)
    sample = ids[:: max(1, npix // 2000)]
    for i in sample:
        assert set(nb[i][nb[i] >= 0]) <= set(idx[i][1:])


def test_known_answer_advanced_tutorial_indices():
    """examples/advanced_tutorial.ipynb:137,211,356: query_disc(nside 64, [1,0,0], 1.5) extended to
    nside_out 8 has 24 832 pixels, then 6 208 -> 1 552 -> 388 through three p=1 reductions."""
    disc = hpx.query_disc(64, [1, 0, 0], 1.5)
    ext = utils.extend_indices(disc, nside_in=64, nside_out=8)
    assert len(ext) == 24832
    sizes = [len(hpx.coarsen_indices(ext, p)) for p in (1, 2, 3)]
    assert sizes == [6208, 1552, 388]


def test_extend_indices_reference_test():
    """reference tests/test_utils.py:7-31 (NEST and RING)."""
    nside_in, nside_out = 4, 2
    npix = hpx.nside2npix(nside_in)
    indices = np.arange(npix)[::4]
    assert len(utils.extend_indices(indices, nside_in=nside_in, nside_out=nside_out)) == npix
    m_ring = np.zeros(npix)
    m_ring[hpx.nest2ring(nside_in, np.arange(npix)[::4])] = 1.0
    indices = np.arange(npix)[m_ring > 0.0]
    assert len(utils.extend_indices(indices, nside_in=nside_in, nside_out=nside_out, nest=False)) == npix


def test_ud_grade_mask_matches_integer_form():
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(hpx.nside2npix(16), 300, replace=False))
    m = np.zeros(hpx.nside2npix(16))
    m[sel] = 1
    down = hpx.ud_grade_mask_nest(m, 4)
    assert np.array_equal(np.arange(len(down))[down > 1e-12], hpx.coarsen_indices(sel, 2))
    up = hpx.ud_grade_mask_nest((down > 1e-12).astype(float), 16)
    assert np.array_equal(np.arange(len(up))[up > 1e-12], utils.extend_indices(sel, 16, 4))


def test_isnsideok():
    assert hpx.isnsideok(64) and hpx.isnsideok(1)
    assert not hpx.isnsideok(12) and not hpx.isnsideok(0) and not hpx.isnsideok(2.5)
