"""HEALPix NESTED-scheme integer geometry, written from the HEALPix definition.

healpy is not installable in this image, and the reference only ever uses it for
``nside2npix``, ``isnsideok``, ``ud_grade`` of 0/1 masks and (through PyGSP) pixel
centres and neighbours (reference call sites: utils.py:27-37,
healpy_networks.py:64,73-78,183-186).  Everything here is host-side and exact:
pixel indices are integers, nest<->(face,x,y) is bit (de)interleaving, and the
pixel centres are evaluated in float64.

All functions are vectorised over numpy integer arrays.  Only the NESTED ordering is
needed by the hot path (the reference assumes it everywhere, healpy_networks.py:37);
``ring2nest``/``nest2ring`` exist for ``extend_indices(nest=False)``.
"""

import numpy as np

# Base-face constants of the HEALPix tessellation (ring of the face's northmost
# corner in units of nside, and its longitude in units of pi/4).
_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)

# Direction order of `neighbours`: SW, W, NW, N, NE, E, SE, S in the face frame.
NB_XOFF = np.array([-1, -1, 0, 1, 1, 1, 0, -1], dtype=np.int64)
NB_YOFF = np.array([0, 1, 1, 1, 0, -1, -1, -1], dtype=np.int64)

# Which face lies in direction (dx,dy) of face f; index = 4 + dx + 3*dy.  -1: no face
# (the 8 valence-3 vertices of the base tessellation).
_NB_FACE = np.array(
    [
        [8, 9, 10, 11, -1, -1, -1, -1, 10, 11, 8, 9],  # S
        [5, 6, 7, 4, 8, 9, 10, 11, 9, 10, 11, 8],  # SE
        [-1, -1, -1, -1, 5, 6, 7, 4, -1, -1, -1, -1],  # E
        [4, 5, 6, 7, 11, 8, 9, 10, 11, 8, 9, 10],  # SW
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11],  # centre
        [1, 2, 3, 0, 0, 1, 2, 3, 5, 6, 7, 4],  # NE
        [-1, -1, -1, -1, 7, 4, 5, 6, -1, -1, -1, -1],  # W
        [3, 0, 1, 2, 3, 0, 1, 2, 4, 5, 6, 7],  # NW
        [2, 3, 0, 1, -1, -1, -1, -1, 0, 1, 2, 3],  # N
    ],
    dtype=np.int64,
)
# Coordinate fix-up when stepping into that face: bit0 flip x, bit1 flip y, bit2 swap.
# Second index: 0 north-polar faces, 1 equatorial, 2 south-polar.
_NB_SWAP = np.array(
    [[0, 0, 3], [0, 0, 6], [0, 0, 0], [0, 0, 5], [0, 0, 0], [5, 0, 0], [0, 0, 0], [6, 0, 0], [3, 0, 0]],
    dtype=np.int64,
)


def isnsideok(nside, nest=True):
    """True if nside is a valid resolution (power of two >= 1 in NESTED)."""
    try:
        n = int(nside)
    except (TypeError, ValueError):
        return False
    if n != nside or n < 1 or n > (1 << 29):
        return False
    return (n & (n - 1)) == 0 if nest else True


def nside2npix(nside):
    return 12 * int(nside) * int(nside)


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError(f"{npix} is not a valid HEALPix pixel count")
    return nside


def nside2order(nside):
    if not isnsideok(nside, nest=True):
        raise ValueError(f"nside {nside} is not a power of two")
    return int(nside).bit_length() - 1


def _spread_bits(v):
    """Insert a zero bit between the low 32 bits of v (x -> even bit positions)."""
    v = np.asarray(v, dtype=np.uint64)
    v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
    return v


def _compress_bits(v):
    """Inverse of _spread_bits: keep the even bit positions."""
    v = np.asarray(v, dtype=np.uint64) & np.uint64(0x5555555555555555)
    v = (v | (v >> np.uint64(1))) & np.uint64(0x3333333333333333)
    v = (v | (v >> np.uint64(2))) & np.uint64(0x0F0F0F0F0F0F0F0F)
    v = (v | (v >> np.uint64(4))) & np.uint64(0x00FF00FF00FF00FF)
    v = (v | (v >> np.uint64(8))) & np.uint64(0x0000FFFF0000FFFF)
    v = (v | (v >> np.uint64(16))) & np.uint64(0x00000000FFFFFFFF)
    return v


def nest2xyf(nside, ipix):
    """NESTED pixel index -> (x, y, face).  x grows towards NE, y towards NW."""
    order = nside2order(nside)
    ipix = np.asarray(ipix, dtype=np.int64)
    face = ipix >> (2 * order)
    inface = (ipix & ((1 << (2 * order)) - 1)).astype(np.uint64)
    x = _compress_bits(inface).astype(np.int64)
    y = _compress_bits(inface >> np.uint64(1)).astype(np.int64)
    return x, y, face


def xyf2nest(nside, x, y, face):
    order = nside2order(nside)
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    face = np.asarray(face, dtype=np.int64)
    inface = _spread_bits(x.astype(np.uint64)) | (_spread_bits(y.astype(np.uint64)) << np.uint64(1))
    return (face << (2 * order)) + inface.astype(np.int64)


def _xyf2ringinfo(nside, x, y, face):
    """Ring number jr (1..4nside-1), pixels-per-quarter nr, and longitude index tmp."""
    jr = _JRLL[face] * nside - x - y - 1
    nr = np.where(jr < nside, jr, np.where(jr > 3 * nside, 4 * nside - jr, nside))
    tmp = _JPLL[face] * nr + x - y
    tmp = np.where(tmp < 0, tmp + 8 * nr, tmp)
    return jr, nr, tmp


def pix2ang(nside, ipix, nest=True):
    """Pixel centre (theta colatitude, phi longitude) in float64."""
    if not nest:
        ipix = ring2nest(nside, ipix)
    x, y, face = nest2xyf(nside, ipix)
    z, phi = _xyf2zphi(nside, x, y, face)
    return np.arccos(z), phi


def _xyf2zphi(nside, x, y, face):
    nside = int(nside)
    jr, nr, tmp = _xyf2ringinfo(nside, x, y, face)
    npix = 12.0 * nside * nside
    fact2 = 4.0 / npix
    fact1 = (2 * nside) * fact2
    nrf = nr.astype(np.float64)
    z = np.where(
        jr < nside,
        1.0 - nrf * nrf * fact2,
        np.where(jr > 3 * nside, nrf * nrf * fact2 - 1.0, (2 * nside - jr) * fact1),
    )
    phi = (0.25 * np.pi) * tmp / nrf
    return z, phi


def pix2vec(nside, ipix, nest=True):
    """Unit vectors of the pixel centres, shape (..., 3), float64."""
    if not nest:
        ipix = ring2nest(nside, ipix)
    x, y, face = nest2xyf(nside, ipix)
    nside = int(nside)
    jr, nr, tmp = _xyf2ringinfo(nside, x, y, face)
    z, phi = _xyf2zphi(nside, x, y, face)
    # sin(theta): in the caps use the cancellation-free form sqrt(t(2-t)), t = nr^2*fact2
    t = (nr.astype(np.float64) ** 2) * (4.0 / (12.0 * nside * nside))
    polar = (jr < nside) | (jr > 3 * nside)
    sth = np.where(polar, np.sqrt(t * (2.0 - t)), np.sqrt((1.0 - z) * (1.0 + z)))
    return np.stack([sth * np.cos(phi), sth * np.sin(phi), z], axis=-1)


def neighbours(nside, ipix):
    """The 8 NESTED neighbours (SW, W, NW, N, NE, E, SE, S) of each pixel; -1 where a
    neighbour does not exist (3 pixels at each of the 8 valence-3 vertices)."""
    nside = int(nside)
    ipix = np.asarray(ipix, dtype=np.int64)
    x, y, face = nest2xyf(nside, ipix)
    out = np.empty(ipix.shape + (8,), dtype=np.int64)
    for d in range(8):
        nx = x + NB_XOFF[d]
        ny = y + NB_YOFF[d]
        nb = np.full(ipix.shape, 4, dtype=np.int64)
        lo = nx < 0
        hi = nx >= nside
        nx = np.where(lo, nx + nside, np.where(hi, nx - nside, nx))
        nb = nb - lo.astype(np.int64) + hi.astype(np.int64)
        lo = ny < 0
        hi = ny >= nside
        ny = np.where(lo, ny + nside, np.where(hi, ny - nside, ny))
        nb = nb - 3 * lo.astype(np.int64) + 3 * hi.astype(np.int64)
        f = _NB_FACE[nb, face]
        bits = _NB_SWAP[nb, face >> 2]
        fx = np.where(bits & 1, nside - nx - 1, nx)
        fy = np.where(bits & 2, nside - ny - 1, ny)
        sx = np.where(bits & 4, fy, fx)
        sy = np.where(bits & 4, fx, fy)
        valid = f >= 0
        out[..., d] = np.where(valid, xyf2nest(nside, sx, sy, np.where(valid, f, 0)), -1)
    return out


def nest2ring(nside, ipix):
    nside = int(nside)
    x, y, face = nest2xyf(nside, ipix)
    jr, nr, tmp = _xyf2ringinfo(nside, x, y, face)
    kshift = np.where(nr == nside, (jr - nside) & 1, 0)
    jp = (_JPLL[face] * nr + x - y + 1 + kshift) // 2
    jp = np.where(jp > 4 * nr, jp - 4 * nr, jp)
    jp = np.where(jp < 1, jp + 4 * nr, jp)
    ncap = 2 * nside * (nside - 1)
    npix = 12 * nside * nside
    n_before = np.where(
        jr < nside,
        2 * nr * (nr - 1),
        np.where(jr > 3 * nside, npix - 2 * (nr + 1) * nr, ncap + (jr - nside) * 4 * nside),
    )
    return n_before + jp - 1


def ring2nest(nside, ipix):
    """Inverse of nest2ring (by table; host-side helper for small/medium nside)."""
    npix = nside2npix(nside)
    table = np.empty(npix, dtype=np.int64)
    table[nest2ring(nside, np.arange(npix, dtype=np.int64))] = np.arange(npix, dtype=np.int64)
    return table[np.asarray(ipix, dtype=np.int64)]


def query_disc(nside, vec, radius, nest=True):
    """Pixels whose centre lies within `radius` (rad) of `vec` (non-inclusive query)."""
    npix = nside2npix(nside)
    ids = np.arange(npix, dtype=np.int64)
    v = pix2vec(nside, ids, nest=True)
    vec = np.asarray(vec, dtype=np.float64)
    vec = vec / np.linalg.norm(vec)
    sel = ids[v @ vec > np.cos(radius)]
    return sel if nest else np.sort(nest2ring(nside, sel))


def ud_grade_mask_nest(mask, nside_out):
    """``hp.ud_grade`` of a NESTED map restricted to what the reference uses it for
    (healpy_networks.py:76-78,183-185; utils.py:31-34): degrade = mean of the 4^p
    children, upgrade = copy the parent to its 4^p children."""
    mask = np.asarray(mask, dtype=np.float64)
    nside_in = npix2nside(mask.shape[-1])
    if nside_out == nside_in:
        return mask.copy()
    if nside_out < nside_in:
        r = (nside_in // nside_out) ** 2
        return mask.reshape(mask.shape[:-1] + (-1, r)).mean(axis=-1)
    r = (nside_out // nside_in) ** 2
    return np.repeat(mask, r, axis=-1)


def coarsen_indices(indices, p):
    """Parent pixels, p levels up, of a NESTED index set (sorted, unique)."""
    return np.unique(np.asarray(indices, dtype=np.int64) >> (2 * int(p)))


def refine_indices(indices, p):
    """All children, p levels down, of a NESTED index set (sorted)."""
    r = 4 ** int(p)
    base = np.sort(np.asarray(indices, dtype=np.int64))[:, None] * r
    return (base + np.arange(r, dtype=np.int64)[None, :]).ravel()
