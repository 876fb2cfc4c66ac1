"""Tile plan for the fused K-hop lattice kernel (host side, numpy).

On the 8-neighbour HEALPix graph every NESTED block of 4^t pixels is a 2^t x 2^t square of a
base face, and - away from the 8 valence-3 vertices of the base tessellation - its H-ring
neighbourhood is a regular (2^t + 2H)^2 lattice even across face boundaries (the neighbouring
face continues the grid, possibly with its x/y frame rotated by a multiple of 90 degrees).
The fused kernel keeps that lattice in shared memory and applies L~ as a 3x3 stencil with
per-pixel weights, so all K-1 recursion hops happen on chip.

This module embeds each tile's neighbourhood into the lattice (walking the HEALPix neighbour
function ring by ring while tracking the frame rotation), VERIFIES the embedding against the
actual sparsity of L~ (a tile that does not verify - the 24 tiles that touch a valence-3
vertex, or anything unexpected in a user-supplied L - is marked irregular and is processed by
the generic kernels instead), and emits per tile:

  pix [LW*LW] int32   row of L~ sitting at each lattice position (-1: hole / outside the mask)
  w   [LW*LW, 9] f32  stencil weights in the tile frame: 8 directions (SW, W, NW, N, NE, E, SE, S)
                      + centre, i.e. w[p, d] = L~[row(p), row(p + off_d)]

There is no reference code for this (the reference multiplies by a tf.SparseTensor,
utils.py:76); correctness is established by `check_plan` below and the parity tests.
"""

import numpy as np
from scipy import sparse

from . import healpix as hpx

# tile-frame lattice offsets (di, dj) of the 8 directions, i along x, j along y
DI = hpx.NB_XOFF
DJ = hpx.NB_YOFF


class LatticePlan:
    def __init__(self, nside, H, tile_side, LW, tile_face, tile_x0, tile_y0, pix, w, exact, n_rows):
        self.nside, self.H, self.tile_side, self.LW = nside, H, tile_side, LW
        self.tile_face, self.tile_x0, self.tile_y0 = tile_face, tile_x0, tile_y0
        self.pix, self.w, self.n_rows = pix, w, n_rows
        # exact [n_tiles, LW*LW] bool: the H-hop recursion on the lattice gives the true value at this position
        # (holes count as exact).  A tile is `regular` when that holds on all of its own pixels, `usable` when it
        # holds on at least one of them (the others are recomputed by the generic path, see irregular_rows).
        self.exact = exact
        own = self.own_mask()
        has = pix[:, own] >= 0
        self.regular = (exact[:, own] | ~has).all(axis=1)
        self.usable = (exact[:, own] & has).any(axis=1)

    @property
    def n_tiles(self):
        return self.pix.shape[0]

    def own_mask(self):
        """Boolean [LW*LW]: lattice positions of the tile's own pixels."""
        H, LW, T = self.H, self.LW, self.tile_side
        m = np.zeros((LW, LW), dtype=bool)
        m[H : H + T, H : H + T] = True  # [j, i]
        return m.ravel()

    def irregular_rows(self):
        """Rows of L~ whose H-hop value the lattice does not reproduce (to be handled by the generic path): the own
        pixels within reach of a valence-3 vertex - a 4 x 4 corner of each of the 24 tiles around them for H = 4 -
        and every own pixel of a tile that is not launched at all."""
        own = self.own_mask()
        rows, ex = self.pix[:, own], self.exact[:, own] & self.usable[:, None]
        return np.sort(rows[(rows >= 0) & ~ex])

    def lattice_rows(self):
        """Rows of L~ whose result comes from the lattice launch (the complement of irregular_rows)."""
        own = self.own_mask()
        rows, ex = self.pix[:, own], self.exact[:, own] & self.usable[:, None]
        return np.sort(rows[(rows >= 0) & ex])


def neighbour_table(nside):
    """The 8 NESTED neighbours of every pixel of the sphere, [npix, 8] (one vectorised pass; the plan builder looks
    pixels up in it ~7 times per pixel, which used to be 60 % of its run time)."""
    return hpx.neighbours(nside, np.arange(hpx.nside2npix(nside), dtype=np.int64))


def graph_is_healpix8(Lt, nside, indices, nb_table=None):
    """True if every off-diagonal non-zero of Lt connects HEALPix 8-neighbours (within `indices`)."""
    npix = hpx.nside2npix(nside)
    indices = np.asarray(indices, dtype=np.int64)
    M = len(indices)
    if Lt.shape != (M, M):
        return False
    lut = np.full(npix + 1, -1, dtype=np.int64)
    lut[indices] = np.arange(M)
    nb = hpx.neighbours(nside, indices) if nb_table is None else nb_table[indices]
    nbrow = lut[nb]  # -1 (missing) indexes the sentinel slot -> -1
    coo = sparse.coo_matrix(Lt)
    off = coo.row != coo.col
    r, c = coo.row[off], coo.col[off]
    return bool(np.all((nbrow[r] == c[:, None]).any(axis=1)))


def build_lattice_plan(Lt, nside, indices, H, tile_order=4):
    """Lt: prepared L~ (scipy sparse, M x M, rows ordered like `indices`, NESTED pixel ids, sorted).
    H: halo rings = K - 1.  Returns a LatticePlan, or None if the graph is not an 8-neighbour
    HEALPix graph or the tile does not fit the resolution."""
    nside = int(nside)
    order = hpx.nside2order(nside)
    if tile_order > order or H < 1:
        return None
    indices = np.asarray(indices, dtype=np.int64)
    if np.any(np.diff(indices) <= 0):
        return None
    NB = neighbour_table(nside)
    if not graph_is_healpix8(Lt, nside, indices, NB):
        return None
    npix = hpx.nside2npix(nside)
    M = len(indices)
    T = 1 << tile_order
    LW = T + 2 * H
    lut = np.full(npix + 1, -1, dtype=np.int64)  # pixel -> row, sentinel slot for pixel -1
    lut[indices] = np.arange(M)

    # tiles: NESTED blocks of T*T pixels holding at least one selected pixel
    tile_ids = np.unique(indices >> (2 * tile_order))
    nt = len(tile_ids)
    tx, ty, tf = hpx.nest2xyf(nside, tile_ids << (2 * tile_order))  # lower-left pixel of each block

    pixel = np.full((nt, LW, LW), -2, dtype=np.int64)  # HEALPix pixel at [tile, j, i]; -1 hole, -2 unknown
    rot = np.zeros((nt, LW, LW), dtype=np.int64)
    jj, ii = np.meshgrid(np.arange(T), np.arange(T), indexing="ij")
    own_pix = hpx.xyf2nest(nside, tx[:, None, None] + ii[None], ty[:, None, None] + jj[None], tf[:, None, None])
    pixel[:, H : H + T, H : H + T] = own_pix

    def nbr_of(p):
        return NB[p]

    # grow ring by ring: position P at ring r is reached from Q = P clamped one step towards the box
    for r in range(1, H + 1):
        lo, hi = H - r, H + T + r - 1
        ring = [(j, i) for j in range(lo, hi + 1) for i in range(lo, hi + 1) if j in (lo, hi) or i in (lo, hi)]
        Pj = np.array([p[0] for p in ring])
        Pi = np.array([p[1] for p in ring])
        Qj = np.clip(Pj, lo + 1, hi - 1)
        Qi = np.clip(Pi, lo + 1, hi - 1)
        d_tile = np.array([int(np.where((DI == pi - qi) & (DJ == pj - qj))[0][0])
                           for pj, pi, qj, qi in zip(Pj, Pi, Qj, Qi)])
        qpix = pixel[:, Qj, Qi]  # [nt, nring]
        qrot = rot[:, Qj, Qi]
        valid = qpix >= 0
        qn = nbr_of(np.where(valid, qpix, 0))  # [nt, nring, 8]
        d_face = (d_tile[None, :] + qrot) % 8
        ppix = np.take_along_axis(qn, d_face[..., None], axis=2)[..., 0]
        ppix = np.where(valid, ppix, -1)
        # rotation of P's face frame relative to the tile frame
        pv = ppix >= 0
        pn = nbr_of(np.where(pv, ppix, 0))
        back = np.argmax(pn == qpix[..., None], axis=2)
        prot = (back - (d_tile[None, :] + 4)) % 8
        pixel[:, Pj, Pi] = ppix
        rot[:, Pj, Pi] = np.where(pv, prot, 0)

    pixel = pixel.reshape(nt, LW * LW)
    rot = rot.reshape(nt, LW * LW)
    row = lut[pixel]  # -1 for holes and for pixels outside the selection

    # L~ by face-frame direction: ldir[r, d] = L~[r, row(nb[pix_r][d])], ldir[r, 8] = L~[r, r]
    csr = sparse.csr_matrix(Lt)
    csr.sort_indices()
    keys = np.repeat(np.arange(M, dtype=np.int64), np.diff(csr.indptr)) * M + csr.indices.astype(np.int64)
    nbrow = lut[NB[indices]]  # [M, 8]
    cols = np.concatenate([nbrow, np.arange(M)[:, None]], axis=1)  # [M, 9]
    q = np.arange(M, dtype=np.int64)[:, None] * M + np.where(cols >= 0, cols, 0)
    loc = np.searchsorted(keys, q)
    loc = np.minimum(loc, len(keys) - 1)
    found = (keys[loc] == q) & (cols >= 0)
    ldir = np.where(found, csr.data[loc], 0.0).astype(np.float32)  # [M, 9]

    # stencil weights in the tile frame
    has = row >= 0
    rsafe = np.where(has, row, 0)
    w = np.zeros((nt, LW * LW, 9), dtype=np.float32)
    for d in range(8):
        w[:, :, d] = np.where(has, ldir[rsafe, (d + rot) % 8], 0.0)
    w[:, :, 8] = np.where(has, ldir[rsafe, 8], 0.0)

    # verification, per lattice position: the lattice neighbour in direction d must be the true graph neighbour in
    # that direction wherever L~ has one (a direction without a true neighbour carries weight 0: whatever sits there is
    # harmless).  A position that fails is `bad`: its value is wrong from hop 1 on ...
    pos_j, pos_i = np.divmod(np.arange(LW * LW), LW)
    inner = (pos_j >= 1) & (pos_j <= LW - 2) & (pos_i >= 1) & (pos_i <= LW - 2)
    bad = np.zeros((nt, LW * LW), dtype=bool)
    present = np.zeros((nt, LW * LW, 8), dtype=bool)
    tgts = []
    nbr_true_all = NB[np.where(pixel >= 0, pixel, 0)]  # [nt, P, 8] face-frame order
    for d in range(8):
        tgt = (pos_j + DJ[d]) * LW + (pos_i + DI[d])
        tgt = np.where(inner, tgt, 0)
        tgts.append(tgt)
        true_pix = np.take_along_axis(nbr_true_all, ((d + rot) % 8)[..., None], axis=2)[..., 0]
        true_row = lut[np.where(true_pix >= 0, true_pix, npix)]  # a neighbour outside the selection == absent
        present[:, :, d] = has & (true_row >= 0)
        bad |= present[:, :, d] & inner[None, :] & (row[:, tgt] != true_row)
    del nbr_true_all
    # ... and the error travels one position per hop: exact_h(p) = p is computed at hop h (inside the shrinking region),
    # not bad, and every true neighbour was exact at hop h - 1.  Away from the valence-3 vertices nothing is bad and the
    # own pixels (depth >= H) are exact by construction, so only the tiles with a bad position are propagated.
    depth = np.minimum(np.minimum(pos_j, LW - 1 - pos_j), np.minimum(pos_i, LW - 1 - pos_i))
    exact = np.broadcast_to(depth >= H, (nt, LW * LW)) | ~has
    hit = np.flatnonzero(bad.any(axis=1))
    if len(hit):
        ex = np.ones((len(hit), LW * LW), dtype=bool)  # hop 0: the gathered input
        for h in range(1, H + 1):
            nxt = (depth >= h)[None, :] & ~bad[hit]
            for d in range(8):
                nxt &= ~present[hit][:, :, d] | ex[:, tgts[d]]
            ex = nxt | ~has[hit]
        exact = exact.copy()
        exact[hit] = ex
    del present
    # every row must be owned by exactly one tile
    plan = LatticePlan(nside, H, T, LW, tf, tx, ty, row.astype(np.int32), w, exact, M)
    own_rows = plan.pix[:, plan.own_mask()]
    owned = np.sort(own_rows[own_rows >= 0])
    if len(owned) != M or np.any(owned != np.arange(M)):
        return None
    return plan


def check_plan(plan, Lt, rng=None, n_tiles=None):
    """Numerical self-check: one application of the stencil on random data equals Lt @ x on the
    computed region of every (sampled) regular tile.  Returns the max abs error.  (The H-hop statement, including
    the exact pixels of the tiles at the valence-3 vertices: tests/test_lattice_cpu.py.)"""
    rng = rng or np.random.default_rng(0)
    M = plan.n_rows
    x = rng.standard_normal(M)
    y = sparse.csr_matrix(Lt, dtype=np.float64) @ x
    LW = plan.LW
    tiles = np.flatnonzero(plan.regular)
    if n_tiles is not None and len(tiles) > n_tiles:
        tiles = rng.choice(tiles, n_tiles, replace=False)
    pos_j, pos_i = np.divmod(np.arange(LW * LW), LW)
    inner = (pos_j >= 1) & (pos_j <= LW - 2) & (pos_i >= 1) & (pos_i <= LW - 2)
    worst = 0.0
    pix = plan.pix[tiles]
    xv = np.where(pix >= 0, x[np.where(pix >= 0, pix, 0)], 0.0)  # [t, P]
    acc = plan.w[tiles][:, :, 8].astype(np.float64) * xv
    for d in range(8):
        tgt = np.where(inner, (pos_j + DJ[d]) * LW + (pos_i + DI[d]), 0)
        acc += plan.w[tiles][:, :, d].astype(np.float64) * xv[:, tgt]
    sel = (pix >= 0) & inner[None, :]
    ref = y[np.where(pix >= 0, pix, 0)]
    worst = float(np.abs(np.where(sel, acc - ref, 0.0)).max()) if sel.any() else 0.0
    return worst


def make_patches(csr, closure, rows_irr, max_rows=1024):
    """The irregular sub-problem as connected patches (ds_plan_attach_patches): the closure of the irregular rows falls
    apart into one neighbourhood per valence-3 vertex (8 x 189 rows for H = 4 on the full sphere); each patch carries
    its rows, L~ restricted to them as a 9-wide ELL with patch-local columns, and the patch-local rows that are wanted.
    None when a row has more than 9 entries or a patch is too large for one thread block."""
    from scipy.sparse.csgraph import connected_components

    sub = sparse.csr_matrix(csr[closure][:, closure])
    sub.sort_indices()
    n = sub.shape[0]
    width = np.diff(sub.indptr)
    if n == 0 or width.max() > 9:
        return None
    n_comp, label = connected_components(sub, directed=False)
    order = np.argsort(label, kind="stable")  # closure-local rows grouped by patch
    counts = np.bincount(label, minlength=n_comp)
    if counts.max() > max_rows:
        return None
    row_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    local = np.empty(n, dtype=np.int64)  # patch-local index of every closure-local row
    local[order] = np.arange(n) - np.repeat(row_ptr[:-1], counts)
    ell_col = np.full((n, 9), -1, dtype=np.int32)
    ell_val = np.zeros((n, 9), dtype=np.float32)
    r_of = np.repeat(np.arange(n), width)
    slot = np.arange(sub.nnz) - np.repeat(sub.indptr[:-1], width)
    ell_col[r_of, slot] = local[sub.indices]
    ell_val[r_of, slot] = sub.data
    own_cl = np.searchsorted(closure, rows_irr)  # closure-local ids of the wanted rows
    own_cl = own_cl[np.argsort(label[own_cl], kind="stable")]
    own_counts = np.bincount(label[own_cl], minlength=n_comp)
    return {
        "n_patches": int(n_comp),
        "row_ptr": np.ascontiguousarray(row_ptr),
        "rows": np.ascontiguousarray(closure[order], dtype=np.int32),
        "ell_col": np.ascontiguousarray(ell_col[order]),
        "ell_val": np.ascontiguousarray(ell_val[order]),
        "own_ptr": np.ascontiguousarray(np.concatenate([[0], np.cumsum(own_counts)]).astype(np.int32)),
        "own_local": np.ascontiguousarray(local[own_cl], dtype=np.int32),
    }


def make_payload(Lt, nside, indices, H, tile_order=4):
    """Everything ds_plan_attach_lattice needs, as contiguous numpy arrays (or None if the fused path does
    not apply): tables of every tile the lattice serves + the compact generic sub-problem that recomputes the rows it
    gets wrong (plan.irregular_rows(): the own pixels within H hops' reach of a valence-3 vertex; the kernels write
    all own pixels of a launched tile and the sub-problem's scatter, which runs after them, overwrites those rows)."""
    plan = build_lattice_plan(Lt, nside, indices, H, tile_order)
    if plan is None or not plan.usable.any():
        return None
    csr = sparse.csr_matrix(Lt)
    # the backward pass re-uses the same stencil for L~^T: symmetric up to rounding (same rule as ds_plan_create_coo)
    if abs(csr - csr.T).max() > 1e-6 * abs(csr).max():
        return None
    M = plan.n_rows
    rows_irr = plan.irregular_rows()
    if len(rows_irr):
        mask = np.zeros(M, dtype=bool)
        mask[rows_irr] = True
        pattern = sparse.csr_matrix((np.ones(csr.nnz, dtype=np.float32), csr.indices, csr.indptr), shape=csr.shape)
        for _ in range(H):
            mask = mask | ((pattern @ mask.astype(np.float32)) > 0)
        closure = np.flatnonzero(mask)
        sub = sparse.coo_matrix(csr[closure][:, closure])
        sub_idx = np.ascontiguousarray(np.column_stack((sub.row, sub.col)).astype(np.int64))
        sub_val = np.ascontiguousarray(sub.data.astype(np.float32))
        own_sub = np.searchsorted(closure, rows_irr).astype(np.int32)
        patches = make_patches(csr, closure, rows_irr)
    else:
        patches = None
        closure = np.zeros(0, dtype=np.int64)
        sub_idx, sub_val = np.zeros((0, 2), np.int64), np.zeros(0, np.float32)
        own_sub = np.zeros(0, dtype=np.int32)
    reg = plan.usable
    return {
        "n_tiles": int(reg.sum()), "LW": int(plan.LW), "H": int(plan.H), "T": int(plan.tile_side),
        "pix": np.ascontiguousarray(plan.pix[reg], dtype=np.int32),
        "w": np.ascontiguousarray(plan.w[reg], dtype=np.float32),
        "sub": (sub_idx, sub_val, int(len(closure))),
        "closure_rows": np.ascontiguousarray(closure, dtype=np.int32),
        "own_sub": np.ascontiguousarray(own_sub, dtype=np.int32),
        "n_irregular_tiles": int((~plan.regular).sum()),
        "lattice_rows": plan.lattice_rows(),
        "patches": patches,
    }
