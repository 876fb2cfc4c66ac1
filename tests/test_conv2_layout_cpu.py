"""Index arithmetic of the register-resident fused kernel (csrc/ds_lattice_conv2.cu), restated in numpy: the
de-interleaved exchange layout, the perimeter offsets of a 3x3 block, the UMMA operand rows <-> lattice positions of the
epilogue, the gather mapping of the loader warps and the stencil direction table.  CPU only - guards the constants
the kernel hard-codes."""
import re
import os

import numpy as np

from deepsphere import healpix as hpx

LW, H, T = 24, 4, 16
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "deepsphere-cosmo-tf2_b200", "csrc", "ds_lattice_conv2.cu")).read()


def pos(j, c, BC=3):
    """float4 index inside a plane of lattice (row j, column c): one pad row above, the columns of the LW / BC blocks
    de-interleaved (kernel: slot_of_col)."""
    NB = LW // BC
    return (j + 1) * LW + (c % BC) * NB + c // BC


def test_layout_is_a_bijection_and_quarter_warps_are_contiguous():
    for BC in (3, 6):
        NB = LW // BC
        p = np.array([[pos(j, c, BC) for c in range(LW)] for j in range(LW)])
        assert len(np.unique(p)) == LW * LW and p.min() == LW and p.max() == (LW + 1) * LW - 1
        # the column blocks cb = 0..NB-1 of one block row and one in-block column cc are NB consecutive positions
        for j in range(LW):
            for cc in range(BC):
                q = [pos(j, BC * cb + cc, BC) for cb in range(NB)]
                assert q == list(range(q[0], q[0] + NB))
        # the inverse map the epilogue and the gather warps use (kernel: col_of_slot)
        for slot in range(LW):
            c = (slot % NB) * BC + slot // NB
            assert pos(0, c, BC) == LW + slot
    assert "return (c % C2_BC) * C2_NB + c / C2_BC;" in SRC and "return (p % C2_NB) * C2_BC + p / C2_NB;" in SRC


def test_perimeter_offsets_of_a_block():
    """Columns -1, 0, .., BC relative to the block's column 0 sit at (BC-1) NB - 1, 0, NB, .., (BC-1) NB, +1 (kernel
    lambda `co`); 3-column blocks: +15, +0, +8, +16, +1."""
    assert "return k == 0 ? (C2_BC - 1) * C2_NB - 1 : (k == C2_BC + 1 ? 1 : (k - 1) * C2_NB);" in SRC
    for BC in (3, 6):
        NB = LW // BC
        co = [(BC - 1) * NB - 1 if k == 0 else (1 if k == BC + 1 else (k - 1) * NB) for k in range(BC + 2)]
        if BC == 3:
            assert co == [15, 0, 8, 16, 1]
        for cb in range(1, NB - 1):  # interior column blocks: all columns exist
            base = pos(5, BC * cb, BC)
            for k, dc in enumerate(range(-1, BC + 1)):
                assert pos(5, BC * cb + dc, BC) == base + co[k]
        # the wrap-around of the outermost blocks stays inside the row (only ever feeds don't-care ring positions)
        assert LW <= pos(0, 0, BC) + co[0] < 2 * LW and LW <= pos(0, LW - BC, BC) + 1 < 2 * LW + 1


def test_operand_rows_map_back_to_lattice_positions():
    """UMMA A operand = positions of lattice rows 4..19 (384 rows, 3 M-tiles); the epilogue maps accumulator row m to
    (row, column) with j = 4 + m / 24, p = m % 24, c = 3 * (p & 7) + (p >> 3)."""
    start = (H + 1) * LW  # plane offset of lattice row 4, position 0
    seen = set()
    for m in range(3 * 128):
        j, p = H + m // LW, m % LW
        c = 3 * (p & 7) + (p >> 3)
        assert pos(j, c) == start + m
        if H <= c < H + T:
            seen.add((j, c))
    assert seen == {(j, c) for j in range(H, H + T) for c in range(H, H + T)}  # every own pixel exactly once


def test_loader_mapping_covers_every_position_once():
    """96 gather threads: t -> (q = t & 1, pl0 = t >> 1), in-row position pl0 % 24, rows 2k + pl0 / 24, k < 12."""
    hit = np.zeros((2, LW, LW), dtype=int)
    for t in range(96):
        q, pl0 = t & 1, t >> 1
        inpos, r0 = pl0 % LW, pl0 // LW
        col = 3 * (inpos & 7) + (inpos >> 3)
        for k in range(LW // 2):
            row = 2 * k + r0
            assert pos(row, col) == (row + 1) * LW + inpos
            hit[q, row, col] += 1
    assert (hit == 1).all()


def test_direction_table_matches_the_plan_builder():
    """dir_of(drow, dcol) of the kernel == index of (di = dcol, dj = drow) in lattice.py's (NB_XOFF, NB_YOFF), 8 = centre."""
    def dir_of(dr, dc):
        if dr == 0:
            return 0 if dc < 0 else (4 if dc > 0 else 8)
        if dr > 0:
            return 1 if dc < 0 else (2 if dc == 0 else 3)
        return 5 if dc > 0 else (6 if dc == 0 else 7)

    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr == 0 and dc == 0:
                assert dir_of(dr, dc) == 8
                continue
            d = int(np.flatnonzero((np.asarray(hpx.NB_XOFF) == dc) & (np.asarray(hpx.NB_YOFF) == dr))[0])
            assert dir_of(dr, dc) == d
    # and the C++ source spells the same table
    assert "dr == 0 ? (dc < 0 ? 0 : (dc > 0 ? 4 : 8))" in SRC


def test_block_recursion_emulation_matches_oracle_and_contains_garbage():
    """numpy emulation of what lattice_conv2_kernel computes on one tile: every one of the 24 x 24 positions is updated on
    every hop from its 3 x 3 lattice neighbourhood with the plan's weights (Chebyshev: weights 2 L~, hop 1 halved,
    T_k = (2L~) T_{k-1} - T_{k-2}); reads beyond the lattice hit the pad rows (NaN here) or wrap inside the row exactly like
    the kernel's position arithmetic.  Claim guarded: the don't-care values (NaN) never reach the tile's own pixels, whose
    T_1..T_4 equal the sparse-matrix oracle."""
    from deepsphere import lattice, utils
    from deepsphere.graph import SphereHealpix
    from scipy import sparse

    nside = 32
    g = SphereHealpix(nside, k=8)
    Lt = utils.rescale_L(sparse.csr_matrix(g.L, dtype=np.float64), lmax=1.9, scale=0.75)
    plan = lattice.build_lattice_plan(Lt, nside, np.arange(12 * nside * nside), H)
    assert plan is not None and plan.LW == LW
    tile = int(np.flatnonzero(plan.regular)[7])
    pix = plan.pix[tile].reshape(LW, LW)
    w = plan.w[tile].astype(np.float64).reshape(LW, LW, 9)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(Lt.shape[0])
    # oracle basis
    basis = [x, Lt @ x]
    for _ in range(2, H + 1):
        basis.append(2 * (Lt @ basis[-1]) - basis[-2])
    # emulation on the padded lattice: row index j + 1, pad rows 0 and LW + 1 hold NaN
    def padded(v):
        out = np.full((LW + 2, LW), np.nan)
        out[1:-1] = v
        return out

    offs = {0: (0, -1), 1: (1, -1), 2: (1, 0), 3: (1, 1), 4: (0, 1), 5: (-1, 1), 6: (-1, 0), 7: (-1, -1), 8: (0, 0)}
    cur = np.where(pix >= 0, x[np.where(pix >= 0, pix, 0)], 0.0)
    old = None
    own = np.zeros((LW, LW), dtype=bool)
    own[H:H + T, H:H + T] = True
    for s in range(1, H + 1):
        P = padded(cur)
        new = np.zeros((LW, LW))
        for j in range(LW):
            for c in range(LW):
                acc = 0.0
                for d, (dr, dc) in offs.items():
                    cc = c + dc
                    if cc < 0:
                        cc = 22      # kernel: position +15 of the first column block -> column 22 of the same row
                    elif cc >= LW:
                        cc = 1       # kernel: position +1 of the last column block -> column 1
                    acc += 2.0 * w[j, c, d] * P[j + 1 + dr, cc]
                new[j, c] = acc
        if s == 1:
            new = 0.5 * new
        else:
            new = new - old
        old, cur = cur, new
        got = cur[own]
        assert np.isfinite(got).all(), f"don't-care values reached the own pixels at hop {s}"
        ref = basis[s][pix[own]]
        # the plan stores the stencil weights in fp32 (as the kernel reads them): agreement to fp32 rounding
        assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max()), (s, np.abs(got - ref).max())
    # and the don't-care region really is contaminated (the test would be vacuous otherwise)
    assert not np.isfinite(cur).all()
