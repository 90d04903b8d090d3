"""CPU: the shared-memory addressing of the TMA-staged tile product (csrc/gram_kernels.cu, gram_syrk_tma_kernel).

The copy unit writes boxes of 16 columns x 16 rows with the 128-byte swizzle (16-byte chunk c of row r lands at chunk
c ^ (r % 8) of that row's 128-byte line); the kernel's fragment loads must undo exactly that, take every row of a stage
once, and touch each 8-byte bank slot at most twice per warp load (256 bytes over a 128-byte bank span: two
wavefronts is the minimum).  This replays both sides in numpy.
"""
import numpy as np

GKB, GT, BOX = 16, 128, 2048


def _stage_as_written_by_the_copy_unit(A):
    sm = np.full(GKB * GT, np.nan)
    for g in range(GT // 16):
        for r in range(GKB):
            for c in range(8):
                for e in range(2):
                    byte = g * BOX + r * 128 + ((c ^ (r % 8)) * 16) + e * 8
                    sm[byte // 8] = A[r, g * 16 + c * 2 + e]
    assert not np.isnan(sm).any()
    return sm


def _lane_offset(lane, t, half):
    fk, fc = lane & 3, lane >> 2
    x = (t & 1) * 2 + (fk & 1) + (fk >> 1) * 4          # row % 8, as in the kernel's loff[][]
    y = (fc >> 1) ^ x
    return x * 128 + ((y ^ (half * 4)) * 16) + (fc & 1) * 8, (t >> 1) * 8 + x


def test_fragment_loads_undo_the_swizzle_and_are_two_wavefronts():
    A = np.random.default_rng(0).standard_normal((GKB, GT))
    sm = _stage_as_written_by_the_copy_unit(A)
    for first_group, n_frag in ((0, 4), (2, 4), (4, 4), (6, 4), (0, 8), (4, 8)):   # operand I per wi, operand J per wj
        for t in range(4):
            for i in range(n_frag):
                slots = []
                for lane in range(32):
                    off, krow = _lane_offset(lane, t, i & 1)
                    byte = (first_group + (i >> 1)) * BOX + (t >> 1) * 1024 + off
                    col = first_group * 16 + i * 8 + (lane >> 2)
                    assert sm[byte // 8] == A[krow, col]
                    slots.append((byte % 128) // 8)
                assert max(slots.count(s) for s in set(slots)) == 2


def test_every_row_of_a_stage_is_taken_once():
    rows = sorted(_lane_offset(fk, t, 0)[1] for t in range(4) for fk in range(4))
    assert rows == list(range(GKB))


def test_path_tile_fragment_loads():
    """path_mainloop_tma: boxes of PM (or TN) rows x 16 k-values, one 128-byte line per tile row; a fragment load takes
    the eight rows r0 + fc at k = 4 q + fk."""
    rows = 32
    T = np.random.default_rng(1).standard_normal((rows, 16))
    sm = np.full(rows * 16, np.nan)
    for r in range(rows):
        for c in range(8):
            for e in range(2):
                sm[(r * 128 + ((c ^ (r % 8)) * 16) + e * 8) // 8] = T[r, c * 2 + e]
    assert not np.isnan(sm).any()
    for r0 in range(0, rows, 8):
        for q in range(4):
            slots = []
            for lane in range(32):
                fk, fc = lane & 3, lane >> 2
                koff = fc * 128 + ((((fk >> 1) + 2 * q) ^ fc) * 16) + (fk & 1) * 8   # the kernel's koff[q]
                byte = r0 * 128 + koff
                assert sm[byte // 8] == T[r0 + fc, 4 * q + fk]
                slots.append((byte % 128) // 8)
            assert max(slots.count(s) for s in set(slots)) == 2
