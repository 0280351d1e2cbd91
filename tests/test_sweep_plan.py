"""Work assignment of the persistent sweep kernel, replayed on the host (ae_sweep_plan runs the kernel's own
Cursor / unit_row / plan_grid code): every (row block, tile) of a sweep set must be visited exactly once, in
sweep order inside a row, with hand-offs only between the CTA pairs the mailbox links; gang members must walk
the same cells in lock step (that is what turns the re-reads of hs/q/cdt into L2 hits).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from alpha_eigen_b200 import _lib


def plan(A, n_neg, G, Zp, resident):
    lib = _lib.load()
    grid, gang = C.c_int32(0), C.c_int32(0)
    n = lib.ae_sweep_plan(A, n_neg, G, Zp, resident, -1, C.byref(grid), C.byref(gang), None, None, 0)
    assert n == 0
    out = []
    for cta in range(grid.value):
        w = lib.ae_sweep_plan(A, n_neg, G, Zp, resident, cta, None, None, None, None, 0)
        rb = np.zeros(w, dtype=np.int32)
        ti = np.zeros(w, dtype=np.int32)
        got = lib.ae_sweep_plan(A, n_neg, G, Zp, resident, cta, None, None, rb.ctypes.data_as(_lib.c_ip),
                                ti.ctypes.data_as(_lib.c_ip), w)
        assert got == w
        out.append((rb, ti))
    return grid.value, gang.value, out


def check_plan(A, n_neg, G, Zp, resident):
    lib = _lib.load()
    grid, gs, ctas = plan(A, n_neg, G, Zp, resident)
    nt = Zp // 256
    P = lib.ae_partial_count(A, n_neg)
    rows = G * P
    seen = np.zeros((rows, nt), dtype=np.int32)
    owner = np.full((rows, nt), -1, dtype=np.int32)
    step = np.zeros((rows, nt), dtype=np.int64)
    for cta, (rb, ti) in enumerate(ctas):
        assert rb.min(initial=0) >= 0 and rb.max(initial=0) < rows and ti.max(initial=0) < nt
        np.add.at(seen, (rb, ti), 1)
        owner[rb, ti] = cta
        step[rb, ti] = np.arange(len(rb))
        # inside a CTA a row's tiles come in sweep order, back to back
        change = np.flatnonzero(np.diff(rb) != 0) + 1
        for a, b in zip(np.r_[0, change], np.r_[change, len(rb)]):
            assert np.all(np.diff(ti[a:b]) == 1)
        assert len(set(rb[np.r_[0, change]].tolist())) == len(np.r_[0, change]) if len(rb) else True
    assert np.all(seen == 1), "every (row block, tile) exactly once"
    sizes = np.array([len(rb) for rb, _ in ctas])
    assert sizes.max() - sizes.min() <= 1, "equal ranges"
    # hand-offs: where a row changes owner, the consumer is the same member of the next gang; the producing piece
    # is the first thing its CTA does (so the consumer's wait is short), the consuming piece the last thing
    for r in range(rows):
        cuts = np.flatnonzero(np.diff(owner[r]) != 0)
        for c in cuts:
            prod, cons = owner[r, c], owner[r, c + 1]
            assert cons == prod + gs
            prb, pti = ctas[prod]
            crb, cti = ctas[cons]
            first_piece_end = np.flatnonzero(np.r_[np.diff(prb) != 0, True])[0]
            assert prb[0] == r and pti[first_piece_end] == c, "producer handles the handed-off piece first"
            last_piece_start = np.r_[0, np.flatnonzero(np.diff(crb) != 0) + 1][-1]
            assert crb[-1] == r and cti[last_piece_start] == c + 1, "consumer handles the continued piece last"
    # gang members: same group, direction and tile at every step
    for gang in range(grid // gs):
        rb0, ti0 = ctas[gang * gs]
        for mb in range(1, gs):
            rb1, ti1 = ctas[gang * gs + mb]
            assert np.array_equal(ti0, ti1)
            assert np.array_equal(rb0 // P, rb1 // P)                                   # same group
            nb_neg = P // 2
            assert np.array_equal(rb0 % P < nb_neg, rb1 % P < nb_neg)                   # same direction
            assert np.all(rb1 - rb0 == mb)
    return grid, gs


def test_plan_c4_uses_gangs_of_four():
    # BASELINE configs[3] on one GPU: 8 + 8 blocks of 16 angles per group, 148 persistent CTAs
    Zp = _lib.load().ae_padded_cells(1_000_000)
    grid, gs = check_plan(256, 128, 70, Zp, 148)
    assert (grid, gs) == (148, 4)


def test_plan_c3_pairs_the_two_blocks_of_a_direction():
    Zp = _lib.load().ae_padded_cells(100_000)
    grid, gs = check_plan(64, 32, 70, Zp, 148)
    assert (grid, gs) == (148, 2)


@pytest.mark.parametrize("A,n_neg,G,cells,resident", [
    (4, 2, 1, 450, 296),            # K-P short: fewer tiles than CTAs
    (196, 98, 1, 1800, 148),        # K-P long
    (6, 3, 5, 5000, 296),           # ragged blocks
    (7, 2, 3, 70_000, 296),         # unequal directions: no gangs
    (64, 32, 12, 300_000, 148),     # gangs, ranges cutting rows
    (128, 64, 9, 40_000, 148),      # four blocks per direction
    (64, 32, 37, 8000, 148),        # shapes of test_ganged_sweep_matches_oracle
    (96, 48, 26, 5000, 148),
    (32, 16, 70, 125_000, 148),     # C4 rows of one rank at 8 GPUs: one block per direction
])
def test_plan_covers_every_tile_once(A, n_neg, G, cells, resident):
    Zp = _lib.load().ae_padded_cells(cells)
    check_plan(A, n_neg, G, Zp, resident)


def test_plan_rejects_bad_arguments():
    lib = _lib.load()
    assert lib.ae_sweep_plan(0, 0, 1, 256, 148, 0, None, None, None, None, 0) == -1
    assert lib.ae_sweep_plan(4, 2, 1, 255, 148, 0, None, None, None, None, 0) == -1


@pytest.mark.parametrize("force,expect", [("1", 1), ("2", 2), ("8", 8), ("3", 1)])
def test_forced_gang_size(force, expect):
    """AE_SWEEP_GANG overrides the automatic choice (read once per process, hence the subprocess); a size that does not
    divide the angle blocks per direction falls back to 1.  Coverage must hold for every forced size."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests.test_sweep_plan import check_plan\n"
            "from alpha_eigen_b200 import _lib\n"
            "print(check_plan(256, 128, 70, _lib.load().ae_padded_cells(200_000), 148)[1])\n") % root
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, AE_SWEEP_GANG=force))
    assert out.returncode == 0, out.stderr[-2000:]
    assert int(out.stdout.strip().splitlines()[-1]) == expect
