"""Host side of the lane-blocked cell storage order (include/alpha_eigen_b200.h).

Logical cell z = tile*256 + lane*8 + j is stored at tile*256 + (j//2)*64 + lane*2 + (j%2).
Pure NumPy; used once at upload / download, never inside a solve.
"""
from __future__ import annotations

import numpy as np

TILE = 256
LANE_CELLS = 8


def padded_cells(Z: int, ranks: int = 1) -> int:
    """Storage length of a row: a multiple of 256 cells, and of 256*ranks when the cells are sharded."""
    unit = TILE * max(int(ranks), 1)
    return (Z + unit - 1) // unit * unit


def storage_index(Z: int) -> np.ndarray:
    """idx[z] = storage offset of logical cell z, for z in range(Z)."""
    z = np.arange(Z, dtype=np.int64)
    tile, r = z // TILE, z % TILE
    lane, j = r // LANE_CELLS, r % LANE_CELLS
    return tile * TILE + (j // 2) * 64 + lane * 2 + (j % 2)


def to_storage(a: np.ndarray, Z: int, Zp: int | None = None) -> np.ndarray:
    """(..., Z) logical order -> (..., Zp) storage order, zero padded."""
    a = np.asarray(a)
    out = np.zeros(a.shape[:-1] + (Zp if Zp is not None else padded_cells(Z),), dtype=a.dtype)
    out[..., storage_index(Z)] = a
    return out


def from_storage(a: np.ndarray, Z: int) -> np.ndarray:
    """(..., Zp) storage order -> (..., Z) logical order."""
    return np.ascontiguousarray(np.asarray(a)[..., storage_index(Z)])
