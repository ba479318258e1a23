"""Host-side logic of the x-slab domain decomposition (one process per GPU).

The reference has no distributed path; its `Threaded` mode splits the same pair loop over cell COLUMNS
(src/integration.jl:159-194).  Here cell columns are split contiguously over ranks: the first `num_cols % world`
ranks own one extra column.  This module mirrors the split used inside libmavi_cuda.so (slab.cu: slab_columns) and
routes particles to their owner.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


def slab_columns(num_cols: int, world: int, rank: int):
    """(first global column, number of columns) owned by `rank`."""
    base, rem = divmod(num_cols, world)
    m = base + (1 if rank < rem else 0)
    lo = rank * base + min(rank, rem)
    return lo, m


def column_of(x, grid_bl_x: float, grid_len: float, num_cols: int):
    """Global cell column of x (0-based) with the reference's clamp of index n+1 (src/chunks.jl:129-142).  Extended
    precision resolves the rounded-multiple cases of Base.div; a mis-routed particle would be rejected loudly by the
    device build (MAVI_ERR_OUT_OF_GRID), never silently accepted."""
    cl = np.longdouble(grid_len) / np.longdouble(num_cols)
    q = np.trunc((np.asarray(x, dtype=np.longdouble) - np.longdouble(grid_bl_x)) / cl).astype(np.int64)
    return np.where(q == num_cols, num_cols - 1, q)


def owner_of_column(col, num_cols: int, world: int):
    base, rem = divmod(num_cols, world)
    col = np.asarray(col)
    split = rem * (base + 1)  # columns below `split` belong to the ranks with one extra column
    return np.where(col < split, col // (base + 1), rem + (col - split) // max(base, 1)).astype(np.int64)


def partition(pos, geometry_cfg, num_cols: int, world: int):
    """Owner rank of every particle."""
    col = column_of(pos[:, 0], geometry_cfg.bottom_left[0], geometry_cfg.length, num_cols)
    return owner_of_column(col, num_cols, world)


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId made by the library (rank 0 calls this; the host broadcasts it to all ranks)."""
    lib = capi.load_library()
    buf = C.create_string_buffer(128)
    st = lib.mavi_nccl_unique_id(buf)
    if st != capi.OK:
        raise capi.MaviError(st, "mavi_nccl_unique_id")
    return buf.raw
