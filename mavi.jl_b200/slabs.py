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
    """Global cell column of x (0-based; -1 outside the grid) by the device's own rule: `mavi_cells_of_points` evaluates
    update_particle_chunk! (src/chunks.jl:120-147: Base.div as trunc of the exact quotient, clamp of index n+1) on the
    host with the arithmetic of the device kernels, so host routing and device binning cannot disagree."""
    lib = capi.load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    pts = np.zeros((len(x), 2))
    pts[:, 0] = x
    p = capi.MaviParams()
    p.struct_size = C.sizeof(capi.MaviParams)
    p.dtype = capi.F64
    p.grid_bl[0], p.grid_bl[1] = float(grid_bl_x), -0.5   # one row of height 1 around y = 0
    p.grid_len, p.grid_h = float(grid_len), 1.0
    p.num_cols, p.num_rows = int(num_cols), 1
    out = np.empty(len(x), dtype=np.int32)
    st = lib.mavi_cells_of_points(C.byref(p), pts.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p))
    if st != capi.OK:
        raise capi.MaviError(st, "mavi_cells_of_points")
    return out.astype(np.int64)


def cells_of_points(pos, space_bbox, num_cols: int, num_rows: int):
    """0-based linear cell ids (col * num_rows + row, row 0 = top; -1 outside) of (n, 2) points — the reference's binning
    evaluated on the host with the device's arithmetic (include/mavi.h: mavi_cells_of_points)."""
    lib = capi.load_library()
    pos = np.ascontiguousarray(pos)
    f32 = pos.dtype == np.float32
    if not f32:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
    p = capi.MaviParams()
    p.struct_size = C.sizeof(capi.MaviParams)
    p.dtype = capi.F32 if f32 else capi.F64
    p.grid_bl[0], p.grid_bl[1] = float(space_bbox.bottom_left[0]), float(space_bbox.bottom_left[1])
    p.grid_len, p.grid_h = float(space_bbox.length), float(space_bbox.height)
    p.num_cols, p.num_rows = int(num_cols), int(num_rows)
    out = np.empty(len(pos), dtype=np.int32)
    st = lib.mavi_cells_of_points(C.byref(p), pos.ctypes.data_as(C.c_void_p), len(pos), out.ctypes.data_as(C.c_void_p))
    if st != capi.OK:
        raise capi.MaviError(st, "mavi_cells_of_points")
    return out


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
