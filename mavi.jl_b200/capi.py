"""ctypes binding of the C ABI declared in include/mavi.h (libmavi_cuda.so).

This is the exact set of entry points the Julia glue (`julia/MaviCUDA.jl`) `ccall`s; Julia is not
installed in this image, so the Python host mirror exercises the same ABI.  There is NO CPU fallback:
if the CUDA library is missing or fails to load, `load_library()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libmavi_cuda.so")

MAVI_MAX_SPACES = 8

# status codes
OK, ERR_BAD_PARAMS, ERR_OUT_OF_GRID, ERR_NAN, ERR_CUDA, ERR_NCCL, ERR_OUTSIDE_SPACE, ERR_CAPACITY, ERR_UNSUPPORTED = range(9)
STATUS_NAMES = ["MAVI_OK", "MAVI_ERR_BAD_PARAMS", "MAVI_ERR_OUT_OF_GRID", "MAVI_ERR_NAN", "MAVI_ERR_CUDA",
                "MAVI_ERR_NCCL", "MAVI_ERR_OUTSIDE_SPACE", "MAVI_ERR_CAPACITY", "MAVI_ERR_UNSUPPORTED"]
F64, F32 = 0, 1
WALL_RIGID, WALL_PERIODIC, WALL_SLIPPERY, WALL_POTENTIAL = range(4)
GEOM_RECT, GEOM_CIRCLE, GEOM_LINES = range(3)
POT_HARMTRUNC, POT_LJ = range(2)
WALLMODE_OUTSIDE, WALLMODE_INSIDE, WALLMODE_REPULSION = range(3)
DYN_LJ, DYN_HARMTRUNC, DYN_SZABO, DYN_RTP, DYN_RINGS = range(5)
RNG_HOST_NOISE, RNG_PHILOX = range(2)
FLAG_RESORT_EVERY_STEP = 1
FLAG_TIGHT_TILES = 2
FLAG_NO_FORCE_CARRY = 4
FLAG_SMALL_BLOCKS = 8
FLAG_SLAB_SELF = 16
FLAG_LEGACY_STAGING = 32  # A/B: round-1 force kernels (include/mavi.h)
NEIGH_OFF, NEIGH_COUNT, NEIGH_LIST = range(3)
SRC_SOURCE, SRC_SINK = range(2)
NEIGH_MAX = 15


class MaviLine(C.Structure):
    _fields_ = [("p1", C.c_double * 2), ("p2", C.c_double * 2)]


MAVI_MAX_POT_TYPES = 4


class MaviSpace(C.Structure):
    _fields_ = [
        ("wall", C.c_int32), ("geom", C.c_int32),
        ("rect_bl", C.c_double * 2), ("rect_len", C.c_double), ("rect_h", C.c_double),
        ("circ_center", C.c_double * 2), ("circ_radius", C.c_double),
        ("lines", C.POINTER(MaviLine)), ("n_lines", C.c_int32),
        ("pot_kind", C.c_int32), ("pot", C.c_double * 4), ("pot_mode", C.c_int32), ("n_pot_types", C.c_int32),
        ("pot_types", (C.c_double * 4) * MAVI_MAX_POT_TYPES),
    ]


class MaviRingsParams(C.Structure):
    _fields_ = [
        ("num_types", C.c_int32), ("n_max", C.c_int32), ("num_rings", C.c_int64),
        ("p0", C.POINTER(C.c_double)), ("relax_time", C.POINTER(C.c_double)), ("vo", C.POINTER(C.c_double)),
        ("mobility", C.POINTER(C.c_double)), ("rot_diff", C.POINTER(C.c_double)), ("k_area", C.POINTER(C.c_double)),
        ("k_spring", C.POINTER(C.c_double)), ("l_spring", C.POINTER(C.c_double)),
        ("num_particles", C.POINTER(C.c_int32)),
        ("interaction", C.POINTER(C.c_double)),
        ("types", C.POINTER(C.c_int32)),
    ]


class MaviSourceSink(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("num_spawn_pos", C.c_int32), ("spawn_pos", C.POINTER(C.c_double)),
        ("bottom_left", C.c_double * 2), ("spawn_pol", C.c_double), ("pad", C.c_double), ("offset", C.c_double * 2),
        ("size", C.c_int32 * 2), ("sink_geom", C.c_int32), ("_pad", C.c_int32),
        ("sink_rect_bl", C.c_double * 2), ("sink_rect_len", C.c_double), ("sink_rect_h", C.c_double),
        ("sink_circ_center", C.c_double * 2), ("sink_circ_radius", C.c_double),
    ]


class MaviParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("n", C.c_int64),
        ("n_spaces", C.c_int32), ("_pad0", C.c_int32),
        ("spaces", MaviSpace * MAVI_MAX_SPACES),
        ("grid_bl", C.c_double * 2), ("grid_len", C.c_double), ("grid_h", C.c_double),
        ("num_cols", C.c_int32), ("num_rows", C.c_int32),
        ("dynamics", C.c_int32), ("_pad1", C.c_int32),
        ("dyn", C.c_double * 8), ("particle_radius", C.c_double),
        ("rings", C.POINTER(MaviRingsParams)),
        ("dt", C.c_double),
        ("rng_mode", C.c_int32), ("n_gpus", C.c_int32), ("seed", C.c_uint64),
        ("device", C.c_int32), ("flags", C.c_int32), ("stream", C.c_void_p),
        ("rank", C.c_int32), ("world", C.c_int32), ("nccl_unique_id", C.c_void_p), ("n_global", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/mavi.h declares
_H = C.c_void_p
SIGNATURES = {
    "mavi_create": (C.c_int32, [C.POINTER(MaviParams), C.POINTER(_H)]),
    "mavi_destroy": (C.c_int32, [_H]),
    "mavi_abi_version": (C.c_int32, []),
    "mavi_upload_state": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mavi_download_state": (C.c_int32, [_H, C.c_void_p, C.c_void_p]),
    "mavi_download_forces": (C.c_int32, [_H, C.c_void_p]),
    "mavi_local_count": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "mavi_download_local": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mavi_nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "mavi_upload_local": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mavi_step": (C.c_int32, [_H, C.c_int64, C.c_void_p]),
    "mavi_calc_forces": (C.c_int32, [_H]),
    "mavi_bin": (C.c_int32, [_H]),
    "mavi_download_cells": (C.c_int32, [_H, C.c_void_p, C.c_void_p]),
    "mavi_download_cell_lists": (C.c_int32, [_H, C.c_void_p, C.c_void_p]),
    "mavi_cell_neighbors": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mavi_cells_of_points": (C.c_int32, [C.POINTER(MaviParams), C.c_void_p, C.c_int64, C.c_void_p]),
    "mavi_energies": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mavi_rings_download_info": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mavi_rings_set_neighbors": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_double]),
    "mavi_rings_download_neighbors": (C.c_int32, [_H, C.c_void_p, C.c_void_p]),
    "mavi_rings_set_sources": (C.c_int32, [_H, C.POINTER(MaviSourceSink), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "mavi_rings_download_active": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "mavi_rings_set_invasions": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "mavi_rings_download_invasions": (C.c_int32, [_H, C.POINTER(C.c_int64), C.c_void_p, C.c_int64]),
    "mavi_get_time": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "mavi_set_time": (C.c_int32, [_H, C.c_int64, C.c_double]),
    "mavi_sync": (C.c_int32, [_H]),
    "mavi_last_error": (C.c_int32, [_H, C.c_char_p, C.c_int32]),
    "mavi_launch_count": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "mavi_rebuild_count": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "mavi_last_step_ms": (C.c_int32, [_H, C.POINTER(C.c_float)]),
    "mavi_set_profiling": (C.c_int32, [_H, C.c_int32]),
    "mavi_counters": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
}


class MaviError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        name = STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else str(status)
        super().__init__(f"{name}: {msg}" if msg else name)


def build_library(verbose=False):
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libmavi_cuda.so failed")
    return LIB_PATH


_lib = None


def load_library(path=None):
    """dlopen libmavi_cuda.so and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("MAVI_LIB_PATH") or LIB_PATH  # MAVI_LIB_PATH: A/B runs against another build
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with __graft_entry__.build() / make -C mavi.jl_b200/csrc. "
            "There is no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
