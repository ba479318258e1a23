"""Host-side mirror of `Mavi.Configs` (reference: src/configs.jl).

Same names, keyword arguments and meaning as the Julia structs so that a reference user finds the
configuration surface unchanged; these objects are lowered to the flat `MaviParams` POD by `params.py`.
Only what parameterises the device hot path is mirrored; host-only helpers (`check_intersection`,
`is_inside`, JSON StructTypes) are out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple, Union

import numpy as np


# ---------------------------------------------------------------- geometries (src/configs.jl:28-229)
class GeometryCfg:
    pass


@dataclass
class RectangleCfg(GeometryCfg):
    """src/configs.jl:34-51.  `size` = (length, height)."""
    length: float
    height: float
    bottom_left: Tuple[float, float] = (0.0, 0.0)

    def __post_init__(self):
        self.length = float(self.length)
        self.height = float(self.height)
        self.bottom_left = (float(self.bottom_left[0]), float(self.bottom_left[1]))

    @property
    def size(self):
        return (self.length, self.height)

    def __add__(self, other: "RectangleCfg") -> "RectangleCfg":
        """Bounding union, src/configs.jl:68-77."""
        max_x = max(self.bottom_left[0] + self.length, other.bottom_left[0] + other.length)
        max_y = max(self.bottom_left[1] + self.height, other.bottom_left[1] + other.height)
        min_x = min(self.bottom_left[0], other.bottom_left[0])
        min_y = min(self.bottom_left[1], other.bottom_left[1])
        return RectangleCfg(length=max_x - min_x, height=max_y - min_y, bottom_left=(min_x, min_y))


@dataclass
class CircleCfg(GeometryCfg):
    """src/configs.jl:145-152."""
    radius: float
    center: Tuple[float, float]

    def __post_init__(self):
        self.radius = float(self.radius)
        self.center = (float(self.center[0]), float(self.center[1]))


@dataclass
class Line2D:
    """src/configs.jl:95-117 (normal/tangent/length are derived on the device exactly like the ctor)."""
    p1: Tuple[float, float]
    p2: Tuple[float, float]


@dataclass
class LinesCfg(GeometryCfg):
    """src/configs.jl:137-143."""
    lines: Sequence
    bbox: Optional[RectangleCfg] = None

    def __post_init__(self):
        self.lines = [l if isinstance(l, Line2D) else Line2D(tuple(map(float, l[0])), tuple(map(float, l[1])))
                      for l in self.lines]


@dataclass
class ManyGeometries(GeometryCfg):
    list: tuple


def get_bounding_box(g: GeometryCfg) -> RectangleCfg:
    """src/configs.jl:188-229."""
    if isinstance(g, RectangleCfg):
        return g
    if isinstance(g, CircleCfg):
        r = g.radius
        return RectangleCfg(length=2 * r, height=2 * r, bottom_left=(g.center[0] - r, g.center[1] - r))
    if isinstance(g, LinesCfg):
        if g.bbox is not None:
            return g.bbox
        xs = [c for l in g.lines for c in (l.p1[0], l.p2[0])]
        ys = [c for l in g.lines for c in (l.p1[1], l.p2[1])]
        return RectangleCfg(length=max(xs) - min(xs), height=max(ys) - min(ys), bottom_left=(min(xs), min(ys)))
    if isinstance(g, ManyGeometries):
        bbox = get_bounding_box(g.list[0])
        for sub in g.list[1:]:
            bbox = bbox + get_bounding_box(sub)
        return bbox
    raise TypeError(g)


# ---------------------------------------------------------------- walls (src/configs.jl:235-279)
class WallType:
    pass


class RigidWalls(WallType):
    pass


class PeriodicWalls(WallType):
    pass


class SlipperyWalls(WallType):
    pass


@dataclass
class ManyWalls(WallType):
    list: tuple


class ForceWalls(WallType):
    pass


@dataclass
class PotentialVector:
    """src/configs.jl:454-463: one wall potential per particle type (get_particle_type: the ring type of a Mavi.Rings state)."""
    vector: list


@dataclass
class PotentialWalls(ForceWalls):
    """src/configs.jl:263-279.  mode in {'outside','inside','repulsion'} (process_dist :259-261)."""
    potential: object
    mode: str = "repulsion"


# ---------------------------------------------------------------- space (src/configs.jl:285-312)
class SpaceCfg:
    def __init__(self, spaces_cfg_list=None, *, wall_type=None, geometry_cfg=None):
        if spaces_cfg_list is not None:  # SpaceCfg(spaces_cfg_list), src/configs.jl:289-302
            walls = tuple(w for w, _ in spaces_cfg_list)
            geoms = tuple(g for _, g in spaces_cfg_list)
            wall_type, geometry_cfg = ManyWalls(walls), ManyGeometries(geoms)
        self.wall_type = wall_type
        self.geometry_cfg = geometry_cfg

    def pairs(self):
        if isinstance(self.wall_type, ManyWalls):
            return list(zip(self.wall_type.list, self.geometry_cfg.list))
        return [(self.wall_type, self.geometry_cfg)]


def get_main_wall(x):
    w = x.wall_type if isinstance(x, SpaceCfg) else x
    return w.list[0] if isinstance(w, ManyWalls) else w


def get_main_geometry(x):
    g = x.geometry_cfg if isinstance(x, SpaceCfg) else x
    return g.list[0] if isinstance(g, ManyGeometries) else g


# ---------------------------------------------------------------- dynamics (src/configs.jl:318-421)
class DynamicCfg:
    pass


class PotentialCfg(DynamicCfg):
    pass


@dataclass
class HarmTruncCfg(PotentialCfg):
    """src/configs.jl:341-368."""
    k_rep: float
    k_atr: float
    dist_eq: float
    dist_max: float


@dataclass
class LenJonesCfg(PotentialCfg):
    """src/configs.jl:380-397 (no cutoff: the cell stencil defines the pair set)."""
    sigma: float
    epsilon: float


@dataclass
class SzaboCfg(DynamicCfg):
    """src/configs.jl:399-408."""
    vo: float
    mobility: float
    relax_time: float
    k_rep: float
    k_adh: float
    r_eq: float
    r_max: float
    rot_diff: float


@dataclass
class RunTumbleCfg(DynamicCfg):
    """src/configs.jl:410-415."""
    vo: float
    sigma: float
    epsilon: float
    tumble_rate: float


def particle_radius(dynamic_cfg):
    """src/configs.jl:417-421 (Rings: src/rings/configs.jl:55-57,215)."""
    if isinstance(dynamic_cfg, (LenJonesCfg, RunTumbleCfg)):
        return float(dynamic_cfg.sigma) * 2 ** (1 / 6) / 2
    if isinstance(dynamic_cfg, SzaboCfg):
        return float(dynamic_cfg.r_eq) / 2
    if isinstance(dynamic_cfg, HarmTruncCfg):
        return float(dynamic_cfg.dist_eq) / 2.0
    if hasattr(dynamic_cfg, "particle_radius"):
        return dynamic_cfg.particle_radius()
    raise TypeError(f"particle_radius: unsupported {type(dynamic_cfg).__name__}")


# ---------------------------------------------------------------- integration cfg (src/configs.jl:469-490)
class DeviceMode:
    pass


class Sequencial(DeviceMode):
    pass


class Threaded(DeviceMode):
    pass


@dataclass
class CUDADevice(DeviceMode):
    """The new `DeviceMode` subtype that routes the hot path to libmavi_cuda.so.

    rng_mode 'host_noise': the caller supplies per-step noise (exact parity); 'philox': device RNG.
    """
    device: int = 0
    rng_mode: str = "philox"
    seed: int = 24042001
    float32: bool = False
    flags: int = 0
    stream: Optional[int] = None
    # ONE process, n_gpus GPUs inside the handle (devices device .. device+n_gpus-1): the library partitions the state
    # into x-slabs on upload and un-permutes on download; System(...) is used exactly as with one GPU
    n_gpus: int = 1
    # x-slab decomposition, one process per GPU
    rank: int = 0
    world: int = 1
    nccl_unique_id: Optional[bytes] = None
    n_global: int = 0


@dataclass
class ChunksCfg:
    num_cols: int
    num_rows: int


@dataclass
class IntCfg:
    dt: float
    chunks_cfg: Optional[ChunksCfg] = None
    device: DeviceMode = field(default_factory=CUDADevice)
    extra: object = None


def has_chunks(int_cfg: IntCfg) -> bool:
    return int_cfg.chunks_cfg is not None
