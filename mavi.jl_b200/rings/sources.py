"""Mirror of `Mavi.Rings.Sources` configs (reference: src/rings/sources.jl:29-45, :227-230) and their lowering to the
`MaviSourceSink` POD of include/mavi.h.  The processing itself (update_area_empty!, add_ring!, remove_ring!,
calc_active_ids!) runs at the head of every device step inside libmavi_cuda.so (csrc/rings.cu)."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Sequence, Tuple, Union

import numpy as np

from .. import capi
from ..configs import CircleCfg, RectangleCfg


@dataclass
class SourceCfg:
    """src/rings/sources.jl:29-45.  spawn_pol: a number or "random" (:random -> rand(rng) * 2 pi at every spawn)."""
    bottom_left: Sequence[float]
    spawn_pos: np.ndarray                 # (num_particles, 2)
    spawn_pol: Union[float, str]
    pad: float = 0.0
    offset: Tuple[float, float] = (0.0, 0.0)
    size: Tuple[int, int] = (1, 1)

    def __post_init__(self):
        self.spawn_pos = np.ascontiguousarray(self.spawn_pos, dtype=np.float64).reshape(-1, 2)
        if isinstance(self.spawn_pol, str):
            if self.spawn_pol.lstrip(":") != "random":
                raise ValueError("spawn_pol must be a number or :random")
            self.spawn_pol = "random"


@dataclass
class SinkCfg:
    """src/rings/sources.jl:227-229."""
    geometry_cfg: object


def lower_sources(source_cfg, keep):
    """[SourceCfg | SinkCfg, ...] -> (MaviSourceSink array, n); buffers the POD points to are appended to `keep`."""
    items = list(source_cfg) if isinstance(source_cfg, (list, tuple)) else [source_cfg]
    arr = (capi.MaviSourceSink * max(len(items), 1))()
    for k, c in enumerate(items):
        e = arr[k]
        if isinstance(c, SourceCfg):
            e.kind = capi.SRC_SOURCE
            e.num_spawn_pos = len(c.spawn_pos)
            keep.append(c.spawn_pos)
            e.spawn_pos = c.spawn_pos.ctypes.data_as(C.POINTER(C.c_double))
            e.bottom_left[:] = [float(c.bottom_left[0]), float(c.bottom_left[1])]
            e.spawn_pol = math.nan if c.spawn_pol == "random" else float(c.spawn_pol)
            e.pad = float(c.pad)
            e.offset[:] = [float(c.offset[0]), float(c.offset[1])]
            e.size[:] = [int(c.size[0]), int(c.size[1])]
        elif isinstance(c, SinkCfg):
            e.kind = capi.SRC_SINK
            g = c.geometry_cfg
            if isinstance(g, RectangleCfg):
                e.sink_geom = capi.GEOM_RECT
                e.sink_rect_bl[:] = [float(g.bottom_left[0]), float(g.bottom_left[1])]
                e.sink_rect_len, e.sink_rect_h = float(g.length), float(g.height)
            elif isinstance(g, CircleCfg):
                e.sink_geom = capi.GEOM_CIRCLE
                e.sink_circ_center[:] = [float(g.center[0]), float(g.center[1])]
                e.sink_circ_radius = float(g.radius)
            else:
                raise TypeError(f"SinkCfg geometry {type(g).__name__} has no is_inside method in the reference")
        else:
            raise TypeError(f"Unknown source configuration type: {type(c).__name__}")   # src/rings/rings.jl:180
    keep.append(arr)
    return arr, len(items)
