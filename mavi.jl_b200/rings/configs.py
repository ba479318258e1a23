"""Host-side mirror of `Mavi.Rings.Configs` (reference: src/rings/configs.jl).

Only the parameter surface of the hot path is mirrored (RingsCfg, the Rings HarmTruncCfg,
InteractionMatrix, RingsIntCfg, get_ring_radius of src/rings/utils.jl).  The NLsolve-based equilibrium
helpers (:221-328) are one-off host setup and out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Union

import numpy as np

from ..configs import ChunksCfg, CUDADevice, DynamicCfg, IntCfg


@dataclass
class HarmTruncCfg:
    """Pairwise ring interaction, src/rings/configs.jl:26-57."""
    k_rep: float
    k_atr: float
    dist_eq: float
    dist_max: float

    def particle_radius(self):
        return self.dist_eq / 2.0


class InteractionMatrix:
    """src/rings/configs.jl:61-89: matrix[type1][type2] of HarmTruncCfg."""

    def __init__(self, matrix):
        self.matrix = [list(row) for row in matrix]

    def get(self, t1, t2):
        return self.matrix[t1][t2]


def list_interactions(finder):
    """src/rings/configs.jl:65-70 (upper triangle, row by row)."""
    if isinstance(finder, HarmTruncCfg):
        return [finder]
    m = finder.matrix
    return [m[i][j] for i in range(len(m)) for j in range(i, len(m))]


def list_self_interactions(finder):
    """src/rings/configs.jl:72-77."""
    if isinstance(finder, HarmTruncCfg):
        return [finder]
    return [finder.matrix[i][i] for i in range(len(finder.matrix))]


def get_ring_radius(p_radius, num_particles):
    """src/rings/utils.jl:5-7."""
    return (p_radius * 2) / (2 * (1 - math.cos(2 * math.pi / num_particles))) ** .5


class RingsCfg(DynamicCfg):
    """src/rings/configs.jl:95-164.  Every physical parameter is a scalar or one value per ring type;
    scalars are broadcast to `num_types` when any parameter is a vector (:112-122)."""
    _names = ("p0", "relax_time", "vo", "mobility", "rot_diff", "k_area", "k_spring", "l_spring")

    def __init__(self, *, p0, relax_time, vo, mobility, rot_diff, k_area, k_spring, l_spring,
                 interaction_finder, num_particles=-1):
        args = dict(p0=p0, relax_time=relax_time, vo=vo, mobility=mobility, rot_diff=rot_diff,
                    k_area=k_area, k_spring=k_spring, l_spring=l_spring)
        lens = [len(v) if isinstance(v, (list, tuple, np.ndarray)) else 1 for v in args.values()]
        self.num_types = max(lens)
        self.has_types = any(isinstance(v, (list, tuple, np.ndarray)) for v in args.values())
        for k, v in args.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                arr = np.asarray(v, dtype=np.float64)
                if len(arr) != self.num_types:
                    raise ValueError(f"{k} has {len(arr)} entries, expected {self.num_types}")
            else:
                arr = np.full(self.num_types, float(v))
            setattr(self, k, arr)
        if num_particles != -1:
            if not self.has_types and not isinstance(num_particles, (int, np.integer)):
                raise ValueError("If U is Number, num_particles must be Int")
            if self.has_types and isinstance(num_particles, (int, np.integer)):
                raise ValueError("If U is AbstractVector, num_particles must be Vector")
            if self.has_types and len(num_particles) != self.num_types:
                raise ValueError(f"length(num_particles)={len(num_particles)}, but there exists {self.num_types} types")
        self.num_particles = num_particles
        self.interaction_finder = interaction_finder

    def interaction(self, t1, t2):
        f = self.interaction_finder
        return f if isinstance(f, HarmTruncCfg) else f.get(t1, t2)

    def particle_radius(self):
        """MaviCfg.particle_radius(::RingsCfg), src/rings/configs.jl:210-215: scalar or per-type vector."""
        r = [self.interaction(t, t).particle_radius() for t in range(self.num_types)]
        return r[0] if len(r) == 1 else r


@dataclass
class NeighborsCfg:
    """src/rings/neighbors.jl:11-15.  `type` is "rings" (:rings, particles of other rings only) or "all" (:all)."""
    only_count: bool = False
    type: str = "all"
    tol: float = 1.1

    def __post_init__(self):
        self.type = str(self.type).lstrip(":")
        if self.type not in ("rings", "all"):
            raise ValueError("NeighborsCfg.type must be :rings or :all")


@dataclass
class InvasionsCfg:
    """src/rings/configs.jl:334-336."""
    steps_to_update: int


@dataclass
class IntCfgExtra:
    """src/rings/configs.jl:338-341."""
    r_chunks_cfg: Optional[ChunksCfg] = None
    invasions_cfg: Optional[InvasionsCfg] = None


def RingsIntCfg(*, dt, p_chunks_cfg=None, r_chunks_cfg=None, invasions_cfg=None, device=None):
    """src/rings/configs.jl:343-351: an IntCfg whose `extra` carries the ring-level chunks and the invasions config."""
    return IntCfg(dt=dt, chunks_cfg=p_chunks_cfg, device=device or CUDADevice(), extra=IntCfgExtra(r_chunks_cfg, invasions_cfg))
