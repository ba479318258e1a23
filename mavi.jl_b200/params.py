"""Lower the Mavi config struct zoo to the flat `MaviParams` POD of include/mavi.h.

This is what the Julia glue does before `ccall(:mavi_create, ...)` (SURVEY.md 5 "Config / flag system").
The returned object keeps every buffer the POD points to alive.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .configs import (CircleCfg, CUDADevice, HarmTruncCfg, LenJonesCfg, LinesCfg, ManyGeometries, PeriodicWalls,
                      PotentialVector, PotentialWalls, RectangleCfg, RigidWalls, RunTumbleCfg, SlipperyWalls, SpaceCfg, SzaboCfg,
                      get_bounding_box, particle_radius)

_MODES = {"outside": capi.WALLMODE_OUTSIDE, "inside": capi.WALLMODE_INSIDE, "repulsion": capi.WALLMODE_REPULSION}


class LoweredParams:
    def __init__(self):
        self.params = capi.MaviParams()
        self.keep = []

    def _darr(self, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        self.keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def _iarr(self, values):
        a = np.ascontiguousarray(values, dtype=np.int32)
        self.keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int32))


def _potential(pot):
    if isinstance(pot, LenJonesCfg):
        return capi.POT_LJ, [pot.sigma, pot.epsilon, 0.0, 0.0]
    if hasattr(pot, "k_rep") and hasattr(pot, "dist_max"):
        return capi.POT_HARMTRUNC, [pot.k_rep, pot.k_atr, pot.dist_eq, pot.dist_max]
    raise TypeError(f"unsupported wall potential {type(pot).__name__}")


def lower(state, space_cfg: SpaceCfg, dynamic_cfg, int_cfg) -> LoweredParams:
    from .rings.configs import RingsCfg

    lp = LoweredParams()
    p = lp.params
    dev = int_cfg.device if isinstance(int_cfg.device, CUDADevice) else CUDADevice()
    p.struct_size = C.sizeof(capi.MaviParams)
    p.dtype = capi.F32 if (dev.float32 or state.pos.dtype == np.float32) else capi.F64
    p.n = len(state.pos)

    pairs = space_cfg.pairs()
    if len(pairs) > capi.MAVI_MAX_SPACES:
        raise ValueError(f"at most {capi.MAVI_MAX_SPACES} (wall, geometry) pairs")
    p.n_spaces = len(pairs)
    for k, (w, g) in enumerate(pairs):
        sp = p.spaces[k]
        if isinstance(w, RigidWalls):
            sp.wall = capi.WALL_RIGID
        elif isinstance(w, PeriodicWalls):
            sp.wall = capi.WALL_PERIODIC
        elif isinstance(w, SlipperyWalls):
            sp.wall = capi.WALL_SLIPPERY
        elif isinstance(w, PotentialWalls):
            sp.wall = capi.WALL_POTENTIAL
            if isinstance(w.potential, PotentialVector):   # one entry per particle (ring) type, all of one kind
                # get_particle_type(state, pid) has a method for RingsState only (src/rings/states.jl:148): with any other
                # state the reference's calc_walls_forces! ends in a MethodError, and types index the vector
                if not isinstance(dynamic_cfg, RingsCfg) or getattr(state, "types", None) is None:
                    raise TypeError("PotentialVector needs a RingsState with types (get_particle_type)")
                if len(w.potential.vector) != dynamic_cfg.num_types:
                    raise ValueError(f"PotentialVector has {len(w.potential.vector)} entries for {dynamic_cfg.num_types} ring types")
                kinds = [_potential(q) for q in w.potential.vector]
                if not kinds or len(kinds) > capi.MAVI_MAX_POT_TYPES or len({k for k, _ in kinds}) != 1:
                    raise ValueError(f"PotentialVector: 1..{capi.MAVI_MAX_POT_TYPES} potentials of one kind")
                sp.pot_kind = kinds[0][0]
                sp.pot[:] = kinds[0][1]
                sp.n_pot_types = len(kinds)
                for t, (_, pot) in enumerate(kinds):
                    sp.pot_types[t][:] = pot
            else:
                sp.pot_kind, pot = _potential(w.potential)
                sp.pot[:] = pot
            sp.pot_mode = _MODES[w.mode]
        else:
            raise TypeError(f"unsupported wall type {type(w).__name__}")
        if isinstance(g, RectangleCfg):
            sp.geom = capi.GEOM_RECT
            sp.rect_bl[:] = g.bottom_left
            sp.rect_len, sp.rect_h = g.length, g.height
        elif isinstance(g, CircleCfg):
            sp.geom = capi.GEOM_CIRCLE
            sp.circ_center[:] = g.center
            sp.circ_radius = g.radius
        elif isinstance(g, LinesCfg):
            sp.geom = capi.GEOM_LINES
            arr = (capi.MaviLine * len(g.lines))()
            for i, l in enumerate(g.lines):
                arr[i].p1[:] = l.p1
                arr[i].p2[:] = l.p2
            lp.keep.append(arr)
            sp.lines = C.cast(arr, C.POINTER(capi.MaviLine))
            sp.n_lines = len(g.lines)
        else:
            raise TypeError(f"unsupported geometry {type(g).__name__}")

    # Chunks over the bounding box of the whole geometry (get_chunks, src/systems.jl:14-28)
    bbox = get_bounding_box(space_cfg.geometry_cfg)
    p.grid_bl[:] = bbox.bottom_left
    p.grid_len, p.grid_h = bbox.length, bbox.height
    if int_cfg.chunks_cfg is not None:
        p.num_cols, p.num_rows = int(int_cfg.chunks_cfg.num_cols), int(int_cfg.chunks_cfg.num_rows)

    dyn = [0.0] * 8
    if isinstance(dynamic_cfg, LenJonesCfg):
        p.dynamics = capi.DYN_LJ
        dyn[:2] = [dynamic_cfg.sigma, dynamic_cfg.epsilon]
    elif isinstance(dynamic_cfg, HarmTruncCfg):
        p.dynamics = capi.DYN_HARMTRUNC
        dyn[:4] = [dynamic_cfg.k_rep, dynamic_cfg.k_atr, dynamic_cfg.dist_eq, dynamic_cfg.dist_max]
    elif isinstance(dynamic_cfg, SzaboCfg):
        p.dynamics = capi.DYN_SZABO
        c = dynamic_cfg
        dyn[:8] = [c.vo, c.mobility, c.relax_time, c.k_rep, c.k_adh, c.r_eq, c.r_max, c.rot_diff]
    elif isinstance(dynamic_cfg, RunTumbleCfg):
        p.dynamics = capi.DYN_RTP
        c = dynamic_cfg
        dyn[:4] = [c.vo, c.sigma, c.epsilon, c.tumble_rate]
    elif isinstance(dynamic_cfg, RingsCfg):
        p.dynamics = capi.DYN_RINGS
        c = dynamic_cfg
        rp = capi.MaviRingsParams()
        rp.num_types = c.num_types
        rp.n_max = state.n_max
        rp.num_rings = state.num_rings
        for name in RingsCfg._names:
            setattr(rp, name, lp._darr(getattr(c, name)))
        nps = c.num_particles if c.has_types else [c.num_particles]
        rp.num_particles = lp._iarr(nps)
        inter = [[getattr(c.interaction(a, b), f) for f in ("k_rep", "k_atr", "dist_eq", "dist_max")]
                 for a in range(c.num_types) for b in range(c.num_types)]
        rp.interaction = lp._darr(np.asarray(inter).ravel())
        if state.types is not None:
            rp.types = lp._iarr(state.types)
        lp.keep.append(rp)
        p.rings = C.pointer(rp)
    else:
        raise TypeError(f"DynamicCfg {type(dynamic_cfg).__name__} has no device kernel; user-defined Julia forces "
                        "cannot cross the C ABI (SURVEY.md 8b)")
    p.dyn[:] = dyn
    pr = particle_radius(dynamic_cfg)
    p.particle_radius = float(np.min(pr))  # minimum(particle_radius(dynamic_cfg)), src/systems.jl:26
    p.dt = float(int_cfg.dt)

    p.rng_mode = capi.RNG_HOST_NOISE if dev.rng_mode == "host_noise" else capi.RNG_PHILOX
    p.n_gpus = int(dev.n_gpus)
    p.seed = dev.seed
    p.device = dev.device
    p.flags = dev.flags
    p.stream = dev.stream
    p.rank, p.world = dev.rank, dev.world
    if dev.nccl_unique_id is not None:
        buf = C.create_string_buffer(dev.nccl_unique_id, len(dev.nccl_unique_id))
        lp.keep.append(buf)
        p.nccl_unique_id = C.cast(buf, C.c_void_p)
    p.n_global = dev.n_global
    return lp
