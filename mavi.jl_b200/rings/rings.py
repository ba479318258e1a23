"""Mirror of `Mavi.Rings.RingsSystem` (reference: src/rings/rings.jl:231-291)."""
from __future__ import annotations

from ..systems import RingsSys, System
from .configs import RingsCfg


def RingsSystem(*, state, space_cfg, dynamic_cfg: RingsCfg, int_cfg, p_neighbors_cfg=None, r_neighbors_cfg=None,
                source_cfg=None, user_data=None, time_info=None, rng=None, spawn_draws=None):
    if dynamic_cfg.has_types != (state.types is not None):
        if dynamic_cfg.has_types:
            raise ValueError("DynamicCfg has multiple types, but state.types is nothing!")
        raise ValueError("DynamicCfg has only one type, but state.types is not nothing!")
    if r_neighbors_cfg is not None:
        raise NotImplementedError("ring-level neighbours are not updated by the reference's own step! "
                                  "(src/rings/integration.jl:363,536): not on the hot path")
    if source_cfg is not None and state.ring_mask is None:
        raise ValueError("sources / sinks need a RingsState built with active_state (VarRingsIds)")
    if isinstance(dynamic_cfg.num_particles, int) and dynamic_cfg.num_particles == -1:
        dynamic_cfg.num_particles = state.num_particles
    if list(_aslist(dynamic_cfg.num_particles)) != list(_aslist(state.num_particles)):
        raise ValueError("`dynamic_cfg.num_particles` is not equal to `state.num_particles`!")
    return System(state=state, space_cfg=space_cfg, dynamic_cfg=dynamic_cfg, int_cfg=int_cfg,
                  info=user_data, time_info=time_info, sys_type=RingsSys(), rng=rng, p_neighbors_cfg=p_neighbors_cfg,
                  source_cfg=source_cfg, spawn_draws=spawn_draws)


def _aslist(x):
    try:
        return list(x)
    except TypeError:
        return [x]
