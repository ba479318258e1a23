"""Host-side mirror of `Mavi.States` (reference: src/states.jl).

Memory contract: `Vector{SVector{2,T}}` == contiguous T[2N]; here `pos`/`vel` are
C-contiguous numpy arrays of shape (N, 2).  Matrix inputs of shape (2, N) are reinterpreted column-wise like
the reference's Matrix constructors (src/states.jl:84-101); integer inputs promote to Float64 (:89-91).
"""
from __future__ import annotations

import numpy as np


class ActiveState:
    """src/states.jl:14-23: a Bool (all particles) or a per-particle mask."""

    def __init__(self, mask=True):
        self.mask = mask

    def get_active_mask(self, num):
        if isinstance(self.mask, (bool, np.bool_)):
            return np.full(num, bool(self.mask), dtype=np.uint8)
        return np.ascontiguousarray(np.asarray(self.mask).ravel(), dtype=np.uint8)


def _as_points(a, dtype=None):
    a = np.asarray(a)
    if a.dtype.kind in "iu":
        a = a.astype(np.float64)
    if a.ndim == 2 and a.shape[0] == 2 and a.shape[1] != 2:
        a = a.T  # Matrix (2, N): columns are particles
    if dtype is not None:
        a = a.astype(dtype)
    return np.ascontiguousarray(a)


class State:
    pass


class SecondLawState(State):
    """Positions and velocities, src/states.jl:75-101."""

    def __init__(self, pos, vel, active_state=None):
        pos, vel = _as_points(pos), _as_points(vel)
        assert pos.shape == vel.shape, "pos and vel must have the same number of particles"
        T = np.promote_types(pos.dtype, vel.dtype)
        self.pos = np.ascontiguousarray(pos, dtype=T)
        self.vel = np.ascontiguousarray(vel, dtype=T)
        self.active_state = active_state

    @property
    def second(self):
        return self.vel

    def active_mask(self):
        return None if self.active_state is None else self.active_state.get_active_mask(len(self.pos))


class SelfPropelledState(State):
    """Positions and polarisation angles, src/states.jl:104-125."""

    def __init__(self, pos, pol_angle, active_state=None):
        pos = _as_points(pos)
        pol_angle = np.asarray(pol_angle)
        T = np.promote_types(pos.dtype, pol_angle.dtype)
        self.pos = np.ascontiguousarray(pos, dtype=T)
        self.pol_angle = np.ascontiguousarray(pol_angle, dtype=T)
        self.active_state = active_state

    @property
    def second(self):
        return self.pol_angle

    def active_mask(self):
        return None if self.active_state is None else self.active_state.get_active_mask(len(self.pos))


def get_num_total_particles(state):
    """src/states.jl:130 (count of active ids)."""
    m = state.active_mask()
    return len(state.pos) if m is None else int(m.sum())


def get_particles_ids(state):
    """src/states.jl:129 (0-based here)."""
    m = state.active_mask()
    return np.arange(len(state.pos)) if m is None else np.flatnonzero(m)
