"""Host-side mirror of `Mavi.Rings.States.RingsState` (reference: src/rings/states.jl:74-124).

`rings_pos` has shape (num_rings, n_max, 2): ring-major, particle within ring, coordinate — the memory order
of the reference's `Matrix{SVector{2,T}}(n_max, num_rings)` (column-major).  `pos` is the flat (num_rings*n_max, 2)
alias (:101); scalar idx = ring*n_max + p (:137-139, 0-based here).
"""
from __future__ import annotations

import numpy as np


class RingsState:
    def __init__(self, *, rings_pos, pol, num_particles=None, types=None, active_state=None):
        rp = np.asarray(rings_pos)
        T = np.float32 if rp.dtype == np.float32 else np.float64  # element type of the state (Float32 mode keeps it)
        rp = rp.astype(T, copy=False)
        if rp.ndim == 3 and rp.shape[0] == 2 and rp.shape[2] != 2:
            rp = np.transpose(rp, (2, 1, 0))  # Julia Array{T,3}(2, n_max, num_rings)
        assert rp.ndim == 3 and rp.shape[2] == 2
        self.rings_pos = np.ascontiguousarray(rp)
        self.num_rings, self.n_max = rp.shape[0], rp.shape[1]
        self.pos = self.rings_pos.reshape(-1, 2)
        self.pol = np.ascontiguousarray(pol, dtype=T)
        if isinstance(num_particles, (list, tuple, np.ndarray)) and types is None:
            raise ValueError("argument 'types' is empty!")
        self.num_particles = self.n_max if num_particles is None else num_particles
        self.types = None if types is None else np.ascontiguousarray(types, dtype=np.int32)  # 1-based like Julia
        # VarRingsIds (src/rings/states.jl:24-43, :104-113): active_state = ActiveState(mask over rings) makes the number of
        # rings variable (sources / sinks); None -> FixRingsIds (every ring active)
        self.ring_mask = None if active_state is None else active_state.get_active_mask(self.num_rings)
        self.uids = None if active_state is None else np.arange(1, self.num_rings + 1, dtype=np.int64)

    @property
    def second(self):
        return self.pol

    def active_mask(self):
        return None

    def ring_num_particles(self, ring):
        """src/rings/states.jl:195-198."""
        if self.types is None:
            return int(self.num_particles)
        return int(self.num_particles[self.types[ring] - 1])
