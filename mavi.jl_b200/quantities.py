"""Mirror of `Mavi.Quantities` (reference: src/quantities.jl) as device block reductions."""
from __future__ import annotations


def kinetic_energy(system):
    """`kinetic_energy(state)`, src/quantities.jl:12-18 (takes the system: the state lives on the device)."""
    return system.energies(0, want_pe=False)[0]


def potential_energy(system, dynamic_cfg=None, stencil_only=False):
    """`potential_energy(system, ::LenJonesCfg)`, src/quantities.jl:46-66: exact all-pairs O(N^2).
    stencil_only=True sums only the cell-stencil pair set (labelled deviation)."""
    return system.energies(1 if stencil_only else 0, want_ke=False)[1]
