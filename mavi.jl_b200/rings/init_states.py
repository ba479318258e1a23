"""Mirror of `Mavi.Rings.InitStates` (reference: src/rings/init_states.jl) — synthetic-input generators."""
from __future__ import annotations

import math

import numpy as np

from ..configs import RectangleCfg
from .configs import get_ring_radius


def create_circle(center, radius, num_p):
    """src/rings/init_states.jl:10-25 -> (num_p, 2)."""
    theta = 2 * math.pi / num_p
    i = np.arange(num_p, dtype=np.float64)
    return np.stack([radius * np.cos(i * theta) + center[0], radius * np.sin(i * theta) + center[1]], axis=1)


def rectangular_grid(*, num_cols, num_rows, num_particles, p_radius, types=None, pad_x=0, pad_y=0, radius_k=1):
    """src/rings/init_states.jl:27-72.  Rings are indexed column-outer / row-inner; the y stride uses pad_x (sic, :52).
    Returns rings_pos (num_rings, n_max, 2) and the RectangleCfg."""
    if types is None:
        num_particles = [num_particles]
        p_radius = [p_radius]
        types = np.ones(num_cols * num_rows, dtype=np.int64)
    ring_r = [get_ring_radius(pr, n) for pr, n in zip(p_radius, num_particles)]
    ring_length = [2 * (rr + pr) for rr, pr in zip(ring_r, p_radius)]
    max_ring_r, max_ring_length, max_num_particles = max(ring_r), max(ring_length), max(num_particles)
    pad_x = pad_x * max_ring_r
    pad_y = pad_y * max_ring_r
    ring_pos = np.zeros((num_cols * num_rows, max_num_particles, 2))
    idx = 0
    for col_id in range(1, num_cols + 1):
        for row_id in range(1, num_rows + 1):
            center_x = pad_x / 2 + (col_id - 1) * (max_ring_length + pad_x) + max_ring_length / 2
            center_y = pad_y / 2 + (row_id - 1) * (max_ring_length + pad_x) + max_ring_length / 2
            t = int(types[idx]) - 1
            num_p = num_particles[t]
            ring_pos[idx, :num_p] = create_circle((center_x, center_y), ring_r[t] * radius_k, num_p)
            idx += 1
    space_l = num_cols * (pad_x + max_ring_length)
    space_h = num_rows * (pad_y + max_ring_length)
    return ring_pos, RectangleCfg(length=space_l, height=space_h)


def random_pol(num_rings, rng=None):
    """src/rings/init_states.jl:74-80 (numpy Generator instead of a Julia RNG)."""
    rng = rng or np.random.default_rng()
    return rng.random(num_rings) * 2 * math.pi
