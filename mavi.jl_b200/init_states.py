"""Mirror of `Mavi.InitStates` (reference: src/init_states.jl) — synthetic-input generators."""
from __future__ import annotations

import numpy as np

from .configs import RectangleCfg


def rectangular_grid(num_p_x, num_p_y, offset, radius, NUM_T=np.float64):
    """src/init_states.jl:34-57.  Row-major fill (y outer, x inner); coordinates are built by repeated
    addition exactly like the reference (`current_x = x[end]`), hence the cumulative sums."""
    step = radius * (2 + offset)
    xs = np.cumsum(np.concatenate([[-radius + step], np.full(num_p_x - 1, step)]))
    ys = np.cumsum(np.concatenate([[radius * (offset + 1)], np.full(num_p_y - 1, step)]))
    pos = np.empty((num_p_y, num_p_x, 2), dtype=NUM_T)
    pos[..., 0] = xs[None, :]
    pos[..., 1] = ys[:, None]
    geometry_cfg = RectangleCfg(
        length=num_p_x * 2 * radius + radius * offset * (num_p_x + 1),
        height=num_p_y * 2 * radius + radius * offset * (num_p_y + 1),
    )
    return pos.reshape(-1, 2), geometry_cfg


def random_vel(num_p, max_value=1, rng=None, NUM_T=np.float64):
    """src/init_states.jl:63-70: each component uniform in [-max, max)."""
    rng = rng or np.random.default_rng()
    return ((rng.random((num_p, 2)) * 2 - 1) * max_value).astype(NUM_T)
