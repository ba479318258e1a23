"""Shared builders: ONE set of configs -> (device System, CPU oracle) pairs on identical seeded inputs."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
from mavi_jl_b200.params import lower  # noqa: E402

SEED = 24042001


def rel_err(a, b):
    """Norm-wise relative error ||a-b||_inf / ||b||_inf (SURVEY.md 7 'Parity metric under cancellation')."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


def lattice(nx, ny, dyn, offset=0.4, jitter=0.0, seed=SEED):
    """rectangular_grid lattice (src/init_states.jl:34-57) optionally jittered by +-jitter*radius (seeded)."""
    r = pkg.particle_radius(dyn)
    pos, geom = pkg.rectangular_grid(nx, ny, offset, r)
    rng = np.random.default_rng(seed)
    if jitter:
        pos = pos + rng.uniform(-jitter * r, jitter * r, pos.shape)
    return pos, geom, rng


def newton_case(nx=32, ny=32, dyn=None, wall="periodic", chunks=True, dt=0.001, jitter=0.05, vmax=0.2, seed=SEED,
                offset=0.4, active_mask=None, cells=None):
    dyn = dyn or pkg.LenJonesCfg(sigma=1.0, epsilon=1.0)
    pos, geom, rng = lattice(nx, ny, dyn, offset=offset, jitter=jitter, seed=seed)
    vel = pkg.random_vel(nx * ny, vmax, rng=rng)
    wall_t = pkg.PeriodicWalls() if wall == "periodic" else pkg.RigidWalls()
    space = pkg.SpaceCfg(wall_type=wall_t, geometry_cfg=geom)
    if cells is None:
        cells = (int(nx * 0.9), int(ny * 0.9))  # examples/chunks.jl:40-46
    ccfg = pkg.ChunksCfg(num_cols=cells[0], num_rows=cells[1]) if chunks else None
    int_cfg = pkg.IntCfg(dt=dt, chunks_cfg=ccfg)
    act = None if active_mask is None else pkg.ActiveState(active_mask)
    mk = lambda: pkg.SecondLawState(pos=pos.copy(), vel=vel.copy(), active_state=act)  # noqa: E731
    return dict(mk=mk, space=space, dyn=dyn, int_cfg=int_cfg, geom=geom)


def sp_case(kind="szabo", nx=24, ny=24, dt=0.01, jitter=0.9, seed=SEED, rot_diff=0.01, chunks=True, wall="periodic"):
    if kind == "szabo":
        dyn = pkg.SzaboCfg(vo=1.0, mobility=1.0, relax_time=1.0, k_rep=10.0, k_adh=0.75, r_eq=1.0, r_max=1.1,
                           rot_diff=rot_diff)  # examples/szabo.jl:20-29
    else:
        dyn = pkg.RunTumbleCfg(vo=1.0, sigma=1.0, epsilon=1.0, tumble_rate=1.0)  # examples/rtp.jl:20-25
        dt = min(dt, 0.001)
    pos, geom, rng = lattice(nx, ny, dyn, offset=1.0, jitter=jitter, seed=seed)
    ang = rng.random(nx * ny) * 2 * np.pi
    wall_t = pkg.PeriodicWalls() if wall == "periodic" else pkg.RigidWalls()
    space = pkg.SpaceCfg(wall_type=wall_t, geometry_cfg=geom)
    ccfg = pkg.ChunksCfg(num_cols=nx - 1, num_rows=ny - 1) if chunks else None
    int_cfg = pkg.IntCfg(dt=dt, chunks_cfg=ccfg, device=pkg.CUDADevice(rng_mode="host_noise"))
    mk = lambda: pkg.SelfPropelledState(pos=pos.copy(), pol_angle=ang.copy())  # noqa: E731
    return dict(mk=mk, space=space, dyn=dyn, int_cfg=int_cfg, geom=geom, rng=rng)


def make_gpu(case):
    return pkg.System(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"])


def make_oracle(case, threads=1):
    oracle = entry.load_oracle()
    return oracle.OracleSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"],
                               int_cfg=case["int_cfg"], lower=lower, threads=threads)


# ---------------------------------------------------------------- Mavi.Rings fixtures (test/tests_rings/rings_utils.jl)
def rings_case(kind="normal", num_cols=13, num_rows=13, use_chunks=True, seed=31415, rot_diff=None, wall="periodic", pad=0.1):
    """`create_system_normal` (:27-99) / `create_system_types` (:193-296) with a seeded numpy RNG in place of Julia's
    MersenneTwister (the reference's golden neighbour lists depend on Julia's RNG streams and cannot be reproduced)."""
    from mavi_jl_b200.rings import configs as rc
    from mavi_jl_b200.rings import init_states as ri
    from mavi_jl_b200.rings.states import RingsState

    rng = np.random.default_rng(seed)
    if kind == "normal":
        inter = rc.HarmTruncCfg(k_rep=20, k_atr=4, dist_eq=1, dist_max=1 + 0.2)
        dyn = rc.RingsCfg(p0=3.5, relax_time=1, vo=1.0, mobility=1, rot_diff=0.05 if rot_diff is None else rot_diff,
                          k_area=1, k_spring=20, l_spring=1, num_particles=10, interaction_finder=inter)
        rings_pos, geom = ri.rectangular_grid(num_cols=num_cols, num_rows=num_rows, num_particles=10,
                                              p_radius=dyn.particle_radius(), pad_x=pad, pad_y=pad)
        types, num_particles = None, None
        max_size = inter.dist_max * 1.1
    else:
        i1 = rc.HarmTruncCfg(k_rep=13, k_atr=0.1, dist_eq=1, dist_max=1 * 1.1)
        i2 = rc.HarmTruncCfg(k_rep=13, k_atr=0.1, dist_eq=0.8, dist_max=0.8 * 1.1)
        pd = i1.dist_eq / 2 + i2.dist_eq / 2
        ip = rc.HarmTruncCfg(k_rep=13, k_atr=10, dist_eq=pd, dist_max=pd * 1.2)
        finder = rc.InteractionMatrix([[i1, ip], [ip, i2]])
        num_particles = [10, 5]
        dyn = rc.RingsCfg(p0=3.5, relax_time=1, vo=1, mobility=1, rot_diff=0.5 if rot_diff is None else rot_diff, k_area=1,
                          k_spring=20, l_spring=[i.dist_eq * 0.8 for i in rc.list_self_interactions(finder)],
                          num_particles=num_particles, interaction_finder=finder)
        types = rng.integers(1, 3, num_cols * num_rows)
        rings_pos, geom = ri.rectangular_grid(num_cols=num_cols, num_rows=num_rows, num_particles=num_particles,
                                              p_radius=[i1.particle_radius(), i2.particle_radius()], types=types,
                                              pad_x=0.1, pad_y=0.1)
        max_size = max(i.dist_max for i in rc.list_interactions(finder)) * 1.1
    pol = ri.random_pol(num_cols * num_rows, rng=rng)
    wall_t = pkg.PeriodicWalls() if wall == "periodic" else pkg.RigidWalls()
    space = pkg.SpaceCfg(wall_type=wall_t, geometry_cfg=geom)
    chunks = pkg.ChunksCfg(int(geom.length // max_size), int(geom.height // max_size)) if use_chunks else None
    int_cfg = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=chunks, device=pkg.CUDADevice(rng_mode="host_noise"))
    mk = lambda: RingsState(rings_pos=rings_pos.copy(), pol=pol.copy(), types=None if types is None else types.copy(),  # noqa: E731
                            num_particles=num_particles)
    return dict(mk=mk, space=space, dyn=dyn, int_cfg=int_cfg, geom=geom, num_rings=num_cols * num_rows, rng=rng)


def make_gpu_rings(case):
    from mavi_jl_b200.rings.rings import RingsSystem
    return RingsSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"])
