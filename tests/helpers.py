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


def rings_source_case(use_chunks=True, spawn_pol="random", seed=24042001, num_slots=100):
    """`create_system_source` (test/tests_rings/rings_utils.jl:101-191): an empty periodic box with 100 inactive ring slots, a
    2 x 2 SourceCfg on the left and a SinkCfg rectangle on the right.  Spawn shape: a circle of the ring's target area
    A0 = (n l0 / p0)^2 (the reference calls get_equilibrium_area, an NLsolve host helper; for these parameters A0 lies
    below its limiting value and is returned as is)."""
    import math
    from mavi_jl_b200.rings import configs as rc
    from mavi_jl_b200.rings import init_states as ri
    from mavi_jl_b200.rings.sources import SinkCfg, SourceCfg
    from mavi_jl_b200.rings.states import RingsState

    n = 10
    inter = rc.HarmTruncCfg(k_rep=30, k_atr=3, dist_eq=1, dist_max=1 * 1.5)
    dyn = rc.RingsCfg(p0=3.545, relax_time=100, vo=1, mobility=1, rot_diff=0, k_area=3, k_spring=30,
                      l_spring=inter.dist_eq * 0.8, num_particles=n, interaction_finder=inter)
    ring_d = rc.get_ring_radius(dyn.particle_radius(), n) * 2
    geom = pkg.RectangleCfg(length=18 * ring_d, height=10 * ring_d)
    ring_area = (n * float(dyn.l_spring[0]) / float(dyn.p0[0])) ** 2
    r = (ring_area / math.pi) ** .5
    spawn_pos = ri.create_circle((0, 0), r, n)
    rng = np.random.default_rng(seed)
    pol = ri.random_pol(num_slots, rng=rng)
    space = pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom)
    max_size = inter.dist_max * 1.2
    chunks = pkg.ChunksCfg(int(geom.length // max_size), int(geom.height // max_size)) if use_chunks else None
    bl = np.asarray(geom.bottom_left, dtype=float)
    source_cfg = [
        SourceCfg(bottom_left=bl + np.array([2 * ring_d, geom.height / 2 - ring_d]), spawn_pos=spawn_pos, size=(2, 2),
                  spawn_pol=spawn_pol, pad=dyn.particle_radius()),
        SinkCfg(pkg.RectangleCfg(length=ring_d * 2, height=ring_d * 2,
                                 bottom_left=bl + np.array([geom.length - 4 * ring_d, geom.height / 2 - ring_d]))),
    ]
    int_cfg = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=chunks, device=pkg.CUDADevice(rng_mode="host_noise"))
    rings_pos = np.zeros((num_slots, n, 2))
    mk = lambda: RingsState(rings_pos=rings_pos.copy(), pol=pol.copy(),  # noqa: E731
                            active_state=pkg.ActiveState(np.zeros(num_slots, dtype=bool)))
    draws = np.random.default_rng(seed + 1).random(4096)   # stands for the rand(system.rng) of spawn_pol = :random
    return dict(mk=mk, space=space, dyn=dyn, int_cfg=int_cfg, geom=geom, num_rings=num_slots, source_cfg=source_cfg,
                spawn_draws=draws, ring_d=ring_d)


def make_oracle_sources(case, threads=1):
    oracle = entry.load_oracle()
    return oracle.OracleSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"],
                               lower=lower, threads=threads, source_cfg=case["source_cfg"], spawn_draws=case["spawn_draws"])


def make_gpu_rings_sources(case):
    from mavi_jl_b200.rings.rings import RingsSystem
    return RingsSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"],
                       source_cfg=case["source_cfg"], spawn_draws=case["spawn_draws"])
