"""Shared builders: ONE set of configs -> (device System, CPU oracle) pairs on identical seeded inputs."""
from __future__ import annotations

import numpy as np

import __graft_entry__ as entry

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
