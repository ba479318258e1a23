"""GPU parity tests of the core hot path (binning, pair forces, Newton / Szabo / RTP steps, walls, quantities):
libmavi_cuda.so through the C ABI vs the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): cell assignment and neighbour sets bit-exact; forces and positions within
1e-12 relative in Float64, measured norm-wise (||a-b||_inf / ||b||_inf; positions relative to the box scale) because
lattice states have near-zero net forces by symmetry (SURVEY.md 7 'Parity metric under cancellation').
"""
import os
from collections import Counter

import numpy as np
import pytest

import helpers as H

pkg = H.pkg
pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pair(case, threads=1):
    return H.make_gpu(case), H.make_oracle(case, threads)


# ---------------------------------------------------------------- binning: bit-exact
@pytest.mark.parametrize("wall", ["periodic", "rigid"])
def test_cell_assignment_bit_exact(cuda_lib, wall):
    case = H.newton_case(nx=64, ny=48, wall=wall, jitter=0.3)
    g, o = _pair(case)
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(cg, co)
    assert np.array_equal(ng, no)
    sg, ig = g.download_cell_lists()
    so, io = o.download_cell_lists()
    assert np.array_equal(sg, so)
    assert np.array_equal(ig, io)  # same cell order AND ascending ids inside every cell (src/chunks.jl:153-155)


def test_cell_assignment_adversarial_points(cuda_lib, oracle):
    """Points exactly on (rounded) cell boundaries, on the box edges, and one ulp either side: the Julia `div`
    semantics (fmod based) must be reproduced bit-exactly by the device's FMA-corrected quotient."""
    dyn = pkg.HarmTruncCfg(k_rep=1.0, k_atr=1.0, dist_eq=0.05, dist_max=0.06)
    L, Hh, nc, nr = 37.3, 21.7, 53, 31
    geom = pkg.RectangleCfg(length=L, height=Hh)
    cl, ch = L / nc, Hh / nr
    pts = []
    for k in range(nc + 1):
        for x in (k * cl, np.nextafter(k * cl, -1), np.nextafter(k * cl, 1e9)):
            pts.append((min(max(x, 0.0), L), 0.37 * Hh))
    for k in range(nr + 1):
        for y in (k * ch, np.nextafter(k * ch, -1), np.nextafter(k * ch, 1e9), Hh - k * ch):
            pts.append((0.61 * L, min(max(y, 0.0), Hh)))
    rng = np.random.default_rng(5)
    pts += list(zip(rng.uniform(0, L, 4000), rng.uniform(0, Hh, 4000)))
    pts = np.array(pts)
    mk = lambda: pkg.SecondLawState(pos=pts.copy(), vel=np.zeros_like(pts))  # noqa: E731
    case = dict(mk=mk, space=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom), dyn=dyn,
                int_cfg=pkg.IntCfg(dt=1e-3, chunks_cfg=pkg.ChunksCfg(num_cols=nc, num_rows=nr)), geom=geom)
    g, o = _pair(case)
    assert np.array_equal(g.download_cells()[0], o.download_cells()[0])
    assert np.array_equal(g.download_cells()[1], o.download_cells()[1])


@pytest.mark.parametrize("wall,rows,cols", [("periodic", 3, 3), ("periodic", 4, 5), ("periodic", 2, 6), ("periodic", 7, 2),
                                            ("rigid", 4, 5), ("rigid", 1, 6), ("rigid", 5, 2)])
def test_neighbor_sets_bit_exact(cuda_lib, wall, rows, cols):
    """The gather stencil == the reference's half stencil united with its mirror image, as a MULTISET per cell
    (2-wide periodic grids double count exactly like the reference tables, src/chunks.jl:61-87)."""
    case = H.newton_case(nx=12, ny=12, wall=wall, cells=(cols, rows))
    g, o = _pair(case)
    full = [Counter() for _ in range(rows * cols)]
    for c in range(rows * cols):
        for n in o.cell_neighbors(c):
            full[c][n] += 1
            full[n][c] += 1
    for c in range(rows * cols):
        assert Counter(g.cell_neighbors(c)) == full[c], c


# ---------------------------------------------------------------- forces
DYNS = {"lj": pkg.LenJonesCfg(sigma=1.0, epsilon=1.0), "harm": pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2)}


@pytest.mark.parametrize("chunks", [True, False])
@pytest.mark.parametrize("wall", ["periodic", "rigid"])
@pytest.mark.parametrize("dyn", ["lj", "harm"])
def test_forces_match_oracle(cuda_lib, dyn, wall, chunks):
    case = H.newton_case(nx=40, ny=36, dyn=DYNS[dyn], wall=wall, chunks=chunks, jitter=0.25)
    g, o = _pair(case)
    g.calc_forces()
    o.calc_forces()
    fg, fo = g.get_forces(), o.get_forces()
    assert H.rel_err(fg, fo) < TOL
    big = np.abs(fo) > 1e-3 * np.abs(fo).max()  # element-wise where no cancellation
    assert np.max(np.abs(fg[big] - fo[big]) / np.abs(fo[big])) < 1e-10


@pytest.mark.parametrize("rows,cols", [(2, 5), (5, 2), (3, 3)])
def test_forces_on_degenerate_periodic_grids(cuda_lib, rows, cols):
    """2-row / 2-column periodic grids: the reference double counts cell pairs; so must the device."""
    case = H.newton_case(nx=10, ny=10, dyn=DYNS["harm"], wall="periodic", cells=(cols, rows), jitter=0.3)
    g, o = _pair(case)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < TOL


@pytest.mark.parametrize("kind", ["szabo", "rtp"])
def test_self_propelled_forces_match_oracle(cuda_lib, kind):
    case = H.sp_case(kind, nx=40, ny=40)
    g, o = _pair(case)
    g.calc_forces()
    o.calc_forces()
    assert np.abs(o.get_forces()).max() > 0
    assert H.rel_err(g.get_forces(), o.get_forces()) < TOL


# ---------------------------------------------------------------- trajectories
@pytest.mark.parametrize("chunks", [True, False])
@pytest.mark.parametrize("wall", ["periodic", "rigid"])
@pytest.mark.parametrize("dyn", ["lj", "harm"])
def test_newton_trajectory(cuda_lib, dyn, wall, chunks):
    n = 32 if chunks else 16
    case = H.newton_case(nx=n, ny=n, dyn=DYNS[dyn], wall=wall, chunks=chunks, dt=0.001)
    g, o = _pair(case)
    for _ in range(4):
        g.step(25)
        o.step(25)
        g.sync_to_host()
        scale = case["geom"].length
        assert np.abs(g.state.pos - o.pos()).max() / scale < TOL
        assert H.rel_err(g.state.vel, o.second()) < 1e-11
        assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-10
    assert g.time_info.num_steps == 100 and g.time_info.time == o.time()[1]
    if chunks:
        o.update_chunks()  # device cells are always those of the current positions (the reference's NEXT update_chunks!)
        cg, _ = g.download_cells()
        co, _ = o.download_cells()
        assert np.array_equal(cg, co)


def test_quick_start_c1(cuda_lib):
    """BASELINE config C1 = README quick start: 10x10 LJ, RigidWalls rectangle, no chunks, dt = 0.01."""
    case = H.newton_case(nx=10, ny=10, wall="rigid", chunks=False, dt=0.01, jitter=0.0)
    g, o = _pair(case)
    pkg.run_system(g, tf=1)
    o.step(g.time_info.num_steps)
    assert g.time_info.num_steps == o.time()[0] == 100
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL
    assert H.rel_err(g.state.vel, o.second()) < 1e-11
    ke_g, pe_g = g.energies()
    ke_o, pe_o = o.energies()
    assert abs(ke_g - ke_o) < 1e-12 * abs(ke_o) and abs(pe_g - pe_o) < 1e-11 * abs(pe_o)


@pytest.mark.parametrize("kind,chunks", [("szabo", True), ("szabo", False), ("rtp", True), ("rtp", False)])
def test_self_propelled_trajectory_host_noise(cuda_lib, kind, chunks):
    """Host-noise mode: the caller passes the per-step draws that stand for the reference's global randn()/rand()."""
    n = 32 if chunks else 14
    case = H.sp_case(kind, nx=n, ny=n, chunks=chunks, jitter=0.9 if kind == "szabo" else 0.6)  # WCA blows up if overlapping
    g, o = _pair(case)
    N = n * n
    rng = np.random.default_rng(11)
    for _ in range(3):
        steps = 20
        noise = rng.standard_normal((steps, N)) if kind == "szabo" else rng.random((steps, N, 2)) * 0.01
        g.step(steps, noise)
        o.step(steps, noise)
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL
        assert np.abs(g.state.pol_angle - o.second()).max() < 1e-11
        assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-10


def test_philox_mode_statistics(cuda_lib):
    """Device RNG mode: rotational diffusion of free Szabo particles has variance 2 D_r t (no oracle stream parity)."""
    dyn = pkg.SzaboCfg(vo=0.0, mobility=0.0, relax_time=1e30, k_rep=0.0, k_adh=0.0, r_eq=1.0, r_max=1.1, rot_diff=0.5)
    pos, geom, rng = H.lattice(100, 100, dyn, offset=1.0)
    st = pkg.SelfPropelledState(pos=pos, pol_angle=np.zeros(len(pos)))
    g = pkg.System(state=st, space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                   int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(99, 99), device=pkg.CUDADevice(rng_mode="philox", seed=7)))
    g.step(100)
    g.sync_to_host()
    var = g.state.pol_angle.var()
    assert abs(var - 2 * 0.5 * 1.0) < 0.05
    assert abs(g.state.pol_angle.mean()) < 0.05


def test_philox_mode_statistics_rtp(cuda_lib):
    """Device RNG mode of update_rtp! (src/integration.jl:467-498): a particle tumbles with probability tumble_rate * dt per
    step and then gets a uniform angle in [0, 2 pi).  Free particles (no overlap, vo = 0), all starting at angle 10 (outside
    the range a tumble can produce): after k steps the fraction still at 10 is (1 - p)^k, the others are uniform."""
    dyn = pkg.RunTumbleCfg(vo=0.0, sigma=1.0, epsilon=1.0, tumble_rate=2.0)
    pos, geom, rng = H.lattice(200, 200, dyn, offset=1.0)
    n = len(pos)
    st = pkg.SelfPropelledState(pos=pos, pol_angle=np.full(n, 10.0))
    g = pkg.System(state=st, space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                   int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(150, 150), device=pkg.CUDADevice(rng_mode="philox", seed=11)))
    k = 40
    g.step(k)
    g.sync_to_host()
    ang = g.state.pol_angle
    untouched = ang == 10.0
    expect = (1 - 2.0 * 0.01) ** k
    assert abs(untouched.mean() - expect) < 4 * np.sqrt(expect * (1 - expect) / n)   # binomial, 4 sigma
    t = ang[~untouched]
    assert t.min() >= 0.0 and t.max() < 2 * np.pi
    m = len(t)
    assert abs(t.mean() - np.pi) < 4 * (2 * np.pi / np.sqrt(12)) / np.sqrt(m)
    assert abs(t.var() - (2 * np.pi) ** 2 / 12) < 0.15
    # the four quadrants are equally likely (chi-square with 3 degrees of freedom, p = 1e-4 at 21.1)
    cnt = np.histogram(t, bins=4, range=(0, 2 * np.pi))[0]
    assert ((cnt - m / 4) ** 2 / (m / 4)).sum() < 21.1
    # different steps draw different numbers: the positions never moved, the tumbled set keeps growing
    g.step(k)
    g.sync_to_host()
    assert (g.state.pol_angle == 10.0).mean() < untouched.mean()


# ---------------------------------------------------------------- walls, force walls, composite spaces
def test_rigid_circle_walls(cuda_lib):
    """examples/print_energy.jl geometry: LJ in a rigid circle, all pairs."""
    pos = np.array([[1, -2.5, 3.3, -4, 5], [-1.7, 2.1, -3.8, 4.4, -5.4]], dtype=float)
    vel = np.array([[0.3, 2, 5.7, 9.8, 3.0], [1, 0, 7.8, .12, 2.2]], dtype=float)
    dyn = pkg.LenJonesCfg(sigma=2, epsilon=4)
    space = pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=pkg.CircleCfg(radius=10, center=(0, 0)))
    mk = lambda: pkg.SecondLawState(pos=pos.copy(), vel=vel.copy())  # noqa: E731
    case = dict(mk=mk, space=space, dyn=dyn, int_cfg=pkg.IntCfg(dt=0.01), geom=None)
    g, o = _pair(case)
    g.step(100)
    o.step(100)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / 10 < TOL
    assert H.rel_err(g.state.vel, o.second()) < 1e-11
    assert not np.allclose(g.state.vel, vel.T)


def _wall_force_case(slippery=False):
    """examples/wall_force.jl: rigid box + PotentialWalls circle (mode outside) + PotentialWalls lines, with chunks."""
    dyn = pkg.HarmTruncCfg(k_rep=10, k_atr=1, dist_eq=1, dist_max=1.2)
    radius = pkg.particle_radius(dyn)
    pos, geom = pkg.rectangular_grid(10, 10, 0.4, radius)
    rng = np.random.default_rng(H.SEED)
    vel = pkg.random_vel(100, 1.0, rng=rng)
    geom = geom + pkg.RectangleCfg(length=geom.length, height=geom.height, bottom_left=(geom.length, 0.0))
    l, h = geom.size
    circle = pkg.CircleCfg(radius=3 * radius, center=(l / 2 + l / 4, h / 2))
    lines = pkg.LinesCfg([[(3 / 4 * l, 1 / 4 * h), (3 / 4 * l, 3 / 4 * h)], [(1 / 2 * l, 1 / 2 * h), (3 / 4 * l, 1 / 2 * h)]])
    wall_pot = pkg.HarmTruncCfg(k_rep=20, k_atr=0, dist_eq=radius, dist_max=radius * 1.1)
    if slippery:
        spaces = [(pkg.RigidWalls(), geom), (pkg.SlipperyWalls(), circle), (pkg.SlipperyWalls(), lines)]
    else:
        spaces = [(pkg.RigidWalls(), geom), (pkg.PotentialWalls(potential=wall_pot, mode="outside"), circle),
                  (pkg.PotentialWalls(potential=wall_pot), lines)]
    int_cfg = pkg.IntCfg(dt=0.001, chunks_cfg=pkg.ChunksCfg(num_cols=18, num_rows=18))
    mk = lambda: pkg.SecondLawState(pos=pos.copy(), vel=vel.copy())  # noqa: E731
    return dict(mk=mk, space=pkg.SpaceCfg(spaces), dyn=dyn, int_cfg=int_cfg, geom=geom)


@pytest.mark.parametrize("slippery", [False, True])
def test_composite_space_with_force_or_slippery_walls(cuda_lib, slippery):
    case = _wall_force_case(slippery)
    g, o = _pair(case)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < TOL
    g.step(2000)
    o.step(2000)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-10
    assert H.rel_err(g.state.vel, o.second()) < 1e-9
    moved = np.abs(g.state.pos - case["mk"]().pos).max()
    assert moved > 0.5  # the obstacles were actually reached


def test_active_mask(cuda_lib):
    """ParticleIds masks (src/states.jl:27-52): inactive slots are not binned and feel no force, but update_verlet!
    still drifts every slot."""
    n = 24
    mask = np.ones(n * n, dtype=bool)
    mask[::7] = False
    case = H.newton_case(nx=n, ny=n, dyn=DYNS["harm"], wall="rigid", active_mask=mask)
    g, o = _pair(case)
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(cg, co) and np.array_equal(ng, no) and (cg[~mask] == -1).all()
    g.step(50)
    o.step(50)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL
    assert np.all(g.get_forces()[~mask] == 0.0)
    assert np.abs(g.state.pos[~mask] - case["mk"]().pos[~mask]).max() > 0


# ---------------------------------------------------------------- quantities
def test_energies(cuda_lib):
    case = H.newton_case(nx=30, ny=30, wall="periodic", jitter=0.2)
    g, o = _pair(case)
    for mode in (0, 1):
        ke_g, pe_g = g.energies(mode)
        ke_o, pe_o = o.energies(mode)
        assert abs(ke_g - ke_o) <= 1e-13 * abs(ke_o)
        assert abs(pe_g - pe_o) <= 1e-11 * abs(pe_o)
    assert abs(g.energies(0)[1] - g.energies(1)[1]) > 1e-6  # stencil sum != exact all-pairs sum (labelled deviation)


# ---------------------------------------------------------------- errors, determinism
def test_out_of_grid_status(cuda_lib):
    dyn = pkg.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=1.0, dist_max=1.2)
    geom = pkg.RectangleCfg(length=10.0, height=6.0)
    pts = np.array([[1.0, 1.0], [2.0, 2.0]])
    vel = np.array([[0.0, 0.0], [-600.0, 0.0]])
    g = pkg.System(state=pkg.SecondLawState(pos=pts, vel=vel), space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom),
                   dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(num_cols=4, num_rows=3)))
    g.step(1)
    with pytest.raises(pkg.MaviError) as e:
        g.step(1)
    assert e.value.status == pkg.capi.ERR_OUT_OF_GRID


def test_outside_space_status(cuda_lib):
    dyn = pkg.LenJonesCfg(sigma=1, epsilon=1)
    geom = pkg.RectangleCfg(length=10.0, height=6.0)
    pts = np.array([[1.0, 1.0], [10.5, 2.0]])
    with pytest.raises(pkg.MaviError) as e:
        pkg.System(state=pkg.SecondLawState(pos=pts, vel=np.zeros_like(pts)),
                   space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom), dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.01))
    assert e.value.status == pkg.capi.ERR_OUTSIDE_SPACE


def test_bad_params_status(cuda_lib):
    case = H.newton_case(nx=8, ny=8, wall="periodic", cells=(1, 4))
    with pytest.raises(pkg.MaviError) as e:
        H.make_gpu(case)
    assert e.value.status == pkg.capi.ERR_BAD_PARAMS


def test_bitwise_reproducible(cuda_lib):
    """The stable (cell, id) order makes runs independent of atomic scheduling."""
    outs = []
    for _ in range(2):
        g = H.make_gpu(H.newton_case(nx=48, ny=48, jitter=0.2))
        g.step(40)
        g.sync_to_host()
        outs.append((g.state.pos.copy(), g.state.vel.copy(), g.get_forces()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("flag", ["tight", "resort"])
def test_repair_paths_agree(cuda_lib, flag):
    """The incremental tile repair, the full rebuild every step (MAVI_FLAG_RESORT_EVERY_STEP) and the
    overflow -> rebuild -> resume path (forced by MAVI_FLAG_TIGHT_TILES: no slack in the tiles, steps enqueued in
    batches and rolled back to the device-side step counter) all give the oracle's trajectory."""
    case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    fl = pkg.capi.FLAG_TIGHT_TILES if flag == "tight" else pkg.capi.FLAG_RESORT_EVERY_STEP
    case["int_cfg"] = pkg.IntCfg(dt=0.002, chunks_cfg=case["int_cfg"].chunks_cfg, device=pkg.CUDADevice(flags=fl))
    g, o = _pair(case)
    g.step(150)
    o.step(150)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL
    assert H.rel_err(g.state.vel, o.second()) < 1e-10
    assert g.time_info.num_steps == 150 and g.time_info.time == o.time()[1]
    if flag == "tight":
        assert g.rebuild_count() > 0  # the overflow path really ran
    # the device keeps the binning of the CURRENT positions (what the reference's next update_chunks! produces)
    o.update_chunks()
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(ng, no) and np.array_equal(cg, co)


def _with_flags(case, flags):
    ic = case["int_cfg"]
    case = dict(case)
    import dataclasses
    case["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=dataclasses.replace(ic.device, flags=flags))
    return case


@pytest.mark.parametrize("kind", ["lj_periodic_hot", "harm_rigid", "force_walls", "slippery", "masked", "tight"])
def test_force_carry_bitwise(cuda_lib, kind):
    """Force carry (default): the second pair pass of step n also produces F1 and the drift of step n+1, and only the
    particles that have a re-binned (or wall-moved) particle in their stencil are recomputed with the fresh cell
    lists.  Must be BIT-identical to running the full first pass every step (MAVI_FLAG_NO_FORCE_CARRY), also across
    downloads, calc_forces! calls, re-uploads and overflow -> rebuild events in the middle of a run."""
    if kind == "lj_periodic_hot":  # many cell changes and periodic wraps per step
        case = H.newton_case(nx=40, ny=36, wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    elif kind == "harm_rigid":
        case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="rigid", jitter=0.3, vmax=3.0, dt=0.002)
    elif kind == "force_walls":
        case = _wall_force_case(False)
    elif kind == "slippery":
        case = _wall_force_case(True)
    elif kind == "masked":
        mask = np.ones(32 * 32, dtype=bool)
        mask[::5] = False
        case = H.newton_case(nx=32, ny=32, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=2.0, dt=0.002, active_mask=mask)
    else:
        case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    base = pkg.capi.FLAG_TIGHT_TILES if kind == "tight" else 0
    a = H.make_gpu(_with_flags(case, base))
    b = H.make_gpu(_with_flags(case, base | pkg.capi.FLAG_NO_FORCE_CARRY))

    def same():
        a.sync_to_host()
        b.sync_to_host()
        assert np.array_equal(a.state.pos, b.state.pos)
        assert np.array_equal(a.state.vel, b.state.vel)
        assert np.array_equal(a.get_forces(), b.get_forces())

    for n in (1, 2, 37, 120):
        a.step(n)
        b.step(n)
        same()
    a.calc_forces()   # calc_forces! between steps must not disturb the carried state
    b.calc_forces()
    same()
    a.step(40)
    b.step(40)
    same()
    a.upload_state()  # re-upload (the host copy is the synced state): carry is invalidated and rebuilt
    b.upload_state()
    a.step(25)
    b.step(25)
    same()
    assert np.array_equal(a.download_cells()[0], b.download_cells()[0])
    if kind == "tight":
        assert a.rebuild_count() > 0


@pytest.mark.parametrize("dyn,wall", [("lj", "periodic"), ("harm", "rigid")])
def test_small_blocks_agree(cuda_lib, dyn, wall):
    """MAVI_FLAG_SMALL_BLOCKS (3 tile columns per CTA: many blocks per tile row, side columns shared between blocks) gives
    bit-identical results to the default block size."""
    case = H.newton_case(nx=64, ny=40, dyn=DYNS[dyn], wall=wall, jitter=0.3, vmax=2.0, dt=0.002)
    a = H.make_gpu(_with_flags(case, 0))
    b = H.make_gpu(_with_flags(case, pkg.capi.FLAG_SMALL_BLOCKS))
    a.step(60)
    b.step(60)
    a.sync_to_host()
    b.sync_to_host()
    assert np.array_equal(a.state.pos, b.state.pos) and np.array_equal(a.state.vel, b.state.vel)
    assert np.array_equal(a.get_forces(), b.get_forces())


@pytest.mark.parametrize("kind", ["lj_periodic_hot", "harm_rigid", "force_walls", "masked", "lj_small_grid", "lj_dense", "szabo", "rtp"])
def test_pipelined_kernels_match_legacy_staging(cuda_lib, kind):
    """The persistent producer/consumer force kernels (bulk-async staging through mbarriers, the default) must be
    BIT-identical to the round-1 kernels (MAVI_FLAG_LEGACY_STAGING: one CTA per tile block, cp.async behind CTA barriers):
    same neighbour order, same arithmetic — through re-binning, wall fix-ups, masks, chunk splitting (small blocks), dense
    columns that do not fit the staging area, and calc_forces! calls."""
    LEG = pkg.capi.FLAG_LEGACY_STAGING
    extra = 0
    sp = kind in ("szabo", "rtp")
    if kind == "lj_periodic_hot":
        case = H.newton_case(nx=70, ny=66, wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
        extra = pkg.capi.FLAG_SMALL_BLOCKS   # many work items per tile row: the work counter and both buffers cycle
    elif kind == "harm_rigid":
        case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="rigid", jitter=0.3, vmax=3.0, dt=0.002)
    elif kind == "force_walls":
        case = _wall_force_case(False)
    elif kind == "masked":
        mask = np.ones(32 * 32, dtype=bool)
        mask[::5] = False
        case = H.newton_case(nx=32, ny=32, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=2.0, dt=0.002, active_mask=mask)
    elif kind == "lj_small_grid":   # 2-row / 2-column periodic wrap
        case = H.newton_case(nx=6, ny=6, wall="periodic", jitter=0.3, vmax=1.0, dt=0.002, cells=(2, 2))
    elif kind == "lj_dense":        # 9 particles per cell, ~290 per tile: a block needs several chunks of the staging area
        case = H.newton_case(nx=96, ny=96, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=1.0, dt=0.001, cells=(32, 32))
    else:
        case = H.sp_case(kind, nx=48, ny=40, rot_diff=0.05, jitter=0.6 if kind == "rtp" else 0.9)
    a = H.make_gpu(_with_flags(case, extra))
    b = H.make_gpu(_with_flags(case, extra | LEG))
    n = len(a.state.pos)
    rng = np.random.default_rng(17)
    for steps in (1, 2, 37, 120):
        noise = None
        if kind == "szabo":
            noise = rng.standard_normal((steps, n))
        elif kind == "rtp":
            noise = rng.random((steps, 2 * n))
            noise[:, 0::2] *= 0.02
        a.step(steps, noise)
        b.step(steps, noise)
        a.sync_to_host()
        b.sync_to_host()
        assert np.array_equal(a.state.pos, b.state.pos)
        assert np.array_equal(a.state.second, b.state.second)
        assert np.array_equal(a.get_forces(), b.get_forces())
    if not sp:
        a.calc_forces()
        b.calc_forces()
        a.step(40)
        b.step(40)
        a.sync_to_host()
        b.sync_to_host()
        assert np.array_equal(a.state.pos, b.state.pos) and np.array_equal(a.get_forces(), b.get_forces())
    assert np.array_equal(a.download_cells()[0], b.download_cells()[0])


def test_kernels_actually_launch(cuda_lib):
    g = H.make_gpu(H.newton_case(nx=16, ny=16))
    n0 = g.launch_count()
    g.step(3)
    assert g.launch_count() - n0 >= 3 * 2


# ---------------------------------------------------------------- full-size properties (BASELINE config C2: 1M LJ periodic)
def test_c2_one_million_properties(cuda_lib):
    """At N = 1M the oracle is too slow for per-step comparison; use size-independent properties: total momentum is
    conserved (periodic, pair forces antisymmetric), the physical order is sorted by cell, the cell histogram sums to
    N, and a sample of cells matches the oracle's binning of the same positions."""
    nx = ny = 1000
    case = H.newton_case(nx=nx, ny=ny, wall="periodic", jitter=0.05)
    g = H.make_gpu(case)
    p0 = g.state.vel.sum(0)
    g.step(20)
    g.sync_to_host()
    f = g.get_forces()
    assert np.abs(f.sum(0)).max() < 1e-9 * np.abs(f).sum()
    assert np.abs(g.state.vel.sum(0) - p0).max() < 1e-9 * np.abs(g.state.vel).sum()
    start, ids = g.download_cell_lists()
    assert start[-1] == nx * ny and np.all(np.diff(start) >= 0)
    cell, counts = g.download_cells()
    assert counts.sum() == nx * ny and np.array_equal(np.bincount(cell, minlength=len(counts)), counts)
    assert np.array_equal(cell[ids], np.repeat(np.arange(len(counts)), counts))  # sorted by cell
    seg = ids[start[12345]:start[12345 + 50]]
    assert np.all(np.diff(cell[seg]) >= 0)


def test_slab_mode_grows_tiles_proactively(cuda_lib):
    """Slab mode cannot roll an overflowed step back (the other ranks have moved on), so tiles grow BEFORE they overflow: at
    every synchronisation point a tile above 85 % of the capacity makes all ranks rebuild with larger tiles.  Forced here by
    MAVI_FLAG_TIGHT_TILES (capacity = fullest tile at upload) in one-rank slab mode; results stay bit-identical."""
    SELF, TIGHT = pkg.capi.FLAG_SLAB_SELF, pkg.capi.FLAG_TIGHT_TILES
    case = H.newton_case(nx=64, ny=40, wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    a = H.make_gpu(_with_flags(case, 0))
    b = H.make_gpu(_with_flags(case, SELF | TIGHT))
    assert b.rebuild_count() >= 1           # grown right after the upload
    cap0 = b.counters()["tile_cap"]
    for steps in (40, 100):
        a.step(steps)
        b.step(steps)
        a.sync_to_host()
        ids, pos, vel, frc = b.download_local()
        o = np.argsort(ids)
        assert np.array_equal(pos[o], a.state.pos) and np.array_equal(vel[o], a.state.vel) and np.array_equal(frc[o], a.get_forces())
    assert b.counters()["tile_cap"] >= cap0


# ---------------------------------------------------------------- the x-slab machinery on ONE GPU (MAVI_FLAG_SLAB_SELF)
@pytest.mark.parametrize("kind,flags", [("lj", 0), ("lj", 8), ("harm", 8), ("szabo", 0), ("szabo_noise", 0), ("rtp", 0)])
def test_slab_self_mode_matches_plain(cuda_lib, kind, flags):
    """One-rank slab mode: halo columns, the two-stream step pipeline (flags=8: several CTAs per tile row), boundary
    recompute and local copies in place of NCCL.  Must reproduce the plain single-domain run: bit-identical for the
    Newton dynamics, to rounding for Szabo (the minimum image is applied through the seam)."""
    SELF = pkg.capi.FLAG_SLAB_SELF
    sp = kind in ("szabo", "szabo_noise", "rtp")
    if sp:
        # szabo_noise / rtp: host noise rows are indexed by the ORIGINAL id (global row in slab mode) through migration
        case = H.sp_case("rtp" if kind == "rtp" else "szabo", nx=40, ny=30, rot_diff=0.0 if kind == "szabo" else 0.05,
                         jitter=0.6 if kind == "rtp" else 0.9)  # WCA blows up if particles overlap
        ic = case["int_cfg"]
        mkdev = lambda f: pkg.CUDADevice(rng_mode="host_noise", flags=f)  # noqa: E731
    else:
        case = H.newton_case(nx=64, ny=40, dyn=DYNS[kind], wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
        ic = case["int_cfg"]
        mkdev = lambda f: pkg.CUDADevice(flags=f)  # noqa: E731
    a_case, b_case = dict(case), dict(case)
    a_case["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=mkdev(flags))
    b_case["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=mkdev(flags | SELF))
    a = H.make_gpu(a_case)
    b = H.make_gpu(b_case)
    n = len(a.state.pos)
    rng = np.random.default_rng(99)
    for steps in (1, 33, 80):
        noise = None
        if kind == "szabo_noise":
            noise = rng.standard_normal((steps, n))
        elif kind == "rtp":
            noise = rng.random((steps, 2 * n))
            noise[:, 0::2] *= 0.02   # u < tumble_rate * dt = 0.001 for ~5 % of the particles per step
        a.step(steps, noise)
        b.step(steps, noise)
        a.sync_to_host()
        ids, pos, second, forces = b.download_local()
        assert len(ids) == n and np.array_equal(np.sort(ids), np.arange(n))
        o = np.argsort(ids)
        if sp:
            assert np.abs(pos[o] - a.state.pos).max() / case["geom"].length < 1e-12
            assert np.abs(second[o] - a.state.pol_angle).max() < 1e-10
        else:
            assert np.array_equal(pos[o], a.state.pos)
            assert np.array_equal(second[o], a.state.vel)
            assert np.array_equal(forces[o], a.get_forces())
