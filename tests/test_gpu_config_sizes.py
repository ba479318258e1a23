"""Device vs oracle AT THE BASELINE CONFIG SIZES (BASELINE.json configs C2-C4 and the 16M headline workload).

The small-system tests (test_gpu_core.py, test_rings.py) cover every code path on <= 72x40 particles; these run the
same comparisons where the numbers of bench.py are measured: 1M / 16M particles, > 100 tile rows, hundreds of tile-block
CTAs, chunk splitting and the magic-number divisions.  The oracle runs in Threaded mode on all host cores (its
force-summation order then differs from the sequential one by a re-association, far below the 1e-12 bar).
Tolerances: positions relative to the box, velocities / forces norm-wise, cell assignment bit-exact.
"""
import os

import numpy as np
import pytest

import helpers as H

pkg = H.pkg
pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _noise(n, steps, seed):
    return np.random.default_rng(seed).standard_normal((steps, n))


def _cells_equal(g, o, rebin=False):
    """Cell assignment and per-cell counts, bit-exact.  After steps the reference's Chunks are STALE (binned at the start
    of the last step, src/integration.jl:507-515) while the device layout is already the fresh binning of the final
    positions: rebin=True runs update_chunks! on both sides first."""
    if rebin:
        g.update_chunks()
        o.update_chunks()
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    return np.array_equal(cg, co) and np.array_equal(ng, no)


def test_c2_lj_one_million_matches_oracle(cuda_lib):
    """BASELINE config C2: LJ gas, 1000 x 1000 = 1M particles, periodic rectangle, 900 x 900 Chunks, Float64.
    calc_forces! and 20 newton_step!s against the oracle: forces / positions / velocities 1e-12, cells bit-exact."""
    case = H.newton_case(nx=1000, ny=1000, wall="periodic", jitter=0.05)
    g, o = H.make_gpu(case), H.make_oracle(case, threads=THREADS)
    assert _cells_equal(g, o)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12
    g.step(20)
    o.step(20)
    g.sync_to_host()
    L = case["geom"].length
    assert np.abs(g.state.pos - o.pos()).max() / L < 1e-12
    assert H.rel_err(g.state.vel, o.second()) < 1e-12
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12
    assert _cells_equal(g, o, rebin=True)
    sg, ig = g.download_cell_lists()
    so, io = o.download_cell_lists()
    assert np.array_equal(sg, so) and np.array_equal(ig, io)  # chunk_particles: ascending ids inside every cell
    assert g.time_info.num_steps == 20


def test_c2_harmtrunc_one_million_rigid_matches_oracle(cuda_lib):
    """test/tests_experiments.jl:11-51 shape (HarmTrunc gas, rigid walls, chunks) at 1M particles, 10 steps."""
    dyn = pkg.HarmTruncCfg(k_rep=10.0, k_atr=1.0, dist_eq=1.0, dist_max=1.3)
    case = H.newton_case(nx=1000, ny=1000, dyn=dyn, wall="rigid", jitter=0.3, vmax=1.0, dt=0.002)
    g, o = H.make_gpu(case), H.make_oracle(case, threads=THREADS)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12   # identical positions: the force evaluation itself
    g.step(10)
    o.step(10)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert H.rel_err(g.state.vel, o.second()) < 1e-12
    # after steps the two sides hold positions that differ in the last bit (ulp(1100) = 2e-13): a stiff law (k_rep = 10)
    # turns that into a few 1e-12 of force
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-11
    assert _cells_equal(g, o, rebin=True)


def test_c3_szabo_one_million_matches_oracle(cuda_lib):
    """BASELINE config C3's dynamics (examples/szabo.jl parameters, lattice offset 1, cells (n-1)^2, dt 0.01) at 1M
    particles, 10 szabo_step!s with host noise = the draws the reference would make (rot_diff = 0.01)."""
    case = H.sp_case("szabo", nx=1000, ny=1000, jitter=0.9, rot_diff=0.01)
    g, o = H.make_gpu(case), H.make_oracle(case, threads=THREADS)
    n = 1000 * 1000
    noise = _noise(n, 10, seed=7)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12   # identical positions
    g.step(10, noise)
    o.step(10, noise)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert np.abs(g.state.pol_angle - o.second()).max() < 1e-11
    # Szabo's repulsion is k_rep / (r_max - r_eq) = 100 per unit length: last-bit position differences (ulp(1500) = 2e-13)
    # show up as 1e-11 of force after steps
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-10
    assert _cells_equal(g, o, rebin=True)


def test_c4_rings_100k_matches_oracle(cuda_lib):
    """BASELINE config C4: Mavi.Rings, 400 x 250 = 100k rings x 10 particles, periodic, parameters of
    test/tests_rings/rings_utils.jl:35-53; constructor state + 5 step!s with host noise."""
    case = H.rings_case("normal", 400, 250)
    g, o = H.make_gpu_rings(case), H.make_oracle(case, threads=THREADS)
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-11   # area force: shoelace cancellation in a 1500-wide box
    assert _cells_equal(g, o)
    noise = _noise(case["num_rings"], 5, seed=11)
    g.step(5, noise)
    o.step(5, noise)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert np.abs(g.state.pol - o.second()).max() < 1e-11
    # after steps positions differ in the last bit (ulp(1500) = 2e-13); the shoelace area of a ring 1500 away from the origin
    # turns that into ~|r| * ulp * n = 3e-9 of area and the area force follows (the reference's formula, src/rings/integration.jl:103-116)
    assert H.rel_err(g.get_forces(), o.get_forces()) < 5e-9
    areas_g, cms_g, cont_g = g.rings_info()
    areas_o, cms_o, cont_o = o.rings_info()
    assert H.rel_err(areas_g, areas_o) < 5e-9 and H.rel_err(cms_g, cms_o) < 1e-12 and H.rel_err(cont_g, cont_o) < 1e-12


def test_c4_rings_two_types_100k_matches_oracle(cuda_lib):
    """The two-type InteractionMatrix fixture (test/tests_rings/rings_utils.jl:193-296) at 100k rings."""
    case = H.rings_case("types", 400, 250)
    g, o = H.make_gpu_rings(case), H.make_oracle(case, threads=THREADS)
    noise = _noise(case["num_rings"], 5, seed=12)
    g.step(5, noise)
    o.step(5, noise)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert H.rel_err(g.get_forces(), o.get_forces()) < 5e-9     # shoelace cancellation, see the one-type test
    areas_g, cms_g, cont_g = g.rings_info()
    areas_o, cms_o, cont_o = o.rings_info()
    assert H.rel_err(areas_g, areas_o) < 5e-9 and H.rel_err(cms_g, cms_o) < 1e-12 and H.rel_err(cont_g, cont_o) < 1e-12


def _ring_bbox_hits(rp, circle):
    """filter_rings_pos of examples/rings_circle_obs.jl:13-26: bounding box of the ring vs bounding box of the circle."""
    lo, hi = rp.min(1), rp.max(1)
    cl = np.asarray(circle.center) - circle.radius
    ch = np.asarray(circle.center) + circle.radius
    return (lo[:, 0] <= ch[0]) & (hi[:, 0] >= cl[0]) & (lo[:, 1] <= ch[1]) & (hi[:, 1] >= cl[1])


@pytest.mark.parametrize("main_wall", ["periodic"])
def test_rings_with_circle_obstacles_matches_oracle(cuda_lib, main_wall):
    """examples/rings_circle_obs.jl:60-79: rings on a grid, two SlipperyWalls circle obstacles in a periodic rectangle, rings
    overlapping an obstacle removed; 150 step!s.  (RigidWalls as the main wall of a RingsSystem is not a valid reference
    configuration: walls!(::RigidWalls, ::RectangleCfg) flips `state.vel`, which a RingsState does not have,
    src/integration.jl:271-285.)"""
    from mavi_jl_b200.rings import configs as rc
    from mavi_jl_b200.rings import init_states as ri
    from mavi_jl_b200.rings.states import RingsState

    inter = rc.HarmTruncCfg(k_rep=40, k_atr=4, dist_eq=1, dist_max=1 + 0.2)
    dyn = rc.RingsCfg(p0=3.5, relax_time=1.0, vo=1.0, mobility=1.0, rot_diff=0.05, k_area=1.0, k_spring=40.0, l_spring=1.0,
                      num_particles=10, interaction_finder=inter)
    rings_pos, geom = ri.rectangular_grid(num_cols=16, num_rows=16, num_particles=10, p_radius=dyn.particle_radius(),
                                          pad_x=0.1, pad_y=0.1)
    ring_r = ri.get_ring_radius(dyn.particle_radius(), 10)
    bl = np.asarray(geom.bottom_left)
    c1 = pkg.CircleCfg(radius=2 * ring_r, center=[bl[0] + geom.length / 2, bl[1] + geom.height / 4])
    c2 = pkg.CircleCfg(radius=2 * ring_r, center=[bl[0] + geom.length / 2, bl[1] + 3 * geom.height / 4])
    keep = ~(_ring_bbox_hits(rings_pos, c1) | _ring_bbox_hits(rings_pos, c2))
    rings_pos = np.ascontiguousarray(rings_pos[keep])
    nr = len(rings_pos)
    assert 0 < nr < 256
    main = pkg.PeriodicWalls() if main_wall == "periodic" else pkg.RigidWalls()
    space = pkg.SpaceCfg([(main, geom), (pkg.SlipperyWalls(), c1), (pkg.SlipperyWalls(), c2)])
    rng = np.random.default_rng(5)
    pol = ri.random_pol(nr, rng=rng)
    max_size = inter.dist_max * 1.1
    chunks = pkg.ChunksCfg(int(geom.length // max_size), int(geom.height // max_size))
    int_cfg = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=chunks, device=pkg.CUDADevice(rng_mode="host_noise"))
    case = dict(mk=lambda: RingsState(rings_pos=rings_pos.copy(), pol=pol.copy()), space=space, dyn=dyn, int_cfg=int_cfg,
                geom=geom, num_rings=nr)
    g, o = H.make_gpu_rings(case), H.make_oracle(case)
    for block in range(3):
        noise = _noise(nr, 50, seed=block)
        g.step(50, noise)
        o.step(50, noise)
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / geom.length < 1e-11
        assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-9
    # the obstacles did something: some particle sits on a circle (slippery walls project onto the surface)
    d1 = np.linalg.norm(g.state.pos - np.asarray(c1.center), axis=1)
    d2 = np.linalg.norm(g.state.pos - np.asarray(c2.center), axis=1)
    assert min(d1.min(), d2.min()) >= c1.radius - 1e-9


def test_headline_16m_three_steps_match_oracle(cuda_lib):
    """The bench.py workload itself (LJ lattice 4000 x 4000 = 16M, periodic, 3600 x 3600 chunks, dt = 0.001, |v| <= 0.2):
    3 newton_step!s (first pass + two carried steps) against the Threaded oracle.  Needs ~8 GB of host memory.
    The lattice is PERFECT, so net forces are cancellation residues (|F| ~ 1e-2 of one pair force): the force error is
    taken relative to max(|F|_inf, one pair force at the lattice spacing) (SURVEY.md 7 "parity metric under cancellation");
    the plain norm-wise figure is bounded at 1e-11."""
    import bench
    w = bench.lj_workload(pkg, 4000, 4000)
    case = dict(mk=lambda: pkg.SecondLawState(pos=w["pos"].copy(), vel=w["vel"].copy()), space=w["space"], dyn=w["dyn"],
                int_cfg=w["int_cfg"], geom=w["geom"])
    g = H.make_gpu(case)
    g.step(3)
    g.sync_to_host()
    gp, gv, gf = g.state.pos, g.state.vel, g.get_forces()
    g.update_chunks()
    cg, ng = g.download_cells()
    g.close()
    o = H.make_oracle(case, threads=THREADS)
    o.step(3)
    of = o.get_forces()
    assert np.abs(gp - o.pos()).max() / w["geom"].length < 1e-12
    assert H.rel_err(gv, o.second()) < 1e-12
    # positions agree to the last bit or two (ulp(4700) = 9e-13) and the LJ force goes like r^-13: one ulp of a coordinate
    # is 13 * 9e-13 / 1.35 = 9e-12 of a pair force.  The force EVALUATION is checked on identical positions below.
    assert np.abs(gf - of).max() / np.abs(of).max() < 2e-11
    # forces from identical positions: the oracle takes the device's state and both run clean + update_chunks + calc_forces
    g2 = H.make_gpu(dict(case, mk=lambda: pkg.SecondLawState(pos=gp.copy(), vel=gv.copy())))
    g2.calc_forces()
    gf2 = g2.get_forces()
    g2.close()
    o2 = H.make_oracle(dict(case, mk=lambda: pkg.SecondLawState(pos=gp.copy(), vel=gv.copy())), threads=THREADS)
    o2.calc_forces()
    assert H.rel_err(gf2, o2.get_forces()) < 1e-12
    o.update_chunks()
    co, no = o.download_cells()
    assert np.array_equal(cg, co) and np.array_equal(ng, no)
