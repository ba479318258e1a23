"""CPU tests pinning the oracle (oracle/) with known answers derived from the reference SOURCE and the
reference's own portable invariants (chunks == all-pairs, Threaded == Sequencial; test/tests_rings/tests_general.jl:16-45).
Julia is unavailable, the reference holds no golden vectors for this path: parity of the oracle itself is UNPINNED.
"""
import math
from fractions import Fraction

import numpy as np
import pytest

import helpers as H

pkg = H.pkg
HT, LJ = 0, 1  # MAVI_POT_*


# ---------------------------------------------------------------- Base.div semantics (src/chunks.jl:129-130)
def _exact_trunc_div(x, y):
    q = Fraction(x) / Fraction(y)
    return float(math.trunc(q))


def test_julia_div_is_trunc_of_exact_quotient(oracle):
    rng = np.random.default_rng(1)
    ys = [1.5, 0.1, 1.0 / 3.0, 1.5231e-3, 1.1111111111111112, 0.7]
    for y in ys:
        xs = list(rng.uniform(0, 4000, 200)) + [k * y for k in range(0, 3000, 7)] + \
             [np.nextafter(k * y, 0) for k in range(1, 3000, 11)] + [np.nextafter(k * y, 1e9) for k in range(1, 3000, 13)]
        for x in xs:
            assert oracle.julia_div(x, y) == _exact_trunc_div(x, y), (x, y)
    # slightly negative coordinates land in cell 1 (round toward zero), SURVEY.md 7
    assert oracle.julia_div(-1e-3, 1.5) == 0.0
    assert oracle.julia_div(-1.6, 1.5) == -1.0


def test_julia_div_differs_from_naive_division(oracle):
    """x = fl(k*y) below k*y: fl(x/y) == k but the exact quotient truncates to k-1 (SURVEY.md 7 'Hard parts')."""
    found = 0
    for y in (0.1, 1.0 / 3.0, 0.7, 1.1111111111111112):
        for k in range(1, 4000):
            x = k * y
            if math.trunc(x / y) != _exact_trunc_div(x, y):
                found += 1
                assert oracle.julia_div(x, y) == _exact_trunc_div(x, y)
    assert found > 0


# ---------------------------------------------------------------- pair laws
def test_lj_force_zero_at_minimum_and_sign(oracle):
    sigma, eps = 1.3, 0.7
    rmin = 2 ** (1 / 6) * sigma
    f = oracle.potential_force(LJ, [sigma, eps], [rmin, 0.0])
    assert abs(f[0]) < 1e-12 and f[1] == 0.0
    assert oracle.potential_force(LJ, [sigma, eps], [0.9 * rmin, 0.0])[0] > 0  # repulsive: pushes i away from j
    assert oracle.potential_force(LJ, [sigma, eps], [1.5 * rmin, 0.0])[0] < 0  # attractive, and NO cutoff
    d = 7.3
    fmod = 4 * eps * (12 * sigma ** 12 / d ** 13 - 6 * sigma ** 6 / d ** 7)
    f = oracle.potential_force(LJ, [sigma, eps], [d * 0.6, d * 0.8])
    assert np.allclose(f, [fmod * 0.6, fmod * 0.8], rtol=1e-13)


def test_harmtrunc_force(oracle):
    par = [10.0, 3.0, 1.0, 1.2]  # k_rep, k_atr, dist_eq, dist_max
    assert np.all(oracle.potential_force(HT, par, [1.0, 0.0]) == 0.0)          # zero at d_eq
    assert np.all(oracle.potential_force(HT, par, [1.2000001, 0.0]) == 0.0)    # zero beyond d_max
    f = oracle.potential_force(HT, par, [1.2, 0.0])                           # at d_max: still attractive (strict >)
    assert np.isclose(f[0], -3.0 * (1.2 / 1.0 - 1))
    f = oracle.potential_force(HT, par, [0.0, 0.8])
    assert np.isclose(f[1], -10.0 * (0.8 - 1))


def test_szabo_force_is_not_normalised(oracle):
    par = [1.0, 1.0, 1.0, 10.0, 0.75, 1.0, 1.1, 0.01]
    dr = np.array([0.3, 0.4]) * 1.8  # d = 0.9 < r_eq
    f = oracle.szabo_interaction(par, dr)
    fmod = 10.0 / (1.1 - 1.0)
    assert np.allclose(f, -fmod * (0.9 - 1.0) * dr, rtol=1e-13)  # multiplies dr, not dr/d (src/integration.jl:85-86)
    dr = np.array([1.05, 0.0])
    assert np.allclose(oracle.szabo_interaction(par, dr), -(0.75 / 1.0) * 0.05 * dr, rtol=1e-12)
    assert np.all(oracle.szabo_interaction(par, [1.1000001, 0.0]) == 0.0)


def test_rtp_is_wca(oracle):
    par = [1.0, 1.0, 1.0, 1.0]
    cutoff = 2 ** (1 / 6)
    assert np.all(oracle.rtp_interaction(par, [cutoff * 1.0000001, 0.0]) == 0.0)
    d = 0.95
    f = oracle.rtp_interaction(par, [d, 0.0])
    assert np.isclose(f[0], 4 * (12 / d ** 13 - 6 / d ** 7), rtol=1e-13)


# ---------------------------------------------------------------- calc_diff / walls
def test_min_image_is_strict_at_half_box(oracle):
    case = H.newton_case(nx=8, ny=8, chunks=False)
    o = H.make_oracle(case)
    L, Hh = case["geom"].length, case["geom"].height
    assert np.array_equal(o.calc_diff([L / 2, 0.0], [0.0, 0.0]), [L / 2, 0.0])       # exactly half: NOT wrapped
    d = o.calc_diff([np.nextafter(L / 2, L), 0.0], [0.0, 0.0])
    assert d[0] < 0 and np.isclose(d[0], -L / 2)
    d = o.calc_diff([0.1, 0.2], [0.1, Hh - 0.1])
    assert np.allclose(d, [0.0, 0.3])


def test_cell_index_conventions(oracle):
    """Rows count downward from the top edge, x == L clamps to the last column, only n+1 is clamped."""
    dyn = pkg.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=1.0, dist_max=1.2)
    geom = pkg.RectangleCfg(length=10.0, height=6.0)
    pts = np.array([[0.0, 6.0], [10.0, 0.0], [9.999, 0.001], [0.5, 5.5], [5.0, 3.0], [2.5, 0.0]])
    st = pkg.SecondLawState(pos=pts, vel=np.zeros_like(pts))
    o = oracle.OracleSystem(state=st, space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                            int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(num_cols=4, num_rows=3)), lower=H.lower)
    cell, counts = o.download_cells()
    rows, cols = cell % 3, cell // 3
    assert list(rows) == [0, 2, 2, 0, 1, 2]  # y=6 (top edge) -> row 1; y=0 -> row n+1 clamped to n
    assert list(cols) == [0, 3, 3, 0, 2, 1]  # x=10 -> col n+1 clamped; x=2.5 = 1*2.5 -> col 2 (0-based 1)
    assert counts.sum() == 6


def test_out_of_grid_is_an_error(oracle):
    dyn = pkg.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=1.0, dist_max=1.2)
    geom = pkg.RectangleCfg(length=10.0, height=6.0)
    pts = np.array([[1.0, 1.0], [2.0, 2.0]])
    vel = np.array([[0.0, 0.0], [-600.0, 0.0]])
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=vel), space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom),
                            dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(num_cols=4, num_rows=3)), lower=H.lower)
    o.step(1)  # particle 2 is now at x = -4 (beyond one cell width below 0): next binning throws BoundsError in the reference
    with pytest.raises(oracle.OracleError) as e:
        o.step(1)
    assert e.value.status == 2


def test_outside_space_rejected_at_construction(oracle):
    dyn = pkg.LenJonesCfg(sigma=1, epsilon=1)
    geom = pkg.RectangleCfg(length=10.0, height=6.0)
    pts = np.array([[1.0, 1.0], [10.5, 2.0]])
    with pytest.raises(oracle.OracleError) as e:
        oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=np.zeros_like(pts)),
                            space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                            int_cfg=pkg.IntCfg(dt=0.01), lower=H.lower)
    assert e.value.status == 6


def _cells(o, rows):
    return lambda lst: sorted((c % rows + 1, c // rows + 1) for c in lst)


@pytest.mark.parametrize("rows,cols", [(3, 3), (4, 5), (5, 4)])
def test_periodic_half_stencil_tables(oracle, rows, cols):
    """src/chunks.jl:61-87: (i+1,j),(i+1,j+1),(i,j+1),(i-1,j+1) with wrap 0->n, n+1->1."""
    case = H.newton_case(nx=12, ny=12, wall="periodic", cells=(cols, rows))
    o = H.make_oracle(case)
    wrap = lambda x, n: n if x == 0 else (1 if x == n + 1 else x)  # noqa: E731
    for i in range(1, rows + 1):
        for j in range(1, cols + 1):
            got = [(c % rows + 1, c // rows + 1) for c in o.cell_neighbors((i - 1) + rows * (j - 1))]
            want = [(wrap(i + 1, rows), wrap(j, cols)), (wrap(i + 1, rows), wrap(j + 1, cols)),
                    (wrap(i, rows), wrap(j + 1, cols)), (wrap(i - 1, rows), wrap(j + 1, cols))]
            assert got == want


def test_walled_half_stencil_tables(oracle):
    """src/chunks.jl:89-118."""
    rows, cols = 4, 5
    case = H.newton_case(nx=12, ny=12, wall="rigid", cells=(cols, rows))
    o = H.make_oracle(case)
    nb = lambda i, j: [(c % rows + 1, c // rows + 1) for c in o.cell_neighbors((i - 1) + rows * (j - 1))]  # noqa: E731
    assert nb(1, 1) == [(2, 1), (2, 2), (1, 2)]
    assert nb(2, 3) == [(3, 2), (3, 3), (3, 4), (2, 4)]
    assert nb(3, 5) == [(4, 4), (4, 5)]
    assert nb(4, 2) == [(4, 3)]
    assert nb(4, 5) == []
    # every unordered pair of 8-adjacent cells appears exactly once
    seen = set()
    for i in range(1, rows + 1):
        for j in range(1, cols + 1):
            for q in nb(i, j):
                key = frozenset([(i, j), q])
                assert key not in seen
                seen.add(key)
    adj = {frozenset([(i, j), (i + a, j + b)]) for i in range(1, rows + 1) for j in range(1, cols + 1)
           for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0) and 1 <= i + a <= rows and 1 <= j + b <= cols}
    assert seen == adj


def test_chunk_capacity_formula(oracle):
    """nc = trunc(ceil(2*(ceil(0.5*cl/r)+1)*(ceil(0.5*ch/r)+1))), src/chunks.jl:32-35."""
    case = H.newton_case(nx=25, ny=25, dyn=pkg.LenJonesCfg(sigma=2, epsilon=4), wall="rigid")  # examples/chunks.jl
    o = H.make_oracle(case)
    r = pkg.particle_radius(case["dyn"])
    cl, ch = case["geom"].length / 22, case["geom"].height / 22
    want = int(math.ceil(2 * (math.ceil(0.5 * cl / r) + 1) * (math.ceil(0.5 * ch / r) + 1)))
    assert o.chunk_capacity() == want == 18  # nc=18 at examples/chunks.jl geometry


# ---------------------------------------------------------------- generators
def test_rectangular_grid_generator():
    """src/init_states.jl:34-57."""
    r, off = 0.56, 0.4
    pos, geom = pkg.rectangular_grid(7, 5, off, r)
    assert pos.shape == (35, 2)
    k = np.arange(7)
    assert np.allclose(pos[:7, 0], r * (1 + off) + k * r * (2 + off), rtol=1e-14)
    assert np.allclose(pos[::7, 1], r * (1 + off) + np.arange(5) * r * (2 + off), rtol=1e-14)
    assert np.isclose(geom.length, 7 * 2 * r + r * off * 8) and np.isclose(geom.height, 5 * 2 * r + r * off * 6)
    # repeated addition, not multiplication (bit pattern of the reference's `current_x = x[end]` loop)
    x = -r
    for i in range(7):
        x = x + r * (2 + off)
        assert pos[i, 0] == x


# ---------------------------------------------------------------- independent numpy restatement of the pair sums
def _numpy_forces(pos, dyn, geom, periodic):
    dr = pos[:, None, :] - pos[None, :, :]
    if periodic:
        size = np.array([geom.length, geom.height])
        dr = dr - (np.abs(dr) > size / 2) * np.copysign(size, dr)
    d = np.sqrt((dr ** 2).sum(-1))
    np.fill_diagonal(d, np.inf)
    if isinstance(dyn, pkg.LenJonesCfg):
        fmod = 4 * dyn.epsilon * (12 * dyn.sigma ** 12 / d ** 13 - 6 * dyn.sigma ** 6 / d ** 7)
        c = fmod / d
    elif isinstance(dyn, pkg.HarmTruncCfg):
        fmod = np.where(d < dyn.dist_eq, -dyn.k_rep * (d / dyn.dist_eq - 1), -dyn.k_atr * (d / dyn.dist_eq - 1))
        with np.errstate(invalid="ignore"):  # inf / inf on the diagonal, masked by the where
            c = np.where(d > dyn.dist_max, 0.0, fmod / d)
    elif isinstance(dyn, pkg.SzaboCfg):
        fm = np.where(d > dyn.r_eq, dyn.k_adh / dyn.r_eq, dyn.k_rep / (dyn.r_max - dyn.r_eq))
        c = np.where(d > dyn.r_max, 0.0, -fm * (d - dyn.r_eq))
    else:
        fmod = -4 * dyn.epsilon * (-12 * dyn.sigma ** 12 / d ** 13 + 6 * dyn.sigma ** 6 / d ** 7)
        c = np.where(d > 2 ** (1 / 6) * dyn.sigma, 0.0, fmod / d)
    c = np.where(np.isfinite(d), c, 0.0)
    return (c[..., None] * dr).sum(1)


@pytest.mark.parametrize("wall", ["periodic", "rigid"])
@pytest.mark.parametrize("dyn", [pkg.LenJonesCfg(sigma=1.0, epsilon=1.0), pkg.HarmTruncCfg(k_rep=10, k_atr=3, dist_eq=1.0, dist_max=1.2)])
def test_allpairs_forces_match_numpy(oracle, wall, dyn):
    case = H.newton_case(nx=12, ny=10, dyn=dyn, wall=wall, chunks=False, jitter=0.15)
    o = H.make_oracle(case)
    o.calc_forces()
    want = _numpy_forces(case["mk"]().pos, dyn, case["geom"], wall == "periodic")
    assert H.rel_err(o.get_forces(), want) < 1e-13


@pytest.mark.parametrize("kind", ["szabo", "rtp"])
def test_self_propelled_forces_match_numpy(oracle, kind):
    case = H.sp_case(kind, nx=10, ny=9, chunks=False, jitter=0.9)
    o = H.make_oracle(case)
    o.calc_forces()
    want = _numpy_forces(case["mk"]().pos, case["dyn"], case["geom"], True)
    assert np.abs(want).max() > 0
    assert H.rel_err(o.get_forces(), want) < 1e-13


# ---------------------------------------------------------------- the reference's own invariants
@pytest.mark.parametrize("wall", ["periodic", "rigid"])
def test_chunks_equal_allpairs_for_cutoff_law(oracle, wall):
    """check_chunks (test/tests_rings/tests_general.jl:16-30) applied to a truncated law: same pair set."""
    dyn = pkg.HarmTruncCfg(k_rep=10, k_atr=1, dist_eq=1.0, dist_max=1.2)
    a = H.make_oracle(H.newton_case(nx=20, ny=20, dyn=dyn, wall=wall, chunks=True))
    b = H.make_oracle(H.newton_case(nx=20, ny=20, dyn=dyn, wall=wall, chunks=False))
    a.step(300)
    b.step(300)
    assert ((a.pos() - b.pos()) ** 2).sum() < 1e-4  # the reference's threshold (test/tests_rings/runtests.jl:5-11)
    assert np.abs(a.pos() - b.pos()).max() < 1e-10


@pytest.mark.parametrize("dynname", ["lj", "harm"])
def test_threaded_equals_sequencial(oracle, dynname):
    """check_threaded (test/tests_rings/tests_general.jl:32-45)."""
    dyn = pkg.LenJonesCfg(sigma=1, epsilon=1) if dynname == "lj" else pkg.HarmTruncCfg(k_rep=10, k_atr=1, dist_eq=1.0, dist_max=1.2)
    a = H.make_oracle(H.newton_case(nx=24, ny=24, dyn=dyn), threads=1)
    b = H.make_oracle(H.newton_case(nx=24, ny=24, dyn=dyn), threads=3)
    a.step(100)
    b.step(100)
    assert np.abs(a.pos() - b.pos()).max() < 1e-11


def test_lj_stale_cells_in_second_pass(oracle):
    """update_verlet! reuses the pre-drift chunks for pass 2 (src/integration.jl:415-431): the pair set is semantics
    for the un-truncated LJ law, so chunks and all-pairs differ at O(1e-3..1e-2) in force."""
    a = H.make_oracle(H.newton_case(nx=16, ny=16, chunks=True))
    b = H.make_oracle(H.newton_case(nx=16, ny=16, chunks=False))
    a.calc_forces()
    b.calc_forces()
    assert H.rel_err(a.get_forces(), b.get_forces()) > 1e-6


# ---------------------------------------------------------------- integrators / walls / quantities
def test_verlet_energy_conservation_c1(oracle):
    """C1: README quick start (README.md:140-174): 10x10 LJ, rigid rectangle, no chunks, dt=0.01."""
    case = H.newton_case(nx=10, ny=10, wall="rigid", chunks=False, dt=0.01, jitter=0.0)
    o = H.make_oracle(case)
    ke0, pe0 = o.energies()
    o.step(100)
    ke1, pe1 = o.energies()
    assert abs((ke1 + pe1) - (ke0 + pe0)) < 2e-3 * abs(ke0 + pe0)
    ns, t = o.time()
    acc = 0.0
    for _ in range(100):
        acc += 0.01  # Float64 accumulation time += dt (src/integration.jl:500-503); note acc != 1.0
    assert ns == 100 and t == acc and t != 1.0


def test_energies_known_answer(oracle):
    dyn = pkg.LenJonesCfg(sigma=1.0, epsilon=2.0)
    geom = pkg.RectangleCfg(length=20.0, height=20.0)
    rmin = 2 ** (1 / 6)
    pts = np.array([[5.0, 5.0], [5.0 + rmin, 5.0]])
    vel = np.array([[3.0, 4.0], [0.0, -1.0]])
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=vel), space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom),
                            dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.01), lower=H.lower)
    ke, pe = o.energies()
    assert ke == (25.0 + 1.0) / 2
    assert np.isclose(pe, -2.0, rtol=1e-13)  # -epsilon at the minimum


def test_rigid_rectangle_flips_velocity_only(oracle):
    """src/integration.jl:271-285: no position fix; the flip repeats every step while the particle overlaps."""
    dyn = pkg.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=1.0, dist_max=1.2)
    geom = pkg.RectangleCfg(length=10.0, height=10.0)
    pts = np.array([[0.3, 5.0], [5.0, 9.9]])
    vel = np.array([[-1.0, 0.5], [0.25, 2.0]])
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=vel), space_cfg=pkg.SpaceCfg(wall_type=pkg.RigidWalls(), geometry_cfg=geom),
                            dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.0), lower=H.lower)
    o.walls()
    assert np.array_equal(o.second(), [[1.0, 0.5], [0.25, -2.0]])
    assert np.array_equal(o.pos(), pts)
    o.walls()
    assert np.array_equal(o.second(), vel)  # flips back


def test_periodic_wrap(oracle):
    case = H.newton_case(nx=8, ny=8, chunks=False)
    g = case["geom"]
    # particles outside are rejected by the constructor, so the wrap is exercised through a step instead
    pts = np.array([[g.length - 1e-4, 1.0], [1.0, 1e-4], [3.0, 3.0]])
    vel = np.array([[1.0, 0.0], [0.0, -1.0], [0.0, 0.0]])
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=vel), space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=g),
                            dynamic_cfg=pkg.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=0.1, dist_max=0.12), int_cfg=pkg.IntCfg(dt=0.001), lower=H.lower)
    o.step(1)
    p = o.pos()
    assert np.isclose(p[0, 0], 9e-4, atol=1e-12) and np.isclose(p[1, 1], g.height - 9e-4, atol=1e-12)


def test_force_walls_circle_and_lines(oracle):
    """calc_walls_forces! (src/integration.jl:228-266) against direct formulas from src/configs.jl:119-163,259-261."""
    dyn = pkg.HarmTruncCfg(k_rep=10, k_atr=1, dist_eq=1.0, dist_max=1.2)
    r = 0.5
    wall_pot = pkg.HarmTruncCfg(k_rep=20, k_atr=0, dist_eq=r, dist_max=r * 1.1)
    geom = pkg.RectangleCfg(length=20.0, height=10.0)
    circle = pkg.CircleCfg(radius=1.5, center=(5.0, 5.0))
    lines = pkg.LinesCfg([[(15.0, 2.5), (15.0, 7.5)]])
    space = pkg.SpaceCfg([(pkg.RigidWalls(), geom), (pkg.PotentialWalls(potential=wall_pot, mode="outside"), circle),
                          (pkg.PotentialWalls(potential=wall_pot), lines)])
    pts = np.array([[5.0 + 1.5 + 0.3, 5.0], [15.2, 5.0], [15.0 + 0.3 * 0.6, 7.5 + 0.3 * 0.8], [2.0, 2.0]])
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=pts, vel=np.zeros_like(pts)), space_cfg=space, dynamic_cfg=dyn,
                            int_cfg=pkg.IntCfg(dt=0.001), lower=H.lower)
    o.clean_forces()
    o.walls_forces()
    f = o.get_forces()
    k = 20.0
    assert np.allclose(f[0], [-k * (0.3 / r - 1), 0.0])              # 0.3 outside the circle, pushed outward
    assert np.allclose(f[1], [-k * (0.2 / r - 1), 0.0])              # 0.2 right of the segment
    assert np.allclose(f[2], -k * (0.3 / r - 1) * np.array([0.6, 0.8]))  # beyond the end point p2: radial from p2
    assert np.all(f[3] == 0.0)
