"""Float32 mode (MaviParams.dtype = MAVI_F32; states whose element type is Float32, src/init_states.jl:34,63 NUM_T):
the mavi_f32 build of libmavi_cuda.so through the C ABI vs the Float64 CPU oracle on the SAME Float32-representable
inputs.

Tolerance (BASELINE.json north_star): forces and positions within 1e-5 relative, norm-wise, over short horizons.
Self-consistency checks that do not involve the oracle are exact: upload/download round trip, force carry == two-pass,
one-rank slab mode == plain run, bitwise reproducibility, and the cell assignment (Float64 arithmetic on the Float32
coordinates, like the reference's Chunks: bit-exact against the oracle and against exact rational arithmetic).
"""
from fractions import Fraction

import numpy as np
import pytest

import helpers as H

pkg = H.pkg
pytestmark = pytest.mark.gpu
TOL32 = 1e-5
F32 = np.float32
DYNS = {"lj": pkg.LenJonesCfg(sigma=1.0, epsilon=1.0),
        "harm": pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2)}


def _cast_state(st, T):
    if isinstance(st, pkg.SecondLawState):
        return pkg.SecondLawState(pos=st.pos.astype(T), vel=st.vel.astype(T), active_state=st.active_state)
    if isinstance(st, pkg.SelfPropelledState):
        return pkg.SelfPropelledState(pos=st.pos.astype(T), pol_angle=st.pol_angle.astype(T), active_state=st.active_state)
    from mavi_jl_b200.rings.states import RingsState
    return RingsState(rings_pos=st.rings_pos.astype(T), pol=st.pol.astype(T), types=st.types, num_particles=st.num_particles
                      if st.types is not None else None)


def _f32_pair(case):
    """(device case with a Float32 state, oracle case with the same values widened to Float64)"""
    mk = case["mk"]
    dev, ora = dict(case), dict(case)
    dev["mk"] = lambda: _cast_state(mk(), F32)
    ora["mk"] = lambda: _cast_state(_cast_state(mk(), F32), np.float64)
    return dev, ora


def _outliers(a, b, tol):
    """fraction of particles whose vector differs by more than tol * ||b||_inf"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b).max(axis=1) > tol * np.abs(b).max()).mean())


def _with_flags(case, flags, **kw):
    c = dict(case)
    ic = case["int_cfg"]
    c["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=pkg.CUDADevice(flags=flags, **kw))
    return c


def test_f32_state_round_trip(cuda_lib):
    dev, _ = _f32_pair(H.newton_case(nx=24, ny=20, jitter=0.3))
    g = H.make_gpu(dev)
    assert g._dtype == F32
    p0, v0 = g.state.pos.copy(), g.state.vel.copy()
    g.state.pos[...] = 0
    g.state.vel[...] = 0
    g.sync_to_host()
    assert g.state.pos.dtype == F32 and np.array_equal(g.state.pos, p0) and np.array_equal(g.state.vel, v0)
    assert g.get_forces().dtype == F32


def test_f32_float64_state_on_float32_device(cuda_lib):
    """CUDADevice(float32=True) with a Float64 host state: converted on upload / download."""
    case = H.newton_case(nx=16, ny=16, jitter=0.2)
    c = dict(case)
    ic = case["int_cfg"]
    c["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=pkg.CUDADevice(float32=True))
    g = H.make_gpu(c)
    assert g._dtype == F32 and g.state.pos.dtype == np.float64
    p0 = g.state.pos.copy()
    g.sync_to_host()
    assert np.array_equal(g.state.pos, p0.astype(F32).astype(np.float64))


@pytest.mark.parametrize("chunks", [True, False])
@pytest.mark.parametrize("wall", ["periodic", "rigid"])
@pytest.mark.parametrize("dyn", ["lj", "harm"])
def test_f32_forces_match_oracle(cuda_lib, dyn, wall, chunks):
    case = H.newton_case(nx=40, ny=36, dyn=DYNS[dyn], wall=wall, chunks=chunks, jitter=0.25)
    dev, ora = _f32_pair(case)
    g, o = H.make_gpu(dev), H.make_oracle(ora)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < TOL32


@pytest.mark.parametrize("dyn,wall", [("lj", "periodic"), ("harm", "rigid"), ("harm", "periodic")])
def test_f32_newton_trajectory(cuda_lib, dyn, wall):
    case = H.newton_case(nx=32, ny=32, dyn=DYNS[dyn], wall=wall, dt=0.001)
    dev, ora = _f32_pair(case)
    g, o = H.make_gpu(dev), H.make_oracle(ora)
    for _ in range(2):
        g.step(25)
        o.step(25)
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL32
        if dyn == "lj":  # smooth law: every element agrees
            assert H.rel_err(g.state.vel, o.second()) < 1e-4
            assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-3
        else:
            # HarmTrunc jumps by k_atr (d_max/d_eq - 1) at the cut-off: a pair that crosses d_max one step earlier or
            # later in Float32 kicks its two particles by ~0.6 dt.  All but a few per cent of the particles
            # must agree (measured: 2.2 % after 50 steps).
            assert _outliers(g.state.vel, o.second(), 1e-4) < 0.06
            assert _outliers(g.get_forces(), o.get_forces(), 1e-3) < 0.06
    ke_g, pe_g = g.energies()
    ke_o, pe_o = o.energies()
    assert abs(ke_g - ke_o) < 1e-4 * abs(ke_o)
    if dyn == "lj":
        assert abs(pe_g - pe_o) < 1e-4 * abs(pe_o)
    assert g.time_info.num_steps == 50 and g.time_info.time == o.time()[1]


@pytest.mark.parametrize("kind", ["szabo", "rtp"])
def test_f32_self_propelled_trajectory(cuda_lib, kind):
    n = 32
    case = H.sp_case(kind, nx=n, ny=n, jitter=0.9 if kind == "szabo" else 0.6)
    dev, ora = _f32_pair(case)
    g, o = H.make_gpu(dev), H.make_oracle(ora)
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < (1e-4 if kind == "szabo" else TOL32)
    rng = np.random.default_rng(11)
    steps = 10
    noise = (rng.standard_normal((steps, n * n)) if kind == "szabo" else rng.random((steps, n * n, 2)) * 0.01).astype(F32)
    g.step(steps, noise)
    o.step(steps, noise.astype(np.float64))
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL32
    assert np.abs(g.state.pol_angle - o.second()).max() < 1e-4


@pytest.mark.parametrize("kind", ["lj_periodic_hot", "harm_rigid", "tight"])
def test_f32_force_carry_bitwise(cuda_lib, kind):
    """carry == two full passes, bit for bit, in the Float32 build as well (same argument as in Float64)."""
    if kind == "lj_periodic_hot":
        case = H.newton_case(nx=40, ny=36, wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    elif kind == "harm_rigid":
        case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="rigid", jitter=0.3, vmax=3.0, dt=0.002)
    else:
        case = H.newton_case(nx=40, ny=40, dyn=DYNS["harm"], wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    dev, _ = _f32_pair(case)
    base = pkg.capi.FLAG_TIGHT_TILES if kind == "tight" else 0
    a = H.make_gpu(_with_flags(dev, base))
    b = H.make_gpu(_with_flags(dev, base | pkg.capi.FLAG_NO_FORCE_CARRY))
    for n in (1, 2, 37, 120):
        a.step(n)
        b.step(n)
        a.sync_to_host()
        b.sync_to_host()
        assert a.state.pos.dtype == F32
        assert np.array_equal(a.state.pos, b.state.pos)
        assert np.array_equal(a.state.vel, b.state.vel)
        assert np.array_equal(a.get_forces(), b.get_forces())
    assert np.array_equal(a.download_cells()[0], b.download_cells()[0])
    if kind == "tight":
        assert a.rebuild_count() > 0


@pytest.mark.parametrize("kind,flags", [("lj", 0), ("harm", 8)])
def test_f32_slab_self_mode_matches_plain(cuda_lib, kind, flags):
    case = H.newton_case(nx=64, ny=40, dyn=DYNS[kind], wall="periodic", jitter=0.3, vmax=3.0, dt=0.002)
    dev, _ = _f32_pair(case)
    a = H.make_gpu(_with_flags(dev, flags))
    b = H.make_gpu(_with_flags(dev, flags | pkg.capi.FLAG_SLAB_SELF))
    n = len(a.state.pos)
    for steps in (1, 33, 80):
        a.step(steps)
        b.step(steps)
        a.sync_to_host()
        ids, pos, second, forces = b.download_local()
        assert pos.dtype == F32 and len(ids) == n and np.array_equal(np.sort(ids), np.arange(n))
        o = np.argsort(ids)
        assert np.array_equal(pos[o], a.state.pos)
        assert np.array_equal(second[o], a.state.vel)
        assert np.array_equal(forces[o], a.get_forces())


@pytest.mark.parametrize("kind,chunks", [("normal", True), ("types", True), ("normal", False)])
def test_f32_rings_trajectory(cuda_lib, kind, chunks):
    n = 13 if kind == "normal" else 5
    case = H.rings_case(kind, n, n, use_chunks=chunks)
    dev, ora = _f32_pair(case)
    g, o = H.make_gpu_rings(dev), H.make_oracle(ora)
    assert g._dtype == F32
    # stiff springs / contacts (k = 20) through the periodic seam: the box size itself is rounded to Float32
    # (ulp(L) ~ 4e-6 at L ~ 40), which bounds the force error by ~ k ulp(L) per pair, a few 1e-4 of the largest force
    assert H.rel_err(g.get_forces(), o.get_forces()) < 5e-4
    noise = np.random.default_rng(3).standard_normal((20, case["num_rings"])).astype(F32)
    g.step(20, noise)
    o.step(20, noise.astype(np.float64))
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < TOL32
    assert np.abs(g.state.pol - o.second()).max() < 1e-4
    areas_g, cms_g, _ = g.rings_info()
    areas_o, cms_o, _ = o.rings_info()
    assert H.rel_err(areas_g, areas_o) < 1e-4 and H.rel_err(cms_g, cms_o) < 1e-4


def test_f32_bitwise_reproducible(cuda_lib):
    outs = []
    for _ in range(2):
        dev, _o = _f32_pair(H.newton_case(nx=48, ny=48, jitter=0.2))
        g = H.make_gpu(dev)
        g.step(40)
        g.sync_to_host()
        outs.append((g.state.pos.copy(), g.state.vel.copy(), g.get_forces()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("wall", ["periodic", "rigid"])
def test_f32_cell_assignment_bit_exact(cuda_lib, wall):
    """A Float32 state is binned with Float64 arithmetic, like the reference: Chunks keeps chunk_length / chunk_height and
    the geometry in Float64 (src/chunks.jl:13,27-30), so div(-pos[2] + bottom_left[2] + space_h, chunk_h) promotes the
    Float32 coordinate.  The cells must therefore equal, bit for bit, those of the Float64 oracle fed with the same
    (Float32-representable) positions, and the trunc of the exact rational quotient."""
    case = H.newton_case(nx=30, ny=26, wall=wall, jitter=0.45)
    dev, ora = _f32_pair(case)
    g, o = H.make_gpu(dev), H.make_oracle(ora)
    cell, counts = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(cell, co) and np.array_equal(counts, no)
    p = g._lowered.params
    cl, ch = p.grid_len / p.num_cols, p.grid_h / p.num_rows
    pos = g.state.pos
    for i in range(0, len(pos), 7):
        x, y = float(pos[i, 0]), float(pos[i, 1])          # Float32 values, widened exactly
        ty = -y + p.grid_bl[1] + p.grid_h                  # Float64, left to right (src/chunks.jl:129)
        tx = x - p.grid_bl[0]
        row = int(Fraction(ty) / Fraction(ch)) + 1         # int() truncates toward zero
        col = int(Fraction(tx) / Fraction(cl)) + 1
        row -= row == p.num_rows + 1
        col -= col == p.num_cols + 1
        assert cell[i] == (col - 1) * p.num_rows + (row - 1), i
    g.step(10)
    o.step(10)
    o.update_chunks()   # device cells are always those of the current positions (the reference's NEXT update_chunks!)
    same = g.download_cells()[0] == o.download_cells()[0]   # trajectories differ at Float32 level: nearly all cells agree
    assert same.mean() > 0.98
