"""Host-side mirror of the reference API (mavi.jl_b200/*.py): argument handling and error behaviour that a reference
user relies on, checked without a device (file:line = where the reference does the same)."""
import numpy as np
import pytest

import helpers as H

pkg = H.pkg


def test_states_reinterpret_matrix_inputs_and_promote_integers():
    # src/states.jl:84-101: a (2, N) Matrix is read column-wise; Int inputs promote to Float64 (:89-91)
    m = np.array([[0, 1, 2, 3], [10, 11, 12, 13]])
    st = pkg.SecondLawState(pos=m, vel=np.zeros((2, 4), dtype=int))
    assert st.pos.shape == (4, 2) and st.pos.dtype == np.float64 and st.pos[2].tolist() == [2.0, 12.0]
    st32 = pkg.SecondLawState(pos=m.T.astype(np.float32), vel=np.zeros((4, 2), dtype=np.float32))
    assert st32.pos.dtype == np.float32 and st32.vel.dtype == np.float32            # Float32 mode keeps the element type
    sp = pkg.SelfPropelledState(pos=m, pol_angle=[0, 1, 2, 3])
    assert sp.pol_angle.dtype == np.float64 and sp.second is sp.pol_angle


def test_active_state_mask_and_particle_counts():
    from mavi_jl_b200.states import get_num_total_particles, get_particles_ids
    mask = np.array([True, False, True, True])
    st = pkg.SecondLawState(pos=np.zeros((4, 2)), vel=np.zeros((4, 2)), active_state=pkg.ActiveState(mask))
    assert get_num_total_particles(st) == 3 and get_particles_ids(st).tolist() == [0, 2, 3]      # src/states.jl:129-130
    assert get_num_total_particles(pkg.SecondLawState(pos=np.zeros((4, 2)), vel=np.zeros((4, 2)))) == 4


def test_particle_radius_of_every_dynamics():
    # src/configs.jl:418-421: LJ / RTP sigma 2^(1/6) / 2, Szabo r_eq / 2, HarmTrunc dist_eq / 2
    assert pkg.particle_radius(pkg.LenJonesCfg(sigma=2.0, epsilon=1.0)) == 2.0 * 2 ** (1 / 6) / 2
    assert pkg.particle_radius(pkg.RunTumbleCfg(vo=1.0, sigma=2.0, epsilon=1.0, tumble_rate=1.0)) == 2.0 * 2 ** (1 / 6) / 2
    assert pkg.particle_radius(pkg.HarmTruncCfg(k_rep=1.0, k_atr=1.0, dist_eq=3.0, dist_max=4.0)) == 1.5
    assert pkg.particle_radius(pkg.SzaboCfg(vo=1, mobility=1, relax_time=1, k_rep=1, k_adh=1, r_eq=5.0, r_max=6.0, rot_diff=0)) == 2.5


def test_bounding_box_of_composite_geometries():
    # src/configs.jl:68-93, src/systems.jl:14-28: the chunk grid spans the bounding box of the main geometry
    from mavi_jl_b200.configs import get_bounding_box
    a = pkg.RectangleCfg(length=4, height=2, bottom_left=(1, 1))
    b = pkg.RectangleCfg(length=1, height=5, bottom_left=(-1, 0))
    u = a + b
    assert (u.bottom_left, u.length, u.height) == ((-1.0, 0.0), 6.0, 5.0)
    c = get_bounding_box(pkg.CircleCfg(radius=2.0, center=(1.0, -1.0)))
    assert (c.bottom_left, c.length, c.height) == ((-1.0, -3.0), 4.0, 4.0)


def test_rings_constructor_errors_mirror_the_reference():
    from mavi_jl_b200.rings import configs as rc
    from mavi_jl_b200.rings.rings import RingsSystem
    from mavi_jl_b200.rings.states import RingsState
    inter = rc.HarmTruncCfg(k_rep=20, k_atr=4, dist_eq=1, dist_max=1.2)
    one = dict(p0=3.5, relax_time=1, vo=1.0, mobility=1, rot_diff=0.05, k_area=1, k_spring=20, l_spring=1)
    # src/rings/configs.jl:112-135: vector parameters need a vector num_particles of the same length, and vice versa
    with pytest.raises(ValueError):
        rc.RingsCfg(**{**one, "vo": [1.0, 2.0]}, num_particles=10, interaction_finder=inter)
    with pytest.raises(ValueError):
        rc.RingsCfg(**one, num_particles=[10, 5], interaction_finder=inter)
    with pytest.raises(ValueError):
        rc.RingsCfg(**{**one, "vo": [1.0, 2.0]}, num_particles=[10, 5, 3], interaction_finder=inter)
    # src/rings/states.jl:82-91: per-type num_particles without types
    with pytest.raises(ValueError):
        RingsState(rings_pos=np.zeros((3, 10, 2)), pol=np.zeros(3), num_particles=[10, 5])
    # src/rings/rings.jl:233-262: types in the config but not in the state (checked before any device call)
    case = H.rings_case("types", 3, 3)
    st = case["mk"]()
    st.types = None
    with pytest.raises(ValueError, match="state.types is nothing"):
        RingsSystem(state=st, space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"])
    # sources / sinks need the variable-ring-count state (src/rings/states.jl:173-227): refused loudly otherwise
    with pytest.raises(ValueError, match="active_state"):
        RingsSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"], source_cfg=object())


def test_neighbors_cfg_accepts_julia_style_symbols():
    from mavi_jl_b200.rings.configs import NeighborsCfg
    assert NeighborsCfg(type=":rings").type == "rings" and NeighborsCfg().type == "all" and NeighborsCfg().tol == 1.1
    with pytest.raises(ValueError):
        NeighborsCfg(type="particles")


def test_cpu_device_modes_are_refused_not_emulated():
    # Sequencial / Threaded are the reference's own CPU paths: this backend has no CPU fallback
    case = H.newton_case(nx=4, ny=4)
    with pytest.raises(TypeError, match="no CPU fallback"):
        pkg.System(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"],
                   int_cfg=pkg.IntCfg(dt=0.001, chunks_cfg=None, device=pkg.Threaded()))


def test_lowering_fills_the_flat_parameter_block():
    from mavi_jl_b200 import capi
    from mavi_jl_b200.params import lower
    case = H.newton_case(nx=10, ny=8, wall="periodic", cells=(7, 5))
    p = lower(case["mk"](), case["space"], case["dyn"], case["int_cfg"]).params
    assert p.struct_size > 0 and p.dtype == capi.F64 and p.n == 80 and p.n_spaces == 1
    assert (p.num_cols, p.num_rows) == (7, 5) and p.dynamics == capi.DYN_LJ and p.dt == 0.001
    assert p.spaces[0].wall == capi.WALL_PERIODIC and p.spaces[0].geom == capi.GEOM_RECT
    assert p.grid_len == case["geom"].length and p.grid_h == case["geom"].height
    st32 = case["mk"]()
    st32.pos, st32.vel = st32.pos.astype(np.float32), st32.vel.astype(np.float32)
    assert lower(st32, case["space"], case["dyn"], case["int_cfg"]).params.dtype == capi.F32
    nochunks = H.newton_case(nx=4, ny=4, chunks=False)
    assert lower(nochunks["mk"](), nochunks["space"], nochunks["dyn"], nochunks["int_cfg"]).params.num_cols == 0
