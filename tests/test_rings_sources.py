"""Sources, sinks and a variable number of rings: src/rings/sources.jl:191-263 (update_area_empty!,
process_sink_source!), src/rings/states.jl:173-227 (add_ring!, remove_ring!, calc_active_ids!), processed at the head of
every Rings step! (src/rings/integration.jl:353-358,523-526).  Fixture: the reference's `:source` system
(test/tests_rings/rings_utils.jl:101-191) with its `check_chunks` / `check_threaded` invariants
(test/tests_rings/tests_general.jl:16-45, threshold 1e-4)."""
import numpy as np
import pytest

import helpers as H

pkg = H.pkg


def _active_cms_x(o, case):
    m, _, _ = o.rings_active()
    p = o.pos().reshape(case["num_rings"], 10, 2)
    return p[m.astype(bool)].mean(1)[:, 0]


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_source_spawns_into_first_free_slots_with_increasing_uids(oracle):
    case = H.rings_source_case()
    o = H.make_oracle_sources(case)
    m, u, na = o.rings_active()
    assert na == 0 and m.sum() == 0 and np.array_equal(u, np.arange(1, 101))
    o.step(1)  # the 2 x 2 spawn areas are empty: four rings appear in slots 0..3 (add_ring!: first free slot)
    m, u, na = o.rings_active()
    assert na == 4 and np.array_equal(np.flatnonzero(m), [0, 1, 2, 3])
    assert list(u[:4]) == [101, 102, 103, 104]   # uid = maximum(uids) + 1, src/rings/states.jl:180
    # spawn polarisation = draw * 2 pi, drawn in spawn order (get_spawn_pol(::RandomPol), src/rings/sources.jl:27)
    assert np.allclose(o.second()[:4] / (2 * np.pi), case["spawn_draws"][:4], atol=0.02)
    # the areas are occupied now: nothing spawns until the rings have left them
    o.step(50)
    assert o.rings_active()[2] == 4
    # spawned ring = spawn_pos shifted into its area: area of a regular 10-gon of the spawn radius
    areas = o.rings_info()[0]
    assert np.allclose(areas[:4], areas[0], rtol=0.05) and areas[0] > 1.0


def test_oracle_sink_removes_rings_and_slots_are_reused(oracle):
    case = H.rings_source_case()
    o = H.make_oracle_sources(case)
    counts, seen_uids = [], set()
    for _ in range(45):
        o.step(100)
        m, u, na = o.rings_active()
        counts.append(na)
        seen_uids |= set(u[m.astype(bool)].tolist())
        assert na == m.sum()
    assert counts[-1] < len(seen_uids)                 # rings were removed by the sink ...
    assert max(counts) <= 30 and counts[-1] >= 15      # ... and the population saturates instead of filling the 100 slots
    geom = case["geom"]
    sink_x = geom.length - 4 * case["ring_d"]
    assert _active_cms_x(o, case).max() < sink_x + 2 * case["ring_d"] + 1.0   # nobody survives beyond the sink


def test_oracle_sources_chunks_equal_allpairs_and_threaded(oracle):
    """check_chunks / check_threaded on the :source fixture (test/tests_rings/tests_general.jl:16-45): ChunksChecker vs
    PosChecker, Sequencial vs Threaded — the same rings appear in the same slots at the same steps."""
    outs = []
    for use_chunks, threads in ((True, 1), (False, 1), (True, 4)):
        case = H.rings_source_case(use_chunks=use_chunks)
        o = H.make_oracle_sources(case, threads=threads)
        o.step(1500)
        outs.append((o.pos(), o.second(), o.rings_active()))
    for pos, pol, (m, u, na) in outs[1:]:
        assert ((pos - outs[0][0]) ** 2).sum() + ((pol - outs[0][1]) ** 2).sum() < 1e-4
        assert np.array_equal(m, outs[0][2][0]) and np.array_equal(u, outs[0][2][1]) and na == outs[0][2][2]


def test_oracle_fixed_spawn_pol_and_full_state(oracle):
    """spawn_pol as a number; when every slot is taken add_ring! finds no space and silently skips (states.jl:185-186)."""
    case = H.rings_source_case(spawn_pol=0.0, num_slots=6)
    case["source_cfg"] = case["source_cfg"][:1]      # no sink
    o = H.make_oracle_sources(case)
    o.step(3000)
    m, u, na = o.rings_active()
    assert na == 6 and m.all()
    assert u.max() == 6 + 6


# ---------------------------------------------------------------- device
@pytest.mark.gpu
@pytest.mark.parametrize("use_chunks", [True, False])
def test_gpu_sources_and_sinks_match_oracle(cuda_lib, use_chunks):
    case = H.rings_source_case(use_chunks=use_chunks)
    g, o = H.make_gpu_rings_sources(case), H.make_oracle_sources(case)
    total = 0
    for steps in (1, 49, 450, 1500, 2000):
        g.step(steps)
        o.step(steps)
        total += steps
        g.sync_to_host()
        mg, ug, ng = g.rings_active()
        mo, uo, no = o.rings_active()
        assert ng == no and np.array_equal(mg, mo) and np.array_equal(ug, uo)
        act = np.repeat(mo.astype(bool), 10)
        assert np.abs(g.state.pos[act] - o.pos()[act]).max() / case["geom"].length < 1e-11
        assert np.abs(g.state.pol[mo.astype(bool)] - o.second()[mo.astype(bool)]).max() < 1e-10
        assert H.rel_err(g.get_forces()[act], o.get_forces()[act]) < 1e-9
        ag, cg, _ = g.rings_info()
        ao, co, _ = o.rings_info()
        assert H.rel_err(ag[mo.astype(bool)], ao[mo.astype(bool)]) < 1e-10
        assert H.rel_err(cg[mo.astype(bool)], co[mo.astype(bool)]) < 1e-11
    assert g.time_info.num_steps == total
    assert no < uo.max() - 100      # the sink removed rings on the way (slots were reused)


@pytest.mark.gpu
def test_gpu_variable_rings_without_sources(cuda_lib):
    """RingsState(active_state=mask) alone: inactive rings take no part in binning, forces or the update."""
    case = H.rings_case("normal", 8, 8)
    from mavi_jl_b200.rings.rings import RingsSystem
    from mavi_jl_b200.rings.states import RingsState
    st0 = case["mk"]()
    mask = np.ones(64, dtype=bool)
    mask[::3] = False

    def mk():
        return RingsState(rings_pos=st0.rings_pos.copy(), pol=st0.pol.copy(), active_state=pkg.ActiveState(mask))

    g = RingsSystem(state=mk(), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"])
    import __graft_entry__ as entry
    from mavi_jl_b200.params import lower
    o = entry.load_oracle().OracleSystem(state=mk(), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"], lower=lower)
    noise = np.random.default_rng(2).standard_normal((200, 64))
    g.step(200, noise)
    o.step(200, noise)
    g.sync_to_host()
    act = np.repeat(mask, 10)
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert np.array_equal(g.state.pos[~act], st0.pos[~act])          # inactive rings never move
    assert H.rel_err(g.get_forces()[act], o.get_forces()[act]) < 1e-10
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    g.update_chunks(); o.update_chunks()
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(cg, co) and np.array_equal(ng, no) and ng.sum() == mask.sum() * 10
