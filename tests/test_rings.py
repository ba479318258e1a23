"""Mavi.Rings: CPU known-answer / invariant tests of the oracle, and GPU parity tests against it.

The reference's Rings tests (test/tests_rings/) hold (a) equivalence invariants `check_chunks` / `check_threaded` with
threshold sum |dpos|^2 < 1e-4 after t = 10 (runtests.jl:5-11, tests_general.jl:16-45) — applied here to the oracle and to
the device path — and (b) golden neighbour-id lists that depend on Julia's MersenneTwister streams (runtests.jl:17-34):
not reproducible without Julia, and contact lists are a "next" row.
"""
import math

import numpy as np
import pytest

import helpers as H

pkg = H.pkg


def _noise(case, steps, seed=5):
    return np.random.default_rng(seed).standard_normal((steps, case["num_rings"]))


# ---------------------------------------------------------------- CPU: oracle
def test_regular_polygon_area_and_ring_radius(oracle):
    """create_circle + get_ring_radius (src/rings/init_states.jl:10-25, src/rings/utils.jl:5-7): a regular n-gon of
    circumradius R has area n/2 R^2 sin(2 pi/n) and side 2*p_radius."""
    case = H.rings_case("normal", 3, 3, rot_diff=0.0)
    o = H.make_oracle(case)
    areas, cms, cont = o.rings_info()
    from mavi_jl_b200.rings.configs import get_ring_radius
    n, pr = 10, 0.5
    R = get_ring_radius(pr, n)
    assert np.allclose(areas, n / 2 * R * R * math.sin(2 * math.pi / n), rtol=1e-13)
    st = case["mk"]()
    side = np.linalg.norm(st.rings_pos[0, 1] - st.rings_pos[0, 0])
    assert np.isclose(side, 2 * pr, rtol=1e-13)
    assert np.allclose(cms, st.rings_pos.mean(1), atol=1e-12)
    assert np.allclose(cont, st.pos)  # nothing wrapped at t = 0


def test_isolated_ring_feels_only_the_area_force(oracle):
    """An isolated regular ring with side == l_spring: springs are at rest, no pair forces, and area_forces!
    (src/rings/integration.jl:140-195) gives F_i = -k_A (A - A0) * (dr.y, -dr.x)/2 with dr = r_{i+1} - r_{i-1},
    i.e. modulus k_A |A - A0| * chord/2 along the outward radius when A < A0."""
    case = H.rings_case("normal", 1, 1, use_chunks=False, rot_diff=0.0, pad=3.0)
    o = H.make_oracle(case)
    f = o.get_forces()
    r0 = case["mk"]().rings_pos[0]
    area = o.rings_info()[0][0]
    A0 = (10 * 1.0 / 3.5) ** 2
    dr = np.roll(r0, -1, 0) - np.roll(r0, 1, 0)
    want = -1.0 * (area - A0) * np.stack([dr[:, 1], -dr[:, 0]], 1) / 2
    assert np.abs(want).max() > 0.1
    assert np.allclose(f, want, rtol=1e-10, atol=1e-12)
    radial = (r0 - r0.mean(0)) / np.linalg.norm(r0 - r0.mean(0), axis=1)[:, None]
    assert np.allclose(np.sign(A0 - area) * f / np.linalg.norm(f, axis=1)[:, None], radial, atol=1e-9)


@pytest.mark.parametrize("kind", ["normal", "types"])
def test_rings_chunks_equal_allpairs(oracle, kind):
    """check_chunks (test/tests_rings/tests_general.jl:16-30): same seed, chunks vs all pairs, threshold 1e-4."""
    n = 8 if kind == "normal" else 5
    ca, cb = H.rings_case(kind, n, n, use_chunks=True), H.rings_case(kind, n, n, use_chunks=False)
    a, b = H.make_oracle(ca), H.make_oracle(cb)
    noise = _noise(ca, 300)
    a.step(300, noise)
    b.step(300, noise)
    assert ((a.pos() - b.pos()) ** 2).sum() + ((a.second() - b.second()) ** 2).sum() < 1e-4
    assert np.abs(a.pos() - b.pos()).max() < 1e-9


def test_rings_threaded_equals_sequencial(oracle):
    """check_threaded (test/tests_rings/tests_general.jl:32-45)."""
    ca = H.rings_case("normal", 8, 8)
    a, b = H.make_oracle(ca, threads=1), H.make_oracle(ca, threads=3)
    noise = _noise(ca, 200)
    a.step(200, noise)
    b.step(200, noise)
    assert ((a.pos() - b.pos()) ** 2).sum() < 1e-4
    assert np.abs(a.pos() - b.pos()).max() < 1e-9


def test_rings_cms_lag_one_step(oracle):
    """update_cms! runs before the unwrap of the current step (src/rings/integration.jl:523-529): after a step, cms is
    the centre of the PREVIOUS step's continuous positions."""
    case = H.rings_case("normal", 4, 4, rot_diff=0.0)
    o = H.make_oracle(case)
    cont0 = o.rings_info()[2].reshape(16, 10, 2)
    o.step(1)
    _, cms1, cont1 = o.rings_info()
    assert np.allclose(cms1, cont0.mean(1), atol=1e-13)
    o.step(1)
    assert np.allclose(o.rings_info()[1], cont1.reshape(16, 10, 2).mean(1), atol=1e-13)


# ---------------------------------------------------------------- GPU parity
@pytest.mark.gpu
@pytest.mark.parametrize("kind,chunks", [("normal", True), ("normal", False), ("types", True), ("types", False)])
def test_rings_gpu_construction_state(cuda_lib, kind, chunks):
    """RingsSystem(...) primes continuos_pos, cms, chunks and runs forces! once (src/rings/rings.jl:280-288)."""
    n = 13 if kind == "normal" else 5
    case = H.rings_case(kind, n, n, use_chunks=chunks)
    g, o = H.make_gpu_rings(case), H.make_oracle(case)
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12
    for a, b in zip(g.rings_info(), o.rings_info()):
        assert H.rel_err(a, b) < 1e-13
    if chunks:
        cg, ng = g.download_cells()
        co, no = o.download_cells()
        assert np.array_equal(cg, co) and np.array_equal(ng, no)
        sg, ig = g.download_cell_lists()
        so, io = o.download_cell_lists()
        assert np.array_equal(sg, so) and np.array_equal(ig, io)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,chunks", [("normal", True), ("normal", False), ("types", True), ("types", False)])
def test_rings_gpu_trajectory(cuda_lib, kind, chunks):
    n = 13 if kind == "normal" else 5
    case = H.rings_case(kind, n, n, use_chunks=chunks)
    g, o = H.make_gpu_rings(case), H.make_oracle(case)
    for block in range(4):
        noise = _noise(case, 25, seed=block)
        g.step(25, noise)
        o.step(25, noise)
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
        assert np.abs(g.state.pol - o.second()).max() < 1e-11
        assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-10
        for a, b in zip(g.rings_info(), o.rings_info()):
            assert H.rel_err(a, b) < 1e-11
    assert g.time_info.num_steps == 100


@pytest.mark.gpu
def test_rings_gpu_reference_invariant_t10(cuda_lib):
    """The reference's own acceptance bar: run to t = 10 (1000 steps, dt = 0.01) and require sum |dpos|^2 < 1e-4
    between two code paths (test/tests_rings/runtests.jl:5-11) — here device (chunks) vs oracle (all pairs)."""
    ca, cb = H.rings_case("normal", 13, 13, use_chunks=True), H.rings_case("normal", 13, 13, use_chunks=False)
    g, o = H.make_gpu_rings(ca), H.make_oracle(cb)
    noise = _noise(ca, 1000)
    g.step(1000, noise)
    o.step(1000, noise)
    g.sync_to_host()
    assert ((g.state.pos - o.pos()) ** 2).sum() + ((g.state.pol - o.second()) ** 2).sum() < 1e-4


@pytest.mark.gpu
def test_rings_gpu_philox_runs(cuda_lib):
    from mavi_jl_b200.rings import configs as rc
    case = H.rings_case("normal", 13, 13)
    case["int_cfg"] = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=case["int_cfg"].chunks_cfg, device=pkg.CUDADevice(rng_mode="philox", seed=3))
    g = H.make_gpu_rings(case)
    pol0 = g.state.pol.copy()
    g.step(50)
    g.sync_to_host()
    assert np.isfinite(g.state.pos).all() and np.abs(g.state.pol - pol0).max() > 1e-3


@pytest.mark.gpu
def test_rings_gpu_philox_statistics(cuda_lib):
    """Device RNG mode of the Rings update! (src/rings/integration.jl:345-346): with the alignment term switched off
    (relax_time -> inf) a ring's polarisation is a random walk with variance 2 D_r t; rings are independent (no correlation
    between neighbouring ring ids) and two seeds give different streams."""
    from mavi_jl_b200.rings import configs as rc

    def run(seed):
        case = H.rings_case("normal", 40, 30, rot_diff=0.5)
        case["dyn"].relax_time = np.full(case["dyn"].num_types, 1e30)
        case["int_cfg"] = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=case["int_cfg"].chunks_cfg, device=pkg.CUDADevice(rng_mode="philox", seed=seed))
        g = H.make_gpu_rings(case)
        pol0 = g.state.pol.copy()
        g.step(100)
        g.sync_to_host()
        assert np.isfinite(g.state.pos).all()
        return g.state.pol - pol0

    d = run(5)
    n = len(d)                                            # 1200 rings, t = 1: variance 2 * 0.5 * 1 = 1
    assert abs(d.var() - 1.0) < 5 * np.sqrt(2.0 / n)      # variance of a normal sample: sigma^2 sqrt(2 / n)
    assert abs(d.mean()) < 5 / np.sqrt(n)
    assert abs(np.corrcoef(d[:-1], d[1:])[0, 1]) < 5 / np.sqrt(n)
    kurt = ((d - d.mean()) ** 4).mean() / d.var() ** 2     # a sum of 100 normal increments is normal: kurtosis 3
    assert abs(kurt - 3.0) < 0.8
    d2 = run(6)
    assert abs(np.corrcoef(d, d2)[0, 1]) < 5 / np.sqrt(n)


@pytest.mark.gpu
def test_rings_asymmetric_matrix_rejected(cuda_lib):
    from mavi_jl_b200.rings import configs as rc
    case = H.rings_case("types", 5, 5)
    m = case["dyn"].interaction_finder.matrix
    m[0][1] = rc.HarmTruncCfg(k_rep=1, k_atr=1, dist_eq=0.9, dist_max=1.0)
    with pytest.raises(pkg.MaviError) as e:
        H.make_gpu_rings(case)
    assert e.value.status == pkg.capi.ERR_UNSUPPORTED


# ---------------------------------------------------------------- particle contact lists (src/rings/neighbors.jl)
def _make_oracle_neigh(case, cfg, threads=1):
    import __graft_entry__ as entry
    from mavi_jl_b200.params import lower
    return entry.load_oracle().OracleSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"],
                                            int_cfg=case["int_cfg"], lower=lower, threads=threads, p_neighbors_cfg=cfg)


def _brute_force_contacts(case, pos, cfg):
    """Independent numpy restatement of neigh_update!(::ParticleNeighbors, ...) (src/rings/neighbors.jl:125-134) over ALL
    pairs: dist < 2 particle_radius(interaction(ti, tj)) * tol and (type == :all or other ring)."""
    st = case["mk"]()
    n_max, n = st.n_max, len(st.pos)
    ring = np.arange(n) // n_max
    active = np.array([(i % n_max) < st.ring_num_particles(i // n_max) for i in range(n)])
    ty = np.zeros(st.num_rings, dtype=int) if st.types is None else np.asarray(st.types) - 1
    dyn = case["dyn"]
    deq = np.array([[dyn.interaction(a, b).dist_eq for b in range(dyn.num_types)] for a in range(dyn.num_types)])
    size = np.array([case["geom"].length, case["geom"].height])
    out = []
    for i in range(n):
        if not active[i]:
            out.append([])
            continue
        dr = pos[i] - pos
        dr = dr - (np.abs(dr) > size / 2) * np.copysign(size, dr)
        dist = np.sqrt((dr ** 2).sum(-1))
        ok = active & (dist < 2 * (deq[ty[ring[i]], ty[ring]] / 2) * cfg.tol) & (np.arange(n) != i)
        if cfg.type == "rings":
            ok &= ring != ring[i]
        out.append(np.flatnonzero(ok).tolist())
    return out


@pytest.mark.parametrize("ntype", ["all", "rings"])
@pytest.mark.parametrize("kind", ["normal", "types"])
def test_oracle_contact_lists_match_brute_force(oracle, kind, ntype):
    from mavi_jl_b200.rings.configs import NeighborsCfg
    n = 6 if kind == "normal" else 5
    case = H.rings_case(kind, n, n)
    cfg = NeighborsCfg(only_count=False, type=ntype, tol=1.1)
    o = _make_oracle_neigh(case, cfg)
    o.step(60, _noise(case, 60))           # rings collide: contacts between different rings appear
    # the lists belong to the forces! of the LAST step, i.e. to the positions before its update!: replay one step less
    o2 = _make_oracle_neigh(case, cfg)
    o2.step(59, _noise(case, 60)[:59])
    want = _brute_force_contacts(case, o2.pos(), cfg)
    count, lists = o.particle_neighbors()
    assert [sorted(l) for l in lists] == want
    assert count.tolist() == [len(w) for w in want]
    if ntype == "all":   # bonded ring neighbours sit at distance l_spring <= dist_eq < dist_eq * tol
        assert min(len(w) for w, a in zip(want, count) if a or w) >= 2
    else:
        assert sum(len(w) for w in want) > 0
    # only_count = true: the same counts, no lists (src/rings/neighbors.jl:102-107)
    oc = _make_oracle_neigh(case, NeighborsCfg(only_count=True, type=ntype, tol=1.1))
    oc.step(60, _noise(case, 60))
    c2, l2 = oc.particle_neighbors()
    assert l2 is None and np.array_equal(c2, count)


def test_oracle_contact_lists_chunks_threads_allpairs_agree(oracle):
    """The reference's own check_particle_neighbors_all loop (test/tests_rings/tests_general.jl:147-180): Sequencial /
    Threaded x chunks / no chunks must give the same sorted lists."""
    from mavi_jl_b200.rings.configs import NeighborsCfg
    cfg = NeighborsCfg(only_count=False, type="rings", tol=1.1)
    outs = []
    for use_chunks, threads in ((True, 1), (False, 1), (True, 4)):
        case = H.rings_case("normal", 5, 5, use_chunks=use_chunks)
        o = _make_oracle_neigh(case, cfg, threads=threads)
        o.step(100, _noise(case, 100))
        count, lists = o.particle_neighbors()
        outs.append((count.tolist(), [sorted(l) for l in lists]))
    assert outs[0] == outs[1] == outs[2]


@pytest.mark.gpu
@pytest.mark.parametrize("ntype", ["all", "rings"])
@pytest.mark.parametrize("kind,chunks", [("normal", True), ("normal", False), ("types", True)])
def test_rings_gpu_contact_lists(cuda_lib, kind, chunks, ntype):
    from mavi_jl_b200.rings.configs import NeighborsCfg
    from mavi_jl_b200.rings.rings import RingsSystem
    n = 8 if kind == "normal" else 5
    case = H.rings_case(kind, n, n, use_chunks=chunks)
    for only_count in (False, True):
        cfg = NeighborsCfg(only_count=only_count, type=ntype, tol=1.1)
        g = RingsSystem(state=case["mk"](), space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"],
                        p_neighbors_cfg=cfg)
        o = _make_oracle_neigh(case, cfg)
        for steps in (0, 1, 60):   # 0: the constructor's forces! already filled them (src/rings/rings.jl:280-288)
            if steps:
                nz = _noise(case, steps, seed=steps)
                g.step(steps, nz)
                o.step(steps, nz)
            cg, lg = g.particle_neighbors()
            co, lo = o.particle_neighbors()
            assert np.array_equal(cg, co)
            if not only_count:
                assert lg == [sorted(l) for l in lo]
        assert cg.sum() > 0


# ---------------------------------------------------------------- PotentialVector: one wall potential per ring type
def _potential_vector_case(potential):
    """create_system_types rings (two ring types) in a periodic box with a PotentialWalls circle (mode outside) in the middle
    and a PotentialWalls line: `potential` is a single HarmTruncCfg or a PotentialVector with one entry per ring type
    (src/configs.jl:454-463; get_particle_type = the ring type, src/rings/states.jl:148)."""
    case = H.rings_case("types", 8, 8)
    geom = case["geom"]
    bl = np.asarray(geom.bottom_left)
    circle = pkg.CircleCfg(radius=1.3, center=[bl[0] + geom.length / 2, bl[1] + geom.height / 2])
    line = pkg.LinesCfg([[(bl[0] + geom.length / 4, bl[1] + geom.height / 4), (bl[0] + geom.length / 4, bl[1] + 3 * geom.height / 4)]])
    case["space"] = pkg.SpaceCfg([(pkg.PeriodicWalls(), geom), (pkg.PotentialWalls(potential=potential, mode="outside"), circle),
                                  (pkg.PotentialWalls(potential=potential), line)])
    return case


def _wall_pots():
    from mavi_jl_b200.rings import configs as rc
    return (rc.HarmTruncCfg(k_rep=30, k_atr=0, dist_eq=0.6, dist_max=0.7), rc.HarmTruncCfg(k_rep=11, k_atr=2, dist_eq=0.9, dist_max=1.2))


def test_oracle_potential_vector_selects_by_ring_type(oracle):
    """The forces of a PotentialVector([A, B]) run are those of a PotentialWalls(A) run for the particles of type-1 rings and
    those of a PotentialWalls(B) run for the particles of type-2 rings (same state; calc_forces! only)."""
    pa, pb = _wall_pots()
    outs = []
    for pot in (pkg.PotentialVector([pa, pb]), pa, pb):
        o = H.make_oracle(_potential_vector_case(pot))
        o.calc_forces()
        outs.append(o.get_forces().copy())
    case = _potential_vector_case(pa)
    st = case["mk"]()
    n_max = st.rings_pos.shape[1]
    ptype = np.repeat(np.asarray(st.types), n_max)          # type of every particle slot (1-based), ring-ordered
    fv, fa, fb = outs
    assert np.array_equal(fv[ptype == 1], fa[ptype == 1]) and np.array_equal(fv[ptype == 2], fb[ptype == 2])
    assert not np.array_equal(fa, fb)                        # the two potentials do act differently on this state


def test_potential_vector_rejected_without_ring_types(oracle):
    """get_particle_type exists for RingsState only: a PotentialVector on a particle system is a constructor error; so is a
    vector whose length is not the number of ring types."""
    pa, pb = _wall_pots()
    case = H.newton_case(nx=8, ny=8)
    geom = case["geom"]
    circle = pkg.CircleCfg(radius=1.0, center=[geom.length / 2, geom.height / 2])
    case["space"] = pkg.SpaceCfg([(pkg.PeriodicWalls(), geom), (pkg.PotentialWalls(potential=pkg.PotentialVector([pa, pb]), mode="outside"), circle)])
    with pytest.raises(TypeError, match="RingsState with types"):
        H.make_oracle(case)
    with pytest.raises(ValueError, match="3 entries for 2 ring types"):
        H.make_oracle(_potential_vector_case(pkg.PotentialVector([pa, pb, pa])))


@pytest.mark.gpu
def test_gpu_potential_vector_matches_oracle(cuda_lib):
    pa, pb = _wall_pots()
    case = _potential_vector_case(pkg.PotentialVector([pa, pb]))
    g, o = H.make_gpu_rings(case), H.make_oracle(case)
    nr = case["num_rings"]
    g.calc_forces()
    o.calc_forces()
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-12
    rng = np.random.default_rng(12)
    for _ in range(2):
        noise = rng.standard_normal((40, nr))
        g.step(40, noise)
        o.step(40, noise)
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-11
        assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-9
