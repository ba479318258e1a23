"""Ring invasions: src/rings/integration.jl:379-520 — every InvasionsCfg.steps_to_update steps the rings
are binned by centre of mass into ring-level chunks and the particles of one ring inside the polygon of a neighbouring ring
are listed (ray casting, point_line_intersect).  The reference has no test of its own for this; the oracle restatement is
pinned by hand-checkable polygons and by its chunks / all-pairs equivalence, the device by the oracle."""
import numpy as np
import pytest

import helpers as H

pkg = H.pkg


def _overlap_case(r_chunks=(4, 4), steps_to_update=5, kind="normal", dt=0.001):
    from mavi_jl_b200.rings import configs as rc
    from mavi_jl_b200.rings.states import RingsState
    n = 8 if kind == "normal" else 6
    case = H.rings_case(kind, n, n)
    st0 = case["mk"]()
    rp = st0.rings_pos.copy()
    rp[1::2] += np.array([0.0, -1.9])   # every second ring (column-outer, row-inner order) pushed into the ring below it
    case["mk"] = lambda: RingsState(rings_pos=rp.copy(), pol=st0.pol.copy(), types=None if st0.types is None else st0.types.copy(),
                                    num_particles=st0.num_particles if st0.types is not None else None)
    ic = case["int_cfg"]
    case["int_cfg"] = rc.RingsIntCfg(dt=dt, p_chunks_cfg=ic.chunks_cfg, r_chunks_cfg=None if r_chunks is None else pkg.ChunksCfg(*r_chunks),
                                     invasions_cfg=rc.InvasionsCfg(steps_to_update=steps_to_update),
                                     device=pkg.CUDADevice(rng_mode="host_noise"))
    return case


def _brute_force(points, nps, n_max):
    """Independent numpy restatement: even-odd rule with the reference's strict comparisons, all ring pairs."""
    out = []
    nr = len(nps)
    for a in range(nr):
        for b in range(nr):
            if a == b:
                continue
            poly = points[b * n_max: b * n_max + nps[b]]
            l1, l2 = poly, np.roll(poly, -1, axis=0)
            for i in range(nps[a]):
                p = points[a * n_max + i]
                dx, dy = l2[:, 0] - l1[:, 0], l2[:, 1] - l1[:, 1]
                ylo, yhi = np.minimum(l1[:, 1], l2[:, 1]), np.maximum(l1[:, 1], l2[:, 1])
                xlo, xhi = np.minimum(l1[:, 0], l2[:, 0]), np.maximum(l1[:, 0], l2[:, 0])
                with np.errstate(divide="ignore", invalid="ignore"):
                    xi = np.where(dx == 0, l1[:, 0], (dy * l1[:, 0] - dx * l1[:, 1] + dx * p[1]) / dy)
                hit = (xi > p[0]) & (ylo < p[1]) & (p[1] < yhi) & ((dx == 0) | ((xlo < xi) & (xi < xhi))) & ((dy != 0) | (dx == 0))
                hit &= ~((dx != 0) & (dy == 0))
                if hit.sum() % 2:
                    out.append((a, b, a * n_max + i))
    return sorted(out)


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_invasions_match_brute_force_and_timing(oracle):
    case = _overlap_case(r_chunks=(4, 4))
    o = H.make_oracle(case)
    o.step(5)
    assert len(o.invasions()) == 0          # num_steps - last_check < steps_to_update until the 6th step! begins
    pts = o.rings_info()[2].copy()          # continuos_pos = ring_points of the step the check runs in ...
    o2 = H.make_oracle(case)
    o2.step(5)
    o.step(1)                               # ... i.e. unwrapped from the positions after 5 steps
    inv = o.invasions()
    assert len(inv) > 50
    # replay: the check of step 6 sees update_continuos_pos! of step 6 (positions after 5 steps)
    from mavi_jl_b200.rings.states import RingsState
    st = case["mk"]()
    p5 = o2.pos()
    size = np.array([case["geom"].length, case["geom"].height])
    cont = p5.reshape(st.num_rings, st.n_max, 2).copy()
    for r in range(st.num_rings):
        for i in range(1, st.n_max):
            d = p5[r * st.n_max + i] - p5[r * st.n_max + i - 1]
            d = d - (np.abs(d) > size / 2) * np.copysign(size, d)
            cont[r, i] = cont[r, i - 1] + d
    want = _brute_force(cont.reshape(-1, 2), [st.n_max] * st.num_rings, st.n_max)
    assert [tuple(t) for t in inv.tolist()] == want
    o.step(3)
    assert np.array_equal(o.invasions(), inv)   # the list stays until the next check (5 steps later)


def test_oracle_invasions_ring_chunks_equal_all_pairs(oracle):
    outs = []
    for r_chunks in ((4, 4), (6, 5), None):
        o = H.make_oracle(_overlap_case(r_chunks=r_chunks))
        o.step(6)
        outs.append(o.invasions().tolist())
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) > 0


# ---------------------------------------------------------------- device
@pytest.mark.gpu
@pytest.mark.parametrize("kind,r_chunks", [("normal", (4, 4)), ("normal", None), ("types", (3, 3))])
def test_gpu_invasions_match_oracle(cuda_lib, kind, r_chunks):
    case = _overlap_case(r_chunks=r_chunks, kind=kind, steps_to_update=7)
    g, o = H.make_gpu_rings(case), H.make_oracle(case)
    nr = case["num_rings"]
    rng = np.random.default_rng(4)
    seen = 0
    for steps in (7, 1, 6, 1, 20):
        noise = rng.standard_normal((steps, nr))
        g.step(steps, noise)
        o.step(steps, noise)
        ig, io = g.invasions(), o.invasions()
        assert np.array_equal(ig, io)
        seen = max(seen, len(ig))
        g.sync_to_host()
        assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-11   # the checks do not disturb the step
    assert seen > 20
