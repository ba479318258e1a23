"""Host hooks of SURVEY.md 8f #4: checkpoint round trip and batched copy-back.

The reference's serder test (test/tests_rings/tests_serder.jl:43-72) runs a system, saves it, loads it and requires the
loaded system to keep evolving as if nothing had happened (state_square_distance below the 1e-4 bar of runtests.jl).
The on-disk format is the host's business (src/serder.jl:41-63 writes `system.state`, the configs and `time_info`); the
device side of the hook is mavi_download_state -> (new handle) mavi_upload_state + mavi_set_time, which is what these
tests exercise: download -> upload -> continue == uninterrupted run.  Experiments (src/experiments.jl:411-488) copy the
state back every few steps: `run_system(sync_every=...)` does one mavi_step + one download per batch.
"""
import numpy as np
import pytest

import helpers as H

pkg = H.pkg
pytestmark = pytest.mark.gpu


def _clone_from_host(case, g, rings=False):
    """A NEW system built from what a checkpoint holds: the downloaded state arrays and TimeInfo."""
    st = g.state
    ti = pkg.TimeInfo(g.time_info.num_steps, g.time_info.time)
    if rings:
        from mavi_jl_b200.rings.rings import RingsSystem
        from mavi_jl_b200.rings.states import RingsState
        new = RingsState(rings_pos=st.rings_pos.copy(), pol=st.pol.copy(), types=None if st.types is None else st.types.copy(),
                         num_particles=st.num_particles if st.types is not None else None)
        return RingsSystem(state=new, space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"], time_info=ti)
    if hasattr(st, "vel"):
        new = pkg.SecondLawState(pos=st.pos.copy(), vel=st.vel.copy())
    else:
        new = pkg.SelfPropelledState(pos=st.pos.copy(), pol_angle=st.pol_angle.copy())
    return pkg.System(state=new, space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=case["int_cfg"], time_info=ti)


@pytest.mark.parametrize("kind", ["lj", "harm_rigid", "szabo", "rtp"])
def test_checkpoint_roundtrip_continues_identically(cuda_lib, kind):
    n1, n2 = 120, 150
    if kind == "lj":
        case = H.newton_case(nx=48, ny=40, wall="periodic", jitter=0.3, vmax=2.0, dt=0.002)
    elif kind == "harm_rigid":
        case = H.newton_case(nx=40, ny=40, dyn=pkg.HarmTruncCfg(k_rep=10.0, k_atr=1.0, dist_eq=1.0, dist_max=1.3), wall="rigid",
                             jitter=0.3, vmax=2.0, dt=0.002)
    else:
        case = H.sp_case(kind, nx=40, ny=32, rot_diff=0.05, jitter=0.6 if kind == "rtp" else 0.9)
    npart = len(case["mk"]().pos)
    rng = np.random.default_rng(8)
    noise = None
    if kind == "szabo":
        noise = rng.standard_normal((n1 + n2, npart))
    elif kind == "rtp":
        noise = rng.random((n1 + n2, 2 * npart))
        noise[:, 0::2] *= 0.02
    a = H.make_gpu(case)                       # uninterrupted
    a.step(n1 + n2, noise)
    a.sync_to_host()
    b = H.make_gpu(case)                       # save at n1, load into a fresh handle, continue
    b.step(n1, None if noise is None else noise[:n1])
    b.sync_to_host()
    c = _clone_from_host(case, b)
    b.close()
    c.step(n2, None if noise is None else noise[n1:])
    c.sync_to_host()
    assert c.time_info.num_steps == a.time_info.num_steps == n1 + n2
    assert c.time_info.time == a.time_info.time          # the same Float64 accumulation time += dt
    # the loaded system re-bins from scratch and primes the force carry again: bit-identical all the same
    assert np.array_equal(c.state.pos, a.state.pos)
    assert np.array_equal(c.state.second, a.state.second)
    assert np.array_equal(c.get_forces(), a.get_forces())


@pytest.mark.parametrize("kind", ["normal", "types"])
def test_rings_checkpoint_roundtrip_reference_bar(cuda_lib, kind):
    """ring_reproducibility_test (test/tests_rings/tests_serder.jl:43-72): run, save, load, run on; the reference's bar is
    state_square_distance < 1e-4.  The loaded RingsSystem re-primes continuos_pos / cms in its constructor."""
    n = 8 if kind == "normal" else 5
    case = H.rings_case(kind, n, n)
    nr = case["num_rings"]
    n1 = n2 = 500
    noise = np.random.default_rng(3).standard_normal((n1 + n2, nr))
    a = H.make_gpu_rings(case)
    a.step(n1 + n2, noise)
    a.sync_to_host()
    b = H.make_gpu_rings(case)
    b.step(n1, noise[:n1])
    b.sync_to_host()
    c = _clone_from_host(case, b, rings=True)
    c.step(n2, noise[n1:])
    c.sync_to_host()
    d2 = ((c.state.pos - a.state.pos) ** 2).sum() + ((c.state.pol - a.state.pol) ** 2).sum()
    assert d2 < 1e-4
    assert np.abs(c.state.pos - a.state.pos).max() < 1e-9
    assert c.time_info.num_steps == n1 + n2


def test_batched_copy_back_equals_uninterrupted(cuda_lib):
    """Experiment-style collection (src/experiments.jl:411-488): state copied back every 16 steps, 7 batches + a short one;
    the device-resident run in between is untouched by the downloads."""
    case = H.newton_case(nx=40, ny=36, wall="periodic", jitter=0.3, vmax=2.0, dt=0.002)
    a, b = H.make_gpu(case), H.make_gpu(case)
    a.step(120)
    a.sync_to_host()
    seen = []
    pkg.run_system(b, num_steps=120, sync_every=16, on_sync=lambda s: seen.append((s.time_info.num_steps, s.state.pos.copy())))
    assert [k for k, _ in seen] == [16, 32, 48, 64, 80, 96, 112, 120]
    assert np.array_equal(b.state.pos, a.state.pos) and np.array_equal(b.state.vel, a.state.vel)
    assert np.array_equal(seen[-1][1], a.state.pos) and not np.array_equal(seen[0][1], seen[1][1])
