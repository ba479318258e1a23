"""Host-side mirror of `Mavi.Integration` / `Mavi.RunSystem` entry points (src/integration.jl:507-548,
src/run_system.jl:7-23).  Julia's `name!` functions drop the bang here."""
from __future__ import annotations

from .configs import RunTumbleCfg, SzaboCfg
from .systems import RingsSys, System


def newton_step(system: System, host_noise=None):
    """`newton_step!`, src/integration.jl:507-515."""
    system.step(1, host_noise)


def szabo_step(system: System, host_noise=None):
    """`szabo_step!`, src/integration.jl:517-525."""
    system.step(1, host_noise)


def rtp_step(system: System, host_noise=None):
    """`rtp_step!`, src/integration.jl:527-535."""
    system.step(1, host_noise)


def rings_step(system: System, host_noise=None):
    """`Rings.Integration.step!`, src/rings/integration.jl:522-543."""
    system.step(1, host_noise)


def calc_forces(system: System):
    """`clean_forces!; update_chunks!; calc_forces!; calc_walls_forces!` as one device call."""
    system.calc_forces()


def update_chunks(system: System):
    system.update_chunks()


def get_step_function(system: System):
    """src/integration.jl:537-548, src/rings/integration.jl:545."""
    if isinstance(system.type, RingsSys):
        return rings_step
    if isinstance(system.dynamic_cfg, SzaboCfg):
        return szabo_step
    if isinstance(system.dynamic_cfg, RunTumbleCfg):
        return rtp_step
    return newton_step


def run_system(system: System, *, tf=None, num_steps=None, step_func=None, sync=True, sync_every=None, on_sync=None,
               host_noise=None):
    """`run_system`, src/run_system.jl:7-23.  With the default step function the whole run is ONE
    `mavi_step(h, nsteps)` call followed by one download; the step count for `tf` is found
    with the reference's own loop condition `while time < tf` on the Float64 accumulation `time += dt`.

    sync_every / on_sync: the copy-back hook of the experiment / checkpoint drivers (src/experiments.jl:411-488 collect
    `deepcopy(system.state)` every few steps; src/serder.jl:41-63 saves it): the run is cut into batches of `sync_every`
    steps, each ONE mavi_step call followed by ONE download into `system.state` and a call of `on_sync(system)` — the
    state stays device-resident in between, results are identical to an uninterrupted run.
    host_noise: (nsteps, stride) rows for host-noise runs (split along the batches)."""
    if step_func is None and sync_every:
        if tf is not None:
            t, n, dt = system.time_info.time, 0, float(system.int_cfg.dt)
            while t < tf:
                t += dt
                n += 1
        else:
            n = num_steps
        done = 0
        while done < n:
            k = min(int(sync_every), n - done)
            system.step(k, None if host_noise is None else host_noise[done:done + k])
            done += k
            system.sync_to_host()
            if on_sync is not None:
                on_sync(system)
        return
    if step_func is not None:
        if tf is not None:
            while system.time_info.time < tf:
                step_func(system)
        else:
            for _ in range(num_steps):
                step_func(system)
    else:
        if tf is not None:
            t, n, dt = system.time_info.time, 0, float(system.int_cfg.dt)
            while t < tf:
                t += dt
                n += 1
        else:
            n = num_steps
        system.step(n, host_noise)
    if sync:
        system.sync_to_host()
