"""Host-side mirror of `Mavi.Systems.System` (reference: src/systems.jl:45-114) bound to a device handle.

`System(...)` keeps the reference's keyword signature.  Construction lowers the configs, creates the device
context (`mavi_create`) and uploads the state (`mavi_upload_state`) — the points where the reference allocates
force buffers, builds `Chunks` and runs the first `update_chunks!`.  Particle state then stays device-resident;
`sync_to_host!` (here `sync_to_host`) is the copy-back hook for GUI / experiment / checkpoint code.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .configs import CUDADevice, SpaceCfg
from .params import lower


class TimeInfo:
    """src/systems.jl:30-33."""

    def __init__(self, num_steps=0, time=0.0):
        self.num_steps = num_steps
        self.time = time


class StandardSys:
    pass


class RingsSys:
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class System:
    def __init__(self, *, state, space_cfg: SpaceCfg, dynamic_cfg, int_cfg, info=None, debug_info=None,
                 time_info=None, sys_type="standard", rng=None, p_neighbors_cfg=None, source_cfg=None, spawn_draws=None):
        self.state = state
        self.space_cfg = space_cfg
        self.dynamic_cfg = dynamic_cfg
        self.int_cfg = int_cfg
        self.info = info
        self.debug_info = debug_info
        self.time_info = time_info or TimeInfo(0, 0.0)
        self.type = StandardSys() if sys_type == "standard" else sys_type
        self.rng = rng
        if not isinstance(int_cfg.device, CUDADevice):
            raise TypeError("this backend implements IntCfg(device=CUDADevice()); Sequencial/Threaded are the "
                            "reference's own CPU paths and are not reimplemented here (no CPU fallback)")
        self._lib = capi.load_library()
        self._lowered = lower(state, space_cfg, dynamic_cfg, int_cfg)
        self._dtype = np.float32 if self._lowered.params.dtype == capi.F32 else np.float64
        self._h = C.c_void_p()
        self._n = len(state.pos)
        self._forces = None
        self._check(self._lib.mavi_create(C.byref(self._lowered.params), C.byref(self._h)))
        self._check(self._lib.mavi_set_time(self._h, self.time_info.num_steps, self.time_info.time))
        self.p_neighbors_cfg = p_neighbors_cfg
        if p_neighbors_cfg is not None:
            # RingsInfo(p_neighbors_cfg=...), src/rings/rings.jl:143-158: before the upload, so that the constructor's
            # first forces! fills the lists like the reference's (src/rings/rings.jl:280-288)
            mode = capi.NEIGH_COUNT if p_neighbors_cfg.only_count else capi.NEIGH_LIST
            self._check(self._lib.mavi_rings_set_neighbors(self._h, mode, int(p_neighbors_cfg.type == "all"),
                                                           float(p_neighbors_cfg.tol)))
        extra = getattr(int_cfg, "extra", None)
        inv = getattr(extra, "invasions_cfg", None)
        if inv is not None:   # RingsIntCfg(invasions_cfg=..., r_chunks_cfg=...), src/rings/configs.jl:343-351
            rc = getattr(extra, "r_chunks_cfg", None)
            self._check(self._lib.mavi_rings_set_invasions(self._h, int(inv.steps_to_update), 0 if rc is None else int(rc.num_cols),
                                                           0 if rc is None else int(rc.num_rows)))
        self.source_cfg = source_cfg
        ring_mask = getattr(state, "ring_mask", None)
        if source_cfg is not None or ring_mask is not None:
            # RingsSystem(source_cfg=...) / RingsState(active_state=...): sources, sinks and the VarRingsIds mask go to the
            # device before the upload (src/rings/rings.jl:162-185; the constructor's update_ids! sees the mask)
            from .rings.sources import lower_sources
            arr, n = lower_sources(source_cfg, self._lowered.keep) if source_cfg is not None else (None, 0)
            draws = None if spawn_draws is None else np.ascontiguousarray(spawn_draws, dtype=np.float64)
            self._check(self._lib.mavi_rings_set_sources(self._h, arr, n, _ptr(ring_mask), _ptr(draws),
                                                         0 if draws is None else len(draws)))
        self._slab = int_cfg.device.world > 1 or bool(int_cfg.device.flags & capi.FLAG_SLAB_SELF)
        if self._slab:
            ids = getattr(state, "ids", None)
            self.local_ids = np.ascontiguousarray(np.arange(self._n) if ids is None else ids, dtype=np.int64)
            self.upload_local()
        else:
            self.upload_state()

    # ------------------------------------------------------------------ plumbing
    def _check(self, status):
        if status != capi.OK:
            buf = C.create_string_buffer(512)
            if self._h:
                self._lib.mavi_last_error(self._h, buf, 512)
            raise capi.MaviError(status, buf.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mavi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state movement
    def upload_state(self):
        """Host state -> device (after construction or after the host edited `system.state`)."""
        st = self.state
        pos = np.ascontiguousarray(st.pos, dtype=self._dtype)
        second = np.ascontiguousarray(st.second, dtype=self._dtype)
        mask = st.active_mask()
        self._check(self._lib.mavi_upload_state(self._h, _ptr(pos), _ptr(second), _ptr(mask), self._n))

    def upload_local(self):
        """Slab mode: upload the particles of this rank's cell columns with their global ids."""
        st = self.state
        pos = np.ascontiguousarray(st.pos, dtype=self._dtype)
        second = np.ascontiguousarray(st.second, dtype=self._dtype)
        self._check(self._lib.mavi_upload_local(self._h, _ptr(self.local_ids), _ptr(pos), _ptr(second), len(pos)))

    def local_count(self):
        n = C.c_int64()
        self._check(self._lib.mavi_local_count(self._h, C.byref(n)))
        return n.value

    def download_local(self, out=None, want_forces=True):
        """Slab mode: (ids, pos, second, forces) of the particles this rank currently owns.  `out` = (ids, pos, second)
        preallocated arrays with room for at least the owned count (e.g. pinned buffers): the download goes straight into
        their prefix and views of it are returned.  want_forces=False skips the force download (forces is None)."""
        n = self.local_count()
        sshape = (n, 2) if self.state.second.ndim == 2 else (n,)
        if out is None:
            ids = np.empty(n, dtype=np.int64)
            pos = np.empty((n, 2), dtype=self._dtype)
            second = np.empty(sshape, dtype=self._dtype)
        else:
            ids, pos, second = out[0][:n], out[1][:n], out[2][:n]
            ok = (ids.dtype == np.int64 and pos.dtype == self._dtype and second.dtype == self._dtype and
                  pos.shape == (n, 2) and second.shape == sshape and len(ids) == n and
                  all(a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"] for a in (ids, pos, second)))
            if not ok:
                raise ValueError("download_local(out=...): need writable C-contiguous (ids int64[>=n], pos T[>=n,2], "
                                 f"second T[>=n{',2' if len(sshape) == 2 else ''}]) with n = {n}")
        forces = np.empty((n, 2), dtype=self._dtype) if want_forces else None
        self._check(self._lib.mavi_download_local(self._h, _ptr(ids), _ptr(pos), _ptr(second), _ptr(forces)))
        return ids, pos, second, forces

    def sync_to_host(self):
        """`sync_to_host!(system)`: device state -> `system.state` arrays, and TimeInfo."""
        st = self.state

        def direct(a):  # download straight into the caller's array when layout and dtype allow (pinned buffers stay pinned)
            return a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"] and a.dtype == self._dtype

        pos = st.pos if direct(st.pos) else np.empty((self._n, 2), dtype=self._dtype)
        second = st.second if direct(st.second) else np.empty(st.second.shape, dtype=self._dtype)
        self._check(self._lib.mavi_download_state(self._h, _ptr(pos), _ptr(second)))
        if pos is not st.pos:
            st.pos[...] = pos
        if second is not st.second:
            st.second[...] = second
        ns, t = C.c_int64(), C.c_double()
        self._check(self._lib.mavi_get_time(self._h, C.byref(ns), C.byref(t)))
        self.time_info.num_steps, self.time_info.time = ns.value, t.value
        return self

    def sync(self):
        self._check(self._lib.mavi_sync(self._h))

    # ------------------------------------------------------------------ hot path
    def step(self, nsteps=1, host_noise=None):
        """nsteps x the step function `get_step_function(system)` selects (src/integration.jl:537-548)."""
        noise = None if host_noise is None else np.ascontiguousarray(host_noise, dtype=self._dtype)
        if noise is not None:
            need = self.noise_stride() * int(nsteps)
            if noise.size < need:
                raise ValueError(f"host_noise has {noise.size} entries, mavi_step reads {need} "
                                 f"({self.noise_stride()} per step x {nsteps} steps)")
        self._check(self._lib.mavi_step(self._h, nsteps, _ptr(noise)))
        # time_info advances on the host exactly like update_time! (time += dt per step, Float64)
        dt = float(self.int_cfg.dt)
        for _ in range(nsteps):
            self.time_info.time += dt
            self.time_info.num_steps += 1

    def noise_stride(self):
        """Host-noise entries mavi_step reads per step (include/mavi.h): Szabo n, RTP 2n, Rings num_rings; in slab mode n
        is the GLOBAL particle count (the kernels index the row with the original id)."""
        p = self._lowered.params
        n = int(p.n_global) if (self._slab and p.n_global > 0) else self._n
        return {capi.DYN_SZABO: n, capi.DYN_RTP: 2 * n,
                capi.DYN_RINGS: getattr(self.state, "num_rings", 0)}.get(p.dynamics, 0)

    def calc_forces(self):
        self._check(self._lib.mavi_calc_forces(self._h))

    def update_chunks(self):
        self._check(self._lib.mavi_bin(self._h))

    def get_forces(self):
        """`get_forces(system)`, src/systems.jl:117 (downloads; (N, 2))."""
        f = np.empty((self._n, 2), dtype=self._dtype)
        self._check(self._lib.mavi_download_forces(self._h, _ptr(f)))
        return f

    # ------------------------------------------------------------------ chunks inspection (parity checks)
    @property
    def num_cells(self):
        p = self._lowered.params
        return p.num_cols * p.num_rows

    def download_cells(self):
        cell = np.empty(self._n, dtype=np.int32)
        counts = np.empty(self.num_cells, dtype=np.int32)
        self._check(self._lib.mavi_download_cells(self._h, _ptr(cell), _ptr(counts)))
        return cell, counts

    def download_cell_lists(self):
        start = np.empty(self.num_cells + 1, dtype=np.int32)
        ids = np.full(self._n, -1, dtype=np.int32)
        self._check(self._lib.mavi_download_cell_lists(self._h, _ptr(start), _ptr(ids)))
        return start, ids[: start[-1]]

    def cell_neighbors(self, cell):
        out = (C.c_int32 * 8)()
        n = C.c_int32()
        self._check(self._lib.mavi_cell_neighbors(self._h, cell, out, C.byref(n)))
        return [out[i] for i in range(n.value)]

    # ------------------------------------------------------------------ quantities / instrumentation
    def energies(self, pe_mode=0, want_ke=True, want_pe=True):
        """(kinetic_energy, potential_energy); a quantity that is not wanted is not computed (NULL pointer) and comes
        back as None — the exact potential energy is O(N^2) like the reference's (src/quantities.jl:46-66)."""
        ke, pe = C.c_double(), C.c_double()
        self._check(self._lib.mavi_energies(self._h, pe_mode, C.byref(ke) if want_ke else None,
                                            C.byref(pe) if want_pe else None))
        return (ke.value if want_ke else None), (pe.value if want_pe else None)

    def launch_count(self):
        n = C.c_int64()
        self._check(self._lib.mavi_launch_count(self._h, C.byref(n)))
        return n.value

    def rebuild_count(self):
        n = C.c_int64()
        self._check(self._lib.mavi_rebuild_count(self._h, C.byref(n)))
        return n.value

    def counters(self):
        """Totals since the last upload (mavi_counters): steps, rebinned particles, inter-tile movers, repaired tiles,
        slab emigrants, overflow rebuilds, tile capacity, tiles."""
        out = (C.c_int64 * 8)()
        self._check(self._lib.mavi_counters(self._h, out))
        keys = ("steps", "rebinned", "tile_movers", "tiles_repaired", "emigrants", "rebuilds", "tile_cap", "tiles")
        return dict(zip(keys, (int(v) for v in out)))

    def set_profiling(self, on=True):
        self._check(self._lib.mavi_set_profiling(self._h, int(on)))

    def last_step_ms(self):
        ms = (C.c_float * 5)()
        self._check(self._lib.mavi_last_step_ms(self._h, ms))
        return list(ms)

    def particle_neighbors(self):
        """(count[n], lists): contact counts and, unless only_count, the neighbour ids of every particle (ascending;
        the reference appends in pair-enumeration order and compares sorted lists)."""
        if self.p_neighbors_cfg is None:
            raise ValueError("the system was built without p_neighbors_cfg")
        count = np.zeros(self._n, dtype=np.int32)
        only = self.p_neighbors_cfg.only_count
        lst = None if only else np.empty((self._n, capi.NEIGH_MAX), dtype=np.int32)
        self._check(self._lib.mavi_rings_download_neighbors(self._h, _ptr(count), _ptr(lst)))
        return count, (None if only else [lst[i, :count[i]].tolist() for i in range(self._n)])

    def invasions(self):
        """`system.info.invasions.list` of the last check (src/rings/integration.jl:509-520) as an (n, 3) int array of
        (invasor ring, invaded ring, scalar particle id), 0-based, sorted."""
        n = C.c_int64()
        self._check(self._lib.mavi_rings_download_invasions(self._h, C.byref(n), None, 0))
        out = np.empty((n.value, 3), dtype=np.int32)
        if n.value:
            self._check(self._lib.mavi_rings_download_invasions(self._h, C.byref(n), _ptr(out), n.value))
        return out

    def rings_active(self):
        """(mask[num_rings], uids[num_rings], num_active): VarRingsIds after the last step (sources / sinks add and remove
        rings on the device); also refreshes `state.ring_mask` / `state.uids`."""
        nr = self.state.num_rings
        mask = np.empty(nr, dtype=np.uint8)
        uids = np.empty(nr, dtype=np.int64)
        na = C.c_int64()
        self._check(self._lib.mavi_rings_download_active(self._h, _ptr(mask), _ptr(uids), C.byref(na)))
        if getattr(self.state, "ring_mask", None) is not None:
            self.state.ring_mask[...] = mask
            self.state.uids[...] = uids
        return mask, uids, na.value

    def rings_info(self):
        nr = self.state.num_rings
        areas = np.empty(nr, dtype=self._dtype)
        cms = np.empty((nr, 2), dtype=self._dtype)
        cont = np.empty((self._n, 2), dtype=self._dtype)
        self._check(self._lib.mavi_rings_download_info(self._h, _ptr(areas), _ptr(cms), _ptr(cont)))
        return areas, cms, cont


def get_forces(system):
    return system.get_forces()


def get_neigh_count(system):
    """`get_neigh_count(system.info.p_neigh)`, src/rings/neighbors.jl:62 (downloads)."""
    return system.particle_neighbors()[0]


def get_neigh_list(system, pid):
    """`get_neigh_list(system.info.p_neigh, id)`, src/rings/neighbors.jl:58-61 (0-based ids, ascending)."""
    return system.particle_neighbors()[1][pid]


def get_num_total_particles(system):
    from .states import get_num_total_particles as g
    return g(system.state)
