"""ctypes binding of the CPU oracle (oracle/libmavi_oracle.so).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: see oracle/mavi_oracle.h.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmavi_oracle.so")
_lib = None
_fast_lib = None


def _cpu_signature():
    """Short hash of this host's CPU model + ISA flags: the -march=native build below is only valid on the CPU it was
    compiled on (the GPU box differs from the build container), so the file name carries it."""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
    except OSError:
        keep = []
    return hashlib.md5("\n".join(keep).encode()).hexdigest()[:10]


def fast_lib_path():
    return os.path.join(_HERE, f"libmavi_oracle_fast_{_cpu_signature()}.so")


def build_fast(verbose=False):
    """The TIMING build of the same source (`-O3 -march=native -fopenmp`), a separate target from the
    parity oracle (`-O2 -ffp-contract=off`): used by bench.py's cpu_baseline / --impl reference legs only."""
    path = fast_lib_path()  # make rebuilds it when mavi_oracle.c is newer (a stale build lacks the newer entry points)
    res = subprocess.run(["make", "-C", _HERE, "fast", f"FAST_SO={os.path.basename(path)}"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-3000:], res.stderr[-3000:])
    if res.returncode != 0:
        raise RuntimeError("building the -O3 -march=native oracle failed")
    return path


def build(verbose=False):
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-3000:], res.stderr[-3000:])
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed")
    return LIB_PATH


def load(fast=False):
    global _lib, _fast_lib
    if fast:
        if _fast_lib is not None:
            return _fast_lib
        lib = C.CDLL(build_fast())
    else:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
    H, vp, i32, i64, dbl = C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "mor_create": (i32, [vp, C.POINTER(H)]), "mor_destroy": (None, [H]), "mor_set_threads": (None, [H, i32]),
        "mor_last_error": (C.c_char_p, [H]),
        "mor_upload_state": (i32, [H, vp, vp, vp, i64]), "mor_download_state": (i32, [H, vp, vp]),
        "mor_download_forces": (i32, [H, vp]), "mor_step": (i32, [H, i64, vp]), "mor_calc_forces": (i32, [H]),
        "mor_bin": (i32, [H]), "mor_download_cells": (i32, [H, vp, vp]), "mor_download_cell_lists": (i32, [H, vp, vp]),
        "mor_cell_neighbors": (i32, [H, i32, C.POINTER(i32), C.POINTER(i32)]), "mor_chunk_capacity": (i64, [H]),
        "mor_energies": (i32, [H, i32, C.POINTER(dbl), C.POINTER(dbl)]),
        "mor_rings_download_info": (i32, [H, vp, vp, vp]),
        "mor_rings_set_neighbors": (i32, [H, i32, i32, dbl]), "mor_rings_download_neighbors": (i32, [H, vp, vp]),
        "mor_rings_set_sources": (i32, [H, vp, i32, vp, vp, i64]), "mor_rings_download_active": (i32, [H, vp, vp, C.POINTER(i64)]),
        "mor_rings_set_invasions": (i32, [H, i32, i32, i32]), "mor_rings_download_invasions": (i32, [H, C.POINTER(i64), vp, i64]),
        "mor_get_time": (i32, [H, C.POINTER(i64), C.POINTER(dbl)]),
        "mor_clean_forces": (None, [H]), "mor_update_chunks": (i32, [H]), "mor_pair_forces": (None, [H]),
        "mor_walls_forces": (None, [H]), "mor_walls": (None, [H]), "mor_update_verlet": (None, [H]),
        "mor_julia_div": (dbl, [dbl, dbl]), "mor_calc_diff": (None, [H, vp, vp, vp]),
        "mor_potential_force": (None, [i32, vp, vp, dbl, vp]),
        "mor_szabo_interaction": (None, [vp, vp, vp]), "mor_rtp_interaction": (None, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if fast:
        _fast_lib = lib
    else:
        _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleError(RuntimeError):
    def __init__(self, status, msg):
        self.status = status
        super().__init__(f"oracle status {status}: {msg}")


class OracleSystem:
    """The oracle behind the same surface as the host mirror's `System` (built from the same configs)."""

    def __init__(self, *, state, space_cfg, dynamic_cfg, int_cfg, lower, threads=1, p_neighbors_cfg=None, fast=False,
                 source_cfg=None, spawn_draws=None):
        self.lib = load(fast=fast)  # fast: the -O3 -march=native timing build (never used for parity checks)
        self.state = state
        self._lowered = lower(state, space_cfg, dynamic_cfg, int_cfg)
        self.n = len(state.pos)
        self.h = C.c_void_p()
        st = self.lib.mor_create(C.byref(self._lowered.params), C.byref(self.h))
        self._check(st)
        if threads > 1:
            self.lib.mor_set_threads(self.h, threads)
        self.p_neighbors_cfg = p_neighbors_cfg
        if p_neighbors_cfg is not None:  # RingsSystem(p_neighbors_cfg=...), src/rings/rings.jl:143-158
            self._check(self.lib.mor_rings_set_neighbors(self.h, 1 if p_neighbors_cfg.only_count else 2,
                                                         int(p_neighbors_cfg.type == "all"), float(p_neighbors_cfg.tol)))
        extra = getattr(int_cfg, "extra", None)
        inv = getattr(extra, "invasions_cfg", None)
        if inv is not None:
            rc = getattr(extra, "r_chunks_cfg", None)
            self._check(self.lib.mor_rings_set_invasions(self.h, int(inv.steps_to_update), 0 if rc is None else int(rc.num_cols),
                                                         0 if rc is None else int(rc.num_rows)))
        ring_mask = getattr(state, "ring_mask", None)
        if source_cfg is not None or ring_mask is not None:  # RingsSystem(source_cfg=...), RingsState(active_state=...)
            import __graft_entry__ as entry
            entry.load_package()
            from mavi_jl_b200.rings.sources import lower_sources
            arr, n = lower_sources(source_cfg, self._lowered.keep) if source_cfg is not None else (None, 0)
            draws = None if spawn_draws is None else np.ascontiguousarray(spawn_draws, dtype=np.float64)
            self._check(self.lib.mor_rings_set_sources(self.h, arr, n, _ptr(ring_mask), _ptr(draws),
                                                       0 if draws is None else len(draws)))
        pos = np.ascontiguousarray(state.pos, dtype=np.float64)
        second = np.ascontiguousarray(state.second, dtype=np.float64)
        mask = state.active_mask()
        self._check(self.lib.mor_upload_state(self.h, _ptr(pos), _ptr(second), _ptr(mask), self.n))

    def _check(self, st):
        if st != 0:
            raise OracleError(st, self.lib.mor_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.mor_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_cells(self):
        p = self._lowered.params
        return p.num_cols * p.num_rows

    def step(self, nsteps=1, host_noise=None):
        noise = None if host_noise is None else np.ascontiguousarray(host_noise, dtype=np.float64)
        self._check(self.lib.mor_step(self.h, nsteps, _ptr(noise)))

    def calc_forces(self):
        self._check(self.lib.mor_calc_forces(self.h))

    def update_chunks(self):
        self._check(self.lib.mor_bin(self.h))

    def pos(self):
        p = np.empty((self.n, 2))
        self._check(self.lib.mor_download_state(self.h, _ptr(p), None))
        return p

    def second(self):
        s = np.empty(self.state.second.shape)
        self._check(self.lib.mor_download_state(self.h, None, _ptr(s)))
        return s

    def get_forces(self):
        f = np.empty((self.n, 2))
        self._check(self.lib.mor_download_forces(self.h, _ptr(f)))
        return f

    def download_cells(self):
        cell = np.empty(self.n, dtype=np.int32)
        counts = np.empty(self.num_cells, dtype=np.int32)
        self._check(self.lib.mor_download_cells(self.h, _ptr(cell), _ptr(counts)))
        return cell, counts

    def download_cell_lists(self):
        start = np.empty(self.num_cells + 1, dtype=np.int32)
        ids = np.full(self.n, -1, dtype=np.int32)
        self._check(self.lib.mor_download_cell_lists(self.h, _ptr(start), _ptr(ids)))
        return start, ids[: start[-1]]

    def cell_neighbors(self, cell):
        out = (C.c_int32 * 4)()
        n = C.c_int32()
        self._check(self.lib.mor_cell_neighbors(self.h, cell, out, C.byref(n)))
        return [out[i] for i in range(n.value)]

    def chunk_capacity(self):
        return self.lib.mor_chunk_capacity(self.h)

    def energies(self, pe_mode=0):
        ke, pe = C.c_double(), C.c_double()
        self._check(self.lib.mor_energies(self.h, pe_mode, C.byref(ke), C.byref(pe)))
        return ke.value, pe.value

    def rings_info(self):
        nr = self.state.num_rings
        areas, cms, cont = np.empty(nr), np.empty((nr, 2)), np.empty((self.n, 2))
        self._check(self.lib.mor_rings_download_info(self.h, _ptr(areas), _ptr(cms), _ptr(cont)))
        return areas, cms, cont

    def invasions(self):
        """info.invasions.list of the last check as (n, 3) rows (invasor, invaded, particle id), sorted."""
        n = C.c_int64()
        self._check(self.lib.mor_rings_download_invasions(self.h, C.byref(n), None, 0))
        out = np.empty((n.value, 3), dtype=np.int32)
        if n.value:
            self._check(self.lib.mor_rings_download_invasions(self.h, C.byref(n), _ptr(out), n.value))
        return out[np.lexsort((out[:, 2], out[:, 1], out[:, 0]))] if n.value else out

    def rings_active(self):
        nr = self.state.num_rings
        mask, uids, na = np.empty(nr, dtype=np.uint8), np.empty(nr, dtype=np.int64), C.c_int64()
        self._check(self.lib.mor_rings_download_active(self.h, _ptr(mask), _ptr(uids), C.byref(na)))
        return mask, uids, na.value

    def particle_neighbors(self):
        """(count[n], lists) — lists[i] in the reference's append order (None with only_count)."""
        count = np.zeros(self.n, dtype=np.int32)
        only = self.p_neighbors_cfg.only_count
        lst = None if only else np.empty((self.n, 15), dtype=np.int32)
        self._check(self.lib.mor_rings_download_neighbors(self.h, _ptr(count), _ptr(lst)))
        return count, (None if only else [lst[i, :count[i]].tolist() for i in range(self.n)])

    def time(self):
        ns, t = C.c_int64(), C.c_double()
        self.lib.mor_get_time(self.h, C.byref(ns), C.byref(t))
        return ns.value, t.value

    # fine-grained operators
    def clean_forces(self):
        self.lib.mor_clean_forces(self.h)

    def pair_forces(self):
        self.lib.mor_pair_forces(self.h)

    def walls_forces(self):
        self.lib.mor_walls_forces(self.h)

    def walls(self):
        self.lib.mor_walls(self.h)

    def update_verlet(self):
        self.lib.mor_update_verlet(self.h)

    def calc_diff(self, r1, r2):
        r1, r2, out = np.asarray(r1, float), np.asarray(r2, float), np.empty(2)
        self.lib.mor_calc_diff(self.h, _ptr(r1), _ptr(r2), _ptr(out))
        return out


def julia_div(x, y):
    return load().mor_julia_div(float(x), float(y))


def potential_force(kind, par, dr, dist=None):
    par = np.asarray(list(par) + [0.0] * (4 - len(par)), dtype=np.float64)
    dr = np.asarray(dr, dtype=np.float64)
    if dist is None:
        dist = float(np.sqrt(dr[0] * dr[0] + dr[1] * dr[1]))
    out = np.empty(2)
    load().mor_potential_force(kind, _ptr(par), _ptr(dr), dist, _ptr(out))
    return out


def szabo_interaction(par8, dr):
    par, dr, out = np.asarray(par8, float), np.asarray(dr, float), np.empty(2)
    load().mor_szabo_interaction(_ptr(par), _ptr(dr), _ptr(out))
    return out


def rtp_interaction(par4, dr):
    par, dr, out = np.asarray(list(par4) + [0.0] * 4, float), np.asarray(dr, float), np.empty(2)
    load().mor_rtp_interaction(_ptr(par), _ptr(dr), _ptr(out))
    return out
