#!/usr/bin/env python
"""bench.py — particle-steps/s of the Mavi.jl per-step hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this framework (libmavi_cuda.so through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...   # the reference ALGORITHM on the host cores (CPU oracle,
                                                            # Threaded mode: the reference is Julia and cannot run here)

Workload (N=1): the configuration the metric is quoted on — LJ lattice gas, 4000x4000 = 16M particles, periodic
rectangle, 3600x3600 Chunks, Float64, dt = 0.001, velocities uniform in [-0.2,0.2]^2 (seed 24042001), SURVEY.md 8d.
A "step" is one newton_step! (bin -> force pass 1 + drift -> force pass 2 + kick + walls) over all particles.
With N>1 ranks every rank owns an x-slab of 16M particles of a box N times as long (weak scaling).

One JSON line on stdout (rank 0).  `value`: state resident in HBM.  `e2e`: every step uploads pos+vel from pinned host
memory, steps once and downloads pos+vel (the stateless drop-in call a host-side caller makes).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

SEED = 24042001
B_ALG_NEWTON = 168.0  # algorithmic bytes per particle-step, Float64 Newton/LJ: bin 24 + A 64 + B 80
B_ALG_PASS_B = 80.0   # pass B: read pos', vel, F1 (48) + write vel', F2 (32)
B_ALG_PASS_A = 64.0   # pass A: read pos, vel (32) + write pos', F1 (32)
# Force carry (default): ONE launch per step does pass B of step n AND pass A of step n+1.  `roofline.achieved` follows
# the bench contract literally: SURVEY.md 8d's per-unit figure for the work the launch does (80 + 64 = 144 B per particle)
# x the particles it processes / its duration.  The launch's OWN compulsory traffic is smaller because the fusion removes
# re-reads (read pos', vel, F1 = 48; write vel', F2 == F1(n+1), pos'' = 48 -> 96 B): reported beside it as
# `roofline.min_traffic`, together with the measured DRAM bytes (`traffic`) — DESIGN.md 4.
B_ALG_FUSED = B_ALG_PASS_B + B_ALG_PASS_A
B_MIN_FUSED = 96.0


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def lj_workload(pkg, nx, ny, cuda_device=None, chunks=True, rank=0, world=1):
    """LJ lattice gas of SURVEY.md 8d (examples/chunks.jl geometry with README sigma=eps=1).  With world > 1 the box is
    `world` times as long (nx*world lattice columns) and only the particles of this rank's x-slab are generated."""
    dyn = pkg.LenJonesCfg(sigma=1.0, epsilon=1.0)
    r = pkg.particle_radius(dyn)
    nxt = nx * world
    ncols, nrows = int(nxt * 0.9), int(ny * 0.9)
    ccfg = pkg.ChunksCfg(num_cols=ncols, num_rows=nrows) if chunks else None
    int_cfg = pkg.IntCfg(dt=0.001, chunks_cfg=ccfg, device=cuda_device or pkg.CUDADevice())
    rng = np.random.default_rng(SEED + rank)
    if world == 1:
        pos, geom = pkg.rectangular_grid(nx, ny, 0.4, r)
        ids = None
    else:
        from mavi_jl_b200 import slabs
        line, geom = pkg.rectangular_grid(nxt, 1, 0.4, r)          # lattice x coordinates (repeated addition)
        colm, g2 = pkg.rectangular_grid(1, ny, 0.4, r)             # lattice y coordinates
        geom = pkg.RectangleCfg(length=geom.length, height=g2.height)
        xs, ys = line[:, 0], colm[:, 1]
        mine_x = np.flatnonzero(slabs.owner_of_column(slabs.column_of(xs, 0.0, geom.length, ncols), ncols, world) == rank)
        pos = np.empty((ny, len(mine_x), 2))
        pos[..., 0] = xs[mine_x][None, :]
        pos[..., 1] = ys[:, None]
        pos = pos.reshape(-1, 2)
        ids = (np.arange(ny)[:, None] * nxt + mine_x[None, :]).reshape(-1)
    vel = pkg.random_vel(len(pos), 1 / 5, rng=rng)
    space = pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom)
    return dict(pos=pos, vel=vel, space=space, dyn=dyn, int_cfg=int_cfg, geom=geom, ids=ids, n_global=nxt * ny)


def szabo_slabs(pkg, dist, torch, nx, ny, rank, world, local_rank, stream, steps, warmup, peak):
    """BASELINE config C3 at N > 1 (side measurement, never part of `value`): Szabo self-propelled particles, nx*world x ny
    lattice (examples/szabo.jl parameters, offset 1), periodic, x-slabs over the ranks, Philox noise; weak scaling like the
    headline.  All ranks call this collectively; returns the entry for `other_configs.szabo_c3`."""
    from mavi_jl_b200 import slabs
    box = [slabs.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    dyn = pkg.SzaboCfg(vo=1.0, mobility=1.0, relax_time=1.0, k_rep=10.0, k_adh=0.75, r_eq=1.0, r_max=1.1, rot_diff=0.01)
    r = pkg.particle_radius(dyn)
    nxt = nx * world
    ncols, nrows = nxt - world, ny - 1
    line, geom = pkg.rectangular_grid(nxt, 1, 1.0, r)
    colm, g2 = pkg.rectangular_grid(1, ny, 1.0, r)
    geom = pkg.RectangleCfg(length=geom.length, height=g2.height)
    xs, ys = line[:, 0], colm[:, 1]
    mine_x = np.flatnonzero(slabs.owner_of_column(slabs.column_of(xs, 0.0, geom.length, ncols), ncols, world) == rank)
    pos = np.empty((ny, len(mine_x), 2))
    pos[..., 0] = xs[mine_x][None, :]
    pos[..., 1] = ys[:, None]
    pos = pos.reshape(-1, 2)
    ids = (np.arange(ny)[:, None] * nxt + mine_x[None, :]).reshape(-1)
    ang = np.random.default_rng(SEED + rank).random(len(pos)) * 2 * np.pi
    dev = pkg.CUDADevice(device=local_rank, stream=stream, rank=rank, world=world, nccl_unique_id=box[0], n_global=nxt * ny,
                         rng_mode="philox")
    st = pkg.SelfPropelledState(pos=pos, pol_angle=ang)
    st.ids = ids
    s = pkg.System(state=st, space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                   int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(ncols, nrows), device=dev))
    try:
        s.step(warmup)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        s.step(steps)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = torch.tensor([float(s.local_count())], device="cuda", dtype=torch.float64)
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        ms = float(t.item()) / steps
        n_all = nxt * ny
        return {"ms_per_step": ms, "steps": steps, "value": n_all / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
                "roofline_step_frac": 88.0 * nx * ny / (ms * 1e-3) / 1e9 / peak, "bytes_per_particle_step": 88.0,
                "count_conserved": bool(int(total.item()) == n_all),
                "what": f"BASELINE config C3 on {world} GPUs: Szabo particles {nxt}x{ny} (16 M per GPU), periodic, f64, szabo_step!, "
                        "Philox noise, x-slabs with NCCL halo + migration (max over ranks)"}
    finally:
        s.close()


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML (nvidia_ml_py) is
    polled every ~5 ms so that a 0.1 s timed region still gets tens of samples; `nvidia-smi` (one sample per ~0.1 s
    process start) is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.samples, self.stop_flag = index, uuid, [], threading.Event()
        self.source = "nvidia-smi"
        self.ready = threading.Event()    # set once the first sample is in: the timed region starts after that
        self.t_begin = self.t_end = None  # samples outside [t_begin, t_end] (perf_counter) are dropped when both are set

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        h = None
        if self.uuid:
            for u in (self.uuid, self.uuid.encode()):
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(u)
                    break
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)  # fail here (-> fallback) rather than inside the loop
        self.source = "nvml"
        while True:  # at least one sample, also when stop is requested early
            sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
            except Exception:
                pw = 0.0
            try:
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                r = 0
            self.samples.append((time.perf_counter(), sm, mx, pw, [bool(r & b) for b in bits]))
            self.ready.set()
            if self.stop_flag.wait(0.005):
                break

    def _smi_loop(self):
        while True:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    s = [x.strip() for x in out.split(",")]
                    self.samples.append((time.perf_counter(), float(s[0]), float(s[1]), float(s[2]),
                                         [s[3 + i].lower().startswith("active") for i in range(4)]))
            except Exception:
                pass
            self.ready.set()
            if self.stop_flag.wait(0.02):
                break

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self.source = "nvidia-smi"
            self._smi_loop()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        smp = self.samples
        if self.t_begin is not None and self.t_end is not None:
            inside = [s for s in smp if self.t_begin <= s[0] <= self.t_end]
            smp = inside or smp  # a region shorter than one sampling period keeps the nearest samples
        if not smp:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(s[1] for s in smp)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[4][i] for s in smp)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smp[0][2], "reasons": reasons,
                "power_w_max": max(s[3] for s in smp), "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------------------------- CPU arm
def workload_config(nx, ny, world):
    """`config` of the JSON line — identical in both arms (the reference arm times the same workload)."""
    n = nx * ny
    return {"workload": f"LJ lattice gas {nx}x{ny}={n} particles per GPU, periodic rectangle, {int(nx * 0.9)}x{int(ny * 0.9)} chunks, "
                        f"f64, dt=0.001, newton_step!", "l2": "state arrays (>=1.5 GB) exceed the 126 MB L2; no flush needed",
            "parallelism": "single GPU" if world == 1 else f"{world} x-slabs of one {nx * world}x{ny} periodic box, NCCL halo+migration"}


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def cpu_reference_rate(pkg, steps, warmup, nx=1000, ny=1000, threads=None, fast=True):
    """The reference algorithm on the host cores: the C restatement (oracle source) in Threaded mode (column-partitioned
    pair loop with private force slices, src/integration.jl:159-194; serial re-bin; two force passes), built with
    `-O3 -march=native -fopenmp` as BASELINE.md 4 prescribes (`fast`; a separate target from the parity oracle)."""
    oracle = entry.load_oracle()
    from mavi_jl_b200.params import lower
    threads = threads or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    w = lj_workload(pkg, nx, ny)
    st = pkg.SecondLawState(pos=w["pos"], vel=w["vel"])
    o = oracle.OracleSystem(state=st, space_cfg=w["space"], dynamic_cfg=w["dyn"], int_cfg=w["int_cfg"], lower=lower,
                            threads=threads, fast=fast)
    if warmup:
        o.step(warmup)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    o.close()
    n = nx * ny
    return n * steps / dt, dt, threads, (f"LJ lattice {nx}x{ny} = {n} particles, {steps} timed steps after {warmup} warm-up, "
                                        f"{'-O3 -march=native' if fast else '-O2 -ffp-contract=off'} -fopenmp, {threads} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()
    # the metric's own workload (16M) when the host has the memory for the reference's cell table (1.9 GB) + state,
    # else a bounded 1M sample of the same lattice (same density and particles per cell)
    full = host_mem_available_gb() > 24.0 and not args.cpu_small
    # ... and when `steps + warmup` passes over it end within a few minutes at the ~7e6 particle-steps/s the host cores reach
    # (16 M x 20 steps = 45 s; 200 steps would be 7 minutes -> the 1 M sample, same lattice, rate-normalised)
    if full and (args.steps + args.warmup) * args.nx * args.ny / 7.0e6 > 150.0:
        full = False
    nx, ny = (args.nx, args.ny) if full else (args.cpu_sample, args.cpu_sample)
    rate, secs, threads, sample = cpu_reference_rate(pkg, args.steps, args.warmup, nx=nx, ny=ny)
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.nx, args.ny, args.gpus),
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": ("reference ALGORITHM restated in C (oracle source, Threaded mode, -O3 -march=native -fopenmp); the reference "
                 "itself is Julia, which is not installed: parity unpinned"),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- parity probe
def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    sc = np.abs(b).max()
    return float(np.abs(a - b).max() / (sc if sc > 0 else 1.0))


def lj_pair_force_scale(pkg, offset=0.4):
    """|LJ pair force| between lattice neighbours of the workload (spacing r(2 + offset), sigma = eps = 1): the scale net
    forces are a cancellation residue of on the perfect lattice."""
    dyn = pkg.LenJonesCfg(sigma=1.0, epsilon=1.0)
    a = pkg.particle_radius(dyn) * (2 + offset)
    return abs(24.0 * (2.0 / a ** 12 - 1.0 / a ** 6) / a)


def parity_probe_single(pkg, nx, ny, stream, local_rank, flags, steps=3):
    """N = 1: the bench workload itself, `steps` newton_step!s from the initial state on the device (first pass + carried
    steps) against the PARITY oracle (-O2 -ffp-contract=off, Threaded) — every particle compared."""
    from mavi_jl_b200.params import lower
    oracle = entry.load_oracle()
    t0 = time.perf_counter()
    w = lj_workload(pkg, nx, ny, cuda_device=pkg.CUDADevice(device=local_rank, stream=stream, flags=flags))
    g = pkg.System(state=pkg.SecondLawState(pos=w["pos"].copy(), vel=w["vel"].copy()), space_cfg=w["space"],
                   dynamic_cfg=w["dyn"], int_cfg=w["int_cfg"])
    try:
        g.step(steps)
        g.sync_to_host()
        gp, gv, gf = g.state.pos, g.state.vel, g.get_forces()
        g.calc_forces()                 # clean_forces! + update_chunks! + calc_forces! at the final positions
        gf_fresh = g.get_forces()
        cg, _ = g.download_cells()
    finally:
        g.close()
    threads = os.cpu_count() or 1
    o = oracle.OracleSystem(state=pkg.SecondLawState(pos=w["pos"], vel=w["vel"]), space_cfg=w["space"], dynamic_cfg=w["dyn"],
                            int_cfg=w["int_cfg"], lower=lower, threads=threads)
    o.step(steps)
    of = o.get_forces()
    # trajectory: positions / velocities after `steps` steps.  Forces are compared on IDENTICAL positions (the oracle takes the
    # device's final state; both run calc_forces!): once positions differ in the last bit (ulp(4700) = 9e-13) the r^-13 law
    # turns one ulp into ~1e-11 of a pair force, which says nothing about the force evaluation (listed as force_after_steps)
    o2 = oracle.OracleSystem(state=pkg.SecondLawState(pos=gp.copy(), vel=gv.copy()), space_cfg=w["space"], dynamic_cfg=w["dyn"],
                             int_cfg=w["int_cfg"], lower=lower, threads=threads)
    o2.calc_forces()
    co, _ = o2.download_cells()     # update_chunks! of the same positions: bit-exact cell assignment
    errs = {"pos": float(np.abs(gp - o.pos()).max() / w["geom"].length), "vel": _rel(gv, o.second()),
            "force": _rel(gf_fresh, o2.get_forces())}
    o2.close()
    out = {"max_rel_err": max(errs.values()), "n_checked": int(nx * ny), "errs": errs, "force_after_steps": _rel(gf, of),
           "force": "calc_forces! on identical positions (the device state after the steps)", "cells_bit_exact": bool(np.array_equal(cg, co)),
           "steps": steps, "against": f"parity oracle (C restatement, Threaded, {threads} threads), every particle, {steps} steps from the initial state",
           "seconds": time.perf_counter() - t0}
    o.close()
    return out


def parity_probe_slabs(pkg, dist, system, w, nx, ny, rank, world, local_rank, stream, total_steps, strip_cols=2):
    """N > 1: after the timed region every rank hands rank 0 the particles it holds within `strip_cols` cell columns of its
    two slab boundaries (ids, positions, velocities, forces); rank 0 re-runs the SAME global system for the same number of
    steps as ONE single-domain device run (no slabs, no NCCL: the path tests/ compare with the oracle) and checks (1) the
    strips hold exactly the particles the single-domain run has there, (2) their state agrees, (3) the owned counts sum to
    the global particle count.  All ranks take part in the collectives unconditionally."""
    import torch
    from mavi_jl_b200 import slabs
    t0 = time.perf_counter()
    geom = w["geom"]
    ncols = int(nx * world * 0.9)
    cl = geom.length / ncols
    ids, pos, vel, frc = system.download_local()
    lo, m = slabs.slab_columns(ncols, world, rank)
    col = slabs.column_of(pos[:, 0], 0.0, geom.length, ncols)
    d = (col - lo) % ncols                      # 0 .. m-1 for owned columns
    sel = (d < strip_cols) | (d >= m - strip_cols)
    mine = {"ids": ids[sel].copy(), "pos": pos[sel].copy(), "vel": vel[sel].copy(), "force": frc[sel].copy(), "n_local": int(len(ids)),
            "foreign": int(((d < 0) | (d >= m)).sum())}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    out = None
    if rank == 0:
        try:
            n_global = nx * world * ny
            gp = np.empty((n_global, 2))
            gv = np.empty((n_global, 2))
            for r in range(world):          # the per-rank generators, placed by global id
                wr = w if r == 0 else lj_workload(pkg, nx, ny, rank=r, world=world)
                gp[wr["ids"]] = wr["pos"]
                gv[wr["ids"]] = wr["vel"]
            dyn = w["dyn"]
            int_cfg = pkg.IntCfg(dt=0.001, chunks_cfg=pkg.ChunksCfg(num_cols=ncols, num_rows=int(ny * 0.9)),
                                 device=pkg.CUDADevice(device=local_rank, stream=stream))
            ref = pkg.System(state=pkg.SecondLawState(pos=gp, vel=gv), space_cfg=w["space"], dynamic_cfg=dyn, int_cfg=int_cfg)
            try:
                ref.step(total_steps)
                ref.sync_to_host()
                rf = ref.get_forces()
            finally:
                ref.close()
            rp, rv = ref.state.pos, ref.state.vel
            gid = np.concatenate([g["ids"] for g in gathered])
            spos = np.concatenate([g["pos"] for g in gathered])
            svel = np.concatenate([g["vel"] for g in gathered])
            sfrc = np.concatenate([g["force"] for g in gathered])
            # the particles the single-domain run has inside the same strips
            rcol = slabs.column_of(rp[:, 0], 0.0, geom.length, ncols)
            in_strip = np.zeros(n_global, dtype=bool)
            for r in range(world):
                lo_r, m_r = slabs.slab_columns(ncols, world, r)
                dr = (rcol - lo_r) % ncols
                in_strip |= (dr < strip_cols) | ((dr >= m_r - strip_cols) & (dr < m_r))
            want = np.flatnonzero(in_strip)
            same_set = bool(len(gid) == len(want) and np.array_equal(np.sort(gid), want))
            errs = {"pos": float(np.abs(spos - rp[gid]).max() / geom.length), "vel": _rel(svel, rv[gid]), "force": _rel(sfrc, rf[gid])}
            total = sum(g["n_local"] for g in gathered)
            out = {"max_rel_err": max(errs.values()), "n_checked": int(len(gid)), "errs": errs, "strip_sets_equal": same_set,
                   "bitwise_equal": bool(np.array_equal(spos, rp[gid]) and np.array_equal(svel, rv[gid]) and np.array_equal(sfrc, rf[gid])),
                   "owned_total": int(total), "n_global": int(n_global), "count_conserved": bool(total == n_global),
                   "foreign_particles": int(sum(g["foreign"] for g in gathered)), "steps": int(total_steps),
                   "against": (f"single-domain run of the same {nx * world}x{ny} system on one GPU (the path tests/ compare with the "
                               f"oracle), particles within {strip_cols} cell columns of every slab boundary, after {total_steps} steps"),
                   "seconds": time.perf_counter() - t0}
        except Exception as exc:  # noqa: BLE001
            out = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmavi_cuda.so has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the library's default for its own communicator (csrc/slab.cu); NCCL reads it once per process, and torch
        # initialises NCCL first here, so it has to be in the environment before that
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "2")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    pkg.load_library()

    nx, ny = args.nx, args.ny
    stream = torch.cuda.current_stream().cuda_stream
    uid = None
    if world > 1:
        from mavi_jl_b200 import slabs
        box = [slabs.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    dev = pkg.CUDADevice(device=local_rank, stream=stream, flags=args.flags, rank=rank, world=world, nccl_unique_id=uid,
                         n_global=nx * world * ny)
    w = lj_workload(pkg, nx, ny, cuda_device=dev, rank=rank, world=world)
    n = nx * ny  # particles per GPU (weak scaling)
    state = pkg.SecondLawState(pos=w["pos"], vel=w["vel"])
    if world > 1:
        state.ids = w["ids"]
    system = pkg.System(state=state, space_cfg=w["space"], dynamic_cfg=w["dyn"], int_cfg=w["int_cfg"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident throughput (`value`) ---------------------------------------------------------------
    system.step(args.warmup)
    system.set_profiling(True)
    launches0 = system.launch_count()
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    sampler.start()
    sampler.ready.wait(timeout=10)  # NVML initialised and the first sample taken before the timed region starts
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    system.step(args.steps)
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    launches = system.launch_count() - launches0
    phase_ms = system.last_step_ms()  # bin+sort, pass A, pass B of the last step (CUDA events on the launching stream)
    system.set_profiling(False)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms * 1e-3)
    counters = system.counters()  # cell changes / tile repairs / emigrants summed over warm-up + timed steps (this rank)

    # ---- parity probe at N > 1: the slab result against a single-domain run of the same global system ------------
    probe = None
    if world > 1 and not args.no_probe:
        probe = parity_probe_slabs(pkg, dist, system, w, nx, ny, rank, world, local_rank, stream, args.warmup + args.steps)

    # ---- end to end through host buffers (`e2e`) -------------------------------------------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    slab_api = world > 1 or bool(args.flags & pkg.capi.FLAG_SLAB_SELF)
    if not slab_api:
        pin_pos = torch.empty((n, 2), dtype=torch.float64).pin_memory()
        pin_vel = torch.empty((n, 2), dtype=torch.float64).pin_memory()
        system.sync_to_host()
        pin_pos.numpy()[...] = system.state.pos
        pin_vel.numpy()[...] = system.state.vel
        system.state.pos, system.state.vel = pin_pos.numpy(), pin_vel.numpy()
        h2d = d2h = 32 * n

        def e2e_step():
            system.upload_state()   # H2D pos+vel from pinned memory (+ constructor-time checks and binning)
            system.step(1)
            system.sync_to_host()   # D2H pos+vel into pinned memory
    else:
        # slab mode: the owned set changes with migration, so ids travel with pos+vel; pinned buffers sized for the
        # staging capacity, forces are not downloaded
        cap = int(n * 1.25) + 4096
        pin = (torch.empty(cap, dtype=torch.int64).pin_memory(), torch.empty((cap, 2), dtype=torch.float64).pin_memory(),
               torch.empty((cap, 2), dtype=torch.float64).pin_memory())
        out = tuple(t_.numpy() for t_ in pin)
        h2d = d2h = 40 * n

        def e2e_step():
            ids, pos, vel, _ = system.download_local(out=out, want_forces=False)   # D2H ids+pos+vel of the owned particles
            system.local_ids, system.state.pos, system.state.vel = ids, pos, vel
            system.upload_local()                        # H2D from the same pinned buffers, re-binning, first halo exchange
            system.step(1)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    peak, peak_kind = measured_peak_hbm()
    szabo_multi = None
    if world > 1 and not args.no_other_configs:   # collective: every rank takes part
        try:
            szabo_multi = szabo_slabs(pkg, dist, torch, nx, ny, rank, world, local_rank, stream, max(10, args.steps // 4), args.warmup, peak)
        except Exception as exc:  # noqa: BLE001
            szabo_multi = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same N only)
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tr = json.load(f)["dram_bytes_per_launch"]
    except Exception:
        tr = {}
    carry = not (args.flags & pkg.capi.FLAG_NO_FORCE_CARRY)
    dom = "pass_b" if phase_ms[2] >= phase_ms[1] else "pass_a"
    dom_ms = phase_ms[2] if dom == "pass_b" else phase_ms[1]
    dom_bytes = (B_ALG_PASS_B if dom == "pass_b" else B_ALG_PASS_A) * n
    if carry:
        dom, dom_bytes = "fused_pass", B_ALG_FUSED * n
    if n == 16_000_000 and world == 1 and dom in tr:   # a capture of THIS kernel on THIS workload exists; null otherwise
        traffic = tr[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_achieved = B_ALG_NEWTON * n * world / (ms * 1e-3 / args.steps) / 1e9
    # ---- side measurements (never part of `value`; each one is fault-isolated so that it cannot cost the line) --------
    def timed_steps(sysx, k):
        barrier()
        e0.record()
        sysx.step(k)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / k

    def side(fn):
        try:
            return fn()
        except Exception as exc:  # noqa: BLE001 — report, do not lose the main measurement
            return {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if world == 1:
        system.close()
        system = None
    mkdev = lambda **kw: pkg.CUDADevice(device=local_rank, stream=stream, **kw)  # noqa: E731

    def m_two_pass():
        # A/B: the same workload with the carry switched off (two full force passes per step), for transparency
        w2 = lj_workload(pkg, nx, ny, cuda_device=mkdev(flags=args.flags | pkg.capi.FLAG_NO_FORCE_CARRY))
        s2 = pkg.System(state=pkg.SecondLawState(pos=w2["pos"], vel=w2["vel"]), space_cfg=w2["space"], dynamic_cfg=w2["dyn"], int_cfg=w2["int_cfg"])
        try:
            s2.step(args.warmup)
            k2 = max(10, args.steps // 4)
            return {"ms_per_step": timed_steps(s2, k2), "steps": k2, "what": "MAVI_FLAG_NO_FORCE_CARRY: full first + second force pass every step"}
        finally:
            s2.close()

    def m_float32():
        # the optional Float32 mode (mavi_f32 build of the same kernels) on the same workload, for reference; the
        # headline metric stays Float64.  88 B per particle-step
        w3 = lj_workload(pkg, nx, ny, cuda_device=mkdev(flags=args.flags))
        s3 = pkg.System(state=pkg.SecondLawState(pos=w3["pos"].astype(np.float32), vel=w3["vel"].astype(np.float32)),
                        space_cfg=w3["space"], dynamic_cfg=w3["dyn"], int_cfg=w3["int_cfg"])
        try:
            s3.step(args.warmup)
            k3 = max(10, args.steps // 2)
            ms3 = timed_steps(s3, k3)
            return {"ms_per_step": ms3, "steps": k3, "value": n / (ms3 * 1e-3), "unit": "particle-steps/s",
                    "roofline_step_frac": 88.0 * n / (ms3 * 1e-3) / 1e9 / peak, "bytes_per_particle_step": 88.0,
                    "what": "MaviParams.dtype = MAVI_F32 (Float32 state and arithmetic), same kernels compiled with real = float"}
        finally:
            s3.close()

    def m_szabo():
        # BASELINE config C3 on this GPU: examples/szabo.jl parameters, lattice offset 1, cells (n-1)^2, dt 0.01, Philox noise
        dyn = pkg.SzaboCfg(vo=1.0, mobility=1.0, relax_time=1.0, k_rep=10.0, k_adh=0.75, r_eq=1.0, r_max=1.1, rot_diff=0.01)
        pos, geom = pkg.rectangular_grid(nx, ny, 1.0, pkg.particle_radius(dyn))
        ang = np.random.default_rng(SEED).random(nx * ny) * 2 * np.pi
        s4 = pkg.System(state=pkg.SelfPropelledState(pos=pos, pol_angle=ang),
                        space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                        int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(nx - 1, ny - 1), device=mkdev(rng_mode="philox")))
        try:
            s4.step(args.warmup)
            k4 = max(10, args.steps // 4)
            ms4 = timed_steps(s4, k4)
            return {"ms_per_step": ms4, "steps": k4, "value": n / (ms4 * 1e-3), "unit": "particle-steps/s",
                    "roofline_step_frac": 88.0 * n / (ms4 * 1e-3) / 1e9 / peak, "bytes_per_particle_step": 88.0,
                    "what": f"BASELINE config C3 on one GPU: Szabo self-propelled particles {nx}x{ny}, periodic, f64, szabo_step!, Philox noise"}
        finally:
            s4.close()

    def m_rings():
        # BASELINE config C4: 400 x 250 rings x 10 particles (test/tests_rings/rings_utils.jl:35-53 parameters), periodic
        from mavi_jl_b200.rings import configs as rc
        from mavi_jl_b200.rings import init_states as ri
        from mavi_jl_b200.rings.rings import RingsSystem
        from mavi_jl_b200.rings.states import RingsState
        inter = rc.HarmTruncCfg(k_rep=20, k_atr=4, dist_eq=1, dist_max=1 + 0.2)
        dyn = rc.RingsCfg(p0=3.5, relax_time=1, vo=1.0, mobility=1, rot_diff=0.05, k_area=1, k_spring=20, l_spring=1,
                          num_particles=10, interaction_finder=inter)
        cols, rows = 400, 250
        rings_pos, geom = ri.rectangular_grid(num_cols=cols, num_rows=rows, num_particles=10, p_radius=dyn.particle_radius(),
                                              pad_x=0.1, pad_y=0.1)
        pol = ri.random_pol(cols * rows, rng=np.random.default_rng(SEED))
        max_size = inter.dist_max * 1.1
        chunks = pkg.ChunksCfg(int(geom.length // max_size), int(geom.height // max_size))
        s5 = RingsSystem(state=RingsState(rings_pos=rings_pos, pol=pol),
                         space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom), dynamic_cfg=dyn,
                         int_cfg=rc.RingsIntCfg(dt=0.01, p_chunks_cfg=chunks, device=mkdev(rng_mode="philox")))
        try:
            s5.step(args.warmup)
            k5 = max(10, args.steps // 4)
            ms5 = timed_steps(s5, k5)
            n5 = cols * rows * 10
            return {"ms_per_step": ms5, "steps": k5, "value": n5 / (ms5 * 1e-3), "unit": "particle-steps/s",
                    "roofline_step_frac": 108.0 * n5 / (ms5 * 1e-3) / 1e9 / peak, "bytes_per_particle_step": 108.0,
                    "what": "BASELINE config C4: Mavi.Rings, 100k rings x 10 particles, periodic, f64, Rings step!, Philox noise"}
        finally:
            s5.close()

    def m_hot():
        # the same box once the lattice has broken up: `hot_steps` more steps (dt = 0.001; the offset-0.4 lattice sits above
        # the LJ minimum, so it is unstable and heats up), then the same timed region.  Re-binning traffic is O(movers):
        # this is the step cost at a realistic mover rate; tile overflows on the way grow the tiles and rebuild.
        w6 = lj_workload(pkg, nx, ny, cuda_device=mkdev(flags=args.flags))
        s6 = pkg.System(state=pkg.SecondLawState(pos=w6["pos"], vel=w6["vel"]), space_cfg=w6["space"], dynamic_cfg=w6["dyn"], int_cfg=w6["int_cfg"])
        try:
            s6.step(args.hot_steps)
            c0 = s6.counters()
            k6 = args.steps
            ms6 = timed_steps(s6, k6)
            c1 = s6.counters()
            ke, _ = s6.energies(want_pe=False)
            return {"ms_per_step": ms6, "steps": k6, "value": n / (ms6 * 1e-3), "unit": "particle-steps/s", "steps_before": args.hot_steps,
                    "rebinned_per_step": (c1["rebinned"] - c0["rebinned"]) / k6, "tiles_repaired_per_step": (c1["tiles_repaired"] - c0["tiles_repaired"]) / k6,
                    "rebuilds_total": c1["rebuilds"], "tile_cap": c1["tile_cap"], "kinetic_energy_per_particle": ke / n,
                    "roofline_step_frac": B_ALG_NEWTON * n / (ms6 * 1e-3) / 1e9 / peak,
                    "what": f"same workload after {args.hot_steps} steps (lattice broken up), force carry + incremental repair"}
        finally:
            s6.close()

    slab_api_main = world > 1 or bool(args.flags & pkg.capi.FLAG_SLAB_SELF)
    hot = side(m_hot) if (world == 1 and args.hot_steps > 0 and not slab_api_main) else None
    if world == 1 and not args.no_probe and not slab_api_main:
        probe = side(lambda: parity_probe_single(pkg, nx, ny, stream, local_rank, args.flags))
    two_pass = side(m_two_pass) if (world == 1 and carry and not args.no_two_pass and not slab_api_main) else None
    float32 = side(m_float32) if (world == 1 and not args.no_f32 and not slab_api_main) else None
    other = None
    if world == 1 and not args.no_other_configs and not slab_api_main:
        other = {"szabo_c3": side(m_szabo), "rings_c4": side(m_rings)}
    elif szabo_multi is not None:
        other = {"szabo_c3": szabo_multi}
    def m_cpu():
        # the C restatement built -O3 -march=native -fopenmp, all host cores, at 1M and (memory permitting) at
        # the metric's own 16M; `value` is the 16M figure when it ran.  Bounded: a few timed steps each.
        cs = args.cpu_sample
        rate, secs, threads, sample = cpu_reference_rate(pkg, args.cpu_steps, 1, nx=cs, ny=cs)
        out = {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample,
               "n1m": {"value": rate, "sample": sample}}
        if host_mem_available_gb() > 24.0 and not args.cpu_small:
            r16, _, _, s16 = cpu_reference_rate(pkg, max(2, args.cpu_steps // 3), 1, nx=nx, ny=ny)
            out.update({"value": r16, "sample": s16, "n16m": {"value": r16, "sample": s16}})
        # the reference's DEFAULT device is Sequencial (one thread, src/configs.jl:471-487): reported beside the Threaded figure
        r1, _, _, s1 = cpu_reference_rate(pkg, max(2, args.cpu_steps // 3), 1, nx=max(100, cs // 2), ny=max(100, cs // 2), threads=1)
        out["sequencial"] = {"value": r1, "cores": 1, "sample": s1}
        # the parity oracle build (-O2 -ffp-contract=off) of the same source, for transparency about what -O3 buys
        r2, _, _, s2 = cpu_reference_rate(pkg, max(2, args.cpu_steps // 3), 1, nx=cs, ny=cs, fast=False)
        out["parity_build"] = {"value": r2, "sample": s2}
        return out

    cpu = side(m_cpu) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(nx, ny, world),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind, "kernel_ms": dom_ms,
                     "bytes_per_particle": dom_bytes / n,
                     "accounting": ("SURVEY.md 8d algorithmic bytes of the work this launch does: pass B of step n (80 B) + pass A "
                                    "of step n+1 (64 B) per particle") if carry else "SURVEY.md 8d algorithmic bytes of this pass",
                     "min_traffic": ({"bytes_per_particle": B_MIN_FUSED, "achieved": B_MIN_FUSED * n / (dom_ms * 1e-3) / 1e9,
                                      "frac": B_MIN_FUSED * n / (dom_ms * 1e-3) / 1e9 / peak,
                                      "what": "the fused launch's own compulsory DRAM traffic (the fusion removes the re-reads "
                                              "between the two passes); compare with `traffic`"} if (carry and dom_ms > 0) else None),
                     "what": ("k_newton_b<CARRY>: force pass 2 + kick + walls! + re-bin decision of step n fused with force pass 1 + drift of "
                              "step n+1 (F1(n+1) = F2(n) except near re-binned particles, which are recomputed sparsely; bit-identical "
                              "to two full passes)") if carry else "two full force passes per step (MAVI_FLAG_NO_FORCE_CARRY)",
                     "step": {"achieved": step_achieved / world, "frac": step_achieved / world / peak, "bytes_per_particle_step": B_ALG_NEWTON},
                     "phase_ms": {"pass_a": phase_ms[1], "pass_b": phase_ms[2], "repair_exchange": phase_ms[3]}},
        "cpu_baseline": cpu,
        "parity_probe": probe,
        "rebinned_per_step": counters["rebinned"] / max(counters["steps"], 1),
        "hot": hot,
        "two_pass": two_pass,
        "float32": float32,
        "other_configs": other,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "steps": e2e_steps,
                "what": ("per step and rank: mavi_download_local(ids,pos,vel into pinned host) + mavi_upload_local(ids,pos,vel) + mavi_step(1)"
                         if slab_api else "per step: mavi_upload_state(pos,vel from pinned host) + mavi_step(1) + mavi_download_state(pos,vel)")},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=4000)
    ap.add_argument("--ny", type=int, default=4000)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1000, help="CPU baseline sample: an n x n lattice")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-small", action="store_true", help="CPU legs: the 1M sample only (skip the 16M run)")
    ap.add_argument("--no-probe", action="store_true", help="skip the parity probe")
    ap.add_argument("--hot-steps", type=int, default=2000, help="steps before the `hot` side measurement (0 = skip)")
    ap.add_argument("--no-two-pass", action="store_true", help="skip the A/B run with the force carry switched off")
    ap.add_argument("--no-f32", action="store_true", help="skip the Float32-mode side measurement")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the Szabo (C3) and Rings (C4) side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
