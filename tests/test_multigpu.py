"""Multi-GPU x-slab path: world_size-2 gloo tests of the host-side logic on CPU, and (with >= 2 GPUs) parity of the
NCCL halo/migration path against the single-domain oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pkg = H.pkg
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- CPU: slab arithmetic
@pytest.mark.parametrize("C,world", [(3600, 8), (3600, 7), (29, 4), (16, 8), (57, 2)])
def test_slab_columns_cover_the_grid(C, world):
    from mavi_jl_b200 import slabs
    nxt = 0
    for r in range(world):
        lo, m = slabs.slab_columns(C, world, r)
        assert lo == nxt and m >= C // world
        assert np.all(slabs.owner_of_column(np.arange(lo, lo + m), C, world) == r)
        nxt = lo + m
    assert nxt == C


def test_partition_matches_oracle_binning(oracle):
    from mavi_jl_b200 import slabs
    case = H.newton_case(nx=60, ny=20, wall="periodic", jitter=0.45)
    o = H.make_oracle(case)
    cell, _ = o.download_cells()
    ccfg = case["int_cfg"].chunks_cfg
    col = cell // ccfg.num_rows
    pos = case["mk"]().pos
    assert np.array_equal(slabs.column_of(pos[:, 0], 0.0, case["geom"].length, ccfg.num_cols), col)
    own = slabs.partition(pos, case["geom"], ccfg.num_cols, 4)
    for r in range(4):
        lo, m = slabs.slab_columns(ccfg.num_cols, 4, r)
        assert np.all((col[own == r] >= lo) & (col[own == r] < lo + m))


_GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import helpers as H
pkg = H.pkg
from mavi_jl_b200 import slabs
from mavi_jl_b200.params import lower
import __graft_entry__ as entry
oracle = entry.load_oracle()
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = H.newton_case(nx=40, ny=24, dyn=pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2), wall="periodic", jitter=0.3)
st = case["mk"]()
ccfg = case["int_cfg"].chunks_cfg
C = ccfg.num_cols
owner = slabs.partition(st.pos, case["geom"], C, world)
col = slabs.column_of(st.pos[:, 0], 0.0, case["geom"].length, C)
lo, m = slabs.slab_columns(C, world, rank)
mine = np.flatnonzero(owner == rank)
# halo = ONE cell column on each side (periodic), exchanged between ranks
send = {}
for side, c in (("left", lo), ("right", lo + m - 1)):
    send[side] = mine[col[mine] == c]
got = [None] * world
dist.all_gather_object(got, (rank, send["left"], send["right"]))
left_rank, right_rank = (rank - 1) % world, (rank + 1) % world
halo = np.concatenate([g[2] for g in got if g[0] == left_rank] + [g[1] for g in got if g[0] == right_rank])
# forces on my particles from (mine + halo) only must equal the global oracle's forces: one halo column suffices
o = H.make_oracle(case)
o.calc_forces()
F = o.get_forces()
sub = np.concatenate([mine, halo])
pos = st.pos[sub]
size = np.array([case["geom"].length, case["geom"].height])
dr = pos[:len(mine), None, :] - pos[None, :, :]
dr = dr - (np.abs(dr) > size / 2) * np.copysign(size, dr)
d = np.sqrt((dr ** 2).sum(-1))
d[np.arange(len(mine)), np.arange(len(mine))] = np.inf
dyn = case["dyn"]
fm = np.where(d < dyn.dist_eq, -dyn.k_rep * (d / dyn.dist_eq - 1), -dyn.k_atr * (d / dyn.dist_eq - 1))
c = np.where(d > dyn.dist_max, 0.0, fm / d)
Floc = (c[..., None] * dr).sum(1)
err = np.abs(Floc - F[mine]).max() / np.abs(F).max()
uid = [os.urandom(128) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ids_all = [None] * world
dist.all_gather_object(ids_all, (mine, uid[0]))
ok = err < 1e-12 and all(u == ids_all[0][1] for _, u in ids_all)
ok = ok and len(np.unique(np.concatenate([i for i, _ in ids_all]))) == len(st.pos)
print(f"rank {rank}: n={len(mine)} halo={len(halo)} err={err:.2e} ok={ok}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slab_host_logic(tmp_path, world):
    """world_size-2 / -3 gloo runs on CPU (3: distinct left and right neighbours, uneven column split): partition, halo
    selection (one cell column per side), id routing and the broadcast of the communicator id behave; forces from
    (owned + halo) equal the global oracle forces."""
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                          "127.0.0.1", "--master-port", str(29529 + world), str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok=True") == world


# ---------------------------------------------------------------- GPU: NCCL slabs vs oracle
def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [0, 8])  # 8 = MAVI_FLAG_SMALL_BLOCKS: several CTAs per tile row -> halo exchange overlapped with interior blocks
@pytest.mark.parametrize("kind,steps", [("lj", 150), ("harm", 150), ("szabo", 40)])  # overlapping Szabo particles are chaotic: short horizon
def test_multigpu_slabs_match_oracle(cuda_lib, kind, steps, flags):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", NCCL_DEBUG="WARN")
    port = 29533 + {"lj": 0, "harm": 1, "szabo": 2}[kind] + (3 if flags else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), kind, str(steps), str(flags)]
    # own process group, so that a hung rank can be killed by its exact pgid without touching anything else
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=150)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail("multi-GPU worker hung:\n" + out[-2000:] + err[-3000:])
    assert "MGPU_OK" in out, out[-3000:] + err[-3000:]


# ---------------------------------------------------------------- host cell rule (mavi_cells_of_points), CPU
def test_cells_of_points_bit_exact_against_oracle(oracle):
    """mavi_cells_of_points (the rule the single-process multi-GPU handle and slabs.column_of partition with) ==
    update_particle_chunk! as the oracle restates it, including points placed exactly on cell edges."""
    from mavi_jl_b200 import slabs
    case = H.newton_case(nx=60, ny=44, wall="periodic", jitter=0.45)
    ccfg = case["int_cfg"].chunks_cfg
    geom = case["geom"]
    st = case["mk"]()
    rng = np.random.default_rng(3)
    pos = st.pos.copy()
    cl, ch = geom.length / ccfg.num_cols, geom.height / ccfg.num_rows
    k = rng.integers(0, ccfg.num_cols + 1, len(pos) // 3)
    pos[: len(k), 0] = k * cl                                       # rounded multiples of the cell length
    pos[len(k): 2 * len(k), 0] = np.nextafter(k * cl, 0)
    r = rng.integers(0, ccfg.num_rows + 1, len(pos) // 3)
    pos[: len(r), 1] = geom.height - r * ch
    pos = np.clip(pos, 0.0, [geom.length, geom.height])
    case2 = dict(case, mk=lambda: pkg.SecondLawState(pos=pos.copy(), vel=st.vel.copy()))
    o = H.make_oracle(case2)
    cell, _ = o.download_cells()
    got = slabs.cells_of_points(pos, geom, ccfg.num_cols, ccfg.num_rows)
    assert np.array_equal(got, cell)
    # outside the grid -> -1 (BoundsError in the reference)
    out = slabs.cells_of_points(np.array([[-2 * cl, 1.0], [1.0, geom.height + 2 * ch], [np.nan, 1.0]]), geom, ccfg.num_cols, ccfg.num_rows)
    assert np.all(out == -1)
    # Float32 points are promoted (Chunks keeps Float64 geometry, src/chunks.jl:13,27-30)
    p32 = pos.astype(np.float32)
    assert np.array_equal(slabs.cells_of_points(p32, geom, ccfg.num_cols, ccfg.num_rows),
                          slabs.cells_of_points(p32.astype(np.float64), geom, ccfg.num_cols, ccfg.num_rows))


def test_single_process_multi_gpu_fails_loudly_without_devices(cuda_lib):
    """MaviParams.n_gpus > 1 with fewer visible devices: MAVI_ERR_CUDA and a message, never a silent single-GPU / CPU run.
    Unsupported configurations are rejected before any device is touched."""
    if _ngpus() >= 2:
        pytest.skip("this box has the devices")
    case = H.newton_case(nx=24, ny=24, wall="periodic")
    case["int_cfg"] = pkg.IntCfg(dt=0.001, chunks_cfg=case["int_cfg"].chunks_cfg, device=pkg.CUDADevice(n_gpus=2))
    with pytest.raises(pkg.MaviError) as e:
        H.make_gpu(case)
    assert e.value.status == pkg.capi.ERR_CUDA and "CUDA devices" in str(e.value)
    rigid = H.newton_case(nx=24, ny=24, wall="rigid")
    rigid["int_cfg"] = pkg.IntCfg(dt=0.001, chunks_cfg=rigid["int_cfg"].chunks_cfg, device=pkg.CUDADevice(n_gpus=2))
    with pytest.raises(pkg.MaviError) as e:
        H.make_gpu(rigid)
    assert e.value.status == pkg.capi.ERR_UNSUPPORTED


# ---------------------------------------------------------------- GPU: ONE process, several GPUs inside the handle
@pytest.mark.gpu
@pytest.mark.parametrize("kind,steps,flags", [("lj", 150, 0), ("lj", 150, 8), ("harm", 150, 8), ("szabo", 40, 0), ("rtp", 40, 0)])
def test_single_process_multi_gpu_matches_oracle(cuda_lib, kind, steps, flags):
    """MaviParams.n_gpus: the plain mavi_upload_state / mavi_step / mavi_download_state calls of ONE
    process drive all GPUs; results against the single-domain oracle, cells bit-exact, time info, energies."""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    G = 4 if n >= 4 else 2
    if kind in ("lj", "harm"):
        dyn = None if kind == "lj" else pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2)
        case = H.newton_case(nx=72, ny=40, dyn=dyn, wall="periodic", jitter=0.2, vmax=3.0, dt=0.002)
        mkdev = lambda: pkg.CUDADevice(n_gpus=G, flags=flags)  # noqa: E731
    else:
        case = H.sp_case(kind, nx=64, ny=40, rot_diff=0.05, jitter=0.6 if kind == "rtp" else 0.9)
        mkdev = lambda: pkg.CUDADevice(n_gpus=G, flags=flags, rng_mode="host_noise")  # noqa: E731
    ic = case["int_cfg"]
    case["int_cfg"] = pkg.IntCfg(dt=ic.dt, chunks_cfg=ic.chunks_cfg, device=mkdev())
    g, o = H.make_gpu(case), H.make_oracle(case)
    npart = len(g.state.pos)
    cg, ng = g.download_cells()
    co, no = o.download_cells()
    assert np.array_equal(cg, co) and np.array_equal(ng, no)
    noise = None
    rng = np.random.default_rng(5)
    if kind == "szabo":
        noise = rng.standard_normal((steps, npart))
    elif kind == "rtp":
        noise = rng.random((steps, 2 * npart))
        noise[:, 0::2] *= 0.02
    g.step(steps, noise)
    o.step(steps, noise)
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert H.rel_err(g.state.second, o.second()) < 1e-10
    assert H.rel_err(g.get_forces(), o.get_forces()) < 1e-9
    assert g.time_info.num_steps == steps and g.time_info.time == o.time()[1]
    if kind in ("lj", "harm"):
        g.update_chunks()   # the reference's Chunks are stale after a step (binned at its start): re-bin both sides
        o.update_chunks()
        cg, ng = g.download_cells()
        co, no = o.download_cells()
        assert np.array_equal(cg, co) and np.array_equal(ng, no)
        sg, ig = g.download_cell_lists()
        so, io = o.download_cell_lists()
        assert np.array_equal(sg, so) and np.array_equal(ig, io)
        ke, _ = g.energies(want_pe=False)
        assert abs(ke - o.energies()[0]) <= 1e-12 * abs(ke)
        c = g.counters()
        assert c["steps"] == steps and c["rebinned"] > 0 and c["emigrants"] > 0   # particles did cross slab boundaries
    # the host edits the state and uploads again (GUI / checkpoint hook): same handle, new partition
    g.state.pos[:] = o.pos()
    g.state.second[:] = o.second()
    g.upload_state()
    g.step(5, None if noise is None else noise[:5])
    o.step(5, None if noise is None else noise[:5])
    g.sync_to_host()
    assert np.abs(g.state.pos - o.pos()).max() / case["geom"].length < 1e-12
    assert g.launch_count() > 0
    g.close()
