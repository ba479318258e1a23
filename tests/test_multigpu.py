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
