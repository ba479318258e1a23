"""Worker of the multi-GPU parity test: torchrun --nproc-per-node N tests/mgpu_worker.py <kind> <steps>.

Every rank owns an x-slab of ONE global periodic system, steps it on its GPU (halo + migration over NCCL inside
libmavi_cuda.so), rank 0 gathers the result and compares it with the single-domain CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as H  # noqa: E402

pkg = H.pkg
from mavi_jl_b200 import slabs  # noqa: E402


def main():
    kind, steps = sys.argv[1], int(sys.argv[2])
    base_flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local_rank)
    if kind == "lj":
        case = H.newton_case(nx=72, ny=40, wall="periodic", jitter=0.2, vmax=3.0, dt=0.002)
    elif kind == "harm":
        case = H.newton_case(nx=72, ny=40, dyn=pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2),
                             wall="periodic", jitter=0.2, vmax=3.0, dt=0.002)
    else:
        case = H.sp_case("szabo", nx=64, ny=40, rot_diff=0.0)
    st0 = case["mk"]()
    ccfg = case["int_cfg"].chunks_cfg
    owner = slabs.partition(st0.pos, case["geom"], ccfg.num_cols, world)
    mine = np.flatnonzero(owner == rank)
    uid = [slabs.nccl_unique_id() if rank == 0 else None, slabs.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dev = pkg.CUDADevice(device=local_rank, rank=rank, world=world, nccl_unique_id=uid[0], n_global=len(st0.pos),
                         rng_mode="host_noise", flags=base_flags)
    int_cfg = pkg.IntCfg(dt=case["int_cfg"].dt, chunks_cfg=ccfg, device=dev)
    if kind in ("lj", "harm"):
        state = pkg.SecondLawState(pos=st0.pos[mine], vel=st0.vel[mine])
    else:
        state = pkg.SelfPropelledState(pos=st0.pos[mine], pol_angle=st0.pol_angle[mine])
    state.ids = mine
    system = pkg.System(state=state, space_cfg=case["space"], dynamic_cfg=case["dyn"], int_cfg=int_cfg)
    n0 = system.local_count()
    system.step(steps)
    ids, pos, second, forces = system.download_local()
    same = True
    if kind in ("lj", "harm"):
        # force carry (default) vs the full first pass every step: bit-identical, also across slab boundaries
        dev2 = pkg.CUDADevice(device=local_rank, rank=rank, world=world, nccl_unique_id=uid[1], n_global=len(st0.pos),
                              rng_mode="host_noise", flags=base_flags | pkg.capi.FLAG_NO_FORCE_CARRY)
        st2 = pkg.SecondLawState(pos=st0.pos[mine], vel=st0.vel[mine])
        st2.ids = mine
        sys2 = pkg.System(state=st2, space_cfg=case["space"], dynamic_cfg=case["dyn"],
                          int_cfg=pkg.IntCfg(dt=case["int_cfg"].dt, chunks_cfg=ccfg, device=dev2))
        sys2.step(steps)
        ids2, pos2, second2, forces2 = sys2.download_local()
        o1, o2 = np.argsort(ids), np.argsort(ids2)
        same = (np.array_equal(ids[o1], ids2[o2]) and np.array_equal(pos[o1], pos2[o2]) and
                np.array_equal(second[o1], second2[o2]) and np.array_equal(forces[o1], forces2[o2]))
        print(f"MGPU rank {rank}: carry == no-carry bitwise: {same}")
        sys2.close()
    sames = [None] * world
    dist.all_gather_object(sames, same)
    gathered = [None] * world
    dist.gather_object((ids, pos, second, forces, n0), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        N = len(st0.pos)
        all_ids = np.concatenate([g[0] for g in gathered])
        assert len(all_ids) == N and len(np.unique(all_ids)) == N, "particles lost or duplicated by migration"
        P = np.empty((N, 2)); F = np.empty((N, 2))
        S = np.empty_like(st0.second)
        for g in gathered:
            P[g[0]], S[g[0]], F[g[0]] = g[1], g[2], g[3]
        o = H.make_oracle(case)
        o.step(steps)
        perr = np.abs(P - o.pos()).max() / case["geom"].length
        serr = H.rel_err(S, o.second())
        ferr = H.rel_err(F, o.get_forces())
        migrated = sum(abs(len(g[0]) - g[4]) for g in gathered)
        owner_now = slabs.partition(o.pos(), case["geom"], ccfg.num_cols, world)
        moved = int((owner_now != owner).sum())
        print(f"MGPU {kind} world={world} steps={steps} pos_err={perr:.3e} second_err={serr:.3e} force_err={ferr:.3e} "
              f"changed_owner={moved} count_delta={migrated}")
        ok = perr < 1e-12 and serr < 1e-10 and ferr < 1e-9 and moved > 0 and all(sames)
        print("MGPU_OK" if ok else "MGPU_FAIL")
    dist.barrier()
    system.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
