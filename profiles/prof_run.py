"""Short workloads for ncu captures: python profiles/prof_run.py [nx] [steps] [lj|lj32|ljself|szabo|rings]."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import __graft_entry__ as entry

pkg = entry.load_package()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "lj"
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # MAVI_FLAG_* (32 = legacy staging kernels)
pre = int(sys.argv[5]) if len(sys.argv) > 5 else 0     # steps before the timed region (2000: the thermalised lattice of bench.py's `hot`)
if kind in ("lj", "ljself", "lj32"):
    # ljself: the x-slab machinery on one GPU (MAVI_FLAG_SLAB_SELF), for profiling the multi-GPU step on one rank
    w = bench.lj_workload(pkg, nx, nx, cuda_device=pkg.CUDADevice(flags=flags | (pkg.capi.FLAG_SLAB_SELF if kind == "ljself" else 0)))
    T = np.float32 if kind == "lj32" else np.float64  # lj32: Float32 mode (mavi_f32 build) on the same workload
    s = pkg.System(state=pkg.SecondLawState(pos=w["pos"].astype(T), vel=w["vel"].astype(T)), space_cfg=w["space"], dynamic_cfg=w["dyn"], int_cfg=w["int_cfg"])
    n = nx * nx
elif kind == "szabo":
    # BASELINE config C3: examples/szabo.jl parameters, lattice offset 1, cells (n-1)^2, dt 0.01, Philox noise
    dyn = pkg.SzaboCfg(vo=1.0, mobility=1.0, relax_time=1.0, k_rep=10.0, k_adh=0.75, r_eq=1.0, r_max=1.1, rot_diff=0.01)
    pos, geom = pkg.rectangular_grid(nx, nx, 1.0, pkg.particle_radius(dyn))
    ang = np.random.default_rng(bench.SEED).random(nx * nx) * 2 * np.pi
    s = pkg.System(state=pkg.SelfPropelledState(pos=pos, pol_angle=ang), space_cfg=pkg.SpaceCfg(wall_type=pkg.PeriodicWalls(), geometry_cfg=geom),
                   dynamic_cfg=dyn, int_cfg=pkg.IntCfg(dt=0.01, chunks_cfg=pkg.ChunksCfg(nx - 1, nx - 1), device=pkg.CUDADevice(rng_mode="philox", flags=flags)))
    n = nx * nx
else:
    # BASELINE config C4: 400 x 250 rings x 10 particles (test/tests_rings/rings_utils.jl:35-53 parameters), two types
    import helpers as H
    case = H.rings_case("normal", 400 if nx >= 400 else nx, 250 if nx >= 400 else nx)
    from mavi_jl_b200.rings import configs as rc
    case["int_cfg"] = rc.RingsIntCfg(dt=0.01, p_chunks_cfg=case["int_cfg"].chunks_cfg, device=pkg.CUDADevice(rng_mode="philox"))
    s = H.make_gpu_rings(case)
    n = case["num_rings"] * 10
s.step(2 + pre)
s.sync()
t0 = time.perf_counter()
s.step(steps)
s.sync()
dt = time.perf_counter() - t0
print(f"done {kind} n={n} steps={steps} {1e3 * dt / steps:.3f} ms/step {n * steps / dt / 1e9:.3f} G particle-steps/s launches={s.launch_count()}")
