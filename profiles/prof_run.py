"""Short workload for ncu captures: python profiles/prof_run.py [nx] [steps] [dyn]  (the bench workload, few steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as entry

pkg = entry.load_package()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.lj_workload(pkg, nx, nx)
s = pkg.System(state=pkg.SecondLawState(pos=w["pos"], vel=w["vel"]), space_cfg=w["space"], dynamic_cfg=w["dyn"], int_cfg=w["int_cfg"])
s.step(steps)
s.sync()
print("done", s.launch_count())
