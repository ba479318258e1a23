"""mavi.jl_b200 — B200-native device backend for the per-step hot path of Mavi.jl.

The product is `csrc/libmavi_cuda.so` (CUDA, sm_100a, plain C ABI: include/mavi.h).  This package is the host-side
mirror of the reference's Julia API for that path (System, State, SpaceCfg, DynamicCfg, IntCfg, step functions,
run_system, quantities) used by tests and bench.py because Julia is not installed here; `julia/MaviCUDA.jl` is the
Julia glue against the same ABI.  The directory name contains a dot, so it is loaded under the module name
`mavi_jl_b200` by `__graft_entry__.load_package()`.
"""
from . import capi
from .capi import MaviError, build_library, load_library
from .configs import *  # noqa: F401,F403
from .configs import particle_radius
from .init_states import random_vel, rectangular_grid
from .integration import (calc_forces, get_step_function, newton_step, rings_step, rtp_step, run_system, szabo_step,
                          update_chunks)
from .quantities import kinetic_energy, potential_energy
from .states import ActiveState, SecondLawState, SelfPropelledState
from .systems import System, TimeInfo, get_forces
