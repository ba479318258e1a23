"""Regenerates tests/golden/oracle_snapshots.npz:  python tests/golden/make_golden.py

What these are — and are not.  The reference's own golden vectors for this path (particle neighbour lists after
t = 10, test/tests_rings/runtests.jl:17-34) depend on Julia's MersenneTwister streams, and Julia cannot run in this
image, so NO fixture here comes from a run of the reference.  The snapshots below are outputs of the CPU oracle
(oracle/mavi_oracle.c) on the seeded cases of tests/helpers.py, frozen at the commit where the oracle agreed with the
independent restatements of tests/test_oracle_kat.py and tests/test_oracle_steps.py.  They pin the ORACLE against silent
drift (tests/test_snapshots.py), nothing more; tolerances are 1e-12 because libm's sin/cos/asin may differ in the last ulp
between machines.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

pkg = H.pkg


def cases():
    harm = pkg.HarmTruncCfg(k_rep=10.0, k_atr=3.0, dist_eq=1.0, dist_max=1.2)
    yield "c1_quick_start", H.newton_case(nx=10, ny=10, wall="rigid", chunks=False, dt=0.01, jitter=0.0), 100, None
    yield "lj_periodic_chunks", H.newton_case(nx=20, ny=18, wall="periodic", jitter=0.2, vmax=2.0, dt=0.002), 60, None
    yield "harm_rigid_chunks", H.newton_case(nx=20, ny=18, dyn=harm, wall="rigid", jitter=0.2, vmax=2.0, dt=0.002), 60, None
    sz = H.sp_case("szabo", nx=16, ny=14)
    yield "szabo_host_noise", sz, 30, np.random.default_rng(101).standard_normal((30, 16 * 14))
    rt = H.sp_case("rtp", nx=16, ny=14, jitter=0.6)
    yield "rtp_host_noise", rt, 30, np.random.default_rng(102).random((30, 16 * 14, 2)) * np.array([0.02, 1.0])
    for kind, n in (("normal", 6), ("types", 5)):
        rc = H.rings_case(kind, n, n)
        yield f"rings_{kind}", rc, 50, np.random.default_rng(103).standard_normal((50, n * n))


def snapshot(case, steps, noise):
    o = H.make_oracle(case)
    o.step(steps, noise)
    out = {"pos": o.pos(), "second": o.second(), "forces": o.get_forces()}
    if case["int_cfg"].chunks_cfg is not None:
        o.update_chunks()
        out["cells"] = o.download_cells()[0]
    if "num_rings" in case:
        areas, cms, _ = o.rings_info()
        out["areas"], out["cms"] = areas, cms
    return out


def main():
    blob = {}
    for name, case, steps, noise in cases():
        for k, v in snapshot(case, steps, noise).items():
            blob[f"{name}/{k}"] = v
    path = os.path.join(HERE, "oracle_snapshots.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
