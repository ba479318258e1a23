"""The CPU oracle against its own frozen snapshots (tests/golden/oracle_snapshots.npz, written by
tests/golden/make_golden.py — read its header: these pin the oracle against drift; no fixture can come from the Julia
reference here), and, on a GPU box, the device library against the same snapshots."""
import os
import sys

import numpy as np
import pytest

import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import make_golden  # noqa: E402

SNAP = np.load(os.path.join(GOLDEN, "oracle_snapshots.npz"))
CASES = {name: (case, steps, noise) for name, case, steps, noise in make_golden.cases()}


def _check(name, got, tol_pos=1e-12, tol=1e-10):
    case = CASES[name][0]
    for k, v in got.items():
        want = SNAP[f"{name}/{k}"]
        if k == "cells":
            assert np.array_equal(v, want), (name, k)
        elif k == "pos":
            assert np.abs(v - want).max() / case["geom"].length < tol_pos, (name, k)
        else:
            assert H.rel_err(v, want) < tol, (name, k)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_its_frozen_snapshots(oracle, name):
    case, steps, noise = CASES[name]
    _check(name, make_golden.snapshot(case, steps, noise))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_frozen_snapshots(cuda_lib, name):
    case, steps, noise = CASES[name]
    g = H.make_gpu_rings(case) if "num_rings" in case else H.make_gpu(case)
    g.step(steps, noise)
    g.sync_to_host()
    got = {"pos": g.state.pos, "second": g.state.second, "forces": g.get_forces()}
    if "num_rings" in case:   # before the re-binning below: the snapshot holds the step's own areas / cms
        areas, cms, _ = g.rings_info()
        got["areas"], got["cms"] = areas, cms
    self_propelled = isinstance(case["dyn"], (H.pkg.SzaboCfg, H.pkg.RunTumbleCfg))
    if case["int_cfg"].chunks_cfg is not None and not self_propelled:  # (cells after Szabo / RTP steps: oracle-side only)
        if "num_rings" in case:
            g.update_chunks()  # Rings bin at the start of a step: re-bin the final positions, like the snapshot
        got["cells"] = g.download_cells()[0]   # core path: device cells are always those of the current positions
    _check(name, got)
