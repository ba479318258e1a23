"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/mavi.h declares."""
import ctypes
import os
import re

import __graft_entry__ as entry

ROOT = entry.ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mavi.h")).read()
    return sorted(set(re.findall(r"\b(mavi_[a-z0-9_]+)\s*\(", text)))


def test_build_and_exports(mavi):
    path = mavi.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mavi.h but not exported"
    # the binding table covers exactly the declared symbols
    assert sorted(mavi.capi.SIGNATURES) == declared
    assert lib.mavi_abi_version() == 1


def test_struct_layout_matches_header(mavi, tmp_path):
    """sizeof/offsetof of the flat POD agree between the C header and the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include "mavi.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(MaviParams),sizeof(MaviSpace),sizeof(MaviRingsParams),offsetof(MaviParams,rings),'
                   'offsetof(MaviParams,dt),offsetof(MaviParams,stream));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = list(map(int, subprocess.check_output([str(exe)]).split()))
    c = mavi.capi
    want = [ctypes.sizeof(c.MaviParams), ctypes.sizeof(c.MaviSpace), ctypes.sizeof(c.MaviRingsParams),
            c.MaviParams.rings.offset, c.MaviParams.dt.offset, c.MaviParams.stream.offset]
    assert got == want


def test_no_cpu_fallback_in_product():
    """The product package never imports the oracle, and the native library is not linked against it."""
    pkg_dir = entry.PKG_DIR
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libmavi_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_missing_library_fails_loudly(mavi, tmp_path):
    import pytest
    with pytest.raises(FileNotFoundError):
        mavi.capi.load_library(str(tmp_path / "nope.so"))


def test_both_arithmetic_builds_are_linked(mavi):
    """libmavi_cuda.so carries the Float64 and the Float32 build of every per-type entry point (csrc/api_decl.inc)."""
    import subprocess
    path = mavi.build_library()
    syms = subprocess.check_output(["nm", "-C", "--defined-only", path], text=True)
    decl = open(os.path.join(entry.PKG_DIR, "csrc", "api_decl.inc")).read()
    names = re.findall(r"\b(api_[a-z0-9_]+)\(", decl)
    assert len(names) == len(_declared_symbols())
    for ns in ("mavi_f64", "mavi_f32"):
        for n in names:
            assert f"{ns}::{n}(" in syms, f"{ns}::{n} missing"


def test_create_without_a_device_fails_loudly(mavi):
    """No CPU fallback: on a box without a CUDA device mavi_create returns MAVI_ERR_CUDA (both dtypes) with a message,
    and the handle can still be queried and destroyed.  (On a GPU box this test is a no-op.)"""
    import numpy as np
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    for T in (np.float64, np.float32):
        pos, geom = mavi.rectangular_grid(4, 4, 0.4, 0.5)
        st = mavi.SecondLawState(pos=pos.astype(T), vel=np.zeros_like(pos, dtype=T))
        with pytest.raises(mavi.MaviError) as e:
            mavi.System(state=st, space_cfg=mavi.SpaceCfg(wall_type=mavi.PeriodicWalls(), geometry_cfg=geom),
                        dynamic_cfg=mavi.LenJonesCfg(sigma=1.0, epsilon=1.0),
                        int_cfg=mavi.IntCfg(dt=0.001, chunks_cfg=mavi.ChunksCfg(num_cols=3, num_rows=3)))
        assert e.value.status == mavi.capi.ERR_CUDA and "no CPU fallback" in str(e.value)
