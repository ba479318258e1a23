"""Executes bench.run_ours end to end on a CPU box with torch.cuda and the device System mocked, to catch Python-level
errors in the orchestration (NameError, bad keys, JSON serialisation)."""
import json, sys, types, io, contextlib
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch, bench
import __graft_entry__ as entry
pkg = entry.load_package()

class FakeEvent:
    def __init__(self, enable_timing=True): pass
    def record(self): pass
    def elapsed_time(self, other): return 1.0
class FakeStream: cuda_stream = 0
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda i: None
torch.cuda.current_stream = lambda: FakeStream()
torch.cuda.synchronize = lambda: None
torch.cuda.Event = FakeEvent
torch.cuda.get_device_properties = lambda i: types.SimpleNamespace(uuid="x")
_orig_tensor = torch.tensor
torch.tensor = lambda data, device=None, dtype=None: _orig_tensor(data, dtype=dtype)
torch.Tensor.pin_memory = lambda self: self

class FakeSystem:
    def __init__(self, *, state, space_cfg, dynamic_cfg, int_cfg, **kw):
        self.state, self.int_cfg = state, int_cfg; self._n = len(state.pos); self.n_launch = 0
        self._slab = bool(int_cfg.device.flags & pkg.capi.FLAG_SLAB_SELF) or int_cfg.device.world > 1
        self.local_ids = np.arange(self._n)
    def step(self, n=1, noise=None): self.n_launch += 8 * n
    def close(self): pass
    def set_profiling(self, on): pass
    def launch_count(self): return self.n_launch
    def last_step_ms(self): return [0.0, 0.003, 0.58, 0.07, 0.66]
    def counters(self): return dict(steps=self.n_launch // 8, rebinned=10 * self.n_launch, tile_movers=1, tiles_repaired=2, emigrants=0, rebuilds=0, tile_cap=96, tiles=12)
    def energies(self, pe_mode=0, want_ke=True, want_pe=True): return 1.0, None
    def get_forces(self): return np.zeros((self._n, 2))
    def download_cells(self): return np.zeros(self._n, dtype=np.int32), np.zeros(4, dtype=np.int32)
    def sync_to_host(self): return self
    def upload_state(self): pass
    def upload_local(self): pass
    def download_local(self, out=None, want_forces=True):
        n = self._n
        return out[0][:n], out[1][:n], out[2][:n], None
pkg.System = FakeSystem
import mavi_jl_b200.rings.rings as rr
rr.System = FakeSystem
pkg.load_library = lambda: None
bench.entry.load_package = lambda: pkg

for argv in (["--nx", "60", "--ny", "50", "--steps", "20", "--cpu-sample", "40", "--cpu-steps", "2", "--no-probe", "--hot-steps", "7"],
             ["--nx", "60", "--ny", "50", "--steps", "20", "--flags", "16", "--no-cpu-baseline"],
             ["--nx", "60", "--ny", "50", "--steps", "20", "--flags", "4", "--no-cpu-baseline", "--no-other-configs"]):
    sys.argv = ["bench.py"] + argv
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    out = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(out) == 1, buf.getvalue()
    d = json.loads(out[0])
    print(argv[-3:], "keys ok:", all(k in d for k in ("metric","value","roofline","cpu_baseline","e2e","gpu_launches","clocks","two_pass","float32","other_configs","parity_probe","hot","rebinned_per_step")))
    for k in ("cpu_baseline", "two_pass", "float32", "other_configs", "hot"):
        v = d[k]
        if isinstance(v, dict) and ("error" in v or any(isinstance(x, dict) and "error" in x for x in v.values())):
            print("  SIDE ERROR in", k, v)
    print("  ", {k: (d[k] if not isinstance(d[k], dict) else "{...}") for k in ("value","ms_per_step","gpu_launches")}, "other:", None if d["other_configs"] is None else {k: v.get("ms_per_step", v) for k, v in d["other_configs"].items()})
