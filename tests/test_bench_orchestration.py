"""bench.py's orchestration (argument handling, side measurements, JSON line) executed end to end on a CPU box with
torch.cuda and the device System mocked (tests/bench_mock.py, run in a subprocess because it monkey-patches torch)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_is_produced_for_the_main_flag_combinations():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_mock.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert res.stdout.count("keys ok: True") == 3 and "SIDE ERROR" not in res.stdout, res.stdout


def test_reference_arm_prints_one_json_line():
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample", "120", "--cpu-small"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms describe the same workload with the same `config` (the driver compares them)
    import bench
    assert d["config"] == bench.workload_config(4000, 4000, 1)
