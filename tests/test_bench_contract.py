"""bench.py's JSON contract, exercised on the CPU through the reference arm on a tiny workload."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "64", "--mask", "128",
           "--steps", "1", "--warmup", "1", "--cpu-sample", "2"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "candidates/s" and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in j["config"]
    # the unmodified reference functions (baseline/_ref, installed where /root/reference exists) -- the port only without them
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "utils", "projection_utils.py"))
    assert j["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert j["cpu_baseline"]["blas_env"]["OPENBLAS_NUM_THREADS"] == "1" and j["product_library_loaded"] is False


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
