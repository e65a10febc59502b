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


def test_parity_gate_helper_detects_mismatches():
    """bench.compare_with_cpu: identical counts/scores pass; a single differing count, or a score off by more than 1e-5
    relative, is reported (bench.py then exits 1)."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeCpu:
        def __init__(self, counts, scores):
            self.counts, self.scores = counts, scores

        def run(self, rows, processes=1):
            return 0.5, self.counts, self.scores

    counts = np.arange(24, dtype=np.int64).reshape(3, 4, 2)
    scores = np.array([0.25, 0.5, 0.75])
    rows = np.zeros((3, 9))
    dt, n, bad, ident = bench.compare_with_cpu(FakeCpu(counts, scores), rows, counts.copy(), scores.copy(), 1)
    assert n == 3 and bad == [] and ident
    c2 = counts.copy()
    c2[1, 2, 0] += 1
    _, _, bad, _ = bench.compare_with_cpu(FakeCpu(counts, scores), rows, c2, scores.copy(), 1)
    assert len(bad) == 1 and bad[0].startswith("counts[1]")
    s2 = scores.copy()
    s2[2] *= 1 + 3e-5
    _, _, bad, ident = bench.compare_with_cpu(FakeCpu(counts, scores), rows, counts.copy(), s2, 1)
    assert len(bad) == 1 and bad[0].startswith("score[2]") and not ident
    s3 = scores.copy()
    s3[0] *= 1 + 1e-7                                        # inside the tolerance: no mismatch, but not bit-identical
    _, _, bad, ident = bench.compare_with_cpu(FakeCpu(counts, scores), rows, counts.copy(), s3, 1)
    assert bad == [] and not ident
