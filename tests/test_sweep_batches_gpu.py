"""Multi-batch sweeps under every batching mode of the library (knobs are read once per process, hence subprocesses):
tests/sweep_batches_check.py compares counts and scores with the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env", [
    {"P3D_MAX_BATCH": "3"},                                   # 14 batches, double-buffered, footprint rectangles
    {"P3D_MAX_BATCH": "3", "P3D_OVERLAP": "0"},               # same batches in sequence on one stream
    {"P3D_MAX_BATCH": "7", "P3D_SCORE_RECT": "0"},            # overlap with whole-image score passes
    {"P3D_MAX_BATCH": "16", "SWEEP_CHECK_SEED": "12"},        # 3 batches, the last one ragged
    {"P3D_MAX_BATCH": "3", "P3D_SEG_MIN_POINTS": "0"},        # the segment splat (ragged runs of 1..3 voxels, dead lanes)
    {"P3D_MAX_BATCH": "16", "P3D_SEG_MIN_POINTS": "0", "P3D_OVERLAP": "0", "SWEEP_CHECK_SEED": "13"},
], ids=["overlap+rect", "serial+rect", "overlap", "ragged", "segments", "segments-serial"])
def test_multi_batch_sweeps_match_oracle(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(HERE, "sweep_batches_check.py")], env=e, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical" in r.stdout
