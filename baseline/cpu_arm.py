"""CPU arm of bench.py: the reference's own scoring path on the host cores, candidate-parallel.

Used (1) as the `--impl reference` arm, (2) as the `cpu_baseline` of the GPU arm and (3) as the checker of the GPU arm's
parity gate (counts and scores of sampled candidates).  The functions timed are the UNMODIFIED reference functions from
baseline/_ref (kind "reference", see baseline/harness.py); only if that directory is absent does it fall back to the
NumPy port under oracle/ (kind "port").  Worker processes run with one BLAS thread each (recorded in the result).
Nothing of the product package is imported here.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_W = {}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _limit_blas_threads():
    try:
        from threadpoolctl import threadpool_limits
        _W["_limit"] = threadpool_limits(1)
    except Exception:
        pass


def _load_backend():
    """(kind, counts_fn): counts_fn(pts, cols, seg, sel, row, H, W) -> ([(inter, union)] per part, mean IoU)."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    try:
        import harness
        if harness.available():
            ref = harness.Reference()
            return "reference", ref.counts, ref
    except Exception as exc:                                          # fall back to the port, say why
        print(f"cpu_arm: reference unavailable ({exc!r}); using the NumPy port", file=sys.stderr)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import np_port
    return "port", lambda pts, cols, seg, sel, row, H, W: np_port.evaluate(pts, cols, seg, sel, row, H, W), None


def _work(rows):
    out = []
    for row in rows:
        counts, mean = _W["fn"](_W["pts"], _W["cols"], _W["seg"], _W["sel"], row, _W["H"], _W["W"])
        out.append((counts, float(mean)))
    return out


class CpuScorer:
    """`evaluate` of launch_smart_aligner (camera_estimation.py:597-603) for rows of candidates, on `processes` host
    processes (fork: the point list is shared copy-on-write)."""

    def __init__(self, pts, cols, gt_image, part_names, part_colors=None):
        self.kind, fn, ref = _load_backend()
        colors = ref.PART_COLORS if ref is not None else part_colors
        if colors is None:
            raise ValueError("part_colors needed with the port backend")
        self.sel = {p: colors[p] for p in part_names}
        if ref is not None:
            seg = ref.mask_parts_from_image(gt_image, colors, list(part_names))        # mask_utils.py:89-97
        else:
            seg = np.zeros_like(gt_image)
            for c in self.sel.values():
                seg[np.all(gt_image == c, axis=-1)] = c
        self.H, self.W = gt_image.shape[:2]
        _W.update(fn=fn, pts=pts, cols=cols, seg=seg, sel=self.sel, H=self.H, W=self.W)
        self.ref = ref

    def render(self, pts, cols, row):
        """project_colored_voxels of the backend (projection_utils.py:5-23)."""
        if self.ref is not None:
            return self.ref.project_colored_voxels(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], self.H, self.W)
        from oracle import np_port
        return np_port.render(pts, cols, row[0:3], row[3:6], row[6], row[7], row[8], self.H, self.W)

    def run(self, cand, processes=1):
        """Returns (seconds, counts (K,P,2) int64, scores (K) float64)."""
        cand = np.asarray(cand, dtype=np.float64).reshape(-1, 9)
        if processes <= 1 or len(cand) <= 1:
            _limit_blas_threads()
            t0 = time.perf_counter()
            res = _work(cand)
            dt = time.perf_counter() - t0
        else:
            chunks = [c for c in np.array_split(cand, min(processes, len(cand))) if len(c)]
            ctx = mp.get_context("fork")
            with ctx.Pool(len(chunks), initializer=_limit_blas_threads) as pool:
                pool.map(_work, [c[:0] for c in chunks])            # start the workers outside the timed region
                t0 = time.perf_counter()
                parts = pool.map(_work, chunks)
                dt = time.perf_counter() - t0
            res = [r for p in parts for r in p]
        counts = np.array([r[0] for r in res], dtype=np.int64).reshape(len(cand), -1, 2)
        scores = np.array([r[1] for r in res], dtype=np.float64)
        return dt, counts, scores


def blas_env():
    return {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
