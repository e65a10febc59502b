"""Run as a script by test_sweep_batches_gpu.py (the library reads its batching knobs once per process): sweeps whose
candidates span many z-buffer batches -- double-buffered splat/score overlap, per-camera footprint rectangles of the
score pass -- against the oracle.  Cameras look at the object from far and near, with principal points shifted so that
the object straddles or leaves the image, and from inside the cloud (no footprint bound: whole image); image widths
with W % 4 == 0, W % 4 != 0 with H*W % 4 == 0, and odd sizes.  Exit code 0 = every count and score identical."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import pkg                      # noqa: E402
from helpers import row_to_args               # noqa: E402
from oracle import oracle                     # noqa: E402


def main():
    oracle.build()
    ce = pkg("utils.camera_estimation")
    rng = np.random.default_rng(int(os.environ.get("SWEEP_CHECK_SEED", "11")))
    names = [k for k in oracle.PART_COLORS if k != "background"]
    checked = 0
    for trial, (H, W) in enumerate(((48, 64), (30, 50), (37, 41), (64, 36))):
        for dt in (np.float64, np.float32):
            A0, A1, A2 = (int(v) for v in rng.integers(8, 20, 3))
            parts = list(rng.choice(names, size=4, replace=False))
            grid = np.zeros((A0, A1, A2, 3), np.uint8)
            lab = rng.integers(0, 7, (A0, A1, A2))
            for k, n in enumerate(parts):
                grid[lab == k + 1] = oracle.PART_COLORS[n]
            image = np.zeros((H, W, 3), np.uint8)
            ilab = rng.integers(0, 5, (H, W))
            for k, n in enumerate(parts):
                image[ilab == k + 1] = oracle.PART_COLORS[n]
            K = 40
            ctr = np.array([A2, A1, A0]) / 2
            size = float(max(A0, A1, A2))
            cand = np.empty((K, 9))
            dist = rng.choice([0.3, 2.0, 4.0, 12.0], K) * size                     # 0.3: inside / next to the cloud
            dirs = rng.normal(0, 1, (K, 3))
            dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
            cand[:, 0:3] = ctr + dirs * dist[:, None]
            cand[:, 3:6] = ctr + rng.normal(0, 1, (K, 3)) * size * 0.3
            cand[:, 6] = rng.uniform(0.5, 4.0, K) * max(H, W)
            cand[:, 7] = W / 2 + rng.choice([0.0, 0.6, -0.6, 1.5], K) * W + rng.normal(0, 2, K)   # object partly / fully outside
            cand[:, 8] = H / 2 + rng.choice([0.0, 0.6, -0.6, 1.5], K) * H + rng.normal(0, 2, K)
            cand = cand.astype(dt)
            scorer = ce.CandidateScorer(grid, image, oracle.PART_COLORS, parts, dtype=dt)
            for rep in range(2):                                                 # the second sweep re-uses the cleared z-buffers
                scores, counts, best = scorer.score(cand)
                pts, cols = oracle.get_voxel_points_by_parts(grid, oracle.PART_COLORS, parts)
                seg = oracle.mask_parts_from_image(image, oracle.PART_COLORS, parts)
                sel = {p: oracle.PART_COLORS[p] for p in parts}
                ref = []
                for k in range(K):
                    cp, tg, f, cx, cy = row_to_args(cand[k], dt)
                    s, inter, uni = oracle.score_candidate(pts, cols, seg, sel,
                                                           {"cam_pos": cp, "target": tg, "f": f, "cx": cx, "cy": cy}, H, W)
                    ref.append(s)
                    if not (np.array_equal(counts[k, :, 0], inter) and np.array_equal(counts[k, :, 1], uni) and scores[k] == s):
                        print(f"MISMATCH trial={trial} dtype={dt.__name__} rep={rep} k={k}: {counts[k].tolist()} vs "
                              f"{list(inter)} / {list(uni)}; {scores[k]} vs {s}")
                        return 1
                    checked += 1
                if best != int(np.argmax(ref)):
                    print(f"MISMATCH best trial={trial}: {best} vs {int(np.argmax(ref))}")
                    return 1
    print(f"ok: {checked} candidate scores identical")
    return 0


if __name__ == "__main__":
    sys.exit(main())
