"""Run as a script by test_fullsize_gpu.py (the library reads its batching knobs once per process): scores 384 candidates
of bench.py's own configuration (512^3 synthetic monument, all parts, 1024x1024 mask -- three z-buffer batches of 128,
both double-buffer slots) and writes counts + scores to the .npz named on the command line."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from conftest import pkg                      # noqa: E402


def main(out_path):
    syn, cfg, ce = pkg("synthetic"), pkg("utils.config"), pkg("utils.camera_estimation")
    N, H, W = 512, 1024, 1024
    dev = torch.device("cuda")
    rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
    base = syn.base_camera(N, H, W)
    hidden = base + np.array([3.0, -2.0, 5.0, 1.0, 2.0, -3.0, 4.0, 1.5, -2.5])
    full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
    gt = full.render(ce.row_to_params(hidden))
    del full
    scorer = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
    cand = syn.candidates(base, 384)
    rng = np.random.default_rng(5)
    far = rng.choice(384, 40, replace=False)                 # some cameras see only part of the object, or none
    cand[far, 7] += rng.uniform(-1.5 * W, 1.5 * W, 40)
    cand[far, 8] += rng.uniform(-1.0 * H, 1.0 * H, 40)
    out = {}
    for rep in range(2):                                     # the second sweep re-uses the cleared z-buffers
        scores, counts, best = scorer.score(cand)
        out[f"scores{rep}"], out[f"counts{rep}"], out[f"best{rep}"] = scores, counts, best
    np.savez(out_path, segs=0 if scorer.segs is None else int(scorer.segs.shape[0]), **out)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
