import hashlib

import numpy as np

GROUP_JOBS = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
              (["small_minarets"], 90), (["dome"], 90)]                       # notebook 1, cell 7
PART_SYMMETRY = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
EXTRUSION_DEPTHS = {"main_door": 20, "windows": 10}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape)


def row_to_args(row, dtype=np.float64):
    """9-vector -> (cam_pos, target, f, cx, cy) typed the way the reference receives them."""
    cp, tg = np.asarray(row[0:3], dtype=dtype), np.asarray(row[3:6], dtype=dtype)
    if dtype == np.float32:
        return cp, tg, np.float32(row[6]), np.float32(row[7]), np.float32(row[8])
    return cp, tg, float(row[6]), float(row[7]), float(row[8])


class OracleScorer:
    """Same interface as CandidateScorer.score, computed by the oracle (joint mode, float64)."""

    def __init__(self, orc, grid, image, part_colors, parts):
        self.orc, self.parts = orc, list(parts)
        self.H, self.W = image.shape[:2]
        self.pts, self.cols = orc.get_voxel_points_by_parts(grid, part_colors, self.parts)
        self.seg = orc.mask_parts_from_image(image, part_colors, self.parts)
        self.sel = {p: part_colors[p] for p in self.parts}

    def score(self, cand):
        scores, counts = [], []
        for row in np.asarray(cand, dtype=np.float64).reshape(-1, 9):
            s, inter, uni = self.orc.score_candidate(self.pts, self.cols, self.seg, self.sel,
                                                     {"cam_pos": row[0:3], "target": row[3:6], "f": row[6], "cx": row[7],
                                                      "cy": row[8]}, self.H, self.W)
            scores.append(s)
            counts.append(np.stack([inter, uni], 1))
        scores = np.array(scores)
        return scores, np.array(counts), (int(np.argmax(scores)) if len(scores) else -1)


def drive_aligner(aligner):
    """The button sequence recorded in tests/golden/aligner_golden.npz (make_golden.aligner_golden)."""
    snap = lambda: np.array([aligner.sliders[k] for k in ["cam_x", "cam_y", "cam_z", "target_x", "target_y", "target_z",
                                                          "f", "cx", "cy"]])
    out = {}
    np.random.seed(1234)
    aligner.run_random(12)
    out["after_random"] = snap()
    aligner.run_coord(4)
    out["after_coord"] = snap()
    aligner.run_powell(2)
    out["after_powell"] = snap()
    s = aligner.save()
    out["saved"] = np.array([*s["cam_pos"], *s["target"], s["f"], s["cx"], s["cy"]])
    return out
