"""Drop-in for the scoring part of the reference's utils/camera_estimation.py.

Reference functions covered: compute_partwise_iou (:770-787), the `evaluate` / `quick_overlay_proj`
closures of launch_smart_aligner (:552-572, :597-603) as a batched scorer, the three optimiser loops
(:606-725) as a headless `SmartAligner`, and the scoring core of visualize_voxel_projection_iou
(:381-403, :433-447).  The ipywidgets UI itself is not reproduced.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from .camera_geometry import candidate_row
from .mask_utils import image_labels
from .voxel_utils import device_points_by_parts

_STEP_SIZES = np.array([50, 50, 100, 50, 50, 100, 50, 20, 20], dtype=np.float64)   # camera_estimation.py:611-617


# --------------------------------------------------------------------------------------------
# compute_partwise_iou
# --------------------------------------------------------------------------------------------
def iou_from_counts(inter, union):
    """camera_estimation.py:783: inter / union, or 0.0 when the union is empty."""
    return inter / union if union > 0 else 0.0


def compute_partwise_iou(proj_mask, gt_mask, part_colors, device=None):
    """camera_estimation.py:770-787 -> (dict part -> IoU, mean IoU over all parts).

    A part whose colour appears in neither image contributes 0.0 to the mean, as in the reference.
    """
    dev = nv.require_cuda(device)
    names = list(part_colors.keys())
    proj = nv.to_device(proj_mask, torch.uint8, dev).reshape(-1, 3)
    gt = nv.to_device(gt_mask, torch.uint8, dev).reshape(-1, 3)
    if proj.shape != gt.shape:
        raise ValueError(f"operands could not be broadcast together: {tuple(proj.shape)} vs {tuple(gt.shape)}")
    if not names:
        return {}, np.mean([])
    rgb = nv.palette_tensor([part_colors[n] for n in names], dev)
    counts = eng.partwise_counts_rgb(proj, gt, rgb).cpu().numpy()
    per_part, total = {}, []
    for name, (inter, union) in zip(names, counts):
        iou = iou_from_counts(inter, union)
        per_part[name] = iou
        total.append(iou)
    return per_part, np.mean(total)


# --------------------------------------------------------------------------------------------
# batched candidate scoring
# --------------------------------------------------------------------------------------------
def params_to_row(p, dtype=np.float64) -> np.ndarray:
    """cam_params dict (camera_estimation.py:97-103) -> 9-vector [cam_pos, target, f, cx, cy]."""
    return candidate_row(p["cam_pos"], p["target"], p["f"], p["cx"], p["cy"], dtype)


def row_to_params(row, H=None, W=None) -> dict:
    p = {"cam_pos": np.array(row[0:3]), "target": np.array(row[3:6]),
         "f": float(row[6]), "cx": float(row[7]), "cy": float(row[8])}
    if H is not None:
        p["H"], p["W"] = H, W
    return p


def random_candidates(base_row, K, rng, include_base=True) -> np.ndarray:
    """K candidates = base (index 0) + perturbations base + U(-1,1) * step sizes, drawn per trial in
    the order cam_pos(3), target(3), f, cx, cy (camera_estimation.py:619-625)."""
    base_row = np.asarray(base_row, dtype=np.float64)
    out = np.empty((K, 9), np.float64)
    start = 0
    if include_base and K > 0:
        out[0] = base_row
        start = 1
    out[start:] = base_row + rng.uniform(-1.0, 1.0, size=(K - start, 9)) * _STEP_SIZES
    return out


class CandidateScorer:
    """Device-resident state for scoring many cameras against one (grid, image, parts) triple.

    mode="joint"    : the aligner's `evaluate` -- the selected parts are rendered into ONE label image
                      (last point in index order wins a pixel) and compared with the image restricted
                      to the selected colours (camera_estimation.py:488-500, 597-603).
    mode="per_part" : every part rendered on its own, plus the combined binary IoU against
                      any(image != background) (camera_estimation.py:381-403, 433-447).
    """

    def __init__(self, voxel_grid, image, part_colors, parts, *, mode="joint", device=None,
                 dtype=np.float64, background=None):
        self.device = nv.require_cuda(device)
        self.mode = {"joint": nv.MODE_JOINT, "per_part": nv.MODE_PER_PART}[mode]
        self.parts = list(parts)
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("dtype must be float32 or float64")
        with torch.cuda.device(self.device):
            self.pts, self.pt_label, self.colours, self.label_of = device_points_by_parts(
                voxel_grid, part_colors, self.parts, self.device)
            if any(c == (0, 0, 0) for c in self.colours):
                raise ValueError("a part colour of (0,0,0) is indistinguishable from empty pixels")
            if len(self.colours) > nv.MAX_PARTS:
                raise ValueError(f"at most {nv.MAX_PARTS} distinct part colours")
            img = nv.to_device(image, torch.uint8, self.device)
            self.H, self.W = int(img.shape[0]), int(img.shape[1])
            self.gt_label = image_labels(img, self.colours, self.device)
            self.gt_any = None
            if self.mode == nv.MODE_PER_PART:
                bg = part_colors.get("background", (0, 0, 0)) if background is None else background
                self.gt_any = (image_labels(img, [tuple(int(v) for v in bg)], self.device) == 0).to(torch.uint8)
            self.pal = nv.palette_tensor(self.colours, self.device)
            self.workspace = eng.SweepWorkspace(self.device)
        self.P = len(self.colours)
        self._cols = [self.label_of[p] - 1 for p in self.parts]
        self._dedup = len(self.colours) != len(self.parts)

    @property
    def n_points(self) -> int:
        return int(self.pts.shape[0])

    def score_device(self, cand: torch.Tensor, want_best=True):
        """cand: (K,9) device tensor of self.dtype.  Returns device tensors (counts (K,rows,2), scores (K),
        best (2)) with rows indexed by DISTINCT colour label (see `label_of`)."""
        with torch.cuda.device(self.device):
            return eng.sweep(self.pts, self.pt_label, cand, self.gt_label, self.H, self.W, self.P, self.mode,
                             gt_any=self.gt_any, workspace=self.workspace, want_best=want_best)

    def score(self, candidates):
        """candidates: (K,9) array-like [cam_pos, target, f, cx, cy].
        Returns (scores (K) f64, counts (K, len(parts)[+1], 2) i64, best_index) as NumPy / int."""
        cand = np.ascontiguousarray(np.asarray(candidates, dtype=self.dtype).reshape(-1, 9))
        K = cand.shape[0]
        if K == 0:
            rows = len(self.parts) + (1 if self.mode == nv.MODE_PER_PART else 0)
            return np.zeros(0), np.zeros((0, rows, 2), np.int64), -1
        if self.P == 0:
            raise ValueError("no parts selected")
        counts_t, scores_t, best_t = self.score_device(torch.from_numpy(cand).to(self.device))
        counts = counts_t.cpu().numpy()
        scores = scores_t.cpu().numpy()
        best = int(best_t[0].item())
        cols = list(self._cols) + ([self.P] if self.mode == nv.MODE_PER_PART else [])
        counts = counts[:, cols, :]
        if self._dedup:          # repeated colours: the mean runs over parts, not over distinct colours
            inter = counts[:, :len(self.parts), 0].astype(np.float64)
            union = counts[:, :len(self.parts), 1].astype(np.float64)
            iou = np.divide(inter, union, out=np.zeros_like(inter), where=union > 0)
            scores = np.array([np.mean(row) for row in iou])
            best = int(np.argmax(scores))
        return scores, counts, best

    def render(self, params) -> np.ndarray:
        """The label image of one camera as RGB (what quick_overlay_proj shows)."""
        with torch.cuda.device(self.device):
            cand = torch.from_numpy(params_to_row(params, self.dtype)[None]).to(self.device)
            cams = eng.setup_cameras(cand)
            zbuf = eng.splat(self.pts, None, cams, self.H, self.W, nv.MODE_JOINT)
            cols = eng.labels_to_rgb(self.pt_label, eng.make_lut(self.pal))
            return eng.resolve_rgb(zbuf[0], cols).cpu().numpy()


def score_camera_candidates(voxel_grid, image, part_colors, parts, candidates, *, mode="joint", device=None,
                            dtype=np.float64):
    """Batched `evaluate`: score K candidate cameras [cam_pos, target, f, cx, cy] against `image`.

    Returns (scores (K) float64 = mean per-part IoU, counts (K,P,2) int64 = (inter, union) per part in
    the order of `parts`, best_index = first candidate with the greatest score)."""
    return CandidateScorer(voxel_grid, image, part_colors, parts, mode=mode, device=device,
                           dtype=dtype).score(candidates)
