"""Drop-in for the scoring part of the reference's utils/camera_estimation.py.

Reference functions covered: compute_partwise_iou (:770-787), the `evaluate` / `quick_overlay_proj`
closures of launch_smart_aligner (:552-572, :597-603) as a batched scorer, the three optimiser loops
(:606-725) as a headless `SmartAligner`, the scoring core of visualize_voxel_projection_iou
(:381-403, :433-447), and the initialisation chain of notebook 2:
auto_compute_initial_params_matching_bbox (:56-108), extract_minaret_voxels_by_label (:176-207),
extract_minaret_masks_by_label (:247-323), extract_top_bottom_voxel_points / _image_points (:329-344),
extract_minaret_kps_for_view (:20-50) and optimize_camera_with_keypoints (:110-170).
The ipywidgets UI itself is not reproduced.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from .camera_geometry import candidate_row
from .mask_utils import image_labels
from .voxel_utils import device_points_by_parts, grid_to_device

_STEP_SIZES = np.array([50, 50, 100, 50, 50, 100, 50, 20, 20], dtype=np.float64)   # camera_estimation.py:611-617


# --------------------------------------------------------------------------------------------
# compute_partwise_iou
# --------------------------------------------------------------------------------------------
def iou_from_counts(inter, union):
    """camera_estimation.py:783: inter / union, or 0.0 when the union is empty."""
    return inter / union if union > 0 else 0.0


@nv.on_device
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
                 dtype=np.float64, background=None, use_segments="auto"):
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
            bg = part_colors.get("background", (0, 0, 0)) if background is None else background
            self._background = tuple(int(v) for v in bg)
            self.set_image(image)
            self.pal = nv.palette_tensor(self.colours, self.device)
            self.workspace = eng.SweepWorkspace(self.device)
            # x-run segments of the point list (ascending flat index => rows are consecutive in x): the sweep's
            # segment splat evaluates the (y, z) part of the projection once per run
            # segment splat evaluates the (y, z) part of the projection once per run.  "auto": only for long lists --
            # measured against the per-point splat (cand/s): 1024^3/2048^2 4.80 k vs 3.99 k, 512^3/1024^2 38.7 k vs 35.2 k,
            # Taj all parts (12 M points) 73.7 k vs 72.6 k, but 256^3/1024^2 (2.7 M) 158 k vs 174 k and the Taj minarets
            # (0.29 M) 1.15 M vs 1.55 M (P3D_SEG_MIN_POINTS moves the threshold)
            if use_segments == "auto":
                use_segments = self.pts.shape[0] >= int(os.environ.get("P3D_SEG_MIN_POINTS", "6000000"))
            self.segs = eng.build_segments(self.pts, self.pt_label) if use_segments else None
        self.P = len(self.colours)
        self._cols = [self.label_of[p] - 1 for p in self.parts]
        self._dedup = len(self.colours) != len(self.parts)

    def set_image(self, image):
        """Score against another 2-D mask (another view of the same monument: the front and the drone mask of notebook 2)
        without rebuilding the point list and its segments: only the ground-truth label image changes."""
        with torch.cuda.device(self.device):
            img = nv.to_device(image, torch.uint8, self.device)
            self.H, self.W = int(img.shape[0]), int(img.shape[1])
            self.gt_label = image_labels(img, self.colours, self.device)
            self.gt_any = None
            if self.mode == nv.MODE_PER_PART:
                self.gt_any = (image_labels(img, [self._background], self.device) == 0).to(torch.uint8)
        return self

    @property
    def n_points(self) -> int:
        return int(self.pts.shape[0])

    def score_device(self, cand: torch.Tensor, want_best=True):
        """cand: (K,9) device tensor of self.dtype.  Returns device tensors (counts (K,rows,2), scores (K),
        best (2)) with rows indexed by DISTINCT colour label (see `label_of`)."""
        with torch.cuda.device(self.device):
            return eng.sweep(self.pts, self.pt_label, cand, self.gt_label, self.H, self.W, self.P, self.mode,
                             gt_any=self.gt_any, workspace=self.workspace, want_best=want_best, segs=self.segs)

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


# --------------------------------------------------------------------------------------------
# launch_smart_aligner without the widgets: the three optimiser buttons as methods
# --------------------------------------------------------------------------------------------
_SLIDER_KEYS = ["cam_x", "cam_y", "cam_z", "target_x", "target_y", "target_z", "f", "cx", "cy"]


class SmartAligner:
    """Headless equivalent of launch_smart_aligner (camera_estimation.py:479-768).

    Holds the same state as the widget version (nine sliders with the reference's ranges, `saved_params`, the
    original init) and exposes the button callbacks as methods: run_random (:606-650), run_coord (:652-686),
    run_powell (:688-725), save / load / reset (:728-739).  Every candidate is scored on the GPU; run_random and
    each coordinate-descent round are scored as ONE batch (the reference's selection only depends on the scores,
    strict `>`, first candidate wins ties), Powell stays sequential because scipy drives it.

    Reference quirks that are reproduced because they change the result:
      * slider values are clamped to the slider ranges whenever parameters are written back (:528-551);
      * run_coord copies the parameter dict shallowly, so a rejected `-20` step on cam_pos/target is undone by the
        following `+20` "trial", which therefore re-evaluates the current point and never tests `+20` (:661-668);
      * run_random draws from the global np.random state in the order cam_pos(3), target(3), f, cx, cy (:621-625).
    `scorer` may be any object with `.score(cand) -> (scores, counts, best)` (tests inject the oracle).
    """

    def __init__(self, voxel_grid, image, part_colors, parts_for_alignment=("plinth", "minarets"), init_params=None,
                 lock_xy_equal=False, device=None, scorer=None, verbose=True):
        if init_params is None:                                   # camera_estimation.py:525-526
            init_params = auto_compute_initial_params_matching_bbox(voxel_grid, image, part_colors,
                                                                    list(parts_for_alignment), device=device)
        self.H, self.W = int(image.shape[0]), int(image.shape[1])
        self.lock_xy_equal = bool(lock_xy_equal)
        self.verbose = verbose
        self.scorer = scorer if scorer is not None else CandidateScorer(voxel_grid, image, part_colors,
                                                                        list(parts_for_alignment), device=device)
        self._ranges = {"cam_x": (-3000, 3000), "cam_y": (-3000, 3000), "cam_z": (-4000, 4000),
                        "target_x": (-2000, 2000), "target_y": (-2000, 2000), "target_z": (-2000, 2000),
                        "f": (100, 3000), "cx": (-self.W, 2 * self.W), "cy": (-self.H, 2 * self.H)}
        self.sliders = {}
        self.set_params(init_params)
        self.orig_init = dict(init_params)
        self.saved_params = {}
        self.evaluations = 0

    # ---- slider state ------------------------------------------------------------------------
    def _set(self, key, value):
        lo, hi = self._ranges[key]
        self.sliders[key] = float(min(max(float(value), lo), hi))

    def set_params(self, p):
        for c, v in zip("xyz", p["cam_pos"]):
            self._set(f"cam_{c}", v)
        for c, v in zip("xyz", p["target"]):
            self._set(f"target_{c}", v)
        for k in ("f", "cx", "cy"):
            self._set(k, p[k])

    def get_params(self):
        cam = np.array([self.sliders[f"cam_{c}"] for c in "xyz"])
        tgt = np.array([self.sliders[f"target_{c}"] for c in "xyz"])
        if self.lock_xy_equal:
            cam[0], cam[1] = tgt[0], tgt[1]
        return {"cam_pos": cam, "target": tgt, "f": self.sliders["f"], "cx": self.sliders["cx"],
                "cy": self.sliders["cy"], "H": self.H, "W": self.W}

    # ---- scoring -------------------------------------------------------------------------------
    def iou_batch(self, param_dicts):
        rows = np.stack([params_to_row(p) for p in param_dicts])
        scores, _, _ = self.scorer.score(rows)
        self.evaluations += len(param_dicts)
        return np.asarray(scores, dtype=np.float64)

    def iou(self, p):
        return float(self.iou_batch([p])[0])

    def evaluate(self, p):
        """The `evaluate` closure (:597-603): negative mean IoU."""
        return -self.iou(p)

    # ---- optimisers ------------------------------------------------------------------------------
    def run_random(self, steps=5):
        base = self.get_params()
        trials = []
        for _ in range(int(steps)):
            t = dict(base)
            t["cam_pos"] = base["cam_pos"] + np.random.uniform(-1, 1, 3) * _STEP_SIZES[0:3]
            t["target"] = base["target"] + np.random.uniform(-1, 1, 3) * _STEP_SIZES[3:6]
            t["f"] = base["f"] + np.random.uniform(-1, 1) * _STEP_SIZES[6]
            t["cx"] = base["cx"] + np.random.uniform(-1, 1) * _STEP_SIZES[7]
            t["cy"] = base["cy"] + np.random.uniform(-1, 1) * _STEP_SIZES[8]
            if self.lock_xy_equal:
                t["cam_pos"][:2] = t["target"][:2]
            trials.append(t)
        scores = self.iou_batch([base] + trials)
        best_iou, best_p = scores[0], base
        for s, t in zip(scores[1:], trials):
            if s > best_iou:
                best_iou, best_p = s, t
        self.set_params(best_p)
        if self.verbose:
            print(f"Random Done | Best IoU: {best_iou:.4f}")
        return float(best_iou)

    def _coord_round_trials(self, cam, tgt, scal):
        """All trials of one coordinate-descent round assuming no improvement, with the reference's aliasing: the
        cam_pos/target arrays are shared by every parameter dict, so each -20 / +20 pair is applied IN PLACE (and its
        floating-point residue, if any, stays).  Returns the parameter snapshots as evaluated."""
        out = []
        for k in _SLIDER_KEYS:
            for delta in (-20, 20):
                sc = dict(scal)
                if k.startswith("cam_") and not self.lock_xy_equal:
                    cam["xyz".index(k[-1])] += delta
                elif k.startswith("target_"):
                    tgt["xyz".index(k[-1])] += delta
                    if self.lock_xy_equal and k in ("target_x", "target_y"):
                        cam["xyz".index(k[-1])] += delta
                elif k in ("f", "cx", "cy"):
                    sc[k] = sc[k] + delta
                else:
                    continue
                out.append({"cam_pos": cam.copy(), "target": tgt.copy(), "f": sc["f"], "cx": sc["cx"], "cy": sc["cy"],
                            "H": self.H, "W": self.W})
        return out

    def run_coord(self, steps=5):
        state = self.get_params()
        best_iou = self.iou(state)
        cam, tgt = state["cam_pos"], state["target"]                 # the arrays every dict of the reference shares
        scal = {k: state[k] for k in ("f", "cx", "cy")}
        for _ in range(int(steps)):
            trials = self._coord_round_trials(cam, tgt, scal)
            if not trials:
                break
            scores = self.iou_batch(trials)
            hit = next((i for i, s in enumerate(scores) if s > best_iou), None)
            if hit is not None:                                       # first improvement ends the round (:679-683)
                best_iou = float(scores[hit])
                cam, tgt = trials[hit]["cam_pos"].copy(), trials[hit]["target"].copy()
                scal = {k: trials[hit][k] for k in ("f", "cx", "cy")}
        self.set_params({"cam_pos": cam, "target": tgt, **scal})
        if self.verbose:
            print(f"Coord Descent Done | Best IoU: {best_iou:.4f}")
        return float(best_iou)

    def _to_vector(self, p):
        if self.lock_xy_equal:
            return np.array([p["cam_pos"][2], p["target"][2], p["f"], p["cx"], p["cy"]])
        return np.concatenate([p["cam_pos"], p["target"], [p["f"], p["cx"], p["cy"]]])

    def _from_vector(self, x):
        if self.lock_xy_equal:
            tx, ty = self.sliders["target_x"], self.sliders["target_y"]
            return {"cam_pos": np.array([tx, ty, x[0]]), "target": np.array([tx, ty, x[1]]), "f": x[2], "cx": x[3],
                    "cy": x[4], "H": self.H, "W": self.W}
        return {"cam_pos": x[:3], "target": x[3:6], "f": x[6], "cx": x[7], "cy": x[8], "H": self.H, "W": self.W}

    def run_powell(self, maxiter=5):
        from scipy.optimize import minimize
        x0 = self._to_vector(self.get_params())
        res = minimize(lambda x: self.evaluate(self._from_vector(x)), x0, method="Powell",
                       options={"maxiter": int(maxiter), "maxfev": int(maxiter) * 10, "xtol": 1e-3, "ftol": 1e-3,
                                "disp": bool(self.verbose)})
        p = self._from_vector(res.x)
        iou = self.iou(p)
        self.set_params(p)
        if self.verbose:
            print(f"Powell Done | Best IoU: {iou:.4f}")
        return iou

    # ---- Save / Load / Init buttons --------------------------------------------------------------
    def save(self):
        self.saved_params.clear()
        self.saved_params.update(self.get_params())
        return self.saved_params

    def load(self):
        self.set_params(self.saved_params)

    def reset(self):
        self.set_params(self.orig_init)


def launch_smart_aligner(voxel_grid, image, part_colors, parts_for_alignment=["plinth", "minarets"], init_params=None,
                         lock_xy_equal=False, device=None):
    """camera_estimation.py:479-768 without the ipywidgets UI: returns the `saved_params` dict like the reference and
    attaches the SmartAligner that fills it (`saved_params.aligner`), so a notebook can drive
    `.run_random() / .run_coord() / .run_powell() / .save()` where the buttons used to be."""
    aligner = SmartAligner(voxel_grid, image, part_colors, parts_for_alignment, init_params, lock_xy_equal, device)

    class _Saved(dict):
        pass

    saved = _Saved()
    aligner.saved_params = saved
    saved.aligner = aligner
    return saved


def partwise_projection_iou(voxel_grid, part_colors, image, cam_params, device=None):
    """Scoring core of visualize_voxel_projection_iou (camera_estimation.py:381-403, 433-447): every part of
    `part_colors` rendered on its own and compared with its colour in `image`, plus the combined binary IoU against
    any(image != background).  Returns (dict part -> IoU for the parts that have voxels, combined IoU, counts)."""
    parts = list(part_colors.keys())
    scorer = CandidateScorer(voxel_grid, image, part_colors, parts, mode="per_part", device=device)
    _, counts, _ = scorer.score(params_to_row(cam_params)[None])
    labels_present = set(torch.unique(scorer.pt_label).cpu().tolist())
    per_part = {}
    for i, name in enumerate(parts):
        if scorer.label_of[name] in labels_present:            # the reference skips parts without voxels (:383-385)
            per_part[name] = iou_from_counts(counts[0, i, 0], counts[0, i, 1])
    combined = iou_from_counts(counts[0, -1, 0], counts[0, -1, 1])
    return per_part, combined, counts[0]


def visualize_voxel_projection_iou(voxel_grid, part_colors, image, cam_params, mode="part_on_whole", save=False,
                                   save_root="visualisation", device=None):
    """camera_estimation.py:346-477.  The figures are out of scope; the numbers the figures are titled with are
    printed instead: per-part IoU for the `part_*` modes, the combined binary IoU for `whole_on_whole`."""
    per_part, combined, _ = partwise_projection_iou(voxel_grid, part_colors, image, cam_params, device=device)
    if mode in ("part_on_whole", "part_on_part"):
        for name, iou in per_part.items():
            print(f"{name} | IoU: {iou:.3f}")
    if mode == "whole_on_whole":
        print("Visualizing combined binary projection vs. binary ground-truth...")
        print(f"Combined Binary | IoU: {combined:.3f}")
    return per_part, combined


# --------------------------------------------------------------------------------------------
# initialisation chain of notebook 2 (camera_estimation.py:20-344)
# --------------------------------------------------------------------------------------------
def _mask_bbox_2d(mask_u8: torch.Tensor):
    """(x_min, y_min, x_max, y_max, count) of the non-zero pixels of an (H,W) u8 device mask."""
    H, W = mask_u8.shape
    lab = mask_u8.to(torch.int32).contiguous()
    bbox = torch.empty((1, 6), dtype=torch.int32, device=mask_u8.device)
    sums = torch.empty((1, 4), dtype=torch.int64, device=mask_u8.device)
    nv.check(nv.lib.p3d_component_stats(nv.ptr(lab), 1, H, W, 1, nv.ptr(bbox), nv.ptr(sums), nv.stream_ptr()),
             "p3d_component_stats")
    eng._launched(2)
    b, n = bbox.cpu().numpy()[0], int(sums[0, 0].item())
    return int(b[2]), int(b[1]), int(b[5]), int(b[4]), n


def auto_compute_initial_params_matching_bbox(voxel_grid, image, part_colors, parts_for_alignment, fov_deg=30,
                                              device=None):
    """camera_estimation.py:56-108: camera on the -z side of the selected parts' bounding box, focal length scaled so
    that the box diagonal matches the diagonal of the mask's bounding box.  The two bounding boxes are reduced on the
    GPU; the scalar arithmetic keeps the reference's NumPy dtypes (float32 voxel box, float64 focal length)."""
    dev = nv.require_cuda(device)
    with torch.cuda.device(dev):
        img = nv.to_device(image, torch.uint8, dev)
        H_img, W_img = int(img.shape[0]), int(img.shape[1])
        pts, _, colours, _ = device_points_by_parts(voxel_grid, part_colors, list(parts_for_alignment), dev)
        if pts.shape[0] == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        box = eng.points_bbox(pts).cpu().numpy()
        x0, y0, x1, y1, n = _mask_bbox_2d((image_labels(img, colours, dev) != 0).to(torch.uint8))
        if n == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
    bbox_min, bbox_max = box[0:3].astype(np.float32), box[3:6].astype(np.float32)
    voxel_center = (bbox_min + bbox_max) / 2
    voxel_size = np.linalg.norm(bbox_max - bbox_min)
    img_bbox_width = np.linalg.norm(np.array([x1, y1]) - np.array([x0, y0]))
    cam_pos = voxel_center + np.array([0, 0, -voxel_size * 2.0])
    f = H_img / (2 * np.tan(np.deg2rad(fov_deg) / 2))
    scale_factor = img_bbox_width / ((voxel_size * f) / (voxel_size * 2.0))
    f_adjusted = f * scale_factor
    print(f"Estimated scale factor: {scale_factor:.4f}")
    print(f"Adjusted focal length: {f_adjusted:.2f}")
    return {"cam_pos": cam_pos, "target": voxel_center, "f": f_adjusted, "cx": W_img / 2, "cy": H_img / 2}


def _component_coords(labels: torch.Tensor, cid: int) -> np.ndarray:
    """np.argwhere(labeled == cid): (n,3) int64 [a0,a1,a2] in raster order."""
    mask = torch.empty(labels.shape, dtype=torch.uint8, device=labels.device)
    nv.check(nv.lib.p3d_label_equals(nv.ptr(labels), labels.numel(), int(cid), nv.ptr(mask), nv.stream_ptr()),
             "p3d_label_equals")
    eng._launched(1)
    pts, _ = eng.compact_points(mask)
    return pts.cpu().numpy()[:, ::-1].astype(np.int64)


def extract_minaret_voxels_by_label(voxel_grid, minaret_colors, device=None):
    """camera_estimation.py:176-207: the four tallest (extent along axis 1) 6-connected components of the minaret
    colours, as coordinate arrays named LM1/LM2 (left, front/back) and RM1/RM2."""
    from .voxel_carving_utils import _colour_mask, _label_components
    dev = nv.require_cuda(device)
    comps = []
    with torch.cuda.device(dev):
        grid = grid_to_device(voxel_grid, dev)
        for colour in minaret_colors:
            labels, n, bbox, sums = _label_components(_colour_mask(grid, colour))
            for cid in range(1, n + 1):
                cnt = sums[cid - 1, 0]
                centroid = sums[cid - 1, 1:4].astype(np.float64) / cnt           # coords.mean(axis=0): exact sums
                height = int(bbox[cid - 1, 4] - bbox[cid - 1, 1])                # coords[:, 1].ptp()
                comps.append((centroid, height, labels, cid))
        if len(comps) < 4:
            raise ValueError(f"Expected ≥4 minarets, found {len(comps)}")
        top4 = sorted(comps, key=lambda c: -c[1])[:4]
        centroids = np.stack([c[0] for c in top4])
        coord_sets = [_component_coords(c[2], c[3]) for c in top4]
    order_x = np.argsort(centroids[:, 0])
    left, right = order_x[:2], order_x[2:]
    left = sorted(left, key=lambda i: centroids[i, 2])
    right = sorted(right, key=lambda i: centroids[i, 2])
    return {"LM1": coord_sets[left[0]], "LM2": coord_sets[left[1]], "RM1": coord_sets[right[0]], "RM2": coord_sets[right[1]]}


def _label_image8(mask_u8: torch.Tensor):
    """skimage.measure.label for an (H,W) mask (8-connectivity): (labels int32, n, bbox (n,6), sums (n,4))."""
    H, W = mask_u8.shape
    dev = mask_u8.device
    labels = torch.empty((1, H, W), dtype=torch.int32, device=dev)
    ncomp = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(nv.lib.p3d_label6_workspace_bytes(H * W))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    nv.check(nv.lib.p3d_label8_2d(nv.ptr(mask_u8), H, W, nv.ptr(labels), nv.ptr(ncomp), nv.ptr(ws), ws_bytes,
                                  nv.stream_ptr()), "p3d_label8_2d")
    eng._launched(7)
    n = int(ncomp.item())
    bbox = torch.empty((max(n, 1), 6), dtype=torch.int32, device=dev)
    sums = torch.empty((max(n, 1), 4), dtype=torch.int64, device=dev)
    nv.check(nv.lib.p3d_component_stats(nv.ptr(labels), 1, H, W, n, nv.ptr(bbox), nv.ptr(sums), nv.stream_ptr()),
             "p3d_component_stats")
    eng._launched(2 if n else 0)
    return labels, n, bbox.cpu().numpy()[:n], sums.cpu().numpy()[:n]


def extract_minaret_masks_by_label(image, minaret_colors, min_area=50, device=None):
    """camera_estimation.py:247-323: 8-connected regions of the minaret colours with at least `min_area` pixels, split
    left/right by centroid column, front/back by colour order then centroid row; (H,W) uint8 masks."""
    dev = nv.require_cuda(device)
    regions = []
    with torch.cuda.device(dev):
        img = nv.to_device(image, torch.uint8, dev)[:, :, :3].contiguous()
        H, W = int(img.shape[0]), int(img.shape[1])
        for color_idx, colour in enumerate(minaret_colors):
            c = tuple(int(v) for v in np.asarray(colour).reshape(3))
            labels, n, _, sums = _label_image8(image_labels(img, [c], dev))
            for cid in range(1, n + 1):
                area = int(sums[cid - 1, 0])
                if area < min_area:
                    continue
                regions.append({"color_idx": color_idx, "centroid": (sums[cid - 1, 2] / area, sums[cid - 1, 3] / area),
                                "area": area, "label": cid, "labels": labels})
        if len(regions) < 2:
            raise ValueError("Not enough minarets for camera alignment")
        regions.sort(key=lambda r: r["centroid"][1])
        mid = len(regions) // 2

        def pick_front_back(side):
            if len(side) == 1:
                return side[0], None
            side = sorted(side, key=lambda r: (r["color_idx"], r["centroid"][0]))
            return side[0], side[1]

        (LM1, LM2), (RM1, RM2) = pick_front_back(regions[:mid]), pick_front_back(regions[mid:])

        def region_to_mask(region):
            m = torch.empty((H, W), dtype=torch.uint8, device=dev)
            nv.check(nv.lib.p3d_label_equals(nv.ptr(region["labels"]), H * W, region["label"], nv.ptr(m), nv.stream_ptr()),
                     "p3d_label_equals")
            eng._launched(1)
            return m.cpu().numpy()

        out = {}
        for key, region in (("LM1", LM1), ("RM1", RM1), ("LM2", LM2), ("RM2", RM2)):
            if region:
                out[key] = region_to_mask(region)
    return out


def _coords_extremes(coords_i32: torch.Tensor, axis: int):
    mm = torch.empty(2, dtype=torch.int32, device=coords_i32.device)
    sums = torch.empty((2, 4), dtype=torch.int64, device=coords_i32.device)
    nv.check(nv.lib.p3d_coords_extremes(nv.ptr(coords_i32), coords_i32.shape[0], axis, nv.ptr(mm), nv.ptr(sums),
                                        nv.stream_ptr()), "p3d_coords_extremes")
    eng._launched(3)
    return mm.cpu().numpy(), sums.cpu().numpy()


def extract_top_bottom_voxel_points(voxel_parts, device=None):
    """camera_estimation.py:329-335: mean coordinate of the voxels at the lowest / highest axis-1 value."""
    dev = nv.require_cuda(device)
    out = {}
    for name, vox in voxel_parts.items():
        v = np.asarray(vox)
        if v.shape[0] == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")
        _, sums = _coords_extremes(torch.from_numpy(np.ascontiguousarray(v, dtype=np.int32)).to(dev), 1)
        out[f"{name}_bottom"] = sums[0, 1:4].astype(np.float64) / sums[0, 0]
        out[f"{name}_top"] = sums[1, 1:4].astype(np.float64) / sums[1, 0]
    return out


def extract_top_bottom_image_points(mask_parts, device=None):
    """camera_estimation.py:338-344: (mean column of the mask's first row, that row) and the same for its last row."""
    dev = nv.require_cuda(device)
    out = {}
    with torch.cuda.device(dev):
        for name, mask in mask_parts.items():
            m = nv.to_device((np.asarray(mask) != 0).astype(np.uint8), torch.uint8, dev).to(torch.int32).contiguous()
            H, W = m.shape
            bbox = torch.empty((1, 6), dtype=torch.int32, device=dev)
            sums = torch.empty((1, 4), dtype=torch.int64, device=dev)
            ext = torch.empty((1, 2, 4), dtype=torch.int64, device=dev)
            nv.check(nv.lib.p3d_component_stats(nv.ptr(m), 1, H, W, 1, nv.ptr(bbox), nv.ptr(sums), nv.stream_ptr()),
                     "p3d_component_stats")
            nv.check(nv.lib.p3d_component_extremes(nv.ptr(m), 1, H, W, 1, 1, nv.ptr(bbox), nv.ptr(ext), nv.stream_ptr()),
                     "p3d_component_extremes")
            eng._launched(3)
            if int(sums[0, 0].item()) == 0:
                raise ValueError("zero-size array to reduction operation minimum which has no identity")
            b, e = bbox.cpu().numpy()[0], ext.cpu().numpy()[0]
            out[f"{name}_top"] = (e[0, 3] / e[0, 0], b[1].astype(np.int64))
            out[f"{name}_bottom"] = (e[1, 3] / e[1, 0], b[4].astype(np.int64))
    return out


def extract_minaret_kps_for_view(voxel_grid, mask_img, minaret_colors, back_top_only=False, device=None):
    """camera_estimation.py:20-50: matching 3-D / 2-D key points of the minarets visible in this view (front minarets:
    top and bottom; back minarets: top only)."""
    voxel_parts = extract_minaret_voxels_by_label(voxel_grid, minaret_colors, device=device)
    mask_parts = extract_minaret_masks_by_label(mask_img, minaret_colors, device=device)
    common = [k for k in voxel_parts if k in mask_parts]       # the reference's list(set & set) has no defined order
    if len(common) < 2:
        raise ValueError("Not enough visible minarets")
    voxel_kps = extract_top_bottom_voxel_points({k: voxel_parts[k] for k in common}, device=device)
    image_kps = extract_top_bottom_image_points({k: mask_parts[k] for k in common}, device=device)
    voxel_sel, image_sel = {}, {}
    for k in voxel_kps:
        m = k.split("_")[0]
        if ("1" in m) or ("2" in m and "top" in k):
            voxel_sel[k] = voxel_kps[k]
            image_sel[k] = image_kps[k]
    if len(voxel_sel) < 2:
        raise ValueError("Not enough keypoints after filtering")
    return voxel_sel, image_sel


def optimize_camera_with_keypoints(voxel_keypoints_dict, image_keypoints_dict, image, init_params, loss_type="L2",
                                   verbose=True):
    """camera_estimation.py:110-170: L-BFGS-B (scipy, on the host like the reference: 9 unknowns, <= 8 key points) on the
    squared / absolute reprojection error; the look-at rotation of every evaluation comes from the device kernel the
    sweep uses (camera_geometry.project)."""
    from scipy.optimize import minimize
    from .camera_geometry import look_at_rotation
    H, W = np.asarray(image).shape[:2] if not isinstance(image, torch.Tensor) else image.shape[:2]
    keys = list(image_keypoints_dict.keys())

    def loss_fn(x):
        cam_pos, target = np.array([x[0], x[1], x[2]]), np.array([x[3], x[4], x[5]])
        R = look_at_rotation(cam_pos, target)             # one device round trip per evaluation, shared by all key points
        total = 0
        for k in keys:
            X, Y, Z = (np.asarray(voxel_keypoints_dict[k]) - cam_pos) @ R.T      # camera_geometry.project :17-27
            Z = max(Z, 1e-8)
            proj_pt = np.array([(X / Z) * x[6] + x[7], -(Y / Z) * x[6] + x[8]])
            gt_pt = image_keypoints_dict[k]
            error = np.abs(proj_pt - gt_pt) if loss_type == "L1" else (proj_pt - gt_pt) ** 2
            total += error.sum()
        return total

    x0 = [*init_params["cam_pos"], *init_params["target"], init_params["f"], init_params["cx"], init_params["cy"]]
    bounds = [(-W, 2 * W), (-H, 2 * H), (-2000, 100), (-W, 2 * W), (-H, 2 * H), (-2000, 100), (10, 2000), (0, W), (0, H)]
    result = minimize(loss_fn, x0, bounds=bounds, method="L-BFGS-B")
    x = result.x
    final_params = {"cam_pos": np.array([x[0], x[1], x[2]]), "target": np.array([x[3], x[4], x[5]]), "f": x[6],
                    "cx": x[7], "cy": x[8]}
    if verbose:
        print("\n\U0001F4F7 Optimized Camera Parameters:")
        for k, v in final_params.items():
            print(f"{k}: {v}")
        print(f"\U0001F4C9 Final Reprojection Loss: {result.fun:.2f}")
    return final_params
