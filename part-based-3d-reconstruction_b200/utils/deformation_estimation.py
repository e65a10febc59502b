"""Drop-in for the reference's utils/deformation_estimation.py (stage 3).

Reference functions covered: the closures of launch_deform_viewer_fixed_camera (:15-356) --
`deform_coords` (:70-103), `update` (:105-145), `save_params` (:263-284), `save_deformed_grid`
(:288-311), `on_part_change` (:313-327), `project_fast` (:32-67) and the commented-out
`run_auto_align` grid search (:148-258) -- as methods of a headless `DeformViewer`; the ipywidgets
UI itself is not reproduced.  All arithmetic runs in csrc/p3d_deform.cu.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from .mask_utils import image_labels
from .voxel_utils import device_points_by_parts, grid_to_device

_KEYS = ("scale_y", "shift_y", "scale_xz", "shift_xz")                      # column order of a deformation row
_RANGES = {"scale_y": (0.5, 2.0), "shift_y": (-100.0, 100.0), "scale_xz": (0.5, 2.0), "shift_xz": (-100.0, 100.0)}   # :22-25
_IDENTITY = {"scale_y": 1.0, "shift_y": 0.0, "scale_xz": 1.0, "shift_xz": 0.0}


def deform_row(deform) -> np.ndarray:
    return np.array([float(deform[k]) for k in _KEYS], dtype=np.float64)


def row_to_deform(row) -> dict:
    return {k: float(v) for k, v in zip(_KEYS, row)}


def _pix2vox(image_shape, voxel_shape) -> np.ndarray:
    """deformation_estimation.py:74-78 (the z factor divides by the image WIDTH, as the reference does)."""
    H_img, W_img = image_shape
    D, H, W = voxel_shape
    return np.array([W / float(W_img), H / float(H_img), D / float(W_img)], dtype=np.float64)


class _PartPoints:
    """Device state of one (part, stride): points, exact coordinate sums and the seven jitter centres."""

    def __init__(self, pts: torch.Tensor, stride: int):
        self.pts, self.stride = pts, int(stride)
        self.n = int(pts.shape[0])
        self.m = (self.n + self.stride - 1) // self.stride
        self.sums = torch.empty(4, dtype=torch.int64, device=pts.device)
        self.centres = torch.empty(9, dtype=torch.float64, device=pts.device)
        nv.check(nv.lib.p3d_deform_centres(nv.ptr(pts), self.n, self.stride, nv.ptr(self.sums), nv.ptr(self.centres),
                                           nv.stream_ptr()), "p3d_deform_centres")
        eng._launched(2 if self.m else 1)
        if int(self.sums[3].item()) != 0:
            raise ValueError("deform_coords: coordinates must be integer voxel indices (as get_voxel_points_by_parts "
                             "returns them)")


def _as_points(coords, device) -> torch.Tensor:
    t = coords if isinstance(coords, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float32))
    if t.dim() != 2 or t.shape[1] != 3:
        raise ValueError(f"expected (N,3) coordinates, got {tuple(t.shape)}")
    return t.to(device=device, dtype=torch.float32).contiguous()


@nv.on_device
def deform_coords(coords, image_shape, voxel_shape, deform, device=None):
    """The `deform_coords` closure (deformation_estimation.py:70-103): seven jittered copies of the coordinates, each
    scaled/shifted about its own mean and rounded half-even; unique rows in lexicographic (x, y, z) order, int64.
    The deformation runs in p3d_deform_points; the row sort behind np.unique is torch.unique on the device."""
    dev = nv.require_cuda(device)
    with torch.cuda.device(dev):
        pts = _as_points(coords, dev)
        if pts.shape[0] == 0:
            raise ValueError("deform_coords: empty coordinate list (the reference's mean of an empty array is NaN)")
        pp = _PartPoints(pts, 1)
        d = torch.from_numpy(deform_row(deform)).to(dev)
        p2v = torch.from_numpy(_pix2vox(image_shape, voxel_shape)).to(dev)
        out = torch.empty((7, pp.m, 3), dtype=torch.int64, device=dev)
        nv.check(nv.lib.p3d_deform_points(nv.ptr(pts), pp.n, 1, nv.ptr(pp.centres), nv.ptr(d), nv.ptr(p2v), nv.ptr(out),
                                          nv.stream_ptr()), "p3d_deform_points")
        eng._launched(1)
        uniq = torch.unique(out.view(-1, 3), dim=0)
    return uniq.cpu().numpy()


class DeformViewer:
    """Headless launch_deform_viewer_fixed_camera: the widget callbacks as methods, plus a batched scorer.

    sliders      dict with 'scale_y', 'shift_y', 'scale_xz', 'shift_xz' (clamped like the FloatSliders :22-25) and 'part'
    saved_params part -> {'deform': {...}, 'iou': float}   (filled by save_params, :263-284)
    grid_storage {'grid': ...}                             (filled by save_deformed_grid, :288-311, 330-333)
    """

    def __init__(self, voxel_grid, part_labels, image, cam_params, part_names, init_params=None, device=None,
                 verbose=True):
        self.device = nv.require_cuda(device)
        self.verbose = verbose
        self.part_labels = dict(part_labels)
        self.part_names = list(part_names)
        with torch.cuda.device(self.device):
            self.grid = grid_to_device(voxel_grid, self.device)
            self.voxel_shape = tuple(int(v) for v in self.grid.shape[:3])
            self.image = nv.to_device(image, torch.uint8, self.device)
            self.H, self.W = int(self.image.shape[0]), int(self.image.shape[1])
            self.p2v = torch.from_numpy(_pix2vox((self.H, self.W), self.voxel_shape)).to(self.device)
            self._set_camera(cam_params)
        self.saved_params = init_params.copy() if init_params else {}        # :30
        self.grid_storage = {"grid": None}
        self.sliders = dict(_IDENTITY, part=self.part_names[0] if self.part_names else None)
        self._points = {}
        self._gt_bits = {}
        self._cov = None
        self.evaluations = 0
        if self.part_names:
            self.on_part_change(self.sliders["part"])                         # :355

    # ---- camera: the projection's working dtype follows the camera arrays (projection_utils.py:7) -------------
    def _set_camera(self, cam_params):
        arrs = [np.asarray(cam_params["cam_pos"]), np.asarray(cam_params["target"])]
        f32 = all(a.dtype == np.float32 for a in arrs)
        self.cam_dtype = np.float32 if f32 else np.float64
        row = np.array([*np.asarray(cam_params["cam_pos"], dtype=self.cam_dtype).reshape(3),
                        *np.asarray(cam_params["target"], dtype=self.cam_dtype).reshape(3),
                        cam_params["f"], cam_params["cx"], cam_params["cy"]], dtype=self.cam_dtype)
        self.cam_params = cam_params
        cand = torch.from_numpy(row[None]).to(self.device)
        self.cam = eng.setup_cameras(cand)
        D, H, W = self.voxel_shape                         # FP32 filter block over the whole grid box (every valid voxel)
        self.bbox = torch.tensor([0, 0, 0, W - 1, H - 1, D - 1, 0, 0], dtype=torch.float32, device=self.device)
        self.fast = torch.empty((1, 16), dtype=torch.float32, device=self.device)
        fn = nv.lib.p3d_fast_cameras_f32 if f32 else nv.lib.p3d_fast_cameras_f64
        nv.check(fn(nv.ptr(self.cam), 1, nv.ptr(self.bbox), self.H, self.W, nv.ptr(self.fast), nv.stream_ptr()),
                 "p3d_fast_cameras")
        eng._launched(1)

    # ---- per-part device state -----------------------------------------------------------------------
    def part_points(self, part, stride=1) -> _PartPoints:
        key = (part, int(stride))
        if key not in self._points:
            base = self._points.get((part, 1))
            if base is None:
                pts, _, _, _ = device_points_by_parts(self.grid, self.part_labels, [part], self.device)
            else:
                pts = base.pts
            if pts.shape[0] == 0:
                raise ZeroDivisionError(f"part {part!r} has no voxels (the reference divides by len(colors) == 0)")
            self._points[key] = _PartPoints(pts, stride)
        return self._points[key]

    def _part_gt_bits(self, part) -> torch.Tensor:
        if part not in self._gt_bits:
            colour = tuple(int(v) for v in np.asarray(self.part_labels[part]).reshape(3))
            lab = image_labels(self.image, [colour], self.device)
            words = (self.H * self.W + 31) // 32
            bits = torch.zeros(words, dtype=torch.int32, device=self.device)
            nv.check(nv.lib.p3d_pack_label_bits(nv.ptr(lab), self.H * self.W, 1, nv.ptr(bits), nv.stream_ptr()),
                     "p3d_pack_label_bits")
            eng._launched(1)
            self._gt_bits[part] = bits
        return self._gt_bits[part]

    # ---- batched scoring ------------------------------------------------------------------------------
    def score(self, part, deforms, stride=1, batch=1024):
        """IoU of `part` after each deformation of `deforms` ((D,4) rows [scale_y, shift_y, scale_xz, shift_xz] or a
        list of dicts), computed as save_params does (:263-284) on every `stride`-th voxel (project_fast :35-38).
        Returns (ious (D) f64, counts (D,2) int64 = inter, union, nvalid (D) int64)."""
        rows = np.stack([deform_row(d) for d in deforms]) if not isinstance(deforms, np.ndarray) else \
            np.ascontiguousarray(deforms, dtype=np.float64).reshape(-1, 4)
        D = rows.shape[0]
        with torch.cuda.device(self.device):
            pp = self.part_points(part, stride)
            gt = self._part_gt_bits(part)
            words = (self.H * self.W + 31) // 32
            batch = max(1, min(batch, D, (256 << 20) // (4 * words)))
            if self._cov is None or self._cov.numel() < batch * words:
                self._cov = torch.zeros(batch * words, dtype=torch.int32, device=self.device)
            dev_rows = torch.from_numpy(rows).to(self.device)
            counts = torch.empty((D, 2), dtype=torch.int64, device=self.device)
            nvalid = torch.empty(D, dtype=torch.int64, device=self.device)
            A0, A1, A2 = self.voxel_shape
            for d0 in range(0, D, batch):
                nd = min(batch, D - d0)
                fn = nv.lib.p3d_deform_sweep_f64 if self.cam_dtype == np.float64 else nv.lib.p3d_deform_sweep_f32
                rc = fn(nv.ptr(pp.pts), pp.n, pp.stride, nv.ptr(pp.centres), nv.ptr(dev_rows[d0:]), nd, nv.ptr(self.p2v),
                        A0, A1, A2, nv.ptr(self.cam), nv.ptr(self.fast), nv.ptr(self.bbox), nv.ptr(gt), self.H, self.W,
                        nv.ptr(self._cov), nv.ptr(counts[d0:]), nv.ptr(nvalid[d0:]), nv.stream_ptr())
                nv.check(rc, "p3d_deform_sweep")
                eng._launched(2)
            c = counts.cpu().numpy()
            nval = nvalid.cpu().numpy()
        self.evaluations += D
        ious = np.array([i / u if u > 0 else 0.0 for i, u in c], dtype=np.float64)    # camera_estimation.py:783
        return ious, c, nval

    def iou(self, part, deform, stride=1) -> float:
        return float(self.score(part, [deform], stride)[0][0])

    # ---- widget callbacks -----------------------------------------------------------------------------
    def set_sliders(self, **values):
        for k, v in values.items():
            if k == "part":
                self.on_part_change(v)
            else:
                lo, hi = _RANGES[k]
                self.sliders[k] = float(min(max(float(v), lo), hi))

    def current_deform(self) -> dict:
        return {k: self.sliders[k] for k in _KEYS}

    def update(self):
        """:105-145 without the figure: prints (and returns) the IoU of the current part at the current sliders."""
        part, deform = self.sliders["part"], self.current_deform()
        try:
            ious, _, nvalid = self.score(part, [deform])
        except ZeroDivisionError:                 # a part without voxels deforms to nothing in the reference (:116-120)
            ious, nvalid = np.zeros(1), np.zeros(1, np.int64)
        if nvalid[0] == 0:
            if self.verbose:
                print("No deformed voxels within bounds. Adjust sliders.")
            return None
        if self.verbose:
            print(f"{part} | IoU: {ious[0]:.4f}")
        return float(ious[0])

    def save_params(self):
        """:263-284."""
        part, deform = self.sliders["part"], self.current_deform()
        iou = self.iou(part, deform)
        self.saved_params[part] = {"deform": deform, "iou": iou}
        if self.verbose:
            print(f"✔ Saved {part} | IoU: {iou:.4f}")
        return iou

    def save_deformed_grid(self, return_tensor=False):
        """:288-311, 330-333: every part with saved parameters re-drawn at its deformed coordinates, in
        `part_labels` order; stores the grid in grid_storage['grid'] and returns it."""
        A0, A1, A2 = self.voxel_shape
        with torch.cuda.device(self.device):
            out = torch.zeros((A0, A1, A2, 3), dtype=torch.uint8, device=self.device)
            for part, colour in self.part_labels.items():
                if part not in self.saved_params:
                    continue
                pp = self.part_points(part)
                d = torch.from_numpy(deform_row(self.saved_params[part]["deform"])).to(self.device)
                r, g, b = (int(v) for v in np.asarray(colour).reshape(3))
                nv.check(nv.lib.p3d_deform_scatter(nv.ptr(pp.pts), pp.n, 1, nv.ptr(pp.centres), nv.ptr(d), nv.ptr(self.p2v),
                                                   A0, A1, A2, r, g, b, nv.ptr(out), None, nv.stream_ptr()),
                         "p3d_deform_scatter")
                eng._launched(1)
            grid = out if return_tensor else out.cpu().numpy()
        self.grid_storage["grid"] = grid
        if self.verbose:
            print("\U0001F4BE Full deformed voxel grid saved.")
        return grid

    def on_part_change(self, part):
        """:313-327: restore the saved sliders of the part (identity otherwise), then update()."""
        self.sliders["part"] = part
        deform = self.saved_params[part]["deform"] if part in self.saved_params else _IDENTITY
        for k in _KEYS:
            self.set_sliders(**{k: deform[k]})
        return self.update()

    # ---- the (commented-out) grid search, batched --------------------------------------------------------
    def project_fast_iou(self, part, deforms, stride=8):
        """`project_fast` (:32-67) for many deformations: IoU on every stride-th voxel, None where fewer than 10
        deformed voxels stay inside the grid (:52-53; decided on the de-duplicated count, as the reference does)."""
        ious, _, nvalid = self.score(part, deforms, stride=stride)
        rows = np.stack([deform_row(d) for d in deforms]) if not isinstance(deforms, np.ndarray) else deforms.reshape(-1, 4)
        out = []
        pp = self.part_points(part, stride)
        for k in range(len(ious)):
            if nvalid[k] < 10:
                out.append(None)
            elif nvalid[k] >= 10 * 7 * 108:       # >= 10 distinct voxels for any scale >= 0.5 (<= 756 pairs per voxel)
                out.append(float(ious[k]))
            else:                                 # ambiguous: count the distinct valid voxels exactly
                cd = deform_coords(pp.pts[::stride], (self.H, self.W), self.voxel_shape, row_to_deform(rows[k]), self.device)
                A0, A1, A2 = self.voxel_shape
                ok = (cd[:, 0] >= 0) & (cd[:, 0] < A2) & (cd[:, 1] >= 0) & (cd[:, 1] < A1) & (cd[:, 2] >= 0) & (cd[:, 2] < A0)
                out.append(float(ious[k]) if int(ok.sum()) >= 10 else None)
        return out

    def run_auto_align(self, part=None):
        """The reference's commented-out `run_auto_align` (:148-258) on the batched scorer: coarse 7x7x9x9 grid at
        stride 6, 5^4 refinement at stride 4, strict `>` selection in loop order; moves the sliders to the best
        deformation and returns (best_deform, best_iou), or (None, -1.0) when no candidate keeps 10 voxels."""
        part = self.sliders["part"] if part is None else part
        scale_vals, shift_vals = np.linspace(0.8, 1.2, 7), np.linspace(-60, 60, 9)
        coarse = [{"scale_y": sy, "shift_y": dy, "scale_xz": sxz, "shift_xz": dxz}
                  for sy in scale_vals for sxz in scale_vals for dy in shift_vals for dxz in shift_vals]
        best_iou, best = -1.0, None
        for d, iou in zip(coarse, self.project_fast_iou(part, coarse, stride=6)):
            if iou is not None and iou > best_iou:
                best_iou, best = iou, dict(d)
        if best is None:
            if self.verbose:
                print(f"❌ Auto-align failed for part '{part}'")
            return None, -1.0
        refine_scales = np.linspace(best["scale_y"] - 0.05, best["scale_y"] + 0.05, 5)       # both axes use scale_y (:211)
        refine_shifts = np.linspace(best["shift_y"] - 10, best["shift_y"] + 10, 5)
        fine = [{"scale_y": sy, "shift_y": dy, "scale_xz": sxz, "shift_xz": dxz}
                for sy in refine_scales for sxz in refine_scales for dy in refine_shifts for dxz in refine_shifts]
        for d, iou in zip(fine, self.project_fast_iou(part, fine, stride=4)):
            if iou is not None and iou > best_iou:
                best_iou, best = iou, dict(d)
        self.sliders["part"] = part
        self.set_sliders(**best)
        if self.verbose:
            print("✔ Auto-align done")
            print("Best IoU:", best_iou)
            print("Best deform:", best)
        return best, best_iou


def launch_deform_viewer_fixed_camera(voxel_grid, part_labels, image, cam_params, part_names, init_params=None,
                                      device=None):
    """deformation_estimation.py:15-356 without the ipywidgets UI: returns `(saved_params, grid_storage)` like the
    reference; both carry the DeformViewer that fills them as `.viewer`, so a notebook drives
    `.set_sliders(...) / .save_params() / .save_deformed_grid() / .run_auto_align()` where the widgets used to be."""
    viewer = DeformViewer(voxel_grid, part_labels, image, cam_params, part_names, init_params, device)

    class _Saved(dict):
        pass

    class _Store(dict):
        pass

    saved, store = _Saved(viewer.saved_params), _Store(viewer.grid_storage)
    viewer.saved_params, viewer.grid_storage = saved, store
    saved.viewer = store.viewer = viewer
    return saved, store
