"""Device-level operations: torch CUDA tensors in, torch CUDA tensors out.

Thin, typed wrappers over the C ABI (include/p3d_b200.h).  The reference-signature
modules next to this file (projection_utils, camera_estimation, voxel_utils, ...)
move NumPy data to the device and call these.
"""
from __future__ import annotations

import torch

from . import _native as nv
from ._native import check, lib, ptr, stream_ptr


def _launched(n: int = 1) -> None:
    nv.launch_count += n


def rgb_to_labels(rgb: torch.Tensor, palette: torch.Tensor) -> torch.Tensor:
    """(..., 3) u8 -> (...) u8, 1 + index of the first matching palette colour, else 0."""
    assert rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.shape[-1] == 3 and rgb.is_contiguous()
    out = torch.empty(rgb.shape[:-1], dtype=torch.uint8, device=rgb.device)
    n = out.numel()
    check(lib.p3d_rgb_to_labels(ptr(rgb), n, ptr(palette), int(palette.shape[0]), ptr(out), stream_ptr()),
          "p3d_rgb_to_labels")
    _launched(1 if n else 0)
    return out


def labels_to_rgb(labels: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
    """(...) u8 -> (..., 3) u8 through a (256,3) u8 lookup table."""
    assert labels.is_cuda and labels.dtype == torch.uint8 and labels.is_contiguous()
    assert lut.shape == (256, 3) and lut.dtype == torch.uint8 and lut.is_contiguous()
    out = torch.empty(tuple(labels.shape) + (3,), dtype=torch.uint8, device=labels.device)
    n = labels.numel()
    check(lib.p3d_labels_to_rgb(ptr(labels), n, ptr(lut), ptr(out), stream_ptr()), "p3d_labels_to_rgb")
    _launched(1 if n else 0)
    return out


def make_lut(palette: torch.Tensor) -> torch.Tensor:
    """(P,3) palette -> (256,3) LUT with label 0 -> black and label k -> palette[k-1]."""
    lut = torch.zeros((256, 3), dtype=torch.uint8, device=palette.device)
    lut[1:1 + palette.shape[0]] = palette
    return lut


def compact_points(labels: torch.Tensor):
    """Dense (A0,A1,A2) u8 label grid -> (pts (N,3) f32 [x=a2,y=a1,z=a0], pt_label (N) u8) in
    ascending flat index (voxel_utils.py:17-19)."""
    assert labels.is_cuda and labels.dtype == torch.uint8 and labels.dim() == 3 and labels.is_contiguous()
    dev = labels.device
    nvox = labels.numel()
    ws_bytes = int(lib.p3d_points_workspace_bytes(nvox))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    check(lib.p3d_points_count(ptr(labels), nvox, ptr(n_out), ptr(ws), ws_bytes, stream_ptr()), "p3d_points_count")
    _launched(2)
    n = int(n_out.item())                       # one 8-byte read-back: sizes the output
    pts = torch.empty((n, 3), dtype=torch.float32, device=dev)
    pt_label = torch.empty((n,), dtype=torch.uint8, device=dev)
    A0, A1, A2 = labels.shape
    check(lib.p3d_points_fill(ptr(labels), A0, A1, A2, ptr(ws), ptr(pts), ptr(pt_label), n, stream_ptr()),
          "p3d_points_fill")
    _launched(1 if n else 0)
    return pts, pt_label


def _elem(dtype: torch.dtype) -> str:
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError(f"camera dtype must be float32 or float64, got {dtype}")


def setup_cameras(cand: torch.Tensor) -> torch.Tensor:
    """(K,9) candidates -> (K,16) camera blocks (cam_pos, R row-major, f, cx, cy, 0)."""
    assert cand.is_cuda and cand.dim() == 2 and cand.shape[1] == 9 and cand.is_contiguous()
    K = cand.shape[0]
    cams = torch.empty((K, 16), dtype=cand.dtype, device=cand.device)
    fn = getattr(lib, f"p3d_setup_cameras_{_elem(cand.dtype)}")
    check(fn(ptr(cand), K, ptr(cams), stream_ptr()), "p3d_setup_cameras")
    _launched(1 if K else 0)
    return cams


def points_bbox(pts: torch.Tensor) -> torch.Tensor:
    """(6,) f32 = min x,y,z, max x,y,z of the point list (feeds the FP32 filter of the f64 splat)."""
    bbox = torch.empty(8, dtype=torch.float32, device=pts.device)
    check(lib.p3d_points_bbox(ptr(pts), pts.shape[0], ptr(bbox), stream_ptr()), "p3d_points_bbox")
    _launched(2 if pts.shape[0] else 1)
    return bbox


def splat(pts: torch.Tensor, pt_label, cams: torch.Tensor, H: int, W: int, mode: int = nv.MODE_JOINT,
          bbox=None) -> torch.Tensor:
    """Scatter the points through K cameras into a fresh (K,H,W) z-buffer (int32 storage of uint32 keys)."""
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.is_contiguous()
    K = cams.shape[0]
    zbuf = torch.zeros((K, H, W), dtype=torch.int32, device=pts.device)
    n = pts.shape[0]
    el = _elem(cams.dtype)
    fast = None
    if n and K:
        if bbox is None:
            bbox = points_bbox(pts)
        fast = torch.empty((K, 16), dtype=torch.float32, device=pts.device)
        check(getattr(lib, f"p3d_fast_cameras_{el}")(ptr(cams), K, ptr(bbox), H, W, ptr(fast), stream_ptr()),
              "p3d_fast_cameras")
        _launched(1)
    check(getattr(lib, f"p3d_splat_{el}")(ptr(pts), ptr(pt_label), n, ptr(cams), K, H, W, mode, ptr(zbuf), ptr(fast),
                                          ptr(bbox if fast is not None else None), stream_ptr()), "p3d_splat")
    _launched(1 if (K and n) else 0)
    return zbuf


def resolve_rgb(zbuf: torch.Tensor, pt_rgb: torch.Tensor) -> torch.Tensor:
    """(H,W) joint-mode z-buffer + per-point colours (N,3) u8 -> (H,W,3) u8 image."""
    assert zbuf.dim() == 2 and zbuf.is_contiguous()
    H, W = zbuf.shape
    img = torch.empty((H, W, 3), dtype=torch.uint8, device=zbuf.device)
    check(lib.p3d_resolve_rgb(ptr(zbuf), ptr(pt_rgb), H * W, ptr(img), stream_ptr()), "p3d_resolve_rgb")
    _launched(1)
    return img


def partwise_counts_rgb(proj: torch.Tensor, gt: torch.Tensor, part_rgb: torch.Tensor) -> torch.Tensor:
    """Two (N,3) u8 images -> (P,2) int64 (inter, union) per part colour."""
    assert proj.is_contiguous() and gt.is_contiguous() and proj.numel() == gt.numel()
    P = int(part_rgb.shape[0])
    counts = torch.empty((P, 2), dtype=torch.int64, device=proj.device)
    check(lib.p3d_partwise_counts_rgb(ptr(proj), ptr(gt), proj.numel() // 3, ptr(part_rgb), P, ptr(counts),
                                      stream_ptr()), "p3d_partwise_counts_rgb")
    _launched(1)
    return counts


class SweepWorkspace:
    """Caller-owned state of p3d_sweep_*: the device scratch (z-buffer batches, camera blocks, raw counters) and the
    host-side context (p3d_sweep_ctx: helper stream + events of the double-buffered batches, launch counter, optional
    splat timing).  One per concurrent caller."""

    def __init__(self, device):
        self.device = device
        self.buf = None
        self._ctx = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.buf

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = lib.p3d_sweep_ctx_create()
            if not self._ctx:
                raise MemoryError("p3d_sweep_ctx_create failed")
        return self._ctx

    def timing(self, on: bool) -> None:
        check(lib.p3d_sweep_ctx_timing(self.ctx, 1 if on else 0), "p3d_sweep_ctx_timing")

    def timing_read(self):
        """(summed splat-launch duration in ms, number of splat launches) since timing(True); synchronises."""
        import ctypes
        ms, n = ctypes.c_double(), ctypes.c_int()
        check(lib.p3d_sweep_ctx_timing_read(self.ctx, ctypes.byref(ms), ctypes.byref(n)), "p3d_sweep_ctx_timing_read")
        return ms.value, n.value

    def __del__(self):
        ctx, self._ctx = self._ctx, None
        if ctx:
            try:
                lib.p3d_sweep_ctx_destroy(ctx)
            except Exception:
                pass


def build_segments(pts: torch.Tensor, pt_label: torch.Tensor):
    """x-run segments of a point list for the segment splat (p3d_segments_*): (S,4) int32 tensor, or None when the list
    cannot be represented (non-integer coordinates, coordinates outside 0..65535, labels outside 1..32) -- the sweep
    then runs the per-point splat."""
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.is_contiguous()
    assert pt_label.dtype == torch.uint8 and pt_label.is_contiguous()
    n = int(pts.shape[0])
    if n == 0:
        return None
    dev = pts.device
    L = int(lib.p3d_segment_length())
    ws_bytes = int(lib.p3d_segments_workspace_bytes(n))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=dev)
    n_out = torch.zeros(2, dtype=torch.int64, device=dev)
    check(lib.p3d_segments_count(ptr(pts), ptr(pt_label), n, L, ptr(n_out), ptr(ws), ws_bytes, stream_ptr()),
          "p3d_segments_count")
    _launched(2)
    n_seg, bad = (int(v) for v in n_out.cpu().numpy())      # one 16-byte read-back: sizes the output
    if bad or n_seg == 0:
        return None
    segs = torch.empty((n_seg, 4), dtype=torch.int32, device=dev)
    check(lib.p3d_segments_fill(ptr(pts), ptr(pt_label), n, L, ptr(ws), ptr(segs), n_seg, stream_ptr()),
          "p3d_segments_fill")
    _launched(1)
    return segs


def sweep(pts: torch.Tensor, pt_label: torch.Tensor, cand: torch.Tensor, gt_label: torch.Tensor, H: int, W: int,
          P: int, mode: int = nv.MODE_JOINT, gt_any=None, workspace: SweepWorkspace | None = None,
          want_best: bool = True, segs=None):
    """Score K candidate cameras.  Returns (counts (K,rows,2) i64, scores (K) f64, best (2) i64 | None).
    `segs` = build_segments(pts, pt_label) selects the segment splat."""
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.is_contiguous()
    assert pt_label.dtype == torch.uint8 and pt_label.is_contiguous()
    assert gt_label.dtype == torch.uint8 and gt_label.is_contiguous() and gt_label.numel() == H * W
    assert cand.dim() == 2 and cand.shape[1] == 9 and cand.is_contiguous()
    dev = pts.device
    K = int(cand.shape[0])
    rows = P + 1 if mode == nv.MODE_PER_PART else P
    counts = torch.empty((K, rows, 2), dtype=torch.int64, device=dev)
    scores = torch.empty((K,), dtype=torch.float64, device=dev)
    best = torch.empty((2,), dtype=torch.int64, device=dev) if want_best else None
    if K == 0:
        return counts, scores, best
    elem = _elem(cand.dtype)
    nbytes = int(lib.p3d_sweep_workspace_bytes(K, H, W, P, 8 if elem == "f64" else 4))
    workspace = workspace or SweepWorkspace(dev)
    ws = workspace.get(nbytes)
    fn = getattr(lib, f"p3d_sweep_{elem}")
    n_seg = int(segs.shape[0]) if segs is not None else 0
    check(fn(ptr(pts), ptr(pt_label), pts.shape[0], ptr(segs), n_seg, ptr(cand), K, ptr(gt_label), ptr(gt_any), H, W, P,
             mode, ptr(counts), ptr(scores), ptr(best), ptr(ws), ws.numel(), workspace.ctx, stream_ptr()), "p3d_sweep")
    _launched(int(lib.p3d_sweep_ctx_launches(workspace.ctx)))
    return counts, scores, best
