"""Drop-in for the reference's utils/voxel_carving_utils.py (stage 1: orthographic semantic voxel carving).

Same function names, argument order, defaults and return conventions as the reference.  NumPy arrays in ->
NumPy arrays out (fresh arrays, inputs never mutated); CUDA tensors in -> CUDA tensors out, so that the stages
can be chained on the device.  All voxel work runs in the CUDA kernels of csrc/p3d_carve.cu; what stays on the
host is what the reference also does with tiny data: the 3x3 inverse rotation (NumPy/LAPACK, :65-69), the
2-D mask preparation (O(H*W)), and the per-component bookkeeping (bounding boxes, the stable sort of component
means).
"""
from __future__ import annotations

import ctypes
import warnings

import functools

import numpy as np
import torch

from . import _native as nv
from ._native import check, lib, ptr, stream_ptr
from .config import PART_COLORS, PART_COLORS_NP  # noqa: F401  (re-exported like the reference module)

try:                                              # progress bar as in the reference (:111-115); optional
    from tqdm import tqdm
except Exception:                                 # pragma: no cover
    def tqdm(it, **kwargs):
        return it


# =========================================================
# helpers
# =========================================================
def _launched(n=1):
    nv.launch_count += n


def _is_tensor(a):
    return isinstance(a, torch.Tensor)


def _to_dev_u8(a, dev, what="array"):
    if _is_tensor(a):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint8:
            return _upload_u8(arr, dev)
        t = torch.from_numpy(arr)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    if t.dtype != torch.uint8:
        # the reference takes any dtype; other integer types are accepted when every value fits a byte (the grids and
        # masks of this pipeline are 0..255 by construction), anything else is refused rather than silently wrapped
        if t.dtype in (torch.int8, torch.int16, torch.int32, torch.int64) and t.numel() and int(t.min()) >= 0 and int(t.max()) <= 255:
            t = t.to(torch.uint8)
        elif t.dtype in (torch.int8, torch.int16, torch.int32, torch.int64) and not t.numel():
            t = t.to(torch.uint8)
        else:
            raise TypeError(f"{what} must hold values 0..255 in an integer or bool dtype (got {t.dtype})")
    return t.to(dev).contiguous()


import threading

_PINNED = threading.local()   # staging buffers for the NumPy boundary (one per direction and thread), grown on demand
_PIN_MIN_BYTES = 1 << 20      # below this a pageable copy is as fast as staging


def _pinned(kind, nbytes):
    buf = getattr(_PINNED, kind, None)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        setattr(_PINNED, kind, buf)
    return buf[:nbytes]


def _upload_u8(arr, dev):
    """Contiguous uint8 NumPy array -> device tensor; large arrays go through a pinned staging buffer (one host memcpy
    + a DMA at full PCIe rate instead of the driver's pageable path)."""
    if arr.nbytes < _PIN_MIN_BYTES:
        return torch.from_numpy(arr).to(dev)
    stage = _pinned("h2d", arr.nbytes)
    stage.numpy()[:] = arr.reshape(-1)
    out = torch.empty(arr.shape, dtype=torch.uint8, device=dev)
    out.view(-1).copy_(stage, non_blocking=True)
    torch.cuda.current_stream().synchronize()              # the staging buffer is reused by the next call
    return out


def _first3(a):
    """Channels 0..2 of an image with more than three (the reference reads `mask[..., c] for c in range(3)`, :134-135,
    so an RGBA mask works there)."""
    if getattr(a, "ndim", 0) == 3 and a.shape[2] > 3:
        return a[..., :3]
    return a


def _ret(t, as_tensor):
    """Device tensor -> what the caller asked for: the tensor itself, or a FRESH NumPy array the caller owns."""
    if as_tensor:
        return t
    if t.numel() * t.element_size() < _PIN_MIN_BYTES or t.dtype != torch.uint8:
        return t.cpu().numpy()
    stage = _pinned("d2h", t.numel())
    stage.copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return stage.numpy().reshape(tuple(t.shape)).copy()


def _mask_to_wh(mask, W, H):
    """voxel_carving_utils.py:19-28.  The (H,W) test comes first, so a square mask is always transposed."""
    if tuple(mask.shape[:2]) == (H, W):
        return mask.T
    if tuple(mask.shape[:2]) == (W, H):
        return mask
    raise ValueError(f"Mask shape {tuple(mask.shape)} incompatible with (W,H)=({W},{H})")


def _rotation_matrix_inv(angle):
    """voxel_carving_utils.py:65-69 (host NumPy, as in the reference)."""
    a = np.deg2rad(angle)
    c, s = np.cos(a), np.sin(a)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.linalg.inv(R)


@functools.lru_cache(maxsize=8192)       # (crop shape, angle): a guided carve asks for 19 angles per component shape
def _pass_transform_cached(shape, angle):
    M, off = _pass_transform_uncached(shape, angle)
    M.setflags(write=False)
    off.setflags(write=False)
    return M, off


def _pass_transform(shape, angle):
    """Inverse rotation and offset of one process_voxel_grid pass; pure in (shape, angle), hence cached."""
    return _pass_transform_cached(tuple(int(v) for v in shape), angle)


def _pass_transform_uncached(shape, angle):
    """Matrix and offset handed to scipy.ndimage.affine_transform at :116-123."""
    M = np.ascontiguousarray(_rotation_matrix_inv(angle), dtype=np.float64)
    ctr = np.array(shape) / 2
    off = np.ascontiguousarray(ctr - M @ ctr, dtype=np.float64)
    return M, off


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _y_decoupled(M, off):
    return M[0, 1] == 0 and M[1, 0] == 0 and M[1, 1] == 1 and M[1, 2] == 0 and M[2, 1] == 0 and off[1] == 0


_FOLD_CACHE = {}          # (n0, n2, M bytes, off bytes, device) -> (table, foldable); a table is a few MB at most


def _fold_table(n0, n2, M, off, dev):
    """(table (n0,n2) int32, foldable) for a y-decoupled pass, else (None, False).  Cached per transform: the
    table depends only on the grid shape and on the host-computed matrix/offset."""
    if not _y_decoupled(M, off) or n0 >= 32768 or n2 >= 65536:
        return None, False
    cache_key = (n0, n2, M.tobytes(), off.tobytes(), str(dev))
    hit = _FOLD_CACHE.get(cache_key)
    if hit is not None:
        return hit
    table = torch.empty((n0, n2), dtype=torch.int32, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.p3d_fold_table(n0, n2, _dptr(M), _dptr(off), ptr(table), ptr(flag), stream_ptr()), "p3d_fold_table")
    _launched()
    res = (table, int(flag.item()) == 0)
    if len(_FOLD_CACHE) >= 16:
        _FOLD_CACHE.pop(next(iter(_FOLD_CACHE)))
    _FOLD_CACHE[cache_key] = res
    return res


_FOLD_BITS_CACHE = {}


def _fold_bits(table, W, D, cache_key):
    """(inside_bits (W, ceil(D/32)) int32 tensor, c, c2 | None) when the table is z-separable (src0 = c - z; c2 is set when
    additionally src2 = x + c2), else None.  Cached."""
    hit = _FOLD_BITS_CACHE.get(cache_key)
    if hit is not None:
        return hit if hit[0] is not None else None
    inside = torch.empty((W, (D + 31) // 32), dtype=torch.int32, device=table.device)
    info = torch.empty(4, dtype=torch.int32, device=table.device)
    check(lib.p3d_fold_analyse(ptr(table), W, D, ptr(inside), ptr(info), stream_ptr()), "p3d_fold_analyse")
    _launched()
    mx, mn, mx2, mn2 = (int(v) for v in info.cpu())
    res = (inside, mx, mx2 if mx2 == mn2 else None) if mx == mn else (None, None, None)
    if len(_FOLD_BITS_CACHE) >= 16:
        _FOLD_BITS_CACHE.pop(next(iter(_FOLD_BITS_CACHE)))
    _FOLD_BITS_CACHE[cache_key] = res
    return res if res[0] is not None else None


_FOLD_PLAN = {}           # (W, H, D, device) -> (table, bits) of the 0/90-degree fold, or None when it does not apply


def _fold_plan(W, H, D, dev):
    """Everything global_carve / part_carve need to know about the 90-degree index fold of a (W,H,D) grid, behind ONE
    dictionary lookup: (table, bits) when pass 0 is the identity and pass 90 is an exact index fold (bits = its
    z-separable bit form, or None), else None.  The tables depend only on the shape."""
    key = (W, H, D, dev.index)
    if key in _FOLD_PLAN:
        return _FOLD_PLAN[key]
    plan = None
    M0, off0 = _pass_transform((W, H, D), 0)
    M, off = _pass_transform((W, H, D), 90)
    if np.array_equal(M0, np.eye(3)) and not off0.any():
        table, foldable = _fold_table(W, D, M, off, dev)
        if foldable:
            # bit form: aligned kernels for D % 32 == 0, the flat-group ("ragged") kernels for any other D >= 16
            bits = _fold_bits(table, W, D, (W, D, M.tobytes(), off.tobytes(), str(dev))) if D >= 16 else None
            plan = (table, bits)
    if len(_FOLD_PLAN) >= 32:
        _FOLD_PLAN.pop(next(iter(_FOLD_PLAN)))
    _FOLD_PLAN[key] = plan
    return plan


def _process_device(vol, mask_wh, angle_interval):
    """process_voxel_grid on device tensors: vol (n0,n1,n2) u8, mask_wh (n0,n1) u8 -> carved (n0,n1,n2) u8.
    Index folds (multiples of 90 degrees, when the table says so) run one by one; every run of consecutive resample
    passes in between goes to the library as ONE call that ping-pongs between two buffers."""
    n0, n1, n2 = vol.shape
    dev = vol.device
    cur = vol
    pending = []                                            # (M, off) of the resample passes not yet issued

    def flush(cur):
        if not pending:
            return cur
        a = cur if cur is not vol else cur.clone()          # never write into the caller's volume
        b = torch.empty_like(a)
        Ms = np.ascontiguousarray(np.stack([m for m, _ in pending]), dtype=np.float64)
        offs = np.ascontiguousarray(np.stack([o for _, o in pending]), dtype=np.float64)
        in_b = ctypes.c_int(0)
        check(lib.p3d_resample_carve_passes(ptr(a), ptr(b), n0, n1, n2, _dptr(Ms), _dptr(offs), len(pending), ptr(mask_wh),
                                            ctypes.byref(in_b), stream_ptr()), "p3d_resample_carve_passes")
        _launched(len(pending))
        pending.clear()
        return b if in_b.value else a

    for angle in tqdm(range(0, 91, angle_interval), desc="90 Carving", leave=True, disable=None):
        M, off = _pass_transform((n0, n1, n2), angle)
        # an index fold is only possible at multiples of 90 degrees; other angles go straight to the resample kernel
        table, foldable = _fold_table(n0, n2, M, off, dev) if angle % 90 == 0 else (None, False)
        if foldable:
            cur = flush(cur)
            out = torch.empty_like(cur)
            check(lib.p3d_fold_gather(ptr(cur), n0, n1, n2, ptr(table), ptr(mask_wh), ptr(out), stream_ptr()),
                  "p3d_fold_gather")
            _launched()
            cur = out
        else:
            pending.append((M, off))
    return flush(cur)


class _PackedMask:
    """A (H,W,3) uint8 semantic mask.  Device side: the RGB image as a tensor (uploaded once) and per-colour (H,W) masks
    from the colour-compare kernel.  Host side (built lazily, only for the general-angle paths): the channels packed
    into one 32-bit key per pixel, so that every colour test is a single integer compare."""

    def __init__(self, semantic_mask):
        self._src = semantic_mask
        self._dev = {}
        self._sem = None
        self._packed = None
        self.shape = tuple(semantic_mask.shape)
        dt = semantic_mask.dtype
        self.is_rgb_u8 = len(self.shape) == 3 and self.shape[2] == 3 and dt in (torch.uint8, np.uint8, np.dtype(np.uint8))

    @property
    def sem(self):
        if self._sem is None:
            self._sem = self._src.cpu().numpy() if _is_tensor(self._src) else np.asarray(self._src)
        return self._sem

    @property
    def packed(self):
        if self._packed is None and self.is_rgb_u8:
            sem = self.sem
            self._packed = (sem[..., 0].astype(np.uint32) | (sem[..., 1].astype(np.uint32) << 8)
                            | (sem[..., 2].astype(np.uint32) << 16))
        return self._packed

    def device_rgb(self, dev):
        """The mask as a contiguous (H,W,3) uint8 device tensor (uploaded once), or None when it is not an RGB u8 image."""
        if not self.is_rgb_u8:
            return None
        t = self._dev.get(str(dev))
        if t is None:
            t = self._src if (_is_tensor(self._src) and self._src.device == dev and self._src.dtype == torch.uint8) \
                else torch.from_numpy(np.ascontiguousarray(self.sem)).to(dev)
            t = self._dev[str(dev)] = t.contiguous()
        return t

    def device_match(self, colour, dev):
        """(H,W) uint8 device mask of all(mask == colour, axis=-1) (cached per colour)."""
        key = (str(dev), tuple(int(v) for v in np.asarray(colour).reshape(3)))
        m = self._dev.get(key)
        if m is None:
            rgb = self.device_rgb(dev)
            if rgb is None:
                m = torch.from_numpy(self.match([colour]).astype(np.uint8)).to(dev)
            else:
                m = _colour_mask(rgb.view(1, *rgb.shape), colour).view(rgb.shape[0], rgb.shape[1])
            self._dev[key] = m
        return m

    def match(self, colours):
        """any over colours of all(mask == colour, axis=-1)  (voxel_carving_utils.py:143-146, :170)."""
        out = np.zeros(self.shape[:2], bool)
        for c in colours:
            if self.packed is None:
                out |= np.all(self.sem == np.asarray(c), axis=-1)
                continue
            r, g, b = (int(v) for v in np.asarray(c).reshape(3))
            if 0 <= r < 256 and 0 <= g < 256 and 0 <= b < 256:
                out |= self.packed == np.uint32(r | (g << 8) | (b << 16))
        return out


def _mask2d_bool(semantic_mask, colours):
    pm = semantic_mask if isinstance(semantic_mask, _PackedMask) else _PackedMask(semantic_mask)
    return pm.match(colours)


def _colour_args(colour):
    c = [int(v) for v in np.asarray(colour).reshape(3)]
    return c


# =========================================================
# Public API (reference names and signatures)
# =========================================================
@nv.on_device
def carve_voxel_grid_with_masks(voxel_grid, combined_mask):
    """voxel_carving_utils.py:76-97: np.where(mask, voxel_grid, 0) with the mask broadcast along depth."""
    as_tensor = _is_tensor(voxel_grid)
    dev = nv.require_cuda(voxel_grid.device if as_tensor and voxel_grid.is_cuda else None)
    is_color = voxel_grid.ndim == 4
    W, H, D = voxel_grid.shape[:3]
    mask = _mask_to_wh(combined_mask, W, H)
    if not ((mask.ndim == 2) or (mask.ndim == 3 and mask.shape[2] == 3)):
        raise ValueError("Unsupported mask shape")
    if mask.ndim == 3 and not is_color:
        raise ValueError("Unsupported mask shape")
    grid = _to_dev_u8(voxel_grid, dev, "voxel_grid")
    m = mask if _is_tensor(mask) else np.asarray(mask)
    m = _to_dev_u8((m != 0), dev, "mask")
    out = torch.empty_like(grid)
    check(lib.p3d_mask_carve(ptr(grid), W, H, D, 3 if is_color else 1, ptr(m), 3 if mask.ndim == 3 else 1, ptr(out),
                             stream_ptr()), "p3d_mask_carve")
    _launched()
    return _ret(out, as_tensor)


@nv.on_device
def process_voxel_grid(voxel_grid, combined_mask, angle_interval=90):
    """voxel_carving_utils.py:104-126: for angle in range(0, 91, angle_interval): rotate the (already rotated)
    grid by `angle` about shape/2 with trilinear interpolation, then carve with the mask."""
    as_tensor = _is_tensor(voxel_grid)
    dev = nv.require_cuda(voxel_grid.device if as_tensor and voxel_grid.is_cuda else None)
    if voxel_grid.ndim != 3:
        raise ValueError("process_voxel_grid expects a (W,H,D) occupancy grid")
    W, H, D = voxel_grid.shape
    mask = _mask_to_wh(combined_mask, W, H)
    if mask.ndim != 2:
        raise ValueError("Unsupported mask shape")
    vol = _to_dev_u8(voxel_grid, dev, "voxel_grid")
    m = mask if _is_tensor(mask) else np.asarray(mask)
    m = _to_dev_u8((m != 0), dev, "mask")
    return _ret(_process_device(vol, m, angle_interval), as_tensor)


@nv.on_device
def apply_colored_mask_to_voxel_grid(carved_voxel_grid, colored_mask):
    """voxel_carving_utils.py:128-136: out[x,y,z,:] = colored_mask[y,x,:] where carved == 1, else 0."""
    as_tensor = _is_tensor(carved_voxel_grid)
    dev = nv.require_cuda(carved_voxel_grid.device if as_tensor and carved_voxel_grid.is_cuda else None)
    W, H, D = carved_voxel_grid.shape
    carved = _to_dev_u8(carved_voxel_grid, dev, "carved_voxel_grid")
    col = _to_dev_u8(_first3(colored_mask), dev, "colored_mask")
    if tuple(col.shape) != (H, W, 3):
        raise ValueError(f"colored_mask {tuple(col.shape)} does not match (H,W,3)=({H},{W},3)")
    out = torch.empty((W, H, D, 3), dtype=torch.uint8, device=dev)
    check(lib.p3d_colourise(ptr(carved), W, H, D, ptr(col), ptr(out), stream_ptr()), "p3d_colourise")
    _launched()
    return _ret(out, as_tensor)


def _group_image(jobs, H, W):
    """(H,W) uint32: bit g where the pixel is in group g's mask m AND in _mask_to_wh(m) -- m itself for W != H, m.T for a
    square image (the reference's quirk), i.e. m & m.T there; bit g of (gm & gm.T) is exactly that, so the square case
    costs one transpose for all groups."""
    gm = np.zeros((H, W), np.uint32)
    for g, (m2d, _) in enumerate(jobs):
        gm |= m2d.view(np.uint8).astype(np.uint32) << np.uint32(g)
    if W == H:
        gm &= np.ascontiguousarray(gm.T)
    return gm


_GROUP_KEYS = {}          # (device, colours per group) -> (keys u32, group_of i32) device tensors
_NAME_COLOURS = {}        # part names per group -> their PART_COLORS as nested int tuples (PART_COLORS is a constant table)


def _job_colours(group_jobs):
    """PART_COLORS of every group's part names as nested tuples of ints (cached per name structure)."""
    key = tuple(tuple(names) for names, _ in group_jobs)
    hit = _NAME_COLOURS.get(key)
    if hit is None:
        if len(_NAME_COLOURS) >= 64:
            _NAME_COLOURS.clear()
        hit = _NAME_COLOURS[key] = _flat_colours([[PART_COLORS[n] for n in names] for names in key])
    return hit


def _flat_colours(group_colours):
    if isinstance(group_colours, tuple) and all(isinstance(g, tuple) and all(isinstance(c, tuple) for c in g)
                                                for g in group_colours):
        return group_colours                                  # already the nested-tuple form of _job_colours
    return tuple(tuple(tuple(int(v) for v in np.asarray(c).reshape(3)) for c in cols) for cols in group_colours)


def _group_image_device(semantic_mask, group_colours, dev):
    """(H,W) int32 device image of group bits (p3d_group_image): bit g where the pixel's colour belongs to group g, for a
    square image also at the transposed pixel (the reference's _mask_to_wh quirk).  Colours outside 0..255 match
    nothing, as in the reference's comparison with a uint8 image."""
    flat = _flat_colours(group_colours)
    hit = _GROUP_KEYS.get((str(dev), flat))
    if hit is None:
        keys, grp = [], []
        for g, cols in enumerate(flat):
            for r, gg, b in cols:
                if 0 <= r < 256 and 0 <= gg < 256 and 0 <= b < 256:
                    keys.append(r | (gg << 8) | (b << 16))
                    grp.append(g)
        if len(keys) > 128:
            raise ValueError("part_carve: more than 128 part colours")
        hit = (torch.tensor(keys or [0], dtype=torch.int32, device=dev), torch.tensor(grp or [0], dtype=torch.int32, device=dev),
               len(keys))
        if len(_GROUP_KEYS) >= 16:
            _GROUP_KEYS.pop(next(iter(_GROUP_KEYS)))
        _GROUP_KEYS[(str(dev), flat)] = hit
    rgb = semantic_mask.device_rgb(dev)
    H, W = int(rgb.shape[0]), int(rgb.shape[1])
    gm = torch.empty((H, W), dtype=torch.int32, device=dev)
    check(lib.p3d_group_image(ptr(rgb), H, W, ptr(hit[0]), ptr(hit[1]), hit[2], ptr(gm), stream_ptr()), "p3d_group_image")
    _launched()
    return gm


@nv.on_device
def part_carve(colored_grid, semantic_mask, group_jobs, visualize=False, *, x_range=None):
    """voxel_carving_utils.py:139-160: per part group, carve the group's voxels with the group's own mask under
    the group's symmetry angle and merge the survivors.

    x_range=(x0, x1) (addition, multi-GPU): return only the output slab [x0, x1) along axis 0, computed from the whole
    (replicated) input grid without any exchange -- `utils.sweep.carve_sharded(lambda a, b: part_carve(g, sem, jobs,
    x_range=(a, b)), W)`.  Slabs of the all-90-degree bit path cost slab-sized traffic; any other job list computes the
    full grid and slices it."""
    as_tensor = _is_tensor(colored_grid)
    dev = nv.require_cuda(colored_grid.device if as_tensor and colored_grid.is_cuda else None)
    grid = _to_dev_u8(colored_grid, dev, "colored_grid")
    W, H, D, _ = grid.shape
    if x_range is not None:
        x0, x1 = int(x_range[0]), int(x_range[1])
        if not (0 <= x0 <= x1 <= W):
            raise ValueError(f"x_range {x_range} outside [0, {W}]")
    semantic_mask = semantic_mask if isinstance(semantic_mask, _PackedMask) else _PackedMask(semantic_mask)
    group_jobs = list(group_jobs)
    out = None
    # All groups at 90 degrees on a cubic grid: every group in ONE fused pass, the per-pixel group bits built on the
    # device (an empty group simply contributes no bits; the reference skips it, :148-149).
    if (group_jobs and all(a == 90 for _, a in group_jobs) and len(group_jobs) <= 32 and D == W
            and semantic_mask.is_rgb_u8 and tuple(semantic_mask.shape[:2]) == (H, W)):
        plan = _fold_plan(W, H, D, dev)
        if plan is not None:
            table, bits = plan
            n_groups = len(group_jobs)
            gm_hw = _group_image_device(semantic_mask, _job_colours(group_jobs), dev)
            if x_range is not None and bits is not None and bits[2] is not None and x1 > x0 and D % 32 == 0:
                ws_bytes = int(lib.p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                slab = torch.empty((x1 - x0, H, D, 3), dtype=torch.uint8, device=dev)
                check(lib.p3d_part_carve_fold_bits_slab(ptr(grid), W, H, D, x0, x1 - x0, ptr(bits[0]), bits[1], bits[2],
                                                        ptr(gm_hw), n_groups, ptr(slab), ptr(ws), ws_bytes, stream_ptr()),
                      "p3d_part_carve_fold_bits_slab")
                _launched(4)
                return _ret(slab, as_tensor)
            out = torch.empty_like(grid)
            if bits is not None and bits[2] is not None:       # z-separable table: bit-packed occupancy / group masks
                ws_bytes = int(lib.p3d_part_carve_bits_workspace_bytes(W, H, D, n_groups))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                check(lib.p3d_part_carve_fold_bits(ptr(grid), W, H, D, ptr(bits[0]), bits[1], bits[2], ptr(gm_hw), n_groups,
                                                   ptr(out), ptr(ws), ws_bytes, stream_ptr()), "p3d_part_carve_fold_bits")
                _launched(3)
            else:
                check(lib.p3d_part_carve_fold(ptr(grid), W, H, D, ptr(table), ptr(gm_hw), ptr(out), stream_ptr()),
                      "p3d_part_carve_fold")
                _launched()
    jobs = []
    if out is None:
        for names, angle in group_jobs:
            m2d = _mask2d_bool(semantic_mask, [PART_COLORS[n] for n in names])          # (H,W)
            if not m2d.any():
                continue
            jobs.append((m2d, angle))                                                   # (H,W) bool; m = m2d.T is the reference's (W,H)
    if out is None:
        out = torch.zeros_like(grid)
        for m2d, angle in jobs:
            m = np.ascontiguousarray(m2d.T)                                          # (W,H) bool
            sel = torch.from_numpy(m.astype(np.uint8)).to(dev)                       # (W,H)
            occ = torch.empty((W, H, D), dtype=torch.uint8, device=dev)
            check(lib.p3d_crop_occupancy(ptr(grid), W, H, D, 0, 0, 0, W, H, D, ptr(sel), ptr(occ), stream_ptr()),
                  "p3d_crop_occupancy")
            mm = _to_dev_u8(np.ascontiguousarray(_mask_to_wh(m, W, H)), dev)
            carved = _process_device(occ, mm, angle)
            check(lib.p3d_accumulate_part(ptr(grid), ptr(carved), W, H, D, ptr(sel), ptr(out), stream_ptr()),
                  "p3d_accumulate_part")
            _launched(2)
    if x_range is not None:
        out = out[x0:x1].contiguous()
    return _ret(out, as_tensor)


class PartCarveSlab:
    """part_carve of ONE x slab of a grid whose rows are sharded over ranks (addition; the reference is single process,
    voxel_carving_utils.py:139-160).  begin(): pass A on this rank's rows (output from the voxel-local terms + the rows'
    occupancy bits); `occ` is the occupancy-bit array of the WHOLE grid, (W, H, D/32) int32, of which only rows
    [x0, x1) are filled -- the caller fills the other rows from the other ranks (utils.sweep.part_carve_sharded: one NCCL
    all-gather, 1/24 of the grid bytes); finish(): pass B clears the runs whose rotated source is empty and returns the
    (x1-x0, H, D, 3) output slab.  Only the all-90-degree bit path can be sharded this way (ValueError otherwise: carve
    a replicated grid with part_carve(..., x_range=...) instead)."""

    def __init__(self, grid_slab, semantic_mask, group_jobs, W, x_range, workspace=None):
        self.as_tensor = _is_tensor(grid_slab)
        dev = nv.require_cuda(grid_slab.device if self.as_tensor and grid_slab.is_cuda else None)
        self.grid = _to_dev_u8(grid_slab, dev, "grid_slab")
        self.x0, self.x1 = int(x_range[0]), int(x_range[1])
        n, H, D, _ = self.grid.shape
        self.W, self.H, self.D = int(W), int(H), int(D)
        if not (0 <= self.x0 <= self.x1 <= self.W) or n != self.x1 - self.x0:
            raise ValueError(f"slab of {n} rows does not match x_range {x_range} of width {W}")
        semantic_mask = semantic_mask if isinstance(semantic_mask, _PackedMask) else _PackedMask(semantic_mask)
        jobs = list(group_jobs)
        ok = (bool(jobs) and all(a == 90 for _, a in jobs) and len(jobs) <= 32 and D == W and D % 32 == 0
              and semantic_mask.is_rgb_u8 and tuple(semantic_mask.shape[:2]) == (H, W))
        bits = None
        if ok:
            plan = _fold_plan(W, H, D, dev)
            bits = plan[1] if plan is not None else None
            ok = bits is not None and bits[2] is not None
        if not ok:
            raise ValueError("sharded-input part_carve needs the all-90-degree bit path (cubic grid, D % 32 == 0, an RGB "
                             "uint8 mask of the grid's (H,W)); carve a replicated grid with part_carve(..., x_range=...)")
        self.gm_hw = _group_image_device(semantic_mask, _job_colours(jobs), dev)
        self.bits, self.n_groups = bits, len(jobs)
        self.ws_bytes = int(lib.p3d_part_carve_bits_workspace_bytes(W, H, D, len(jobs)))
        if workspace is not None:                             # caller-owned scratch, e.g. symmetric (peer-mapped) memory
            if workspace.numel() < self.ws_bytes or workspace.dtype != torch.uint8 or workspace.device != dev:
                raise ValueError(f"workspace must be a uint8 tensor of >= {self.ws_bytes} bytes on {dev}")
            self.ws = workspace[:self.ws_bytes]
        else:
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        words = D // 32
        self.occ = self.ws[:W * H * words * 4].view(torch.int32).view(W, H, words)
        self.out = torch.empty_like(self.grid)
        # the per-group mask bits depend on the mask and the job list only: packed once per object, so that begin() /
        # finish() can be repeated (e.g. on new contents of the same input tensor) with two launches
        check(lib.p3d_part_carve_pack_groups(ptr(self.gm_hw), self.W, self.H, self.D, self.n_groups, ptr(self.ws),
                                             self.ws_bytes, stream_ptr()), "p3d_part_carve_pack_groups")
        _launched()

    def begin(self):
        if self.x1 > self.x0:
            check(lib.p3d_part_carve_slab_pass_a_packed(ptr(self.grid), self.W, self.H, self.D, self.x0, self.x1 - self.x0,
                                                        ptr(self.bits[0]), self.bits[1], ptr(self.gm_hw), self.n_groups,
                                                        ptr(self.out), ptr(self.ws), self.ws_bytes, stream_ptr()),
                  "p3d_part_carve_slab_pass_a_packed")
            _launched()
        return self

    @staticmethod
    def workspace_bytes(W, H, D, n_groups):
        return int(lib.p3d_part_carve_bits_workspace_bytes(int(W), int(H), int(D), int(n_groups)))

    def needed_words(self, x0, x1):
        """Word range [lo, hi) of every occupancy row that pass B of the output slab [x0, x1) reads: the fold takes
        occ[c - z, y, x + c2], i.e. the z-bit range [x0 + c2, x1 + c2) of EVERY source row -- 1/world of the bits."""
        c2, words = self.bits[2], self.D // 32
        lo = max(0, (x0 + c2) // 32 - 1)
        hi = min(words, (x1 + c2 + 31) // 32 + 1)
        return lo, max(hi, lo)

    def finish(self, peers=None, n_ranks=1):
        """peers: int64 device tensor of `n_ranks` workspace pointers (one per rank, peer-mapped): pass B then reads the
        other ranks' occupancy rows in place instead of from this rank's copy."""
        if self.x1 > self.x0:
            if peers is not None:
                check(lib.p3d_part_carve_slab_pass_b_peers(self.W, self.H, self.D, self.x0, self.x1 - self.x0, self.bits[1],
                                                           self.bits[2], ptr(self.out), ptr(self.ws), self.ws_bytes, ptr(peers),
                                                           int(n_ranks), stream_ptr()), "p3d_part_carve_slab_pass_b_peers")
            else:
                check(lib.p3d_part_carve_slab_pass_b(self.W, self.H, self.D, self.x0, self.x1 - self.x0, self.bits[1],
                                                     self.bits[2], ptr(self.out), ptr(self.ws), self.ws_bytes, stream_ptr()),
                      "p3d_part_carve_slab_pass_b")
            _launched()
        return _ret(self.out, self.as_tensor)


_STATS_CAP = 1024         # components whose statistics are gathered before the one host read-back


def _label_components(mask_u8, extra=None, conn=6):
    """scipy.ndimage.label (6-connectivity) on device: (labels int32, n, bbox (n,6) ndarray, sums (n,4) ndarray).
    Labelling and the per-component statistics of up to _STATS_CAP components are enqueued back to back and read with
    ONE synchronisation (more components: a second statistics pass).  `extra`: a small device tensor whose host copy is
    wanted from the same synchronisation (returned as a fifth value)."""
    n0, n1, n2 = mask_u8.shape
    dev = mask_u8.device
    nvox = mask_u8.numel()
    labels = torch.empty((n0, n1, n2), dtype=torch.int32, device=dev)
    ncomp = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.p3d_label6_workspace_bytes(nvox))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    fn = lib.p3d_label26 if conn == 26 else lib.p3d_label6
    check(fn(ptr(mask_u8), n0, n1, n2, ptr(labels), ptr(ncomp), ptr(ws), ws_bytes, stream_ptr()), "p3d_label")
    _launched(7)

    def stats(cap):
        bbox = torch.empty((cap, 6), dtype=torch.int32, device=dev)
        sums = torch.empty((cap, 4), dtype=torch.int64, device=dev)
        check(lib.p3d_component_stats(ptr(labels), n0, n1, n2, cap, ptr(bbox), ptr(sums), stream_ptr()), "p3d_component_stats")
        _launched(2)
        return bbox, sums

    bbox, sums = stats(_STATS_CAP)
    n = int(ncomp.cpu().item())                               # the one synchronisation
    if n > _STATS_CAP:
        bbox, sums = stats(n)
    res = (labels, n, bbox[:n].cpu().numpy() if n else np.zeros((0, 6), np.int32),
           sums[:n].cpu().numpy() if n else np.zeros((0, 4), np.int64))
    return res + (extra.cpu().numpy(),) if extra is not None else res


def _colour_mask(grid, colour):
    n = grid.numel() // 3
    mask = torch.empty(grid.shape[:3], dtype=torch.uint8, device=grid.device)
    r, g, b = _colour_args(colour)
    check(lib.p3d_colour_mask(ptr(grid), n, r, g, b, ptr(mask), stream_ptr()), "p3d_colour_mask")
    _launched()
    return mask


def _lr_tables(boxes, angle):
    """Host tables of one left_right_guided_carve call: per pass the inverse rotation handed to scipy at :116-123, per
    (component, pass) the offset for that crop's shape -- the very NumPy expressions the per-component path evaluates."""
    angles = list(range(0, 91, angle))
    Ms = np.stack([_pass_transform((1, 1, 1), a)[0] for a in angles])
    offs = np.stack([np.stack([_pass_transform((w, h, d), a)[1] for a in angles]) for (_, _, _, w, h, d) in boxes])
    return np.ascontiguousarray(Ms, dtype=np.float64), np.ascontiguousarray(offs, dtype=np.float64)


def _boxes_overlap(boxes):
    """True when two bounding boxes intersect (then the paste order of the reference's loop matters)."""
    n = len(boxes)
    if n < 2:
        return False
    b = np.asarray(boxes, dtype=np.int64)
    lo, hi = b[:, 0:3], b[:, 0:3] + b[:, 3:6]
    for i in range(n - 1):
        if np.any(np.all((lo[i + 1:] < hi[i]) & (lo[i] < hi[i + 1:]), axis=1)):
            return True
    return False


@nv.on_device
def left_right_guided_carve(colored_grid, semantic_mask, target_color, angle=60, visualize=False, stride=2):
    """voxel_carving_utils.py:163-210: for every 6-connected 3-D component of `target_color`, carve the
    component's bounding-box crop with the matching crop of the part's 2-D mask under `angle` symmetry.

    Device schedule: colour masks -> 3-D labelling + component boxes -> ONE host read (boxes; sizes the scratch) ->
    one crop launch, one launch per rotate-and-carve pass and one paste launch for ALL components (blockIdx.y = the
    component) -> one host read of the per-component counts for the log."""
    as_tensor = _is_tensor(colored_grid)
    dev = nv.require_cuda(colored_grid.device if as_tensor and colored_grid.is_cuda else None)
    grid = _to_dev_u8(colored_grid, dev, "colored_grid")
    W, H, D, _ = grid.shape
    carved = grid.clone()
    pm = semantic_mask if isinstance(semantic_mask, _PackedMask) else _PackedMask(semantic_mask)
    mask2d = pm.device_match(target_color, dev)                                  # (H,W) u8 on the device
    if tuple(mask2d.shape) != (H, W):
        raise ValueError(f"semantic_mask {tuple(pm.shape)} does not match the grid's (H,W)=({H},{W})")
    any2d = mask2d.max().reshape(1)
    labels, n, bbox, _, any_host = _label_components(_colour_mask(grid, target_color), extra=any2d)
    if not int(any_host[0]):
        print(f"[SKIP] No mask for color {target_color}")
        return _ret(carved, as_tensor)
    print(f"[{target_color}] 3D components: {n}")
    if n:
        boxes = [(int(b[0]), int(b[1]), int(b[2]), int(b[3]) + 1 - int(b[0]), int(b[4]) + 1 - int(b[1]), int(b[5]) + 1 - int(b[2]))
                 for b in bbox]
        vols = np.array([w * h * d for (_, _, _, w, h, d) in boxes], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(vols)[:-1]])
        comps = np.zeros((n, 8), np.int32)
        comps[:, 0:6] = np.asarray(boxes, dtype=np.int32)
        comps[:, 6] = (offsets & 0xffffffff).astype(np.uint32).view(np.int32)
        comps[:, 7] = (offsets >> 32).astype(np.int32)
        Ms, offs = _lr_tables(boxes, angle)
        total = int(vols.sum())
        comps_d = torch.from_numpy(comps).to(dev)
        tab = torch.from_numpy(np.concatenate([Ms.reshape(-1), offs.reshape(-1)])).to(dev)
        Ms_d, offs_d = tab[:Ms.size], tab[Ms.size:]
        buf = torch.empty((2, max(total, 1)), dtype=torch.uint8, device=dev)
        counts = torch.empty(n, dtype=torch.int64, device=dev)
        check(lib.p3d_lr_carve_components(ptr(grid), ptr(labels), W, H, D, ptr(mask2d), ptr(comps_d), n, int(vols.max()),
                                          ptr(Ms_d), ptr(offs_d), Ms.shape[0], ptr(buf[0]), ptr(buf[1]),
                                          1 if _boxes_overlap(boxes) else 0, ptr(carved), ptr(counts), stream_ptr()),
              "p3d_lr_carve_components")
        _launched(2 + Ms.shape[0])
        for i, ((x0, y0, z0, w, h, d), c) in enumerate(zip(boxes, counts.cpu().tolist()), start=1):
            print(f"  - Component {i}: bbox ({x0},{y0},{z0}) → ({x0 + w},{y0 + h},{z0 + d})")
            print(f"    carved voxels: {int(c)}")
    if visualize:
        warnings.warn("visualize=True is ignored: plotting is outside this package's scope")
    return _ret(carved, as_tensor)


def _extrude_inplace(out, mask_2d, axis, direction, depth, fill_color):
    W, H, D, _ = out.shape
    if axis not in (0, 2):
        return
    if _is_tensor(mask_2d) and mask_2d.is_cuda and mask_2d.dtype == torch.uint8:
        md = mask_2d.contiguous()                             # device mask (any non-zero byte selects the column)
    else:
        m = mask_2d.cpu().numpy() if _is_tensor(mask_2d) else np.asarray(mask_2d)
        md = torch.from_numpy(np.ascontiguousarray(m != 0).astype(np.uint8)).to(out.device)
    want = (H, W) if axis == 2 else (H, D)
    if tuple(md.shape) != want:
        raise ValueError(f"operands could not be broadcast together: mask {tuple(md.shape)} vs {want}")
    colour = (0, 0, 0) if fill_color is None else _colour_args(fill_color)
    sign = 1 if direction == "+" else -1
    check(lib.p3d_extrude(ptr(out), W, H, D, ptr(md), int(md.shape[0]), int(md.shape[1]), axis, sign, int(depth), colour[0],
                          colour[1], colour[2], stream_ptr()), "p3d_extrude")
    _launched()


@nv.on_device
def extrude_from_surface(grid, mask_2d, axis, direction="+", depth=5, fill_color=None):
    """voxel_carving_utils.py:213-248: from the first occupied voxel of each masked column (index 0 / last when
    the column is empty), paint `depth` voxels along `axis` in `direction`."""
    as_tensor = _is_tensor(grid)
    dev = nv.require_cuda(grid.device if as_tensor and grid.is_cuda else None)
    out = _to_dev_u8(grid, dev, "grid").clone()
    _extrude_inplace(out, mask_2d, axis, direction, depth, fill_color)
    return _ret(out, as_tensor)


@nv.on_device
def recolor_backward_components(voxel_grid, color, new_color, k=4, sort_axis=2):
    """voxel_carving_utils.py:252-266: keep the k components of `color` with the smallest mean coordinate on
    `sort_axis` (stable: ties keep the lower scipy id), recolour the others to `new_color`."""
    as_tensor = _is_tensor(voxel_grid)
    dev = nv.require_cuda(voxel_grid.device if as_tensor and voxel_grid.is_cuda else None)
    if as_tensor:
        grid = voxel_grid.to(dev).contiguous().clone()
    else:
        grid = torch.from_numpy(np.ascontiguousarray(voxel_grid)).to(dev)
    if grid.dtype != torch.uint8:
        raise TypeError("voxel_grid must be uint8")
    labels, n, _, sums = _label_components(_colour_mask(grid, color))
    if n:
        means = [(i, np.float64(sums[i - 1, 1 + sort_axis]) / np.float64(sums[i - 1, 0])) for i in range(1, n + 1)]
        keep = {i for i, _ in sorted(means, key=lambda t: t[1])[:k]}
        flags = np.array([0 if i in keep else 1 for i in range(1, n + 1)], np.uint8)
        if flags.any():
            r, g, b = _colour_args(new_color)
            fl = torch.from_numpy(flags).to(dev)
            check(lib.p3d_recolour_components(ptr(labels), ptr(fl), labels.numel(), r, g, b, ptr(grid), stream_ptr()),
                  "p3d_recolour_components")
            _launched()
    return _ret(grid, as_tensor)


@nv.on_device
def global_carve(binary_mask, semantic_mask_exterior, angle_interval=90, stride=4, visualize=False,
                 device=None, return_tensor=False, x_range=None):
    """voxel_carving_utils.py:269-298: start from a full (w,h,w) grid, carve it with the binary front mask under
    4-way symmetry (rotate-and-carve every `angle_interval` degrees), then colour every surviving voxel with
    the semantic colour of its (x,y) pixel.  Returns the (W,H,D=W,3) uint8 grid.

    Extension (keyword-only in spirit): `x_range=(x0, x1)` returns only the x-slab `[x0:x1]` of that grid -- the unit
    of multi-GPU sharding (utils.sweep.carve_sharded); supported on the 90-degree fast path."""
    as_tensor = return_tensor or _is_tensor(binary_mask)
    dev = nv.require_cuda(device if device is not None else (binary_mask.device if _is_tensor(binary_mask) and binary_mask.is_cuda else None))
    bm_dev = binary_mask if (_is_tensor(binary_mask) and binary_mask.is_cuda) else None     # device mask: never copied to the host
    bm = None if bm_dev is not None else (binary_mask.numpy() if _is_tensor(binary_mask) else np.asarray(binary_mask))
    if (bm_dev if bm_dev is not None else bm).ndim != 2:
        raise ValueError("binary_mask must be 2-D (H,W)")
    h, w = (int(v) for v in (bm_dev if bm_dev is not None else bm).shape)
    W, H, D = w, h, w
    col = _to_dev_u8(_first3(semantic_mask_exterior), dev, "semantic_mask_exterior")
    if tuple(col.shape) != (H, W, 3):
        raise ValueError(f"semantic_mask_exterior {tuple(col.shape)} does not match (H,W,3)=({H},{W},3)")
    # (W,H) mask of the reference (:279-283).  W, H come from binary_mask itself, so _mask_to_wh always takes its
    # (H,W) -> .T branch here and the (H,W) form the fold kernels want is simply `bm != 0`.
    m_hw_np = None if bm is None else bm != 0

    def mask_hw_device():
        if bm_dev is not None:
            return bm_dev.contiguous() if bm_dev.dtype == torch.uint8 else (bm_dev != 0).to(torch.uint8)
        return torch.from_numpy(m_hw_np.view(np.uint8) if m_hw_np.flags.c_contiguous else
                                np.ascontiguousarray(m_hw_np).view(np.uint8)).to(dev)
    x0, x1 = (0, W) if x_range is None else (int(x_range[0]), int(x_range[1]))
    if not (0 <= x0 <= x1 <= W):
        raise ValueError(f"x_range {x_range} outside [0, {W}]")
    out = None
    plan = _fold_plan(W, H, D, dev) if angle_interval == 90 else None
    if plan is not None:
        table, bits = plan
        m_hw = mask_hw_device()                           # any non-zero byte counts as foreground
        if bits is not None:                              # z-separable table: bit-packed mask, 10x fewer loads
            out = torch.empty((x1 - x0, H, D, 3), dtype=torch.uint8, device=dev)
            wpr = (W + 31) // 32 + 2
            mbits = torch.empty((H, wpr), dtype=torch.int32, device=dev)
            st = stream_ptr()
            check(lib.p3d_pack_mask_bits(ptr(m_hw), H, W, ptr(mbits), wpr, st), "p3d_pack_mask_bits")
            check(lib.p3d_global_carve_fold_bits(W, H, D, x0, x1 - x0, ptr(bits[0]), bits[1], ptr(mbits), wpr,
                                                 ptr(col), 1, ptr(out), st), "p3d_global_carve_fold_bits")
            _launched(2)
            x0, x1 = 0, out.shape[0]                      # the slab has been applied
        else:
            out = torch.empty((W, H, D, 3), dtype=torch.uint8, device=dev)
            check(lib.p3d_global_carve_fold(W, H, D, ptr(table), ptr(m_hw), ptr(col), 1, ptr(out),
                                            stream_ptr()), "p3d_global_carve_fold")
            _launched()
    if out is None:
        vol = torch.ones((W, H, D), dtype=torch.uint8, device=dev)
        m_wh_d = (_mask_to_wh(mask_hw_device(), W, H) != 0).to(torch.uint8).contiguous()   # (W,H)
        carved = _process_device(vol, m_wh_d, angle_interval)
        out = torch.empty((W, H, D, 3), dtype=torch.uint8, device=dev)
        check(lib.p3d_colourise(ptr(carved), W, H, D, ptr(col), ptr(out), stream_ptr()), "p3d_colourise")
        _launched()
    if (x0, x1) != (0, out.shape[0]):
        out = out[x0:x1].contiguous()                     # general path: slab cut from the full grid
    if visualize:
        warnings.warn("visualize=True is ignored: plotting is outside this package's scope")
    return _ret(out, as_tensor)


@nv.on_device
def partwise_carve(colored_voxel_grid, semantic_mask_exterior, semantic_mask_full, part_colors_np, group_jobs,
                   part_symmetry, extrusion_depths, recolor_back_minarets=True, visualize=False, stride=4):
    """voxel_carving_utils.py:302-400: part_carve -> left_right_guided_carve per symmetric part -> interior
    extrusion (4 directions per part) -> re-orientation to (D, H-flipped, W) and back-minaret recolouring.
    The whole chain stays on the device; only the result is copied back for NumPy callers."""
    as_tensor = _is_tensor(colored_voxel_grid)
    dev = nv.require_cuda(colored_voxel_grid.device if as_tensor and colored_voxel_grid.is_cuda else None)
    grid = _to_dev_u8(colored_voxel_grid, dev, "colored_voxel_grid")

    semantic_mask_exterior = _PackedMask(semantic_mask_exterior)
    semantic_mask_full = _PackedMask(semantic_mask_full)
    grid = part_carve(grid, semantic_mask_exterior, group_jobs, visualize=False)     # fresh tensor from here on
    for part, angle in part_symmetry.items():
        grid = left_right_guided_carve(colored_grid=grid, semantic_mask=semantic_mask_exterior,
                                       target_color=part_colors_np[part], angle=angle, visualize=False, stride=stride)
    for part, depth in extrusion_depths.items():
        mask = semantic_mask_full.device_match(part_colors_np[part], dev)    # one device mask for the four directions
        for axis, direction in ((2, "+"), (2, "-"), (0, "+"), (0, "-")):     # extrude_4dirs :356-361
            _extrude_inplace(grid, mask, axis, direction, depth, part_colors_np[part])
    if recolor_back_minarets:
        W, H, D, _ = grid.shape
        oriented = torch.empty((D, H, W, 3), dtype=torch.uint8, device=dev)
        check(lib.p3d_reorient(ptr(grid), W, H, D, ptr(oriented), stream_ptr()), "p3d_reorient")
        _launched()
        grid = recolor_backward_components(oriented, part_colors_np["front_minarets"],
                                           new_color=part_colors_np["back_minarets"], k=2, sort_axis=0)
    if visualize:
        warnings.warn("visualize=True is ignored: plotting is outside this package's scope")
    return _ret(grid, as_tensor)
