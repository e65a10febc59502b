"""Drop-in for the reference's utils/projection_utils.py (project_colored_voxels)."""
from __future__ import annotations

import numpy as np
import torch

from . import _engine as eng
from . import _native as nv
from .camera_geometry import candidate_row, working_dtype


def _points_f32(pts3d, device) -> torch.Tensor:
    """The kernels take float32 points, which is what get_voxel_points_by_parts produces
    (voxel_utils.py:18).  Other dtypes are accepted when every value is float32-representable."""
    if isinstance(pts3d, torch.Tensor):
        t = pts3d
        if t.dtype != torch.float32:
            t32 = t.to(torch.float32)
            if not torch.equal(t32.to(t.dtype), t):
                raise NotImplementedError("points must be exactly representable in float32")
            t = t32
        return t.to(device).reshape(-1, 3).contiguous()
    a = np.asarray(pts3d)
    if a.dtype != np.float32:
        a32 = a.astype(np.float32)
        if not np.array_equal(a32.astype(a.dtype), a):
            raise NotImplementedError("points must be exactly representable in float32")
        a = a32
    return torch.from_numpy(np.ascontiguousarray(a.reshape(-1, 3))).to(device)


@nv.on_device
def project_colored_voxels(pts3d, colors, cam_pos, target, f, cx, cy, H, W, device=None, return_tensor=False):
    """projection_utils.py:5-23.  Projects the coloured points through one look-at camera and
    returns the (H, W, 3) uint8 image in which, per pixel, the LAST point in array order wins
    (NumPy fancy-assignment semantics), untouched pixels are (0,0,0).

    Arithmetic dtype follows NumPy promotion of (pts3d, cam_pos, target): float32 when all are
    float32, else float64.  Accepts NumPy arrays or torch tensors; returns a NumPy array unless
    `return_tensor=True`.
    """
    dev = nv.require_cuda(device)
    H, W = int(H), int(W)
    pts = _points_f32(pts3d, dev)
    dt = working_dtype(pts3d, np.asarray(cam_pos), np.asarray(target))
    cand = torch.from_numpy(candidate_row(cam_pos, target, f, cx, cy, dt)[None]).to(dev)
    cols = nv.to_device(colors, torch.uint8, dev).reshape(-1, 3)
    if cols.shape[0] != pts.shape[0]:
        raise ValueError(f"shape mismatch: {pts.shape[0]} points but {cols.shape[0]} colours")
    cams = eng.setup_cameras(cand)
    zbuf = eng.splat(pts, None, cams, H, W, nv.MODE_JOINT)
    img = eng.resolve_rgb(zbuf[0], cols)
    return img if return_tensor else img.cpu().numpy()


def visualize_reprojection(image, voxel_keypoints_dict, image_keypoints_dict, cam_params, title="Reprojection"):
    """projection_utils.py:26-67 without the matplotlib figure: the printed key-point comparison table."""
    from .camera_geometry import project
    projected = {name: tuple(project(pt3d, cam_params["cam_pos"], cam_params["target"], cam_params["f"], cam_params["cx"],
                                     cam_params["cy"])) for name, pt3d in voxel_keypoints_dict.items()}
    print(f"\n{'Keypoint':<6} | {'GT (x, y)':<30} | {'Projected (x, y)':<30} | Error (L2)")
    print("-" * 80)
    total_err = 0
    for name in image_keypoints_dict:
        gt = np.array(image_keypoints_dict[name])
        pr = np.array(projected[name])
        err = np.linalg.norm(gt - pr)
        total_err += err
        print(f"{name:<6} | {tuple(np.round(gt, 2))!s:<30} | {tuple(np.round(pr, 2))!s:<30} | {err:.2f}")
    avg_err = total_err / len(image_keypoints_dict)
    print(f"\nAverage Reprojection Error: {avg_err:.2f} pixels")
